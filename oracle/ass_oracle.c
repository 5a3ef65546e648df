/* oracle/ass_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  See ass_oracle.h for status and rules.
 *
 * Build with -ffp-contract=off: the only fused multiply-adds are the explicit fma() calls below, which
 * reproduce what gcc emits for the reference under its own flags:
 *   - src_spring/*.cpp are GNU C++ (contraction on): Vector::len()/dot() in matrix_math.cpp:158-166 become
 *     fma(z,z, fma(x,x, y*y)); cross() becomes fma(a,b, -(c*d));
 *   - src/*.c are -std=c99 (contraction off): gravity.c, integrator_leapfrog.c, tools.c are evaluated
 *     left-to-right, unfused.
 * Each function cites the reference lines it restates.
 */
#define _GNU_SOURCE
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "ass_oracle.h"

#define P(i, f) p[(size_t)(i) * ORC_STRIDE + (f)]
static const double L_EPS = 1e-6; /* springs.cpp:28 */

/* dot(lhs, rhs) as compiled: matrix_math.cpp:163-166 under GNU C++ -O3 with FMA hardware */
static inline double dot3(double lx, double ly, double lz, double rx, double ry, double rz) {
	return fma(lz, rz, fma(lx, rx, ly * ry));
}
/* Vector::len(), matrix_math.cpp:158-160 */
static inline double len3(double x, double y, double z) {
	return sqrt(dot3(x, y, z, x, y, z));
}

/* ------------------------------------------------------------------------------------------------ */
/* per-step path                                                                                    */
/* ------------------------------------------------------------------------------------------------ */

void orc_zero_accel(double* p, int n) { /* physics.cpp:557-564 */
	for (int i = 0; i < n; i++) {
		P(i, ORC_AX) = 0.0;
		P(i, ORC_AY) = 0.0;
		P(i, ORC_AZ) = 0.0;
	}
}

void orc_spring_forces(double* p, const orc_spring* s, int ns) { /* springs.cpp:288-346 */
	for (int q = 0; q < ns; q++) {
		const double len0 = s[q].rs0, k = s[q].k;
		const int i = s[q].particle_1, j = s[q].particle_2;
		/* spring_r: x_i - x_j  (springs.cpp:220-232) */
		const double dx = P(i, ORC_X) - P(j, ORC_X);
		const double dy = P(i, ORC_Y) - P(j, ORC_Y);
		const double dz = P(i, ORC_Z) - P(j, ORC_Z);
		const double len = len3(dx, dy, dz) + L_EPS;          /* :302 */
		const double hx = dx / len, hy = dy / len, hz = dz / len; /* :303 */
		const double m_i = P(i, ORC_M), m_j = P(j, ORC_M);
		const double c = -k * (len - len0);                   /* :312, (-k*(len-len0)) * len_hat */
		const double fx = hx * c, fy = hy * c, fz = hz * c;
		P(i, ORC_AX) += fx / m_i;                             /* :315-320 */
		P(j, ORC_AX) -= fx / m_j;
		P(i, ORC_AY) += fy / m_i;
		P(j, ORC_AY) -= fy / m_j;
		P(i, ORC_AZ) += fz / m_i;
		P(j, ORC_AZ) -= fz / m_j;
		const double gamma = s[q].gamma;
		if (gamma > 0.0) {                                    /* :324 */
			const double dvx = P(i, ORC_VX) - P(j, ORC_VX);
			const double dvy = P(i, ORC_VY) - P(j, ORC_VY);
			const double dvz = P(i, ORC_VZ) - P(j, ORC_VZ);
			const double strain_rate = dot3(dvx, dvy, dvz, hx, hy, hz) / len; /* :331 */
			const double m_red = m_i * m_j / (m_i + m_j);     /* :334 */
			const double d = gamma * m_red * strain_rate;     /* :337 */
			const double gx = hx * d, gy = hy * d, gz = hz * d;
			P(i, ORC_AX) -= gx / m_i;                         /* :338-343 */
			P(j, ORC_AX) += gx / m_j;
			P(i, ORC_AY) -= gy / m_i;
			P(j, ORC_AY) += gy / m_j;
			P(i, ORC_AZ) -= gz / m_i;
			P(j, ORC_AZ) += gz / m_j;
		}
	}
}

void orc_gravity_basic_slab(double* p, int n, int n_active, double G, double softening, int i_lo, int i_hi) {
	/* gravity.c:72-103 with nghost = 0, gb.shift = 0, gravity_ignore_terms = 0, N_var = 0 */
	const double softening2 = softening * softening;
	const int na = (n_active == -1) ? n : n_active;
	for (int i = i_lo; i < i_hi; i++) {
		P(i, ORC_AX) = 0;
		P(i, ORC_AY) = 0;
		P(i, ORC_AZ) = 0;
	}
	for (int i = i_lo; i < i_hi; i++) {
		for (int j = 0; j < na; j++) {
			if (i == j) continue;
			const double dx = (0.0 + P(i, ORC_X)) - P(j, ORC_X);
			const double dy = (0.0 + P(i, ORC_Y)) - P(j, ORC_Y);
			const double dz = (0.0 + P(i, ORC_Z)) - P(j, ORC_Z);
			const double r = sqrt(dx * dx + dy * dy + dz * dz + softening2);
			const double prefact = -G / (r * r * r) * P(j, ORC_M);
			P(i, ORC_AX) += prefact * dx;
			P(i, ORC_AY) += prefact * dy;
			P(i, ORC_AZ) += prefact * dz;
		}
	}
}

void orc_gravity_basic(double* p, int n, int n_active, double G, double softening) {
	orc_gravity_basic_slab(p, n, n_active, G, softening, 0, n);
}

void orc_leapfrog_part1(double* p, int n, orc_sim* sim) { /* integrator_leapfrog.c:39-51 */
	const double dt = sim->dt;
	for (int i = 0; i < n; i++) {
		P(i, ORC_X) += 0.5 * dt * P(i, ORC_VX);
		P(i, ORC_Y) += 0.5 * dt * P(i, ORC_VY);
		P(i, ORC_Z) += 0.5 * dt * P(i, ORC_VZ);
	}
	sim->t += dt / 2.;
}

void orc_leapfrog_part2(double* p, int n, orc_sim* sim) { /* integrator_leapfrog.c:52-67 */
	const double dt = sim->dt;
	for (int i = 0; i < n; i++) {
		P(i, ORC_VX) += dt * P(i, ORC_AX);
		P(i, ORC_VY) += dt * P(i, ORC_AY);
		P(i, ORC_VZ) += dt * P(i, ORC_AZ);
		P(i, ORC_X) += 0.5 * dt * P(i, ORC_VX);
		P(i, ORC_Y) += 0.5 * dt * P(i, ORC_VY);
		P(i, ORC_Z) += 0.5 * dt * P(i, ORC_VZ);
	}
	sim->t += dt / 2.;
	sim->dt_last_done = sim->dt;
}

void orc_step(double* p, int n, const orc_spring* s, int ns, orc_sim* sim) { /* rebound.c:69-144 */
	orc_leapfrog_part1(p, n, sim);                                   /* :72 */
	if (sim->gravity_basic)                                          /* :109, gravity.c:68-70 NONE = no-op */
		orc_gravity_basic(p, n, sim->n_active, sim->G, sim->softening);
	if (sim->zero_first) orc_zero_accel(p, n);                       /* :114 additional_forces */
	orc_spring_forces(p, s, ns);
	orc_leapfrog_part2(p, n, sim);                                   /* :119 */
}

void orc_heartbeat(double* p, int n, orc_sim* sim) {
	if (sim->center_heartbeat) orc_center_sim(p, n, 0, n - sim->n_perts);
}

/* reb_check_exit, rebound.c:509-569 (no MPI, no pause) */
static int check_exit(orc_sim* r, int n, double tmax, double* last_full_dt) {
	const double dtsign = copysign(1., r->dt);
	if (r->status >= 0) {
	} else if (tmax != INFINITY) {
		if (r->exact_finish_time == 1) {
			if ((r->t + r->dt) * dtsign >= tmax * dtsign) {
				double tscale = 1e-12 * fabs(tmax);
				if (tscale < 1e-200) tscale = 1e-12;
				if (r->t == tmax) {
					r->status = 0;
				} else if (r->status == -2) {
					if (fabs(r->t - tmax) < tscale) {
						r->status = 0;
					} else {
						r->dt = tmax - r->t;
					}
				} else {
					r->status = -2;
					if (r->dt_last_done != 0.) *last_full_dt = r->dt_last_done;
					r->dt = tmax - r->t;
				}
			} else {
				if (r->status == -2) r->status = -1;
			}
		} else {
			if (r->t * dtsign >= tmax * dtsign) r->status = 0;
		}
	}
	if (n <= 0) r->status = 2;
	return r->status;
}

int orc_integrate(double* p, int n, const orc_spring* s, int ns, orc_sim* sim, double tmax) { /* rebound.c:621-704 */
	double last_full_dt = sim->dt;
	sim->dt_last_done = 0.;
	sim->status = -1;
	orc_heartbeat(p, n, sim);
	while (check_exit(sim, n, tmax, &last_full_dt) < 0) {
		orc_step(p, n, s, ns, sim);
		orc_heartbeat(p, n, sim);
	}
	if (sim->exact_finish_time == 1) sim->dt = last_full_dt;
	return sim->status;
}

/* ------------------------------------------------------------------------------------------------ */
/* spring build                                                                                     */
/* ------------------------------------------------------------------------------------------------ */

void orc_free(void* ptr) { free(ptr); }

static void push_spring(orc_spring** springs, int* ns, int* cap, orc_spring spr) { /* springs.cpp:96-106 */
	if (*ns >= *cap) {
		*cap = *cap ? *cap * 2 : 1024;
		*springs = (orc_spring*) realloc(*springs, sizeof(orc_spring) * (size_t) *cap);
	}
	(*springs)[(*ns)++] = spr;
}

int orc_add_spring(const double* p, int i, int j, double k, double gamma, orc_spring** springs, int* ns, int* cap) {
	/* springs.cpp:52-92 */
	if (i == j) return -1; /* reference throws */
	const int lo = j < i ? j : i, hi = j < i ? i : j;
	for (int q = 0; q < *ns; q++)
		if ((*springs)[q].particle_1 == lo && (*springs)[q].particle_2 == hi) return q;
	orc_spring spr;
	memset(&spr, 0, sizeof(spr));
	spr.k = k;
	spr.gamma = gamma;
	spr.particle_1 = lo;
	spr.particle_2 = hi;
	spr.rs0 = len3(P(lo, ORC_X) - P(hi, ORC_X), P(lo, ORC_Y) - P(hi, ORC_Y), P(lo, ORC_Z) - P(hi, ORC_Z)); /* :84 */
	push_spring(springs, ns, cap, spr);
	return *ns - 1;
}

int orc_connect_springs_dist_n2(const double* p, double max_dist, int i_low, int i_high, double k, double gamma,
		orc_spring** springs, int* ns, int* cap) { /* springs.cpp:110-143 */
	if (i_high <= i_low) return *ns;
	for (int i = i_low; i < i_high - 1; i++) {
		for (int j = i + 1; j < i_high; j++) {
			const double dist = len3(P(i, ORC_X) - P(j, ORC_X), P(i, ORC_Y) - P(j, ORC_Y), P(i, ORC_Z) - P(j, ORC_Z));
			if (dist < max_dist) orc_add_spring(p, i, j, k, gamma, springs, ns, cap);
		}
	}
	return *ns;
}

/* -- uniform grid over particles [lo, hi) -- */
typedef struct {
	double x0, y0, z0, inv_h;
	int nx, ny, nz;
	int* head; /* per cell: first particle, -1 = empty */
	int* next; /* per particle (index - lo) */
	int lo;
} grid_t;

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
/* a non-finite coordinate (a blown-up run) goes to a border cell: such a particle is never within any cutoff of another */
static inline int cell_axis(double x, double x0, double inv_h, int n) {
	const double q = floor((x - x0) * inv_h);
	if (!(q >= 0.0)) return 0; /* also NaN */
	if (q > (double) (n - 1)) return n - 1;
	return (int) q;
}
static inline int cell_of(const grid_t* g, double x, double y, double z, int* cx, int* cy, int* cz) {
	*cx = cell_axis(x, g->x0, g->inv_h, g->nx);
	*cy = cell_axis(y, g->y0, g->inv_h, g->ny);
	*cz = cell_axis(z, g->z0, g->inv_h, g->nz);
	return (*cz * g->ny + *cy) * g->nx + *cx;
}

/* Cell edge h is inflated by 1e-9 so that |x_i - x_j| < cutoff always lands within +-1 cell despite rounding
 * in the index computation. */
static int grid_alloc(grid_t* g, double xmin, double ymin, double zmin, double xmax, double ymax, double zmax,
		double h, int lo, int count, double max_cells) {
	h *= (1.0 + 1e-9);
	/* bound the cell count (sparse clouds): coarsening the grid never changes which pairs pass the cutoff */
	if (!(xmax >= xmin)) xmin = xmax = 0.0; /* no finite particle at all: one cell */
	if (!(ymax >= ymin)) ymin = ymax = 0.0;
	if (!(zmax >= zmin)) zmin = zmax = 0.0;
	for (int tries = 0; tries < 4000; tries++) {
		double cells = (floor((xmax - xmin) / h) + 1) * (floor((ymax - ymin) / h) + 1) * (floor((zmax - zmin) / h) + 1);
		if (cells <= max_cells) break;
		h *= 1.5;
	}
	g->x0 = xmin; g->y0 = ymin; g->z0 = zmin;
	g->inv_h = 1.0 / h;
	g->nx = (int) floor((xmax - xmin) / h) + 1;
	g->ny = (int) floor((ymax - ymin) / h) + 1;
	g->nz = (int) floor((zmax - zmin) / h) + 1;
	g->lo = lo;
	size_t nc = (size_t) g->nx * g->ny * g->nz;
	g->head = (int*) malloc(sizeof(int) * nc);
	g->next = (int*) malloc(sizeof(int) * (size_t) (count > 0 ? count : 1));
	if (!g->head || !g->next) return -1;
	for (size_t c = 0; c < nc; c++) g->head[c] = -1;
	return 0;
}
static void grid_free(grid_t* g) { free(g->head); free(g->next); }
static inline void grid_insert(grid_t* g, const double* p, int idx) {
	int cx, cy, cz;
	int c = cell_of(g, P(idx, ORC_X), P(idx, ORC_Y), P(idx, ORC_Z), &cx, &cy, &cz);
	g->next[idx - g->lo] = g->head[c];
	g->head[c] = idx;
}
static void bbox(const double* p, int lo, int hi, double* mn, double* mx) {
	for (int d = 0; d < 3; d++) { mn[d] = DBL_MAX; mx[d] = -DBL_MAX; }
	for (int i = lo; i < hi; i++) {
		if (!(isfinite(P(i, 0)) && isfinite(P(i, 1)) && isfinite(P(i, 2)))) continue; /* pairs with nobody: stays out of the box */
		for (int d = 0; d < 3; d++) {
			if (P(i, d) < mn[d]) mn[d] = P(i, d);
			if (P(i, d) > mx[d]) mx[d] = P(i, d);
		}
	}
}

static int cmp_int(const void* a, const void* b) { return (*(const int*) a > *(const int*) b) - (*(const int*) a < *(const int*) b); }

/* open-addressing set of (i,j) pairs, for the duplicate rule of add_spring across repeated calls */
typedef struct { uint64_t* slot; size_t mask; } pairset_t;
static inline uint64_t mix64(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x; }
static void pairset_init(pairset_t* s, size_t n) {
	size_t cap = 16;
	while (cap < 2 * n + 1) cap <<= 1;
	s->slot = (uint64_t*) malloc(sizeof(uint64_t) * cap);
	memset(s->slot, 0xff, sizeof(uint64_t) * cap);
	s->mask = cap - 1;
}
static void pairset_add(pairset_t* s, int i, int j) {
	uint64_t key = ((uint64_t) (uint32_t) i << 32) | (uint32_t) j;
	size_t h = mix64(key) & s->mask;
	while (s->slot[h] != UINT64_MAX && s->slot[h] != key) h = (h + 1) & s->mask;
	s->slot[h] = key;
}
static int pairset_has(const pairset_t* s, int i, int j) {
	uint64_t key = ((uint64_t) (uint32_t) i << 32) | (uint32_t) j;
	size_t h = mix64(key) & s->mask;
	while (s->slot[h] != UINT64_MAX) {
		if (s->slot[h] == key) return 1;
		h = (h + 1) & s->mask;
	}
	return 0;
}

int orc_connect_springs_dist(const double* p, double max_dist, int i_low, int i_high, double k, double gamma,
		orc_spring** springs, int* ns, int* cap) {
	/* Same result as orc_connect_springs_dist_n2: springs appended in lexicographic (i, j) order, strict `<`
	 * on the same rounded length, rs0 = that length, pairs already present skipped. */
	if (i_high <= i_low) return *ns;
	if (!(max_dist > 0.0)) return *ns; /* len >= 0 is never < a non-positive cutoff */
	double mn[3], mx[3];
	bbox(p, i_low, i_high, mn, mx);
	grid_t g;
	if (grid_alloc(&g, mn[0], mn[1], mn[2], mx[0], mx[1], mx[2], max_dist, i_low, i_high - i_low, 8.0 * (i_high - i_low) + 4096.0)) return -1;
	for (int i = i_high - 1; i >= i_low; i--) grid_insert(&g, p, i);
	pairset_t pre;
	const int had = *ns;
	if (had) {
		pairset_init(&pre, (size_t) had);
		for (int q = 0; q < had; q++) pairset_add(&pre, (*springs)[q].particle_1, (*springs)[q].particle_2);
	}
	int* cand = NULL;
	int ccap = 0;
	for (int i = i_low; i < i_high - 1; i++) {
		int cx, cy, cz, nc = 0;
		cell_of(&g, P(i, ORC_X), P(i, ORC_Y), P(i, ORC_Z), &cx, &cy, &cz);
		for (int z = cz - 1; z <= cz + 1; z++) {
			if (z < 0 || z >= g.nz) continue;
			for (int y = cy - 1; y <= cy + 1; y++) {
				if (y < 0 || y >= g.ny) continue;
				for (int x = cx - 1; x <= cx + 1; x++) {
					if (x < 0 || x >= g.nx) continue;
					for (int j = g.head[(z * g.ny + y) * g.nx + x]; j >= 0; j = g.next[j - g.lo]) {
						if (j <= i) continue;
						const double dist = len3(P(i, ORC_X) - P(j, ORC_X), P(i, ORC_Y) - P(j, ORC_Y), P(i, ORC_Z) - P(j, ORC_Z));
						if (dist < max_dist) {
							if (nc >= ccap) { ccap = ccap ? 2 * ccap : 256; cand = (int*) realloc(cand, sizeof(int) * (size_t) ccap); }
							cand[nc++] = j;
						}
					}
				}
			}
		}
		qsort(cand, (size_t) nc, sizeof(int), cmp_int);
		for (int q = 0; q < nc; q++) {
			const int j = cand[q];
			if (had && pairset_has(&pre, i, j)) continue;
			orc_spring spr;
			memset(&spr, 0, sizeof(spr));
			spr.k = k; spr.gamma = gamma; spr.particle_1 = i; spr.particle_2 = j;
			spr.rs0 = len3(P(i, ORC_X) - P(j, ORC_X), P(i, ORC_Y) - P(j, ORC_Y), P(i, ORC_Z) - P(j, ORC_Z));
			push_spring(springs, ns, cap, spr);
		}
	}
	free(cand);
	if (had) free(pre.slot);
	grid_free(&g);
	return *ns;
}

/* ------------------------------------------------------------------------------------------------ */
/* node generators                                                                                  */
/* ------------------------------------------------------------------------------------------------ */

double orc_random_uniform(double lo, double hi) { /* tools.c:44-46, ISO C: unfused */
	return ((double) rand()) / ((double) (RAND_MAX)) * (hi - lo) + lo;
}

static double* grow_rows(double* p, int* cap, int need) {
	if (need <= *cap) return p;
	while (*cap < need) *cap = *cap ? *cap * 2 : 1024;
	return (double*) realloc(p, sizeof(double) * ORC_STRIDE * (size_t) *cap);
}

static void finish_masses(double* p, int i_0, int n, double total_mass) {
	const double particle_mass = total_mass / (n - i_0);
	for (int i = i_0; i < n; i++) P(i, ORC_M) = particle_mass;
}

int orc_rand_ellipsoid_n2(double** pp, int n0, double min_dist, double a, double b, double c, double total_mass) {
	/* shapes.cpp:210-278, literal */
	double* p = *pp;
	int cap = n0, n = n0;
	const int i_0 = n0;
	const int n_part = 40 * pow(2.0 * a / min_dist, 3.0); /* :214, double -> int truncation */
	const double radius = (min_dist / 2.0) / 2.0;       /* :227-228 */
	for (int t = 0; t < n_part; t++) {
		const double x = orc_random_uniform(-a, a);
		const double y = orc_random_uniform(-b, b);
		const double z = orc_random_uniform(-c, c);
		if (len3(x / a, y / b, z / c) < 1.0) {
			int too_near = 0;
			for (int j = i_0; !too_near && j < n; j++)
				if (len3(x - P(j, ORC_X), y - P(j, ORC_Y), z - P(j, ORC_Z)) < min_dist) too_near = 1;
			if (!too_near) {
				p = grow_rows(p, &cap, n + 1);
				memset(&P(n, 0), 0, sizeof(double) * ORC_STRIDE);
				P(n, ORC_X) = x; P(n, ORC_Y) = y; P(n, ORC_Z) = z;
				P(n, ORC_M) = 1.0; P(n, ORC_R) = radius;
				n++;
			}
		}
	}
	finish_masses(p, i_0, n, total_mass);
	*pp = p;
	return n;
}

int orc_rand_ellipsoid(double** pp, int n0, double min_dist, double a, double b, double c, double total_mass) {
	/* Same candidate stream, same accept tests as above; the "is any accepted node within min_dist" scan is
	 * answered from a uniform grid instead of a linear scan (the answer is an OR, so order does not matter). */
	double* p = *pp;
	int cap = n0, n = n0;
	const int i_0 = n0;
	const int n_part = 40 * pow(2.0 * a / min_dist, 3.0);
	const double radius = (min_dist / 2.0) / 2.0;
	grid_t g;
	int gcap = 1024;
	if (grid_alloc(&g, -a, -b, -c, a, b, c, min_dist, i_0, gcap, 2.0e8)) return -1;
	for (int t = 0; t < n_part; t++) {
		const double x = orc_random_uniform(-a, a);
		const double y = orc_random_uniform(-b, b);
		const double z = orc_random_uniform(-c, c);
		if (len3(x / a, y / b, z / c) < 1.0) {
			int cx, cy, cz, too_near = 0;
			const int cc = cell_of(&g, x, y, z, &cx, &cy, &cz);
			for (int zz = cz - 1; zz <= cz + 1 && !too_near; zz++) {
				if (zz < 0 || zz >= g.nz) continue;
				for (int yy = cy - 1; yy <= cy + 1 && !too_near; yy++) {
					if (yy < 0 || yy >= g.ny) continue;
					for (int xx = cx - 1; xx <= cx + 1 && !too_near; xx++) {
						if (xx < 0 || xx >= g.nx) continue;
						for (int j = g.head[(zz * g.ny + yy) * g.nx + xx]; j >= 0; j = g.next[j - i_0])
							if (len3(x - P(j, ORC_X), y - P(j, ORC_Y), z - P(j, ORC_Z)) < min_dist) { too_near = 1; break; }
					}
				}
			}
			if (!too_near) {
				p = grow_rows(p, &cap, n + 1);
				memset(&P(n, 0), 0, sizeof(double) * ORC_STRIDE);
				P(n, ORC_X) = x; P(n, ORC_Y) = y; P(n, ORC_Z) = z;
				P(n, ORC_M) = 1.0; P(n, ORC_R) = radius;
				if (n - i_0 >= gcap) { gcap *= 2; g.next = (int*) realloc(g.next, sizeof(int) * (size_t) gcap); }
				g.next[n - i_0] = g.head[cc];
				g.head[cc] = n;
				n++;
			}
		}
	}
	grid_free(&g);
	finish_masses(p, i_0, n, total_mass);
	*pp = p;
	return n;
}

static int lattice_common(double** pp, int n0, double a, double b, double c, double total_mass, int n) {
	double* p = *pp;
	const int i_0 = n0;
	finish_masses(p, i_0, n, total_mass);
	const double min_d = orc_mindist(p, i_0, n);
	for (int i = i_0; i < n; i++) P(i, ORC_R) = min_d / 4.0;
	(void) a; (void) b; (void) c;
	return n;
}

int orc_hcp_ellipsoid(double** pp, int n0, double min_dist, double a, double b, double c, double total_mass) {
	/* shapes.cpp:282-352.  shapes.cpp is GNU C++: `dx*i + xfac*(j%2) + xfac*(k%2)` is contracted by gcc to
	 * fma(dx, i, t_j) + t_k with the loop invariants t_j = xfac*(j%2), t_k = xfac*(k%2) hoisted (found by matching
	 * the reference's output bit for bit, tests/test_oracle.py::test_golden_lattices);
	 * `yfac*(j + 0.5*(k%2))` to yfac * fma(0.5, k%2, j). */
	double* p = *pp;
	int cap = n0, n = n0;
	const double dx = 1.08 * min_dist;
	const double xfac = dx / 2.0, yfac = dx * sin(M_PI / 3.0), zfac = dx * sqrt(2.0 / 3.0);
	const int nx = (int) (1.2 * a / dx), ny = (int) (1.2 * b / yfac), nz = (int) (1.2 * c / zfac);
	for (int k = -nz; k <= nz; k++) {
		const double z = zfac * k;
		for (int j = -ny; j <= ny; j++) {
			const double y = yfac * fma(0.5, (double) (k % 2), (double) j);
			for (int i = -nx; i <= nx; i++) {
				const double x = fma(dx, (double) i, xfac * (double) (j % 2)) + xfac * (double) (k % 2);
				if (len3(x / a, y / b, z / c) <= 1.0) {
					p = grow_rows(p, &cap, n + 1);
					memset(&P(n, 0), 0, sizeof(double) * ORC_STRIDE);
					P(n, ORC_X) = x; P(n, ORC_Y) = y; P(n, ORC_Z) = z;
					P(n, ORC_M) = 1.0;
					n++;
				}
			}
		}
	}
	*pp = p;
	return lattice_common(pp, n0, a, b, c, total_mass, n);
}

int orc_cubic_ellipsoid(double** pp, int n0, double min_dist, double a, double b, double c, double total_mass) {
	/* shapes.cpp:356-431 */
	double* p = *pp;
	int cap = n0, n = n0;
	const int nx = (int) (1.2 * a / min_dist), ny = (int) (1.2 * b / min_dist), nz = (int) (1.2 * c / min_dist);
	const double fac = min_dist;
	for (int k = -nz; k <= nz; k++) {
		const double z = fac * k;
		for (int j = -ny; j <= ny; j++) {
			const double y = fac * j;
			for (int i = -nx; i <= nx; i++) {
				const double x = fac * i;
				if (len3(x / a, y / b, z / c) <= 1.0) {
					p = grow_rows(p, &cap, n + 1);
					memset(&P(n, 0), 0, sizeof(double) * ORC_STRIDE);
					P(n, ORC_X) = x; P(n, ORC_Y) = y; P(n, ORC_Z) = z;
					P(n, ORC_M) = 1.0;
					n++;
				}
			}
		}
	}
	*pp = p;
	return lattice_common(pp, n0, a, b, c, total_mass, n);
}

double orc_mindist_n2(const double* p, int i_low, int i_high) { /* shapes.cpp:814-833 */
	double min_dist = DBL_MAX;
	for (int i = i_low; i < i_high - 1; i++)
		for (int j = i + 1; j < i_high; j++) {
			const double d = len3(P(i, ORC_X) - P(j, ORC_X), P(i, ORC_Y) - P(j, ORC_Y), P(i, ORC_Z) - P(j, ORC_Z));
			if (d < min_dist) min_dist = d;
		}
	return min_dist;
}

double orc_mindist(const double* p, int i_low, int i_high) {
	/* min is exact and order-free, so any search that visits the closest pair returns the same double.
	 * Visit pairs in adjacent grid cells; if the best found is <= the cell edge it is the global minimum
	 * (a closer pair would share or neighbour a cell), otherwise coarsen and retry.
	 * NB the reference evaluates len(x_i - x_j) with i < j; len is even in its argument bit-for-bit. */
	const int count = i_high - i_low;
	if (count < 2) return DBL_MAX;
	if (count <= 2048) return orc_mindist_n2(p, i_low, i_high);
	double mn[3], mx[3];
	bbox(p, i_low, i_high, mn, mx);
	double vol = 1.0;
	for (int d = 0; d < 3; d++) vol *= (mx[d] - mn[d] > 0 ? mx[d] - mn[d] : 1e-300);
	double h = cbrt(vol / count);
	if (!(h > 0)) return orc_mindist_n2(p, i_low, i_high);
	for (int attempt = 0; attempt < 40; attempt++, h *= 2.0) {
		grid_t g;
		if (grid_alloc(&g, mn[0], mn[1], mn[2], mx[0], mx[1], mx[2], h, i_low, count, 8.0 * count + 4096.0)) return orc_mindist_n2(p, i_low, i_high);
		for (int i = i_low; i < i_high; i++) grid_insert(&g, p, i);
		const double h_eff = 1.0 / g.inv_h;
		double best = DBL_MAX;
		for (int i = i_low; i < i_high; i++) {
			int cx, cy, cz;
			cell_of(&g, P(i, ORC_X), P(i, ORC_Y), P(i, ORC_Z), &cx, &cy, &cz);
			for (int z = cz - 1; z <= cz + 1; z++) {
				if (z < 0 || z >= g.nz) continue;
				for (int y = cy - 1; y <= cy + 1; y++) {
					if (y < 0 || y >= g.ny) continue;
					for (int x = cx - 1; x <= cx + 1; x++) {
						if (x < 0 || x >= g.nx) continue;
						for (int j = g.head[(z * g.ny + y) * g.nx + x]; j >= 0; j = g.next[j - g.lo]) {
							if (j <= i) continue;
							const double d = len3(P(i, ORC_X) - P(j, ORC_X), P(i, ORC_Y) - P(j, ORC_Y), P(i, ORC_Z) - P(j, ORC_Z));
							if (d < best) best = d;
						}
					}
				}
			}
		}
		grid_free(&g);
		if (best < h_eff * (1.0 - 1e-6)) return best;
	}
	return orc_mindist_n2(p, i_low, i_high);
}

/* ------------------------------------------------------------------------------------------------ */
/* heartbeat helpers / diagnostics                                                                  */
/* ------------------------------------------------------------------------------------------------ */

double orc_sum_mass(const double* p, int i_low, int i_high) { /* orb.cpp:438-449 */
	double m_tot = 0.0;
	for (int i = i_low; i < i_high; i++) m_tot += P(i, ORC_M);
	return m_tot;
}

void orc_compute_com(const double* p, int i_low, int i_high, double* out3) { /* physics.cpp:41-57 */
	const double m_tot = orc_sum_mass(p, i_low, i_high);
	double sx = 0, sy = 0, sz = 0;
	for (int i = i_low; i < i_high; i++) {
		sx += P(i, ORC_X) * P(i, ORC_M);
		sy += P(i, ORC_Y) * P(i, ORC_M);
		sz += P(i, ORC_Z) * P(i, ORC_M);
	}
	out3[0] = sx / m_tot; out3[1] = sy / m_tot; out3[2] = sz / m_tot;
}

void orc_compute_cov(const double* p, int i_low, int i_high, double* out3) { /* physics.cpp:60-74 */
	const double m_tot = orc_sum_mass(p, i_low, i_high);
	double sx = 0, sy = 0, sz = 0;
	for (int i = i_low; i < i_high; i++) {
		sx += P(i, ORC_VX) * P(i, ORC_M);
		sy += P(i, ORC_VY) * P(i, ORC_M);
		sz += P(i, ORC_VZ) * P(i, ORC_M);
	}
	out3[0] = sx / m_tot; out3[1] = sy / m_tot; out3[2] = sz / m_tot;
}

void orc_center_sim(double* p, int n, int i_low, int i_high) { /* physics.cpp:98-109 */
	double c[3];
	orc_compute_com(p, i_low, i_high, c);
	for (int i = 0; i < n; i++) {
		P(i, ORC_X) -= c[0];
		P(i, ORC_Y) -= c[1];
		P(i, ORC_Z) -= c[2];
	}
}

void orc_move_resolved(double* p, const double* dx, const double* dv, int i_low, int i_high) { /* physics.cpp:387-402 */
	for (int i = i_low; i < i_high; i++) {
		P(i, ORC_X) += dx[0]; P(i, ORC_Y) += dx[1]; P(i, ORC_Z) += dx[2];
		P(i, ORC_VX) += dv[0]; P(i, ORC_VY) += dv[1]; P(i, ORC_VZ) += dv[2];
	}
}

void orc_subtract_com(double* p, int i_low, int i_high) { /* physics.cpp:77-84 */
	double c[3], z[3] = { 0, 0, 0 };
	orc_compute_com(p, i_low, i_high, c);
	c[0] = -c[0]; c[1] = -c[1]; c[2] = -c[2];
	orc_move_resolved(p, c, z, i_low, i_high);
}

void orc_subtract_cov(double* p, int i_low, int i_high) { /* physics.cpp:88-94 */
	double c[3], z[3] = { 0, 0, 0 };
	orc_compute_cov(p, i_low, i_high, c);
	c[0] = -c[0]; c[1] = -c[1]; c[2] = -c[2];
	orc_move_resolved(p, z, c, i_low, i_high);
}

/* cross(), matrix_math.cpp:169-174 as compiled: a*b - c*d -> fma(a, b, -(c*d)) */
static inline void cross3(const double* l, const double* r, double* o) {
	o[0] = fma(l[1], r[2], -(l[2] * r[1]));
	o[1] = fma(l[2], r[0], -(l[0] * r[2]));
	o[2] = fma(l[0], r[1], -(l[1] * r[0]));
}

void orc_spin_body(double* p, int i_low, int i_high, double ox, double oy, double oz) { /* physics.cpp:184-210 */
	double com[3], cov[3], om[3] = { ox, oy, oz };
	orc_compute_com(p, i_low, i_high, com);
	orc_compute_cov(p, i_low, i_high, cov);
	if (len3(ox, oy, oz) > 1e-5) {
		for (int i = i_low; i < i_high; i++) {
			double dx[3] = { P(i, ORC_X) - com[0], P(i, ORC_Y) - com[1], P(i, ORC_Z) - com[2] }, w[3];
			cross3(om, dx, w);
			P(i, ORC_VX) = cov[0] + w[0];
			P(i, ORC_VY) = cov[1] + w[1];
			P(i, ORC_VZ) = cov[2] + w[2];
		}
	}
}

void orc_set_gamma(orc_spring* s, int ns, double g) { /* springs.cpp:146-154 */
	for (int i = 0; i < ns; i++) s[i].gamma = g;
}

void orc_diagnostics(const double* p, int i_low, int i_high, const orc_spring* s, int ns, double G, int want_gpe, double* out) {
	double com[3], cov[3];
	orc_compute_com(p, i_low, i_high, com);
	orc_compute_cov(p, i_low, i_high, cov);
	memcpy(out, com, sizeof(com));
	memcpy(out + 3, cov, sizeof(cov));
	/* measure_L, physics.cpp:136-162 */
	double L[3] = { 0, 0, 0 };
	for (int i = i_low; i < i_high; i++) {
		double dx[3] = { P(i, ORC_X) - com[0], P(i, ORC_Y) - com[1], P(i, ORC_Z) - com[2] };
		double dv[3] = { P(i, ORC_VX) - cov[0], P(i, ORC_VY) - cov[1], P(i, ORC_VZ) - cov[2] }, dL[3];
		cross3(dx, dv, dL);
		L[0] += dL[0] * P(i, ORC_M);
		L[1] += dL[1] * P(i, ORC_M);
		L[2] += dL[2] * P(i, ORC_M);
	}
	memcpy(out + 6, L, sizeof(L));
	/* compute_rot_kin, physics.cpp:508-528 */
	double KE = 0.0;
	for (int i = i_low; i < i_high; i++) {
		const double l = len3(P(i, ORC_VX) - cov[0], P(i, ORC_VY) - cov[1], P(i, ORC_VZ) - cov[2]);
		KE += 0.5 * P(i, ORC_M) * (l * l);
	}
	out[9] = KE;
	/* spring_potential_energy, physics.cpp:457-476 */
	double SPE = 0.0;
	for (int q = 0; q < ns; q++) {
		const int i = s[q].particle_1, j = s[q].particle_2;
		const double len = len3(P(i, ORC_X) - P(j, ORC_X), P(i, ORC_Y) - P(j, ORC_Y), P(i, ORC_Z) - P(j, ORC_Z)) + L_EPS;
		const double e = len - s[q].rs0;
		SPE += s[q].k * (e * e) / 2.0;
	}
	out[10] = SPE;
	/* grav_potential_energy, physics.cpp:479-504: unsoftened, i<j */
	double GPE = 0.0;
	if (want_gpe)
		for (int i = i_low; i < i_high; i++)
			for (int j = i + 1; j < i_high; j++)
				GPE -= G * P(i, ORC_M) * P(j, ORC_M) / len3(P(i, ORC_X) - P(j, ORC_X), P(i, ORC_Y) - P(j, ORC_Y), P(i, ORC_Z) - P(j, ORC_Z));
	out[11] = GPE;
	/* dEdt_total, physics.cpp:409-454 */
	double power = 0.0;
	for (int q = 0; q < ns; q++) {
		if (s[q].gamma == 0.0) continue;
		const int i = s[q].particle_1, j = s[q].particle_2;
		const double dx = P(i, ORC_X) - P(j, ORC_X), dy = P(i, ORC_Y) - P(j, ORC_Y), dz = P(i, ORC_Z) - P(j, ORC_Z);
		const double len = len3(dx, dy, dz) + L_EPS;
		const double hx = dx / len, hy = dy / len, hz = dz / len;
		const double m_i = P(i, ORC_M), m_j = P(j, ORC_M);
		const double sr = dot3(P(i, ORC_VX) - P(j, ORC_VX), P(i, ORC_VY) - P(j, ORC_VY), P(i, ORC_VZ) - P(j, ORC_VZ), hx, hy, hz) / len;
		const double m_red = m_i * m_j / (m_i + m_j);
		power += s[q].gamma * m_red * (sr * sr);
	}
	out[12] = power;
	/* mom_inertia, physics.cpp:116-134 */
	double I[6] = { 0, 0, 0, 0, 0, 0 };
	for (int i = i_low; i < i_high; i++) {
		const double dx = P(i, ORC_X) - com[0], dy = P(i, ORC_Y) - com[1], dz = P(i, ORC_Z) - com[2];
		const double l = len3(dx, dy, dz), m = P(i, ORC_M);
		const double r2m = m * (l * l);
		I[0] += r2m; I[1] += r2m; I[2] += r2m;
		I[0] -= m * (dx * dx); I[1] -= m * (dy * dy); I[2] -= m * (dz * dz);
		I[3] -= m * (dx * dy); I[4] -= m * (dx * dz); I[5] -= m * (dy * dz);
	}
	memcpy(out + 13, I, sizeof(I));
	out[19] = orc_sum_mass(p, i_low, i_high);
}

/* ------------------------------------------------------------------------------------------------ */
/* the other per-step terms of the phobos* modules (src_spring/orb.cpp)                             */
/* ------------------------------------------------------------------------------------------------ */

/* quadrupole_accel, orb.cpp:311-366.  Operator order of
 *     Vector a = constant * r / r5 * (5.0 * cos2 - diff) / particles[i].m;
 * is ((((constant * r) / r5) * (5 cos2 - diff)) / m), all component-wise (matrix_math.cpp:418-424, 453-460);
 * `scalar - Vector` is -(Vector - scalar) (matrix_math.cpp:453-455).  The reaction on i_p is accumulated in particle
 * order.  The racy `#pragma omp parallel for` of the reference is off. */
void orc_quadrupole_accel(double* p, int n, double G, double J2, double R, double phi, double theta, int i_p) {
	if (i_p >= n) return;       /* :315-316 */
	if (J2 == 0.0) return;      /* :319-320 */
	const double Gmc = G * P(i_p, ORC_M);
	const double constant = 1.5 * Gmc * J2 * pow(R, 2.0);
	const double xhat[3] = { sin(theta) * cos(phi), sin(theta) * sin(phi), cos(theta) };
	const double diff[3] = { 1.0, 1.0, 3.0 };
	for (int i = 0; i < n; i++) {
		if (i == i_p) continue;
		const double r[3] = { P(i, ORC_X) - P(i_p, ORC_X), P(i, ORC_Y) - P(i_p, ORC_Y), P(i, ORC_Z) - P(i_p, ORC_Z) };
		const double rl = len3(r[0], r[1], r[2]);
		const double r5 = pow(rl, 5.0) + 0.5;
		const double cos2 = pow(dot3(r[0], r[1], r[2], xhat[0], xhat[1], xhat[2]) / rl, 2.0);
		double a[3];
		for (int c = 0; c < 3; c++) a[c] = (constant * r[c]) / r5 * (-(diff[c] - 5.0 * cos2)) / P(i, ORC_M);
		P(i, ORC_AX) += a[0]; P(i, ORC_AY) += a[1]; P(i, ORC_AZ) += a[2];
		const double mfac = P(i, ORC_M) / P(i_p, ORC_M);
		P(i_p, ORC_AX) -= a[0] * mfac; P(i_p, ORC_AY) -= a[1] * mfac; P(i_p, ORC_AZ) -= a[2] * mfac;
	}
}

/* the common tail of drift_bin / drift_resolved (orb.cpp:218-238, 271-291): dv from the relative state (x, v) */
static void drift_dv(const double* x, const double* v, double GM, double timestep, double inv_tau_a, double inv_tau_e, double* dv) {
	const double rad = len3(x[0], x[1], x[2]);
	const double v_cent = sqrt(GM / rad);
	double xv[3], rl[3];
	cross3(x, v, xv);
	cross3(x, xv, rl);
	const double n = len3(rl[0], rl[1], rl[2]);
	for (int c = 0; c < 3; c++) {
		const double dir = rl[c] / n;
		const double delta_v = v[c] - v_cent * dir;
		dv[c] = timestep * (v[c] * inv_tau_a / 2.0 + delta_v * inv_tau_e);
	}
}

void orc_drift_resolved(double* p, double G, double timestep, double inv_tau_a, double inv_tau_e, int i_part, int i_low, int i_high) {
	const double m1 = P(i_part, ORC_M), m2 = orc_sum_mass(p, i_low, i_high);
	double com[3], cov[3], x[3], v[3], dv[3];
	orc_compute_com(p, i_low, i_high, com);
	orc_compute_cov(p, i_low, i_high, cov);
	for (int c = 0; c < 3; c++) {
		x[c] = com[c] - P(i_part, ORC_X + c);
		v[c] = cov[c] - P(i_part, ORC_VX + c);
	}
	drift_dv(x, v, G * (m1 + m2), timestep, inv_tau_a, inv_tau_e, dv);
	double dv2[3], zero[3] = { 0, 0, 0 };
	for (int c = 0; c < 3; c++) {
		P(i_part, ORC_VX + c) -= m1 * dv[c] / (m1 + m2);   /* dv_1, orb.cpp:293-298 */
		dv2[c] = -(m2 * dv[c] / (m1 + m2));                /* move_resolved(zero, -dv_2), orb.cpp:301 */
	}
	orc_move_resolved(p, zero, dv2, i_low, i_high);
}

void orc_drift_bin(double* p, double G, double timestep, double inv_tau_a, double inv_tau_e, int part_1, int part_2) {
	const double m1 = P(part_1, ORC_M), m2 = P(part_2, ORC_M);
	double x[3], v[3], dv[3];
	for (int c = 0; c < 3; c++) {
		x[c] = P(part_2, ORC_X + c) - P(part_1, ORC_X + c);
		v[c] = P(part_2, ORC_VX + c) - P(part_1, ORC_VX + c);
	}
	drift_dv(x, v, G * (m1 + m2), timestep, inv_tau_a, inv_tau_e, dv);
	for (int c = 0; c < 3; c++) {
		P(part_1, ORC_VX + c) -= m1 * dv[c] / (m1 + m2);
		P(part_2, ORC_VX + c) -= m2 * dv[c] / (m1 + m2);
	}
}

/* ------------------------------------------------------------------------------------------------ */
/* generator variants: literal restatements of shapes.cpp (one libc rand() stream, linear "too near" scan) */
/* ------------------------------------------------------------------------------------------------ */
typedef int (*inside_fn)(const void* ctx, const double* p, double x, double y, double z);

static int rand_fill_n2(double** pp, int n0, int n_part, const double* lo, const double* hi, inside_fn inside, const void* ctx,
		double min_dist, double radius, double total_mass) {
	double* p = *pp;
	int cap = n0, n = n0;
	const int i_0 = n0;
	for (int t = 0; t < n_part; t++) {
		const double x = orc_random_uniform(lo[0], hi[0]);
		const double y = orc_random_uniform(lo[1], hi[1]);
		const double z = orc_random_uniform(lo[2], hi[2]);
		if (inside && !inside(ctx, p, x, y, z)) continue;
		int too_near = 0;
		for (int j = i_0; !too_near && j < n; j++)
			if (len3(x - P(j, ORC_X), y - P(j, ORC_Y), z - P(j, ORC_Z)) < min_dist) too_near = 1;
		if (too_near) continue;
		p = grow_rows(p, &cap, n + 1);
		memset(&P(n, 0), 0, sizeof(double) * ORC_STRIDE);
		P(n, ORC_X) = x; P(n, ORC_Y) = y; P(n, ORC_Z) = z;
		P(n, ORC_M) = 1.0; P(n, ORC_R) = radius;
		n++;
	}
	if (n > i_0) finish_masses(p, i_0, n, total_mass);
	*pp = p;
	return n;
}

int orc_rand_rectangle_n2(double** pp, int n0, double min_dist, double x, double y, double z, double total_mass) { /* shapes.cpp:66-128 */
	const int n_part = 40.0 * x * y * z / pow(min_dist, 3.0);
	const double lo[3] = { -x / 2, -y / 2, -z / 2 }, hi[3] = { x / 2, y / 2, z / 2 };
	return rand_fill_n2(pp, n0, n_part, lo, hi, NULL, NULL, min_dist, (min_dist / 2.0) / 3.0, total_mass);
}

typedef struct { double radius, height, slope; } cone_t;
static int in_cone(const void* ctx, const double* p, double x, double y, double z) { /* shapes.cpp:168-173, GNU C++: contracted */
	(void) p;
	const cone_t* c = (const cone_t*) ctx;
	const double cur_rad = sqrt(fma(x, x, y * y));
	const double max_height = fma(-cur_rad, c->slope, c->height);
	return (cur_rad < c->radius) && (z < max_height);
}
int orc_rand_cone_n2(double** pp, int n0, double min_dist, double radius, double height, double total_mass) { /* shapes.cpp:132-206 */
	cone_t c = { radius, height, height / radius };
	const int n_part = 40 * pow(2.0 * radius / min_dist, 3.0);
	const double lo[3] = { -radius, -radius, 0.0 }, hi[3] = { radius, radius, height };
	return rand_fill_n2(pp, n0, n_part, lo, hi, in_cone, &c, min_dist, (min_dist / 2.0) / 2.0, total_mass);
}

typedef struct { int n_shape; double r_min, r_max; } shape_t;
static int in_shape(const void* ctx, const double* p, double x, double y, double z) { /* within_shape + nearest_to_point, shapes.cpp:720-774 */
	const shape_t* s = (const shape_t*) ctx;
	const double rad = len3(x, y, z);
	if (rad < s->r_min) return 1;
	if (rad > s->r_max) return 0;
	double dist = DBL_MAX;
	int i_0 = -1;
	for (int i = 0; i < s->n_shape; i++) {
		const double r = len3(P(i, ORC_X) - x, P(i, ORC_Y) - y, P(i, ORC_Z) - z);
		if (r < dist) { i_0 = i; dist = r; }
	}
	return !(rad > len3(P(i_0, ORC_X), P(i_0, ORC_Y), P(i_0, ORC_Z)));
}
int orc_rand_shape_n2(double** pp, int n0, double min_dist, double total_mass) { /* shapes.cpp:435-516 */
	const double* p = *pp;
	shape_t s = { n0, DBL_MAX, 0.0 };
	for (int i = 0; i < n0; i++) { /* min_radius / max_radius, shapes.cpp:777-811 */
		const double r = len3(P(i, ORC_X), P(i, ORC_Y), P(i, ORC_Z));
		if (r < s.r_min) s.r_min = r;
		if (r > s.r_max) s.r_max = r;
	}
	const int n_part = 40 * pow(2.0 * s.r_max / min_dist, 3.0);
	const double lo[3] = { -s.r_max, -s.r_max, -s.r_max }, hi[3] = { s.r_max, s.r_max, s.r_max };
	return rand_fill_n2(pp, n0, n_part, lo, hi, in_shape, &s, min_dist, min_dist / 4.0, total_mass);
}

/* ------------------------------------------------------------------------------------------------ */
/* set-up rotations, physics.cpp:249-331; Matrix arithmetic as gcc compiles matrix_math.cpp:442-450, 582-592 (res += a*b  */
/* chains contracted into fused multiply-adds, the first product plain)                                                    */
/* ------------------------------------------------------------------------------------------------ */
static void mat_mul(const double* a, const double* b, double* out) {
	for (int i = 0; i < 3; i++)
		for (int j = 0; j < 3; j++) out[3 * i + j] = fma(a[3 * i + 2], b[6 + j], fma(a[3 * i + 1], b[3 + j], a[3 * i] * b[j]));
}
static void mat_vec(const double* a, const double* v, double* out) {
	for (int i = 0; i < 3; i++) out[i] = fma(a[3 * i + 2], v[2], fma(a[3 * i + 1], v[1], a[3 * i] * v[0]));
}
void orc_euler_matrix(double alpha, double beta, double gamma, double* out9) {
	const double Rz1[9] = { cos(alpha), -sin(alpha), 0, sin(alpha), cos(alpha), 0, 0, 0, 1 };
	const double Rx[9] = { 1, 0, 0, 0, cos(beta), -sin(beta), 0, sin(beta), cos(beta) };
	const double Rz2[9] = { cos(gamma), -sin(gamma), 0, sin(gamma), cos(gamma), 0, 0, 0, 1 };
	double t[9];
	mat_mul(Rz2, Rx, t);     /* Rz2 * Rx * Rz1 associates left to right (physics.cpp:274) */
	mat_mul(t, Rz1, out9);
}
static void rotate_about(double* p, int i_low, int i_high, const double* R, const double* c, const double* u) {
	for (int i = i_low; i < i_high; i++) {
		const double x0[3] = { P(i, ORC_X) - c[0], P(i, ORC_Y) - c[1], P(i, ORC_Z) - c[2] };
		const double v0[3] = { P(i, ORC_VX) - u[0], P(i, ORC_VY) - u[1], P(i, ORC_VZ) - u[2] };
		double xr[3], vr[3];
		mat_vec(R, x0, xr);
		mat_vec(R, v0, vr);
		P(i, ORC_X) = xr[0] + c[0]; P(i, ORC_Y) = xr[1] + c[1]; P(i, ORC_Z) = xr[2] + c[2];
		P(i, ORC_VX) = vr[0] + u[0]; P(i, ORC_VY) = vr[1] + u[1]; P(i, ORC_VZ) = vr[2] + u[2];
	}
}
void orc_rotate_body(double* p, int i_low, int i_high, double alpha, double beta, double gamma) { /* physics.cpp:249-290 */
	double com[3], cov[3], R[9];
	orc_compute_com(p, i_low, i_high, com);
	orc_compute_cov(p, i_low, i_high, cov);
	orc_euler_matrix(alpha, beta, gamma, R);
	rotate_about(p, i_low, i_high, R, com, cov);
}
void orc_rotate_origin(double* p, int i_low, int i_high, double alpha, double beta, double gamma) { /* physics.cpp:294-331 */
	double R[9];
	const double zero[3] = { 0.0, 0.0, 0.0 };
	orc_euler_matrix(alpha, beta, gamma, R);
	for (int i = i_low; i < i_high; i++) { /* no centre: nothing is subtracted or added back */
		const double x0[3] = { P(i, ORC_X), P(i, ORC_Y), P(i, ORC_Z) }, v0[3] = { P(i, ORC_VX), P(i, ORC_VY), P(i, ORC_VZ) };
		double xr[3], vr[3];
		mat_vec(R, x0, xr);
		mat_vec(R, v0, vr);
		P(i, ORC_X) = xr[0]; P(i, ORC_Y) = xr[1]; P(i, ORC_Z) = xr[2];
		P(i, ORC_VX) = vr[0]; P(i, ORC_VY) = vr[1]; P(i, ORC_VZ) = vr[2];
	}
	(void) zero;
}

/* ------------------------------------------------------------------------------------------------ */
/* stress tensors, stress.cpp:32-85 with spring_i_force (springs.cpp:350-392)                         */
/* ------------------------------------------------------------------------------------------------ */
void orc_stress(const double* p, int n, const orc_spring* s, int ns, double* tensors, double* max_force, int* max_spring) {
	for (int i = 0; i < n; i++) {
		for (int q = 0; q < 9; q++) tensors[9 * (size_t) i + q] = 0.0;
		max_force[i] = 0.0;
		max_spring[i] = -1;
	}
	for (int k = 0; k < ns; k++) {
		const int i = s[k].particle_1, j = s[k].particle_2;
		const double L[3] = { P(i, ORC_X) - P(j, ORC_X), P(i, ORC_Y) - P(j, ORC_Y), P(i, ORC_Z) - P(j, ORC_Z) };
		const double len = len3(L[0], L[1], L[2]);
		const double hat[3] = { L[0] / len, L[1] / len, L[2] / len };
		const double c = -s[k].k * (len - s[k].rs0);
		double F[3] = { hat[0] * c, hat[1] * c, hat[2] * c };
		if (s[k].gamma > 0.0) {
			const double dv[3] = { P(i, ORC_VX) - P(j, ORC_VX), P(i, ORC_VY) - P(j, ORC_VY), P(i, ORC_VZ) - P(j, ORC_VZ) };
			const double strain_rate = fma(dv[2], hat[2], fma(dv[0], hat[0], dv[1] * hat[1])) / len;
			const double m_red = P(i, ORC_M) * P(j, ORC_M) / (P(i, ORC_M) + P(j, ORC_M));
			const double g = s[k].gamma * m_red * strain_rate;
			for (int a = 0; a < 3; a++) F[a] -= hat[a] * g;
		}
		for (int a = 0; a < 3; a++)
			for (int b = 0; b < 3; b++) {
				const double o = F[a] * L[b];
				tensors[9 * (size_t) i + 3 * a + b] += o;
				tensors[9 * (size_t) j + 3 * a + b] -= o;
			}
		const double F_mag = len3(F[0], F[1], F[2]);
		if (F_mag > max_force[i]) { max_force[i] = F_mag; max_spring[i] = k; }
		if (F_mag > max_force[j]) { max_force[j] = F_mag; max_spring[j] = k; }
	}
	const double vol = 4.0 * M_PI / (3.0 * (double) n);
	for (int i = 0; i < n; i++)
		for (int q = 0; q < 9; q++) tensors[9 * (size_t) i + q] /= vol;
}

/* ------------------------------------------------------------------------------------------------ */
/* module-local force terms                                                                          */
/* ------------------------------------------------------------------------------------------------ */
int orc_apply_impact(double* p, int n_body, const unsigned char* is_surf, double lat, double longi, double dangle, double force, double envelope) {
	/* modules/bennu2/problem.cpp:386-447 without the display-radius bookkeeping; envelope = 1 - cos(2 pi t / tau) */
	const double x0[3] = { cos(longi) * cos(lat), sin(longi) * cos(lat), sin(lat) };
	int num_pushed = 0;
	for (int pass = 0; pass < 2; pass++)
		for (int i = 0; i < n_body; i++) {
			if (!is_surf[i]) continue;
			const double len = len3(P(i, ORC_X), P(i, ORC_Y), P(i, ORC_Z));
			const double hat[3] = { P(i, ORC_X) / len, P(i, ORC_Y) / len, P(i, ORC_Z) / len };
			const double projection = fma(x0[2], hat[2], fma(x0[0], hat[0], x0[1] * hat[1]));
			if (!(acos(projection) < dangle)) continue;
			if (pass == 0) { num_pushed++; continue; }
			const double den = P(i, ORC_M) * num_pushed * len;
			P(i, ORC_AX) -= hat[0] * force / den * envelope;
			P(i, ORC_AY) -= hat[1] * force / den * envelope;
			P(i, ORC_AZ) -= hat[2] * force / den * envelope;
		}
	return num_pushed;
}
void orc_top_push(double* p, int n, const unsigned char* top, double pressure) { /* modules/cube/problem.cpp:230-234 */
	int n_surf = 0;
	for (int i = 0; i < n; i++) n_surf += top[i] ? 1 : 0;
	for (int i = 0; i < n; i++)
		if (top[i]) P(i, ORC_AY) -= pressure / n_surf / P(i, ORC_M);
}
void orc_table_top(double* p, int n, const unsigned char* bottom) { /* modules/cube/problem.cpp:259-263 */
	for (int i = 0; i < n; i++)
		if (bottom[i]) P(i, ORC_AY) = 0.0;
}
