// generators.cpp -- host side of the build path: the node generators of reference src_spring/shapes.cpp.
//
// rand_ellipsoid (shapes.cpp:210-278) is sequential by construction -- one libc rand() stream, and each accept test
// depends on every earlier accept -- so it stays on the host.  What made it O(n_candidates x N) in the reference
// (80 s at N = 9.3k) is the linear "is any accepted node within min_dist" scan; here that question is
// answered from a hash of occupied grid cells, which changes the cost, not the answer: the candidate stream, both
// accept tests and their rounding (Vector::len = sqrt(fma(z,z,fma(x,x,y*y)))) are the reference's, so the
// same srand() seed yields the same particles bit for bit.
//
// Rows are packed (n,11) doubles: x y z vx vy vz ax ay az m r  (the leading fields of struct reb_particle).
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <unordered_map>
#include <vector>

#include "../../include/ass_b200.h"

void ass_set_error(const char* fmt, ...);

namespace {

constexpr int STRIDE = 11;

inline double vlen(double x, double y, double z) { return sqrt(fma(z, z, fma(x, x, y * y))); }

// reb_random_uniform, src/tools.c:44-46 (ISO C build: no contraction)
inline double random_uniform(double lo, double hi) {
	volatile double u = ((double) rand()) / ((double) (RAND_MAX));
	volatile double w = u * (hi - lo);
	return w + lo;
}

struct Rows {
	double** base;
	int* cap;
	int n;
	bool push(double x, double y, double z, double m, double r) {
		if (n >= *cap) {
			int nc = std::max(1024, *cap * 2);
			double* nb = (double*) realloc(*base, sizeof(double) * STRIDE * (size_t) nc);
			if (!nb) return false;
			*base = nb;
			*cap = nc;
		}
		double* row = *base + (size_t) n * STRIDE;
		memset(row, 0, sizeof(double) * STRIDE);
		row[0] = x; row[1] = y; row[2] = z; row[9] = m; row[10] = r;
		n++;
		return true;
	}
};

// Dense 3-D bucket grid over a box; cell edge slightly above the query radius so a 3x3x3 stencil is exhaustive.
class BucketGrid {
public:
	BucketGrid(const double lo[3], const double hi[3], double radius, double max_cells) {
		h_ = radius * (1.0 + 1e-9);
		for (;;) {
			double cells = 1.0;
			for (int d = 0; d < 3; d++) cells *= floor((hi[d] - lo[d]) / h_) + 1.0;
			if (cells <= max_cells) break;
			h_ *= 1.5;
		}
		for (int d = 0; d < 3; d++) {
			lo_[d] = lo[d];
			n_[d] = (int) floor((hi[d] - lo[d]) / h_) + 1;
		}
		head_.assign((size_t) n_[0] * n_[1] * n_[2], -1);
	}
	void coords(const double p[3], int c[3]) const {
		for (int d = 0; d < 3; d++) {
			int v = (int) floor((p[d] - lo_[d]) / h_);
			c[d] = v < 0 ? 0 : (v >= n_[d] ? n_[d] - 1 : v);
		}
	}
	size_t flat(const int c[3]) const { return ((size_t) c[2] * n_[1] + c[1]) * n_[0] + c[0]; }
	void insert(const double p[3], int id) {
		int c[3];
		coords(p, c);
		if ((size_t) id >= next_.size()) next_.resize(std::max<size_t>(1024, 2 * (size_t) id + 1));
		next_[id] = head_[flat(c)];
		head_[flat(c)] = id;
	}
	template <class F>
	bool any_neighbour(const double p[3], F&& pred) const {  // true as soon as pred(id) is
		int c[3], q[3];
		coords(p, c);
		for (q[2] = std::max(c[2] - 1, 0); q[2] <= std::min(c[2] + 1, n_[2] - 1); q[2]++)
			for (q[1] = std::max(c[1] - 1, 0); q[1] <= std::min(c[1] + 1, n_[1] - 1); q[1]++)
				for (q[0] = std::max(c[0] - 1, 0); q[0] <= std::min(c[0] + 1, n_[0] - 1); q[0]++)
					for (int id = head_[flat(q)]; id >= 0; id = next_[id])
						if (pred(id)) return true;
		return false;
	}
	double edge() const { return h_; }
	bool single_cell() const { return n_[0] == 1 && n_[1] == 1 && n_[2] == 1; }

private:
	double h_, lo_[3];
	int n_[3];
	std::vector<int> head_, next_;
};

// mindist (shapes.cpp:814-833) on host rows [lo,hi): min is exact and order-free.
double host_mindist(const double* rows, int lo, int hi) {
	const int cnt = hi - lo;
	if (cnt < 2) return DBL_MAX;
	double bl[3] = { DBL_MAX, DBL_MAX, DBL_MAX }, bh[3] = { -DBL_MAX, -DBL_MAX, -DBL_MAX };
	for (int i = lo; i < hi; i++)
		for (int d = 0; d < 3; d++) {
			bl[d] = std::min(bl[d], rows[(size_t) i * STRIDE + d]);
			bh[d] = std::max(bh[d], rows[(size_t) i * STRIDE + d]);
		}
	double h = 1e-300;
	for (int attempt = 0; attempt < 64; attempt++) {
		BucketGrid g(bl, bh, h, 8.0 * cnt + 4096.0);
		double best = DBL_MAX;
		for (int i = lo; i < hi; i++) {
			const double* p = rows + (size_t) i * STRIDE;
			g.any_neighbour(p, [&](int id) {
				const double* q = rows + (size_t) (lo + id) * STRIDE;
				best = std::min(best, vlen(q[0] - p[0], q[1] - p[1], q[2] - p[2]));
				return false;
			});
			g.insert(p, i - lo);  // earlier particles only: every pair once
		}
		if (best < g.edge() * (1.0 - 1e-6) || g.single_cell()) return best;
		h = g.edge() * 2.0;
	}
	return DBL_MAX;
}

int finish_lattice(Rows& R, int i_0, double total_mass) {
	// shapes.cpp:337-346 / 416-425: equal masses; radius = (smallest pair distance) / 4
	const int n = R.n;
	const double particle_mass = total_mass / (n - i_0);
	const double min_d = host_mindist(*R.base, i_0, n);
	for (int i = i_0; i < n; i++) {
		(*R.base)[(size_t) i * STRIDE + 9] = particle_mass;
		(*R.base)[(size_t) i * STRIDE + 10] = min_d / 4.0;
	}
	return n;
}

}  // namespace

extern "C" int ass_rand_ellipsoid(double** rows, int* cap_rows, int n0, double min_dist, double a, double b, double c, double total_mass) {
	if (!rows || !cap_rows || n0 < 0 || !(min_dist > 0.0) || !(a > 0.0) || !(b > 0.0) || !(c > 0.0)) {
		ass_set_error("ass_rand_ellipsoid: bad argument");
		return ASS_EINVAL;
	}
	Rows R{ rows, cap_rows, n0 };
	const int i_0 = n0;
	const int n_part = 40 * pow(2.0 * a / min_dist, 3.0);  // shapes.cpp:214 (double -> int truncation)
	const double radius = (min_dist / 2.0) / 2.0;          // shapes.cpp:227-228
	const double lo[3] = { -a, -b, -c }, hi[3] = { a, b, c };
	BucketGrid grid(lo, hi, min_dist, 2.5e8);
	for (int t = 0; t < n_part; t++) {
		double p[3];
		p[0] = random_uniform(-a, a);  // shapes.cpp:238-240: x, then y, then z
		p[1] = random_uniform(-b, b);
		p[2] = random_uniform(-c, c);
		if (!(vlen(p[0] / a, p[1] / b, p[2] / c) < 1.0)) continue;  // :244
		const double* base = *rows;
		const bool too_near = grid.any_neighbour(p, [&](int id) {  // :249-255, strict <
			const double* q = base + (size_t) (i_0 + id) * STRIDE;
			return vlen(p[0] - q[0], p[1] - q[1], p[2] - q[2]) < min_dist;
		});
		if (too_near) continue;
		if (!R.push(p[0], p[1], p[2], 1.0, radius)) {
			ass_set_error("ass_rand_ellipsoid: out of memory");
			return ASS_ENOMEM;
		}
		grid.insert(p, R.n - 1 - i_0);
	}
	if (R.n > i_0) {
		const double particle_mass = total_mass / (R.n - i_0);  // :264-268
		for (int i = i_0; i < R.n; i++) (*rows)[(size_t) i * STRIDE + 9] = particle_mass;
	}
	return R.n;
}

// The common loop of rand_rectangle / rand_cone / rand_shape (shapes.cpp:66-208, 435-516): n_candidates darts, each
// x = U(lo0, hi0), y = U(lo1, hi1), z = U(lo2, hi2) in that order from libc rand(); kept if `inside` says so and no
// particle created by this call lies within min_dist (strict <, Vector::len); m = 1 and r = `radius` until the end,
// then m = total_mass / count.  The "too near" scan of the reference is linear in the particles created so far; here
// it is a 27-cell lookup -- same answer.
extern "C" int ass_rand_fill(double** rows, int* cap_rows, int n0, long long n_candidates, const double lo[3], const double hi[3],
		ass_inside_fn inside, void* ctx, double min_dist, double radius, double total_mass) {
	if (!rows || !cap_rows || n0 < 0 || !lo || !hi || !(min_dist > 0.0) || n_candidates < 0) {
		ass_set_error("ass_rand_fill: bad argument");
		return ASS_EINVAL;
	}
	Rows R{ rows, cap_rows, n0 };
	const int i_0 = n0;
	BucketGrid grid(lo, hi, min_dist, 2.5e8);
	for (long long t = 0; t < n_candidates; t++) {
		double p[3];
		p[0] = random_uniform(lo[0], hi[0]);
		p[1] = random_uniform(lo[1], hi[1]);
		p[2] = random_uniform(lo[2], hi[2]);
		if (inside && !inside(ctx, p[0], p[1], p[2])) continue;
		const double* base = *rows;
		const bool too_near = grid.any_neighbour(p, [&](int id) {
			const double* q = base + (size_t) (i_0 + id) * STRIDE;
			return vlen(p[0] - q[0], p[1] - q[1], p[2] - q[2]) < min_dist;
		});
		if (too_near) continue;
		if (!R.push(p[0], p[1], p[2], 1.0, radius)) {
			ass_set_error("ass_rand_fill: out of memory");
			return ASS_ENOMEM;
		}
		grid.insert(p, R.n - 1 - i_0);
	}
	if (R.n > i_0) {
		const double particle_mass = total_mass / (R.n - i_0);
		for (int i = i_0; i < R.n; i++) (*rows)[(size_t) i * STRIDE + 9] = particle_mass;
	}
	return R.n;
}

// rand_rectangle, shapes.cpp:66-128: box of edge lengths x, y, z about the origin
extern "C" int ass_rand_rectangle(double** rows, int* cap_rows, int n0, double min_dist, double x, double y, double z, double total_mass) {
	if (!(x > 0.0) || !(y > 0.0) || !(z > 0.0) || !(min_dist > 0.0)) {
		ass_set_error("ass_rand_rectangle: bad argument");
		return ASS_EINVAL;
	}
	const int n_part = 40.0 * x * y * z / pow(min_dist, 3.0);  // :70 (double -> int truncation)
	const double lo[3] = { -x / 2, -y / 2, -z / 2 }, hi[3] = { x / 2, y / 2, z / 2 };
	return ass_rand_fill(rows, cap_rows, n0, n_part, lo, hi, nullptr, nullptr, min_dist, (min_dist / 2.0) / 3.0, total_mass);  // r: :83-84
}

namespace {
struct Cone { double radius, height, slope; };
int inside_cone(void* ctx, double x, double y, double z) {  // shapes.cpp:168-173, with the contractions gcc applies in GNU C++ mode
	const Cone* c = (const Cone*) ctx;
	const double cur_rad = sqrt(fma(x, x, y * y));           // gcc contracts pt.x*pt.x + pt.y*pt.y into fma(x, x, y*y)
	const double max_height = fma(-cur_rad, c->slope, c->height);  // height - cur_rad * slope, contracted
	return (cur_rad < c->radius) && (z < max_height);
}
}  // namespace

// rand_cone, shapes.cpp:132-206: base radius and height, apex up, base in the plane z = 0
extern "C" int ass_rand_cone(double** rows, int* cap_rows, int n0, double min_dist, double radius, double height, double total_mass) {
	if (!(radius > 0.0) || !(height > 0.0) || !(min_dist > 0.0)) {
		ass_set_error("ass_rand_cone: bad argument");
		return ASS_EINVAL;
	}
	Cone c{ radius, height, height / radius };  // :136
	const int n_part = 40 * pow(2.0 * radius / min_dist, 3.0);  // :139
	const double lo[3] = { -radius, -radius, 0.0 }, hi[3] = { radius, radius, height };
	return ass_rand_fill(rows, cap_rows, n0, n_part, lo, hi, inside_cone, &c, min_dist, (min_dist / 2.0) / 2.0, total_mass);  // r: :153-154
}

namespace {
// within_shape (shapes.cpp:720-748) over the vertex rows [0, n_shape): inside the smallest vertex radius -> in; outside
// the largest -> out; otherwise compare with the radius of the NEAREST vertex (nearest_to_point, :752-774: first index
// of the strictly smallest distance).  The nearest vertex is found through a bucket grid over the vertices instead of a
// scan over all of them: cells are visited in shells around the point until no closer vertex can exist, ties go to the
// lowest index -- the vertex the linear scan returns.
struct Shape {
	const double* rows;
	int n;
	double r_min, r_max, h, lo[3];
	int dim[3];
	std::vector<int> start, ids;
	int cell(const double p[3], int c[3]) const {
		for (int d = 0; d < 3; d++) {
			int v = (int) floor((p[d] - lo[d]) / h);
			c[d] = v < 0 ? 0 : (v >= dim[d] ? dim[d] - 1 : v);
		}
		return (c[2] * dim[1] + c[1]) * dim[0] + c[0];
	}
	int nearest(double x, double y, double z) const {
		const double p[3] = { x, y, z };
		int c[3];
		cell(p, c);
		double best = DBL_MAX;
		int best_i = -1;
		const int max_shell = std::max(dim[0], std::max(dim[1], dim[2]));
		for (int sh = 0; sh <= max_shell; sh++) {
			// every vertex in a cell of shell `sh` is at least (sh - 1) * h away: along some axis its cell index differs by sh
			// from the point's (clamped) cell, and a point outside the box is only farther from that side
			if (best_i >= 0 && (sh - 1) * h > best) break;
			for (int k = c[2] - sh; k <= c[2] + sh; k++) {
				if (k < 0 || k >= dim[2]) continue;
				for (int j = c[1] - sh; j <= c[1] + sh; j++) {
					if (j < 0 || j >= dim[1]) continue;
					const bool edge_kj = (k == c[2] - sh || k == c[2] + sh || j == c[1] - sh || j == c[1] + sh);
					for (int i = c[0] - sh; i <= c[0] + sh; i += (edge_kj ? 1 : std::max(1, 2 * sh))) {
						if (i < 0 || i >= dim[0]) continue;
						const int cc = (k * dim[1] + j) * dim[0] + i;
						for (int q = start[cc]; q < start[cc + 1]; q++) {
							const int id = ids[q];
							const double* v = rows + (size_t) id * STRIDE;
							const double r = vlen(v[0] - x, v[1] - y, v[2] - z);  // Vector dx = part_pos - pos; dx.len()
							if (r < best || (r == best && id < best_i)) { best = r; best_i = id; }
						}
					}
				}
			}
		}
		return best_i;
	}
};
int inside_shape(void* ctx, double x, double y, double z) {
	const Shape* S = (const Shape*) ctx;
	const double rad = vlen(x, y, z);
	if (rad < S->r_min) return 1;
	if (rad > S->r_max) return 0;
	const int i = S->nearest(x, y, z);
	const double* v = S->rows + (size_t) i * STRIDE;
	return !(rad > vlen(v[0], v[1], v[2]));
}
}  // namespace

// rand_shape, shapes.cpp:435-516: rows [0, n0) are the vertices of the shape model (already "read in"); new particles are
// appended after them.
extern "C" int ass_rand_shape(double** rows, int* cap_rows, int n0, double min_dist, double total_mass) {
	if (!rows || !cap_rows || n0 <= 0 || !*rows || !(min_dist > 0.0)) {
		ass_set_error("ass_rand_shape: needs the shape's vertices as rows [0, n0) and a positive min_dist");
		return ASS_EINVAL;
	}
	// the vertex rows may move when the array grows: work on a copy
	std::vector<double> verts(*rows, *rows + (size_t) n0 * STRIDE);
	Shape S;
	S.rows = verts.data();
	S.n = n0;
	S.r_min = DBL_MAX;  // min_radius / max_radius, shapes.cpp:777-811
	S.r_max = 0.0;
	double bl[3] = { DBL_MAX, DBL_MAX, DBL_MAX }, bh[3] = { -DBL_MAX, -DBL_MAX, -DBL_MAX };
	for (int i = 0; i < n0; i++) {
		const double* v = verts.data() + (size_t) i * STRIDE;
		const double r = vlen(v[0], v[1], v[2]);
		if (r < S.r_min) S.r_min = r;
		if (r > S.r_max) S.r_max = r;
		for (int d = 0; d < 3; d++) { bl[d] = std::min(bl[d], v[d]); bh[d] = std::max(bh[d], v[d]); }
	}
	// ~2 vertices per cell
	double vol = 1.0;
	for (int d = 0; d < 3; d++) vol *= std::max(bh[d] - bl[d], 1e-300);
	S.h = std::max(cbrt(vol * 2.0 / n0), 1e-12 * S.r_max + 1e-300);
	for (int d = 0; d < 3; d++) {
		S.lo[d] = bl[d];
		S.dim[d] = std::min(1024, (int) floor((bh[d] - bl[d]) / S.h) + 1);
	}
	const size_t nc = (size_t) S.dim[0] * S.dim[1] * S.dim[2];
	S.start.assign(nc + 1, 0);
	std::vector<int> cell_of((size_t) n0);
	for (int i = 0; i < n0; i++) {
		int c[3];
		cell_of[(size_t) i] = S.cell(verts.data() + (size_t) i * STRIDE, c);
		S.start[(size_t) cell_of[(size_t) i] + 1]++;
	}
	for (size_t c = 0; c < nc; c++) S.start[c + 1] += S.start[c];
	S.ids.resize((size_t) n0);
	{
		std::vector<int> fill(S.start.begin(), S.start.end() - 1);
		for (int i = 0; i < n0; i++) S.ids[(size_t) fill[(size_t) cell_of[(size_t) i]]++] = i;
	}
	const int n_part = 40 * pow(2.0 * S.r_max / min_dist, 3.0);  // :447
	const double lo[3] = { -S.r_max, -S.r_max, -S.r_max }, hi[3] = { S.r_max, S.r_max, S.r_max };
	return ass_rand_fill(rows, cap_rows, n0, n_part, lo, hi, inside_shape, &S, min_dist, min_dist / 4.0, total_mass);  // r: :462
}

extern "C" int ass_hcp_ellipsoid(double** rows, int* cap_rows, int n0, double min_dist, double a, double b, double c, double total_mass) {
	if (!rows || !cap_rows || n0 < 0 || !(min_dist > 0.0) || !(a > 0.0) || !(b > 0.0) || !(c > 0.0)) {
		ass_set_error("ass_hcp_ellipsoid: bad argument");
		return ASS_EINVAL;
	}
	Rows R{ rows, cap_rows, n0 };
	// shapes.cpp:300-310
	const double dx = 1.08 * min_dist;
	const double xfac = dx / 2.0, yfac = dx * sin(M_PI / 3.0), zfac = dx * sqrt(2.0 / 3.0);
	const int nx = (int) (1.2 * a / dx), ny = (int) (1.2 * b / yfac), nz = (int) (1.2 * c / zfac);
	for (int k = -nz; k <= nz; k++) {
		const double z = zfac * k;
		for (int j = -ny; j <= ny; j++) {
			// shapes.cpp is GNU C++: gcc contracts these sums into the FMAs written out here
			const double y = yfac * fma(0.5, (double) (k % 2), (double) j);                                   // :323
			for (int i = -nx; i <= nx; i++) {
				// :328 -- gcc hoists the two loop-invariant products and fuses only dx*i: fma(dx, i, t_j) + t_k
				const double x = fma(dx, (double) i, xfac * (double) (j % 2)) + xfac * (double) (k % 2);
				if (vlen(x / a, y / b, z / c) <= 1.0)                                                           // :331-332
					if (!R.push(x, y, z, 1.0, dx / 2.0)) { ass_set_error("ass_hcp_ellipsoid: out of memory"); return ASS_ENOMEM; }
			}
		}
	}
	if (R.n == n0) return R.n;
	return finish_lattice(R, n0, total_mass);
}

extern "C" int ass_cubic_ellipsoid(double** rows, int* cap_rows, int n0, double min_dist, double a, double b, double c, double total_mass) {
	if (!rows || !cap_rows || n0 < 0 || !(min_dist > 0.0) || !(a > 0.0) || !(b > 0.0) || !(c > 0.0)) {
		ass_set_error("ass_cubic_ellipsoid: bad argument");
		return ASS_EINVAL;
	}
	Rows R{ rows, cap_rows, n0 };
	// shapes.cpp:373-404
	const int nx = (int) (1.2 * a / min_dist), ny = (int) (1.2 * b / min_dist), nz = (int) (1.2 * c / min_dist);
	for (int k = -nz; k <= nz; k++) {
		const double z = min_dist * k;
		for (int j = -ny; j <= ny; j++) {
			const double y = min_dist * j;
			for (int i = -nx; i <= nx; i++) {
				const double x = min_dist * i;
				if (vlen(x / a, y / b, z / c) <= 1.0)
					if (!R.push(x, y, z, 1.0, 1.0)) { ass_set_error("ass_cubic_ellipsoid: out of memory"); return ASS_ENOMEM; }
			}
		}
	}
	if (R.n == n0) return R.n;
	return finish_lattice(R, n0, total_mass);
}
