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
