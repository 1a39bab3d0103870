// Host-side grid providers: stand-ins for Stuff::Grid::Providers::Cube<SGrid<2,2>> and for the
// ALUGrid<2,2,simplex,conforming> refinement ladder of the test cases (testcases/ESV2007.hh:50-59,123-134,
// testcases/base.hh:92-103), including the [px py 1] subdomain partition of grid::Multiscale
// (testcases/ESV2007.hh:150-163).  Closed form, O(n), no recursion; tests/ compares the simplex generator
// against an actual recursive longest-edge bisection.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <vector>

#include "common.hpp"

namespace {

// subdomain of a point: axis-aligned px x py boxes, x fastest (assumption: box numbering is decided upstream by
// dune-grid-multiscale and no test pins it).
inline int box_of(double c, double lo, double hi, int parts) {
  int b = int((c - lo) / (hi - lo) * parts);
  return std::min(std::max(b, 0), parts - 1);
}

// Stable renumbering that makes cells subdomain-major. perm[old] = new.
std::vector<int32_t> subdomain_major(const std::vector<int32_t>& sub, int n_sub) {
  std::vector<int64_t> start(size_t(n_sub) + 1, 0);
  for (int32_t s : sub) ++start[size_t(s) + 1];
  for (int s = 0; s < n_sub; ++s) start[size_t(s) + 1] += start[size_t(s)];
  std::vector<int32_t> perm(sub.size());
  for (size_t c = 0; c < sub.size(); ++c) perm[c] = int32_t(start[size_t(sub[c])]++);
  return perm;
}

}  // namespace

extern "C" {

int hdd_grid_cube_sizes(int64_t nx, int64_t ny, int64_t* n_cells, int64_t* n_verts) {
  return hdd::guarded([&] {
    if (nx < 1 || ny < 1 || nx * ny > INT32_MAX / 4) HDD_THROW(HDD_ERR_WRONG_INPUT, "bad grid size " << nx << " x " << ny);
    *n_cells = nx * ny;
    *n_verts = (nx + 1) * (ny + 1);
  });
}

int hdd_grid_cube(int64_t nx, int64_t ny, double x0, double x1, double y0, double y1, int px, int py, double* xy,
                  int32_t* cv, int32_t* nb, int32_t* cell_subdomain) {
  return hdd::guarded([&] {
    if (nx < 1 || ny < 1 || px < 1 || py < 1) HDD_THROW(HDD_ERR_WRONG_INPUT, "bad grid size");
    for (int64_t j = 0; j <= ny; ++j)
      for (int64_t i = 0; i <= nx; ++i) {
        xy[2 * (j * (nx + 1) + i)] = x0 + (x1 - x0) * double(i) / double(nx);
        xy[2 * (j * (nx + 1) + i) + 1] = y0 + (y1 - y0) * double(j) / double(ny);
      }
    const bool part = px > 1 || py > 1;
    std::vector<int32_t> perm;
    if (part) {
      std::vector<int32_t> sub(size_t(nx * ny));
      for (int64_t j = 0; j < ny; ++j)
        for (int64_t i = 0; i < nx; ++i)
          sub[size_t(j * nx + i)] = box_of(j + 0.5, 0, double(ny), py) * px + box_of(i + 0.5, 0, double(nx), px);
      perm = subdomain_major(sub, px * py);
      if (cell_subdomain)
        for (size_t c = 0; c < sub.size(); ++c) cell_subdomain[perm[c]] = sub[c];
    } else if (cell_subdomain) {
      std::fill(cell_subdomain, cell_subdomain + nx * ny, 0);
    }
    auto id = [&](int64_t i, int64_t j) -> int32_t {
      if (i < 0 || j < 0 || i >= nx || j >= ny) return -1;
      const int64_t c = j * nx + i;
      return part ? perm[size_t(c)] : int32_t(c);
    };
    for (int64_t j = 0; j < ny; ++j)
      for (int64_t i = 0; i < nx; ++i) {
        const int64_t c = id(i, j);
        cv[4 * c + 0] = int32_t(j * (nx + 1) + i);
        cv[4 * c + 1] = int32_t(j * (nx + 1) + i + 1);
        cv[4 * c + 2] = int32_t((j + 1) * (nx + 1) + i);
        cv[4 * c + 3] = int32_t((j + 1) * (nx + 1) + i + 1);
        nb[4 * c + 0] = id(i - 1, j);
        nb[4 * c + 1] = id(i + 1, j);
        nb[4 * c + 2] = id(i, j - 1);
        nb[4 * c + 3] = id(i, j + 1);
      }
  });
}

int hdd_grid_simplex_sizes(int64_t s, int64_t* n_cells, int64_t* n_verts) {
  return hdd::guarded([&] {
    if (s < 1 || 8 * s * s > INT32_MAX / 4) HDD_THROW(HDD_ERR_WRONG_INPUT, "bad grid size " << s);
    *n_cells = 8 * s * s;
    *n_verts = (2 * s + 1) * (2 * s + 1);
  });
}

int hdd_grid_simplex(int64_t s, double x0, double x1, double y0, double y1, int px, int py, double* xy, int32_t* cv,
                     int32_t* nb, int32_t* cell_subdomain) {
  return hdd::guarded([&] {
    if (s < 1 || px < 1 || py < 1) HDD_THROW(HDD_ERR_WRONG_INPUT, "bad grid size");
    const int64_t L = 2 * s + 1;
    for (int64_t b = 0; b < L; ++b)
      for (int64_t a = 0; a < L; ++a) {
        xy[2 * (b * L + a)] = x0 + (x1 - x0) * double(a) / double(2 * s);
        xy[2 * (b * L + a) + 1] = y0 + (y1 - y0) * double(b) / double(2 * s);
      }
    // perimeter of a square, counter-clockwise from its lower-left corner, in half-square lattice steps
    static const int pa[8] = {0, 1, 2, 2, 2, 1, 0, 0};
    static const int pb[8] = {0, 0, 0, 1, 2, 2, 2, 1};
    static const int opp[8] = {5, 4, 7, 6, 1, 0, 3, 2};  // triangle across the perimeter edge, in the next square
    static const int di[8] = {0, 0, 1, 1, 0, 0, -1, -1};
    static const int dj[8] = {-1, -1, 0, 0, 1, 1, 0, 0};
    const int64_t nc = 8 * s * s;
    const bool part = px > 1 || py > 1;
    std::vector<int32_t> perm;
    if (part) {
      std::vector<int32_t> sub(size_t(nc), 0);
      for (int64_t J = 0; J < s; ++J)
        for (int64_t I = 0; I < s; ++I)
          for (int t = 0; t < 8; ++t) {
            // centroid in lattice units
            const double ca = ((2 * I + 1) + (2 * I + pa[t]) + (2 * I + pa[(t + 1) % 8])) / 3.0;
            const double cb = ((2 * J + 1) + (2 * J + pb[t]) + (2 * J + pb[(t + 1) % 8])) / 3.0;
            sub[size_t((J * s + I) * 8 + t)] =
                box_of(cb, 0, double(2 * s), py) * px + box_of(ca, 0, double(2 * s), px);
          }
      perm = subdomain_major(sub, px * py);
      if (cell_subdomain)
        for (size_t c = 0; c < sub.size(); ++c) cell_subdomain[perm[c]] = sub[c];
    } else if (cell_subdomain) {
      std::fill(cell_subdomain, cell_subdomain + nc, 0);
    }
    auto id = [&](int64_t I, int64_t J, int t) -> int32_t {
      if (I < 0 || J < 0 || I >= s || J >= s) return -1;
      const int64_t c = (J * s + I) * 8 + t;
      return part ? perm[size_t(c)] : int32_t(c);
    };
    for (int64_t J = 0; J < s; ++J)
      for (int64_t I = 0; I < s; ++I)
        for (int t = 0; t < 8; ++t) {
          const int64_t c = id(I, J, t);
          const int t1 = (t + 1) % 8;
          cv[3 * c + 0] = int32_t((2 * J + 1) * L + (2 * I + 1));
          cv[3 * c + 1] = int32_t((2 * J + pb[t]) * L + (2 * I + pa[t]));
          cv[3 * c + 2] = int32_t((2 * J + pb[t1]) * L + (2 * I + pa[t1]));
          nb[3 * c + 0] = id(I, J, (t + 7) % 8);            // face {0,1}: centre - P_t
          nb[3 * c + 1] = id(I, J, t1);                     // face {0,2}: centre - P_{t+1}
          nb[3 * c + 2] = id(I + di[t], J + dj[t], opp[t]); // face {1,2}: perimeter edge
        }
  });
}

// Father cells for the prolongation of the convergence studies.  The reference walks ALUGrid's father() pointers
// (test/linearelliptic-swipdg.hh:186-194) or searches the coarse grid view for the centre of each fine cell
// (Stuff::Grid::EntityInlevelSearch, test/linearelliptic-block-swipdg.hh:169-177); flat arrays carry no hierarchy, so
// this is the search: coarse cells are binned into a uniform bucket grid by bounding box, every fine centre tests the
// cells of its bucket and takes the one it lies deepest inside (ties on shared faces resolve to the lowest id).
int hdd_grid_fathers(int kind, int64_t n_coarse, int64_t n_coarse_verts, const double* xy_c, const int32_t* cv_c,
                     int64_t n_fine, int64_t n_fine_verts, const double* xy_f, const int32_t* cv_f, int32_t* father) {
  return hdd::guarded([&] {
    if (kind != HDD_SIMPLEX2D && kind != HDD_CUBE2D) HDD_THROW(HDD_ERR_WRONG_INPUT, "unknown element kind " << kind);
    if (!xy_c || !cv_c || !xy_f || !cv_f || !father) HDD_THROW(HDD_ERR_WRONG_INPUT, "NULL argument");
    if (n_coarse < 1 || n_fine < 0 || n_coarse > INT32_MAX) HDD_THROW(HDD_ERR_WRONG_INPUT, "bad grid size");
    const int nl = kind == HDD_SIMPLEX2D ? 3 : 4;
    for (int64_t t = 0; t < n_coarse * nl; ++t)
      if (cv_c[t] < 0 || cv_c[t] >= n_coarse_verts) HDD_THROW(HDD_ERR_INDEX_OUT_OF_RANGE, "coarse vertex index out of range");
    for (int64_t t = 0; t < n_fine * nl; ++t)
      if (cv_f[t] < 0 || cv_f[t] >= n_fine_verts) HDD_THROW(HDD_ERR_INDEX_OUT_OF_RANGE, "fine vertex index out of range");
    // bounding boxes
    std::vector<double> bb(size_t(4) * n_coarse);
    double lo[2] = {HUGE_VAL, HUGE_VAL}, hi[2] = {-HUGE_VAL, -HUGE_VAL};
    for (int64_t c = 0; c < n_coarse; ++c) {
      double b[4] = {HUGE_VAL, HUGE_VAL, -HUGE_VAL, -HUGE_VAL};
      for (int i = 0; i < nl; ++i) {
        const double* p = xy_c + 2 * size_t(cv_c[c * nl + i]);
        b[0] = std::min(b[0], p[0]); b[1] = std::min(b[1], p[1]);
        b[2] = std::max(b[2], p[0]); b[3] = std::max(b[3], p[1]);
      }
      for (int d = 0; d < 4; ++d) bb[size_t(4) * c + d] = b[d];
      lo[0] = std::min(lo[0], b[0]); lo[1] = std::min(lo[1], b[1]);
      hi[0] = std::max(hi[0], b[2]); hi[1] = std::max(hi[1], b[3]);
    }
    const double ext[2] = {hi[0] - lo[0], hi[1] - lo[1]};
    if (!(ext[0] > 0.0) || !(ext[1] > 0.0)) HDD_THROW(HDD_ERR_WRONG_INPUT, "degenerate coarse grid");
    // about one cell per bucket, buckets as square as the domain allows
    const double per_area = double(n_coarse) / (ext[0] * ext[1]);
    const int64_t B[2] = {std::max<int64_t>(1, int64_t(std::sqrt(per_area) * ext[0])),
                          std::max<int64_t>(1, int64_t(std::sqrt(per_area) * ext[1]))};
    auto bucket = [&](double v, int d) {
      const int64_t b = int64_t((v - lo[d]) / ext[d] * double(B[d]));
      return std::min(std::max<int64_t>(b, 0), B[d] - 1);
    };
    std::vector<int64_t> start(size_t(B[0] * B[1]) + 1, 0);
    for (int64_t c = 0; c < n_coarse; ++c)
      for (int64_t by = bucket(bb[4 * c + 1], 1); by <= bucket(bb[4 * c + 3], 1); ++by)
        for (int64_t bx = bucket(bb[4 * c], 0); bx <= bucket(bb[4 * c + 2], 0); ++bx) ++start[size_t(by * B[0] + bx) + 1];
    for (size_t t = 1; t < start.size(); ++t) start[t] += start[t - 1];
    std::vector<int32_t> items(size_t(start.back()));
    {  // cells in increasing id inside every bucket
      std::vector<int64_t> fill(start.begin(), start.end() - 1);
      for (int64_t c = 0; c < n_coarse; ++c)
        for (int64_t by = bucket(bb[4 * c + 1], 1); by <= bucket(bb[4 * c + 3], 1); ++by)
          for (int64_t bx = bucket(bb[4 * c], 0); bx <= bucket(bb[4 * c + 2], 0); ++bx)
            items[size_t(fill[size_t(by * B[0] + bx)]++)] = int32_t(c);
    }
    // depth of a point inside a coarse cell: >= 0 inside, scaled to the cell (barycentric / relative box distance)
    auto depth = [&](int64_t c, double x, double y) {
      if (kind == HDD_CUBE2D) {
        const double* b = &bb[size_t(4) * c];
        return std::min(std::min(x - b[0], b[2] - x) / (b[2] - b[0]), std::min(y - b[1], b[3] - y) / (b[3] - b[1]));
      }
      const double* p0 = xy_c + 2 * size_t(cv_c[3 * c]);
      const double* p1 = xy_c + 2 * size_t(cv_c[3 * c + 1]);
      const double* p2 = xy_c + 2 * size_t(cv_c[3 * c + 2]);
      const double det = (p1[0] - p0[0]) * (p2[1] - p0[1]) - (p2[0] - p0[0]) * (p1[1] - p0[1]);
      const double l1 = ((x - p0[0]) * (p2[1] - p0[1]) - (p2[0] - p0[0]) * (y - p0[1])) / det;
      const double l2 = ((p1[0] - p0[0]) * (y - p0[1]) - (x - p0[0]) * (p1[1] - p0[1])) / det;
      return std::min(std::min(l1, l2), 1.0 - l1 - l2);
    };
    std::atomic<int64_t> lost{-1};
    hdd::parallel_for(n_fine, [&](int64_t a, int64_t e) {
      for (int64_t f = a; f < e; ++f) {
        double x = 0.0, y = 0.0;
        for (int i = 0; i < nl; ++i) {
          x += xy_f[2 * size_t(cv_f[f * nl + i])];
          y += xy_f[2 * size_t(cv_f[f * nl + i]) + 1];
        }
        x /= nl;
        y /= nl;
        const size_t b = size_t(bucket(y, 1) * B[0] + bucket(x, 0));
        int32_t best = -1;
        double best_depth = -1e-9;  // tolerate centres on a coarse face up to rounding
        for (int64_t t = start[b]; t < start[b + 1]; ++t) {
          const double d = depth(items[size_t(t)], x, y);
          if (d > best_depth) {
            best_depth = d;
            best = items[size_t(t)];
          }
        }
        father[f] = best;
        if (best < 0) lost.store(f);
      }
    });
    if (lost.load() >= 0)
      HDD_THROW(HDD_ERR_WRONG_INPUT, "the centre of fine cell " << lost.load() << " lies in no coarse cell: the grids do not cover "
                                                                   "the same domain");
  });
}

}  // extern "C"
