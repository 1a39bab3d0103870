// Host-side grid providers: stand-ins for Stuff::Grid::Providers::Cube<SGrid<2,2>> and for the
// ALUGrid<2,2,simplex,conforming> refinement ladder of the test cases (testcases/ESV2007.hh:50-59,123-134,
// testcases/base.hh:92-103), including the [px py 1] subdomain partition of grid::Multiscale
// (testcases/ESV2007.hh:150-163).  Closed form, O(n), no recursion; tests/ compares the simplex generator
// against an actual recursive longest-edge bisection.
#include <algorithm>
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

}  // extern "C"
