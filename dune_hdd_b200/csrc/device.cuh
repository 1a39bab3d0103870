// Device-side types and element helpers shared by the kernels (sm_100a).
#pragma once
#include <cstdint>

#include "expr.hpp"
#include "quadrature.hpp"

namespace hdd {

constexpr int kMaxParts = 8;

// A data function localised for the device (see hdd_function): constant, per local cell, or postfix program.
struct DevFn {
  int kind;
  int order;
  double value;
  const double* cell;  // indexed by LOCAL cell id
  Program prog;
  int separable;       // expression == px(x[0]) * py(x[1])
  Program px, py;
  FastFn fast;         // the expression as a sum of products of elementary functions of affine arguments, if it fits
};

// sum_k theta[k] * fn[k]: a frozen affinely decomposed function (problem.with_mu(mu), estimators/swipdg.hh:134).
struct DevCombo {
  int n;
  int order;
  double theta[kMaxParts];
  int idx[kMaxParts];  // indices into the handle's function table
};

// Local mesh of one rank: cells [0, n_loc) sorted by global id = [lower halo | owned | upper halo].
struct MeshView {
  int kind;
  int nl;                   // local DoFs per cell: P1 3, Q1 4, P2 6, Q2 9
  int nf;                   // faces per cell
  int32_t n_loc;
  int32_t own0;
  int32_t n_own;
  const double* cgeo;       // simplex: x0,y0,x1,y1,x2,y2 per cell; cube: x0,y0,x1,y1 (vertices 0 and 3)
  const int32_t* neigh;     // [n_own*nf] local neighbour ids, -1 = domain boundary
  const uint8_t* btype;     // [n_own*nf] 1 Dirichlet, 2 Neumann
  const double* tensor;     // [n_loc*4] or nullptr (identity)
  const int64_t* blk_start; // [n_own+1] number of matrix blocks in front of owned cell k
  const int32_t* cgid;      // [n_loc] global cell id
  // logically structured (tensor-product) cube grid, nullptr / 0 otherwise: vertex 0 of every local cell and the
  // one-dimensional geometry tables {x0, hx, 1/hx, -} per column [tnx] followed by {y0, hy, 1/hy, -} per row [tny]
  const int32_t* cell_v0;
  const double* tgeo;
  int tnx, tny;
};

__device__ __forceinline__ double fn_eval(const DevFn& f, int cell, double x, double y) {
  if (f.kind == HDD_FN_CONSTANT) return f.value;
  if (f.kind == HDD_FN_CELLWISE) return __ldg(f.cell + cell);
  if (f.fast.n_terms > 0) return eval_fast(f.fast, x, y);
  const double v[2] = {x, y};
  return eval_program(f.prog, v);
}

__device__ __forceinline__ double combo_eval(const DevCombo& c, const DevFn* table, int cell, double x, double y) {
  double s = 0.0;
  for (int k = 0; k < c.n; ++k) s += c.theta[k] * fn_eval(table[c.idx[k]], cell, x, y);
  return s;
}

__device__ __forceinline__ void load_tensor(const double* tensor, int cell, double K[4]) {
  if (tensor) {
    const double2 a = __ldg(reinterpret_cast<const double2*>(tensor + 4 * size_t(cell)));
    const double2 b = __ldg(reinterpret_cast<const double2*>(tensor + 4 * size_t(cell)) + 1);
    K[0] = a.x; K[1] = a.y; K[2] = b.x; K[3] = b.y;
  } else {
    K[0] = 1.0; K[1] = 0.0; K[2] = 0.0; K[3] = 1.0;
  }
}

// ---- elements ------------------------------------------------------------------------------------------
template <int KIND>
struct Geo;

// P1 on a triangle, Dune numbering: vertices (0,0),(1,0),(0,1); faces {0,1},{0,2},{1,2}.
template <>
struct Geo<HDD_SIMPLEX2D> {
  static constexpr int NL = 3, NF = 3, NGEO = 6;
  double vx[3], vy[3];
  double i00, i01, i10, i11, detj;

  __device__ __forceinline__ void load(const double* cgeo, int cell) {
    const double2* p = reinterpret_cast<const double2*>(cgeo + size_t(NGEO) * cell);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const double2 v = __ldg(p + i);
      vx[i] = v.x;
      vy[i] = v.y;
    }
    const double j00 = vx[1] - vx[0], j10 = vy[1] - vy[0], j01 = vx[2] - vx[0], j11 = vy[2] - vy[0];
    const double det = j00 * j11 - j01 * j10;
    detj = fabs(det);
    const double idet = 1.0 / det;
    i00 = j11 * idet; i01 = -j01 * idet; i10 = -j10 * idet; i11 = j00 * idet;
  }
  __device__ __forceinline__ void to_global(double xi, double eta, double& x, double& y) const {
    x = vx[0] + (vx[1] - vx[0]) * xi + (vx[2] - vx[0]) * eta;
    y = vy[0] + (vy[1] - vy[0]) * xi + (vy[2] - vy[0]) * eta;
  }
  __device__ __forceinline__ void to_local(double x, double y, double& xi, double& eta) const {
    const double dx = x - vx[0], dy = y - vy[0];
    xi = i00 * dx + i01 * dy;
    eta = i10 * dx + i11 * dy;
  }
  __device__ __forceinline__ void basis(double xi, double eta, double* phi, double* gx, double* gy) const {
    phi[0] = 1.0 - xi - eta; phi[1] = xi; phi[2] = eta;
    gx[0] = -i00 - i10; gy[0] = -i01 - i11;
    gx[1] = i00;        gy[1] = i01;
    gx[2] = i10;        gy[2] = i11;
  }
  __device__ __forceinline__ void centroid(double& cx, double& cy) const {
    cx = (vx[0] + vx[1] + vx[2]) / 3.0;
    cy = (vy[0] + vy[1] + vy[2]) / 3.0;
  }
  // faces {0,1}, {0,2}, {1,2}; written with selects so that a face index known only at run time does not turn vx / vy
  // into dynamically indexed (local-memory) arrays
  __device__ __forceinline__ void face_ends(int f, double& ax, double& ay, double& bx, double& by) const {
    ax = f == 2 ? vx[1] : vx[0];
    ay = f == 2 ? vy[1] : vy[0];
    bx = f == 0 ? vx[1] : vx[2];
    by = f == 0 ? vy[1] : vy[2];
  }
  __device__ __forceinline__ double diameter() const {
    double h = 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = i + 1; j < 3; ++j) h = fmax(h, hypot(vx[i] - vx[j], vy[i] - vy[j]));
    return h;
  }
};

// Q1 on an axis-parallel rectangle, Dune numbering: vertices (0,0),(1,0),(0,1),(1,1); faces {0,2},{1,3},{0,1},{2,3}.
template <>
struct Geo<HDD_CUBE2D> {
  static constexpr int NL = 4, NF = 4, NGEO = 4;
  double x0, y0, x1, y1, hx, hy, ihx, ihy, detj;

  __device__ __forceinline__ void load(const double* cgeo, int cell) {
    const double2* p = reinterpret_cast<const double2*>(cgeo + size_t(NGEO) * cell);
    const double2 a = __ldg(p), b = __ldg(p + 1);
    x0 = a.x; y0 = a.y; x1 = b.x; y1 = b.y;
    hx = x1 - x0; hy = y1 - y0;
    ihx = 1.0 / hx; ihy = 1.0 / hy;
    detj = fabs(hx * hy);
  }
  __device__ __forceinline__ void to_global(double xi, double eta, double& x, double& y) const {
    x = x0 + hx * xi;
    y = y0 + hy * eta;
  }
  __device__ __forceinline__ void to_local(double x, double y, double& xi, double& eta) const {
    xi = (x - x0) * ihx;
    eta = (y - y0) * ihy;
  }
  __device__ __forceinline__ void basis(double xi, double eta, double* phi, double* gx, double* gy) const {
    phi[0] = (1.0 - xi) * (1.0 - eta); phi[1] = xi * (1.0 - eta);
    phi[2] = (1.0 - xi) * eta;         phi[3] = xi * eta;
    gx[0] = -(1.0 - eta) * ihx; gx[1] = (1.0 - eta) * ihx; gx[2] = -eta * ihx; gx[3] = eta * ihx;
    gy[0] = -(1.0 - xi) * ihy;  gy[1] = -xi * ihy;         gy[2] = (1.0 - xi) * ihy; gy[3] = xi * ihy;
  }
  __device__ __forceinline__ void centroid(double& cx, double& cy) const {
    cx = 0.5 * (x0 + x1);
    cy = 0.5 * (y0 + y1);
  }
  __device__ __forceinline__ void face_ends(int f, double& ax, double& ay, double& bx, double& by) const {
    // {0,2} left, {1,3} right, {0,1} bottom, {2,3} top
    ax = (f == 1) ? x1 : x0;
    ay = (f == 3) ? y1 : y0;
    bx = (f == 0) ? x0 : x1;
    by = (f == 2) ? y0 : y1;
  }
  __device__ __forceinline__ double diameter() const { return hypot(hx, hy); }
};

// Geometry + nodal Lagrange basis of order P.  P = 1 is Geo<KIND> itself (nodes = vertices in Dune order).  P = 2: P2 / Q2
// with the nodes in lexicographic order - (0,0),(1/2,0),(1,0),(0,1/2),(1/2,1/2),(0,1) on the triangle, i + 3 j on the
// square - which is what the generic Lagrange point sets of the reference's space backend produce (dune-fem,
// discretizations/swipdg.hh:67-71); no reference test instantiates polOrder = 2, so this numbering is unpinned.
template <int KIND, int P>
struct Elem : Geo<KIND> {};

template <>
struct Elem<HDD_SIMPLEX2D, 2> : Geo<HDD_SIMPLEX2D> {
  static constexpr int NL = 6;
  __device__ __forceinline__ void basis(double xi, double eta, double* phi, double* gx, double* gy) const {
    const double l0 = 1.0 - xi - eta, l1 = xi, l2 = eta;
    // physical gradients of the barycentric coordinates
    const double g1x = i00, g1y = i01, g2x = i10, g2y = i11, g0x = -i00 - i10, g0y = -i01 - i11;
    phi[0] = l0 * (2.0 * l0 - 1.0); phi[1] = 4.0 * l0 * l1; phi[2] = l1 * (2.0 * l1 - 1.0);
    phi[3] = 4.0 * l0 * l2;         phi[4] = 4.0 * l1 * l2; phi[5] = l2 * (2.0 * l2 - 1.0);
    const double d0 = 4.0 * l0 - 1.0, d1 = 4.0 * l1 - 1.0, d2 = 4.0 * l2 - 1.0;
    gx[0] = d0 * g0x;                     gy[0] = d0 * g0y;
    gx[1] = 4.0 * (l1 * g0x + l0 * g1x);  gy[1] = 4.0 * (l1 * g0y + l0 * g1y);
    gx[2] = d1 * g1x;                     gy[2] = d1 * g1y;
    gx[3] = 4.0 * (l2 * g0x + l0 * g2x);  gy[3] = 4.0 * (l2 * g0y + l0 * g2y);
    gx[4] = 4.0 * (l2 * g1x + l1 * g2x);  gy[4] = 4.0 * (l2 * g1y + l1 * g2y);
    gx[5] = d2 * g2x;                     gy[5] = d2 * g2y;
  }
};

template <>
struct Elem<HDD_CUBE2D, 2> : Geo<HDD_CUBE2D> {
  static constexpr int NL = 9;
  __device__ __forceinline__ static void lagrange3(double t, double* l, double* d) {
    l[0] = (1.0 - t) * (1.0 - 2.0 * t); l[1] = 4.0 * t * (1.0 - t); l[2] = t * (2.0 * t - 1.0);
    d[0] = 4.0 * t - 3.0;               d[1] = 4.0 - 8.0 * t;       d[2] = 4.0 * t - 1.0;
  }
  __device__ __forceinline__ void basis(double xi, double eta, double* phi, double* gx, double* gy) const {
    double lx[3], dx[3], ly[3], dy[3];
    lagrange3(xi, lx, dx);
    lagrange3(eta, ly, dy);
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        phi[i + 3 * j] = lx[i] * ly[j];
        gx[i + 3 * j] = dx[i] * ly[j] * ihx;
        gy[i + 3 * j] = lx[i] * dy[j] * ihy;
      }
  }
};

// local DoFs of the nodal Lagrange space of order p
__host__ __device__ inline int n_local_dofs(int kind, int p) {
  return kind == HDD_SIMPLEX2D ? (p + 1) * (p + 2) / 2 : (p + 1) * (p + 1);
}

struct FaceGeo {
  double ax, ay, bx, by, nx, ny, h, ih;
};

// axis-parallel rectangle: exact normals, no sqrt / division
__device__ __forceinline__ FaceGeo make_face(const Geo<HDD_CUBE2D>& g, int f) {
  FaceGeo e;
  g.face_ends(f, e.ax, e.ay, e.bx, e.by);
  const bool vertical = f < 2;  // faces 0,1 are x = const
  e.h = vertical ? fabs(g.hy) : fabs(g.hx);
  e.ih = vertical ? fabs(g.ihy) : fabs(g.ihx);
  e.nx = vertical ? (f == 0 ? -1.0 : 1.0) : 0.0;
  e.ny = vertical ? 0.0 : (f == 2 ? -1.0 : 1.0);
  if (g.hx < 0.0) e.nx = -e.nx;
  if (g.hy < 0.0) e.ny = -e.ny;
  return e;
}

__device__ __forceinline__ FaceGeo make_face(const Geo<HDD_SIMPLEX2D>& g, int f) {
  FaceGeo e;
  g.face_ends(f, e.ax, e.ay, e.bx, e.by);
  const double tx = e.bx - e.ax, ty = e.by - e.ay;
  e.h = sqrt(tx * tx + ty * ty);
  e.ih = 1.0 / e.h;
  e.nx = ty * e.ih;
  e.ny = -tx * e.ih;
  double cx, cy;
  g.centroid(cx, cy);
  if (e.nx * (0.5 * (e.ax + e.bx) - cx) + e.ny * (0.5 * (e.ay + e.by) - cy) < 0.0) {
    e.nx = -e.nx;
    e.ny = -e.ny;
  }
  return e;
}

// upstream GDT::LocalEvaluation::SWIPDG::internal::{inner,boundary}_sigma(polOrder)
__host__ __device__ inline double sigma_inner(int p) { return p <= 1 ? 8.0 : p <= 2 ? 20.0 : p <= 3 ? 38.0 : 62.0; }
__host__ __device__ inline double sigma_boundary(int p) { return p <= 1 ? 14.0 : p <= 2 ? 38.0 : p <= 3 ? 74.0 : 122.0; }

// Sorted block slots of an owned cell: the row block of cell T holds one n_loc x n_loc block per member of
// sort({T} u neighbours) (EllipticSWIPDG::pattern + SparsityPatternDefault ordering).
template <int NF>
__device__ __forceinline__ int block_slot(int self, const int* nb, int target) {
  int s = self < target ? 1 : 0;
#pragma unroll
  for (int f = 0; f < NF; ++f) s += (nb[f] >= 0 && nb[f] < target) ? 1 : 0;
  return s;
}

template <int NF>
__device__ __forceinline__ int block_count(const int* nb) {
  int s = 1;
#pragma unroll
  for (int f = 0; f < NF; ++f) s += nb[f] >= 0 ? 1 : 0;
  return s;
}

}  // namespace hdd
