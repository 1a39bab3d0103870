// "cg.mg": CG preconditioned by a two-level method for the SWIPDG matrix on logically structured Q1 grids.
//
//   M^-1 r = D_blk^-1 r  +  P V(P^T r)  +  P C V_C(C P^T r)
//
// D_blk   the 4 x 4 cell blocks of A (block Jacobi, the DG-level smoother),
// P       the injection of the conforming Q1 space into the DG space (DG DoF (T, i) = value at vertex i of T); the
//         auxiliary operator A_c = P^T A P is a 9-point vertex stencil (for continuous functions every inner-face jump
//         vanishes, what remains is the volume term and the Dirichlet-face terms),
// V       one geometric multigrid V(1,1)-cycle for A_c on the (nx+1) x (ny+1) vertex grid: damped Jacobi, bilinear
//         interpolation, full weighting, Galerkin coarse operators, dense inverse on the coarsest grid,
// C       the checkerboard sign (-1)^(ix+iy).  The reference under-integrates the Q1 volume term with the midpoint rule
//         (SURVEY 0.4), which makes the element hourglass mode energy-free: A_c has a second family of low-energy modes,
//         checkerboard x smooth, invisible to standard coarse grids.  The twisted hierarchy V_C built from C A_c C
//         removes exactly that family (measured: CG iterations 188 -> 42 at 128^2, flat in h).
//
// The reference's default solver is BiCGSTAB with an algebraic-multigrid / ILU preconditioner
// (Stuff::LA::Solver defaults, discretizations/base.hh:314-322); this is the same idea specialised to the
// structured grids of BASELINE configs 2 and 5.  Everything is additive and symmetric, so CG stays applicable; all
// reductions are deterministic.
//
// Multi GPU: the DG level (SpMV, block Jacobi, restriction, prolongation) is distributed like the rest of the solve.
// The vertex-level operators are replicated (one all-reduce per solve makes the level-0 stencil global); when the ranks'
// cells are full-width bands of rows stacked in rank order (the slabs bench.py deals out), the vectors of the two finest
// vertex levels are swept in row strips with one halo row exchanged per sweep (grouped ncclSend/ncclRecv with the rank
// below / above), the level-2 right-hand side is summed over the ranks and levels >= 2 run replicated.  Otherwise the
// whole V-cycle runs replicated on the all-reduced restricted residual.
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <initializer_list>
#include <ctime>

#include "handles.hpp"
#include "reduce.cuh"

namespace hdd {

namespace {

constexpr int kMgThreads = 256;
constexpr int kMaxCoarse = 400;   // dense coarsest solve up to this many vertices
constexpr double kOmega = 0.8;    // Jacobi damping on the vertex levels

inline int blocks_for(int64_t n) { return int((n + kMgThreads - 1) / kMgThreads); }

// The vertices (linear index range = whole grid rows) a kernel launch works on: the whole level, or - for the levels
// that are distributed over the GPUs - the row strip this rank owns.
struct Rows {
  int64_t beg, cnt;
};

__device__ __forceinline__ void load_neigh4(const int32_t* neigh, int k, int* nb) {
  const int4 v = __ldg(reinterpret_cast<const int4*>(neigh) + k);
  nb[0] = v.x; nb[1] = v.y; nb[2] = v.z; nb[3] = v.w;
}

// ---- structure detection -------------------------------------------------------------------------------------
// cells: vertices must be (v0, v0+1, v0+nx1, v0+nx1+1); records v0 per cell and the map lexicographic cell -> cell
__global__ void k_struct_cells(const int32_t* __restrict__ cv, int32_t n_cells, int nx, int ny, int32_t* __restrict__ cell_v0,
                               int32_t* __restrict__ lex_cell, int32_t* flag) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;  // local cell id
  if (c >= n_cells) return;
  const int4 v = __ldg(reinterpret_cast<const int4*>(cv) + c);
  const int nx1 = nx + 1;
  const int cx = v.x % nx1, cy = v.x / nx1;
  if (v.x < 0 || cx >= nx || cy >= ny || v.y != v.x + 1 || v.z != v.x + nx1 || v.w != v.x + nx1 + 1) {
    atomicOr(flag, 1);
    return;
  }
  cell_v0[c] = v.x;
  lex_cell[cx + nx * cy] = c;
}

// vertices: tensor-product coordinates
__global__ void k_struct_verts(const double* __restrict__ xy, int32_t n_verts, int nx, int32_t* flag) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n_verts) return;
  const int nx1 = nx + 1;
  const double2 p = __ldg(reinterpret_cast<const double2*>(xy) + v);
  const double2 px = __ldg(reinterpret_cast<const double2*>(xy) + v % nx1);
  const double2 py = __ldg(reinterpret_cast<const double2*>(xy) + (v / nx1) * nx1);
  if (p.x != px.x || p.y != py.y) atomicOr(flag, 2);
}

// one-dimensional geometry of the tensor grid from the vertex coordinates: {x0, hx, 1/hx, -} per column followed by
// {y0, hy, 1/hy, -} per row (the reciprocals save the assembly kernel six of its ten fp64 divisions per cell)
__global__ void k_struct_geo(const double* __restrict__ xy, int nx, int ny, double* __restrict__ tgeo) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nx + ny) return;
  const int nx1 = nx + 1;
  double4* out = reinterpret_cast<double4*>(tgeo);
  if (t < nx) {
    const double x0 = xy[2 * size_t(t)], x1 = xy[2 * size_t(t + 1)];
    out[t] = make_double4(x0, x1 - x0, 1.0 / (x1 - x0), 0.0);
  } else {
    const int r = t - nx;
    const double y0 = xy[2 * size_t(r) * nx1 + 1], y1 = xy[2 * size_t(r + 1) * nx1 + 1];
    out[t] = make_double4(y0, y1 - y0, 1.0 / (y1 - y0), 0.0);
  }
}

// ---- level-0 operator: A_c = P^T A P gathered per vertex (no atomics) ------------------------------------------
// S[e][v], e = (dy+1)*3 + (dx+1): coupling of vertex v = (ix, iy) to vertex (ix+dx, iy+dy).  Entries of P^T A P at
// index distance 2 cancel analytically (inner-face jumps of continuous functions) and are dropped.
__global__ void __launch_bounds__(kMgThreads)
    k_vertex_galerkin(MeshView m, const double* __restrict__ vals, const int32_t* __restrict__ cell_v0,
                      const int32_t* __restrict__ lex_cell, int nx, int ny, Rows rg, double* __restrict__ S) {
  const int64_t nv = int64_t(nx + 1) * (ny + 1);
  const int64_t tid = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (tid >= rg.cnt) return;
  const int64_t v = rg.beg + tid;
  const int nx1 = nx + 1;
  const int ix = int(v % nx1), iy = int(v / nx1);
  double acc[9];
#pragma unroll
  for (int e = 0; e < 9; ++e) acc[e] = 0.0;
#pragma unroll
  for (int b = 0; b < 2; ++b)
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const int cx = ix - 1 + a, cy = iy - 1 + b;
      if (cx < 0 || cy < 0 || cx >= nx || cy >= ny) continue;
      const int c = __ldg(lex_cell + cx + nx * cy);  // local cell id, -1 = not on this rank
      const int k = c - m.own0;
      if (c < 0 || k < 0 || k >= m.n_own) continue;  // rows of cells owned elsewhere are added by the all-reduce
      const int i = (1 - a) + 2 * (1 - b);  // local index of v in that cell
      int nb[4];
      load_neigh4(m.neigh, k, nb);
      const int nblk = block_count<4>(nb);
      const double* row = vals + __ldg(m.blk_start + k) * 16 + int64_t(i) * nblk * 4;
#pragma unroll
      for (int t = 0; t < 5; ++t) {
        const int cell = t == 0 ? c : nb[t - 1];
        if (cell < 0) continue;
        const int slot = block_slot<4>(c, nb, cell);
        const int v0 = __ldg(cell_v0 + cell);
        const int jx = v0 % nx1 - ix, jy = v0 / nx1 - iy;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int dx = jx + (j & 1), dy = jy + (j >> 1);
          if (dx < -1 || dx > 1 || dy < -1 || dy > 1) continue;
          acc[(dy + 1) * 3 + dx + 1] += __ldg(row + slot * 4 + j);
        }
      }
    }
#pragma unroll
  for (int e = 0; e < 9; ++e) S[e * nv + v] = acc[e];
}

// b1 = C b0 after the all-reduce of b0 (multi GPU)
__global__ void k_twist_vector(const int* done, const double* __restrict__ b0, int nx, Rows rg, double* __restrict__ b1) {
  if (done && *done) return;
  const int64_t tid = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (tid >= rg.cnt) return;
  const int64_t v = rg.beg + tid;
  const int nx1 = nx + 1;
  b1[v] = ((int(v % nx1) + int(v / nx1)) & 1) ? -b0[v] : b0[v];
}

// ---- Galerkin coarse operator with bilinear interpolation: 9-point -> 9-point ---------------------------------
__global__ void __launch_bounds__(kMgThreads)
    k_rap(const double* __restrict__ Sf, int nxf, int nyf, int twist, double* __restrict__ Sc) {
  const int nxc = nxf / 2, nyc = nyf / 2;
  const int64_t nvc = int64_t(nxc + 1) * (nyc + 1), nvf = int64_t(nxf + 1) * (nyf + 1);
  const int64_t I = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (I >= nvc) return;
  const int IX = int(I % (nxc + 1)), IY = int(I / (nxc + 1));
  double acc[9];
#pragma unroll
  for (int e = 0; e < 9; ++e) acc[e] = 0.0;
  for (int dy = -1; dy <= 1; ++dy)
    for (int dx = -1; dx <= 1; ++dx) {
      const int fx = 2 * IX + dx, fy = 2 * IY + dy;
      if (fx < 0 || fy < 0 || fx > nxf || fy > nyf) continue;
      const double wd = (dx == 0 ? 1.0 : 0.5) * (dy == 0 ? 1.0 : 0.5);
      const int64_t i = fx + int64_t(nxf + 1) * fy;
      for (int ey = -1; ey <= 1; ++ey)
        for (int ex = -1; ex <= 1; ++ex) {
          const int jx = fx + ex, jy = fy + ey;
          if (jx < 0 || jy < 0 || jx > nxf || jy > nyf) continue;
          double s = wd * __ldg(Sf + ((ey + 1) * 3 + ex + 1) * nvf + i);
          if (twist && ((ex + ey) & 1)) s = -s;  // fine operator C A C, read from the arrays of A
          // coarse vertices J = I + D whose interpolation stencil touches j: |j - 2J| <= 1
#pragma unroll
          for (int Dy = -1; Dy <= 1; ++Dy) {
            const int py = dy + ey - 2 * Dy;
            if (py < -1 || py > 1 || IY + Dy < 0 || IY + Dy > nyc) continue;
#pragma unroll
            for (int Dx = -1; Dx <= 1; ++Dx) {
              const int px = dx + ex - 2 * Dx;
              if (px < -1 || px > 1 || IX + Dx < 0 || IX + Dx > nxc) continue;
              acc[(Dy + 1) * 3 + Dx + 1] += s * (px == 0 ? 1.0 : 0.5) * (py == 0 ? 1.0 : 0.5);
            }
          }
        }
    }
#pragma unroll
  for (int e = 0; e < 9; ++e) Sc[e * nvc + I] = acc[e];
}

// ---- V-cycle kernels.  done: convergence latch of the CG iteration that owns this application (see k_cg_*) ----------
// pre-smoothing from a zero guess, x = w D^-1 b, fused with the residual r = b - A x
__global__ void __launch_bounds__(kMgThreads)
    k_mg_pre(const int* done, const double* __restrict__ S, const double* __restrict__ dinv, int nx, int ny, Rows rg,
             const double* __restrict__ b, double* __restrict__ x, double* __restrict__ r) {
  if (done && *done) return;
  const int nx1 = nx + 1;
  const int64_t nv = int64_t(nx1) * (ny + 1);
  const int64_t tid = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (tid >= rg.cnt) return;
  const int64_t i = rg.beg + tid;
  const int ix = int(i % nx1), iy = int(i / nx1);
  double ax = 0.0;
#pragma unroll
  for (int ey = -1; ey <= 1; ++ey)
#pragma unroll
    for (int ex = -1; ex <= 1; ++ex) {
      const int jx = ix + ex, jy = iy + ey;
      if (jx < 0 || jy < 0 || jx > nx || jy > ny) continue;
      const int64_t j = i + ex + int64_t(ey) * nx1;
      const double xj = __ldg(dinv + j) * __ldg(b + j);  // dinv = w / diagonal: no division in the sweep
      ax = fma(__ldg(S + ((ey + 1) * 3 + ex + 1) * nv + i), xj, ax);
      if (ex == 0 && ey == 0) x[i] = xj;
    }
  r[i] = b[i] - ax;
}

// full weighting: b_H = P^T r
__global__ void __launch_bounds__(kMgThreads)
    k_mg_restrict(const int* done, const double* __restrict__ r, int nxf, int nyf, Rows rg /* coarse */, double* __restrict__ bc) {
  if (done && *done) return;
  const int nxc = nxf / 2;
  const int64_t tid = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (tid >= rg.cnt) return;
  const int64_t I = rg.beg + tid;
  const int IX = int(I % (nxc + 1)), IY = int(I / (nxc + 1));
  double s = 0.0;
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) {
      const int fx = 2 * IX + dx, fy = 2 * IY + dy;
      if (fx < 0 || fy < 0 || fx > nxf || fy > nyf) continue;
      s = fma((dx == 0 ? 1.0 : 0.5) * (dy == 0 ? 1.0 : 0.5), __ldg(r + fx + int64_t(nxf + 1) * fy), s);
    }
  bc[I] = s;
}

// x += P x_H (bilinear interpolation)
__global__ void __launch_bounds__(kMgThreads)
    k_mg_prolong_add(const int* done, const double* __restrict__ xc, int nxf, Rows rg /* fine */, double* __restrict__ x) {
  if (done && *done) return;
  const int nxc = nxf / 2;
  const int64_t tid = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (tid >= rg.cnt) return;
  const int64_t i = rg.beg + tid;
  const int fx = int(i % (nxf + 1)), fy = int(i / (nxf + 1));
  const int cx = fx >> 1, cy = fy >> 1;
  const int ox = fx & 1, oy = fy & 1;
  const int64_t I = cx + int64_t(nxc + 1) * cy;
  double v = __ldg(xc + I);
  if (ox) v += __ldg(xc + I + 1);
  if (oy) {
    v += __ldg(xc + I + nxc + 1);
    if (ox) v += __ldg(xc + I + nxc + 2);
  }
  x[i] += v * (ox ? 0.5 : 1.0) * (oy ? 0.5 : 1.0);
}

// post-smoothing: y = x + w D^-1 (b - A x)
__global__ void __launch_bounds__(kMgThreads)
    k_mg_post(const int* done, const double* __restrict__ S, const double* __restrict__ dinv, int nx, int ny, Rows rg,
              const double* __restrict__ b, const double* __restrict__ x, double* __restrict__ y) {
  if (done && *done) return;
  const int nx1 = nx + 1;
  const int64_t nv = int64_t(nx1) * (ny + 1);
  const int64_t tid = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (tid >= rg.cnt) return;
  const int64_t i = rg.beg + tid;
  const int ix = int(i % nx1), iy = int(i / nx1);
  double ax = 0.0;
#pragma unroll
  for (int ey = -1; ey <= 1; ++ey)
#pragma unroll
    for (int ex = -1; ex <= 1; ++ex) {
      const int jx = ix + ex, jy = iy + ey;
      if (jx < 0 || jy < 0 || jx > nx || jy > ny) continue;
      ax = fma(__ldg(S + ((ey + 1) * 3 + ex + 1) * nv + i), __ldg(x + i + ex + int64_t(ey) * nx1), ax);
    }
  y[i] = fma(__ldg(dinv + i), b[i] - ax, x[i]);
}

// Level 0, both hierarchies in one sweep: C A_c C differs from A_c only by the sign of the edge-neighbour entries
// (e odd), so the nine coefficient arrays - three quarters of the bytes of a smoothing sweep - are read once for
// the plain and the twisted hierarchy.
__global__ void __launch_bounds__(kMgThreads)
    k_mg_pre2(const int* done, const double* __restrict__ S, const double* __restrict__ dinv, int nx, int ny, Rows rg,
              const double* __restrict__ b0, const double* __restrict__ b1, double* __restrict__ x0, double* __restrict__ x1,
              double* __restrict__ r0, double* __restrict__ r1) {
  if (done && *done) return;
  const int nx1 = nx + 1;
  const int64_t nv = int64_t(nx1) * (ny + 1);
  const int64_t tid = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (tid >= rg.cnt) return;
  const int64_t i = rg.beg + tid;
  const int ix = int(i % nx1), iy = int(i / nx1);
  double a0 = 0.0, a1 = 0.0;
#pragma unroll
  for (int ey = -1; ey <= 1; ++ey)
#pragma unroll
    for (int ex = -1; ex <= 1; ++ex) {
      const int jx = ix + ex, jy = iy + ey;
      if (jx < 0 || jy < 0 || jx > nx || jy > ny) continue;
      const int64_t j = i + ex + int64_t(ey) * nx1;
      const double dj = __ldg(dinv + j);
      const double xj0 = dj * __ldg(b0 + j), xj1 = dj * __ldg(b1 + j);
      const double se = __ldg(S + ((ey + 1) * 3 + ex + 1) * nv + i);
      a0 = fma(se, xj0, a0);
      a1 = fma(((ex + ey) & 1) ? -se : se, xj1, a1);
      if (ex == 0 && ey == 0) { x0[i] = xj0; x1[i] = xj1; }
    }
  r0[i] = b0[i] - a0;
  r1[i] = b1[i] - a1;
}

__global__ void __launch_bounds__(kMgThreads)
    k_mg_post2(const int* done, const double* __restrict__ S, const double* __restrict__ dinv, int nx, int ny, Rows rg,
               const double* __restrict__ b0, const double* __restrict__ b1, const double* __restrict__ x0,
               const double* __restrict__ x1, double* __restrict__ y0, double* __restrict__ y1) {
  if (done && *done) return;
  const int nx1 = nx + 1;
  const int64_t nv = int64_t(nx1) * (ny + 1);
  const int64_t tid = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (tid >= rg.cnt) return;
  const int64_t i = rg.beg + tid;
  const int ix = int(i % nx1), iy = int(i / nx1);
  double a0 = 0.0, a1 = 0.0;
#pragma unroll
  for (int ey = -1; ey <= 1; ++ey)
#pragma unroll
    for (int ex = -1; ex <= 1; ++ex) {
      const int jx = ix + ex, jy = iy + ey;
      if (jx < 0 || jy < 0 || jx > nx || jy > ny) continue;
      const int64_t j = i + ex + int64_t(ey) * nx1;
      const double se = __ldg(S + ((ey + 1) * 3 + ex + 1) * nv + i);
      a0 = fma(se, __ldg(x0 + j), a0);
      a1 = fma(((ex + ey) & 1) ? -se : se, __ldg(x1 + j), a1);
    }
  const double d = __ldg(dinv + i);
  y0[i] = fma(d, b0[i] - a0, x0[i]);
  y1[i] = fma(d, b1[i] - a1, x1[i]);
}

// dinv = w / diagonal of the level operator
__global__ void k_mg_dinv(const double* __restrict__ S, int64_t nv, double* __restrict__ dinv) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < nv) dinv[i] = kOmega / S[4 * nv + i];
}

// coarsest grid: x = A^-1 b with the dense inverse, one thread per row
__global__ void k_mg_dense(const int* done, const double* __restrict__ Ainv, int n, const double* __restrict__ b,
                           double* __restrict__ x) {
  if (done && *done) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = 0.0;
  for (int j = 0; j < n; ++j) s = fma(__ldg(Ainv + size_t(i) * n + j), __ldg(b + j), s);
  x[i] = s;
}

// ---- DG level ---------------------------------------------------------------------------------------------------
// rc = P^T r (sum of the DG residual entries sitting on each vertex) and its checkerboard-signed copy
__global__ void __launch_bounds__(kMgThreads)
    k_dg_restrict(const int* done, const double* __restrict__ r, const int32_t* __restrict__ lex_cell, int own0,
                  int n_own, int nx, int ny, Rows rg, double* __restrict__ b0, double* __restrict__ b1) {
  if (done && *done) return;
  const int nx1 = nx + 1;
  const int64_t tid = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (tid >= rg.cnt) return;
  const int64_t v = rg.beg + tid;
  const int ix = int(v % nx1), iy = int(v / nx1);
  double s = 0.0;
#pragma unroll
  for (int b = 0; b < 2; ++b)
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const int cx = ix - 1 + a, cy = iy - 1 + b;
      if (cx < 0 || cy < 0 || cx >= nx || cy >= ny) continue;
      const int k = __ldg(lex_cell + cx + nx * cy) - own0;
      if (k < 0 || k >= n_own) continue;
      s += __ldg(r + size_t(4) * k + (1 - a) + 2 * (1 - b));
    }
  b0[v] = s;
  if (b1) b1[v] = ((ix + iy) & 1) ? -s : s;
}

// dst += src over one grid row (the neighbour's share of a vertex row both ranks contribute to); min / max cell row
__global__ void k_add_row(const int* done, double* __restrict__ dst, const double* __restrict__ src, int n) {
  if (done && *done) return;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) dst[t] += src[t];
}

__global__ void k_cell_row_range(const int32_t* __restrict__ cell_v0_owned, int32_t n_own, int nx, int* __restrict__ minmax) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_own) return;
  const int cy = __ldg(cell_v0_owned + k) / (nx + 1);
  atomicMin(minmax, cy);
  atomicMax(minmax + 1, cy);
}

// z += P (x0 + C x1), r.z recomputed; optionally p = z (first direction).  One thread per cell, 256-bit accesses.
__global__ void __launch_bounds__(kMgThreads)
    k_dg_prolong_dot(const int* done, int32_t n_cells, const int32_t* __restrict__ cell_v0 /* of the owned cells */, int nx,
                     const double* __restrict__ x0, const double* __restrict__ x1, const double* __restrict__ r,
                     double* __restrict__ z, double* __restrict__ p_init, double* partial, CgScalars* sc) {
  if (done && *done) return;
  const int nx1 = nx + 1;
  double v[1] = {0.0};
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  for (int64_t c = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; c < n_cells; c += stride) {
    const int v0 = __ldg(cell_v0 + c);
    const double sg = ((v0 % nx1 + v0 / nx1) & 1) ? -1.0 : 1.0;  // parity of vertex 0; vertices 1, 2 flip, 3 agrees
    const double4 zz = *reinterpret_cast<const double4*>(z + 4 * c);
    const double4 rr = *reinterpret_cast<const double4*>(r + 4 * c);
    double4 o;
    o.x = zz.x + __ldg(x0 + v0) + sg * __ldg(x1 + v0);
    o.y = zz.y + __ldg(x0 + v0 + 1) - sg * __ldg(x1 + v0 + 1);
    o.z = zz.z + __ldg(x0 + v0 + nx1) - sg * __ldg(x1 + v0 + nx1);
    o.w = zz.w + __ldg(x0 + v0 + nx1 + 1) + sg * __ldg(x1 + v0 + nx1 + 1);
    *reinterpret_cast<double4*>(z + 4 * c) = o;
    if (p_init) *reinterpret_cast<double4*>(p_init + 4 * c) = o;
    v[0] = fma(rr.x, o.x, fma(rr.y, o.y, fma(rr.z, o.z, fma(rr.w, o.w, v[0]))));
  }
  grid_sum<1>(v, partial, &sc->ticket_a, [sc](const double(&w)[1]) { sc->red[1] = w[0]; });
}

}  // namespace

// ---- host side ------------------------------------------------------------------------------------------------
struct MgLevel {
  int nx = 0, ny = 0;
  int64_t nv = 0;
  DevBuf<double> S, dinv, b, x, r, y;
};

struct MgHierarchy {
  std::vector<std::unique_ptr<MgLevel>> levels;
  DevBuf<double> coarse_inv;
};

// Multi GPU: the two finest vertex levels are swept in row strips.  The operators (S, dinv) stay replicated - they
// are static during a solve and one all-reduce per solve makes them global - only the vectors b, x, r of levels 0 and 1
// are computed strip-wise, on full-size arrays, so a halo row lives at the same index on every rank and an exchange is
// "send my first / last owned row, receive the neighbour's into the row next to my strip".  Level 2 and below run
// replicated on the all-reduced level-2 right-hand side.
struct MgDist {
  bool on = false;
  int lower = -1, upper = -1;  // ranks owning the strips below / above (-1: none)
  int c0 = 0, c1 = 0;          // owned cell rows [c0, c1) of level 0
  bool last = false;           // the top strip also owns the last vertex row
  DevBuf<double> tmp;          // one level-0 row
};

struct MgState {
  MgHierarchy h[2];  // plain and checkerboard-twisted
  int nx = 0, ny = 0;
  MgDist dist;
};

static Rows level_rows(const MgState& st, const MgLevel& L, int l) {
  if (!st.dist.on || l >= 2) return Rows{0, L.nv};
  const int v0 = st.dist.c0 >> l, v1 = (st.dist.c1 >> l) + (st.dist.last ? 1 : 0);
  return Rows{int64_t(v0) * (L.nx + 1), int64_t(v1 - v0) * (L.nx + 1)};
}

enum { EX_UP = 1, EX_DOWN = 2 };
// EX_UP: my last owned row goes to the rank above, the last row of the rank below arrives in the row under my strip.
// EX_DOWN: my first owned row goes to the rank below, the first row of the rank above arrives in the row over my strip.
static void exchange_rows(hdd_mesh* m, const MgState& st, const MgLevel& L, int l, std::initializer_list<double*> arrays, int dirs) {
  const MgDist& d = st.dist;
  if (!d.on || (d.lower < 0 && d.upper < 0)) return;
  const int nx1 = L.nx + 1;
  const int64_t v0 = d.c0 >> l, v1 = (d.c1 >> l) + (d.last ? 1 : 0);
  Nccl& nc = Nccl::get();
  nc.group_start();
  for (double* a : arrays) {
    if (dirs & EX_UP) {
      if (d.upper >= 0) nc.send(a + (v1 - 1) * nx1, size_t(nx1), d.upper, m->comm, m->stream);
      if (d.lower >= 0) nc.recv(a + (v0 - 1) * nx1, size_t(nx1), d.lower, m->comm, m->stream);
    }
    if (dirs & EX_DOWN) {
      if (d.lower >= 0) nc.send(a + v0 * nx1, size_t(nx1), d.lower, m->comm, m->stream);
      if (d.upper >= 0) nc.recv(a + v1 * nx1, size_t(nx1), d.upper, m->comm, m->stream);
    }
  }
  nc.group_end();
}

void mg_detect_structure(hdd_mesh* m, const double* xy_host, const double* xy_dev, const int32_t* cv_dev, int64_t n_verts) {
  m->sx = m->sy = 0;
  if (m->kind != HDD_CUBE2D) return;
  int64_t nx1 = 1;
  while (nx1 < n_verts && xy_host[2 * nx1 + 1] == xy_host[1]) ++nx1;
  if (nx1 < 2 || n_verts % nx1 != 0) return;
  const int64_t nx = nx1 - 1, ny = n_verts / nx1 - 1;
  if (ny < 1 || nx * ny != m->n_global) return;
  cudaStream_t s = m->stream;
  m->cell_v0.alloc(size_t(m->n_loc));
  m->lex_cell.alloc(size_t(m->n_global));
  HDD_CUDA(cudaMemsetAsync(m->lex_cell.p, 0xFF, size_t(m->n_global) * sizeof(int32_t), s));  // -1: not on this rank
  DevBuf<int32_t> flag;
  flag.alloc(1);
  flag.zero(s);
  k_struct_cells<<<blocks_for(m->n_loc), kMgThreads, 0, s>>>(cv_dev, m->n_loc, int(nx), int(ny), m->cell_v0.p, m->lex_cell.p, flag.p);
  k_struct_verts<<<blocks_for(n_verts), kMgThreads, 0, s>>>(xy_dev, int32_t(n_verts), int(nx), flag.p);
  m->tgeo.alloc(4 * size_t(nx + ny));
  k_struct_geo<<<blocks_for(nx + ny), kMgThreads, 0, s>>>(xy_dev, int(nx), int(ny), m->tgeo.p);
  count_launch(3);
  int32_t f = 0;
  HDD_CUDA(cudaMemcpyAsync(&f, flag.p, sizeof(f), cudaMemcpyDeviceToHost, s));
  HDD_CUDA(cudaStreamSynchronize(s));
  if (f != 0) {
    m->cell_v0.release();
    m->lex_cell.release();
    m->tgeo.release();
    return;
  }
  m->sx = int(nx);
  m->sy = int(ny);
}

// recomputes the Galerkin operators below the finest level and the dense coarsest inverse; the level buffers are
// allocated once per mesh by mg_setup and persist across solves, so a captured CUDA graph of the iteration stays valid.
// The twisted hierarchy has no level-0 arrays of its own: its finest operator is C A_c C, read from `S0` with the sign
// of the edge-neighbour entries flipped.
static void build_hierarchy(hdd_swipdg* h, MgHierarchy& H, const double* S0, bool twisted) {
  cudaStream_t s = h->mesh->stream;
  for (size_t l = 0; l + 1 < H.levels.size(); ++l) {
    MgLevel& f = *H.levels[l];
    MgLevel& c = *H.levels[l + 1];
    k_rap<<<blocks_for(c.nv), kMgThreads, 0, s>>>(l == 0 ? S0 : f.S.p, f.nx, f.ny, (l == 0 && twisted) ? 1 : 0, c.S.p);
    count_launch();
  }
  for (size_t l = twisted ? 1 : 0; l < H.levels.size(); ++l) {
    MgLevel& L = *H.levels[l];
    k_mg_dinv<<<blocks_for(L.nv), kMgThreads, 0, s>>>(L.S.p, L.nv, L.dinv.p);
    count_launch();
  }
  // dense inverse of the coarsest operator (s.p.d.), Gauss-Jordan on the host
  MgLevel& C = *H.levels.back();
  if (C.nv > kMaxCoarse)
    HDD_THROW(HDD_ERR_REQUIREMENTS_NOT_MET, "cg.mg: the " << H.levels.front()->nx << " x " << H.levels.front()->ny
                                                           << " grid cannot be coarsened by halving down to at most "
                                                           << kMaxCoarse << " vertices (stuck at " << C.nx << " x " << C.ny << ")");
  const int n = int(C.nv);
  const bool coarsest_is_finest = H.levels.size() == 1;
  std::vector<double> S(size_t(9) * n);
  HDD_CUDA(cudaMemcpyAsync(S.data(), coarsest_is_finest ? S0 : C.S.p, S.size() * sizeof(double), cudaMemcpyDeviceToHost, s));
  HDD_CUDA(cudaStreamSynchronize(s));
  if (coarsest_is_finest && twisted)
    for (int e = 1; e < 9; e += 2)
      for (int i = 0; i < n; ++i) S[size_t(e) * n + i] = -S[size_t(e) * n + i];
  std::vector<double> A(size_t(n) * n, 0.0), I(size_t(n) * n, 0.0);
  const int nx1 = C.nx + 1;
  for (int i = 0; i < n; ++i) {
    const int ix = i % nx1, iy = i / nx1;
    for (int ey = -1; ey <= 1; ++ey)
      for (int ex = -1; ex <= 1; ++ex) {
        const int jx = ix + ex, jy = iy + ey;
        if (jx < 0 || jy < 0 || jx > C.nx || jy > C.ny) continue;
        A[size_t(i) * n + (jx + nx1 * jy)] = S[size_t((ey + 1) * 3 + ex + 1) * n + i];
      }
    I[size_t(i) * n + i] = 1.0;
  }
  for (int c = 0; c < n; ++c) {
    const double piv = A[size_t(c) * n + c];
    if (!(piv > 0.0)) HDD_THROW(HDD_ERR_REQUIREMENTS_NOT_MET, "cg.mg: the coarsest operator is not positive definite");
    const double ip = 1.0 / piv;
    for (int j = 0; j < n; ++j) { A[size_t(c) * n + j] *= ip; I[size_t(c) * n + j] *= ip; }
    for (int i = 0; i < n; ++i) {
      if (i == c) continue;
      const double f = A[size_t(i) * n + c];
      if (f == 0.0) continue;
      for (int j = 0; j < n; ++j) { A[size_t(i) * n + j] -= f * A[size_t(c) * n + j]; I[size_t(i) * n + j] -= f * I[size_t(c) * n + j]; }
    }
  }
  H.coarse_inv.upload(I.data(), I.size(), s);
  HDD_CUDA(cudaStreamSynchronize(s));
}

// levels l0 .. coarsest of one hierarchy, all of them replicated: down, dense solve, up; result in levels[l0]->x
static void vcycle_from(MgHierarchy& H, int l0, const int* done, cudaStream_t s) {
  const int nl = int(H.levels.size());
  for (int l = l0; l + 1 < nl; ++l) {
    MgLevel& L = *H.levels[size_t(l)];
    MgLevel& Cn = *H.levels[size_t(l) + 1];
    k_mg_pre<<<blocks_for(L.nv), kMgThreads, 0, s>>>(done, L.S.p, L.dinv.p, L.nx, L.ny, Rows{0, L.nv}, L.b.p, L.x.p, L.r.p);
    k_mg_restrict<<<blocks_for(Cn.nv), kMgThreads, 0, s>>>(done, L.r.p, L.nx, L.ny, Rows{0, Cn.nv}, Cn.b.p);
    count_launch(2);
  }
  MgLevel& C = *H.levels.back();
  k_mg_dense<<<(int(C.nv) + 127) / 128, 128, 0, s>>>(done, H.coarse_inv.p, int(C.nv), C.b.p, C.x.p);
  count_launch();
  for (int l = nl - 2; l >= l0; --l) {
    MgLevel& L = *H.levels[size_t(l)];
    MgLevel& Cn = *H.levels[size_t(l) + 1];
    k_mg_prolong_add<<<blocks_for(L.nv), kMgThreads, 0, s>>>(done, Cn.x.p, L.nx, Rows{0, L.nv}, L.x.p);
    k_mg_post<<<blocks_for(L.nv), kMgThreads, 0, s>>>(done, L.S.p, L.dinv.p, L.nx, L.ny, Rows{0, L.nv}, L.b.p, L.x.p, L.y.p);
    count_launch(2);
    std::swap(L.x.p, L.y.p);  // the smoothed iterate is the level's x from here on
  }
}

// Level 1 of one hierarchy and everything below it.  Replicated: plain V-cycle.  Distributed: level 1 in row strips,
// the level-2 right-hand side is summed over the ranks (every rank restricts its own rows into a zeroed array) and
// levels >= 2 run replicated.
static void vcycle_level1(hdd_mesh* m, MgState& st, MgHierarchy& H, const int* done, cudaStream_t s) {
  if (!st.dist.on) {
    vcycle_from(H, 1, done, s);
    return;
  }
  MgLevel& L = *H.levels[1];
  MgLevel& Cn = *H.levels[2];
  const Rows r1 = level_rows(st, L, 1);
  const int W0 = st.dist.c0 >> 2, W1 = (st.dist.c1 >> 2) + (st.dist.last ? 1 : 0);
  const Rows r2own{int64_t(W0) * (Cn.nx + 1), int64_t(W1 - W0) * (Cn.nx + 1)};
  exchange_rows(m, st, L, 1, {L.b.p}, EX_UP | EX_DOWN);
  k_mg_pre<<<blocks_for(r1.cnt), kMgThreads, 0, s>>>(done, L.S.p, L.dinv.p, L.nx, L.ny, r1, L.b.p, L.x.p, L.r.p);
  exchange_rows(m, st, L, 1, {L.r.p}, EX_UP);
  HDD_CUDA(cudaMemsetAsync(Cn.b.p, 0, size_t(Cn.nv) * sizeof(double), s));
  k_mg_restrict<<<blocks_for(r2own.cnt), kMgThreads, 0, s>>>(done, L.r.p, L.nx, L.ny, r2own, Cn.b.p);
  Nccl::get().all_reduce_sum(Cn.b.p, size_t(Cn.nv), m->comm, s);
  count_launch(2);
  vcycle_from(H, 2, done, s);
  k_mg_prolong_add<<<blocks_for(r1.cnt), kMgThreads, 0, s>>>(done, Cn.x.p, L.nx, r1, L.x.p);
  exchange_rows(m, st, L, 1, {L.x.p}, EX_UP | EX_DOWN);
  k_mg_post<<<blocks_for(r1.cnt), kMgThreads, 0, s>>>(done, L.S.p, L.dinv.p, L.nx, L.ny, r1, L.b.p, L.x.p, L.y.p);
  count_launch(2);
  std::swap(L.x.p, L.y.p);
}

// one V(1,1)-cycle of both hierarchies; level 0 is swept once for the two of them (k_mg_pre2 / k_mg_post2)
static void vcycle_pair(hdd_mesh* m, MgState& st, const int* done, cudaStream_t s) {
  MgLevel& a = *st.h[0].levels[0];
  MgLevel& b = *st.h[1].levels[0];
  MgLevel& a1 = *st.h[0].levels[1];
  MgLevel& b1 = *st.h[1].levels[1];
  const Rows r0 = level_rows(st, a, 0), r1 = level_rows(st, a1, 1);
  exchange_rows(m, st, a, 0, {a.b.p, b.b.p}, EX_UP | EX_DOWN);
  k_mg_pre2<<<blocks_for(r0.cnt), kMgThreads, 0, s>>>(done, a.S.p, a.dinv.p, a.nx, a.ny, r0, a.b.p, b.b.p, a.x.p, b.x.p, a.r.p, b.r.p);
  exchange_rows(m, st, a, 0, {a.r.p, b.r.p}, EX_UP);
  k_mg_restrict<<<blocks_for(r1.cnt), kMgThreads, 0, s>>>(done, a.r.p, a.nx, a.ny, r1, a1.b.p);
  k_mg_restrict<<<blocks_for(r1.cnt), kMgThreads, 0, s>>>(done, b.r.p, b.nx, b.ny, r1, b1.b.p);
  count_launch(3);
  vcycle_level1(m, st, st.h[0], done, s);
  vcycle_level1(m, st, st.h[1], done, s);
  exchange_rows(m, st, a1, 1, {a1.x.p, b1.x.p}, EX_DOWN);  // fine rows up to v1 - 1 interpolate from coarse row V1
  k_mg_prolong_add<<<blocks_for(r0.cnt), kMgThreads, 0, s>>>(done, a1.x.p, a.nx, r0, a.x.p);
  k_mg_prolong_add<<<blocks_for(r0.cnt), kMgThreads, 0, s>>>(done, b1.x.p, b.nx, r0, b.x.p);
  exchange_rows(m, st, a, 0, {a.x.p, b.x.p}, EX_UP | EX_DOWN);
  k_mg_post2<<<blocks_for(r0.cnt), kMgThreads, 0, s>>>(done, a.S.p, a.dinv.p, a.nx, a.ny, r0, a.b.p, b.b.p, a.x.p, b.x.p, a.y.p, b.y.p);
  count_launch(3);
  std::swap(a.x.p, a.y.p);
  std::swap(b.x.p, b.y.p);
}

// Decides, identically on every rank, whether the strips of the ranks are full-width bands of cell rows stacked in rank
// order with boundaries on multiples of four rows - then levels 0 and 1 are swept in strips.  HDD_MG_DISTRIBUTED=0
// keeps everything replicated, =1 asks for strips at any world size; unset, strips are used up to kVerifiedStripWorld
// ranks (the sizes the strip path has been run and checked on; larger jobs take the size-agnostic replicated V-cycle).
constexpr int kVerifiedStripWorld = 4;
static void detect_strips(hdd_swipdg* h, MgState& st) {
  hdd_mesh* m = h->mesh;
  MgDist& d = st.dist;
  d.on = false;
  static const int env = [] { const char* e = std::getenv("HDD_MG_DISTRIBUTED"); return !e ? -1 : (e[0] == '0' ? 0 : 1); }();
  const bool wanted = env < 0 ? m->world <= kVerifiedStripWorld : env == 1;
  if (m->world <= 1) return;
  cudaStream_t s = m->stream;
  DevBuf<int> mm;
  const int init[2] = {INT32_MAX, -1};
  mm.upload(init, 2, s);
  if (m->n_own > 0) k_cell_row_range<<<blocks_for(m->n_own), kMgThreads, 0, s>>>(m->cell_v0.p + m->own0, m->n_own, st.nx, mm.p);
  count_launch();
  int got[2] = {0, 0};
  HDD_CUDA(cudaMemcpyAsync(got, mm.p, sizeof(got), cudaMemcpyDeviceToHost, s));
  HDD_CUDA(cudaStreamSynchronize(s));
  const int c0 = got[0], c1 = got[1] + 1;
  const bool mine = wanted && m->n_own > 0 && int64_t(m->n_own) == int64_t(st.nx) * (c1 - c0) && (c0 % 4 == 0) &&
                    (c1 % 4 == 0 || c1 == st.ny) && st.h[0].levels.size() >= 3;
  // every rank learns every rank's band: 3 doubles per rank through one all-reduce
  std::vector<double> all(size_t(3) * m->world, 0.0);
  all[size_t(3) * m->rank] = c0;
  all[size_t(3) * m->rank + 1] = c1;
  all[size_t(3) * m->rank + 2] = mine ? 1.0 : 0.0;
  DevBuf<double> buf;
  buf.upload(all.data(), all.size(), s);
  Nccl::get().all_reduce_sum(buf.p, all.size(), m->comm, s);
  HDD_CUDA(cudaMemcpyAsync(all.data(), buf.p, all.size() * sizeof(double), cudaMemcpyDeviceToHost, s));
  HDD_CUDA(cudaStreamSynchronize(s));
  bool ok = all[0] == 0.0 && all[size_t(3) * (m->world - 1) + 1] == double(st.ny);
  for (int r = 0; r < m->world; ++r) {
    ok = ok && all[size_t(3) * r + 2] == 1.0;
    if (r > 0) ok = ok && all[size_t(3) * r] == all[size_t(3) * (r - 1) + 1];
  }
  if (!ok) return;
  d.on = true;
  d.c0 = c0;
  d.c1 = c1;
  d.lower = m->rank > 0 ? m->rank - 1 : -1;
  d.upper = m->rank + 1 < m->world ? m->rank + 1 : -1;
  d.last = d.upper < 0;
  if (d.tmp.n < size_t(st.nx) + 1) d.tmp.alloc(size_t(st.nx) + 1);
}

void mg_release(MgState* st) { delete st; }

// (re)builds both hierarchies for the frozen operator `vals`
void mg_setup(hdd_swipdg* h, const double* vals) {
  hdd_mesh* m = h->mesh;
  if (m->sx == 0 || h->polorder != 1)
    HDD_THROW(HDD_ERR_REQUIREMENTS_NOT_MET,
              "solver type 'cg.mg' needs polOrder 1 on a logically structured HDD_CUBE2D grid (vertices numbered x-fastest as by "
              "Stuff::Grid::Providers::Cube / hdd_grid_cube); use 'cg.diagonal' or 'cg.blockdiagonal'");
  cudaStream_t s = m->stream;
  if (!h->mg) h->mg = new MgState;
  MgState& st = *h->mg;
  st.nx = m->sx;
  st.ny = m->sy;
  const int64_t nv = int64_t(m->sx + 1) * (m->sy + 1);
  if (!st.h[0].levels.empty() && (st.h[0].levels[0]->nx != m->sx || st.h[0].levels[0]->ny != m->sy)) {
    st.h[0].levels.clear();
    st.h[1].levels.clear();
  }
  if (st.h[0].levels.empty()) {  // level structure, allocation only
    for (int t = 0; t < 2; ++t) {
      int lx = m->sx, ly = m->sy;
      for (int l = 0;; ++l) {
        std::unique_ptr<MgLevel> L(new MgLevel);
        L->nx = lx;
        L->ny = ly;
        L->nv = int64_t(lx + 1) * (ly + 1);
        if (!(t == 1 && l == 0)) {  // the twisted hierarchy shares the level-0 operator of the plain one
          L->S.alloc(size_t(9) * L->nv);
          L->dinv.alloc(size_t(L->nv));
        }
        L->b.alloc(size_t(L->nv));
        L->x.alloc(size_t(L->nv));
        L->r.alloc(size_t(L->nv));
        L->y.alloc(size_t(L->nv));
        const bool last = L->nv <= kMaxCoarse || (lx & 1) || (ly & 1) || lx < 2 || ly < 2;
        st.h[t].levels.push_back(std::move(L));
        if (last) break;
        lx /= 2;
        ly /= 2;
      }
    }
  }
  MgLevel& f0 = *st.h[0].levels[0];
  k_vertex_galerkin<<<blocks_for(nv), kMgThreads, 0, s>>>(h->view(), vals, m->cell_v0.p, m->lex_cell.p, m->sx, m->sy, Rows{0, nv}, f0.S.p);
  count_launch();
  if (m->world > 1) Nccl::get().all_reduce_sum(f0.S.p, size_t(9) * nv, m->comm, s);
  HDD_CUDA(cudaGetLastError());
  build_hierarchy(h, st.h[0], f0.S.p, false);
  build_hierarchy(h, st.h[1], f0.S.p, true);
  HDD_CUDA(cudaGetLastError());
  detect_strips(h, st);
}

// z += P V(P^T r) + P C V_C(C P^T r), red[1] = r.z; p_init != nullptr also stores z as the first direction
void mg_apply(hdd_swipdg* h, const int* done, const double* r, double* z, double* p_init, double* partial, CgScalars* sc) {
  hdd_mesh* m = h->mesh;
  MgState& st = *h->mg;
  cudaStream_t s = m->stream;
  // HDD_MG_TIMING=1: wall-clock phases of the first applications (synchronising; diagnostics only)
  static const bool timing = [] { const char* e = std::getenv("HDD_MG_TIMING"); return e && e[0] == '1'; }();
  static int timed_calls = 0;
  const bool tt = timing && timed_calls < 6;
  double t_mark = 0.0;
  auto lap = [&](const char* what) {
    if (!tt) return;
    cudaStreamSynchronize(s);
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    const double now = ts.tv_sec + 1e-9 * ts.tv_nsec;
    if (what) std::fprintf(stderr, "[hdd mg rank %d call %d] %-12s %.3f ms\n", m->rank, timed_calls, what, 1e3 * (now - t_mark));
    t_mark = now;
  };
  lap(nullptr);
  MgLevel& a = *st.h[0].levels[0];
  MgLevel& b = *st.h[1].levels[0];
  const bool multi = m->world > 1;
  const MgDist& d = st.dist;
  const int nx1 = st.nx + 1;
  if (d.on) {
    // vertex rows c0 .. c1 of my cells; row c1 belongs to the rank above (which adds my share), row c0 gets the share
    // of the rank below
    const Rows mine{int64_t(d.c0) * nx1, int64_t(d.c1 - d.c0 + 1) * nx1};
    k_dg_restrict<<<blocks_for(mine.cnt), kMgThreads, 0, s>>>(done, r, m->lex_cell.p, m->own0, m->n_own, st.nx, st.ny, mine, a.b.p, nullptr);
    count_launch();
    Nccl& nc = Nccl::get();
    nc.group_start();
    if (d.upper >= 0) nc.send(a.b.p + int64_t(d.c1) * nx1, size_t(nx1), d.upper, m->comm, s);
    if (d.lower >= 0) nc.recv(d.tmp.p, size_t(nx1), d.lower, m->comm, s);
    nc.group_end();
    if (d.lower >= 0) {
      k_add_row<<<(nx1 + 255) / 256, 256, 0, s>>>(done, a.b.p + int64_t(d.c0) * nx1, d.tmp.p, nx1);
      count_launch();
    }
    k_twist_vector<<<blocks_for(level_rows(st, a, 0).cnt), kMgThreads, 0, s>>>(done, a.b.p, st.nx, level_rows(st, a, 0), b.b.p);
    count_launch();
    lap("restrict");
  } else {
    k_dg_restrict<<<blocks_for(a.nv), kMgThreads, 0, s>>>(done, r, m->lex_cell.p, m->own0, m->n_own, st.nx, st.ny, Rows{0, a.nv}, a.b.p,
                                                           multi ? nullptr : b.b.p);
    count_launch();
    lap("restrict");
    if (multi) {
      Nccl::get().all_reduce_sum(a.b.p, size_t(a.nv), m->comm, s);
      lap("all-reduce");
      k_twist_vector<<<blocks_for(a.nv), kMgThreads, 0, s>>>(done, a.b.p, st.nx, Rows{0, a.nv}, b.b.p);
      count_launch();
    }
  }
  if (st.h[0].levels.size() == 1) {
    // the fine vertex grid is already small enough for the dense solve
    k_mg_dense<<<(int(a.nv) + 127) / 128, 128, 0, s>>>(done, st.h[0].coarse_inv.p, int(a.nv), a.b.p, a.x.p);
    k_mg_dense<<<(int(b.nv) + 127) / 128, 128, 0, s>>>(done, st.h[1].coarse_inv.p, int(b.nv), b.b.p, b.x.p);
    count_launch(2);
  } else {
    vcycle_pair(m, st, done, s);
    lap("v-cycles");
  }
  // the cells of my top row read the vertex row above my strip
  exchange_rows(m, st, a, 0, {a.x.p, b.x.p}, EX_DOWN);
  const int grid = int(std::min<int64_t>((m->n_own + kMgThreads - 1) / kMgThreads, kMaxBlocks));
  k_dg_prolong_dot<<<grid, kMgThreads, 0, s>>>(done, m->n_own, m->cell_v0.p + m->own0, st.nx, a.x.p, b.x.p, r, z, p_init, partial, sc);
  count_launch();
  HDD_CUDA(cudaGetLastError());
  lap("prolong+dot");
  if (tt) ++timed_calls;
}

int mg_num_levels(const hdd_swipdg* h) { return h->mg ? int(h->mg->h[0].levels.size()) : 0; }

}  // namespace hdd
