// K5 (SpMV) and K6 (fused CG vector updates) for sm_100a.
//
// Matrix storage: the CSR value array of the owned rows (bit-compatible with what hdd_component_values hands out),
// but the kernels never read CSR column indices or row pointers.  The DG matrix is block structured - every row of
// cell T has one dense n_loc-wide run per member of sort({T} u neighbours) - so the column structure is recovered
// from the 16-byte neighbour record of the cell and one 8-byte block offset per cell.  That removes the 4 B/nnz
// index stream of textbook CSR: 8.05 B/nnz instead of 12 B/nnz of HBM traffic for Q1.
//
// CG is the classic Hestenes-Stiefel recurrence (same algebra as oracle/or_cg) with device-resident scalars:
// no host synchronisation inside the iteration, dot products reduced deterministically (fixed per-block partials,
// last-arriving block sums them in a fixed order), convergence latched in a ping-pong flag so that iterations
// launched after convergence are no-ops.
#include "kernels.hpp"

namespace hdd {

namespace {

constexpr int kCgThreads = 256;
constexpr int kMaxBlocks = 148 * 8;

template <int NF>
__device__ __forceinline__ void load_neigh(const int32_t* neigh, int k, int* nb) {
  if constexpr (NF == 4) {
    const int4 v = __ldg(reinterpret_cast<const int4*>(neigh) + k);
    nb[0] = v.x; nb[1] = v.y; nb[2] = v.z; nb[3] = v.w;
  } else {
#pragma unroll
    for (int f = 0; f < NF; ++f) nb[f] = __ldg(neigh + size_t(NF) * k + f);
  }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-level sum of up to 3 values; result valid in thread 0.
template <int N>
__device__ __forceinline__ void block_sum(double (&v)[N], double* smem /* [N*32] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int k = 0; k < N; ++k) v[k] = warp_sum(v[k]);
  __syncthreads();
  if (lane == 0)
#pragma unroll
    for (int k = 0; k < N; ++k) smem[k * 32 + warp] = v[k];
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < N; ++k) {
      double t = lane < nw ? smem[k * 32 + lane] : 0.0;
      v[k] = warp_sum(t);
    }
  }
}

// Deterministic grid-wide sum: every block stores its partial, the last block to arrive adds all partials in a
// fixed order and hands the totals to `finish` (executed by thread 0 of that block only).
template <int N, class Finish>
__device__ __forceinline__ void grid_sum(double (&v)[N], double* partial, unsigned int* ticket, Finish finish) {
  __shared__ double red[N * 32];
  __shared__ bool is_last;
  block_sum<N>(v, red);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < N; ++k) partial[k * kMaxBlocks + blockIdx.x] = v[k];
    __threadfence();
    const unsigned int t = atomicInc(ticket, gridDim.x - 1);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  double w[N];
#pragma unroll
  for (int k = 0; k < N; ++k) {
    w[k] = 0.0;
    for (int b = threadIdx.x; b < int(gridDim.x); b += blockDim.x) w[k] += __ldcg(partial + k * kMaxBlocks + b);
  }
  block_sum<N>(w, red);
  if (threadIdx.x == 0) finish(w);
}

// Row kernel shared by SpMV and the CG step: y_t = sum over the blocks of row t, ascending column order.
template <int KIND>
__device__ __forceinline__ double row_times_x(const MeshView& m, const double* __restrict__ vals,
                                              const double* __restrict__ x, int64_t t) {
  constexpr int NL = Geo<KIND>::NL, NF = Geo<KIND>::NF;
  const int k = int(t / NL), i = int(t % NL);
  int nb[NF];
  load_neigh<NF>(m.neigh, k, nb);
  const int self = m.own0 + k;
  int cells[NF + 1];
  int nblk = 1;
#pragma unroll
  for (int f = 0; f < NF; ++f) nblk += nb[f] >= 0 ? 1 : 0;
  // scatter cells into their sorted slots
#pragma unroll
  for (int s = 0; s < NF + 1; ++s) cells[s] = 0;
  {
    const int slot = block_slot<NF>(self, nb, self);
#pragma unroll
    for (int s = 0; s < NF + 1; ++s)
      if (s == slot) cells[s] = self;
  }
#pragma unroll
  for (int f = 0; f < NF; ++f)
    if (nb[f] >= 0) {
      const int slot = block_slot<NF>(self, nb, nb[f]);
#pragma unroll
      for (int s = 0; s < NF + 1; ++s)
        if (s == slot) cells[s] = nb[f];
    }
  const double* row = vals + m.blk_start[k] * (NL * NL) + int64_t(i) * nblk * NL;
  double sum = 0.0;
#pragma unroll
  for (int s = 0; s < NF + 1; ++s) {
    if (s < nblk) {
      const double* xs = x + size_t(NL) * cells[s];
      if constexpr (NL == 4) {
        double a0, a1, a2, a3, x0, x1, x2, x3;
        asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
                     : "=d"(a0), "=d"(a1), "=d"(a2), "=d"(a3)
                     : "l"(row + s * 4));
        asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(x0), "=d"(x1), "=d"(x2), "=d"(x3) : "l"(xs));
        sum = fma(a0, x0, sum); sum = fma(a1, x1, sum); sum = fma(a2, x2, sum); sum = fma(a3, x3, sum);
      } else {
#pragma unroll
        for (int j = 0; j < NL; ++j) sum = fma(row[s * NL + j], xs[j], sum);
      }
    }
  }
  return sum;
}

template <int KIND>
__global__ void __launch_bounds__(kCgThreads)
    k_spmv(MeshView m, const double* __restrict__ vals, const double* __restrict__ x, double* __restrict__ y) {
  constexpr int NL = Geo<KIND>::NL;
  const int64_t rows = int64_t(m.n_own) * NL;
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  for (int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; t < rows; t += stride)
    y[t] = row_times_x<KIND>(m, vals, x, t);
}

__global__ void __launch_bounds__(kCgThreads) k_cg_init(int64_t rows, int64_t own_off, CgBuffers c) {
  double v[2] = {0.0, 0.0};
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  for (int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; t < rows; t += stride) {
    const double b = c.b[t], z = c.dinv[t] * b;
    c.x[t] = 0.0;
    c.r[t] = b;
    c.p[own_off + t] = z;
    v[0] += b * z;
    v[1] += b * b;
  }
  CgScalars* sc = c.sc;
  grid_sum<2>(v, c.partial, &sc->ticket_a, [sc](const double(&w)[2]) {
    sc->red[1] = w[0];  // local r.z
    sc->red[2] = w[1];  // local b.b
  });
}

// after the (optional) all-reduce of red[1..2]
__global__ void k_cg_init_finish(CgScalars* sc) {
  sc->rz[0] = sc->red[1];
  sc->bb = sc->red[2];
  sc->rr = sc->red[2];
  sc->it[0] = 0;
  sc->done[0] = (sc->red[2] == 0.0 || sc->max_it <= 0) ? 1 : 0;
}

template <int KIND>
__global__ void __launch_bounds__(kCgThreads) k_cg_spmv(MeshView m, CgBuffers c, int par) {
  constexpr int NL = Geo<KIND>::NL;
  CgScalars* sc = c.sc;
  if (sc->done[par]) return;
  const int64_t rows = int64_t(m.n_own) * NL, own_off = int64_t(m.own0) * NL;
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  double v[1] = {0.0};
  for (int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; t < rows; t += stride) {
    const double y = row_times_x<KIND>(m, c.values, c.p, t);
    c.q[t] = y;
    v[0] = fma(c.p[own_off + t], y, v[0]);
  }
  grid_sum<1>(v, c.partial, &sc->ticket_a, [sc](const double(&w)[1]) { sc->red[0] = w[0]; });
}

__global__ void __launch_bounds__(kCgThreads) k_cg_update(int64_t rows, int64_t own_off, CgBuffers c, int par) {
  CgScalars* sc = c.sc;
  if (sc->done[par]) return;
  const double alpha = sc->rz[par] / sc->red[0];
  double v[2] = {0.0, 0.0};
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  for (int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; t < rows; t += stride) {
    c.x[t] = fma(alpha, c.p[own_off + t], c.x[t]);
    const double r = fma(-alpha, c.q[t], c.r[t]);
    c.r[t] = r;
    const double z = c.dinv[t] * r;
    v[0] = fma(r, z, v[0]);
    v[1] = fma(r, r, v[1]);
  }
  grid_sum<2>(v, c.partial, &sc->ticket_b, [sc](const double(&w)[2]) {
    sc->red[1] = w[0];
    sc->red[2] = w[1];
  });
}

__global__ void __launch_bounds__(kCgThreads) k_cg_direction(int64_t rows, int64_t own_off, CgBuffers c, int par) {
  CgScalars* sc = c.sc;
  const int done = sc->done[par];
  if (!done) {
    const double beta = sc->red[1] / sc->rz[par];
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    for (int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; t < rows; t += stride)
      c.p[own_off + t] = fma(beta, c.p[own_off + t], c.dinv[t] * c.r[t]);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    if (done) {
      sc->rz[par ^ 1] = sc->rz[par];
      sc->it[par ^ 1] = sc->it[par];
      sc->done[par ^ 1] = 1;
    } else {
      const int it = sc->it[par] + 1;
      sc->rz[par ^ 1] = sc->red[1];
      sc->rr = sc->red[2];
      sc->it[par ^ 1] = it;
      sc->done[par ^ 1] = (sc->red[2] <= sc->tol2 * sc->bb || it >= sc->max_it) ? 1 : 0;
    }
  }
}

__global__ void k_pack(const double* __restrict__ v, const int32_t* __restrict__ idx, int64_t n,
                       double* __restrict__ out) {
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  for (int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; t < n; t += stride) out[t] = v[idx[t]];
}

inline int cg_grid(int64_t rows) {
  const int64_t need = (rows + kCgThreads - 1) / kCgThreads;
  return int(std::max<int64_t>(1, std::min<int64_t>(need, kMaxBlocks)));
}

}  // namespace

int cg_partial_capacity() { return 3 * kMaxBlocks; }

void launch_spmv(const MeshView& m, const double* values, const double* x_local, double* y, cudaStream_t s) {
  const int64_t rows = int64_t(m.n_own) * m.nl;
  if (rows == 0) return;
  if (m.kind == HDD_SIMPLEX2D)
    k_spmv<HDD_SIMPLEX2D><<<cg_grid(rows), kCgThreads, 0, s>>>(m, values, x_local, y);
  else
    k_spmv<HDD_CUBE2D><<<cg_grid(rows), kCgThreads, 0, s>>>(m, values, x_local, y);
  count_launch();
  HDD_CUDA(cudaGetLastError());
}

void launch_cg_init(const MeshView& m, const CgBuffers& c, double precision, int max_it, cudaStream_t s) {
  const int64_t rows = int64_t(m.n_own) * m.nl;
  CgScalars h{};
  h.tol2 = precision * precision;
  h.max_it = max_it;
  HDD_CUDA(cudaMemcpyAsync(c.sc, &h, sizeof(h), cudaMemcpyHostToDevice, s));
  k_cg_init<<<cg_grid(rows), kCgThreads, 0, s>>>(rows, int64_t(m.own0) * m.nl, c);
  count_launch();
  HDD_CUDA(cudaGetLastError());
}

void launch_cg_init_finish(const MeshView&, const CgBuffers& c, cudaStream_t s) {
  k_cg_init_finish<<<1, 1, 0, s>>>(c.sc);
  count_launch();
  HDD_CUDA(cudaGetLastError());
}

void launch_cg_spmv(const MeshView& m, const CgBuffers& c, int parity, cudaStream_t s) {
  const int64_t rows = int64_t(m.n_own) * m.nl;
  if (m.kind == HDD_SIMPLEX2D)
    k_cg_spmv<HDD_SIMPLEX2D><<<cg_grid(rows), kCgThreads, 0, s>>>(m, c, parity);
  else
    k_cg_spmv<HDD_CUBE2D><<<cg_grid(rows), kCgThreads, 0, s>>>(m, c, parity);
  count_launch();
  HDD_CUDA(cudaGetLastError());
}

void launch_cg_update(const MeshView& m, const CgBuffers& c, int parity, cudaStream_t s) {
  const int64_t rows = int64_t(m.n_own) * m.nl;
  k_cg_update<<<cg_grid(rows), kCgThreads, 0, s>>>(rows, int64_t(m.own0) * m.nl, c, parity);
  count_launch();
  HDD_CUDA(cudaGetLastError());
}

void launch_cg_direction(const MeshView& m, const CgBuffers& c, int parity, cudaStream_t s) {
  const int64_t rows = int64_t(m.n_own) * m.nl;
  k_cg_direction<<<cg_grid(rows), kCgThreads, 0, s>>>(rows, int64_t(m.own0) * m.nl, c, parity);
  count_launch();
  HDD_CUDA(cudaGetLastError());
}

void launch_pack(const double* v_local, const int32_t* dof_idx, int64_t n, double* out, cudaStream_t s) {
  if (n == 0) return;
  k_pack<<<int(std::min<int64_t>((n + 255) / 256, 1024)), 256, 0, s>>>(v_local, dof_idx, n, out);
  count_launch();
  HDD_CUDA(cudaGetLastError());
}

}  // namespace hdd
