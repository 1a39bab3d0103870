// Deterministic block / grid reductions shared by the CG and multigrid kernels (sm_100a).
#pragma once
#include "kernels.hpp"

namespace hdd {

constexpr int kMaxBlocks = 148 * 8;  // capacity of the per-block partial arrays (see cg_partial_capacity)

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-level sum of N values; result valid in thread 0.  Works for up to 32 warps.
template <int N>
__device__ __forceinline__ void block_sum(double (&v)[N], double* smem /* [N*32] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
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

}  // namespace hdd
