// Half solve against stored factors (one level): x_k = D^{-1} y[0::2], yn = y[1::2] - U x_k.
// Replaces one iteration of the reference's halfsolve (cyclic_gps/cyclic_reduction.py:318-333,
// Ux :40-60).  This entry serves the stand-alone halfsolve() / solve() / mahal() API (factor once,
// solve many); the likelihood hot path gets x_k from the fused forward kernel instead.
//
// Mapping: a group of LG lanes per node, lane r = row r of the node's blocks.  A block is contiguous in global
// memory and every lane reads one whole row of it, so the group's loads are coalesced and each factor byte is read
// exactly once (the first version used one thread per node with block-strided reads and ell registers of state per
// thread, which collapsed for large blocks).  x = D^{-1} y is a forward substitution over the lanes of the group
// (one shuffle + one FMA per column, every lane collects the finished x); the products F x and G x are row dot
// products against the two x vectors of the neighbouring even nodes, which every lane of the group reads as a
// broadcast.
#pragma once
#include "cr_common.cuh"

namespace crb200 {

using HalfSolveArgs = ::crb200_hs_args;

constexpr int kHsThreads = 256;

// one row of L elements from global memory (16-byte loads when the row is a multiple of 16 bytes and aligned)
template <typename T, int L>
__device__ __forceinline__ void ldg_row(T (&v)[L], const T* __restrict__ g, const bool vec_ok) {
  constexpr int VE = 16 / (int)sizeof(T);
  if constexpr ((L % VE) == 0) {
    if (vec_ok) {
#pragma unroll
      for (int i = 0; i < L; i += VE) {
        if constexpr (sizeof(T) == 4) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(g + i));
          v[i] = t.x; v[i + 1] = t.y; v[i + 2] = t.z; v[i + 3] = t.w;
        } else {
          const double2 t = __ldg(reinterpret_cast<const double2*>(g + i));
          v[i] = t.x; v[i + 1] = t.y;
        }
      }
      return;
    }
  }
#pragma unroll
  for (int i = 0; i < L; ++i) v[i] = __ldg(g + i);
}

template <typename T, int L>
__global__ void __launch_bounds__(kHsThreads) cr_hs_x_kernel(const HalfSolveArgs a) {
  constexpr int LG = GroupLanes<L>::value, GPB = kHsThreads / LG;
  __shared__ double sred[32];
  const int E = (a.m + 1) >> 1;
  const long long total = (long long)a.batch * E;
  const long long node_raw = (long long)blockIdx.x * GPB + threadIdx.x / LG;
  const bool act = node_raw < total;
  const long long node = act ? node_raw : total - 1;          // idle groups shadow the last node (shuffles need every lane)
  const int r = threadIdx.x % LG;
  const bool rowok = r < L;
  const int rr = rowok ? r : 0;
  const int b = (int)(node / E);
  const int e = (int)(node - (long long)b * E);
  const T* D = static_cast<const T*>(a.D) + (size_t)node * (L * L) + rr * L;
  T drow[L];
  ldg_row<T, L>(drow, D, is_aligned16(a.D));
  T acc = static_cast<const T*>(a.y)[(size_t)b * a.stridey + (size_t)(2 * e) * L + rr];
  T inv = T(1);
#pragma unroll
  for (int k = 0; k < L; ++k) if (k == rr) inv = T(1) / drow[k];
  T xr = T(0);
#pragma unroll
  for (int k = 0; k < L; ++k) {
    const T xk = __shfl_sync(0xffffffffu, acc * inv, k, LG);   // lane k holds the finished x_k
    if (rr == k) xr = xk;
    if (rr > k) acc = fma(-drow[k], xk, acc);
  }
  double part = 0.0;
  if (act && rowok) {
    static_cast<T*>(a.xk)[(size_t)node * L + r] = xr;
    part = (double)xr * (double)xr;
  }
  if (a.mahal != nullptr) {
    // one atomic per CTA when all its nodes belong to one series, else one per group
    const long long first = (long long)blockIdx.x * GPB;
    const long long last = (first + GPB - 1 < total ? first + GPB - 1 : total - 1);
    const bool uniform = (first / E) == (last / E);
    if (uniform) {
      const double t = block_sum(part, sred);
      if (threadIdx.x == 0) atomicAdd(a.mahal + (int)(first / E), t);
    } else {
#pragma unroll
      for (int off = LG / 2; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off, LG);
      if (act && r == 0) atomicAdd(a.mahal + b, part);
    }
  }
}

template <typename T, int L>
__global__ void __launch_bounds__(kHsThreads) cr_hs_y_kernel(const HalfSolveArgs a) {
  constexpr int LG = GroupLanes<L>::value, GPB = kHsThreads / LG;
  const int E = (a.m + 1) >> 1, o = a.m >> 1, gcnt = (a.m - 1) >> 1;
  const long long node = (long long)blockIdx.x * GPB + threadIdx.x / LG;   // flat (series, odd node j)
  const int r = threadIdx.x % LG;
  if (node >= (long long)a.batch * o || r >= L) return;
  const int b = (int)(node / o);
  const int j = (int)(node - (long long)b * o);
  const T* x0 = static_cast<const T*>(a.xk) + ((size_t)b * E + j) * L;
  T s = static_cast<const T*>(a.y)[(size_t)b * a.stridey + (size_t)(2 * j + 1) * L + r];
  {
    T row[L], xv[L];
    ldg_row<T, L>(row, static_cast<const T*>(a.F) + (size_t)node * (L * L) + r * L, is_aligned16(a.F));
    ldg_row<T, L>(xv, x0, is_aligned16(a.xk));
#pragma unroll
    for (int k = 0; k < L; ++k) s = fma(-row[k], xv[k], s);
  }
  if (j < gcnt) {
    T row[L], xv[L];
    ldg_row<T, L>(row, static_cast<const T*>(a.G) + ((size_t)b * gcnt + j) * (L * L) + r * L, is_aligned16(a.G));
    ldg_row<T, L>(xv, x0 + L, is_aligned16(a.xk));
#pragma unroll
    for (int k = 0; k < L; ++k) s = fma(-row[k], xv[k], s);
  }
  static_cast<T*>(a.yn)[(size_t)node * L + r] = s;
}

template <typename T, int L>
cudaError_t launch_level_halfsolve(const HalfSolveArgs& a, cudaStream_t stream) {
  constexpr int GPB = kHsThreads / GroupLanes<L>::value;
  const long long E = (a.m + 1) / 2, o = a.m / 2;
  const long long n1 = E * a.batch, n2 = o * a.batch;
  if (n1 > 0) {
    cr_hs_x_kernel<T, L><<<(unsigned)((n1 + GPB - 1) / GPB), kHsThreads, 0, stream>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  if (n2 > 0 && a.yn != nullptr) {
    cr_hs_y_kernel<T, L><<<(unsigned)((n2 + GPB - 1) / GPB), kHsThreads, 0, stream>>>(a);
    return cudaGetLastError();
  }
  return cudaSuccess;
}

}  // namespace crb200
