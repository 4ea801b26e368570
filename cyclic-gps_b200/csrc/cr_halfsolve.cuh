// Half solve against stored factors (one level): x_k = D^{-1} y[0::2], yn = y[1::2] - U x_k.
// Replaces one iteration of the reference's halfsolve (cyclic_gps/cyclic_reduction.py:318-333,
// Ux :40-60).  One thread per node, factors read straight from global memory: this entry is
// only used by the stand-alone halfsolve()/solve()/mahal() API (factor-once, solve-many);
// the likelihood hot path gets x_k from the fused forward kernel instead.
#pragma once
#include "cr_common.cuh"

namespace crb200 {

using HalfSolveArgs = ::crb200_hs_args;

template <typename T, int L>
__global__ void __launch_bounds__(128) cr_hs_x_kernel(const HalfSolveArgs a) {
  __shared__ double sred[32];
  const int E = (a.m + 1) >> 1;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool act = idx < (long long)a.batch * E;
  double part = 0.0;
  int b = 0;
  if (act) {
    b = (int)(idx / E);
    const int e = (int)(idx - (long long)b * E);
    const T* D = static_cast<const T*>(a.D) + (size_t)idx * (L * L);
    const T* y = static_cast<const T*>(a.y) + (size_t)b * a.stridey + (size_t)(2 * e) * L;
    T x[L];
#pragma unroll
    for (int c = 0; c < L; ++c) {
      T s = y[c];
#pragma unroll
      for (int k = 0; k < c; ++k) s = fma(-x[k], D[c * L + k], s);
      x[c] = s / D[c * L + c];
    }
    T* xo = static_cast<T*>(a.xk) + (size_t)idx * L;
#pragma unroll
    for (int c = 0; c < L; ++c) { xo[c] = x[c]; part += (double)x[c] * (double)x[c]; }
  }
  if (a.mahal != nullptr) {
    // series may differ inside a CTA only when E < blockDim.x; use per-thread atomics then
    const long long first = (long long)blockIdx.x * blockDim.x;
    const long long last = first + blockDim.x - 1;
    const bool uniform = (first / E) == ((last < (long long)a.batch * E ? last : (long long)a.batch * E - 1) / E);
    if (uniform) {
      const double t = block_sum(part, sred);
      if (threadIdx.x == 0) atomicAdd(a.mahal + (int)(first / E), t);
    } else if (act) {
      atomicAdd(a.mahal + b, part);
    }
  }
}

template <typename T, int L>
__global__ void __launch_bounds__(128) cr_hs_y_kernel(const HalfSolveArgs a) {
  const int E = (a.m + 1) >> 1, o = a.m >> 1, gcnt = (a.m - 1) >> 1;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)a.batch * o) return;
  const int b = (int)(idx / o);
  const int j = (int)(idx - (long long)b * o);
  const T* F = static_cast<const T*>(a.F) + (size_t)idx * (L * L);
  const T* x0 = static_cast<const T*>(a.xk) + ((size_t)b * E + j) * L;
  const T* y = static_cast<const T*>(a.y) + (size_t)b * a.stridey + (size_t)(2 * j + 1) * L;
  T acc[L];
#pragma unroll
  for (int r = 0; r < L; ++r) {
    T s = y[r];
#pragma unroll
    for (int k = 0; k < L; ++k) s = fma(-F[r * L + k], x0[k], s);
    acc[r] = s;
  }
  if (j < gcnt) {
    const T* G = static_cast<const T*>(a.G) + ((size_t)b * gcnt + j) * (L * L);
    const T* x1 = x0 + L;
#pragma unroll
    for (int r = 0; r < L; ++r) {
      T s = acc[r];
#pragma unroll
      for (int k = 0; k < L; ++k) s = fma(-G[r * L + k], x1[k], s);
      acc[r] = s;
    }
  }
  T* yo = static_cast<T*>(a.yn) + (size_t)idx * L;
#pragma unroll
  for (int r = 0; r < L; ++r) yo[r] = acc[r];
}

template <typename T, int L>
cudaError_t launch_level_halfsolve(const HalfSolveArgs& a, cudaStream_t stream) {
  const long long E = (a.m + 1) / 2, o = a.m / 2;
  const long long n1 = E * a.batch, n2 = o * a.batch;
  if (n1 > 0) {
    cr_hs_x_kernel<T, L><<<(unsigned)((n1 + 127) / 128), 128, 0, stream>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  if (n2 > 0 && a.yn != nullptr) {
    cr_hs_y_kernel<T, L><<<(unsigned)((n2 + 127) / 128), 128, 0, stream>>>(a);
    return cudaGetLastError();
  }
  return cudaSuccess;
}

}  // namespace crb200
