// Common device helpers for the cyclic-reduction level kernels (sm_100a).
//
// Vocabulary: a level with m block-rows has E = ceil(m/2) even
// (eliminated) nodes, o = floor(m/2) odd (surviving) nodes and g = floor((m-1)/2) G links.
// A "group" is LG consecutive lanes that own one even node; lane r of a group owns row r of
// every ell x ell block of that node.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include "crb200.h"

namespace crb200 {

constexpr int kMaxEll = 32;
constexpr int kThreads = 256;

template <int L>
struct GroupLanes {
  static constexpr int value = L <= 1 ? 1 : L <= 2 ? 2 : L <= 4 ? 4 : L <= 8 ? 8 : L <= 16 ? 16 : 32;
};

constexpr int kMaxDevices = 64;

// Opt a kernel into its dynamic shared-memory size once per device.  Thread-safe (the flags are atomics; a race only
// repeats the idempotent attribute call); devices beyond kMaxDevices set the attribute on every launch.  The CURRENT
// device must be the one that owns the stream the kernel is launched on (include/crb200.h, "Threading").
template <typename Kernel>
inline cudaError_t ensure_dynamic_smem(Kernel kernel, int bytes, std::atomic<unsigned char>* done) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const bool tracked = dev >= 0 && dev < kMaxDevices;
  if (tracked && done[dev].load(std::memory_order_acquire)) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return e;
  if (tracked) done[dev].store(1, std::memory_order_release);
  return cudaSuccess;
}

__host__ __device__ constexpr int cmin(int a, int b) { return a < b ? a : b; }
__host__ __device__ constexpr int cmax(int a, int b) { return a > b ? a : b; }
__host__ __device__ constexpr size_t align16(size_t x) { return (x + 15) & ~size_t(15); }

// ----------------------------------------------------------------------------------------
// async global -> shared copies (cp.async / LDGSTS).  Every thread of the CTA calls these
// with the same arguments; completion is cp_async_wait_all() + __syncthreads().
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}

template <typename T>
__device__ __forceinline__ void cp_async_elem(T* smem, const T* gmem) {
  if constexpr (sizeof(T) == 8) cp_async8(smem, gmem); else cp_async4(smem, gmem);
}

// Copy `nunits` units of UE contiguous elements each from global (units contiguous) to
// shared memory where unit u lands at sdst + u * SMULT * UE (SMULT = 1: flat, 2: every
// second slot).  `vec_ok` = global base is 16-byte aligned (checked by the caller).
template <typename T, int UE, int SMULT>
__device__ __forceinline__ void tile_g2s(T* sdst, const T* __restrict__ gsrc, int nunits, bool vec_ok) {
  constexpr int VE = 16 / (int)sizeof(T);
  if (nunits <= 0) return;
  if constexpr ((UE % VE) == 0) {
    if (vec_ok) {
      constexpr int CPU = UE / VE;  // 16-byte chunks per unit
      const int total = nunits * CPU;
      for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int u = i / CPU, c = i - u * CPU;
        cp_async16(sdst + (size_t)u * (SMULT * UE) + c * VE, gsrc + (size_t)i * VE);
      }
      return;
    }
  }
  const int total = nunits * UE;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int u = i / UE, c = i - u * UE;
    cp_async_elem(sdst + (size_t)u * (SMULT * UE) + c, gsrc + i);
  }
}

// Shared -> global, same addressing convention.
template <typename T, int UE, int SMULT>
__device__ __forceinline__ void tile_s2g(T* __restrict__ gdst, const T* ssrc, int nunits, bool vec_ok) {
  constexpr int VE = 16 / (int)sizeof(T);
  if (nunits <= 0) return;
  if constexpr ((UE % VE) == 0) {
    if (vec_ok) {
      constexpr int CPU = UE / VE;
      const int total = nunits * CPU;
      for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int u = i / CPU, c = i - u * CPU;
        const int4 v = *reinterpret_cast<const int4*>(ssrc + (size_t)u * (SMULT * UE) + c * VE);
        *reinterpret_cast<int4*>(gdst + (size_t)i * VE) = v;
      }
      return;
    }
  }
  const int total = nunits * UE;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int u = i / UE, c = i - u * UE;
    gdst[i] = ssrc[(size_t)u * (SMULT * UE) + c];
  }
}

// ----------------------------------------------------------------------------------------
// TMA 1-D bulk copies (cp.async.bulk, SASS UBLKCP) + mbarrier completion.  Addresses and
// sizes must be multiples of 16 bytes.
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* sdst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(sdst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* ssrc, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_commit_and_wait_read() {
  asm volatile("cp.async.bulk.commit_group;\ncp.async.bulk.wait_group.read 0;\n" ::: "memory");
}

__device__ __forceinline__ bool is_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ----------------------------------------------------------------------------------------
// row load / store between shared memory and a register array (vectorised when possible)
// ----------------------------------------------------------------------------------------
template <typename T, int L>
__device__ __forceinline__ void lds_row(T (&a)[L], const T* p) {
  constexpr int VE = 16 / (int)sizeof(T);
  if constexpr ((L % VE) == 0) {
#pragma unroll
    for (int c = 0; c < L; c += VE) {
      if constexpr (sizeof(T) == 4) {
        const float4 v = *reinterpret_cast<const float4*>(p + c);
        a[c] = v.x; a[c + 1] = v.y; a[c + 2] = v.z; a[c + 3] = v.w;
      } else {
        const double2 v = *reinterpret_cast<const double2*>(p + c);
        a[c] = v.x; a[c + 1] = v.y;
      }
    }
  } else {
#pragma unroll
    for (int c = 0; c < L; ++c) a[c] = p[c];
  }
}

template <typename T, int L>
__device__ __forceinline__ void sts_row(T* p, const T (&a)[L]) {
  constexpr int VE = 16 / (int)sizeof(T);
  if constexpr ((L % VE) == 0) {
#pragma unroll
    for (int c = 0; c < L; c += VE) {
      if constexpr (sizeof(T) == 4) {
        *reinterpret_cast<float4*>(p + c) = make_float4(a[c], a[c + 1], a[c + 2], a[c + 3]);
      } else {
        *reinterpret_cast<double2*>(p + c) = make_double2(a[c], a[c + 1]);
      }
    }
  } else if constexpr (sizeof(T) == 4 && (L % 2) == 0) {
#pragma unroll
    for (int c = 0; c < L; c += 2) *reinterpret_cast<float2*>(p + c) = make_float2(a[c], a[c + 1]);
  } else {
#pragma unroll
    for (int c = 0; c < L; ++c) p[c] = a[c];
  }
}

// acc[c] += sum_k coef[k] * M[k*L + c]   (M: L x L row-major in shared memory, rows are
// broadcast reads: every lane of the group reads the same addresses)
template <typename T, int L>
__device__ __forceinline__ void row_times_mat(T (&acc)[L], const T (&coef)[L], const T* M) {
#pragma unroll
  for (int k = 0; k < L; ++k) {
    T row[L];
    lds_row<T, L>(row, M + k * L);
#pragma unroll
    for (int c = 0; c < L; ++c) acc[c] = fma(coef[k], row[c], acc[c]);
  }
}

// acc[c] += sum_k coef[k] * M[c*L + k]   (acc = coef * M^T, i.e. dot products with rows of M)
template <typename T, int L>
__device__ __forceinline__ void row_times_matT(T (&acc)[L], const T (&coef)[L], const T* M) {
#pragma unroll
  for (int c = 0; c < L; ++c) {
    T row[L];
    lds_row<T, L>(row, M + c * L);
    T s = acc[c];
#pragma unroll
    for (int k = 0; k < L; ++k) s = fma(coef[k], row[k], s);
    acc[c] = s;
  }
}

template <typename T>
__device__ __forceinline__ T shfl_grp(T v, int src, int width) {
  return __shfl_sync(0xffffffffu, v, src, width);
}

// CTA-wide sum of a double; result valid on thread 0.  `scratch` holds >= 32 doubles.
__device__ __forceinline__ double block_sum(double v, double* scratch) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = (blockDim.x + 31) >> 5;
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  double t = 0.0;
  if (warp == 0) {
    t = lane < nwarps ? scratch[lane] : 0.0;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
  }
  __syncthreads();
  return t;
}

}  // namespace crb200
