// Record-based staging for the thread-per-node kernels: one padded shared-memory record per
// node, filled from / drained to flat coalesced global ranges by a single warp.
//
// All loops address shared memory through 32-bit shared-window addresses and walk global
// memory with a running pointer, so a 16-byte chunk costs ~4 instructions (cp.async or
// LDS.128 + STG.128, two integer ops, loop control).
#pragma once
#include "cr_common.cuh"
#include "cr_multi.h"

#ifndef CRB200_SMEM_PAD
#define CRB200_SMEM_PAD 0      // extra dynamic shared memory per CTA (lowers the CTAs/SM; profiling experiments)
#endif
#ifndef CRB200_CS_MAX_CTAS
#define CRB200_CS_MAX_CTAS 8   // cap of the resident-CTA hint of the column-split kernels: at 9-10 CTAs ptxas limits the fp64 ell = 8
                               // kernels to 168 registers and spills ~200 B; at 8 they get 224 / 242 registers, no spills (+1 %)
#endif
#ifndef CRB200_TPN_MAX_BLOCK_BYTES
#define CRB200_TPN_MAX_BLOCK_BYTES 400   // largest block the thread-per-node kernels take: fp32 ell <= 10, fp64 ell <= 7.  Up to 324 B (fp32 ell = 9,
                                          // fp64 ell = 6) P, Q and Sigma_ee fit in 254 registers; at 392-400 B the backward kernel spills ~0.5 KB and
                                          // is still 1.2-1.5x faster than the lane-per-row kernels (profiles/r1_summary.md)
#endif
#ifndef CRB200_TPN_WARPS
#define CRB200_TPN_WARPS 1     // independent single-tile warps per CTA of the thread-per-node kernels
#endif

namespace crb200 {

__device__ __forceinline__ void cp_async16_u32(unsigned saddr, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(saddr), "l"(gmem) : "memory");
}
template <typename T>
__device__ __forceinline__ void cp_async_elem_u32(unsigned saddr, const T* gmem) {
  if constexpr (sizeof(T) == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(saddr), "l"(gmem) : "memory");
  else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(saddr), "l"(gmem) : "memory");
}
template <typename T>
__device__ __forceinline__ T lds_elem_u32(unsigned saddr) {
  T v;
  if constexpr (sizeof(T) == 8) asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(v) : "r"(saddr));
  else asm volatile("ld.shared.f32 %0, [%1];\n" : "=f"(v) : "r"(saddr));
  return v;
}
// Element-by-element walk over records (blocks that are not a multiple of 16 bytes): lane starts at element
// k0 = lane + first, every step advances 32 elements; (record, offset) are carried along instead of being
// recomputed with a division per element (that arithmetic cost as much as the block algebra at ell = 7).
template <int EPR>
struct ElemWalk {
  static constexpr unsigned SR = 32 / EPR, SC = 32 % EPR;
  unsigned c, saddr;
  __device__ __forceinline__ ElemWalk(unsigned srec0, unsigned nsb, unsigned es, unsigned first) {
    const unsigned k0 = (threadIdx.x & 31) + first;
    const unsigned rec = k0 / EPR;
    c = k0 - rec * EPR;
    saddr = srec0 + rec * nsb + c * es;
  }
  __device__ __forceinline__ void next(unsigned nsb, unsigned es) {
    c += SC;
    saddr += SR * nsb + SC * es;
    if (c >= EPR) { c -= EPR; saddr += nsb - EPR * es; }
  }
};

__device__ __forceinline__ int4 lds128_u32(unsigned saddr) {
  int4 v;
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];\n" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
  return v;
}

// Chunk walk for units of CPU 16-byte chunks when neither 32 % CPU nor CPU % 32 is zero (packed lower triangles: CPU = 9), one unit per
// record: fully unrolled over the at most MAXU units of a tile, so that a chunk costs one compare + select for the record wrap and
// immediates for everything else (the generic path divides per chunk).  f(saddr, j, wrap): lane's chunk number lane + 32 j.
// Used only where a caller asks for it (template flag WALK of the rec_* functions: the packed-triangle fields).  Measured on B200: for the
// 25-chunk blocks of float32 ell = 10 the unrolled walk cost registers that kernel does not have (14.9 -> 25.4 ms per step) and it is neutral for
// the 9- and 18-chunk full blocks of ell = 6, so every other field keeps the generic loop.
constexpr int kMaxTileUnits = 33;
template <int CPU, int MAXU, typename F>
__device__ __forceinline__ void chunk_walk(unsigned srec0, unsigned nsb, int kstart, int nunits, F f) {
  const unsigned lane = threadIdx.x & 31;
  const unsigned k0 = lane + (unsigned)kstart * CPU;
  const unsigned q0 = k0 / CPU, c0 = k0 - q0 * CPU;
  const unsigned base = srec0 + q0 * nsb + c0 * 16;
  const int total = nunits * CPU;
  constexpr int J = (MAXU * CPU + 31) / 32;
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const unsigned a = (32u * j) / CPU, b = (32u * j) % CPU;
    const bool wrap = (b != 0) && (c0 >= CPU - b);
    const unsigned saddr = base + a * nsb + b * 16 + (wrap ? nsb - CPU * 16 : 0u);
    if ((int)lane + 32 * j < total) f(saddr, j, wrap);
  }
}

// Global units (UE elements each, contiguous) -> records.  Units are grouped GRP per record
// (GRP = 1: one unit per record; GRP = 2: units 2t, 2t+1 are adjacent fields of record t).
// Unit index k = kstart + j goes to record k / GRP, sub-field k % GRP.
// `srec0` = shared address of the field in record 0, `nsb` = record stride in bytes.
template <typename T, int UE, int GRP, bool WALK = false>
__device__ __forceinline__ void rec_g2s(unsigned srec0, unsigned nsb, const T* __restrict__ g, int kstart, int nunits, bool vec_ok) {
  constexpr int VE = 16 / (int)sizeof(T);
  if (nunits <= 0) return;
  const int lane = threadIdx.x & 31;
  if constexpr ((UE % VE) == 0) {
    if (vec_ok) {
      constexpr int CPU = UE / VE;          // chunks per unit
      constexpr int CPR = CPU * GRP;        // chunks per record (for this field group)
      const int total = nunits * CPU;
      const char* gp = reinterpret_cast<const char*>(g) + lane * 16;
      if constexpr ((32 % CPR) == 0 || (CPR % 32) == 0) {
        // a warp step of 32 chunks advances by a whole number of records (or stays inside one): the shared
        // address is a plain induction variable -> 2 integer adds per 16-byte chunk
        const unsigned k0 = (unsigned)(lane + kstart * CPU);
        unsigned rec = k0 / CPR, c = k0 - rec * CPR;
        unsigned saddr = srec0 + rec * nsb + c * 16;
        if constexpr ((32 % CPR) == 0) {
          const unsigned step = (32 / CPR) * nsb;
          for (int i = lane; i < total; i += 32, gp += 512, saddr += step) cp_async16_u32(saddr, gp);
        } else {
          // CPR is a multiple of 32: stay in the record for CPR/32 steps, then jump to the next one
          constexpr unsigned SPR = CPR / 32;
          unsigned j = c / 32;
          for (int i = lane; i < total; i += 32, gp += 512) {
            cp_async16_u32(saddr, gp);
            if (++j == SPR) { j = 0; saddr += nsb - (SPR - 1) * 512; } else { saddr += 512; }
          }
        }
        return;
      }
      if constexpr (WALK && GRP == 1 && CPR < 32) if (nunits <= kMaxTileUnits) {
        chunk_walk<CPR, kMaxTileUnits>(srec0, nsb, kstart, nunits, [&](unsigned saddr, int j, bool) { cp_async16_u32(saddr, gp + 512 * j); });
        return;
      }
      for (int i = lane; i < total; i += 32, gp += 512) {
        const unsigned k = (unsigned)(i + kstart * CPU);
        const unsigned rec = k / CPR, c = k - rec * CPR;
        cp_async16_u32(srec0 + rec * nsb + c * 16, gp);
      }
      return;
    }
  }
  constexpr int EPR = UE * GRP;
  const int total = nunits * UE;
  ElemWalk<EPR> w(srec0, nsb, (unsigned)sizeof(T), (unsigned)(kstart * UE));
  const T* gp = g + lane;
  for (int i = lane; i < total; i += 32, gp += 32, w.next(nsb, (unsigned)sizeof(T))) cp_async_elem_u32<T>(w.saddr, gp);
}

// Every GS-th global unit -> one unit per record (unit j = global unit j * GS lands in record kstart + j).
// Used to stage only the even (or only the odd) rows of a level.
template <typename T, int UE, int GS, bool WALK = false>
__device__ __forceinline__ void rec_g2s_strided(unsigned srec0, unsigned nsb, const T* __restrict__ g, int kstart, int nunits, bool vec_ok) {
  constexpr int VE = 16 / (int)sizeof(T);
  if (nunits <= 0) return;
  const int lane = threadIdx.x & 31;
  if constexpr ((UE % VE) == 0) {
    if (vec_ok) {
      constexpr int CPU = UE / VE;
      const int total = nunits * CPU;
      if constexpr ((32 % CPU) == 0) {
        const unsigned j0 = (unsigned)lane / CPU, c = (unsigned)lane - j0 * CPU;
        unsigned saddr = srec0 + (kstart + j0) * nsb + c * 16;
        const char* gp = reinterpret_cast<const char*>(g) + ((size_t)j0 * GS * CPU + c) * 16;
        const unsigned sstep = (32 / CPU) * nsb;
        constexpr size_t gstep = (size_t)(32 / CPU) * GS * CPU * 16;
        for (int i = lane; i < total; i += 32, gp += gstep, saddr += sstep) cp_async16_u32(saddr, gp);
        return;
      }
      if constexpr (WALK && CPU < 32) if (nunits <= kMaxTileUnits) {
        // unit u = q0 + a + wrap, chunk c = c0 + b - CPU * wrap: the global side skips (GS - 1) units at every wrap
        const unsigned q0 = (unsigned)lane / CPU, c0 = (unsigned)lane - q0 * CPU;
        const char* g0 = reinterpret_cast<const char*>(g) + ((size_t)q0 * GS * CPU + c0) * 16;
        const char* g1 = g0 + (size_t)(GS - 1) * CPU * 16;
        chunk_walk<CPU, kMaxTileUnits>(srec0 + kstart * nsb, nsb, 0, nunits, [&](unsigned saddr, int j, bool wrap) {
          const unsigned a = (32u * j) / CPU, b = (32u * j) % CPU;
          cp_async16_u32(saddr, (wrap ? g1 : g0) + ((size_t)a * GS * CPU + b) * 16);
        });
        return;
      }
      for (int i = lane; i < total; i += 32) {
        const unsigned j = (unsigned)i / CPU, c = (unsigned)i - j * CPU;
        cp_async16_u32(srec0 + (kstart + j) * nsb + c * 16, reinterpret_cast<const char*>(g) + ((size_t)j * GS * CPU + c) * 16);
      }
      return;
    }
  }
  // one unit per record: the global side skips (GS - 1) units whenever the walk moves to the next record
  const int total = nunits * UE;
  ElemWalk<UE> w(srec0 + kstart * nsb, nsb, (unsigned)sizeof(T), 0u);
  const unsigned j0 = (unsigned)lane / UE;
  const T* gp = g + (size_t)j0 * GS * UE + ((unsigned)lane - j0 * UE);
  constexpr unsigned SR = 32 / UE, SC = 32 % UE;
  for (int i = lane; i < total; i += 32) {
    cp_async_elem_u32<T>(w.saddr, gp);
    const unsigned cb = w.c;
    w.next(nsb, (unsigned)sizeof(T));
    gp += (size_t)SR * GS * UE + SC + ((cb + SC >= (unsigned)UE) ? (size_t)(GS - 1) * UE : 0);
  }
}

// one row of L elements from registers straight to global memory (16-byte stores when aligned)
template <typename T, int L>
__device__ __forceinline__ void stg_row(T* p, const T (&a)[L], bool vec_ok) {
  constexpr int VE = 16 / (int)sizeof(T);
  if constexpr ((L % VE) == 0) {
    if (vec_ok) {
#pragma unroll
      for (int c = 0; c < L; c += VE) {
        if constexpr (sizeof(T) == 4) *reinterpret_cast<float4*>(p + c) = make_float4(a[c], a[c + 1], a[c + 2], a[c + 3]);
        else *reinterpret_cast<double2*>(p + c) = make_double2(a[c], a[c + 1]);
      }
      return;
    }
  }
  if constexpr (sizeof(T) == 4 && (L % 2) == 0) {
    if (vec_ok) {                       // 8-byte stores: rows start at multiples of ell elements from a 16-byte aligned base
#pragma unroll
      for (int c = 0; c < L; c += 2) *reinterpret_cast<float2*>(p + c) = make_float2(a[c], a[c + 1]);
      return;
    }
  }
#pragma unroll
  for (int c = 0; c < L; ++c) p[c] = a[c];
}

// Records -> global units, same addressing.
template <typename T, int UE, int GRP, bool WALK = false>
__device__ __forceinline__ void rec_s2g(T* __restrict__ g, unsigned srec0, unsigned nsb, int kstart, int nunits, bool vec_ok) {
  constexpr int VE = 16 / (int)sizeof(T);
  if (nunits <= 0) return;
  const int lane = threadIdx.x & 31;
  if constexpr ((UE % VE) == 0) {
    if (vec_ok) {
      constexpr int CPU = UE / VE;
      constexpr int CPR = CPU * GRP;
      const int total = nunits * CPU;
      char* gp = reinterpret_cast<char*>(g) + lane * 16;
      if constexpr ((32 % CPR) == 0 || (CPR % 32) == 0) {
        const unsigned k0 = (unsigned)(lane + kstart * CPU);
        unsigned rec = k0 / CPR, c = k0 - rec * CPR;
        unsigned saddr = srec0 + rec * nsb + c * 16;
        if constexpr ((32 % CPR) == 0) {
          const unsigned step = (32 / CPR) * nsb;
          for (int i = lane; i < total; i += 32, gp += 512, saddr += step) *reinterpret_cast<int4*>(gp) = lds128_u32(saddr);
        } else {
          constexpr unsigned SPR = CPR / 32;
          unsigned j = c / 32;
          for (int i = lane; i < total; i += 32, gp += 512) {
            *reinterpret_cast<int4*>(gp) = lds128_u32(saddr);
            if (++j == SPR) { j = 0; saddr += nsb - (SPR - 1) * 512; } else { saddr += 512; }
          }
        }
        return;
      }
      if constexpr (WALK && GRP == 1 && CPR < 32) if (nunits <= kMaxTileUnits) {
        chunk_walk<CPR, kMaxTileUnits>(srec0, nsb, kstart, nunits, [&](unsigned saddr, int j, bool) { *reinterpret_cast<int4*>(gp + 512 * j) = lds128_u32(saddr); });
        return;
      }
      for (int i = lane; i < total; i += 32, gp += 512) {
        const unsigned k = (unsigned)(i + kstart * CPU);
        const unsigned rec = k / CPR, c = k - rec * CPR;
        *reinterpret_cast<int4*>(gp) = lds128_u32(srec0 + rec * nsb + c * 16);
      }
      return;
    }
  }
  constexpr int EPR = UE * GRP;
  const int total = nunits * UE;
  ElemWalk<EPR> w(srec0, nsb, (unsigned)sizeof(T), (unsigned)(kstart * UE));
  T* gp = g + lane;
  for (int i = lane; i < total; i += 32, gp += 32, w.next(nsb, (unsigned)sizeof(T))) *gp = lds_elem_u32<T>(w.saddr);
}

// One unit per record -> every GS-th global unit (record kstart + j -> global unit j * GS): the mirror of
// rec_g2s_strided, used to write only the odd rows of an interleaved output.
template <typename T, int UE, int GS, bool WALK = false>
__device__ __forceinline__ void rec_s2g_strided(T* __restrict__ g, unsigned srec0, unsigned nsb, int kstart, int nunits, bool vec_ok) {
  constexpr int VE = 16 / (int)sizeof(T);
  if (nunits <= 0) return;
  const int lane = threadIdx.x & 31;
  if constexpr ((UE % VE) == 0) {
    if (vec_ok) {
      constexpr int CPU = UE / VE;
      const int total = nunits * CPU;
      if constexpr ((32 % CPU) == 0) {
        const unsigned j0 = (unsigned)lane / CPU, c = (unsigned)lane - j0 * CPU;
        unsigned saddr = srec0 + (kstart + j0) * nsb + c * 16;
        char* gp = reinterpret_cast<char*>(g) + ((size_t)j0 * GS * CPU + c) * 16;
        const unsigned sstep = (32 / CPU) * nsb;
        constexpr size_t gstep = (size_t)(32 / CPU) * GS * CPU * 16;
        for (int i = lane; i < total; i += 32, gp += gstep, saddr += sstep) *reinterpret_cast<int4*>(gp) = lds128_u32(saddr);
        return;
      }
      if constexpr (WALK && CPU < 32) if (nunits <= kMaxTileUnits) {
        const unsigned q0 = (unsigned)lane / CPU, c0 = (unsigned)lane - q0 * CPU;
        char* g0 = reinterpret_cast<char*>(g) + ((size_t)q0 * GS * CPU + c0) * 16;
        char* g1 = g0 + (size_t)(GS - 1) * CPU * 16;
        chunk_walk<CPU, kMaxTileUnits>(srec0 + kstart * nsb, nsb, 0, nunits, [&](unsigned saddr, int j, bool wrap) {
          const unsigned a = (32u * j) / CPU, b = (32u * j) % CPU;
          *reinterpret_cast<int4*>((wrap ? g1 : g0) + ((size_t)a * GS * CPU + b) * 16) = lds128_u32(saddr);
        });
        return;
      }
      for (int i = lane; i < total; i += 32) {
        const unsigned j = (unsigned)i / CPU, c = (unsigned)i - j * CPU;
        *reinterpret_cast<int4*>(reinterpret_cast<char*>(g) + ((size_t)j * GS * CPU + c) * 16) = lds128_u32(srec0 + (kstart + j) * nsb + c * 16);
      }
      return;
    }
  }
  const int total = nunits * UE;
  ElemWalk<UE> w(srec0 + kstart * nsb, nsb, (unsigned)sizeof(T), 0u);
  const unsigned j0 = (unsigned)lane / UE;
  T* gp = g + (size_t)j0 * GS * UE + ((unsigned)lane - j0 * UE);
  constexpr unsigned SR = 32 / UE, SC = 32 % UE;
  for (int i = lane; i < total; i += 32) {
    *gp = lds_elem_u32<T>(w.saddr);
    const unsigned cb = w.c;
    w.next(nsb, (unsigned)sizeof(T));
    gp += (size_t)SR * GS * UE + SC + ((cb + SC >= (unsigned)UE) ? (size_t)(GS - 1) * UE : 0);
  }
}

// Boundary lanes (no node, no odd neighbour, no left link) get neutral operands written into their own
// shared-memory record once, so that the dense algebra needs no per-element selects.
template <typename T, int N>
__device__ __forceinline__ void smem_fill_zero(T* p) {
#pragma unroll
  for (int i = 0; i < N; ++i) p[i] = T(0);
}
template <typename T, int L>
__device__ __forceinline__ void smem_fill_identity(T* p) {
#pragma unroll
  for (int r = 0; r < L; ++r)
#pragma unroll
    for (int c = 0; c < L; ++c) p[r * L + c] = (r == c) ? T(1) : T(0);
}

// ------------------------------------------------------------------------------------------------------------
// Packed lower triangles (crb200_*_args.tri): the symmetric / triangular blocks that never leave the library in the
// fused likelihood path -- D (Cholesky factors), R~ (reduced diagonal blocks) and Sigma_d of the inner levels -- travel
// through HBM as the ell (ell + 1) / 2 entries of their lower triangle, row after row, padded to a multiple of 16 bytes
// (PKS elements per block).  Shared-memory slots keep their full size: the packed payload lands at the start of the
// slot and is expanded / compacted there by the thread that owns the node.  Offered where the thread-per-node kernels
// use the three-blocks-per-node layout and both block sizes are multiples of 16 bytes: ell = 8 in float32, the
// configuration the headline metric is quoted on (36 instead of 64 elements: -14 % bytes over a whole step).
// ------------------------------------------------------------------------------------------------------------
template <typename T, int L>
struct TriPack {
  static constexpr int VE = 16 / (int)sizeof(T);
  static constexpr int PK = L * (L + 1) / 2;
  static constexpr int PKS = (PK + VE - 1) / VE * VE;
  static constexpr bool OK = tri_stride_elems((int)sizeof(T), L) != 0;
  static_assert(!OK || (tri_stride_elems((int)sizeof(T), L) == PKS && (L * L) % VE == 0 && L % VE == 0), "packed blocks and full blocks must be 16-byte multiples");
};
__host__ __device__ constexpr int tri_index(int r, int c) { return r * (r + 1) / 2 + c; }

// packed block in shared memory -> flat registers
template <typename T, int L>
__device__ __forceinline__ void lds_tri_flat(T (&f)[TriPack<T, L>::PKS], const T* p) {
  lds_row<T, TriPack<T, L>::PKS>(f, p);
}
// packed block in shared memory -> lower triangle of M (entries above the diagonal are left alone)
template <typename T, int L>
__device__ __forceinline__ void lds_tri(T (&M)[L][L], const T* p) {
  T f[TriPack<T, L>::PKS];
  lds_tri_flat<T, L>(f, p);
#pragma unroll
  for (int r = 0; r < L; ++r)
#pragma unroll
    for (int c = 0; c <= r; ++c) M[r][c] = f[tri_index(r, c)];
}
// lower triangle of M -> packed block (shared or global memory, 16-byte aligned)
template <typename T, int L>
__device__ __forceinline__ void st_tri(T* p, const T (&M)[L][L]) {
  T f[TriPack<T, L>::PKS];
#pragma unroll
  for (int i = TriPack<T, L>::PK; i < TriPack<T, L>::PKS; ++i) f[i] = T(0);
#pragma unroll
  for (int r = 0; r < L; ++r)
#pragma unroll
    for (int c = 0; c <= r; ++c) f[tri_index(r, c)] = M[r][c];
  sts_row<T, TriPack<T, L>::PKS>(p, f);
}
// in place, one block owned by the calling thread: packed payload at the start of the slot -> full symmetric L x L rows
template <typename T, int L>
__device__ __forceinline__ void tri_expand_inplace(T* p, T* copy_to = nullptr) {
  T f[TriPack<T, L>::PKS];
  lds_tri_flat<T, L>(f, p);
  if (copy_to != nullptr) sts_row<T, TriPack<T, L>::PKS>(copy_to, f);      // the packed block passes through (16-byte stores, any state space)
#pragma unroll
  for (int r = 0; r < L; ++r) {
    T row[L];
#pragma unroll
    for (int c = 0; c < L; ++c) row[c] = (c <= r) ? f[tri_index(r, c)] : f[tri_index(c, r)];
    sts_row<T, L>(p + r * L, row);
  }
}
// The same for ONE block by the whole warp (a record no lane owns): lane j < 4 L gathers 16 bytes of the full block, the warp
// syncs, the lanes store.
template <typename T, int L>
__device__ __forceinline__ void tri_expand_inplace_warp(T* p, const int lane) {
  constexpr int VE = TriPack<T, L>::VE;
  static_assert(L * L / VE <= 32, "one 16-byte chunk of the full block per lane");
  T v[VE];
  const int r = (lane * VE) / L, c0 = (lane * VE) % L;
  if (lane < L * L / VE) {
#pragma unroll
    for (int i = 0; i < VE; ++i) {
      const int c = c0 + i;
      v[i] = p[c <= r ? tri_index(r, c) : tri_index(c, r)];
    }
  }
  __syncwarp();
  if (lane < L * L / VE) sts_row<T, VE>(p + lane * VE, v);
}
// ... and back: lower triangle of the full rows -> packed payload at the start of the slot
template <typename T, int L>
__device__ __forceinline__ void tri_compact_inplace(T* p) {
  T f[TriPack<T, L>::PKS];
#pragma unroll
  for (int i = TriPack<T, L>::PK; i < TriPack<T, L>::PKS; ++i) f[i] = T(0);
#pragma unroll
  for (int r = 0; r < L; ++r) {
    T row[L];
    lds_row<T, L>(row, p + r * L);
#pragma unroll
    for (int c = 0; c <= r; ++c) f[tri_index(r, c)] = row[c];
  }
  sts_row<T, TriPack<T, L>::PKS>(p, f);
}
template <typename T, int L>
__device__ __forceinline__ void smem_fill_identity_tri(T* p) {
#pragma unroll
  for (int r = 0; r < L; ++r)
#pragma unroll
    for (int c = 0; c <= r; ++c) p[tri_index(r, c)] = (r == c) ? T(1) : T(0);
}

// compiler-only fence: stops the scheduler from hoisting later shared-memory loads above this
// point (keeps the register live set of the fully unrolled row loops bounded)
__device__ __forceinline__ void sched_fence() { asm volatile("" ::: "memory"); }

// acc[c] += s * row[c] for a register-resident row.  A packed FFMA2 (fma.rn.f32x2, new on sm_100)
// variant was measured on B200 and is NOT used: it issues at half the rate of FFMA, the aligned register
// pairs cost spills, and the l=8 backward level got 3 % slower (profiles/r1_summary.md).
#ifndef CRB200_USE_FFMA2
#define CRB200_USE_FFMA2 0
#endif
template <typename T, int L>
__device__ __forceinline__ void axpy_row(T (&acc)[L], T s, const T (&row)[L]) {
  if constexpr (CRB200_USE_FFMA2 && sizeof(T) == 4 && (L % 2) == 0) {
    const float2 ss = make_float2(s, s);
#pragma unroll
    for (int c = 0; c < L; c += 2) {
      const float2 r = __ffma2_rn(ss, make_float2(row[c], row[c + 1]), make_float2(acc[c], acc[c + 1]));
      acc[c] = r.x;
      acc[c + 1] = r.y;
    }
  } else {
#pragma unroll
    for (int c = 0; c < L; ++c) acc[c] = fma(s, row[c], acc[c]);
  }
}

// record stride (elements).  Blocks that are a multiple of 16 bytes are accessed with 16-byte vector loads / stores:
// the payload is rounded up to an ODD number of 16-byte chunks, so that the 8 lanes of a quarter warp hit 8 different
// 16-byte bank groups.  Other blocks (odd ell) are accessed element by element: an ODD number of ELEMENTS makes
// the 32 lanes (fp32) / the 16 lanes of a half warp (fp64) hit distinct banks -- with the 16-byte rule those scalar
// accesses were 4-way (fp32) / 2-way (fp64) bank conflicted (ell = 7 fp32 ran slower than ell = 8).
template <typename T>
__host__ __device__ constexpr int record_stride(int payload_elems, int block_elems) {
  return ((block_elems * (int)sizeof(T)) % 16 != 0)
             ? (payload_elems | 1)
             : ((((payload_elems * (int)sizeof(T)) + 15) / 16) | 1) * 16 / (int)sizeof(T);
}

}  // namespace crb200
