// Warp-per-node building blocks of the large-block kernel family (cr_mma_fwd.cuh / cr_mma_bwd.cuh):
// blocks that no longer fit one thread's registers (automatic for fp64 ell >= 10 and fp32 ell >= 17, see cr_inst.cu).
//
// One WARP owns one even node.  Every ell x ell block lives in shared memory as a padded
// LP x LD row-major matrix of DOUBLES (LP = ell rounded up to a multiple of 8, LD = LP + 4), whatever the
// storage type: fp32 data is widened while it is staged and narrowed when it leaves, so the arithmetic type of
// this family is fp64 for both dtypes.  All ell^3 work (triangular solves restated as products with the explicit
// triangular inverse, Schur products, selected-inverse products) runs on the FP64 tensor path,
// mma.sync.m8n8k4.f64 (SASS: DMMA), which on B200 sustains the full FP64 rate with one warp per SM sub-partition
// (profiles/r2_pipe_peaks.json: 18.6 TFMA/s DMMA vs 17.1 DFMA).  The O(ell^2)-per-step serial parts -- Cholesky and
// triangular inverse -- run with lane = row / lane = column and operands broadcast from shared memory.
//
// Fragment addressing.  m8n8k4 wants A[lane/4][lane%4], B[lane%4][lane/4] and gives C[lane/4][2(lane%4)+{0,1}].
// With LD = LP + 4 (LD mod 16 in {4, 12}) the 16 lanes of a half warp hit 16 different 8-byte bank pairs for any of
// the four transposition combinations, so operands are read straight from the row-major blocks, transposed or not.
#pragma once
#include "cr_common.cuh"

namespace crb200 {

template <int L>
struct MmaGeom {
  static constexpr int LP = (L + 7) / 8 * 8;     // padded block size
  static constexpr int LD = LP + 4;              // leading dimension (doubles)
  static constexpr int BLK = LP * LD;            // doubles per block
  static constexpr int NTL = LP / 8;             // 8 x 8 tiles per side
  static constexpr int KP = (L + 3) / 4 * 4;     // inner dimension of the products: the padding beyond it is zero (or a unit diagonal against zeros)
  static constexpr int LA = (L + 1) / 2 * 2;     // active size of the serial parts (Cholesky, triangular inverse): the padding is the identity
};

__device__ __forceinline__ void dmma884(double& c0, double& c1, const double a, const double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// which k (inner index) ranges contribute to output tile (mt, nt); lets the products skip the zero halves of
// triangular operands at tile granularity
enum KRange {
  K_FULL = 0,
  K_LE_N = 1,       // inner tile <= column tile: second operand is (lower-triangular)^T
  K_GE_N = 2,       // inner tile >= column tile: second operand is lower-triangular
  K_GE_MAX_MN = 3   // both operands lower-triangular, first one transposed (Di^T Di)
};

template <int KR>
__device__ __forceinline__ constexpr bool k_active(int kt, int mt, int nt) {
  return KR == K_FULL ? true : KR == K_LE_N ? (kt <= nt) : KR == K_GE_N ? (kt >= nt) : (kt >= (mt > nt ? mt : nt));
}

// acc (+)= op(A) op(B) over the whole LP x LP block by ONE warp.
//   op(A)[m][k] = TA ? A[k][m] : A[m][k]        op(B)[k][n] = TB ? B[n][k] : B[k][n]
// acc[mt][nt][0..1] is the C fragment of tile (mt, nt).  LOWER: only tiles with mt >= nt are computed.
// KP: inner dimension actually summed (a multiple of 4, >= ell): blocks are zero beyond ell -- or carry a unit diagonal that only ever
// meets zeros there -- so the k-steps over the padding contribute nothing and are skipped.
template <int LP, bool TA, bool TB, int KR, bool LOWER, int KP = LP>
__device__ __forceinline__ void warp_gemm(double (&acc)[LP / 8][LP / 8][2], const double* __restrict__ A, const double* __restrict__ B, const int lane) {
  constexpr int LD = LP + 4, NTL = LP / 8;
  const int lr = lane >> 2, lc = lane & 3;
  const double* pa = TA ? A + lc * LD + lr : A + lr * LD + lc;
  const double* pb = TB ? B + lr * LD + lc : B + lc * LD + lr;
#pragma unroll
  for (int k0 = 0; k0 < KP; k0 += 4) {
    const int kt = k0 >> 3;
    double af[NTL], bf[NTL];
#pragma unroll
    for (int mt = 0; mt < NTL; ++mt) af[mt] = TA ? pa[k0 * LD + mt * 8] : pa[mt * 8 * LD + k0];
#pragma unroll
    for (int nt = 0; nt < NTL; ++nt) bf[nt] = TB ? pb[nt * 8 * LD + k0] : pb[k0 * LD + nt * 8];
#pragma unroll
    for (int mt = 0; mt < NTL; ++mt)
#pragma unroll
      for (int nt = 0; nt < NTL; ++nt)
        if ((!LOWER || mt >= nt) && k_active<KR>(kt, mt, nt)) dmma884(acc[mt][nt][0], acc[mt][nt][1], af[mt], bf[nt]);
  }
}

template <int LP>
__device__ __forceinline__ void acc_zero(double (&acc)[LP / 8][LP / 8][2]) {
#pragma unroll
  for (int mt = 0; mt < LP / 8; ++mt)
#pragma unroll
    for (int nt = 0; nt < LP / 8; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
}

// C fragments -> padded shared-memory block, scaled by `s` (16-byte stores)
template <int LP>
__device__ __forceinline__ void acc_to_smem(double* S, const double (&acc)[LP / 8][LP / 8][2], const double s, const int lane) {
  constexpr int LD = LP + 4;
  const int lr = lane >> 2, lc = lane & 3;
#pragma unroll
  for (int mt = 0; mt < LP / 8; ++mt)
#pragma unroll
    for (int nt = 0; nt < LP / 8; ++nt)
      *reinterpret_cast<double2*>(S + (mt * 8 + lr) * LD + nt * 8 + 2 * lc) = make_double2(s * acc[mt][nt][0], s * acc[mt][nt][1]);
}

// mirror the strictly-lower tiles of a symmetric result (computed with LOWER) into the upper tiles, through the
// shared-memory copy that acc_to_smem has just written (caller syncs the warp in between)
template <int LP>
__device__ __forceinline__ void acc_mirror_from_smem(double (&acc)[LP / 8][LP / 8][2], const double* S, const double s, const int lane) {
  constexpr int LD = LP + 4;
  const int lr = lane >> 2, lc = lane & 3;
#pragma unroll
  for (int mt = 0; mt < LP / 8; ++mt)
#pragma unroll
    for (int nt = 0; nt < LP / 8; ++nt)
      if (nt > mt) {     // C[row][col] = C[col][row], the latter sits in the stored lower tile (nt, mt)
        acc[mt][nt][0] = s * S[(nt * 8 + 2 * lc) * LD + mt * 8 + lr];
        acc[mt][nt][1] = s * S[(nt * 8 + 2 * lc + 1) * LD + mt * 8 + lr];
      }
}

// one pair (row, col), (row, col + 1) of a block -> global storage type T, dense ell x ell rows
template <typename T, int L>
__device__ __forceinline__ void frag_pair_store(T* __restrict__ g, const int row, const int col, const double v0, const double v1, const bool vec_ok) {
  if constexpr (L % 8 != 0) {           // padded block: rows / columns >= L do not exist in global memory
    if (row >= L || col >= L) return;
  }
  T* p = g + row * L + col;
  if constexpr ((L % 2) == 0) {
    if (vec_ok) {
      if constexpr (sizeof(T) == 8) *reinterpret_cast<double2*>(p) = make_double2(v0, v1);
      else *reinterpret_cast<float2*>(p) = make_float2((float)v0, (float)v1);
      return;
    }
  }
  p[0] = (T)v0;
  if (col + 1 < L) p[1] = (T)v1;
}

// C fragments (scaled) -> a dense ell x ell global block
template <typename T, int L, int LP>
__device__ __forceinline__ void acc_to_global(T* __restrict__ g, const double (&acc)[LP / 8][LP / 8][2], const double s, const int lane, const bool vec_ok) {
  const int lr = lane >> 2, lc = lane & 3;
#pragma unroll
  for (int mt = 0; mt < LP / 8; ++mt)
#pragma unroll
    for (int nt = 0; nt < LP / 8; ++nt)
      frag_pair_store<T, L>(g, mt * 8 + lr, nt * 8 + 2 * lc, s * acc[mt][nt][0], s * acc[mt][nt][1], vec_ok);
}

// ---------------------------------------------------------------------------------------------------------
// staging: dense ell x ell global block (type T) <-> padded LP x LD block of doubles, one warp
// ---------------------------------------------------------------------------------------------------------
// zero the padding (rows / columns >= L) of a block; `identity` puts ones on the padded diagonal
template <int L, int LP>
__device__ __forceinline__ void mma_pad_block(double* S, const bool identity, const int lane) {
  constexpr int LD = LP + 4;
  if constexpr (L < LP) {
    for (int i = lane; i < LP * LP; i += 32) {
      const int r = i / LP, c = i - r * LP;
      if (r >= L || c >= L) S[r * LD + c] = (identity && r == c) ? 1.0 : 0.0;
    }
  }
}

template <int LP>
__device__ __forceinline__ void mma_fill_block(double* S, const bool identity, const int lane) {
  constexpr int LD = LP + 4;
  for (int i = lane; i < LP * LP; i += 32) {
    const int r = i / LP, c = i - r * LP;
    S[r * LD + c] = (identity && r == c) ? 1.0 : 0.0;
  }
}

// global -> shared.  fp64 goes by cp.async (the caller waits with cp_async_wait_all + __syncwarp); fp32 is loaded,
// widened and stored by the lanes (several loads in flight per lane).
template <typename T, int L, int LP>
__device__ __forceinline__ void mma_stage_block(double* S, const T* __restrict__ g, const int lane, const bool vec_ok) {
  constexpr int LD = LP + 4;
  if constexpr (sizeof(T) == 8) {
    if constexpr ((L % 2) == 0) {
      if (vec_ok) {
        constexpr int CPR = L / 2, TOT = L * CPR;
        if constexpr ((32 % CPR) == 0 && (TOT % 32) == 0) {
          // a warp step of 32 chunks covers 32 / CPR whole rows: both addresses are plain induction variables
          const int r0 = lane / CPR, c0 = lane - r0 * CPR;
          unsigned sa = smem_u32(S + r0 * LD + 2 * c0);
          const char* gp = reinterpret_cast<const char*>(g) + (size_t)lane * 16;
          constexpr unsigned sstep = (32 / CPR) * LD * 8;
#pragma unroll
          for (int i = 0; i < TOT / 32; ++i, sa += sstep, gp += 512)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gp) : "memory");
          return;
        }
        for (int i = lane; i < TOT; i += 32) {
          const int r = i / CPR, c = i - r * CPR;
          cp_async16(S + r * LD + 2 * c, g + r * L + 2 * c);
        }
        return;
      }
    }
    for (int i = lane; i < L * L; i += 32) {
      const int r = i / L, c = i - r * L;
      cp_async8(S + r * LD + c, g + i);
    }
  } else {
    if constexpr ((L % 4) == 0) {
      if (vec_ok) {
        constexpr int CPR = L / 4, TOT = L * CPR;
#pragma unroll 4
        for (int i = lane; i < TOT; i += 32) {
          const int r = i / CPR, c = i - r * CPR;
          const float4 v = __ldg(reinterpret_cast<const float4*>(g + r * L + 4 * c));
          double* d = S + r * LD + 4 * c;
          *reinterpret_cast<double2*>(d) = make_double2((double)v.x, (double)v.y);
          *reinterpret_cast<double2*>(d + 2) = make_double2((double)v.z, (double)v.w);
        }
        return;
      }
    }
#pragma unroll 4
    for (int i = lane; i < L * L; i += 32) {
      const int r = i / L, c = i - r * L;
      S[r * LD + c] = (double)__ldg(g + i);
    }
  }
}

// fp32 storage, asynchronous: cp.async the raw ell x ell floats into the TAIL of the block's own shared memory
// (mma_stage_raw_f32), and after the copy has landed widen them in place (mma_widen_f32: every lane first reads all its
// elements into registers, then the warp writes the padded doubles -- the two regions overlap).
template <int L, int LP>
__device__ __forceinline__ float* mma_raw_f32(double* S) {
  constexpr int RAWF = (L * L + 3) / 4 * 4;
  return reinterpret_cast<float*>(S + LP * (LP + 4)) - RAWF;
}
template <int L, int LP>
__device__ __forceinline__ void mma_stage_raw_f32(double* S, const float* __restrict__ g, const int lane, const bool vec_ok) {
  float* raw = mma_raw_f32<L, LP>(S);
  if constexpr ((L * L) % 4 == 0) {
    if (vec_ok) {
      for (int i = lane; i < L * L / 4; i += 32) cp_async16(raw + 4 * i, g + 4 * i);
      return;
    }
  }
  for (int i = lane; i < L * L; i += 32) cp_async4(raw + i, g + i);
}
template <int L, int LP>
__device__ __forceinline__ void mma_widen_f32(double* S, const bool identity, const int lane) {
  constexpr int LD = LP + 4, PER = (L * L + 31) / 32;
  const float* raw = mma_raw_f32<L, LP>(S);
  float v[PER];
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    const int i = lane + 32 * k;
    v[k] = i < L * L ? raw[i] : 0.f;
  }
  __syncwarp();
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    const int i = lane + 32 * k;
    if (i < L * L) {
      const int r = i / L, c = i - r * L;
      S[r * LD + c] = (double)v[k];
    }
  }
  mma_pad_block<L, LP>(S, identity, lane);
}

// one call per block for either storage type: issue the copy now ...
template <typename T, int L, int LP>
__device__ __forceinline__ void mma_stage_issue(double* S, const T* __restrict__ g, const bool identity, const int lane, const bool vec_ok) {
  if constexpr (sizeof(T) == 8) {
    mma_pad_block<L, LP>(S, identity, lane);
    mma_stage_block<T, L, LP>(S, g, lane, vec_ok);
  } else {
    mma_stage_raw_f32<L, LP>(S, g, lane, vec_ok);
  }
}
// ... and finish it after cp_async_wait_all + __syncwarp (fp32: widen in place; fp64: nothing left to do)
template <typename T, int L, int LP>
__device__ __forceinline__ void mma_stage_finish(double* S, const bool identity, const int lane) {
  if constexpr (sizeof(T) == 4) mma_widen_f32<L, LP>(S, identity, lane);
}

// CTA-internal producer / consumer hand-over between two warps on a named barrier (ids 1..15)
__device__ __forceinline__ void pair_arrive(const int id) {
  __threadfence_block();
  asm volatile("bar.arrive %0, 64;\n" ::"r"(id) : "memory");
}
__device__ __forceinline__ void pair_wait(const int id) { asm volatile("bar.sync %0, 64;\n" ::"r"(id) : "memory"); }

template <typename T, int L>
__device__ __forceinline__ void mma_stage_vec(double* S, const T* __restrict__ g, const int lane) {
  for (int i = lane; i < L; i += 32) S[i] = (double)g[i];
}

// shared -> global (dense ell x ell, type T), optionally through an element transform f(r, c, v)
template <typename T, int L, int LP, typename F>
__device__ __forceinline__ void mma_store_block_f(T* __restrict__ g, const double* S, const int lane, const bool vec_ok, F f) {
  constexpr int LD = LP + 4;
  if constexpr (sizeof(T) == 8 && (L % 2) == 0) {
    if (vec_ok) {
      constexpr int CPR = L / 2, TOT = L * CPR;
      if constexpr ((32 % CPR) == 0 && (TOT % 32) == 0) {
        const int r0 = lane / CPR, c0 = 2 * (lane - r0 * CPR);
        const double* sp = S + r0 * LD + c0;
        double2* gp = reinterpret_cast<double2*>(g) + lane;
#pragma unroll
        for (int i = 0; i < TOT / 32; ++i, sp += (32 / CPR) * LD, gp += 32) {
          const double2 v = *reinterpret_cast<const double2*>(sp);
          const int r = r0 + i * (32 / CPR);
          *gp = make_double2(f(r, c0, v.x), f(r, c0 + 1, v.y));
        }
        return;
      }
      for (int i = lane; i < TOT; i += 32) {
        const int r = i / CPR, c = 2 * (i - r * CPR);
        const double2 v = *reinterpret_cast<const double2*>(S + r * LD + c);
        *reinterpret_cast<double2*>(g + r * L + c) = make_double2(f(r, c, v.x), f(r, c + 1, v.y));
      }
      return;
    }
  }
  if constexpr (sizeof(T) == 4 && (L % 4) == 0) {
    if (vec_ok) {
      constexpr int CPR = L / 4, TOT = L * CPR;
      for (int i = lane; i < TOT; i += 32) {
        const int r = i / CPR, c = 4 * (i - r * CPR);
        const double2 v0 = *reinterpret_cast<const double2*>(S + r * LD + c);
        const double2 v1 = *reinterpret_cast<const double2*>(S + r * LD + c + 2);
        *reinterpret_cast<float4*>(g + r * L + c) =
            make_float4((float)f(r, c, v0.x), (float)f(r, c + 1, v0.y), (float)f(r, c + 2, v1.x), (float)f(r, c + 3, v1.y));
      }
      return;
    }
  }
  for (int i = lane; i < L * L; i += 32) {
    const int r = i / L, c = i - r * L;
    g[i] = (T)f(r, c, S[r * LD + c]);
  }
}

template <typename T, int L, int LP>
__device__ __forceinline__ void mma_store_block(T* __restrict__ g, const double* S, const int lane, const bool vec_ok) {
  mma_store_block_f<T, L, LP>(g, S, lane, vec_ok, [](int, int, double v) { return v; });
}

// ---------------------------------------------------------------------------------------------------------
// serial parts, one warp, lane = row (Cholesky) / lane = column (inverse)
// ---------------------------------------------------------------------------------------------------------
// In-place Cholesky of the SPD block in S (lower triangle read; exact zeros written above the diagonal), right-looking:
// lane r keeps row r in registers; once column j is final it is published through a small double-buffered vector
// (`colbuf`, 2 * LP doubles) and every lane applies the rank-1 update to the rest of its row -- LP - j independent
// FMAs, so the serial chain per column is only pivot broadcast -> rsqrt -> scale -> publish.
// invd[j] = 1 / K[j][j] on every lane.  Returns true when a pivot was not positive.
// LA (even, >= ell): the block is the identity beyond it, so columns j >= LA need no work (invd = 1) and row updates stop at LA.
template <int LP, int LA = LP>
__device__ __forceinline__ bool warp_cholesky(double* S, double* colbuf, double (&invd)[LP], const int lane) {
  constexpr int LD = LP + 4;
  const bool act = lane < LP;
  const int r = act ? lane : LP - 1;        // idle lanes shadow the last row and never store
  double a[LP];
#pragma unroll
  for (int c = 0; c < LP; c += 2) {
    const double2 v = *reinterpret_cast<const double2*>(S + r * LD + c);
    a[c] = v.x; a[c + 1] = v.y;
  }
  bool bad = false;
#pragma unroll
  for (int j = LA; j < LP; ++j) invd[j] = 1.0;
#pragma unroll
  for (int j = 0; j < LA; ++j) {
    const double d = __shfl_sync(0xffffffffu, a[j], j);
    if (!(d > 0.0)) bad = true;
    const double inv = rsqrt(d);
    invd[j] = inv;
    const double l = (r > j) ? a[j] * inv : (r == j ? d * inv : 0.0);
    a[j] = l;
    double* cb = colbuf + (j & 1) * LP;
    if (act) cb[r] = l;
    __syncwarp();
    // a[c] -= L[r][j] L[c][j] for c > j (entries with c > r are never used)
    if (((j + 1) & 1) != 0 && j + 1 < LA) a[j + 1] = fma(-l, cb[j + 1], a[j + 1]);
#pragma unroll
    for (int c = (j + 2) & ~1; c < LA; c += 2) {
      const double2 t = *reinterpret_cast<const double2*>(cb + c);
      a[c] = fma(-l, t.x, a[c]);
      a[c + 1] = fma(-l, t.y, a[c + 1]);
    }
  }
  if (act) {
#pragma unroll
    for (int c = 0; c < LP; c += 2)
      *reinterpret_cast<double2*>(S + r * LD + c) = make_double2(c <= r ? a[c] : 0.0, c + 1 <= r ? a[c + 1] : 0.0);
  }
  __syncwarp();
  return bad;
}

// S <- S^{-1} for lower-triangular S with 1 / diag in invd.  Lane r builds ROW r of the inverse by a column sweep from
// the right (y^T K = e_r^T): y_c = acc_c / K_cc, then acc_c' -= y_c K[c][c'] for c' < c -- the updates of one step are
// independent of each other and read row c of K as a broadcast, so the serial chain per step is one multiply + one FMA.
template <int LP, int LA = LP>
__device__ __forceinline__ void warp_tri_inverse(double* S, const double (&invd)[LP], const int lane) {
  constexpr int LD = LP + 4;
  const bool act = lane < LP;
  const int r = act ? lane : LP - 1;
  double acc[LP];
#pragma unroll
  for (int c = 0; c < LP; ++c) acc[c] = (c == r) ? 1.0 : 0.0;
#pragma unroll
  for (int c = LA - 1; c >= 0; --c) {       // (columns >= LA: identity rows and columns, nothing to eliminate)
    const double y = acc[c] * invd[c];      // Ki[r][c] (exactly zero for c > r)
    acc[c] = y;
#pragma unroll
    for (int k = 0; k + 1 < c; k += 2) {
      const double2 kc = *reinterpret_cast<const double2*>(S + c * LD + k);
      acc[k] = fma(-y, kc.x, acc[k]);
      acc[k + 1] = fma(-y, kc.y, acc[k + 1]);
    }
    if (c & 1) acc[c - 1] = fma(-y, S[c * LD + c - 1], acc[c - 1]);
  }
  __syncwarp();        // every lane is done reading the factor
  if (act) {
#pragma unroll
    for (int c = 0; c < LP; c += 2) *reinterpret_cast<double2*>(S + r * LD + c) = make_double2(acc[c], acc[c + 1]);
  }
  __syncwarp();
}

// out = M v (TR = false) or M^T v (TR = true) for lane r = output row; lanes >= LP return 0
template <int LP, bool TR>
__device__ __forceinline__ double warp_matvec(const double* M, const double* v, const int lane) {
  constexpr int LD = LP + 4;
  const int r = lane < LP ? lane : 0;
  double s0 = 0.0, s1 = 0.0;
#pragma unroll
  for (int k = 0; k < LP; k += 2) {
    const double2 vk = *reinterpret_cast<const double2*>(v + k);
    if constexpr (TR) {
      s0 = fma(M[k * LD + r], vk.x, s0);
      s1 = fma(M[(k + 1) * LD + r], vk.y, s1);
    } else {
      const double2 mk = *reinterpret_cast<const double2*>(M + r * LD + k);
      s0 = fma(mk.x, vk.x, s0);
      s1 = fma(mk.y, vk.y, s1);
    }
  }
  return lane < LP ? s0 + s1 : 0.0;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

}  // namespace crb200
