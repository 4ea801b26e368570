// Warp-per-node building blocks of the large-block kernel family (cr_mma_fwd.cuh / cr_mma_bwd.cuh):
// fp32 ell >= 11 and fp64 ell >= 9, where a block no longer fits one thread's registers.
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
template <int LP, bool TA, bool TB, int KR, bool LOWER>
__device__ __forceinline__ void warp_gemm(double (&acc)[LP / 8][LP / 8][2], const double* __restrict__ A, const double* __restrict__ B, const int lane) {
  constexpr int LD = LP + 4, NTL = LP / 8;
  const int lr = lane >> 2, lc = lane & 3;
  const double* pa = TA ? A + lc * LD + lr : A + lr * LD + lc;
  const double* pb = TB ? B + lr * LD + lc : B + lc * LD + lr;
#pragma unroll
  for (int k0 = 0; k0 < LP; k0 += 4) {
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
  if (row >= L || col >= L) return;
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
// In-place Cholesky of the SPD block in S (lower triangle read; exact zeros written above the diagonal).
// invd[j] = 1 / K[j][j] on every lane.  Returns true when a pivot was not positive.
template <int LP>
__device__ __forceinline__ bool warp_cholesky(double* S, double (&invd)[LP], const int lane) {
  constexpr int LD = LP + 4;
  const bool act = lane < LP;
  const int r = act ? lane : LP - 1;        // idle lanes shadow the last row and never store
  double Lr[LP];
#pragma unroll
  for (int c = 0; c < LP; c += 2) {
    const double2 v = *reinterpret_cast<const double2*>(S + r * LD + c);
    Lr[c] = v.x; Lr[c + 1] = v.y;
  }
  bool bad = false;
#pragma unroll
  for (int j = 0; j < LP; ++j) {
    // left-looking: s = A[r][j] - sum_{k<j} L[r][k] L[j][k]; row j of L is complete in shared memory (broadcast reads)
    double s0 = Lr[j], s1 = 0.0;
#pragma unroll
    for (int k = 0; k + 1 < j; k += 2) {
      const double2 lj = *reinterpret_cast<const double2*>(S + j * LD + k);
      s0 = fma(-Lr[k], lj.x, s0);
      s1 = fma(-Lr[k + 1], lj.y, s1);
    }
    if (j & 1) s0 = fma(-Lr[j - 1], S[j * LD + j - 1], s0);
    const double s = s0 + s1;
    const double d = __shfl_sync(0xffffffffu, s, j);
    if (!(d > 0.0)) bad = true;
    const double inv = rsqrt(d);
    invd[j] = inv;
    Lr[j] = (r > j) ? s * inv : (r == j ? d * inv : 0.0);
    if (act) S[r * LD + j] = Lr[j];
    __syncwarp();
  }
  return bad;
}

// S <- S^{-1} for lower-triangular S with 1 / diag in invd; lane c builds column c by forward substitution
template <int LP>
__device__ __forceinline__ void warp_tri_inverse(double* S, const double (&invd)[LP], const int lane) {
  constexpr int LD = LP + 4;
  const int c = lane;
  double z[LP];
#pragma unroll
  for (int r = 0; r < LP; ++r) {
    double s0 = (r == c) ? 1.0 : 0.0, s1 = 0.0;
#pragma unroll
    for (int k = 0; k + 1 < r; k += 2) {
      const double2 kr = *reinterpret_cast<const double2*>(S + r * LD + k);
      s0 = fma(-kr.x, z[k], s0);
      s1 = fma(-kr.y, z[k + 1], s1);
    }
    if (r & 1) s0 = fma(-S[r * LD + r - 1], z[r - 1], s0);
    z[r] = (r >= c) ? (s0 + s1) * invd[r] : 0.0;
  }
  __syncwarp();        // every lane is done reading the factor
  if (c < LP) {
#pragma unroll
    for (int r = 0; r < LP; ++r) S[r * LD + c] = z[r];
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
