// Precision-block builder of the LEG / PEG process on the device, with its own backward.
//
// Replaces LEGFamily.compute_PEG_precision of the reference (cyclic_gps/models.py:181-239; posterior shift of
// compute_posterior_precision :254-268 folded in) and the torch-autograd backward through it.  For every gap
// d_g = t_{g+1} - t_g of a series, with A = exp(-d_g/2 G):
//   B = (I - A A^T)^{-1} A            O_g     = -B                                   (:219-226)
//   P - I = B A^T (= A (I - A^T A)^{-1} A^T)      added to R_{g+1}                   (Dcontrib1 :228)
//   Q - I = A^T B                                   added to R_g                     (Dcontrib2 :229)
//   R_r = I + (P_{r-1} - I) + (Q_r - I) + shift                                       (:231-239, :265)
// A comes from ONE eigendecomposition G = V diag(lam) V^{-1} made on the host (the reference's compute_eG,
// model_utils.py:12-29): A - I = Re sum_k (e^{c lam_k} - 1) M_k with M_k = V[:,k] V^{-1}[k,:], c = -d/2.  Working with
// A - I and expm1 keeps I - A A^T = -(D + D^T + D D^T), D = A - I, free of cancellation for small gaps (fp32 safe).
//
// Forward: one THREAD per gap, everything in registers; a warp covers 32 consecutive gaps and finishes the 31 rows
// between them (P - I of the left gap arrives by __shfl_up), so neighbouring warps overlap by one gap.
//
// Backward (cotangents U = gR_{g+1}, W = gR_g, H = gO_g; Us = U + U^T, Ws = W + W^T):
//   gA = X1 + B (A^T X1 + X2),   X1 = (Us - H A^T) B - H,   X2 = Ws + (Ws A^T - H^T) B
// (derivation in DESIGN.md; uses P = I + B A^T, Q = I + A^T B, so nothing is factorised again), then the adjoint of the
// matrix exponential in the eigenbasis (Daleckii-Krein):  Y = V^{-1} gA^T V,  Z_jk += Y_kj Phi_jk(d),
// Phi_jk = (e^{c lam_j} - e^{c lam_k}) / (lam_j - lam_k)  (c e^{c lam_j} on the diagonal and for equal eigenvalues).
// Z (ell x ell complex, fp64) is summed over all gaps: warp butterfly per entry, per-lane fp64 accumulators across the
// tiles of a persistent warp, one atomicAdd per entry and warp at the end.  The host finishes gG = Re(V^{-T} Z V^T).
#pragma once
#include "cr_common.cuh"
#include "cr_tpn_common.cuh"   // record_stride

namespace crb200 {

using PegFwdArgs = ::crb200_peg_fwd_args;
using PegBwdArgs = ::crb200_peg_bwd_args;

constexpr int kPegMaxEll = 8;
constexpr int kPegThreads = 128;

template <typename CT> __device__ __forceinline__ CT peg_exp(CT x);
template <> __device__ __forceinline__ float peg_exp<float>(float x) { return expf(x); }
template <> __device__ __forceinline__ double peg_exp<double>(double x) { return exp(x); }
template <typename CT> __device__ __forceinline__ CT peg_expm1(CT x);
template <> __device__ __forceinline__ float peg_expm1<float>(float x) { return expm1f(x); }
template <> __device__ __forceinline__ double peg_expm1<double>(double x) { return expm1(x); }
__device__ __forceinline__ void peg_sincos(float x, float* s, float* c) { sincosf(x, s, c); }
__device__ __forceinline__ void peg_sincos(double x, double* s, double* c) { sincos(x, s, c); }
__device__ __forceinline__ float peg_sin(float x) { return sinf(x); }
__device__ __forceinline__ double peg_sin(double x) { return sin(x); }

// constants of one model in shared memory (compute type CT)
template <typename CT, int L>
struct PegConsts {
  CT lre[L], lim[L];
  CT Mre[L][L * L], Mim[L][L * L];
  CT shift[L * L];
};

template <typename CT, int L>
__device__ __forceinline__ void peg_load_consts(PegConsts<CT, L>* s, const double* lam_re, const double* lam_im, const double* M_re,
                                                const double* M_im, const double* shift) {
  for (int i = threadIdx.x; i < L; i += blockDim.x) { s->lre[i] = (CT)lam_re[i]; s->lim[i] = (CT)lam_im[i]; }
  for (int i = threadIdx.x; i < L * L * L; i += blockDim.x) {
    (&s->Mre[0][0])[i] = (CT)M_re[i];
    (&s->Mim[0][0])[i] = (CT)M_im[i];
  }
  for (int i = threadIdx.x; i < L * L; i += blockDim.x) s->shift[i] = shift != nullptr ? (CT)shift[i] : CT(0);
}

// D = A - I = Re sum_k (e^{c lam_k} - 1) M_k ; also returns e^{c lam_k} (needed by the backward pass)
template <typename CT, int L>
__device__ __forceinline__ void peg_expm_minus_I(CT (&D)[L][L], const PegConsts<CT, L>* s, const CT c, const int nterms) {
#pragma unroll
  for (int r = 0; r < L; ++r)
#pragma unroll
    for (int q = 0; q < L; ++q) D[r][q] = CT(0);
#pragma unroll
  for (int k = 0; k < L; ++k) {
    if (k >= nterms) break;                         // uniform: conjugate pairs folded by the caller
    const CT a = c * s->lre[k], b = c * s->lim[k];
    const CT em1 = peg_expm1<CT>(a);
    CT sb, cb;
    peg_sincos(b, &sb, &cb);
    const CT sh = peg_sin(CT(0.5) * b);
    const CT re = em1 * cb - CT(2) * sh * sh;       // Re(e^{a+ib} - 1) = expm1(a) cos b + (cos b - 1)
    const CT im = (em1 + CT(1)) * sb;
#pragma unroll
    for (int r = 0; r < L; ++r)
#pragma unroll
      for (int q = 0; q < L; ++q) D[r][q] = fma(re, s->Mre[k][r * L + q], fma(-im, s->Mim[k][r * L + q], D[r][q]));
  }
}

template <typename T, int L>
__device__ __forceinline__ void peg_store_block(T* g, const T (&M)[L][L], const bool vec_ok) {
  constexpr int VE = 16 / (int)sizeof(T);
  if constexpr ((L * L) % VE == 0) {
    if (vec_ok) {
      const T* flat = &M[0][0];
#pragma unroll
      for (int i = 0; i < L * L; i += VE) {
        if constexpr (sizeof(T) == 4) *reinterpret_cast<float4*>(g + i) = make_float4(flat[i], flat[i + 1], flat[i + 2], flat[i + 3]);
        else *reinterpret_cast<double2*>(g + i) = make_double2(flat[i], flat[i + 1]);
      }
      return;
    }
  }
#pragma unroll
  for (int r = 0; r < L; ++r)
#pragma unroll
    for (int q = 0; q < L; ++q) g[r * L + q] = M[r][q];
}

template <typename T, int L>
__device__ __forceinline__ void peg_load_block(T (&M)[L][L], const T* __restrict__ g, const bool vec_ok) {
  constexpr int VE = 16 / (int)sizeof(T);
  if constexpr ((L * L) % VE == 0) {
    if (vec_ok) {
      T* flat = &M[0][0];
#pragma unroll
      for (int i = 0; i < L * L; i += VE) {
        if constexpr (sizeof(T) == 4) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(g + i));
          flat[i] = v.x; flat[i + 1] = v.y; flat[i + 2] = v.z; flat[i + 3] = v.w;
        } else {
          const double2 v = __ldg(reinterpret_cast<const double2*>(g + i));
          flat[i] = v.x; flat[i + 1] = v.y;
        }
      }
      return;
    }
  }
#pragma unroll
  for (int r = 0; r < L; ++r)
#pragma unroll
    for (int q = 0; q < L; ++q) M[r][q] = g[r * L + q];
}

template <typename T, int L>
__device__ __forceinline__ void peg_load_row(T (&v)[L], const T* __restrict__ g, const bool vec_ok) {
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

// ------------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------------
template <typename T, int L>
__global__ void __launch_bounds__(kPegThreads, sizeof(T) == 4 ? 3 : 1) cr_peg_fwd_kernel(const PegFwdArgs a) {
  using CT = T;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  PegConsts<CT, L>* S = reinterpret_cast<PegConsts<CT, L>*>(smem_raw);
  peg_load_consts<CT, L>(S, a.lam_re, a.lam_im, a.M_re, a.M_im, a.shift);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = a.n;
  const int tiles = (n + 30) / 31;                              // a warp finishes 31 rows
  const long long vb = (long long)blockIdx.x * (kPegThreads / 32) + warp;
  if (vb >= (long long)tiles * a.batch) return;
  const int b = (int)(vb / tiles);
  const int t = (int)(vb - (long long)b * tiles);
  const int g = 31 * t - 1 + lane;                              // gap between rows g and g + 1 (virtual at both ends)
  const bool real = g >= 0 && g <= n - 2;
  const T* gaps = static_cast<const T*>(a.gaps) + (size_t)b * a.stride_gaps;

  CT Pm1[L][L], Qm1[L][L];                  // P - I (-> row g + 1), Q - I (-> row g): symmetric, only the lower triangles (q <= r) are formed
#pragma unroll
  for (int r = 0; r < L; ++r)
#pragma unroll
    for (int q = 0; q <= r; ++q) Pm1[r][q] = Qm1[r][q] = CT(0);
  bool bad = false;
  if (real) {
    const CT c = CT(-0.5) * (CT)gaps[g];
    CT D[L][L];
    peg_expm_minus_I<CT, L>(D, S, c, a.nterms > 0 ? a.nterms : L);
    // K = chol(I - A A^T), I - A A^T = -(D + D^T + D D^T)  (lower triangle)
    CT K[L][L];
#pragma unroll
    for (int r = 0; r < L; ++r)
#pragma unroll
      for (int q = 0; q <= r; ++q) {
        CT s = D[r][q] + D[q][r];
#pragma unroll
        for (int k = 0; k < L; ++k) s = fma(D[r][k], D[q][k], s);
        K[r][q] = -s;
      }
    CT inv[L];
#pragma unroll
    for (int k = 0; k < L; ++k) {
      const CT d = K[k][k];
      if (!(d > CT(0))) bad = true;
      const CT lkk = sqrt(d);
      inv[k] = CT(1) / lkk;
      K[k][k] = lkk;
#pragma unroll
      for (int r = k + 1; r < L; ++r) K[r][k] *= inv[k];
#pragma unroll
      for (int q = k + 1; q < L; ++q)
#pragma unroll
        for (int r = q; r < L; ++r) K[r][q] = fma(-K[r][k], K[q][k], K[r][q]);
    }
    // A = I + D;  B = (K K^T)^{-1} A by two triangular solves per column block (rows of the solve are independent columns)
    CT Bm[L][L];
#pragma unroll
    for (int r = 0; r < L; ++r)
#pragma unroll
      for (int q = 0; q < L; ++q) Bm[r][q] = D[r][q] + (r == q ? CT(1) : CT(0));
#pragma unroll
    for (int q = 0; q < L; ++q) {                                // forward: K z = A[:, q]
#pragma unroll
      for (int r = 0; r < L; ++r) {
        CT s = Bm[r][q];
#pragma unroll
        for (int k = 0; k < r; ++k) s = fma(-K[r][k], Bm[k][q], s);
        Bm[r][q] = s * inv[r];
      }
#pragma unroll
      for (int r = L - 1; r >= 0; --r) {                         // backward: K^T b = z
        CT s = Bm[r][q];
#pragma unroll
        for (int k = r + 1; k < L; ++k) s = fma(-K[k][r], Bm[k][q], s);
        Bm[r][q] = s * inv[r];
      }
    }
    // P - I = B A^T,  Q - I = A^T B   (A = I + D); both are symmetric
#pragma unroll
    for (int r = 0; r < L; ++r)
#pragma unroll
      for (int q = 0; q <= r; ++q) {
        CT sp = Bm[r][q], sq = Bm[r][q];                         // the identity part of A
#pragma unroll
        for (int k = 0; k < L; ++k) {
          sp = fma(Bm[r][k], D[q][k], sp);
          sq = fma(D[k][r], Bm[k][q], sq);
        }
        Pm1[r][q] = sp;
        Qm1[r][q] = sq;
      }
    // O_g = -B  (lane 0's gap belongs to the previous warp's tile)
    if (lane >= 1) {
      T* Og = static_cast<T*>(a.O) + (size_t)b * a.strideO + (size_t)g * (L * L);
#pragma unroll
      for (int r = 0; r < L; ++r)
#pragma unroll
        for (int q = 0; q < L; ++q) Bm[r][q] = -Bm[r][q];
      peg_store_block<T, L>(Og, Bm, is_aligned16(static_cast<T*>(a.O) + (size_t)b * a.strideO));
    }
  }
  if (bad && a.info != nullptr) atomicMax(a.info, 1);
  // row r = g (lanes 1..31): R_r = I + (P_{g-1} - I) + (Q_g - I) + shift
  CT Rr[L][L];
#pragma unroll
  for (int r = 0; r < L; ++r)
#pragma unroll
    for (int q = 0; q <= r; ++q) {
      const CT sym = __shfl_up_sync(0xffffffffu, Pm1[r][q], 1) + Qm1[r][q];
      Rr[r][q] = (r == q ? CT(1) : CT(0)) + sym + S->shift[r * L + q];
      if (q < r) Rr[q][r] = sym + S->shift[q * L + r];
    }
  if (lane >= 1 && g <= n - 1) {
    T* Rg = static_cast<T*>(a.R) + (size_t)b * a.strideR + (size_t)g * (L * L);
    peg_store_block<T, L>(Rg, Rr, is_aligned16(static_cast<T*>(a.R) + (size_t)b * a.strideR));
  }
}

// ------------------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------------------
template <typename CT, int L>
struct PegBwdConsts {
  PegConsts<CT, L> f;
  CT Vre[L * L], Vim[L * L], Wre[L * L], Wim[L * L];       // V and V^{-1}
  // the divided differences Phi_jk are formed in fp64 whatever CT is: e^{c lam_j} - e^{c lam_k} cancels like 1 / |c (lam_j - lam_k)|,
  // which for small gaps and close eigenvalues leaves no digits in fp32
  double lred[L], limd[L];
  double idre[L * L], idim[L * L], deg[L * L];               // 1 / (lam_j - lam_k) (0 where degenerate), degenerate mask
};

template <typename T, int L>
__global__ void __launch_bounds__(kPegThreads) cr_peg_bwd_kernel(const PegBwdArgs a) {
  using CT = T;
  constexpr int BS = L * L, NZ = 2 * BS, SLOTS = (NZ + 31) / 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  PegBwdConsts<CT, L>* S = reinterpret_cast<PegBwdConsts<CT, L>*>(smem_raw);
  peg_load_consts<CT, L>(&S->f, a.lam_re, a.lam_im, a.M_re, a.M_im, nullptr);
  for (int i = threadIdx.x; i < BS; i += blockDim.x) {
    S->Vre[i] = (CT)a.V_re[i]; S->Vim[i] = (CT)a.V_im[i];
    S->Wre[i] = (CT)a.Vinv_re[i]; S->Wim[i] = (CT)a.Vinv_im[i];
    S->idre[i] = a.invdl_re[i]; S->idim[i] = a.invdl_im[i]; S->deg[i] = a.degenerate[i];
  }
  for (int i = threadIdx.x; i < L; i += blockDim.x) { S->lred[i] = a.lamfull_re[i]; S->limd[i] = a.lamfull_im[i]; }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = a.n;
  const int ngap = n - 1;
  const long long tiles_per_series = (ngap + 31) / 32;
  const long long ntiles = tiles_per_series * a.batch;
  const long long nwarps = (long long)gridDim.x * (kPegThreads / 32);
  double accZ[SLOTS];
#pragma unroll
  for (int i = 0; i < SLOTS; ++i) accZ[i] = 0.0;
  // per-thread record [ A | B ] behind the constants; the stride keeps the threads of a warp on different banks
  constexpr int RS = record_stride<CT>(2 * BS, BS);
  CT* sA = reinterpret_cast<CT*>(smem_raw + align16(sizeof(PegBwdConsts<CT, L>))) + (size_t)threadIdx.x * RS;
  CT* sB = sA + BS;
  const int ncols = a.nterms > 0 ? a.nterms : L;     // columns j of Z the kernel produces (the caller mirrors the conjugate ones)

  for (long long vt = (long long)blockIdx.x * (kPegThreads / 32) + warp; vt < ntiles; vt += nwarps) {
    const int b = (int)(vt / tiles_per_series);
    const int g = (int)(vt - (long long)b * tiles_per_series) * 32 + lane;
    const bool real = g < ngap;
    CT gA[L][L];
    CT c = CT(0);
#pragma unroll
    for (int r = 0; r < L; ++r)
#pragma unroll
      for (int q = 0; q < L; ++q) gA[r][q] = CT(0);
    if (real) {
      const T* gR = static_cast<const T*>(a.gR) + (size_t)b * a.stride_gR;
      const T* gO = static_cast<const T*>(a.gO) + (size_t)b * a.stride_gO;
      const T* Og = static_cast<const T*>(a.O) + (size_t)b * a.strideO;
      const T* Ug = gR + (size_t)(g + 1) * BS;       // U = gR_{g+1}
      const T* Wg = gR + (size_t)g * BS;             // W = gR_g
      const T* Hg = gO + (size_t)g * BS;             // H = gO_g
      const bool vR = is_aligned16(gR), vH = is_aligned16(gO);
      c = CT(-0.5) * (CT) static_cast<const T*>(a.gaps)[(size_t)b * a.stride_gaps + g];
      // A and B = -O_g live in this thread's shared-memory record (rows are re-read many times; registers hold X1 and Y3)
      {
        CT A[L][L];
        peg_expm_minus_I<CT, L>(A, &S->f, c, a.nterms > 0 ? a.nterms : L);
#pragma unroll
        for (int r = 0; r < L; ++r) {
          A[r][r] += CT(1);
          sts_row<CT, L>(sA + r * L, A[r]);
        }
        peg_load_block<T, L>(A, Og + (size_t)g * BS, is_aligned16(Og));
#pragma unroll
        for (int r = 0; r < L; ++r) {
#pragma unroll
          for (int q = 0; q < L; ++q) A[r][q] = -A[r][q];
          sts_row<CT, L>(sB + r * L, A[r]);
        }
      }
      CT Y3[L][L];
      // X1 = (Us - H A^T) B - H,  Us = U + U^T   (rows of U, H stream from global memory / L1; gA holds X1)
#pragma unroll
      for (int r = 0; r < L; ++r) {
        CT hrow[L], us[L], tt[L];
        peg_load_row<T, L>(hrow, Hg + r * L, vH);
        peg_load_row<T, L>(us, Ug + r * L, vR);
#pragma unroll
        for (int q = 0; q < L; ++q) us[q] += __ldg(Ug + q * L + r);
#pragma unroll
        for (int q = 0; q < L; ++q) {
          CT arow[L];
          lds_row<CT, L>(arow, sA + q * L);
          CT sacc = us[q];
#pragma unroll
          for (int k = 0; k < L; ++k) sacc = fma(-hrow[k], arow[k], sacc);
          tt[q] = sacc;
        }
#pragma unroll
        for (int j = 0; j < L; ++j) gA[r][j] = -hrow[j];
#pragma unroll
        for (int q = 0; q < L; ++q) {
          CT brow[L];
          lds_row<CT, L>(brow, sB + q * L);
#pragma unroll
          for (int j = 0; j < L; ++j) gA[r][j] = fma(tt[q], brow[j], gA[r][j]);
        }
      }
      // Y3 = X2 + A^T X1,  X2 = Ws + (Ws A^T - H^T) B,  Ws = W + W^T
#pragma unroll
      for (int r = 0; r < L; ++r) {
        CT ws[L], tt[L];
        peg_load_row<T, L>(ws, Wg + r * L, vR);
#pragma unroll
        for (int q = 0; q < L; ++q) ws[q] += __ldg(Wg + q * L + r);
#pragma unroll
        for (int q = 0; q < L; ++q) {
          CT arow[L];
          lds_row<CT, L>(arow, sA + q * L);
          CT sacc = -(CT)__ldg(Hg + q * L + r);
#pragma unroll
          for (int k = 0; k < L; ++k) sacc = fma(ws[k], arow[k], sacc);
          tt[q] = sacc;
        }
#pragma unroll
        for (int j = 0; j < L; ++j) Y3[r][j] = ws[j];
#pragma unroll
        for (int q = 0; q < L; ++q) {
          CT brow[L];
          lds_row<CT, L>(brow, sB + q * L);
#pragma unroll
          for (int j = 0; j < L; ++j) Y3[r][j] = fma(tt[q], brow[j], Y3[r][j]);
        }
      }
#pragma unroll
      for (int k = 0; k < L; ++k) {
        CT arow[L];
        lds_row<CT, L>(arow, sA + k * L);
#pragma unroll
        for (int r = 0; r < L; ++r)
#pragma unroll
          for (int j = 0; j < L; ++j) Y3[r][j] = fma(arow[r], gA[k][j], Y3[r][j]);
      }
      // gA = X1 + B Y3   (in place, row by row)
#pragma unroll
      for (int r = 0; r < L; ++r) {
        CT brow[L];
        lds_row<CT, L>(brow, sB + r * L);
#pragma unroll
        for (int j = 0; j < L; ++j) {
          CT sacc = gA[r][j];
#pragma unroll
          for (int k = 0; k < L; ++k) sacc = fma(brow[k], Y3[k][j], sacc);
          gA[r][j] = sacc;
        }
      }
    }
    // Daleckii-Krein in the eigenbasis: T = gA^T V (column j), Y_kj = sum_m Vinv[k][m] T[m][j], Z_jk = Y_kj Phi_jk
    double ere[L], eim[L];                                         // e^{c lam_k} in fp64 (see PegBwdConsts)
    const double cd = (double)c;
#pragma unroll
    for (int k = 0; k < L; ++k) {
      const double ea = exp(cd * S->lred[k]);
      double sb, cb;
      sincos(cd * S->limd[k], &sb, &cb);
      ere[k] = ea * cb;
      eim[k] = ea * sb;
    }
#pragma unroll
    for (int j = 0; j < L; ++j) {
      if (j >= ncols) break;                                         // uniform
      CT Tre[L], Tim[L];
#pragma unroll
      for (int m = 0; m < L; ++m) {
        CT sr = CT(0), si = CT(0);
#pragma unroll
        for (int p = 0; p < L; ++p) {
          sr = fma(gA[p][m], S->Vre[p * L + j], sr);
          si = fma(gA[p][m], S->Vim[p * L + j], si);
        }
        Tre[m] = sr; Tim[m] = si;
      }
#pragma unroll
      for (int k = 0; k < L; ++k) {
        CT yr = CT(0), yi = CT(0);
#pragma unroll
        for (int m = 0; m < L; ++m) {
          const CT wr = S->Wre[k * L + m], wi = S->Wim[k * L + m];
          yr = fma(wr, Tre[m], fma(-wi, Tim[m], yr));
          yi = fma(wr, Tim[m], fma(wi, Tre[m], yi));
        }
        // Phi_jk = (e_j - e_k) / (lam_j - lam_k), or c e_j where the eigenvalues coincide
        const double dr = ere[j] - ere[k], di = eim[j] - eim[k];
        const double ir = S->idre[j * L + k], ii = S->idim[j * L + k], dg = S->deg[j * L + k] * cd;
        const CT pr = (CT)(dr * ir - di * ii + dg * ere[j]);
        const CT pi = (CT)(dr * ii + di * ir + dg * eim[j]);
        CT zr = yr * pr - yi * pi, zi = yr * pi + yi * pr;          // butterfly over the 32 gaps of the tile in CT,
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {                      // accumulation across tiles in fp64
          zr += __shfl_xor_sync(0xffffffffu, zr, off);
          zi += __shfl_xor_sync(0xffffffffu, zi, off);
        }
        const int s0 = 2 * (j * L + k), s1 = s0 + 1;
        if (lane == (s0 & 31)) accZ[s0 >> 5] += (double)zr;
        if (lane == (s1 & 31)) accZ[s1 >> 5] += (double)zi;
      }
    }
  }
  // one atomic per entry and warp: entry s lives on lane s % 32, slot s / 32
#pragma unroll
  for (int i = 0; i < SLOTS; ++i) {
    const int s = i * 32 + lane;
    if (s < NZ) atomicAdd(a.Z + s, accZ[i]);
  }
}

template <typename T, int L>
cudaError_t launch_peg_fwd(const PegFwdArgs& a, cudaStream_t stream) {
  const size_t smem = sizeof(PegConsts<T, L>);
  static std::atomic<unsigned char> attr_done[kMaxDevices];
  if (cudaError_t e = ensure_dynamic_smem(cr_peg_fwd_kernel<T, L>, (int)smem, attr_done); e != cudaSuccess) return e;
  const long long tiles = (long long)((a.n + 30) / 31) * a.batch;
  if (tiles <= 0) return cudaSuccess;
  const long long grid = (tiles + kPegThreads / 32 - 1) / (kPegThreads / 32);
  if (grid > 0x7fffffffLL) return cudaErrorInvalidValue;
  cr_peg_fwd_kernel<T, L><<<(unsigned)grid, kPegThreads, smem, stream>>>(a);
  return cudaGetLastError();
}

template <typename T, int L>
cudaError_t launch_peg_bwd(const PegBwdArgs& a, cudaStream_t stream) {
  const size_t smem = align16(sizeof(PegBwdConsts<T, L>)) + (size_t)kPegThreads * record_stride<T>(2 * L * L, L * L) * sizeof(T);
  static std::atomic<unsigned char> attr_done[kMaxDevices];
  if (cudaError_t e = ensure_dynamic_smem(cr_peg_bwd_kernel<T, L>, (int)smem, attr_done); e != cudaSuccess) return e;
  if (a.n < 2 || a.batch <= 0) return cudaSuccess;
  const long long tiles = (long long)((a.n - 1 + 31) / 32) * a.batch;
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long grid = (tiles + kPegThreads / 32 - 1) / (kPegThreads / 32);
  const long long cap = (long long)sms * 8;                      // persistent warps: at most 8 CTAs per SM worth of blocks
  if (grid > cap) grid = cap;
  cr_peg_bwd_kernel<T, L><<<(unsigned)grid, kPegThreads, smem, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace crb200
