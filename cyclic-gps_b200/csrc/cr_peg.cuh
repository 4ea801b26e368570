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
  constexpr int RSF = record_stride<T>(L * L, L * L);            // staging record stride (elements)
  T* wstage = reinterpret_cast<T*>(smem_raw + align16(sizeof(PegConsts<CT, L>))) + (size_t)warp * 32 * RSF;
  T* myrec = wstage + (size_t)lane * RSF;
  const unsigned srec0 = smem_u32(wstage), nsb = (unsigned)(RSF * sizeof(T));

  CT Pm1[L][L], Qm1[L][L];                  // P - I (-> row g + 1), Q - I (-> row g): symmetric, only the lower triangles (q <= r) are formed
#pragma unroll
  for (int r = 0; r < L; ++r)
#pragma unroll
    for (int q = 0; q <= r; ++q) Pm1[r][q] = Qm1[r][q] = CT(0);
  bool bad = false;
  double ld = 0.0;                            // -logdet(I - A A^T) of this gap (lanes >= 1: the gap belongs to this tile)
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
    double kprod = 1.0;
#pragma unroll
    for (int k = 0; k < L; ++k) {
      const CT d = K[k][k];
      if (!(d > CT(0))) bad = true;
      const CT lkk = sqrt(d);
      inv[k] = CT(1) / lkk;
      K[k][k] = lkk;
      kprod *= (double)lkk;
#pragma unroll
      for (int r = k + 1; r < L; ++r) K[r][k] *= inv[k];
#pragma unroll
      for (int q = k + 1; q < L; ++q)
#pragma unroll
        for (int r = q; r < L; ++r) K[r][q] = fma(-K[r][k], K[q][k], K[r][q]);
    }
    if (a.logdet != nullptr && lane >= 1) ld = -2.0 * log(kprod);
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
    // O_g = -B into this lane's staging record (lane 0's gap belongs to the previous warp's tile)
    if (lane >= 1) {
#pragma unroll
      for (int r = 0; r < L; ++r) {
#pragma unroll
        for (int q = 0; q < L; ++q) Bm[r][q] = -Bm[r][q];
        sts_row<T, L>(myrec + r * L, Bm[r]);
      }
    }
  }
  // The results leave through a per-warp staging area (one padded record per lane) as flat, fully coalesced 16-byte stores:
  // per-thread 256-byte block stores cost the LSU 32 sectors per instruction and made the kernel LSU-bound (ncu: STG = 20 % of the stalls).
  __syncwarp();
  {
    const int g0 = 31 * t;                                        // first gap / row of this tile (lane 1)
    T* Og = static_cast<T*>(a.O) + (size_t)b * a.strideO + (size_t)g0 * (L * L);
    if (a.O != nullptr) rec_s2g<T, L * L, 1>(Og, srec0, nsb, 1, cmin(31, (n - 1) - g0), is_aligned16(Og));
  }
  __syncwarp();
  if (bad && a.info != nullptr) atomicMax(a.info, 1);
  if (a.logdet != nullptr) {                  // logdet of the whole block-tridiagonal precision = -sum_g logdet(I - A_g A_g^T)
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) ld += __shfl_xor_sync(0xffffffffu, ld, off);
    if (lane == 0 && ld != 0.0) atomicAdd(a.logdet + b, ld);
  }
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
#pragma unroll
    for (int r = 0; r < L; ++r) sts_row<T, L>(myrec + r * L, Rr[r]);
  }
  __syncwarp();
  {
    const int g0 = 31 * t;
    T* Rg = static_cast<T*>(a.R) + (size_t)b * a.strideR + (size_t)g0 * (L * L);
    rec_s2g<T, L * L, 1>(Rg, srec0, nsb, 1, cmin(31, n - g0), is_aligned16(Rg));
  }
}

// ------------------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------------------
// One THREAD per gap forms gA (the cotangent of A = exp(cG)), a warp covers 32 consecutive gaps of one series.  The adjoint of
// the matrix exponential is NOT taken per gap: with Y_g = V^{-1} gA_g^T V and Phi_jk(d) = (e^{c lam_j} - e^{c lam_k}) / (lam_j - lam_k),
//   Z_jk = sum_g Y_g[k][j] Phi_jk(d_g) = ( T_j[k][j] - T_k[k][j] ) / (lam_j - lam_k),     T_m = V^{-1} ( sum_g e^{c_g lam_m} gA_g^T ) V,
// (and sum_g c_g e^{c_g lam_m} gA_g^T for the diagonal / equal eigenvalues), so the kernel only accumulates the 2 ell real
// weighted sums  S[row] = sum_g E[row][g] gA_g^T  in fp64 (the differences cancel like 1 / |c (lam_j - lam_k)|: fp64 weights,
// fp64 accumulation, exact fp32 x fp64 products) and the host finishes with ell x ell complex algebra.  Rows of E, in order,
// for every m < nterms:  Re e_m, c Re e_m, and for a complex eigenvalue (lam_im != 0) also Im e_m, c Im e_m  -- 2 ell rows in all
// because the caller folds conjugate pairs.
//
// Data movement: gR / gO / O tiles arrive by cp.async (coalesced 16-byte chunks) in one padded shared-memory record per gap,
// [ W = gR_g | H = gO_g | O_g ] (+ [ A ] when A and B do not both fit in registers); U = gR_{g+1} is the W field of the next
// record (33 records per warp).  B = -O_g and A are register-resident, every other operand streams by rows:
//   pass 1 (rows r):    X1[r] = (Us[r] - H[r] A^T) B - H[r]                     -> O slot (dead once B is in registers)
//   pass 2 (rows r):    Y3[r] = Ws[r] + (Ws[r] A^T - H[:,r]^T) B + (A^T X1)[r]   -> COLUMN r of the H slot (that column is dead)
//   pass 3 (columns j): gA[:,j] = X1[:,j] + B Y3[:,j] (+ 2 gld B[:,j])          -> row j of the H slot (= gA^T, entry order of S)
// Then the warps of the CTA share the accumulation: warp w owns rows [w RPW, (w+1) RPW) of E and walks the gA^T records of
// all NW tiles (lane = entry), so the fp64 accumulators are RPW x ceil(ell^2 / 32) registers per lane instead of 2 ell x that.
template <typename T, int L>
struct PegBwdSizes {
  static constexpr int BS = L * L;
  static constexpr bool A_REGS = 2 * BS * (int)sizeof(T) <= 512;            // A and B both register-resident
  static constexpr int NSLOT = A_REGS ? 3 : 4;
  static constexpr int RS = record_stride<T>(NSLOT * BS, BS);                // record stride (elements)
  static constexpr int NE = 2 * L;                                           // rows of E
  static constexpr size_t consts_bytes = align16(2 * L * sizeof(double) + 2 * (size_t)L * BS * sizeof(T));
  static constexpr size_t warp_bytes = align16((size_t)33 * RS * sizeof(T));
  static constexpr size_t e_bytes(int nw) { return (size_t)nw * ((NE + nw - 1) / nw) * 32 * sizeof(double); }   // per tile
  static constexpr size_t cta_bytes(int nw) { return consts_bytes + nw * (warp_bytes + e_bytes(nw)); }
  static constexpr int warps_per_sm(int nw) { return (int)((size_t)(227 * 1024) / (cta_bytes(nw) + 1024)) * nw; }
  static constexpr int pick_nw() {
    int best = 1;
    for (int nw = 2; nw <= 4; ++nw)
      if (cta_bytes(nw) + 1024 <= (size_t)227 * 1024 && warps_per_sm(nw) >= warps_per_sm(best)) best = nw;
    return best;
  }
};
template <typename T, int L>
struct PegBwdCfg : PegBwdSizes<T, L> {
  using Sz = PegBwdSizes<T, L>;
  static constexpr int NW = Sz::pick_nw();
  static constexpr int RPW = (Sz::NE + NW - 1) / NW;                         // rows of E per warp
  static constexpr int SLOTS = (Sz::BS + 31) / 32;
  static constexpr int CTAS_PER_SM = cmax(1, cmin(8, Sz::warps_per_sm(NW) / NW));
  static constexpr size_t E_TILE_BYTES = Sz::e_bytes(NW);                    // E rows of one tile
  static constexpr size_t REC_OFFSET = Sz::consts_bytes + NW * E_TILE_BYTES; // first warp's records
  static constexpr size_t CTA_BYTES = Sz::cta_bytes(NW);
};

template <typename T, int L>
struct PegBwdConsts {
  double lre[L], lim[L];
  T Mre[L][L * L], Mim[L][L * L];
};

template <typename T, int L>
__global__ void __launch_bounds__(PegBwdCfg<T, L>::NW * 32, 1) cr_peg_bwd_kernel(const PegBwdArgs a) {
  using Cfg = PegBwdCfg<T, L>;
  constexpr int BS = Cfg::BS, RS = Cfg::RS, NW = Cfg::NW, RPW = Cfg::RPW, NE = Cfg::NE, SLOTS = Cfg::SLOTS;
  constexpr bool A_REGS = Cfg::A_REGS;
  constexpr int AL = A_REGS ? L : 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  PegBwdConsts<T, L>* S = reinterpret_cast<PegBwdConsts<T, L>*>(smem_raw);
  double* sE = reinterpret_cast<double*>(smem_raw + Cfg::consts_bytes);                       // [NW tiles][NW row chunks][32 gaps][RPW rows]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  T* wrec = reinterpret_cast<T*>(smem_raw + Cfg::REC_OFFSET + (size_t)warp * Cfg::warp_bytes);
  const int nterms = a.nterms > 0 ? a.nterms : L;
  for (int i = threadIdx.x; i < L; i += blockDim.x) { S->lre[i] = a.lam_re[i]; S->lim[i] = a.lam_im[i]; }
  for (int i = threadIdx.x; i < L * BS; i += blockDim.x) {
    (&S->Mre[0][0])[i] = (T)a.M_re[i];
    (&S->Mim[0][0])[i] = (T)a.M_im[i];
  }
  for (int i = threadIdx.x; i < NW * NW * RPW * 32; i += blockDim.x) sE[i] = 0.0;              // rows past NE stay zero
  __syncthreads();

  const int n = a.n, ngap = n - 1;
  const long long tiles_per_series = (ngap + 31) / 32;
  const long long ntiles = tiles_per_series * a.batch;
  const long long stride_tiles = (long long)gridDim.x * NW;
  const unsigned nsb = (unsigned)(RS * sizeof(T));
  const unsigned srec0 = smem_u32(wrec);
  T* rec = wrec + (size_t)lane * RS;                      // this gap's record
  T* sW = rec;
  T* sH = rec + BS;
  T* sX = rec + 2 * BS;                                   // O_g, then X1
  T* sA = rec + (A_REGS ? 0 : 3 * BS);                    // only used when !A_REGS
  const T* sU = rec + RS;                                 // W field of the next record = gR_{g+1}
  double* myE = sE + (size_t)warp * (NW * RPW * 32) + (size_t)lane * RPW;       // E of this gap: [row chunk of warp w][gap][RPW rows]
  auto e_slot = [&](int row) -> double& { return myE[(row / RPW) * (32 * RPW) + (row % RPW)]; };

  double acc[RPW][SLOTS], accW[SLOTS];
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) {
    accW[s] = 0.0;
#pragma unroll
    for (int i = 0; i < RPW; ++i) acc[i][s] = 0.0;
  }

  // stage the W (33 rows of gR) and O fields, or the H field, of tile vt
  auto stage_WO = [&](long long vt) {
    const int b = (int)(vt / tiles_per_series);
    const int g0 = (int)(vt - (long long)b * tiles_per_series) * 32;
    const T* gR = static_cast<const T*>(a.gR) + (size_t)b * a.stride_gR + (size_t)g0 * BS;
    const T* Og = static_cast<const T*>(a.O) + (size_t)b * a.strideO + (size_t)g0 * BS;
    rec_g2s<T, BS, 1>(srec0, nsb, gR, 0, cmin(33, n - g0), is_aligned16(gR));
    rec_g2s<T, BS, 1>(srec0 + 2 * BS * (unsigned)sizeof(T), nsb, Og, 0, cmin(32, ngap - g0), is_aligned16(Og));
  };
  auto stage_H = [&](long long vt) {
    const int b = (int)(vt / tiles_per_series);
    const int g0 = (int)(vt - (long long)b * tiles_per_series) * 32;
    const T* gO = static_cast<const T*>(a.gO) + (size_t)b * a.stride_gO + (size_t)g0 * BS;
    rec_g2s<T, BS, 1>(srec0 + BS * (unsigned)sizeof(T), nsb, gO, 0, cmin(32, ngap - g0), is_aligned16(gO));
  };

  const long long base0 = (long long)blockIdx.x * NW;
  if (base0 + warp < ntiles) { stage_WO(base0 + warp); stage_H(base0 + warp); }
  cp_async_commit();

  for (long long base = base0; base < ntiles; base += stride_tiles) {
    const long long vt = base + warp;
    const bool tile_ok = vt < ntiles;
    int g = 0;
    bool real = false;
    T Ar[AL][AL], Bm[L][L];
    double cd = 0.0;
    T gl2 = T(0);                                                     // 2 x cotangent of this series' logdet output
    if (tile_ok) {
      const int b = (int)(vt / tiles_per_series);
      if (a.g_logdet != nullptr) gl2 = (T)(2.0 * a.g_logdet[b]);
      g = (int)(vt - (long long)b * tiles_per_series) * 32 + lane;
      real = g < ngap;
      // weights of E and the coefficients of A = I + Re sum_m (e_m - 1) M_m, from fp64 exponentials (the staged tile lands meanwhile)
      T cre[L], cim[L];
      if (real) cd = (double)(T(-0.5) * static_cast<const T*>(a.gaps)[(size_t)b * a.stride_gaps + g]);
      {
        int row = 0;
#pragma unroll
        for (int m = 0; m < L; ++m) {
          if (m >= nterms) break;                                     // uniform
          const bool cplx = S->lim[m] != 0.0;                         // uniform
          if (row + (cplx ? 4 : 2) > NE) break;                       // conjugate pairs not folded by the caller: rows beyond 2 ell are dropped
          const double ea = exp(cd * S->lre[m]);
          double sb = 0.0, cb = 1.0;
          if (cplx) sincos(cd * S->lim[m], &sb, &cb);
          const double er = ea * cb, ei = ea * sb;
          cre[m] = (T)(er - 1.0);
          cim[m] = (T)ei;
          const double w = real ? 1.0 : 0.0;
          e_slot(row + 0) = w * er;
          e_slot(row + 1) = w * cd * er;
          row += 2;
          if (cplx) {
            e_slot(row + 0) = w * ei;
            e_slot(row + 1) = w * cd * ei;
            row += 2;
          }
        }
      }
      if (real) {
#pragma unroll
        for (int r = 0; r < L; ++r) {
          T arow[L];
#pragma unroll
          for (int q = 0; q < L; ++q) arow[q] = (r == q) ? T(1) : T(0);
#pragma unroll
          for (int m = 0; m < L; ++m) {
            if (m >= nterms) break;                                   // uniform
#pragma unroll
            for (int q = 0; q < L; ++q) arow[q] = fma(cre[m], S->Mre[m][r * L + q], fma(-cim[m], S->Mim[m][r * L + q], arow[q]));
          }
          if constexpr (A_REGS) {
#pragma unroll
            for (int q = 0; q < L; ++q) Ar[r][q] = arow[q];
          } else {
            sts_row<T, L>(sA + r * L, arow);
          }
        }
      }
    }
    cp_async_wait_all();
    __syncwarp();
    if (tile_ok) {
      if (real) {
#pragma unroll
        for (int r = 0; r < L; ++r) {
          lds_row<T, L>(Bm[r], sX + r * L);
#pragma unroll
          for (int q = 0; q < L; ++q) Bm[r][q] = -Bm[r][q];
        }
        // pass 1: X1 rows -> O slot
#pragma unroll
        for (int r = 0; r < L; ++r) {
          T hrow[L], us[L], t1[L], x[L];
          lds_row<T, L>(hrow, sH + r * L);
          lds_row<T, L>(us, sU + r * L);
#pragma unroll
          for (int q = 0; q < L; ++q) us[q] += sU[q * L + r];
#pragma unroll
          for (int q = 0; q < L; ++q) {
            T arow[L];
            if constexpr (A_REGS) {
#pragma unroll
              for (int k = 0; k < L; ++k) arow[k] = Ar[q % AL][k % AL];
            } else {
              lds_row<T, L>(arow, sA + q * L);
            }
            T sacc = us[q];
#pragma unroll
            for (int k = 0; k < L; ++k) sacc = fma(-hrow[k], arow[k], sacc);
            t1[q] = sacc;
          }
#pragma unroll
          for (int j = 0; j < L; ++j) x[j] = -hrow[j];
#pragma unroll
          for (int q = 0; q < L; ++q)
#pragma unroll
            for (int j = 0; j < L; ++j) x[j] = fma(t1[q], Bm[q][j], x[j]);
          sts_row<T, L>(sX + r * L, x);
          sched_fence();
        }
        // pass 2: Y3 rows -> columns of the H slot
#pragma unroll
        for (int r = 0; r < L; ++r) {
          T ws[L], t2[L], y[L];
          lds_row<T, L>(ws, sW + r * L);
#pragma unroll
          for (int q = 0; q < L; ++q) ws[q] += sW[q * L + r];
#pragma unroll
          for (int q = 0; q < L; ++q) {
            T arow[L];
            if constexpr (A_REGS) {
#pragma unroll
              for (int k = 0; k < L; ++k) arow[k] = Ar[q % AL][k % AL];
            } else {
              lds_row<T, L>(arow, sA + q * L);
            }
            T sacc = -sH[q * L + r];
#pragma unroll
            for (int k = 0; k < L; ++k) sacc = fma(ws[k], arow[k], sacc);
            t2[q] = sacc;
          }
#pragma unroll
          for (int j = 0; j < L; ++j) y[j] = ws[j];
#pragma unroll
          for (int q = 0; q < L; ++q)
#pragma unroll
            for (int j = 0; j < L; ++j) y[j] = fma(t2[q], Bm[q][j], y[j]);
#pragma unroll
          for (int k = 0; k < L; ++k) {
            T xr[L];
            lds_row<T, L>(xr, sX + k * L);
            T akr;
            if constexpr (A_REGS) akr = Ar[k % AL][r % AL]; else akr = sA[k * L + r];
#pragma unroll
            for (int j = 0; j < L; ++j) y[j] = fma(akr, xr[j], y[j]);
          }
#pragma unroll
          for (int j = 0; j < L; ++j) sH[j * L + r] = y[j];
          sched_fence();
        }
        // pass 3: gA columns -> rows of the H slot (gA^T)
#pragma unroll
        for (int j = 0; j < L; ++j) {
          T yc[L], ga[L];
          lds_row<T, L>(yc, sH + j * L);
#pragma unroll
          for (int r = 0; r < L; ++r) ga[r] = fma(gl2, Bm[r][j], sX[r * L + j]);    // d(-logdet(I - A A^T)) / dA = 2 B
#pragma unroll
          for (int r = 0; r < L; ++r)
#pragma unroll
            for (int k = 0; k < L; ++k) ga[r] = fma(Bm[r][k], yc[k], ga[r]);
          sts_row<T, L>(sH + j * L, ga);
          sched_fence();
        }
      } else {
#pragma unroll
        for (int i = 0; i < BS; ++i) sH[i] = T(0);                    // no gap here: contributes nothing (and no NaN x 0)
      }
    }
    __syncwarp();
    if (tile_ok) {
      // cotangent of `shift` = sum of all gR rows: the rows of this tile (row g0 + 32 is the next tile's first one, unless the series ends here)
      const int g0 = g - lane;
      const int cnt = (g0 + 32 >= ngap) ? n - g0 : 32;
#pragma unroll 4
      for (int gg = 0; gg < cnt; ++gg)
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) {
          const int entry = lane + 32 * s;
          if (entry < BS) accW[s] += (double)wrec[(size_t)gg * RS + entry];
        }
    }
    __syncwarp();
    const long long next = vt + stride_tiles;
    if (next < ntiles) stage_WO(next);                                // W and X1 are dead for every lane of this warp
    __syncthreads();
    // accumulation: this warp's rows of E against the gA^T records of every tile of the CTA
#pragma unroll 1
    for (int tt = 0; tt < NW; ++tt) {
      if (base + tt >= ntiles) break;                                 // uniform
      const double* Et = sE + (size_t)tt * (NW * RPW * 32) + (size_t)warp * RPW * 32;
      const T* grec = reinterpret_cast<const T*>(smem_raw + Cfg::REC_OFFSET + (size_t)tt * Cfg::warp_bytes) + BS;
#pragma unroll 4
      for (int gg = 0; gg < 32; ++gg) {
        double e[RPW];
        lds_row<double, RPW>(e, Et + gg * RPW);
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) {
          const int entry = lane + 32 * s;
          const double v = (entry < BS) ? (double)grec[(size_t)gg * RS + entry] : 0.0;
#pragma unroll
          for (int i = 0; i < RPW; ++i) acc[i][s] = fma(e[i], v, acc[i][s]);
        }
      }
    }
    __syncthreads();
    if (next < ntiles) stage_H(next);
    cp_async_commit();
  }
  cp_async_wait_all();
#pragma unroll
  for (int i = 0; i < RPW; ++i) {
    const int row = warp * RPW + i;
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
      const int entry = lane + 32 * s;
      if (row < NE && entry < BS) atomicAdd(a.S + (size_t)row * BS + entry, acc[i][s]);
    }
  }
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) {
    const int entry = lane + 32 * s;
    if (entry < BS) atomicAdd(a.S + (size_t)NE * BS + entry, accW[s]);
  }
}

template <typename T, int L>
cudaError_t launch_peg_fwd(const PegFwdArgs& a, cudaStream_t stream) {
  const size_t smem = align16(sizeof(PegConsts<T, L>)) + (size_t)kPegThreads * record_stride<T>(L * L, L * L) * sizeof(T);
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
  using Cfg = PegBwdCfg<T, L>;
  const size_t smem = Cfg::CTA_BYTES;
  static std::atomic<unsigned char> attr_done[kMaxDevices];
  if (cudaError_t e = ensure_dynamic_smem(cr_peg_bwd_kernel<T, L>, (int)smem, attr_done); e != cudaSuccess) return e;
  if (a.n < 2 || a.batch <= 0) return cudaSuccess;
  const long long tiles = (long long)((a.n - 1 + 31) / 32) * a.batch;
  int dev = 0, sms = 148, resident = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, cr_peg_bwd_kernel<T, L>, Cfg::NW * 32, smem) != cudaSuccess || resident < 1)
    resident = Cfg::CTAS_PER_SM;
  long long grid = (tiles + Cfg::NW - 1) / Cfg::NW;
  const long long cap = (long long)sms * resident;               // persistent CTAs: exactly one resident wave
  if (grid > cap) grid = cap;
  cr_peg_bwd_kernel<T, L><<<(unsigned)grid, Cfg::NW * 32, smem, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace crb200
