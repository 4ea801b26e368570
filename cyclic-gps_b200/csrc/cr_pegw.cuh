// Precision-block builder for LARGE ranks (9 <= ell <= 32): one WARP per gap, block algebra on the FP64 tensor path.
//
// Same contract as the thread-per-gap kernels of cr_peg.cuh (reference cyclic_gps/models.py:181-239, :254-268 and the torch
// autograd backward through them), for blocks that do not fit one thread's registers.  Every ell x ell matrix of a gap lives
// in shared memory as a padded LP x (LP + 4) block of DOUBLES (cr_mma_common.cuh), whatever the storage type, and everything
// of order ell^3 is a product on mma.sync.m8n8k4.f64:
//   forward   D = A - I = Re sum_m (e^{c lam_m} - 1) M_m        (lane m forms the coefficients, the lanes share the entries)
//             K = chol(-(D + D^T + D D^T)),  Ki = K^{-1},  logdet -= 2 sum log K_jj
//             B = Ki^T (Ki A),   P - I = B A^T,   Q - I = A^T B,   O_g = -B,
//             R_g = I + shift + (P_{g-1} - I) + (Q_g - I)         (P - I of the left gap: the neighbouring warp, via shared memory)
//   backward  gA = X1 + B (A^T X1 + X2) (+ 2 gld B),  X1 = (Us - H A^T) B - H,  X2 = Ws + (Ws A^T - H^T) B      (cr_peg.cuh)
// The backward kernel stores gA_g per gap (fp64); the sum over gaps S[row] = sum_g E[row][g] gA_g^T of the thread-per-gap
// backward is a plain (2 ell x gaps) x (gaps x ell^2) GEMM here, left to the caller (cuBLAS): with ell^2 * 2 ell accumulators
// it does not fit the registers of a CTA.
#pragma once
#include "cr_mma_common.cuh"
#include "cr_peg.cuh"

namespace crb200 {

template <typename T, int L>
struct PegwCfg {
  using Geo = MmaGeom<L>;
  static constexpr int LP = Geo::LP, LD = Geo::LD, BLK = Geo::BLK;
  static constexpr int VEC = 4 * LP;                                    // coefficients (re, im) + column buffer of the Cholesky
  // forward: A | K -> Ki -> P - I | T -> B
  static constexpr int REC_F = 3 * BLK + VEC;
  static constexpr int W_F = LP <= 16 ? 8 : 4;                          // warps (gaps) per CTA incl. the halo gap
  static constexpr size_t SMEM_F = (size_t)W_F * REC_F * sizeof(double);
  // backward: A | B | U -> T1 -> T2 | W -> Ws | H -> Y3 | X1
  static constexpr int REC_B = 6 * BLK + VEC;
  static constexpr int W_B = LP <= 16 ? 6 : (LP <= 24 ? 6 : 4);
  static constexpr size_t SMEM_B = (size_t)W_B * REC_B * sizeof(double);
};

// coefficients of the eigen-expansion of A - I for one gap: lane m forms e^{c lam_m} - 1 in fp64 and publishes (re, im)
__device__ __forceinline__ void pegw_coeffs(double* cre, double* cim, const double c, const double* __restrict__ lam_re,
                                            const double* __restrict__ lam_im, const int nterms, const int lane) {
  if (lane < nterms) {
    const double a = c * lam_re[lane], b = c * lam_im[lane];
    const double em1 = expm1(a);
    double sb, cb;
    sincos(b, &sb, &cb);
    const double sh = sin(0.5 * b);
    cre[lane] = em1 * cb - 2.0 * sh * sh;             // Re(e^{a + ib} - 1)
    cim[lane] = (em1 + 1.0) * sb;
  }
  __syncwarp();
}

// D = A - I (diag == 0) or A (diag == 1) into the padded block S; padding rows / columns are zero
template <int L, int LP>
__device__ __forceinline__ void pegw_expansion(double* S, const double* cre, const double* cim, const double* __restrict__ M_re,
                                               const double* __restrict__ M_im, const int nterms, const double diag, const int lane) {
  constexpr int LD = LP + 4, BS = L * L;
  for (int i = lane; i < LP * LP; i += 32) {
    const int r = i / LP, q = i - r * LP;
    double v = 0.0;
    if (r < L && q < L) {
      const int e = r * L + q;
      v = (r == q) ? diag : 0.0;
      for (int m = 0; m < nterms; ++m) v = fma(cre[m], __ldg(M_re + (size_t)m * BS + e), fma(-cim[m], __ldg(M_im + (size_t)m * BS + e), v));
    }
    S[r * LD + q] = v;
  }
  __syncwarp();
}

// ------------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------------
template <typename T, int L>
__global__ void __launch_bounds__(32 * PegwCfg<T, L>::W_F, 1) cr_pegw_fwd_kernel(const PegFwdArgs a) {
  using C = PegwCfg<T, L>;
  constexpr int KP = MmaGeom<L>::KP, LA = MmaGeom<L>::LA;
  constexpr int LP = C::LP, LD = C::LD, BLK = C::BLK, NTL = LP / 8, W = C::W_F, OWN = W - 1, BS = L * L;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int lr = lane >> 2, lc = lane & 3;
  double* N = reinterpret_cast<double*>(smem_raw) + (size_t)warp * C::REC_F;
  double* SA = N;                       // D -> A
  double* SK = N + BLK;                 // K -> Ki -> P - I (read by the warp to the right)
  double* SB = N + 2 * BLK;             // Ki A -> B
  double* cre = N + 3 * BLK;
  double* cim = cre + LP;
  double* CB = cim + LP;                // column buffer of the Cholesky (2 LP)
  const int n = a.n;
  const int nterms = a.nterms > 0 ? a.nterms : L;
  const int tiles = (n + OWN - 1) / OWN;                 // a CTA finishes OWN rows
  const int b = blockIdx.x / tiles;
  const int t = blockIdx.x - b * tiles;
  const int g = OWN * t - 1 + warp;                      // gap between rows g and g + 1 (virtual at both ends of the series)
  const bool real = g >= 0 && g <= n - 2;
  double acc[NTL][NTL][2], accQ[NTL][NTL][2];
  acc_zero<LP>(accQ);
  bool bad = false;
  double ld = 0.0;
  if (real) {
    const double c = (double)(T(-0.5) * static_cast<const T*>(a.gaps)[(size_t)b * a.stride_gaps + g]);
    pegw_coeffs(cre, cim, c, a.lam_re, a.lam_im, nterms, lane);
    pegw_expansion<L, LP>(SA, cre, cim, a.M_re, a.M_im, nterms, 0.0, lane);          // D
    // K = -(D + D^T + D D^T), identity on the padded diagonal
    acc_zero<LP>(acc);
    warp_gemm<LP, false, true, K_FULL, false, KP>(acc, SA, SA, lane);
#pragma unroll
    for (int mt = 0; mt < NTL; ++mt)
#pragma unroll
      for (int nt = 0; nt < NTL; ++nt) {
        const int row = mt * 8 + lr, col = nt * 8 + 2 * lc;
        double v0 = -(acc[mt][nt][0] + SA[row * LD + col] + SA[col * LD + row]);
        double v1 = -(acc[mt][nt][1] + SA[row * LD + col + 1] + SA[(col + 1) * LD + row]);
        if (L < LP) {
          if (row >= L && row == col) v0 = 1.0;
          if (row >= L && row == col + 1) v1 = 1.0;
        }
        *reinterpret_cast<double2*>(SK + row * LD + col) = make_double2(v0, v1);
      }
    __syncwarp();
    double invd[LP];
    bad = warp_cholesky<LP, LA>(SK, CB, invd, lane);
    if (a.logdet != nullptr && warp >= 1 && lane == 0) {
      double p = 1.0;
#pragma unroll
      for (int j = 0; j < L; ++j) p *= invd[j];
      ld = 2.0 * log(p);                                   // -logdet(I - A A^T) = -2 sum log K_jj
    }
    warp_tri_inverse<LP, LA>(SK, invd, lane);                  // SK = Ki
    if (lane < L) SA[lane * LD + lane] += 1.0;             // SA = A
    __syncwarp();
    acc_zero<LP>(acc);
    warp_gemm<LP, false, false, K_FULL, false, KP>(acc, SK, SA, lane);        // Ki A
    acc_to_smem<LP>(SB, acc, 1.0, lane);
    __syncwarp();
    acc_zero<LP>(acc);
    warp_gemm<LP, true, false, K_FULL, false, KP>(acc, SK, SB, lane);         // B = Ki^T (Ki A)
    __syncwarp();
    acc_to_smem<LP>(SB, acc, 1.0, lane);
    if (warp >= 1) {                                       // O_g = -B (the halo warp's gap belongs to the previous CTA)
      T* Og = static_cast<T*>(a.O) + (size_t)b * a.strideO + (size_t)g * BS;
      acc_to_global<T, L, LP>(Og, acc, -1.0, lane, is_aligned16(static_cast<T*>(a.O) + (size_t)b * a.strideO) && (BS * sizeof(T)) % 16 == 0);
    }
    __syncwarp();
    acc_zero<LP>(acc);
    warp_gemm<LP, false, true, K_FULL, false, KP>(acc, SB, SA, lane);         // P - I = B A^T
    acc_to_smem<LP>(SK, acc, 1.0, lane);                                  // (Ki is dead)
    warp_gemm<LP, true, false, K_FULL, false, KP>(accQ, SA, SB, lane);        // Q - I = A^T B
  } else {
    mma_fill_block<LP>(SK, false, lane);
  }
  if (bad && a.info != nullptr && lane == 0) atomicMax(a.info, 1);
  if (a.logdet != nullptr && lane == 0 && ld != 0.0) atomicAdd(a.logdet + b, ld);
  __syncthreads();
  // row g (warps 1..W-1): R_g = I + shift + (P_{g-1} - I) + (Q_g - I)
  if (warp >= 1 && g <= n - 1) {
    const double* Pp = N - C::REC_F + BLK;                 // SK of the warp to the left
    T* Rg = static_cast<T*>(a.R) + (size_t)b * a.strideR + (size_t)g * BS;
    const bool vec = is_aligned16(static_cast<T*>(a.R) + (size_t)b * a.strideR) && (BS * sizeof(T)) % 16 == 0;
#pragma unroll
    for (int mt = 0; mt < NTL; ++mt)
#pragma unroll
      for (int nt = 0; nt < NTL; ++nt) {
        const int row = mt * 8 + lr, col = nt * 8 + 2 * lc;
        const double2 p = *reinterpret_cast<const double2*>(Pp + row * LD + col);
        double v0 = accQ[mt][nt][0] + p.x + (row == col ? 1.0 : 0.0);
        double v1 = accQ[mt][nt][1] + p.y + (row == col + 1 ? 1.0 : 0.0);
        if (a.shift != nullptr && row < L) {
          if (col < L) v0 += a.shift[row * L + col];
          if (col + 1 < L) v1 += a.shift[row * L + col + 1];
        }
        frag_pair_store<T, L>(Rg, row, col, v0, v1, vec);
      }
  }
}

// ------------------------------------------------------------------------------------------------------------
// backward: gA_g per gap (fp64, row-major ell x ell)
// ------------------------------------------------------------------------------------------------------------
template <typename T, int L>
__global__ void __launch_bounds__(32 * PegwCfg<T, L>::W_B, 1) cr_pegw_bwd_kernel(const PegBwdArgs a) {
  using C = PegwCfg<T, L>;
  constexpr int KP = MmaGeom<L>::KP, LA = MmaGeom<L>::LA;
  constexpr int LP = C::LP, LD = C::LD, BLK = C::BLK, NTL = LP / 8, W = C::W_B, BS = L * L;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int lr = lane >> 2, lc = lane & 3;
  double* N = reinterpret_cast<double*>(smem_raw) + (size_t)warp * C::REC_B;
  double* SA = N;
  double* SB = N + BLK;
  double* ST = N + 2 * BLK;             // U -> T1 -> T2
  double* SW = N + 3 * BLK;             // W -> Ws
  double* SH = N + 4 * BLK;             // H -> Y3
  double* SX = N + 5 * BLK;             // X1
  double* cre = N + 6 * BLK;
  double* cim = cre + LP;
  const int n = a.n, ngap = n - 1;
  const int nterms = a.nterms > 0 ? a.nterms : L;
  const long long total = (long long)ngap * a.batch;
  for (long long vg = (long long)blockIdx.x * W + warp; vg < total; vg += (long long)gridDim.x * W) {
    const int b = (int)(vg / ngap);
    const int g = (int)(vg - (long long)b * ngap);
    const T* gR = static_cast<const T*>(a.gR) + (size_t)b * a.stride_gR;
    const T* gO = static_cast<const T*>(a.gO) + (size_t)b * a.stride_gO;
    const T* Ob = static_cast<const T*>(a.O) + (size_t)b * a.strideO;
    const bool vR = is_aligned16(gR) && (BS * sizeof(T)) % 16 == 0, vH = is_aligned16(gO) && (BS * sizeof(T)) % 16 == 0,
               vO = is_aligned16(Ob) && (BS * sizeof(T)) % 16 == 0;
    mma_stage_issue<T, L, LP>(ST, gR + (size_t)(g + 1) * BS, false, lane, vR);      // U = gR_{g+1}
    mma_stage_issue<T, L, LP>(SW, gR + (size_t)g * BS, false, lane, vR);            // W = gR_g
    mma_stage_issue<T, L, LP>(SH, gO + (size_t)g * BS, false, lane, vH);            // H = gO_g
    mma_stage_issue<T, L, LP>(SB, Ob + (size_t)g * BS, false, lane, vO);            // O_g = -B
    const double c = (double)(T(-0.5) * static_cast<const T*>(a.gaps)[(size_t)b * a.stride_gaps + g]);
    const double gl2 = a.g_logdet != nullptr ? 2.0 * a.g_logdet[b] : 0.0;
    pegw_coeffs(cre, cim, c, a.lam_re, a.lam_im, nterms, lane);
    pegw_expansion<L, LP>(SA, cre, cim, a.M_re, a.M_im, nterms, 1.0, lane);          // A
    cp_async_wait_all();
    __syncwarp();
    mma_stage_finish<T, L, LP>(ST, false, lane);
    mma_stage_finish<T, L, LP>(SW, false, lane);
    mma_stage_finish<T, L, LP>(SH, false, lane);
    mma_stage_finish<T, L, LP>(SB, false, lane);
    __syncwarp();
    double acc[NTL][NTL][2], sym[NTL][NTL][2];
    // B = -O in place; Us = U + U^T into registers (its slot then takes T1), Ws = W + W^T in place
    for (int i = lane; i < LP * LP; i += 32) {
      const int r = i / LP, q = i - r * LP;
      SB[r * LD + q] = -SB[r * LD + q];
    }
#pragma unroll
    for (int mt = 0; mt < NTL; ++mt)
#pragma unroll
      for (int nt = 0; nt < NTL; ++nt) {
        const int row = mt * 8 + lr, col = nt * 8 + 2 * lc;
        sym[mt][nt][0] = ST[row * LD + col] + ST[col * LD + row];
        sym[mt][nt][1] = ST[row * LD + col + 1] + ST[(col + 1) * LD + row];
      }
    __syncwarp();
    // T1 = Us - H A^T  -> ST
    acc_zero<LP>(acc);
    warp_gemm<LP, false, true, K_FULL, false, KP>(acc, SH, SA, lane);
#pragma unroll
    for (int mt = 0; mt < NTL; ++mt)
#pragma unroll
      for (int nt = 0; nt < NTL; ++nt) {
        const int row = mt * 8 + lr, col = nt * 8 + 2 * lc;
        *reinterpret_cast<double2*>(ST + row * LD + col) = make_double2(sym[mt][nt][0] - acc[mt][nt][0], sym[mt][nt][1] - acc[mt][nt][1]);
      }
    // Ws into registers (written back after every lane has read W)
#pragma unroll
    for (int mt = 0; mt < NTL; ++mt)
#pragma unroll
      for (int nt = 0; nt < NTL; ++nt) {
        const int row = mt * 8 + lr, col = nt * 8 + 2 * lc;
        sym[mt][nt][0] = SW[row * LD + col] + SW[col * LD + row];
        sym[mt][nt][1] = SW[row * LD + col + 1] + SW[(col + 1) * LD + row];
      }
    __syncwarp();
    acc_to_smem<LP>(SW, sym, 1.0, lane);                                   // SW = Ws
    // X1 = T1 B - H  -> SX
    acc_zero<LP>(acc);
    warp_gemm<LP, false, false, K_FULL, false, KP>(acc, ST, SB, lane);
#pragma unroll
    for (int mt = 0; mt < NTL; ++mt)
#pragma unroll
      for (int nt = 0; nt < NTL; ++nt) {
        const int row = mt * 8 + lr, col = nt * 8 + 2 * lc;
        const double2 h = *reinterpret_cast<const double2*>(SH + row * LD + col);
        *reinterpret_cast<double2*>(SX + row * LD + col) = make_double2(acc[mt][nt][0] - h.x, acc[mt][nt][1] - h.y);
      }
    __syncwarp();                                                          // T1 consumed, Ws and X1 visible
    // T2 = Ws A^T - H^T  -> ST
    acc_zero<LP>(acc);
    warp_gemm<LP, false, true, K_FULL, false, KP>(acc, SW, SA, lane);
#pragma unroll
    for (int mt = 0; mt < NTL; ++mt)
#pragma unroll
      for (int nt = 0; nt < NTL; ++nt) {
        const int row = mt * 8 + lr, col = nt * 8 + 2 * lc;
        *reinterpret_cast<double2*>(ST + row * LD + col) =
            make_double2(acc[mt][nt][0] - SH[col * LD + row], acc[mt][nt][1] - SH[(col + 1) * LD + row]);
      }
    __syncwarp();                                                          // T2 visible, H consumed
    // Y3 = Ws + T2 B + A^T X1  -> SH
    acc_zero<LP>(acc);
    warp_gemm<LP, false, false, K_FULL, false, KP>(acc, ST, SB, lane);
    warp_gemm<LP, true, false, K_FULL, false, KP>(acc, SA, SX, lane);
#pragma unroll
    for (int mt = 0; mt < NTL; ++mt)
#pragma unroll
      for (int nt = 0; nt < NTL; ++nt) {
        const int row = mt * 8 + lr, col = nt * 8 + 2 * lc;
        const double2 w2 = *reinterpret_cast<const double2*>(SW + row * LD + col);
        *reinterpret_cast<double2*>(SH + row * LD + col) = make_double2(acc[mt][nt][0] + w2.x, acc[mt][nt][1] + w2.y);
      }
    __syncwarp();
    // gA = X1 + B Y3 + 2 gld B
    acc_zero<LP>(acc);
    warp_gemm<LP, false, false, K_FULL, false, KP>(acc, SB, SH, lane);
    double* out = a.gA + (size_t)vg * BS;
#pragma unroll
    for (int mt = 0; mt < NTL; ++mt)
#pragma unroll
      for (int nt = 0; nt < NTL; ++nt) {
        const int row = mt * 8 + lr, col = nt * 8 + 2 * lc;
        const double2 x = *reinterpret_cast<const double2*>(SX + row * LD + col);
        const double2 bb = *reinterpret_cast<const double2*>(SB + row * LD + col);
        frag_pair_store<double, L>(out, row, col, acc[mt][nt][0] + x.x + gl2 * bb.x, acc[mt][nt][1] + x.y + gl2 * bb.y, (L % 2) == 0);
      }
    __syncwarp();                                                          // the record is reused by the next gap of this warp
  }
}

template <typename T, int L>
cudaError_t launch_pegw_fwd(const PegFwdArgs& a, cudaStream_t stream) {
  using C = PegwCfg<T, L>;
  static std::atomic<unsigned char> attr_done[kMaxDevices];
  if (cudaError_t e = ensure_dynamic_smem(cr_pegw_fwd_kernel<T, L>, (int)C::SMEM_F, attr_done); e != cudaSuccess) return e;
  const long long tiles = (long long)((a.n + C::W_F - 2) / (C::W_F - 1)) * a.batch;
  if (tiles <= 0) return cudaSuccess;
  if (tiles > 0x7fffffffLL) return cudaErrorInvalidValue;
  cr_pegw_fwd_kernel<T, L><<<(unsigned)tiles, 32 * C::W_F, C::SMEM_F, stream>>>(a);
  return cudaGetLastError();
}

template <typename T, int L>
cudaError_t launch_pegw_bwd(const PegBwdArgs& a, cudaStream_t stream) {
  using C = PegwCfg<T, L>;
  static std::atomic<unsigned char> attr_done[kMaxDevices];
  if (cudaError_t e = ensure_dynamic_smem(cr_pegw_bwd_kernel<T, L>, (int)C::SMEM_B, attr_done); e != cudaSuccess) return e;
  if (a.n < 2 || a.batch <= 0) return cudaSuccess;
  if (a.gA == nullptr) return cudaErrorInvalidValue;
  const long long total = (long long)(a.n - 1) * a.batch;
  int dev = 0, sms = 148, resident = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, cr_pegw_bwd_kernel<T, L>, 32 * C::W_B, C::SMEM_B) != cudaSuccess || resident < 1) resident = 1;
  long long grid = (total + C::W_B - 1) / C::W_B;
  const long long cap = (long long)sms * resident * 4;
  if (grid > cap) grid = cap;
  cr_pegw_bwd_kernel<T, L><<<(unsigned)grid, 32 * C::W_B, C::SMEM_B, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace crb200
