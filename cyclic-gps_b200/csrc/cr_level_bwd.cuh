// One cyclic-reduction level, backward direction (deepest level first): back-half-solve and
// selected inverse fused, with optional gradient assembly at the top level.
//
// Replaces, per level, the reference's backhalfsolve step (cyclic_gps/cyclic_reduction.py
// :362-373, U_Tx :63-87, interleave :181-200) and inverse_blocks step (:478-501, SigU :90-136,
// UtV_diags :139-178).  For even node e of a level with factors D_e, F_e, G_{e-1}:
//   Di = D_e^{-1};  P = F_e Di;  Q = G_{e-1} Di                              (:484-490)
//   S_d[e]   = -(S~_d[e] P + S~_o[e-1] Q)        = Sigma_{2e+1,2e}            (:493, SigU diag)
//   S_o[e-1] = -(S~_d[e-1] Q + S~_o[e-1]^T P)    = Sigma_{2e-1,2e}            (:493, SigU upper)
//   Sigma_{2e,2e} = Di^T Di - P^T S_d[e] - Q^T S_o[e-1]                       (:496)
//   w_{2e} = Di^T x_e - P^T w~_e - Q^T w~_{e-1}                               (:366-369)
// and the surviving (odd) nodes are copied through (interleave, :498-501 / :373).
// Outputs of the level: Sd_out (m blocks), So_out (m-1 lower blocks, So_out[i] = Sigma_{i+1,i}),
// w_out (m rows).
//
// Gradient assembly (closed forms verified against reference autograd): with
// per-series cotangents gm (for mahal) and gd (for logdet) the top level writes
//   gR_i = gd Sigma_ii - gm w_i w_i^T,  gO_i = 2 gd Sigma_{i+1,i} - 2 gm w_{i+1} w_i^T,  gx = 2 gm w.
//
// Work mapping: one group of LG lanes per even node, lane r = row r; a CTA owns NG consecutive
// even nodes of one series; no exchange between groups is needed.  Sigma~ and w~ of the deeper
// level are staged directly into the odd slots of the output staging area, so the interleave
// costs nothing.
#pragma once
#include "cr_common.cuh"

namespace crb200 {

using LevelBwdArgs = ::crb200_bwd_args;   // include/crb200.h

template <typename T, int L>
struct BwdCfg {
  static constexpr int LG = GroupLanes<L>::value;
  static constexpr int BS = L * L;
  static constexpr int NODE_ELEMS = 8 * BS + 3 * L;
  static constexpr int GRAN = 32 / LG;
  static constexpr int NG_FIT = (100 * 1024 - (BS + L) * (int)sizeof(T)) / (NODE_ELEMS * (int)sizeof(T));
  static constexpr int NG_RAW = cmin(kThreads / LG, NG_FIT);
  static constexpr int NG = cmax(cmax(1, GRAN), (NG_RAW / GRAN) * GRAN);
  static constexpr int THREADS = NG * LG;
  static constexpr size_t D_OFF = 0;
  static constexpr size_t F_OFF = align16(D_OFF + sizeof(T) * NG * BS);
  static constexpr size_t G_OFF = align16(F_OFF + sizeof(T) * NG * BS);
  static constexpr size_t SD_OFF = align16(G_OFF + sizeof(T) * NG * BS);
  static constexpr size_t SO_OFF = align16(SD_OFF + sizeof(T) * (2 * NG + 1) * BS);
  static constexpr size_t X_OFF = align16(SO_OFF + sizeof(T) * (2 * NG) * BS);
  static constexpr size_t W_OFF = align16(X_OFF + sizeof(T) * NG * L);
  static constexpr size_t SMEM = align16(W_OFF + sizeof(T) * (2 * NG + 1) * L);
};

template <typename T, int L>
__global__ void __launch_bounds__(BwdCfg<T, L>::THREADS)
cr_level_bwd_kernel(const LevelBwdArgs a) {
  using C = BwdCfg<T, L>;
  constexpr int LG = C::LG, NG = C::NG, BS = C::BS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* sD = reinterpret_cast<T*>(smem_raw + C::D_OFF);
  T* sF = reinterpret_cast<T*>(smem_raw + C::F_OFF);
  T* sG = reinterpret_cast<T*>(smem_raw + C::G_OFF);
  T* sSd = reinterpret_cast<T*>(smem_raw + C::SD_OFF);   // position p <-> Sd_out row 2e0-1+p
  T* sSo = reinterpret_cast<T*>(smem_raw + C::SO_OFF);   // position p <-> So_out row 2e0-1+p
  T* sx = reinterpret_cast<T*>(smem_raw + C::X_OFF);
  T* sw = reinterpret_cast<T*>(smem_raw + C::W_OFF);     // position p <-> w_out row 2e0-1+p

  const int m = a.m;
  const int E = (m + 1) >> 1, o = m >> 1, gcnt = (m - 1) >> 1;
  const int tiles = (E + NG - 1) / NG;
  const int b = blockIdx.x / tiles;
  const int tile = blockIdx.x - b * tiles;
  const int e0 = tile * NG;
  const int nE = cmin(NG, E - e0);
  const bool do_sigma = a.Sd_out != nullptr;
  const bool do_w = a.w_out != nullptr;
  const bool halo = a.G_halo != nullptr;

  // ---------------- stage in ----------------
  {
    tile_g2s<T, BS, 1>(sD, static_cast<const T*>(a.D) + ((size_t)b * E + e0) * BS, nE, is_aligned16(a.D));
    const int nF = cmax(0, cmin(NG, o - e0));
    tile_g2s<T, BS, 1>(sF, static_cast<const T*>(a.F) + ((size_t)b * o + e0) * BS, nF, is_aligned16(a.F));
    const int gf = (e0 == 0) ? 1 : 0;
    tile_g2s<T, BS, 1>(sG + (size_t)gf * BS, static_cast<const T*>(a.G) + ((size_t)b * gcnt + (e0 + gf - 1)) * BS,
                       nE - gf, is_aligned16(a.G));
    if (e0 == 0 && halo) tile_g2s<T, BS, 1>(sG, static_cast<const T*>(a.G_halo) + (size_t)b * BS, 1, is_aligned16(a.G_halo));
    // deeper-level odd nodes e0-1 .. e0+NG-1 go to even positions 0,2,..,2NG
    const int ilo = (e0 == 0) ? 1 : 0;
    const int nodd = cmin(e0 + NG, o) - (e0 - 1 + ilo);      // nodes e0-1+ilo .. min(e0+NG-1, o-1)
    if (do_sigma) {
      tile_g2s<T, BS, 2>(sSd + (size_t)(2 * ilo) * BS, static_cast<const T*>(a.Sd_in) + ((size_t)b * o + (e0 - 1 + ilo)) * BS,
                         nodd, is_aligned16(a.Sd_in));
      const int nso = cmin(e0 + NG - 1, o - 1) - (e0 - 1 + ilo);   // links e0-1+ilo .. min(e0+NG-2, o-2)
      tile_g2s<T, BS, 2>(sSo + (size_t)(2 * ilo) * BS, static_cast<const T*>(a.So_in) + ((size_t)b * (o - 1) + (e0 - 1 + ilo)) * BS,
                         nso, is_aligned16(a.So_in));
      if (e0 == 0 && halo) {
        tile_g2s<T, BS, 1>(sSd, static_cast<const T*>(a.Sd_halo) + (size_t)b * BS, 1, is_aligned16(a.Sd_halo));
        if (o > 0) tile_g2s<T, BS, 1>(sSo, static_cast<const T*>(a.So_halo_in) + (size_t)b * BS, 1, is_aligned16(a.So_halo_in));
      }
    }
    if (do_w) {
      tile_g2s<T, L, 1>(sx, static_cast<const T*>(a.xk) + ((size_t)b * E + e0) * L, nE, is_aligned16(a.xk));
      tile_g2s<T, L, 2>(sw + (size_t)(2 * ilo) * L, static_cast<const T*>(a.w_in) + ((size_t)b * o + (e0 - 1 + ilo)) * L,
                        nodd, is_aligned16(a.w_in));
      if (e0 == 0 && halo) tile_g2s<T, L, 1>(sw, static_cast<const T*>(a.w_halo) + (size_t)b * L, 1, is_aligned16(a.w_halo));
    }
    cp_async_wait_all();
    __syncthreads();
  }

  // ---------------- per-node compute ----------------
  const int g = threadIdx.x / LG;
  const int r = threadIdx.x - g * LG;
  const int e = e0 + g;
  const bool valid = e < E;
  const bool rowok = r < L;
  const int rr = rowok ? r : 0;
  const bool has_odd = valid && (e < o);
  const bool has_left = valid && (e >= 1 || halo);
  const bool has_so = has_left && has_odd;
  T* Db = sD + (size_t)g * BS;
  T* Fb = sF + (size_t)g * BS;
  T* Gb = sG + (size_t)g * BS;
  T* SdL = sSd + (size_t)(2 * g) * BS;       // S~_d[e-1]
  T* SdE = sSd + (size_t)(2 * g + 1) * BS;   // Sigma_{2e,2e}   (output)
  T* SdR = sSd + (size_t)(2 * g + 2) * BS;   // S~_d[e]
  T* SoL = sSo + (size_t)(2 * g) * BS;       // S~_o[e-1] in -> Sigma_{2e,2e-1} out
  T* SoR = sSo + (size_t)(2 * g + 1) * BS;   // Sigma_{2e+1,2e} out

  // Di = D^{-1}, row r per lane (row-oriented back substitution on e_r^T D^{-1})
  T di[L];
  {
    T s[L];
#pragma unroll
    for (int c = 0; c < L; ++c) s[c] = (c == rr) ? T(1) : T(0);
    if (valid) {
#pragma unroll
      for (int k = L - 1; k >= 0; --k) {
        T drow[L];
        lds_row<T, L>(drow, Db + k * L);
        di[k] = s[k] / drow[k];
#pragma unroll
        for (int c = 0; c < k; ++c) s[c] = fma(-di[k], drow[c], s[c]);
      }
    } else {
#pragma unroll
      for (int c = 0; c < L; ++c) di[c] = s[c];
    }
  }
  T f[L], gq[L];
#pragma unroll
  for (int c = 0; c < L; ++c) { f[c] = T(0); gq[c] = T(0); }
  if (has_odd) lds_row<T, L>(f, Fb + rr * L);
  if (has_left) lds_row<T, L>(gq, Gb + rr * L);
  __syncwarp();
  if (valid && rowok) sts_row<T, L>(Db + r * L, di);
  __syncwarp();
  {
    T p[L], q[L];
#pragma unroll
    for (int c = 0; c < L; ++c) { p[c] = T(0); q[c] = T(0); }
    if (has_odd) row_times_mat<T, L>(p, f, Db);
    if (has_left) row_times_mat<T, L>(q, gq, Db);
    if (valid && rowok) { sts_row<T, L>(Fb + r * L, p); sts_row<T, L>(Gb + r * L, q); }
  }
  __syncwarp();   // Di, P, Q visible inside the group

  T pcol[L], qcol[L], dcol[L];
#pragma unroll
  for (int k = 0; k < L; ++k) {
    pcol[k] = valid ? Fb[k * L + rr] : T(0);
    qcol[k] = valid ? Gb[k * L + rr] : T(0);
    dcol[k] = valid ? Db[k * L + rr] : T(0);
  }

  if (do_sigma) {
    T sd[L], st[L];
#pragma unroll
    for (int c = 0; c < L; ++c) { sd[c] = T(0); st[c] = T(0); }
    if (has_odd) {
      T sig[L];
      lds_row<T, L>(sig, SdR + rr * L);
      row_times_mat<T, L>(sd, sig, Fb);
    }
    if (has_so) {
      T so[L];
      lds_row<T, L>(so, SoL + rr * L);
      row_times_mat<T, L>(sd, so, Gb);
    }
    if (has_left) row_times_matT<T, L>(st, qcol, SdL);   // Q^T S~_d[e-1]^T
    if (has_so) row_times_mat<T, L>(st, pcol, SoL);
#pragma unroll
    for (int c = 0; c < L; ++c) { sd[c] = -sd[c]; st[c] = -st[c]; }
    __syncwarp();   // every lane is done reading S~_o[e-1]
    if (has_left && rowok) sts_row<T, L>(SoL + r * L, st);
    if (has_odd && rowok) sts_row<T, L>(SoR + r * L, sd);
    __syncwarp();
    T se[L];
#pragma unroll
    for (int c = 0; c < L; ++c) se[c] = T(0);
    row_times_mat<T, L>(se, dcol, Db);
#pragma unroll
    for (int k = 0; k < L; ++k) { pcol[k] = -pcol[k]; qcol[k] = -qcol[k]; }
    if (has_odd) row_times_mat<T, L>(se, pcol, SoR);
    if (has_left) row_times_matT<T, L>(se, qcol, SoL);
    if (valid && rowok) sts_row<T, L>(SdE + r * L, se);
  } else {
#pragma unroll
    for (int k = 0; k < L; ++k) { pcol[k] = -pcol[k]; qcol[k] = -qcol[k]; }
  }
  if (do_w && valid) {
    T xs[L];
    lds_row<T, L>(xs, sx + (size_t)g * L);
    T acc = T(0);
#pragma unroll
    for (int k = 0; k < L; ++k) acc = fma(dcol[k], xs[k], acc);
    if (has_odd) {
      T we[L];
      lds_row<T, L>(we, sw + (size_t)(2 * g + 2) * L);
#pragma unroll
      for (int k = 0; k < L; ++k) acc = fma(pcol[k], we[k], acc);
    }
    if (has_left) {
      T wl[L];
      lds_row<T, L>(wl, sw + (size_t)(2 * g) * L);
#pragma unroll
      for (int k = 0; k < L; ++k) acc = fma(qcol[k], wl[k], acc);
    }
    if (rowok) sw[(size_t)(2 * g + 1) * L + r] = acc;
  }
  __syncthreads();

  // ---------------- gradient assembly (top level) ----------------
  const int row_lo = 2 * e0;                                    // first Sd / w row of this tile
  const int nrows = cmin(2 * nE, m - row_lo);                   // Sd / w rows produced
  const int so_plo = (e0 == 0) ? 1 : 0;                         // first So position that is a real row
  const int nso_rows = cmin(2 * e0 + 2 * nE - 1, m - 1) - (2 * e0 - 1 + so_plo);
  if (a.grad_mode) {
    const T gm = (T)(a.gm != nullptr ? a.gm[b] : 0.0);
    const T gd = (T)(a.gd != nullptr ? a.gd[b] : 0.0);
    const bool use_w = do_w;
    if (do_sigma) {
      for (int i = threadIdx.x; i < nrows * BS; i += blockDim.x) {
        const int p = 1 + i / BS, rc = i % BS, rw = rc / L, cl = rc % L;
        T v = gd * sSd[(size_t)p * BS + rc];
        if (use_w) v -= gm * sw[p * L + rw] * sw[p * L + cl];
        sSd[(size_t)p * BS + rc] = v;
      }
      const int pstart = (e0 == 0 && halo) ? 0 : so_plo;
      const int pend = so_plo + nso_rows;
      for (int i = threadIdx.x + pstart * BS; i < pend * BS; i += blockDim.x) {
        const int p = i / BS, rc = i % BS, rw = rc / L, cl = rc % L;
        T v = T(2) * gd * sSo[(size_t)p * BS + rc];
        if (use_w) v -= T(2) * gm * sw[(p + 1) * L + rw] * sw[p * L + cl];
        sSo[(size_t)p * BS + rc] = v;
      }
    }
    __syncthreads();
    if (do_w) {
      for (int i = threadIdx.x; i < nrows * L; i += blockDim.x) sw[L + i] = T(2) * gm * sw[L + i];
    }
    __syncthreads();
  }

  // ---------------- stage out ----------------
  if (do_sigma) {
    T* Sd = static_cast<T*>(a.Sd_out) + (size_t)b * a.strideSd + (size_t)row_lo * BS;
    tile_s2g<T, BS, 1>(Sd, sSd + BS, nrows, is_aligned16(static_cast<T*>(a.Sd_out) + (size_t)b * a.strideSd));
    if (nso_rows > 0) {
      T* So = static_cast<T*>(a.So_out) + (size_t)b * a.strideSo + (size_t)(2 * e0 - 1 + so_plo) * BS;
      tile_s2g<T, BS, 1>(So, sSo + (size_t)so_plo * BS, nso_rows, is_aligned16(static_cast<T*>(a.So_out) + (size_t)b * a.strideSo));
    }
    if (e0 == 0 && halo && a.So_halo_out != nullptr)
      tile_s2g<T, BS, 1>(static_cast<T*>(a.So_halo_out) + (size_t)b * BS, sSo, 1, is_aligned16(a.So_halo_out));
  }
  if (do_w) {
    T* W = static_cast<T*>(a.w_out) + (size_t)b * a.stridew + (size_t)row_lo * L;
    tile_s2g<T, L, 1>(W, sw + L, nrows, is_aligned16(static_cast<T*>(a.w_out) + (size_t)b * a.stridew));
  }
}

template <typename T, int L>
cudaError_t launch_level_bwd(const LevelBwdArgs& a, cudaStream_t stream) {
  using C = BwdCfg<T, L>;
  static std::atomic<unsigned char> attr_done[kMaxDevices];
  if (cudaError_t e = ensure_dynamic_smem(cr_level_bwd_kernel<T, L>, (int)C::SMEM, attr_done); e != cudaSuccess) return e;
  const int E = (a.m + 1) / 2;
  const long long tiles = (E + C::NG - 1) / C::NG;
  const long long grid = tiles * a.batch;
  if (grid <= 0) return cudaSuccess;
  if (grid > 0x7fffffffLL) return cudaErrorInvalidValue;
  cr_level_bwd_kernel<T, L><<<(unsigned)grid, C::THREADS, C::SMEM, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace crb200
