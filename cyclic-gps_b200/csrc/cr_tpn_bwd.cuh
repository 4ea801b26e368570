// Thread-per-node backward level kernel for small blocks (sizeof(T) * ell^2 <= 256 bytes).
//
// Same contract as cr_level_bwd_kernel (cr_level_bwd.cuh: back-half-solve + selected inverse +
// optional gradient assembly; reference cyclic_gps/cyclic_reduction.py:362-373, :478-501),
// different mapping: ONE THREAD owns one even node e and keeps P = F D^{-1} and Q = G D^{-1}
// in registers for the whole kernel; every other operand is streamed row by row from the
// thread's own shared-memory record with conflict-free 16-byte accesses.  A CTA is one warp
// (32 even nodes); the only data shared between threads are S~_d[e-1] and w~_{e-1}, read from
// the left neighbour's record.
//
// Record layout (stride NS, NS/16 B odd), with the fields that leave as interleaved rows kept
// adjacent so that the interleave (reference interleave(), :181-200) is a plain pair copy:
//   [ A : D_e      -> Sigma_{2e,2e}   | SD : S~_d[e] (Sigma_{2e+1,2e+1})        ]  -> Sd_out rows 2e, 2e+1
//   [ C : G_{e-1}  -> Sigma_{2e,2e-1} | B  : F_e -> Sigma_{2e+1,2e}             ]  -> So_out rows 2e-1, 2e
//   [ SO: S~_o[e-1] (input only) ]
//   [ X : x_e -> w_{2e}               | WT : w~_e (w_{2e+1})                    ]  -> w_out rows 2e, 2e+1
// Record 0 is the left neighbour (deeper-level node e0-1) and only carries SD and WT.
#pragma once
#include "cr_level_bwd.cuh"
#include "cr_tpn_common.cuh"

namespace crb200 {

template <typename T, int L>
struct TpnBwdCfg {
  static constexpr bool ELIGIBLE = (sizeof(T) * L * L <= 256);
  static constexpr int BS = L * L;
  static constexpr int NT = 32;
  static constexpr int A = 0, SD = BS, C = 2 * BS, B = 3 * BS, SO = 4 * BS, X = 5 * BS, WT = 5 * BS + L;
  static constexpr int RAW = 5 * BS + 2 * L;
  static constexpr int NS = record_stride<T>(RAW);
  static constexpr size_t SMEM_W = (size_t)(NT + 1) * NS * sizeof(T);           // per warp
  static constexpr int NW = cmax(1, cmin(CRB200_TPN_WARPS, (int)((220 * 1024) / (SMEM_W * 1 + 1024))));
  static constexpr size_t SMEM = SMEM_W * NW + CRB200_SMEM_PAD;   // CRB200_SMEM_PAD: occupancy experiments only
  static constexpr int MIN_CTAS = cmin(16, cmax(1, (int)((226 * 1024) / (SMEM + 1024))));
};


template <typename T, int L>
__global__ void __launch_bounds__(32 * TpnBwdCfg<T, L>::NW, TpnBwdCfg<T, L>::MIN_CTAS)
cr_tpn_bwd_kernel(const LevelBwdArgs a) {
  using Cf = TpnBwdCfg<T, L>;
  constexpr int BS = Cf::BS, NS = Cf::NS, NT = Cf::NT;
  constexpr unsigned ES = sizeof(T);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5;
  T* S = reinterpret_cast<T*>(smem_raw + (size_t)warp * TpnBwdCfg<T, L>::SMEM_W);
  const unsigned s0 = smem_u32(S);
  const unsigned nsb = NS * ES;
  const unsigned rec1 = s0 + nsb;

  const int m = a.m;
  const int E = (m + 1) >> 1, o = m >> 1, gcnt = (m - 1) >> 1;
  const int tiles = (E + NT - 1) / NT;
  const long long vb = (long long)blockIdx.x * TpnBwdCfg<T, L>::NW + warp;   // virtual block = one warp's tile
  if (vb >= (long long)tiles * a.batch) return;
  const int b = (int)(vb / tiles);
  const int tile = (int)(vb - (long long)b * tiles);
  const int e0 = tile * NT;
  const int nE = cmin(NT, E - e0);
  const bool do_sigma = a.Sd_out != nullptr;
  const bool do_w = a.w_out != nullptr;
  const bool halo = a.G_halo != nullptr;
  const int lane = threadIdx.x & 31;

  // ---------------- stage in: two cp.async groups ----------------
  // group 0 = factors D, F, G and the vectors (needed first: D^{-1}, P, Q, w); group 1 = S~_d, S~_o of the
  // deeper level.  The triangular inverse and the P / Q products overlap the arrival of group 1.
  {
    rec_g2s<T, BS, 1>(rec1 + Cf::A * ES, nsb, static_cast<const T*>(a.D) + ((size_t)b * E + e0) * BS, 0, nE, is_aligned16(a.D));
    const int nF = cmax(0, cmin(NT, o - e0));
    rec_g2s<T, BS, 1>(rec1 + Cf::B * ES, nsb, static_cast<const T*>(a.F) + ((size_t)b * o + e0) * BS, 0, nF, is_aligned16(a.F));
    const int gf = (e0 == 0) ? 1 : 0;
    rec_g2s<T, BS, 1>(rec1 + Cf::C * ES, nsb, static_cast<const T*>(a.G) + ((size_t)b * gcnt + (e0 + gf - 1)) * BS, gf, nE - gf,
                      is_aligned16(a.G));
    if (e0 == 0 && halo)
      rec_g2s<T, BS, 1>(rec1 + Cf::C * ES, nsb, static_cast<const T*>(a.G_halo) + (size_t)b * BS, 0, 1, is_aligned16(a.G_halo));
    const int ilo = (e0 == 0) ? 1 : 0;
    const int nodd = cmin(e0 + NT, o) - (e0 - 1 + ilo);          // deeper node e0-1+i -> record i
    if (do_w) {
      rec_g2s<T, L, 1>(rec1 + Cf::X * ES, nsb, static_cast<const T*>(a.xk) + ((size_t)b * E + e0) * L, 0, nE, is_aligned16(a.xk));
      rec_g2s<T, L, 1>(s0 + Cf::WT * ES, nsb, static_cast<const T*>(a.w_in) + ((size_t)b * o + (e0 - 1 + ilo)) * L, ilo, nodd,
                       is_aligned16(a.w_in));
      if (e0 == 0 && halo)
        rec_g2s<T, L, 1>(s0 + Cf::WT * ES, nsb, static_cast<const T*>(a.w_halo) + (size_t)b * L, 0, 1, is_aligned16(a.w_halo));
    }
    cp_async_commit();
    if (do_sigma) {
      rec_g2s<T, BS, 1>(s0 + Cf::SD * ES, nsb, static_cast<const T*>(a.Sd_in) + ((size_t)b * o + (e0 - 1 + ilo)) * BS, ilo, nodd,
                        is_aligned16(a.Sd_in));
      const int nso = cmin(e0 + NT - 1, o - 1) - (e0 - 1 + ilo);  // link e0-1+i -> record i+1
      rec_g2s<T, BS, 1>(rec1 + Cf::SO * ES, nsb, static_cast<const T*>(a.So_in) + ((size_t)b * (o - 1) + (e0 - 1 + ilo)) * BS, ilo, nso,
                        is_aligned16(a.So_in));
      if (e0 == 0 && halo) {
        rec_g2s<T, BS, 1>(s0 + Cf::SD * ES, nsb, static_cast<const T*>(a.Sd_halo) + (size_t)b * BS, 0, 1, is_aligned16(a.Sd_halo));
        if (o > 0)
          rec_g2s<T, BS, 1>(rec1 + Cf::SO * ES, nsb, static_cast<const T*>(a.So_halo_in) + (size_t)b * BS, 0, 1, is_aligned16(a.So_halo_in));
      }
    }
    cp_async_commit();
    cp_async_wait_group<1>();      // factors and vectors have landed
    __syncwarp();
  }

  // ---------------- stage out, piece by piece ----------------
  const int row_lo = 2 * e0;
  const int nrows = cmin(2 * nE, m - row_lo);
  const int so_plo = (e0 == 0) ? 1 : 0;
  const int nso_rows = cmin(2 * e0 + 2 * nE - 1, m - 1) - (2 * e0 - 1 + so_plo);
  auto out_Sd = [&]() {
    if (do_sigma) {
      T* Sd = static_cast<T*>(a.Sd_out) + (size_t)b * a.strideSd;
      rec_s2g<T, BS, 2>(Sd + (size_t)row_lo * BS, rec1 + Cf::A * ES, nsb, 0, nrows, is_aligned16(Sd));
    }
  };
  auto out_So = [&]() {
    if (do_sigma) {
      if (nso_rows > 0) {
        T* So = static_cast<T*>(a.So_out) + (size_t)b * a.strideSo;
        rec_s2g<T, BS, 2>(So + (size_t)(2 * e0 - 1 + so_plo) * BS, rec1 + Cf::C * ES, nsb, so_plo, nso_rows, is_aligned16(So));
      }
      if (e0 == 0 && halo && a.So_halo_out != nullptr)
        rec_s2g<T, BS, 1>(static_cast<T*>(a.So_halo_out) + (size_t)b * BS, rec1 + Cf::C * ES, nsb, 0, 1, is_aligned16(a.So_halo_out));
    }
  };
  auto out_w = [&]() {
    if (do_w) {
      T* W = static_cast<T*>(a.w_out) + (size_t)b * a.stridew;
      rec_s2g<T, L, 2>(W + (size_t)row_lo * L, rec1 + Cf::X * ES, nsb, 0, nrows, is_aligned16(W));
    }
  };
  const bool early = (a.grad_mode == 0) && (a.variant != CRB200_COPY_ONLY);   // inner levels: results are final as produced

  // variant CRB200_COPY_ONLY (profiling aid): stage in, stage out, no arithmetic -> the memory-system
  // ceiling of this access pattern
  if (a.variant != CRB200_COPY_ONLY) {
  // ---------------- per-node compute ----------------
  T* N = S + (size_t)(lane + 1) * NS;     // this node's record
  const T* Lf = S + (size_t)lane * NS;    // left neighbour's record (S~_d[e-1], w~_{e-1})
  const int e = e0 + lane;
  const bool valid = e < E;
  const bool has_odd = valid && (e < o);
  const bool has_left = valid && (e >= 1 || halo);
  const bool has_so = has_left && has_odd;
  const bool grad = a.grad_mode != 0;
  T gm = T(0), gd = T(1);
  if (grad) {
    gm = (T)(a.gm != nullptr ? a.gm[b] : 0.0);
    gd = (T)(a.gd != nullptr ? a.gd[b] : 0.0);
  }

  // neutral operands for boundary lanes (rare, hence divergent code is fine): no node -> D = I; no odd
  // neighbour -> F = 0, w~_e = 0; no left link -> G = 0 (and the never-loaded left record is cleared)
  if (!valid) smem_fill_identity<T, L>(N + Cf::A);
  if (!has_odd) { smem_fill_zero<T, BS>(N + Cf::B); smem_fill_zero<T, L>(N + Cf::WT); }
  if (!has_left) smem_fill_zero<T, BS>(N + Cf::C);
  if (!valid) smem_fill_zero<T, L>(N + Cf::X);
  if (lane == 0 && e0 == 0 && !halo) { smem_fill_zero<T, L>(S + Cf::WT); }
  __syncwarp();

  T P[L][L], Q[L][L];
  T dxs[L];
  {
    // Di = D^{-1} (lower triangular), in registers
    T Di[L][L];
#pragma unroll
    for (int r = 0; r < L; ++r) {
      T row[L];
      lds_row<T, L>(row, N + Cf::A + r * L);
#pragma unroll
      for (int c = 0; c < L; ++c) Di[r][c] = row[c];   // holds K for now
    }
    T dinv[L];
#pragma unroll
    for (int c = 0; c < L; ++c) dinv[c] = T(1) / Di[c][c];
#pragma unroll
    for (int c = 0; c < L; ++c) {
      // column c of the inverse, top to bottom; K[r][k] for k >= c is still intact in Di[r][k]
      // only for k > c, so keep the column being replaced in a temporary
      T col[L];
      col[c] = dinv[c];
#pragma unroll
      for (int r = c + 1; r < L; ++r) {
        T s = T(0);
#pragma unroll
        for (int k = c; k < r; ++k) s = fma(Di[r][k], col[k], s);
        col[r] = -s * dinv[r];
      }
#pragma unroll
      for (int r = c; r < L; ++r) Di[r][c] = col[r];
    }
    // note: column c of K is overwritten only after every later column's recurrence no longer needs it:
    // the recurrence for column c' > c reads K[r][k] with k >= c' > c.
#pragma unroll
    for (int c = 0; c < L; ++c) dxs[c] = T(0);
    if (do_w) {
      T xs[L];
      lds_row<T, L>(xs, N + Cf::X);
#pragma unroll
      for (int c = 0; c < L; ++c) {
        T s = T(0);
#pragma unroll
        for (int k = c; k < L; ++k) s = fma(Di[k][c], xs[k], s);
        dxs[c] = s;
      }
    }
    if (do_sigma && valid) {
#pragma unroll
      for (int r = 0; r < L; ++r) {
        T row[L];
#pragma unroll
        for (int c = 0; c < L; ++c) {
          T s = T(0);
#pragma unroll
          for (int k = (r > c ? r : c); k < L; ++k) s = fma(Di[k][r], Di[k][c], s);
          row[c] = s;
        }
        sts_row<T, L>(N + Cf::A + r * L, row);    // Di^T Di, the start of Sigma_{2e,2e}
      }
      sched_fence();
    }
#pragma unroll
    for (int r = 0; r < L; ++r) {
      T f[L], g[L];
      lds_row<T, L>(f, N + Cf::B + r * L);
      lds_row<T, L>(g, N + Cf::C + r * L);
#pragma unroll
      for (int c = 0; c < L; ++c) {
        T sp = T(0), sq = T(0);
#pragma unroll
        for (int k = c; k < L; ++k) { sp = fma(f[k], Di[k][c], sp); sq = fma(g[k], Di[k][c], sq); }
        P[r][c] = sp;
        Q[r][c] = sq;
      }
      sched_fence();
    }
  }

  // w_{2e} = Di^T x_e - P^T w~_e - Q^T w~_{e-1}; parked in X (x_e is consumed) and re-read when needed
  if (do_w) {
    T we[L], wl[L], wv[L];
    lds_row<T, L>(we, N + Cf::WT);
    lds_row<T, L>(wl, Lf + Cf::WT);
#pragma unroll
    for (int c = 0; c < L; ++c) {
      T s = dxs[c];
#pragma unroll
      for (int k = 0; k < L; ++k) { s = fma(-P[k][c], we[k], s); s = fma(-Q[k][c], wl[k], s); }
      wv[c] = s;
    }
    if (valid) sts_row<T, L>(N + Cf::X, wv);
  }

  cp_async_wait_group<0>();        // S~_d, S~_o have landed
  __syncwarp();
  if (do_sigma) {
    if (!has_so) smem_fill_zero<T, BS>(N + Cf::SO);
    if (lane == 0 && e0 == 0 && !halo) smem_fill_zero<T, BS>(S + Cf::SD);
    __syncwarp();
    // The three big products run as ROLLED loops over a row / column index that only addresses
    // shared memory (P and Q keep static register indices): 8x less code than full unrolling, which
    // matters because five single-warp CTAs at different program counters share one SM's i-cache.
    // Sigma_{2e+1,2e} = -(S~_d[e] P + S~_o[e-1] Q), row by row into B
    if (has_odd) {
#pragma unroll 1
      for (int r = 0; r < L; ++r) {
        T sig[L], so[L], out[L];
        lds_row<T, L>(sig, N + Cf::SD + r * L);
        lds_row<T, L>(so, N + Cf::SO + r * L);
#pragma unroll
        for (int c = 0; c < L; ++c) out[c] = T(0);
#pragma unroll
        for (int k = 0; k < L; ++k) {
          axpy_row<T, L>(out, -sig[k], P[k]);
          axpy_row<T, L>(out, -so[k], Q[k]);
        }
        sts_row<T, L>(N + Cf::B + r * L, out);
      }
    }
    // Sigma_{2e,2e-1} = -(Q^T S~_d[e-1]^T + P^T S~_o[e-1]), COLUMN by column into C:
    // column c needs row c of S~_d[e-1] (left record) and column c of S~_o[e-1]
    if (has_left) {
#pragma unroll 1
      for (int c = 0; c < L; ++c) {
        T a0[L], socol[L], st[L];
        lds_row<T, L>(a0, Lf + Cf::SD + c * L);
#pragma unroll
        for (int k = 0; k < L; ++k) socol[k] = N[Cf::SO + k * L + c];
#pragma unroll
        for (int r = 0; r < L; ++r) st[r] = T(0);
#pragma unroll
        for (int k = 0; k < L; ++k) {
          axpy_row<T, L>(st, -a0[k], Q[k]);
          axpy_row<T, L>(st, -socol[k], P[k]);
        }
#pragma unroll
        for (int r = 0; r < L; ++r) N[Cf::C + r * L + c] = st[r];
      }
    }
    if (early) {
      __syncwarp();
      out_So();                    // Sigma_off rows and w are final: let them leave while Sigma_{2e,2e} is computed
      out_w();
    }
    // Sigma_{2e,2e} = Di^T Di - S_d^T P - (S_o^T) Q, row by row in place in A:
    // row r needs column r of S_d (B) and row r of Sigma_{2e,2e-1} (C)
    if (valid) {
      T wv[L];
#pragma unroll
      for (int c = 0; c < L; ++c) wv[c] = T(0);
      if (grad && do_w) lds_row<T, L>(wv, N + Cf::X);
#pragma unroll 1
      for (int r = 0; r < L; ++r) {
        T acc[L], sdcol[L], st[L];
        lds_row<T, L>(acc, N + Cf::A + r * L);
        lds_row<T, L>(st, N + Cf::C + r * L);
#pragma unroll
        for (int k = 0; k < L; ++k) sdcol[k] = N[Cf::B + k * L + r];
#pragma unroll
        for (int k = 0; k < L; ++k) {
          axpy_row<T, L>(acc, -sdcol[k], P[k]);
          axpy_row<T, L>(acc, -st[k], Q[k]);
        }
        if (grad) {
          T wr = T(0);
#pragma unroll
          for (int c = 0; c < L; ++c) wr = (c == r) ? wv[c] : wr;
#pragma unroll
          for (int c = 0; c < L; ++c) acc[c] = gd * acc[c] - gm * wr * wv[c];
        }
        sts_row<T, L>(N + Cf::A + r * L, acc);
      }
    }
  }
  __syncwarp();   // neighbours are done reading this record's SD / WT

  if (grad) {
    T wv[L], we[L], wl[L];
#pragma unroll
    for (int c = 0; c < L; ++c) { wv[c] = T(0); we[c] = T(0); wl[c] = T(0); }
    if (do_w) {
      if (valid) lds_row<T, L>(wv, N + Cf::X);
      if (has_odd) lds_row<T, L>(we, N + Cf::WT);
      if (has_left) lds_row<T, L>(wl, Lf + Cf::WT);
    }
    if (do_sigma) {
      if (has_odd) {
#pragma unroll
        for (int r = 0; r < L; ++r) {
          T v[L];
          lds_row<T, L>(v, N + Cf::SD + r * L);
#pragma unroll
          for (int c = 0; c < L; ++c) v[c] = gd * v[c] - gm * we[r] * we[c];
          sts_row<T, L>(N + Cf::SD + r * L, v);
          lds_row<T, L>(v, N + Cf::B + r * L);
#pragma unroll
          for (int c = 0; c < L; ++c) v[c] = T(2) * (gd * v[c] - gm * we[r] * wv[c]);
          sts_row<T, L>(N + Cf::B + r * L, v);
        }
      }
      if (has_left) {
#pragma unroll
        for (int r = 0; r < L; ++r) {
          T v[L];
          lds_row<T, L>(v, N + Cf::C + r * L);
#pragma unroll
          for (int c = 0; c < L; ++c) v[c] = T(2) * (gd * v[c] - gm * wv[r] * wl[c]);
          sts_row<T, L>(N + Cf::C + r * L, v);
        }
      }
    }
    __syncwarp();   // every lane has read its neighbours' untransformed w~ before anyone rescales it
    if (do_w) {
#pragma unroll
      for (int c = 0; c < L; ++c) { wv[c] = T(2) * gm * wv[c]; we[c] = T(2) * gm * we[c]; }
      if (valid) sts_row<T, L>(N + Cf::X, wv);
      if (has_odd) sts_row<T, L>(N + Cf::WT, we);
    }
  }
  __syncwarp();

  }

  cp_async_wait_group<0>();        // (copy-only path) everything staged
  __syncwarp();
  // ---------------- stage out (what has not left yet) ----------------
  out_Sd();
  if (!early || !do_sigma) {
    out_So();
    out_w();
  }
}

template <typename T, int L>
cudaError_t launch_tpn_bwd(const LevelBwdArgs& a, cudaStream_t stream) {
  using C = TpnBwdCfg<T, L>;
  static bool attr_done[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 64 && !attr_done[dev]) {
    cudaError_t e = cudaFuncSetAttribute(cr_tpn_bwd_kernel<T, L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
    if (e != cudaSuccess) return e;
    attr_done[dev] = true;
  }
  const int E = (a.m + 1) / 2;
  const long long tiles = (E + C::NT - 1) / C::NT;
  const long long total = tiles * a.batch;
  if (total <= 0) return cudaSuccess;
  const long long grid = (total + C::NW - 1) / C::NW;
  if (grid > 0x7fffffffLL) return cudaErrorInvalidValue;
  cr_tpn_bwd_kernel<T, L><<<(unsigned)grid, 32 * C::NW, C::SMEM, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace crb200
