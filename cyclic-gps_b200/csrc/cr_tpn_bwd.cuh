// Thread-per-node backward level kernel for small blocks (sizeof(T) * ell^2 <= CRB200_TPN_MAX_BLOCK_BYTES = 400 bytes).
//
// Same contract as cr_level_bwd_kernel (cr_level_bwd.cuh: back-half-solve + selected inverse +
// optional gradient assembly; reference cyclic_gps/cyclic_reduction.py:362-373, :478-501),
// different mapping: ONE THREAD owns one even node e and keeps P = F D^{-1}, Q = G D^{-1} and the
// lower triangle of Sigma_{2e,2e} in registers for the whole kernel; every other operand is streamed
// row by row from the thread's own shared-memory record with conflict-free 16-byte accesses.  A CTA
// is one warp (32 even nodes); the only data shared between threads are S~_d[e-1] and w~_{e-1}, read
// from the left neighbour's record.
//
// COMPACT record layout (stride NS, NS/16 B odd; 848 B at ell = 8 fp32), THREE blocks per node, each used twice
// (chosen when five blocks per node would leave fewer than 8 CTAs per SM: fp32 ell = 7, 8 and fp64 ell = 5):
//   [ A : D_e, then S~_d[e] (= Sigma_{2e+1,2e+1})                 ]  -> Sd_out row 2e+1
//   [ C : G_{e-1}        -> Sigma_{2e,2e-1} | B : F_e, then S~_o[e-1] -> Sigma_{2e+1,2e} ]  -> So_out rows 2e-1, 2e
//   [ X : x_e -> w_{2e}                     | WT : w~_e (w_{2e+1})    ]  -> w_out rows 2e, 2e+1
// D is turned into D^{-1} in registers at once and S~_d is staged over it; F is consumed into P and S~_o
// is staged over it; both arrive while P, Q and w are being computed.  Sigma_{2e,2e} never touches shared
// memory: it starts as D^{-T} D^{-1} in registers, is finished from the rows of Sigma_{2e+1,2e} and
// Sigma_{2e,2e-1}, and its rows go from registers straight to global memory.  Three blocks instead of five
// per node means 8 resident single-warp CTAs per SM instead of 5, which is what this latency-bound kernel
// needs: 2.71 -> 2.18 ms for the level-0 launch of configs[1] (profiles/r1_summary.md).
// Small blocks keep FIVE blocks per node, [ A | SD | C | B | SO | X | WT ]: everything is requested up front,
// Sigma_{2e,2e} is written into A, and (A, SD), (C, B), (X, WT) leave as interleaved pairs.
// Record 0 is the left neighbour (deeper-level node e0-1) and only carries S~_d and WT.
// Packed lower triangles (crb200_bwd_args.tri, TriPack in cr_tpn_common.cuh; COMPACT layout, float32 ell = 8): D and S~_d arrive as 36 elements at
// the start of the A slot; S~_d is expanded to full symmetric rows in place once it has landed (the products read rows of it), and on inner
// levels both halves of Sigma_d leave packed straight from registers (the odd rows at the expansion, the even rows at the end).
#pragma once
#include "cr_level_bwd.cuh"
#include "cr_tpn_common.cuh"

namespace crb200 {

template <typename T, int L>
struct TpnBwdCfg {
  static constexpr bool ELIGIBLE = (sizeof(T) * L * L <= CRB200_TPN_MAX_BLOCK_BYTES);
  static constexpr int BS = L * L;
  static constexpr int NT = 32;
  // COMPACT (three blocks per node, A and B slots used twice) when five blocks per node would leave fewer than
  // 8 CTAs on an SM (fp32 ell = 7, 8; fp64 ell = 5); otherwise S~_d and S~_o get slots of their own, everything is
  // requested up front and Sigma_{2e,2e} leaves through the A slot (measured: at ell = 4 fp32, 17 CTAs/SM with five
  // blocks beat 26 CTAs/SM with three blocks and three dependent load phases, 5.34 vs 6.03 ms on the long series)
  static constexpr int CTAS5 = (int)((228 * 1024) / ((size_t)(NT + 1) * record_stride<T>(5 * BS + 2 * L, BS) * sizeof(T) + 1024));
  static constexpr bool COMPACT = CTAS5 < 8;
  static constexpr int A = 0, SD = COMPACT ? 0 : BS, C = COMPACT ? BS : 2 * BS, B = C + BS, SO = COMPACT ? B : 4 * BS;
  static constexpr int X = (COMPACT ? 3 : 5) * BS, WT = X + L;
  static constexpr int RAW = X + 2 * L;
  static constexpr int NS = record_stride<T>(RAW, BS);
  static constexpr size_t SMEM_W = (size_t)(NT + 1) * NS * sizeof(T);           // per warp
  static constexpr int NW = cmax(1, cmin(CRB200_TPN_WARPS, (int)((220 * 1024) / (SMEM_W * 1 + 1024))));
  static constexpr size_t SMEM = SMEM_W * NW + CRB200_SMEM_PAD;   // CRB200_SMEM_PAD: occupancy experiments only
  static constexpr int MIN_CTAS = cmin(16, cmax(1, (int)((228 * 1024) / (SMEM + 1024))));
};


// One tile (NT even nodes of series b starting at node tile * NT) by one warp; see tpn_fwd_tile.
template <typename T, int L>
__device__ __forceinline__ void tpn_bwd_tile(const LevelBwdArgs& a, unsigned char* smem_warp, const int b, const int tile) {
  using Cf = TpnBwdCfg<T, L>;
  constexpr int BS = Cf::BS, NS = Cf::NS, NT = Cf::NT;
  constexpr unsigned ES = sizeof(T);
  T* S = reinterpret_cast<T*>(smem_warp);
  const unsigned s0 = smem_u32(S);
  const unsigned nsb = NS * ES;
  const unsigned rec1 = s0 + nsb;

  const int m = a.m;
  const int E = (m + 1) >> 1, o = m >> 1, gcnt = (m - 1) >> 1;
  const int e0 = tile * NT;
  const int nE = cmin(NT, E - e0);
  const bool do_sigma = a.Sd_out != nullptr;
  const bool do_w = a.w_out != nullptr;
  const bool halo = a.G_halo != nullptr;
  const int lane = threadIdx.x & 31;
  const int nF = cmax(0, cmin(NT, o - e0));                      // odd neighbours e0 .. of this tile
  const int ilo = (e0 == 0) ? 1 : 0;
  const int nodd = cmin(e0 + NT, o) - (e0 - 1 + ilo);            // deeper node e0-1+i -> record i
  // packed lower triangles (crb200_bwd_args.tri, cr_tpn_common.cuh): D and S~_d arrive packed / Sigma_d leaves packed (inner levels)
  using TP = TriPack<T, L>;
  constexpr bool TRI = TP::OK && Cf::COMPACT;
  constexpr int PKS = TRI ? TP::PKS : BS;
  const bool tri_in = TRI && (a.tri & 1) != 0;
  const bool tri_out = TRI && (a.tri & 2) != 0;

  // ---------------- stage in, first group: factors D, F, G and the vectors ----------------
  {
    if (tri_in) rec_g2s<T, PKS, 1, true>(rec1 + Cf::A * ES, nsb, static_cast<const T*>(a.D) + ((size_t)b * E + e0) * PKS, 0, nE, is_aligned16(a.D));
    else rec_g2s<T, BS, 1>(rec1 + Cf::A * ES, nsb, static_cast<const T*>(a.D) + ((size_t)b * E + e0) * BS, 0, nE, is_aligned16(a.D));
    rec_g2s<T, BS, 1>(rec1 + Cf::B * ES, nsb, static_cast<const T*>(a.F) + ((size_t)b * o + e0) * BS, 0, nF, is_aligned16(a.F));
    const int gf = (e0 == 0) ? 1 : 0;
    rec_g2s<T, BS, 1>(rec1 + Cf::C * ES, nsb, static_cast<const T*>(a.G) + ((size_t)b * gcnt + (e0 + gf - 1)) * BS, gf, nE - gf,
                      is_aligned16(a.G));
    if (e0 == 0 && halo)
      rec_g2s<T, BS, 1>(rec1 + Cf::C * ES, nsb, static_cast<const T*>(a.G_halo) + (size_t)b * BS, 0, 1, is_aligned16(a.G_halo));
    if (do_w) {
      rec_g2s<T, L, 1>(rec1 + Cf::X * ES, nsb, static_cast<const T*>(a.xk) + ((size_t)b * E + e0) * L, 0, nE, is_aligned16(a.xk));
      rec_g2s<T, L, 1>(s0 + Cf::WT * ES, nsb, static_cast<const T*>(a.w_in) + ((size_t)b * o + (e0 - 1 + ilo)) * L, ilo, nodd,
                       is_aligned16(a.w_in));
      if (e0 == 0 && halo)
        rec_g2s<T, L, 1>(s0 + Cf::WT * ES, nsb, static_cast<const T*>(a.w_halo) + (size_t)b * L, 0, 1, is_aligned16(a.w_halo));
    }
    cp_async_commit();
  }
  // S~_d of the deeper level (record i <- deeper node e0-1+i); COMPACT: second use of the A slots
  auto stage_Sd = [&]() {
    if (do_sigma) {
      if (tri_in)
        rec_g2s<T, PKS, 1, true>(s0 + Cf::SD * ES, nsb, static_cast<const T*>(a.Sd_in) + ((size_t)b * o + (e0 - 1 + ilo)) * PKS, ilo, nodd,
                           is_aligned16(a.Sd_in));
      else
        rec_g2s<T, BS, 1>(s0 + Cf::SD * ES, nsb, static_cast<const T*>(a.Sd_in) + ((size_t)b * o + (e0 - 1 + ilo)) * BS, ilo, nodd,
                          is_aligned16(a.Sd_in));
      if (e0 == 0 && halo)
        rec_g2s<T, BS, 1>(s0 + Cf::SD * ES, nsb, static_cast<const T*>(a.Sd_halo) + (size_t)b * BS, 0, 1, is_aligned16(a.Sd_halo));
    }
    cp_async_commit();
  };
  // S~_o of the deeper level (link e0-1+i -> record i+1); COMPACT: second use of the B slots
  auto stage_So = [&]() {
    if (do_sigma) {
      const int nso = cmin(e0 + NT - 1, o - 1) - (e0 - 1 + ilo);
      rec_g2s<T, BS, 1>(rec1 + Cf::SO * ES, nsb, static_cast<const T*>(a.So_in) + ((size_t)b * (o - 1) + (e0 - 1 + ilo)) * BS, ilo, nso,
                        is_aligned16(a.So_in));
      if (e0 == 0 && halo && o > 0)
        rec_g2s<T, BS, 1>(rec1 + Cf::SO * ES, nsb, static_cast<const T*>(a.So_halo_in) + (size_t)b * BS, 0, 1, is_aligned16(a.So_halo_in));
    }
    cp_async_commit();
  };

  // ---------------- stage out, piece by piece ----------------
  const int row_lo = 2 * e0;
  const int nrows = cmin(2 * nE, m - row_lo);
  const int so_plo = (e0 == 0) ? 1 : 0;
  const int nso_rows = cmin(2 * e0 + 2 * nE - 1, m - 1) - (2 * e0 - 1 + so_plo);
  T* const gSd = do_sigma ? static_cast<T*>(a.Sd_out) + (size_t)b * a.strideSd : nullptr;
  const bool sd_vec = is_aligned16(gSd);
  auto out_Sd = [&]() {
    if (do_sigma) {
      if constexpr (Cf::COMPACT) { // Sigma_{2e+1,2e+1} only (record i+1 -> Sd_out row 2(e0+i)+1); the even rows left from registers
        if (!tri_out) rec_s2g_strided<T, BS, 2>(gSd + (size_t)(row_lo + 1) * BS, rec1 + Cf::SD * ES, nsb, 0, nF, sd_vec);   // (packed: left at the expansion)
      } else                         // A and SD are adjacent: rows 2e, 2e+1 leave as pairs
        rec_s2g<T, BS, 2>(gSd + (size_t)row_lo * BS, rec1 + Cf::A * ES, nsb, 0, nrows, sd_vec);
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

  T* N = S + (size_t)(lane + 1) * NS;     // this node's record
  const T* Lf = S + (size_t)lane * NS;    // left neighbour's record (S~_d[e-1], w~_{e-1})
  const int e = e0 + lane;
  const bool valid = e < E;

  // variant CRB200_COPY_ONLY (profiling aid): stage in, stage out, no arithmetic -> the memory-system
  // ceiling of this access pattern (same bytes, same two uses of the A and B slots)
  if (a.variant == CRB200_COPY_ONLY) {
    if constexpr (Cf::COMPACT) {
      cp_async_wait_group<0>();
      __syncwarp();
      if (do_sigma && valid)
        for (int r = 0; r < L; ++r) {
          T z[L];
          lds_row<T, L>(z, N + Cf::A + r * L);
          stg_row<T, L>(gSd + (size_t)(2 * e) * BS + r * L, z, sd_vec);
        }
      __syncwarp();
    }
    stage_Sd();
    stage_So();
    cp_async_wait_group<0>();
    __syncwarp();
    out_Sd(); out_So(); out_w();
    return;
  }

  // ---------------- per-node compute ----------------
  const bool has_odd = valid && (e < o);
  const bool has_left = valid && (e >= 1 || halo);
  const bool has_so = has_left && has_odd;
  const bool grad = a.grad_mode != 0;
  const bool early = !grad;               // inner levels: results are final as produced
  T gm = T(0), gd = T(1);
  if (grad) {
    gm = (T)(a.gm != nullptr ? a.gm[b] : 0.0);
    gd = (T)(a.gd != nullptr ? a.gd[b] : 0.0);
  }
  if constexpr (Cf::COMPACT) {
    cp_async_wait_group<0>();      // factors and vectors have landed
  } else {
    stage_Sd();                    // own slots: S~_d, S~_o are requested now and arrive during D^{-1}, P, Q
    stage_So();
    cp_async_wait_group<2>();      // factors and vectors have landed
  }
  __syncwarp();

  // neutral operands for boundary lanes (rare, hence divergent code is fine): no node -> D = I; no odd
  // neighbour -> F = 0, w~_e = 0; no left link -> G = 0 (and the never-loaded left record is cleared)
  if (!valid) {
    if constexpr (TRI) { if (tri_in) smem_fill_identity_tri<T, L>(N + Cf::A); else smem_fill_identity<T, L>(N + Cf::A); }
    else smem_fill_identity<T, L>(N + Cf::A);
  }
  if (!has_odd) { smem_fill_zero<T, BS>(N + Cf::B); smem_fill_zero<T, L>(N + Cf::WT); }
  if (!has_left) smem_fill_zero<T, BS>(N + Cf::C);
  if (!valid) smem_fill_zero<T, L>(N + Cf::X);
  if (lane == 0 && e0 == 0 && !halo) { smem_fill_zero<T, L>(S + Cf::WT); }
  __syncwarp();

  T P[L][L], Q[L][L];
  T Mx[L][L];   // Di = D^{-1} (lower triangular), later the lower triangle of Sigma_{2e,2e}; only r >= c is used
  T dxs[L];
  {
    bool k_loaded = false;
    if constexpr (TRI) {
      if (tri_in) {
#pragma unroll
        for (int r = 0; r < L; ++r)
#pragma unroll
          for (int c = 0; c < L; ++c) Mx[r][c] = T(0);
        lds_tri<T, L>(Mx, N + Cf::A);
        k_loaded = true;
      }
    }
    if (!k_loaded) {
#pragma unroll
      for (int r = 0; r < L; ++r) {
        T row[L];
        lds_row<T, L>(row, N + Cf::A + r * L);
#pragma unroll
        for (int c = 0; c < L; ++c) Mx[r][c] = row[c];   // holds K for now
      }
    }
    if constexpr (Cf::COMPACT) {
      __syncwarp();                // every lane has its D in registers: the A slots are free
      stage_Sd();                  // S~_d arrives while D^{-1}, P and Q are computed
    }
    T dinv[L];
#pragma unroll
    for (int c = 0; c < L; ++c) dinv[c] = T(1) / Mx[c][c];
#pragma unroll
    for (int c = 0; c < L; ++c) {
      // column c of the inverse, top to bottom; the recurrence for column c reads K[r][k] with k >= c only,
      // and column c is replaced after its own recurrence, so later columns still see K where they need it
      // (all loops run over the full constant range with compile-time guards: with triangular bounds the
      // compiler left them rolled for ell = 5, 6 and put the matrix into local memory)
      T col[L];
#pragma unroll
      for (int r = 0; r < L; ++r) {
        if (r == c) col[r] = dinv[c];
        if (r > c) {
          T s = T(0);
#pragma unroll
          for (int k = 0; k < L; ++k)
            if (k >= c && k < r) s = fma(Mx[r][k], col[k], s);
          col[r] = -s * dinv[r];
        }
      }
#pragma unroll
      for (int r = 0; r < L; ++r)
        if (r >= c) Mx[r][c] = col[r];
    }
#pragma unroll
    for (int c = 0; c < L; ++c) dxs[c] = T(0);
    if (do_w) {
      T xs[L];
      lds_row<T, L>(xs, N + Cf::X);
#pragma unroll
      for (int c = 0; c < L; ++c) {
        T s = T(0);
#pragma unroll
        for (int k = c; k < L; ++k) s = fma(Mx[k][c], xs[k], s);
        dxs[c] = s;
      }
    }
    // P = F Di
#pragma unroll
    for (int r = 0; r < L; ++r) {
      T f[L];
      lds_row<T, L>(f, N + Cf::B + r * L);
#pragma unroll
      for (int c = 0; c < L; ++c) {
        T sp = T(0);
#pragma unroll
        for (int k = c; k < L; ++k) sp = fma(f[k], Mx[k][c], sp);
        P[r][c] = sp;
      }
    }
    if constexpr (Cf::COMPACT) {
      __syncwarp();                // every lane has consumed its F: the B slots are free
      stage_So();                  // S~_o arrives while Q and w are computed
    }
    // Q = G Di
#pragma unroll
    for (int r = 0; r < L; ++r) {
      T g[L];
      lds_row<T, L>(g, N + Cf::C + r * L);
#pragma unroll
      for (int c = 0; c < L; ++c) {
        T sq = T(0);
#pragma unroll
        for (int k = c; k < L; ++k) sq = fma(g[k], Mx[k][c], sq);
        Q[r][c] = sq;
      }
    }
    // Di -> lower triangle of Di^T Di, in place: entry (r, c) needs Di[k][r], Di[k][c] for k >= r only, and
    // within row r the entries are replaced left to right with (r, r) last
#pragma unroll
    for (int r = 0; r < L; ++r)
#pragma unroll
      for (int c = 0; c <= r; ++c) {
        T s = T(0);
#pragma unroll
        for (int k = r; k < L; ++k) s = fma(Mx[k][r], Mx[k][c], s);
        Mx[r][c] = s;
      }
  }

  // w_{2e} = Di^T x_e - P^T w~_e - Q^T w~_{e-1}; kept in registers and parked in X (x_e is consumed)
  T wv[L];
#pragma unroll
  for (int c = 0; c < L; ++c) wv[c] = T(0);
  if (do_w) {
    T we[L], wl[L];
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
  if constexpr (TRI) {
    if (tri_in && do_sigma) {      // packed S~_d -> full symmetric rows, in place (the products below read rows of this and of the left record)
      // inner levels: Sigma_{2e+1,2e+1} = S~_d[e] leaves right here, still packed, from the registers of the expansion
      tri_expand_inplace<T, L>(N + Cf::SD, (tri_out && valid && e < o) ? gSd + (size_t)(2 * e + 1) * PKS : nullptr);
      tri_expand_inplace_warp<T, L>(S + Cf::SD, lane);
      __syncwarp();
    }
  }
  if (do_sigma) {
    if (!has_so) smem_fill_zero<T, BS>(N + Cf::SO);
    if (lane == 0 && e0 == 0 && !halo) smem_fill_zero<T, BS>(S + Cf::SD);
    __syncwarp();
    // The two big products run as ROLLED loops over a row / column index that only addresses shared
    // memory (P and Q keep static register indices): 8x less code than full unrolling, which matters
    // because the single-warp CTAs of an SM sit at different program counters and share its instruction
    // caches (fully unrolled, this kernel stalled 1.3 cycles per instruction on instruction fetch).
    // Sigma_{2e,2e-1} = -(Q^T S~_d[e-1]^T + P^T S~_o[e-1]), COLUMN by column into C: column c needs row c of
    // S~_d[e-1] (left record) and column c of S~_o[e-1].  (This product goes first: the next one overwrites S~_o.)
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
    // Sigma_{2e+1,2e} = -(S~_d[e] P + S~_o[e-1] Q), row by row in place over S~_o
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
    if (early) {
      __syncwarp();
      out_So();                    // Sigma_off rows and w are final: let them leave while Sigma_{2e,2e} is computed
      out_w();
    }
    // Sigma_{2e,2e} = Di^T Di - Sigma_{2e+1,2e}^T P - Sigma_{2e,2e-1} Q: lower triangle only (it is symmetric),
    // accumulated in the registers that hold Di^T Di; operands are rows of B and C.  Unrolled: the
    // accumulators need static indices.  Rows then go from registers straight to global memory
    // (gR_{2e} = gd Sigma - gm w w^T at the top level).
    // At the top level the rows of the two off-diagonal results are turned into gradients in the same pass, right
    // after their last use (gO_{2e} = 2 (gd Sigma_{2e+1,2e} - gm w~_e w_{2e}^T), gO_{2e-1} = 2 (gd Sigma_{2e,2e-1} - gm w_{2e} w~_{e-1}^T)).
    if (valid) {
      T we[L], wl[L];
#pragma unroll
      for (int c = 0; c < L; ++c) { we[c] = T(0); wl[c] = T(0); }
      if (grad && do_w) {
        if (has_odd) lds_row<T, L>(we, N + Cf::WT);
        if (has_left) lds_row<T, L>(wl, Lf + Cf::WT);
      }
#pragma unroll
      for (int k = 0; k < L; ++k) {
        T bk[L];
        lds_row<T, L>(bk, N + Cf::B + k * L);
#pragma unroll
        for (int r = 0; r < L; ++r)
#pragma unroll
          for (int c = 0; c <= r; ++c) Mx[r][c] = fma(-bk[r], P[k][c], Mx[r][c]);
        if (grad && has_odd) {
#pragma unroll
          for (int c = 0; c < L; ++c) bk[c] = T(2) * (gd * bk[c] - gm * we[k] * wv[c]);
          sts_row<T, L>(N + Cf::B + k * L, bk);
        }
      }
      T* dst = gSd + (size_t)(2 * e) * BS;
#pragma unroll
      for (int r = 0; r < L; ++r) {
        T cr[L];
        lds_row<T, L>(cr, N + Cf::C + r * L);
#pragma unroll
        for (int c = 0; c <= r; ++c) {
          T s = Mx[r][c];
#pragma unroll
          for (int k = 0; k < L; ++k) s = fma(-cr[k], Q[k][c], s);
          Mx[r][c] = s;
        }
        if (grad && has_left) {
#pragma unroll
          for (int c = 0; c < L; ++c) cr[c] = T(2) * (gd * cr[c] - gm * wv[r] * wl[c]);
          sts_row<T, L>(N + Cf::C + r * L, cr);
        }
      }
#pragma unroll
      for (int r = 0; r < L; ++r) {
        T row[L];
#pragma unroll
        for (int c = 0; c < L; ++c) row[c] = (c <= r) ? Mx[r][c] : Mx[c][r];
        if (grad) {
#pragma unroll
          for (int c = 0; c < L; ++c) row[c] = gd * row[c] - gm * wv[r] * wv[c];
        }
        if constexpr (Cf::COMPACT) { if (!tri_out) stg_row<T, L>(dst + r * L, row, sd_vec); }
        else sts_row<T, L>(N + Cf::A + r * L, row);
      }
      if constexpr (TRI) {
        if (tri_out) st_tri<T, L>(gSd + (size_t)(2 * e) * PKS, Mx);     // (inner level: no gradient scaling)
      }
    }
  }
  __syncwarp();   // neighbours are done reading this record's S~_d / WT

  if (grad) {
    T we[L];
#pragma unroll
    for (int c = 0; c < L; ++c) we[c] = T(0);
    if (do_w && has_odd) lds_row<T, L>(we, N + Cf::WT);
    if (do_sigma && has_odd) {     // gR_{2e+1} = gd S~_d[e] - gm w~_e w~_e^T   (gO rows were finished above)
#pragma unroll
      for (int r = 0; r < L; ++r) {
        T v[L];
        lds_row<T, L>(v, N + Cf::SD + r * L);
#pragma unroll
        for (int c = 0; c < L; ++c) v[c] = gd * v[c] - gm * we[r] * we[c];
        sts_row<T, L>(N + Cf::SD + r * L, v);
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

  // ---------------- stage out (what has not left yet) ----------------
  out_Sd();
  if (!early || !do_sigma) {
    out_So();
    out_w();
  }
}

template <typename T, int L>
__global__ void __launch_bounds__(32 * TpnBwdCfg<T, L>::NW, TpnBwdCfg<T, L>::MIN_CTAS)
cr_tpn_bwd_kernel(const LevelBwdArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5;
  const int E = (a.m + 1) >> 1;
  const int tiles = (E + TpnBwdCfg<T, L>::NT - 1) / TpnBwdCfg<T, L>::NT;
  const long long vb = (long long)blockIdx.x * TpnBwdCfg<T, L>::NW + warp;   // virtual block = one warp's tile
  if (vb >= (long long)tiles * a.batch) return;
  const int b = (int)(vb / tiles);
  tpn_bwd_tile<T, L>(a, smem_raw + (size_t)warp * TpnBwdCfg<T, L>::SMEM_W, b, (int)(vb - (long long)b * tiles));
}

// Deep levels fused (deepest first): one CTA per series, see cr_tpn_fwd_multi_kernel.
template <typename T, int L, int NWM>
__global__ void __launch_bounds__(32 * NWM, 1)
cr_tpn_bwd_multi_kernel(const __grid_constant__ MultiArgs<LevelBwdArgs> ma) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5;
  const int b = blockIdx.x;
  for (int k = 0; k < ma.count; ++k) {
    const LevelBwdArgs& a = ma.lv[k];
    const int E = (a.m + 1) >> 1;
    const int tiles = (E + TpnBwdCfg<T, L>::NT - 1) / TpnBwdCfg<T, L>::NT;
    for (int tile = warp; tile < tiles; tile += NWM) {
      tpn_bwd_tile<T, L>(a, smem_raw + (size_t)warp * TpnBwdCfg<T, L>::SMEM_W, b, tile);
      __syncwarp();
    }
    __syncthreads();
  }
}

template <typename T, int L>
cudaError_t launch_tpn_bwd(const LevelBwdArgs& a, cudaStream_t stream) {
  using C = TpnBwdCfg<T, L>;
  static std::atomic<unsigned char> attr_done[kMaxDevices];
  if (cudaError_t e = ensure_dynamic_smem(cr_tpn_bwd_kernel<T, L>, (int)C::SMEM, attr_done); e != cudaSuccess) return e;
  const int E = (a.m + 1) / 2;
  const long long tiles = (E + C::NT - 1) / C::NT;
  const long long total = tiles * a.batch;
  if (total <= 0) return cudaSuccess;
  const long long grid = (total + C::NW - 1) / C::NW;
  if (grid > 0x7fffffffLL) return cudaErrorInvalidValue;
  cr_tpn_bwd_kernel<T, L><<<(unsigned)grid, 32 * C::NW, C::SMEM, stream>>>(a);
  return cudaGetLastError();
}

template <typename T, int L, int NWM>
cudaError_t launch_tpn_bwd_multi_w(const MultiArgs<LevelBwdArgs>& ma, cudaStream_t stream) {
  using C = TpnBwdCfg<T, L>;
  constexpr int SMEM = (int)(C::SMEM_W * NWM);
  static std::atomic<unsigned char> attr_done[kMaxDevices];
  if (cudaError_t e = ensure_dynamic_smem(cr_tpn_bwd_multi_kernel<T, L, NWM>, SMEM, attr_done); e != cudaSuccess) return e;
  if (ma.count <= 0 || ma.lv[0].batch <= 0) return cudaSuccess;
  cr_tpn_bwd_multi_kernel<T, L, NWM><<<(unsigned)ma.lv[0].batch, 32 * NWM, SMEM, stream>>>(ma);
  return cudaGetLastError();
}
template <typename T, int L>
cudaError_t launch_tpn_bwd_multi(const MultiArgs<LevelBwdArgs>& ma, cudaStream_t stream) {
  return ma.warps == 1 ? launch_tpn_bwd_multi_w<T, L, 1>(ma, stream) : launch_tpn_bwd_multi_w<T, L, kMultiWarps>(ma, stream);
}

}  // namespace crb200
