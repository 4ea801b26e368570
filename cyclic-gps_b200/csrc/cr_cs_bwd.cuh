// Column-split backward level kernel: LPN lanes per even node, each lane owns CW = L / LPN columns
// of the register-resident operands P = F D^{-1} and Q = G D^{-1}.
//
// Same contract and record layout as cr_tpn_bwd_kernel (cr_tpn_bwd.cuh; reference
// cyclic_gps/cyclic_reduction.py:362-373, :478-501).  Why: the thread-per-node kernel keeps ~1.3 KB of
// shared memory per node in flight, so only five warps fit on an SM and every stall is exposed.  Splitting
// a node over LPN lanes keeps the shared-memory footprint per node but multiplies the resident warps by
// LPN and divides the per-lane registers and FMA chain by LPN; every output is separable by columns:
//   P[:,c], Q[:,c]                        from D^{-1}[:,c]
//   Sigma_{2e+1,2e}[:,c]                  = -(S~_d[e] P[:,c] + S~_o[e-1] Q[:,c])
//   Sigma_{2e,2e-1}[r,:]  (ROWS r = my c) = -(Q[:,r]^T S~_d[e-1]^T + P[:,r]^T S~_o[e-1])
//   Sigma_{2e,2e}[:,c]                    = (D^{-T} D^{-1})[:,c] - S_d^T P[:,c] - S_o^T-rows Q[:,c]
//   w_{2e}[c]                             = (D^{-T} x)[c] - P[:,c]^T w~_e - Q[:,c]^T w~_{e-1}
// The column offset c0 of a lane only ever enters shared-memory ADDRESSES (never a register index), so all
// lanes of a warp run the same instruction stream.  D^{-1} is computed redundantly by every lane of a node
// (36 elements at L = 8) and parked in shared memory so that the column slices can be read back.
#pragma once
#include "cr_tpn_bwd.cuh"

namespace crb200 {

// record layout of the column-split kernel (five blocks; see cr_tpn_bwd.cuh for the roles)
template <typename T, int L>
struct CsBwdRec {
  static constexpr int BS = L * L;
  static constexpr int A = 0, SD = BS, C = 2 * BS, B = 3 * BS, SO = 4 * BS, X = 5 * BS, WT = 5 * BS + L;
  static constexpr int RAW = 5 * BS + 2 * L;
  static constexpr int NS = record_stride<T>(RAW, BS);
};

template <typename T, int L, int LPN>
struct CsBwdCfg {
  static constexpr int CW = L / LPN;
  static constexpr bool ELIGIBLE = (L % LPN == 0) && (LPN > 1) && (L <= 8) && ((CW * (int)sizeof(T)) % 16 == 0);
  static constexpr int BS = L * L;
  static constexpr int NT = 32 / LPN;          // nodes per warp
  using Rec = CsBwdRec<T, L>;
  static constexpr int NS = Rec::NS;
  static constexpr size_t SMEM = (size_t)(NT + 1) * NS * sizeof(T);
  static constexpr int MIN_CTAS = cmin(CRB200_CS_MAX_CTAS, cmax(1, (int)((226 * 1024) / (SMEM + 1024))));
};

// slice of CW consecutive elements (16-byte multiple) from / to shared memory
template <typename T, int CW>
__device__ __forceinline__ void lds_slice(T (&a)[CW], const T* p) { lds_row<T, CW>(a, p); }
template <typename T, int CW>
__device__ __forceinline__ void sts_slice(T* p, const T (&a)[CW]) { sts_row<T, CW>(p, a); }

template <typename T, int L, int LPN>
__global__ void __launch_bounds__(32, CsBwdCfg<T, L, LPN>::MIN_CTAS)
cr_cs_bwd_kernel(const LevelBwdArgs a) {
  using Cf = CsBwdCfg<T, L, LPN>;
  using R = typename Cf::Rec;
  constexpr int BS = Cf::BS, NS = Cf::NS, NT = Cf::NT, CW = Cf::CW;
  constexpr unsigned ES = sizeof(T);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* S = reinterpret_cast<T*>(smem_raw);
  const unsigned s0 = smem_u32(S);
  const unsigned nsb = NS * ES;
  const unsigned rec1 = s0 + nsb;

  const int m = a.m;
  const int E = (m + 1) >> 1, o = m >> 1, gcnt = (m - 1) >> 1;
  const int tiles = (E + NT - 1) / NT;
  const int b = blockIdx.x / tiles;
  const int tile = blockIdx.x - b * tiles;
  const int e0 = tile * NT;
  const int nE = cmin(NT, E - e0);
  const bool do_sigma = a.Sd_out != nullptr;
  const bool do_w = a.w_out != nullptr;
  const bool halo = a.G_halo != nullptr;
  const int lane = threadIdx.x;

  // ---------------- stage in (identical to the thread-per-node kernel, NT records) ----------------
  {
    rec_g2s<T, BS, 1>(rec1 + R::A * ES, nsb, static_cast<const T*>(a.D) + ((size_t)b * E + e0) * BS, 0, nE, is_aligned16(a.D));
    const int nF = cmax(0, cmin(NT, o - e0));
    rec_g2s<T, BS, 1>(rec1 + R::B * ES, nsb, static_cast<const T*>(a.F) + ((size_t)b * o + e0) * BS, 0, nF, is_aligned16(a.F));
    const int gf = (e0 == 0) ? 1 : 0;
    rec_g2s<T, BS, 1>(rec1 + R::C * ES, nsb, static_cast<const T*>(a.G) + ((size_t)b * gcnt + (e0 + gf - 1)) * BS, gf, nE - gf,
                      is_aligned16(a.G));
    if (e0 == 0 && halo)
      rec_g2s<T, BS, 1>(rec1 + R::C * ES, nsb, static_cast<const T*>(a.G_halo) + (size_t)b * BS, 0, 1, is_aligned16(a.G_halo));
    const int ilo = (e0 == 0) ? 1 : 0;
    const int nodd = cmin(e0 + NT, o) - (e0 - 1 + ilo);
    if (do_sigma) {
      rec_g2s<T, BS, 1>(s0 + R::SD * ES, nsb, static_cast<const T*>(a.Sd_in) + ((size_t)b * o + (e0 - 1 + ilo)) * BS, ilo, nodd,
                        is_aligned16(a.Sd_in));
      const int nso = cmin(e0 + NT - 1, o - 1) - (e0 - 1 + ilo);
      rec_g2s<T, BS, 1>(rec1 + R::SO * ES, nsb, static_cast<const T*>(a.So_in) + ((size_t)b * (o - 1) + (e0 - 1 + ilo)) * BS, ilo, nso,
                        is_aligned16(a.So_in));
      if (e0 == 0 && halo) {
        rec_g2s<T, BS, 1>(s0 + R::SD * ES, nsb, static_cast<const T*>(a.Sd_halo) + (size_t)b * BS, 0, 1, is_aligned16(a.Sd_halo));
        if (o > 0)
          rec_g2s<T, BS, 1>(rec1 + R::SO * ES, nsb, static_cast<const T*>(a.So_halo_in) + (size_t)b * BS, 0, 1, is_aligned16(a.So_halo_in));
      }
    }
    if (do_w) {
      rec_g2s<T, L, 1>(rec1 + R::X * ES, nsb, static_cast<const T*>(a.xk) + ((size_t)b * E + e0) * L, 0, nE, is_aligned16(a.xk));
      rec_g2s<T, L, 1>(s0 + R::WT * ES, nsb, static_cast<const T*>(a.w_in) + ((size_t)b * o + (e0 - 1 + ilo)) * L, ilo, nodd,
                       is_aligned16(a.w_in));
      if (e0 == 0 && halo)
        rec_g2s<T, L, 1>(s0 + R::WT * ES, nsb, static_cast<const T*>(a.w_halo) + (size_t)b * L, 0, 1, is_aligned16(a.w_halo));
    }
    cp_async_wait_all();
    __syncwarp();
  }

  // ---------------- per-node compute, LPN lanes per node ----------------
  const int t = lane / LPN;                  // node within the tile
  const int c0 = (lane - t * LPN) * CW;      // first column owned by this lane (only used in addresses)
  const bool lead = (c0 == 0);               // one lane per node does the non-separable stores
  T* N = S + (size_t)(t + 1) * NS;
  const T* Lf = S + (size_t)t * NS;
  const int e = e0 + t;
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

  T Pm[L][CW], Qm[L][CW];      // my columns of P and Q
  {
    // D^{-1}, full, redundantly on every lane of the node (registers), then parked in slot A
    T Di[L][L];
#pragma unroll
    for (int r = 0; r < L; ++r) {
      T row[L];
      lds_row<T, L>(row, N + R::A + r * L);
#pragma unroll
      for (int c = 0; c < L; ++c) Di[r][c] = valid ? row[c] : (r == c ? T(1) : T(0));
    }
    T dinv[L];
#pragma unroll
    for (int c = 0; c < L; ++c) dinv[c] = T(1) / Di[c][c];
#pragma unroll
    for (int c = 0; c < L; ++c) {
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
    __syncwarp();                                   // every lane of the node has read D
    if (lead && valid) {
#pragma unroll
      for (int r = 0; r < L; ++r) {
        T row[L];
#pragma unroll
        for (int c = 0; c < L; ++c) row[c] = (c <= r) ? Di[r][c] : T(0);
        sts_row<T, L>(N + R::A + r * L, row);
      }
    }
    __syncwarp();
    T Dm[L][CW];                                    // my columns of D^{-1}
#pragma unroll
    for (int k = 0; k < L; ++k) lds_slice<T, CW>(Dm[k], N + R::A + k * L + c0);

    // P[:, mine] = F D^{-1}[:, mine],  Q[:, mine] = G D^{-1}[:, mine]
#pragma unroll
    for (int r = 0; r < L; ++r) {
      T f[L], g[L];
      lds_row<T, L>(f, N + R::B + r * L);
      lds_row<T, L>(g, N + R::C + r * L);
#pragma unroll
      for (int cc = 0; cc < CW; ++cc) {
        T sp = T(0), sq = T(0);
#pragma unroll
        for (int k = 0; k < L; ++k) { sp = fma(f[k], Dm[k][cc], sp); sq = fma(g[k], Dm[k][cc], sq); }
        Pm[r][cc] = has_odd ? sp : T(0);
        Qm[r][cc] = has_left ? sq : T(0);
      }
      sched_fence();
    }
    // w_{2e}[mine] = (D^{-T} x)[mine] - P^T w~_e - Q^T w~_{e-1}
    T wm[CW];
#pragma unroll
    for (int cc = 0; cc < CW; ++cc) wm[cc] = T(0);
    if (do_w) {
      T xs[L], we[L], wl[L];
      lds_row<T, L>(xs, N + R::X);
      lds_row<T, L>(we, N + R::WT);
      lds_row<T, L>(wl, Lf + R::WT);
#pragma unroll
      for (int cc = 0; cc < CW; ++cc) {
        T s = T(0);
#pragma unroll
        for (int k = 0; k < L; ++k) {
          s = fma(Dm[k][cc], xs[k], s);
          s = fma(-Pm[k][cc], has_odd ? we[k] : T(0), s);
          s = fma(-Qm[k][cc], has_left ? wl[k] : T(0), s);
        }
        wm[cc] = s;
      }
    }
    // (D^{-T} D^{-1})[:, mine] -> slot A (after every lane has fetched its D^{-1} columns)
    T dtd[L][CW];
    if (do_sigma) {
#pragma unroll
      for (int r = 0; r < L; ++r)
#pragma unroll
        for (int cc = 0; cc < CW; ++cc) {
          T s = T(0);
#pragma unroll
          for (int k = r; k < L; ++k) s = fma(Di[k][r], Dm[k][cc], s);
          dtd[r][cc] = s;
        }
    }
    __syncwarp();                                   // all D^{-1} column reads and x reads are done
    if (do_sigma && valid) {
#pragma unroll
      for (int r = 0; r < L; ++r) sts_slice<T, CW>(N + R::A + r * L + c0, dtd[r]);
    }
    if (do_w && valid) sts_slice<T, CW>(N + R::X + c0, wm);
  }

  if (do_sigma) {
    // Sigma_{2e+1,2e}[:, mine] = -(S~_d[e] P + S~_o[e-1] Q)[:, mine], row by row into B
    __syncwarp();                                   // F rows (slot B) are no longer needed by any lane
    if (has_odd) {
#pragma unroll 1
      for (int r = 0; r < L; ++r) {
        T sig[L], so[L], out[CW];
        lds_row<T, L>(sig, N + R::SD + r * L);
        lds_row<T, L>(so, N + R::SO + r * L);
#pragma unroll
        for (int cc = 0; cc < CW; ++cc) {
          T s = T(0);
#pragma unroll
          for (int k = 0; k < L; ++k) { s = fma(-sig[k], Pm[k][cc], s); s = fma(has_so ? -so[k] : T(0), Qm[k][cc], s); }
          out[cc] = s;
        }
        sts_slice<T, CW>(N + R::B + r * L + c0, out);
      }
    }
    // Sigma_{2e,2e-1}[rows mine, :] = -(Q[:,r]^T S~_d[e-1]^T + P[:,r]^T S~_o[e-1]), column by column into C
    if (has_left) {
#pragma unroll 1
      for (int c = 0; c < L; ++c) {
        T a0[L], socol[L], st[CW];
        lds_row<T, L>(a0, Lf + R::SD + c * L);
#pragma unroll
        for (int k = 0; k < L; ++k) socol[k] = has_so ? N[R::SO + k * L + c] : T(0);
#pragma unroll
        for (int rr = 0; rr < CW; ++rr) {
          T s = T(0);
#pragma unroll
          for (int k = 0; k < L; ++k) { s = fma(-Qm[k][rr], a0[k], s); s = fma(-Pm[k][rr], socol[k], s); }
          st[rr] = s;
        }
        // G (slot C) was consumed before the first __syncwarp of this block; rows c0.. are mine
#pragma unroll
        for (int rr = 0; rr < CW; ++rr) N[R::C + (c0 + rr) * L + c] = st[rr];
      }
    }
    __syncwarp();                                   // S_d and Sigma_{2e,2e-1} complete (both lanes' parts)
    // Sigma_{2e,2e}[:, mine] = DtD - S_d^T P - (S_o^T) Q, row by row in place in A
    if (valid) {
      T wv[L];
#pragma unroll
      for (int c = 0; c < L; ++c) wv[c] = T(0);
      if (grad && do_w) lds_row<T, L>(wv, N + R::X);
#pragma unroll 1
      for (int r = 0; r < L; ++r) {
        T acc[CW], sdcol[L], st[L];
        lds_slice<T, CW>(acc, N + R::A + r * L + c0);
        lds_row<T, L>(st, N + R::C + r * L);
#pragma unroll
        for (int k = 0; k < L; ++k) sdcol[k] = has_odd ? N[R::B + k * L + r] : T(0);
#pragma unroll
        for (int cc = 0; cc < CW; ++cc) {
          T s = acc[cc];
#pragma unroll
          for (int k = 0; k < L; ++k) { s = fma(-sdcol[k], Pm[k][cc], s); s = fma(has_left ? -st[k] : T(0), Qm[k][cc], s); }
          acc[cc] = s;
        }
        if (grad) {
          T wmine[CW];
          lds_slice<T, CW>(wmine, N + R::X + c0);
          const T wr = do_w ? N[R::X + r] : T(0);
#pragma unroll
          for (int cc = 0; cc < CW; ++cc) acc[cc] = gd * acc[cc] - gm * wr * (do_w ? wmine[cc] : T(0));
        }
        sts_slice<T, CW>(N + R::A + r * L + c0, acc);
      }
    }
  }
  __syncwarp();   // neighbours are done reading this record's SD / WT; all of A, B, C is final (untransformed B, C)

  if (grad) {
    // gradient assembly on my column slices; w~ of the neighbours is read before anyone rescales it
    T wmine[CW], wem[CW], wlm[CW];
#pragma unroll
    for (int cc = 0; cc < CW; ++cc) { wmine[cc] = T(0); wem[cc] = T(0); wlm[cc] = T(0); }
    if (do_w) {
      if (valid) lds_slice<T, CW>(wmine, N + R::X + c0);
      if (has_odd) lds_slice<T, CW>(wem, N + R::WT + c0);
      if (has_left) lds_slice<T, CW>(wlm, Lf + R::WT + c0);
    }
    if (do_sigma) {
#pragma unroll 1
      for (int r = 0; r < L; ++r) {
        const T we_r = (do_w && has_odd) ? N[R::WT + r] : T(0);
        const T wv_r = (do_w && valid) ? N[R::X + r] : T(0);
        if (has_odd) {
          T v[CW];
          lds_slice<T, CW>(v, N + R::SD + r * L + c0);
#pragma unroll
          for (int cc = 0; cc < CW; ++cc) v[cc] = gd * v[cc] - gm * we_r * wem[cc];
          sts_slice<T, CW>(N + R::SD + r * L + c0, v);
          lds_slice<T, CW>(v, N + R::B + r * L + c0);
#pragma unroll
          for (int cc = 0; cc < CW; ++cc) v[cc] = T(2) * (gd * v[cc] - gm * we_r * wmine[cc]);
          sts_slice<T, CW>(N + R::B + r * L + c0, v);
        }
        if (has_left) {
          T v[CW];
          lds_slice<T, CW>(v, N + R::C + r * L + c0);
#pragma unroll
          for (int cc = 0; cc < CW; ++cc) v[cc] = T(2) * (gd * v[cc] - gm * wv_r * wlm[cc]);
          sts_slice<T, CW>(N + R::C + r * L + c0, v);
        }
      }
    }
    __syncwarp();
    if (do_w) {
#pragma unroll
      for (int cc = 0; cc < CW; ++cc) { wmine[cc] = T(2) * gm * wmine[cc]; wem[cc] = T(2) * gm * wem[cc]; }
      if (valid) sts_slice<T, CW>(N + R::X + c0, wmine);
      if (has_odd) sts_slice<T, CW>(N + R::WT + c0, wem);
    }
  }
  __syncwarp();

  // ---------------- stage out ----------------
  const int row_lo = 2 * e0;
  const int nrows = cmin(2 * nE, m - row_lo);
  const int so_plo = (e0 == 0) ? 1 : 0;
  const int nso_rows = cmin(2 * e0 + 2 * nE - 1, m - 1) - (2 * e0 - 1 + so_plo);
  if (do_sigma) {
    T* Sd = static_cast<T*>(a.Sd_out) + (size_t)b * a.strideSd;
    rec_s2g<T, BS, 2>(Sd + (size_t)row_lo * BS, rec1 + R::A * ES, nsb, 0, nrows, is_aligned16(Sd));
    if (nso_rows > 0) {
      T* So = static_cast<T*>(a.So_out) + (size_t)b * a.strideSo;
      rec_s2g<T, BS, 2>(So + (size_t)(2 * e0 - 1 + so_plo) * BS, rec1 + R::C * ES, nsb, so_plo, nso_rows, is_aligned16(So));
    }
    if (e0 == 0 && halo && a.So_halo_out != nullptr)
      rec_s2g<T, BS, 1>(static_cast<T*>(a.So_halo_out) + (size_t)b * BS, rec1 + R::C * ES, nsb, 0, 1, is_aligned16(a.So_halo_out));
  }
  if (do_w) {
    T* W = static_cast<T*>(a.w_out) + (size_t)b * a.stridew;
    rec_s2g<T, L, 2>(W + (size_t)row_lo * L, rec1 + R::X * ES, nsb, 0, nrows, is_aligned16(W));
  }
}

template <typename T, int L, int LPN>
cudaError_t launch_cs_bwd(const LevelBwdArgs& a, cudaStream_t stream) {
  using C = CsBwdCfg<T, L, LPN>;
  static std::atomic<unsigned char> attr_done[kMaxDevices];
  if (cudaError_t e = ensure_dynamic_smem(cr_cs_bwd_kernel<T, L, LPN>, (int)C::SMEM, attr_done); e != cudaSuccess) return e;
  const int E = (a.m + 1) / 2;
  const long long tiles = (E + C::NT - 1) / C::NT;
  const long long grid = tiles * a.batch;
  if (grid <= 0) return cudaSuccess;
  if (grid > 0x7fffffffLL) return cudaErrorInvalidValue;
  cr_cs_bwd_kernel<T, L, LPN><<<(unsigned)grid, 32, C::SMEM, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace crb200
