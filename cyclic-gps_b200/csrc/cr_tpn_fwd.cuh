// Thread-per-node forward level kernel for small blocks (sizeof(T) * ell^2 <= CRB200_TPN_MAX_BLOCK_BYTES = 400 bytes).
//
// Same contract as cr_level_fwd_kernel (see cr_level_fwd.cuh for the maths and the reference
// lines it replaces: cyclic_gps/cyclic_reduction.py:204-259, :412-427), different mapping:
// ONE THREAD owns one even node and keeps K, F, G and the Schur products entirely in
// registers, so there are no broadcast shared-memory reads and no shuffles inside the dense
// algebra.  A CTA is a single warp: 31 owned even nodes + 1 read-only halo node whose
// G G^T / G x contributions reach the neighbouring lane by __shfl_down.  Several such CTAs are
// resident per SM (the kernel is limited by shared memory, ~1.3 KB per node at ell=8 fp32) and
// overlap each other's load / compute / store phases.
//
// Shared memory holds one padded record per node,
//     [ R | O_left | O_right | y_even | y_odd ]   (stride NS, NS/16B odd, 848 B at ell = 8 fp32)
// filled by cp.async from coalesced global ranges; results are written in place (R_even->K, O_left->G,
// O_right->F, y_even->x, y_odd->y~) and leave as flat coalesced ranges.  The R slot is used TWICE: it
// first holds R_even (-> K, copied out as soon as it exists), then R_odd is staged into it while F and G
// are being computed, and R~ is finished there.  O~ rows go from registers straight to global memory.
// Three blocks per node instead of five means 8 resident CTAs per SM instead of 5; measured throughput of
// this kernel scales almost linearly with resident CTAs (profiles/r1_summary.md).
// The odd record stride makes the per-thread 16-byte accesses conflict free.
// Packed lower triangles (crb200_fwd_args.tri, TriPack in cr_tpn_common.cuh; float32 ell = 8): R of the inner levels arrives as 36 elements at
// the start of the R slot, K and R~ are written there packed and leave packed; everything else is unchanged.
#pragma once
#include "cr_common.cuh"
#include "cr_level_fwd.cuh"
#include "cr_tpn_common.cuh"

namespace crb200 {

template <typename T, int L>
struct TpnFwdCfg {
  static constexpr bool ELIGIBLE = (sizeof(T) * L * L <= CRB200_TPN_MAX_BLOCK_BYTES);
  static constexpr int BS = L * L;
  static constexpr int NT = 32, OWN = 31;
  static constexpr int RE = 0, OL = BS, OR_ = 2 * BS, YE = 3 * BS, YO = 3 * BS + L;   // RE: R_even, later R_odd -> R~
  static constexpr int RAW = 3 * BS + 2 * L;
  static constexpr int NS = record_stride<T>(RAW, BS);
  static constexpr size_t SMEM_W = (size_t)NT * NS * sizeof(T);                 // per warp
  // independent warps per CTA (each warp = one tile): warps of a CTA run the same code at about the same time
  static constexpr int NW = cmax(1, cmin(CRB200_TPN_WARPS, (int)((220 * 1024) / (SMEM_W * 1 + 1024))));
  static constexpr size_t SMEM = SMEM_W * NW + CRB200_SMEM_PAD;   // CRB200_SMEM_PAD: occupancy experiments only
  static constexpr int MIN_CTAS = cmin(16, cmax(1, (int)((224 * 1024) / (SMEM + 1024))));
};

// One tile (OWN even nodes of series b starting at node tile * OWN) by one warp; `smem_warp` is this warp's
// SMEM_W bytes.  Called by the level kernel (one tile per warp) and by the fused deep-level kernel.
template <typename T, int L>
__device__ __forceinline__ void tpn_fwd_tile(const LevelFwdArgs& a, unsigned char* smem_warp, const int b, const int tile) {
  using C = TpnFwdCfg<T, L>;
  constexpr int BS = C::BS, NS = C::NS, NT = C::NT, OWN = C::OWN;
  T* S = reinterpret_cast<T*>(smem_warp);
  constexpr unsigned ES = sizeof(T);
  const unsigned s0 = smem_u32(S);
  const unsigned nsb = NS * ES;

  const int m = a.m;
  const int E = (m + 1) >> 1, o = m >> 1, gcnt = (m - 1) >> 1;
  const int e0 = tile * OWN;
  const bool has_y = a.y != nullptr;
  const bool halo = a.O_halo != nullptr;
  const int lane = threadIdx.x & 31;
  // packed lower triangles (crb200_fwd_args.tri, cr_tpn_common.cuh): R comes in packed (levels >= 1 of a fused sweep) / D and R~ leave packed
  using TP = TriPack<T, L>;
  constexpr int PKS = TP::OK ? TP::PKS : BS;
  const bool tri_in = TP::OK && (a.tri & 1) != 0;
  const bool tri_out = TP::OK && (a.tri & 2) != 0;

  const T* gR = static_cast<const T*>(a.R) + (size_t)b * a.strideR;
  const T* gO = static_cast<const T*>(a.O) + (size_t)b * a.strideO;
  const T* gy = has_y ? static_cast<const T*>(a.y) + (size_t)b * a.stridey : nullptr;

  // ---------------- stage in: two cp.async groups ----------------
  // group 0 = R tile + y (needed first: Cholesky, half solve); group 1 = O tile.  The Cholesky of this CTA
  // overlaps the arrival of group 1, and the factor blocks leave for global memory as soon as they exist,
  // so one CTA keeps the memory pipe busy through most of its life (8 such CTAs fit on an SM at ell = 8 fp32).
  {
    const int r0 = 2 * e0;
    const int nR = cmin(2 * NT - 1, m - r0);
    if (tri_in) rec_g2s_strided<T, PKS, 2, true>(s0 + C::RE * ES, nsb, gR + (size_t)r0 * PKS, 0, (nR + 1) >> 1, is_aligned16(gR));
    else rec_g2s_strided<T, BS, 2>(s0 + C::RE * ES, nsb, gR + (size_t)r0 * BS, 0, (nR + 1) >> 1, is_aligned16(gR));   // even rows
    if (has_y) rec_g2s<T, L, 2>(s0 + C::YE * ES, nsb, gy + (size_t)r0 * L, 0, nR, is_aligned16(gy));
    cp_async_commit();
    const int pfirst = (r0 == 0) ? 1 : 0;
    const int nO = cmin(2 * NT - 1, m - r0) - pfirst;
    rec_g2s<T, BS, 2>(s0 + C::OL * ES, nsb, gO + (size_t)(r0 - 1 + pfirst) * BS, pfirst, nO, is_aligned16(gO));
    if (r0 == 0 && halo)
      rec_g2s<T, BS, 2>(s0 + C::OL * ES, nsb, static_cast<const T*>(a.O_halo) + (size_t)b * BS, 0, 1, is_aligned16(a.O_halo));
    cp_async_commit();
  }

  // ---------------- stage out, piece by piece ----------------
  const int n_own = cmin(OWN, E - e0);
  const int n_odd = cmax(0, cmin(OWN, o - e0));
  auto out_D_x = [&]() {
    if (a.D != nullptr) {
      if (tri_out) rec_s2g<T, PKS, 1, true>(static_cast<T*>(a.D) + ((size_t)b * E + e0) * PKS, s0 + C::RE * ES, nsb, 0, n_own, is_aligned16(a.D));
      else rec_s2g<T, BS, 1>(static_cast<T*>(a.D) + ((size_t)b * E + e0) * BS, s0 + C::RE * ES, nsb, 0, n_own, is_aligned16(a.D));
    }
    if (a.xk != nullptr && has_y)
      rec_s2g<T, L, 1>(static_cast<T*>(a.xk) + ((size_t)b * E + e0) * L, s0 + C::YE * ES, nsb, 0, n_own, is_aligned16(a.xk));
  };
  auto out_F = [&]() {
    if (a.D != nullptr)
      rec_s2g<T, BS, 1>(static_cast<T*>(a.F) + ((size_t)b * o + e0) * BS, s0 + C::OR_ * ES, nsb, 0, n_odd, is_aligned16(a.F));
  };
  auto out_G = [&]() {
    if (a.D != nullptr) {
      const int gfirst = (e0 == 0) ? 1 : 0;
      rec_s2g<T, BS, 1>(static_cast<T*>(a.G) + ((size_t)b * gcnt + (e0 + gfirst - 1)) * BS, s0 + C::OL * ES, nsb, gfirst, n_own - gfirst,
                        is_aligned16(a.G));
    }
    if (halo && e0 == 0 && a.G_halo != nullptr)
      rec_s2g<T, BS, 1>(static_cast<T*>(a.G_halo) + (size_t)b * BS, s0 + C::OL * ES, nsb, 0, 1, is_aligned16(a.G_halo));
  };
  auto out_reduced = [&]() {
    if (a.Rn != nullptr && n_odd > 0) {
      if (tri_out) rec_s2g<T, PKS, 1, true>(static_cast<T*>(a.Rn) + ((size_t)b * o + e0) * PKS, s0 + C::RE * ES, nsb, 0, n_odd, is_aligned16(a.Rn));
      else rec_s2g<T, BS, 1>(static_cast<T*>(a.Rn) + ((size_t)b * o + e0) * BS, s0 + C::RE * ES, nsb, 0, n_odd, is_aligned16(a.Rn));
      if (has_y && a.yn != nullptr)
        rec_s2g<T, L, 1>(static_cast<T*>(a.yn) + ((size_t)b * o + e0) * L, s0 + C::YO * ES, nsb, 0, n_odd, is_aligned16(a.yn));
    }
  };
  // R_odd of this tile -> the R slot (second use), issued once K has been copied out
  auto stage_R_odd = [&]() {
    const int r0 = 2 * e0;
    const int nR = cmin(2 * NT - 1, m - r0);
    if (tri_in) rec_g2s_strided<T, PKS, 2, true>(s0 + C::RE * ES, nsb, gR + (size_t)(r0 + 1) * PKS, 0, nR >> 1, is_aligned16(gR));
    else rec_g2s_strided<T, BS, 2>(s0 + C::RE * ES, nsb, gR + (size_t)(r0 + 1) * BS, 0, nR >> 1, is_aligned16(gR));
    cp_async_commit();
  };

  // variant CRB200_COPY_ONLY (profiling aid): stage in, stage out, no arithmetic -> the memory-system
  // ceiling of this access pattern
  if (a.variant == CRB200_COPY_ONLY) {
    cp_async_wait_group<0>();
    __syncwarp();
    out_D_x(); out_F(); out_G();
    __syncwarp();
    stage_R_odd();
    cp_async_wait_group<0>();
    __syncwarp();
    out_reduced();
    if (a.On != nullptr && n_odd > 1 && lane < n_odd - 1)      // same bytes as the real kernel's direct O~ stores
      for (int r = 0; r < L; ++r) {
        T z[L];
        lds_row<T, L>(z, S + (size_t)lane * NS + C::OL + r * L);
        stg_row<T, L>(static_cast<T*>(a.On) + ((size_t)b * (o - 1) + e0 + lane) * BS + r * L, z, is_aligned16(a.On));
      }
    return;
  }
  {
  // ---------------- per-node compute (everything in registers) ----------------
  cp_async_wait_group<1>();      // R tile and y have landed
  __syncwarp();
  T* N = S + (size_t)lane * NS;
  const int e = e0 + lane;
  const bool valid = e < E;
  const bool own = valid && (lane < OWN);
  const bool do_f = own && (e < o);
  const bool has_left = valid && (e >= 1 || halo);

  // neutral operands for lanes without a node (R = I, y = 0), written once into the lane's own record so that
  // the dense algebra below needs no per-element selects
  if (!valid) {
    if constexpr (TP::OK) { if (tri_in) smem_fill_identity_tri<T, L>(N + C::RE); else smem_fill_identity<T, L>(N + C::RE); }
    else smem_fill_identity<T, L>(N + C::RE);
    smem_fill_zero<T, L>(N + C::YE);
  }
  T K[L][L];
  T inv[L];
  bool bad = false;
  bool k_loaded = false;
  if constexpr (TP::OK) {
    if (tri_in) {                // only the lower triangle is ever read
#pragma unroll
      for (int r = 0; r < L; ++r)
#pragma unroll
        for (int c = 0; c < L; ++c) K[r][c] = T(0);
      lds_tri<T, L>(K, N + C::RE);
      k_loaded = true;
    }
  }
  if (!k_loaded) {
#pragma unroll
    for (int r = 0; r < L; ++r) {
      T row[L];
      lds_row<T, L>(row, N + C::RE + r * L);
#pragma unroll
      for (int c = 0; c < L; ++c) K[r][c] = row[c];
    }
  }
  double dprod = 1.0;
#pragma unroll
  for (int k = 0; k < L; ++k) {
    const T d = K[k][k];
    if (!(d > T(0))) bad = true;
    const T lkk = sqrt(d);
    inv[k] = T(1) / lkk;
    K[k][k] = lkk;
    dprod *= (double)lkk;
#pragma unroll
    for (int r = k + 1; r < L; ++r) K[r][k] *= inv[k];
#pragma unroll
    for (int c = k + 1; c < L; ++c)
#pragma unroll
      for (int r = c; r < L; ++r) K[r][c] = fma(-K[r][k], K[c][k], K[r][c]);
  }
  if (valid) {
    bool k_stored = false;
    if constexpr (TP::OK) {
      if (tri_out) { st_tri<T, L>(N + C::RE, K); k_stored = true; }
    }
    if (!k_stored) {
#pragma unroll
      for (int r = 0; r < L; ++r) {
        T row[L];
#pragma unroll
        for (int c = 0; c < L; ++c) row[c] = (c <= r) ? K[r][c] : T(0);
        sts_row<T, L>(N + C::RE + r * L, row);
      }
    }
  }
  if (bad && own && a.info != nullptr) {
    const long long flat = (long long)b * E + e;
    atomicMax(a.info, 0x7fffffff - (int)(flat > 0x7ffffffeLL ? 0x7ffffffeLL : flat));
  }
  double ld_part = (own && a.logdet != nullptr) ? log(dprod) : 0.0;

  // x = K^{-1} y_even
  T x[L];
  double mh_part = 0.0;
#pragma unroll
  for (int c = 0; c < L; ++c) x[c] = T(0);
  if (has_y) {
    lds_row<T, L>(x, N + C::YE);
#pragma unroll
    for (int c = 0; c < L; ++c) {
      T s = x[c];
#pragma unroll
      for (int k = 0; k < c; ++k) s = fma(-x[k], K[c][k], s);
      x[c] = s * inv[c];
    }
    if (valid) sts_row<T, L>(N + C::YE, x);
    if (own) {
#pragma unroll
      for (int c = 0; c < L; ++c) mh_part += (double)x[c] * (double)x[c];
    }
  }
  __syncwarp();
  out_D_x();                     // K and x leave now; the stores overlap the rest of the kernel
  __syncwarp();                  // the R slot has been read out by every lane
  cp_async_wait_group<0>();      // O tile has landed
  stage_R_odd();                 // R_odd -> R slot, arrives while F and G are computed
  __syncwarp();
  // no odd neighbour (or halo lane) -> O_right = 0 ; no left link -> O_left = 0   (boundary lanes only)
  if (!do_f) smem_fill_zero<T, BS>(N + C::OR_);
  if (!has_left) smem_fill_zero<T, BS>(N + C::OL);

  // F = O_right K^{-T} (row by row), A = F F^T (lower), u = F x
  T A[L][L];   // only r >= c used
  T u[L];
  {
    T F[L][L];
#pragma unroll
    for (int r = 0; r < L; ++r) {
      T f[L];
      lds_row<T, L>(f, N + C::OR_ + r * L);
#pragma unroll
      for (int c = 0; c < L; ++c) {
        T s = f[c];
#pragma unroll
        for (int k = 0; k < c; ++k) s = fma(-f[k], K[c][k], s);
        f[c] = s * inv[c];
      }
      if (do_f) sts_row<T, L>(N + C::OR_ + r * L, f);
      T ur = T(0);
#pragma unroll
      for (int c = 0; c < L; ++c) { F[r][c] = f[c]; ur = fma(f[c], x[c], ur); }
      u[r] = ur;
    }
#pragma unroll
    for (int r = 0; r < L; ++r)
#pragma unroll
      for (int c = 0; c <= r; ++c) {
        T s = T(0);
#pragma unroll
        for (int k = 0; k < L; ++k) s = fma(F[r][k], F[c][k], s);
        A[r][c] = s;
      }
  }

  __syncwarp();
  out_F();                       // F is final in shared memory

  // G = O_left^T K^{-T}: row r of G solves against column r of O_left
  T G[L][L];
  {
#pragma unroll
    for (int c = 0; c < L; ++c) {
      T row[L];
      lds_row<T, L>(row, N + C::OL + c * L);
#pragma unroll
      for (int r = 0; r < L; ++r) G[r][c] = row[r];
    }
#pragma unroll
    for (int r = 0; r < L; ++r) {
#pragma unroll
      for (int c = 0; c < L; ++c) {
        T s = G[r][c];
#pragma unroll
        for (int k = 0; k < c; ++k) s = fma(-G[r][k], K[c][k], s);
        G[r][c] = s * inv[c];
      }
    }
    if (has_left) {
#pragma unroll
      for (int r = 0; r < L; ++r) sts_row<T, L>(N + C::OL + r * L, G[r]);
    }
  }
  __syncwarp();
  out_G();                       // G is final in shared memory

  // O~_{e-1} = -F G^T (F rows re-read from shared memory), B = G G^T (lower), v = G x
  if (do_f && has_left) {
#pragma unroll
    for (int r = 0; r < L; ++r) {
      T f[L], on[L];
      lds_row<T, L>(f, N + C::OR_ + r * L);
#pragma unroll
      for (int c = 0; c < L; ++c) {
        T s = T(0);
#pragma unroll
        for (int k = 0; k < L; ++k) s = fma(-f[k], G[c][k], s);
        on[c] = s;
      }
      // O~_{e-1}[r,:] straight to global memory (one 32-byte sector per row at ell = 8 fp32)
      T* base = static_cast<T*>(e >= 1 ? a.On : a.On_halo);
      if (base != nullptr) {
        T* dst = base + ((e >= 1) ? ((size_t)b * (o - 1) + (e - 1)) * BS : (size_t)b * BS);
        stg_row<T, L>(dst + r * L, on, is_aligned16(base));
      }
    }
  }
  T Bn[L][L];   // lower: G G^T of the NEXT even node after the shuffle
  T vn[L];
#pragma unroll
  for (int r = 0; r < L; ++r) {
#pragma unroll
    for (int c = 0; c <= r; ++c) {
      T s = T(0);
#pragma unroll
      for (int k = 0; k < L; ++k) s = fma(G[r][k], G[c][k], s);
      Bn[r][c] = s;
    }
    T s = T(0);
#pragma unroll
    for (int k = 0; k < L; ++k) s = fma(G[r][k], x[k], s);
    vn[r] = s;
  }
  if (halo && e0 == 0 && lane == 0) {
    // link to the virtual node -1: accumulate -G G^T and -G x there
    if (a.Rh_acc != nullptr) {
      T* acc = static_cast<T*>(a.Rh_acc) + (size_t)b * BS;
#pragma unroll
      for (int r = 0; r < L; ++r)
#pragma unroll
        for (int c = 0; c < L; ++c) acc[r * L + c] -= (r >= c) ? Bn[r][c] : Bn[c][r];
    }
    if (a.yh_acc != nullptr && has_y) {
      T* acc = static_cast<T*>(a.yh_acc) + (size_t)b * L;
#pragma unroll
      for (int r = 0; r < L; ++r) acc[r] -= vn[r];
    }
  }
#pragma unroll
  for (int r = 0; r < L; ++r) {
#pragma unroll
    for (int c = 0; c <= r; ++c) Bn[r][c] = __shfl_down_sync(0xffffffffu, Bn[r][c], 1);
    vn[r] = __shfl_down_sync(0xffffffffu, vn[r], 1);
  }

  // R~_e = R_odd - A - B_{e+1},  y~_e = y_odd - u - v_{e+1}
  cp_async_wait_group<0>();      // R_odd has landed in the R slot
  __syncwarp();
  if (do_f) {
    // (no even node e+1: the lane to the right holds G = 0, so B and v arrive as zeros)
    bool r_done = false;
    if constexpr (TP::OK) {
      if (tri_out) {               // lower triangle only: R~ leaves packed (the next level's Cholesky reads nothing else)
        T Rm[L][L];
        if (tri_in) {
          lds_tri<T, L>(Rm, N + C::RE);
        } else {
#pragma unroll
          for (int r = 0; r < L; ++r) {
            T row[L];
            lds_row<T, L>(row, N + C::RE + r * L);
#pragma unroll
            for (int c = 0; c <= r; ++c) Rm[r][c] = row[c];
          }
        }
#pragma unroll
        for (int r = 0; r < L; ++r)
#pragma unroll
          for (int c = 0; c <= r; ++c) Rm[r][c] = Rm[r][c] - A[r][c] - Bn[r][c];
        st_tri<T, L>(N + C::RE, Rm);
        r_done = true;
      }
    }
    if (!r_done) {
#pragma unroll
      for (int r = 0; r < L; ++r) {
        T row[L];
        lds_row<T, L>(row, N + C::RE + r * L);
#pragma unroll
        for (int c = 0; c < L; ++c) {
          const T av = (r >= c) ? A[r][c] : A[c][r];
          const T bv = (r >= c) ? Bn[r][c] : Bn[c][r];
          row[c] = row[c] - av - bv;
        }
        sts_row<T, L>(N + C::RE + r * L, row);
      }
    }
    if (has_y) {
      T yo[L];
      lds_row<T, L>(yo, N + C::YO);
#pragma unroll
      for (int r = 0; r < L; ++r) yo[r] = yo[r] - u[r] - vn[r];
      sts_row<T, L>(N + C::YO, yo);
    }
  }

  // scalars: warp reduce, one atomic per CTA
  if (a.logdet != nullptr) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) ld_part += __shfl_xor_sync(0xffffffffu, ld_part, off);
    if (lane == 0) atomicAdd(a.logdet + acc_index(a, b, tile), ld_part);
  }
  if (a.mahal != nullptr && has_y) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) mh_part += __shfl_xor_sync(0xffffffffu, mh_part, off);
    if (lane == 0) atomicAdd(a.mahal + acc_index(a, b, tile), mh_part);
  }
  __syncwarp();

  }

  out_reduced();                 // R~, y~, O~ (and the halo coupling)
}

template <typename T, int L>
__global__ void __launch_bounds__(32 * TpnFwdCfg<T, L>::NW, TpnFwdCfg<T, L>::MIN_CTAS)
cr_tpn_fwd_kernel(const LevelFwdArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5;
  const int E = (a.m + 1) >> 1;
  const int tiles = (E + TpnFwdCfg<T, L>::OWN - 1) / TpnFwdCfg<T, L>::OWN;
  const long long vb = (long long)blockIdx.x * TpnFwdCfg<T, L>::NW + warp;   // virtual block = one warp's tile
  if (vb >= (long long)tiles * a.batch) return;
  const int b = (int)(vb / tiles);
  tpn_fwd_tile<T, L>(a, smem_raw + (size_t)warp * TpnFwdCfg<T, L>::SMEM_W, b, (int)(vb - (long long)b * tiles));
}

// Deep levels fused: ONE CTA per series runs levels lv[0..count) back to back, its warps sharing the tiles of a
// level; levels communicate through global memory (L2) and a __syncthreads.  Used by crb200_sweep_fwd for the
// levels with only a few tiles per series, where a launch per level costs more than its work.
template <typename T, int L, int NWM>
__global__ void __launch_bounds__(32 * NWM, 1)
cr_tpn_fwd_multi_kernel(const __grid_constant__ MultiArgs<LevelFwdArgs> ma) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5;
  const int b = blockIdx.x;
  for (int k = 0; k < ma.count; ++k) {
    const LevelFwdArgs& a = ma.lv[k];
    const int E = (a.m + 1) >> 1;
    const int tiles = (E + TpnFwdCfg<T, L>::OWN - 1) / TpnFwdCfg<T, L>::OWN;
    for (int tile = warp; tile < tiles; tile += NWM) {
      tpn_fwd_tile<T, L>(a, smem_raw + (size_t)warp * TpnFwdCfg<T, L>::SMEM_W, b, tile);
      __syncwarp();
    }
    __syncthreads();
  }
}

template <typename T, int L>
cudaError_t launch_tpn_fwd(const LevelFwdArgs& a, cudaStream_t stream) {
  using C = TpnFwdCfg<T, L>;
  static std::atomic<unsigned char> attr_done[kMaxDevices];
  if (cudaError_t e = ensure_dynamic_smem(cr_tpn_fwd_kernel<T, L>, (int)C::SMEM, attr_done); e != cudaSuccess) return e;
  const int E = (a.m + 1) / 2;
  const long long tiles = (E + C::OWN - 1) / C::OWN;
  const long long total = tiles * a.batch;
  if (total <= 0) return cudaSuccess;
  const long long grid = (total + C::NW - 1) / C::NW;
  if (grid > 0x7fffffffLL) return cudaErrorInvalidValue;
  cr_tpn_fwd_kernel<T, L><<<(unsigned)grid, 32 * C::NW, C::SMEM, stream>>>(a);
  return cudaGetLastError();
}

template <typename T, int L, int NWM>
cudaError_t launch_tpn_fwd_multi_w(const MultiArgs<LevelFwdArgs>& ma, cudaStream_t stream) {
  using C = TpnFwdCfg<T, L>;
  constexpr int SMEM = (int)(C::SMEM_W * NWM);
  static std::atomic<unsigned char> attr_done[kMaxDevices];
  if (cudaError_t e = ensure_dynamic_smem(cr_tpn_fwd_multi_kernel<T, L, NWM>, SMEM, attr_done); e != cudaSuccess) return e;
  if (ma.count <= 0 || ma.lv[0].batch <= 0) return cudaSuccess;
  cr_tpn_fwd_multi_kernel<T, L, NWM><<<(unsigned)ma.lv[0].batch, 32 * NWM, SMEM, stream>>>(ma);
  return cudaGetLastError();
}
template <typename T, int L>
cudaError_t launch_tpn_fwd_multi(const MultiArgs<LevelFwdArgs>& ma, cudaStream_t stream) {
  return ma.warps == 1 ? launch_tpn_fwd_multi_w<T, L, 1>(ma, stream) : launch_tpn_fwd_multi_w<T, L, kMultiWarps>(ma, stream);
}

}  // namespace crb200
