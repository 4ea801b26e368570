// One cyclic-reduction level, forward (factor + reduce + half-solve + log-det / Mahalanobis
// partial sums) as ONE fused kernel.
//
// Replaces, per level, the reference's decompose_step (cyclic_gps/cyclic_reduction.py:204-259)
// together with the per-level body of mahal_and_det (:412-427) / halfsolve (:318-333):
//   K_e  = chol(R_{2e})                       D output          (:225-227)
//   F_e  = O_{2e}   K_e^{-T}                  (e < o)           (:242-244)
//   G_{e-1} = O_{2e-1}^T K_e^{-T}             (e >= 1)          (:246-248)
//   x_e  = K_e^{-1} y_{2e}                                       (:420-423)
//   R~_e = R_{2e+1} - F_e F_e^T - G_e G_e^T   (G term iff e+1<E) (:250-253, UU_T :15-37)
//   O~_{e-1} = -F_e G_{e-1}^T                 (1 <= e < o)       (:254)
//   y~_e = y_{2e+1} - F_e x_e - G_e x_{e+1}                      (:427, Ux :40-60)
//   logdet += sum log diag K_e ; mahal += |x_e|^2                (:417,:424)
//
// Work mapping: one group of LG lanes per EVEN node e, lane r = row r of every block.  A CTA
// owns NG-1 consecutive even nodes of one series plus one read-only halo node (the next even
// node, recomputed so that G_e G_e^T and G_e x_{e+1} are available without a grid-wide
// exchange).  Inputs are staged with cp.async as three flat, fully coalesced ranges
// (R rows 2e0.., O rows 2e0-1.., y rows 2e0..); all results are produced IN PLACE in that
// staging area and leave as flat coalesced ranges (block stride 1 or 2).
//
// Optional left halo (chunk-partitioned long series): the series has a virtual
// odd node -1 owned by another chunk; O_halo couples it to row 0.  The kernel then also
// emits G_{-1}, the next-level coupling O~_{-1}, and accumulates the updates -G G^T / -G x
// destined for that node into Rh_acc / yh_acc.
#pragma once
#include "cr_common.cuh"

namespace crb200 {

using LevelFwdArgs = ::crb200_fwd_args;   // include/crb200.h

// slot of the per-series scalar accumulators this CTA adds into (see crb200_fwd_args.acc_slots)
__device__ __forceinline__ size_t acc_index(const LevelFwdArgs& a, int b, int tile) {
  const int slots = a.acc_slots > 0 ? a.acc_slots : 1;
  return (size_t)b * slots + (tile % slots);
}

template <typename T, int L>
struct FwdCfg {
  static constexpr int LG = GroupLanes<L>::value;
  static constexpr int BS = L * L;
  static constexpr int NODE_ELEMS = 6 * BS + 3 * L;
  static constexpr int GRAN = 32 / LG;  // groups per warp: NG must be a multiple
  // shared-memory budget per CTA: 100 KB (two CTAs per SM), or 200 KB when fewer than four nodes would fit in 100 KB
  // (fp64 ell >= 24): every CTA recomputes one halo node, so with NG = 2 half of the work would be redundant
  static constexpr int CAP = (4 * NODE_ELEMS * (int)sizeof(T) > 100 * 1024) ? 200 * 1024 : 100 * 1024;
  static constexpr int NG_FIT = CAP / (NODE_ELEMS * (int)sizeof(T));
  static constexpr int NG_RAW = cmin(kThreads / LG, NG_FIT);
  static constexpr int NG = cmax(cmax(2, GRAN), (NG_RAW / GRAN) * GRAN);
  static constexpr int THREADS = NG * LG;
  // shared memory carve-up (element offsets), every region 16-byte aligned
  static constexpr size_t R_OFF = 0;
  static constexpr size_t O_OFF = align16(R_OFF + sizeof(T) * (2 * NG - 1) * BS);
  static constexpr size_t Y_OFF = align16(O_OFF + sizeof(T) * (2 * NG - 1) * BS);
  static constexpr size_t ON_OFF = align16(Y_OFF + sizeof(T) * (2 * NG - 1) * L);
  static constexpr size_t B_OFF = align16(ON_OFF + sizeof(T) * NG * BS);
  static constexpr size_t V_OFF = align16(B_OFF + sizeof(T) * NG * BS);
  static constexpr size_t RED_OFF = align16(V_OFF + sizeof(T) * NG * L);
  static constexpr size_t SMEM = RED_OFF + 32 * sizeof(double);
};

template <typename T, int L>
__global__ void __launch_bounds__(FwdCfg<T, L>::THREADS)
cr_level_fwd_kernel(const LevelFwdArgs a) {
  using C = FwdCfg<T, L>;
  constexpr int LG = C::LG, NG = C::NG, BS = C::BS, OWN = NG - 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* sR = reinterpret_cast<T*>(smem_raw + C::R_OFF);
  T* sO = reinterpret_cast<T*>(smem_raw + C::O_OFF);
  T* sy = reinterpret_cast<T*>(smem_raw + C::Y_OFF);
  T* sOn = reinterpret_cast<T*>(smem_raw + C::ON_OFF);
  T* sB = reinterpret_cast<T*>(smem_raw + C::B_OFF);
  T* sv = reinterpret_cast<T*>(smem_raw + C::V_OFF);
  double* sred = reinterpret_cast<double*>(smem_raw + C::RED_OFF);

  const int m = a.m;
  const int E = (m + 1) >> 1, o = m >> 1, gcnt = (m - 1) >> 1;
  const int tiles = (E + OWN - 1) / OWN;
  const int b = blockIdx.x / tiles;
  const int tile = blockIdx.x - b * tiles;
  const int e0 = tile * OWN;
  const bool has_y = a.y != nullptr;
  const bool halo = a.O_halo != nullptr;

  const T* gR = static_cast<const T*>(a.R) + (size_t)b * a.strideR;
  const T* gO = static_cast<const T*>(a.O) + (size_t)b * a.strideO;
  const T* gy = has_y ? static_cast<const T*>(a.y) + (size_t)b * a.stridey : nullptr;

  // ---------------- stage in ----------------
  {
    const int r0 = 2 * e0;
    const int nR = cmin(2 * NG - 1, m - r0);                  // R rows r0 .. r0+nR
    tile_g2s<T, BS, 1>(sR, gR + (size_t)r0 * BS, nR, is_aligned16(gR));
    // O rows: position p <-> O_{r0-1+p}; p = 0 is the left link of the first even node
    const int pfirst = (r0 == 0) ? 1 : 0;
    const int nO = cmin(2 * NG - 1, (m - 1) - (r0 - 1)) - pfirst;   // rows up to O_{m-2}
    tile_g2s<T, BS, 1>(sO + (size_t)pfirst * BS, gO + (size_t)(r0 - 1 + pfirst) * BS, nO, is_aligned16(gO));
    if (r0 == 0 && halo)
      tile_g2s<T, BS, 1>(sO, static_cast<const T*>(a.O_halo) + (size_t)b * BS, 1, is_aligned16(a.O_halo));
    if (has_y) tile_g2s<T, L, 1>(sy, gy + (size_t)r0 * L, nR, is_aligned16(gy));
    cp_async_wait_all();
    __syncthreads();
  }

  // ---------------- per-node compute ----------------
  const int g = threadIdx.x / LG;          // group in CTA
  const int r = threadIdx.x - g * LG;      // row owned by this lane
  const int e = e0 + g;
  const bool valid = e < E;
  const bool own = valid && (g < OWN);
  const bool rowok = r < L;
  const bool has_right = own && (e < o);                      // odd node e exists (halo group never forms F)
  const bool has_left = valid && (e >= 1 || halo);            // link to odd node e-1 exists
  T* Kb = sR + (size_t)(2 * g) * BS;                          // R_{2e} -> K_e
  T* Ob_left = sO + (size_t)(2 * g) * BS;                     // O_{2e-1} -> G_{e-1}
  T* Ob_right = sO + (size_t)(2 * g + 1) * BS;                // O_{2e}   -> F_e
  const int rr = rowok ? r : 0;                               // idle lanes shadow row 0 (never store)

  double ld_part = 0.0, mh_part = 0.0;
  T inv[L];
  bool bad = false;
  {
    // Cholesky, right-looking, lane = row; pivots / multipliers move by warp shuffle.
    T arow[L];
    T mydiag = T(1);
    if (valid) lds_row<T, L>(arow, Kb + rr * L);
    else {
#pragma unroll
      for (int c = 0; c < L; ++c) arow[c] = (c == rr) ? T(1) : T(0);
    }
#pragma unroll
    for (int k = 0; k < L; ++k) {
      const T dk = shfl_grp(arow[k], k, LG);
      if (!(dk > T(0))) bad = true;
      const T lkk = sqrt(dk);
      inv[k] = T(1) / lkk;
      const T lrk = (rr == k) ? lkk : arow[k] * inv[k];
      arow[k] = lrk;
      if (rr == k) mydiag = lkk;
#pragma unroll
      for (int c = k + 1; c < L; ++c) {
        const T lck = shfl_grp(lrk, c, LG);
        arow[c] = fma(-lrk, lck, arow[c]);
      }
    }
#pragma unroll
    for (int c = 0; c < L; ++c) if (c > rr) arow[c] = T(0);    // exact zeros above the diagonal
    if (valid && rowok) sts_row<T, L>(Kb + r * L, arow);
    if (own && rowok && a.logdet != nullptr) ld_part = log((double)mydiag);
  }
  if (bad && valid && own && a.info != nullptr && r == 0)
    atomicMax(a.info, 0x7fffffff - (b * E + e > 0x7ffffffe ? 0x7ffffffe : b * E + e));

  // read phase: everything other lanes will later overwrite
  T f[L], gl[L], xv[L];
#pragma unroll
  for (int c = 0; c < L; ++c) { f[c] = T(0); gl[c] = T(0); xv[c] = T(0); }
  if (has_right) lds_row<T, L>(f, Ob_right + rr * L);            // row r of O_{2e}
  if (has_left) {
#pragma unroll
    for (int c = 0; c < L; ++c) gl[c] = Ob_left[c * L + rr];     // column r of O_{2e-1}
  }
  if (valid && has_y) lds_row<T, L>(xv, sy + (size_t)(2 * g) * L);
  __syncwarp();   // K rows visible; all column / y reads done

  // forward substitutions against K (broadcast reads of K rows)
#pragma unroll
  for (int c = 0; c < L; ++c) {
    T krow[L];
    lds_row<T, L>(krow, Kb + c * L);
    T sf = f[c], sg = gl[c], sx = xv[c];
#pragma unroll
    for (int k = 0; k < c; ++k) {
      sf = fma(-f[k], krow[k], sf);
      sg = fma(-gl[k], krow[k], sg);
      sx = fma(-xv[k], krow[k], sx);
    }
    f[c] = sf * inv[c]; gl[c] = sg * inv[c]; xv[c] = sx * inv[c];
  }
  if (!valid) {
#pragma unroll
    for (int c = 0; c < L; ++c) { f[c] = T(0); gl[c] = T(0); xv[c] = T(0); }
  }
  if (has_right && rowok) sts_row<T, L>(Ob_right + r * L, f);    // F_e in place
  if (has_left && rowok) sts_row<T, L>(Ob_left + r * L, gl);     // G_{e-1} in place
  if (valid && has_y && rowok) {
    T xr = T(0);
#pragma unroll
    for (int c = 0; c < L; ++c) if (c == r) xr = xv[c];
    sy[(size_t)(2 * g) * L + r] = xr;
    if (own) mh_part = (double)xr * (double)xr;
  }
  __syncwarp();   // F, G rows visible inside the group

  // products
  T arow2[L];   // row r of F F^T
  T u = T(0), v = T(0);
  {
    T brow[L], orow[L];
#pragma unroll
    for (int c = 0; c < L; ++c) { arow2[c] = T(0); brow[c] = T(0); orow[c] = T(0); }
    if (has_right) row_times_matT<T, L>(arow2, f, Ob_right);
    if (has_left) {
#pragma unroll
      for (int c = 0; c < L; ++c) {
        T grow[L];
        lds_row<T, L>(grow, Ob_left + c * L);
        T sb = T(0), so = T(0);
#pragma unroll
        for (int k = 0; k < L; ++k) { sb = fma(gl[k], grow[k], sb); so = fma(-f[k], grow[k], so); }
        brow[c] = sb; orow[c] = so;
      }
    }
#pragma unroll
    for (int k = 0; k < L; ++k) { u = fma(f[k], xv[k], u); v = fma(gl[k], xv[k], v); }
    if (rowok) {
      sts_row<T, L>(sB + (size_t)g * BS + r * L, brow);
      sts_row<T, L>(sOn + (size_t)g * BS + r * L, orow);
      sv[g * L + r] = v;
    }
  }
  __syncthreads();   // B_{e+1}, v_{e+1} of the neighbouring group visible

  if (own && has_right && rowok) {
    const bool next_even = (e + 1) < E;
    T rt[L];
    lds_row<T, L>(rt, sR + (size_t)(2 * g + 1) * BS + r * L);
    if (next_even) {
      T bn[L];
      lds_row<T, L>(bn, sB + (size_t)(g + 1) * BS + r * L);
#pragma unroll
      for (int c = 0; c < L; ++c) rt[c] = rt[c] - arow2[c] - bn[c];
    } else {
#pragma unroll
      for (int c = 0; c < L; ++c) rt[c] = rt[c] - arow2[c];
    }
    sts_row<T, L>(sR + (size_t)(2 * g + 1) * BS + r * L, rt);
    if (has_y) {
      T yt = sy[(size_t)(2 * g + 1) * L + r] - u;
      if (next_even) yt -= sv[(g + 1) * L + r];
      sy[(size_t)(2 * g + 1) * L + r] = yt;
    }
  }

  // scalars: one atomic per CTA per accumulator
  if (a.logdet != nullptr) {
    const double t = block_sum(ld_part, sred);
    if (threadIdx.x == 0) atomicAdd(a.logdet + acc_index(a, b, tile), t);
  }
  if (a.mahal != nullptr && has_y) {
    const double t = block_sum(mh_part, sred);
    if (threadIdx.x == 0) atomicAdd(a.mahal + acc_index(a, b, tile), t);
  }
  __syncthreads();

  // ---------------- stage out ----------------
  const int n_own = cmin(OWN, E - e0);                         // even nodes owned by this CTA
  const int n_odd = cmax(0, cmin(OWN, o - e0));                // odd nodes e0.. with e < o
  if (a.D != nullptr) {
    T* D = static_cast<T*>(a.D) + ((size_t)b * E + e0) * BS;
    tile_s2g<T, BS, 2>(D, sR, n_own, is_aligned16(a.D));
    T* F = static_cast<T*>(a.F) + ((size_t)b * o + e0) * BS;
    tile_s2g<T, BS, 2>(F, sO + BS, n_odd, is_aligned16(a.F));
    // G_{e-1} for own e >= 1 : staged at position 2(e-e0)
    const int gfirst = (e0 == 0) ? 1 : 0;
    const int n_g = cmax(0, cmin(e0 + n_own, gcnt + 1) - (e0 + gfirst));
    T* G = static_cast<T*>(a.G) + ((size_t)b * gcnt + (e0 + gfirst - 1)) * BS;
    tile_s2g<T, BS, 2>(G, sO + (size_t)(2 * gfirst) * BS, n_g, is_aligned16(a.G));
  }
  if (a.xk != nullptr && has_y) {
    T* X = static_cast<T*>(a.xk) + ((size_t)b * E + e0) * L;
    tile_s2g<T, L, 2>(X, sy, n_own, is_aligned16(a.xk));
  }
  if (a.Rn != nullptr && n_odd > 0) {
    T* Rn = static_cast<T*>(a.Rn) + ((size_t)b * o + e0) * BS;
    tile_s2g<T, BS, 2>(Rn, sR + BS, n_odd, is_aligned16(a.Rn));
    if (has_y && a.yn != nullptr) {
      T* yn = static_cast<T*>(a.yn) + ((size_t)b * o + e0) * L;
      tile_s2g<T, L, 2>(yn, sy + L, n_odd, is_aligned16(a.yn));
    }
    // O~_{e-1} for own e with 1 <= e < o : staged flat at sOn[e-e0]
    const int ofirst = (e0 == 0) ? 1 : 0;
    const int n_on = cmax(0, cmin(e0 + n_own, o) - (e0 + ofirst));
    if (a.On != nullptr && n_on > 0) {
      T* On = static_cast<T*>(a.On) + ((size_t)b * (o - 1) + (e0 + ofirst - 1)) * BS;
      tile_s2g<T, BS, 1>(On, sOn + (size_t)ofirst * BS, n_on, is_aligned16(a.On));
    }
  }
  if (halo && e0 == 0) {
    // group 0 owns the link to the virtual node -1
    if (a.G_halo != nullptr) tile_s2g<T, BS, 1>(static_cast<T*>(a.G_halo) + (size_t)b * BS, sO, 1, is_aligned16(a.G_halo));
    if (a.On_halo != nullptr && o > 0) tile_s2g<T, BS, 1>(static_cast<T*>(a.On_halo) + (size_t)b * BS, sOn, 1, is_aligned16(a.On_halo));
    if (a.Rh_acc != nullptr) {
      T* acc = static_cast<T*>(a.Rh_acc) + (size_t)b * BS;
      for (int i = threadIdx.x; i < BS; i += blockDim.x) acc[i] -= sB[i];
    }
    if (a.yh_acc != nullptr && has_y) {
      T* acc = static_cast<T*>(a.yh_acc) + (size_t)b * L;
      for (int i = threadIdx.x; i < L; i += blockDim.x) acc[i] -= sv[i];
    }
  }
}

template <typename T, int L>
cudaError_t launch_level_fwd(const LevelFwdArgs& a, cudaStream_t stream) {
  using C = FwdCfg<T, L>;
  static std::atomic<unsigned char> attr_done[kMaxDevices];
  if (cudaError_t e = ensure_dynamic_smem(cr_level_fwd_kernel<T, L>, (int)C::SMEM, attr_done); e != cudaSuccess) return e;
  const int E = (a.m + 1) / 2;
  const long long tiles = (E + (C::NG - 1) - 1) / (C::NG - 1);
  const long long grid = tiles * a.batch;
  if (grid <= 0) return cudaSuccess;
  if (grid > 0x7fffffffLL) return cudaErrorInvalidValue;
  cr_level_fwd_kernel<T, L><<<(unsigned)grid, C::THREADS, C::SMEM, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace crb200
