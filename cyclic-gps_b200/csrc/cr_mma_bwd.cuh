// Large-block backward level kernel: one WARP per even node, block products on the FP64 tensor path (DMMA).
//
// Same contract as cr_level_bwd_kernel (cr_level_bwd.cuh: back-half-solve + selected inverse + optional gradient
// assembly; reference cyclic_gps/cyclic_reduction.py:362-373, :478-501).  Per even node e, with Di = D_e^{-1}:
//   P = F_e Di ; Q = G_{e-1} Di                                             DMMA, skipping the zero half of Di
//   w_{2e}          = Di^T x_e - P^T w~_e - Q^T w~_{e-1}
//   N1 = S~_d[e] P + S~_o[e-1] Q               Sigma_{2e+1,2e} = -N1        DMMA
//   N2 = Q^T S~_d[e-1]^T + P^T S~_o[e-1]       Sigma_{2e,2e-1} = -N2        DMMA
//   Sigma_{2e,2e}   = Di^T Di + P^T N1 + Q^T N2^T                           DMMA, lower tiles only, then mirrored
// Shared memory per node: five padded blocks of doubles, reused as their first content dies:
//   T0: D -> Di -> N1     T1: F -> P (-> mirror scratch)     T2: G -> Q     T3: S~_d[e] (read by the right neighbour)
//   T4: S~_o[e-1] -> N2
// The three results leave straight from the accumulator fragments (with the gradient transform applied on the way at
// the top level); the odd rows (S~_d[e], w~_e) are copied through from shared memory.  A CTA is W warps = W
// consecutive even nodes; the only data shared between warps are S~_d[e-1] and w~_{e-1}, read from the left
// neighbour's record after ONE __syncthreads that follows the stage-in (warp 0 stages its own copy).
//
// Two warps per node for LP >= 24 (TW = 2).  There a node costs 25-46 KB of shared memory, so with one warp per node an SM holds
// one warp per sub-partition and DMMA time, the serial triangular inverse, staging latency and stores simply add up (ncu, LP = 32:
// 29.6k cycles per node of which 13.9k are DMMA).  The products of a node come in independent pairs, so a TEAM of two warps shares it:
//   warp A: Di, P, Di^T Di, N1, P^T N1, Sigma_{2e+1,2e} and Sigma_{2e,2e} out      warp B: Q, w, N2, Q^T N2^T, Sigma_{2e,2e-1} out, odd rows
// with five named-barrier hand-overs (Di; P and Q; w and the end of Di; S~_o read before N2 overwrites it; B's part of Sigma_ee through
// the dead Q slot).  Same shared memory per node, twice the warps per SM.  With TW = 1 one warp runs both roles in the same order.
#pragma once
#include "cr_level_bwd.cuh"
#include "cr_mma_common.cuh"

namespace crb200 {

template <typename T, int L>
struct MmaBwdCfg {
  using Geo = MmaGeom<L>;
  static constexpr bool ELIGIBLE = (L >= 8);
  static constexpr int LP = Geo::LP, LD = Geo::LD, BLK = Geo::BLK;
  static constexpr int VEC = 4 * LP;                          // x_e | w~_e | w_{2e} | spare
  static constexpr int REC = 5 * BLK + VEC;                   // doubles per node
  static constexpr int LEFT = BLK + VEC;                      // record of the left neighbour of warp 0: S~_d[e0-1], w~_{e0-1}
  // warps (= nodes) per CTA, chosen so that TWO CTAs fit on an SM (their load / compute / store phases then overlap; see cr_mma_fwd.cuh)
  static constexpr int W = LP <= 16 ? 8 : (LP <= 24 ? 3 : 2);
  static constexpr int TW = LP >= 24 ? 2 : 1;                 // warps per node (see the header)
  static constexpr int NT = W;
  static constexpr size_t SMEM = (size_t)(LEFT + W * REC) * sizeof(double);
  static constexpr int MIN_CTAS = (2 * (SMEM + 1024) <= 227 * 1024) ? 2 : 1;
};

template <typename T, int L>
__global__ void __launch_bounds__(32 * MmaBwdCfg<T, L>::W * MmaBwdCfg<T, L>::TW, MmaBwdCfg<T, L>::MIN_CTAS)
cr_mma_bwd_kernel(const LevelBwdArgs a) {
  using C = MmaBwdCfg<T, L>;
  constexpr int KP = MmaGeom<L>::KP, LA = MmaGeom<L>::LA;
  constexpr int LP = C::LP, LD = C::LD, BLK = C::BLK, NT = C::NT, NTL = LP / 8, BS = L * L, TW = C::TW;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = (threadIdx.x >> 5) / TW;               // node of this CTA
  const int sub = (threadIdx.x >> 5) % TW;                // role inside the node's team
  const bool roleA = TW == 1 || sub == 0, roleB = TW == 1 || sub == 1;
  auto team_sync = [&]() {
    if constexpr (TW == 1) __syncwarp();
    else asm volatile("bar.sync %0, 64;\n" ::"r"(warp + 1) : "memory");
  };
  double* base = reinterpret_cast<double*>(smem_raw);
  double* N = base + C::LEFT + (size_t)warp * C::REC;
  double* T0 = N;
  double* T1 = N + BLK;
  double* T2 = N + 2 * BLK;
  double* T3 = N + 3 * BLK;
  double* T4 = N + 4 * BLK;
  double* X = N + 5 * BLK;
  double* WT = X + LP;
  double* WV = WT + LP;
  // left neighbour: S~_d[e-1], w~_{e-1}
  const double* LSD = warp == 0 ? base : (N - C::REC + 3 * BLK);
  const double* LWT = warp == 0 ? base + BLK : (N - C::REC + 5 * BLK + LP);

  const int m = a.m;
  const int E = (m + 1) >> 1, o = m >> 1, gcnt = (m - 1) >> 1;
  const int tiles = (E + NT - 1) / NT;
  const int b = blockIdx.x / tiles;
  const int tile = blockIdx.x - b * tiles;
  const int e0 = tile * NT;
  const int e = e0 + warp;
  const bool do_sigma = a.Sd_out != nullptr;
  const bool do_w = a.w_out != nullptr;
  const bool halo = a.G_halo != nullptr;
  const bool valid = e < E;
  const bool has_odd = valid && e < o;
  const bool has_left = valid && (e >= 1 || halo);
  const bool has_so = has_left && has_odd;
  const bool grad = a.grad_mode != 0;
  double gm = 0.0, gd = 1.0;
  if (grad) {
    gm = a.gm != nullptr ? a.gm[b] : 0.0;
    gd = a.gd != nullptr ? a.gd[b] : 0.0;
  }

  // ---------------- stage in (role A: D, F, S~_d and the vectors; role B: G, S~_o and the left neighbour of the CTA) ----------------
  // vectors first (plain loads whose latency is covered by the block copies issued next)
  double x_v = 0.0, wt_v = 0.0, lw_v = 0.0;
  if (roleA && do_w && lane < L) {
    if (valid) x_v = (double)static_cast<const T*>(a.xk)[((size_t)b * E + e) * L + lane];
    if (has_odd) wt_v = (double)static_cast<const T*>(a.w_in)[((size_t)b * o + e) * L + lane];
    if (warp == 0) {
      if (e0 >= 1) lw_v = (double)static_cast<const T*>(a.w_in)[((size_t)b * o + (e0 - 1)) * L + lane];
      else if (halo) lw_v = (double)static_cast<const T*>(a.w_halo)[(size_t)b * L + lane];
    }
  }
  const bool st3 = do_sigma && has_odd, st4 = do_sigma && has_so, stl = do_sigma && warp == 0 && (e0 >= 1 || halo);
  if (roleA) {
    if (valid) mma_stage_issue<T, L, LP>(T0, static_cast<const T*>(a.D) + ((size_t)b * E + e) * BS, true, lane, is_aligned16(a.D));
    else mma_fill_block<LP>(T0, true, lane);
    if (has_odd) mma_stage_issue<T, L, LP>(T1, static_cast<const T*>(a.F) + ((size_t)b * o + e) * BS, false, lane, is_aligned16(a.F));
    else mma_fill_block<LP>(T1, false, lane);
    if (do_sigma) {
      if (st3) mma_stage_issue<T, L, LP>(T3, static_cast<const T*>(a.Sd_in) + ((size_t)b * o + e) * BS, false, lane, is_aligned16(a.Sd_in));
      else mma_fill_block<LP>(T3, false, lane);
    }
    if (lane < LP) {
      X[lane] = x_v;
      WT[lane] = wt_v;
      WV[lane] = 0.0;
      if (warp == 0) base[BLK + lane] = lw_v;
    }
  }
  if (roleB) {
    if (has_left) {
      if (e >= 1) mma_stage_issue<T, L, LP>(T2, static_cast<const T*>(a.G) + ((size_t)b * gcnt + (e - 1)) * BS, false, lane, is_aligned16(a.G));
      else mma_stage_issue<T, L, LP>(T2, static_cast<const T*>(a.G_halo) + (size_t)b * BS, false, lane, is_aligned16(a.G_halo));
    } else {
      mma_fill_block<LP>(T2, false, lane);
    }
    if (do_sigma) {
      if (st4) {
        if (e >= 1) mma_stage_issue<T, L, LP>(T4, static_cast<const T*>(a.So_in) + ((size_t)b * (o - 1) + (e - 1)) * BS, false, lane, is_aligned16(a.So_in));
        else mma_stage_issue<T, L, LP>(T4, static_cast<const T*>(a.So_halo_in) + (size_t)b * BS, false, lane, is_aligned16(a.So_halo_in));
      } else {
        mma_fill_block<LP>(T4, false, lane);
      }
      if (warp == 0) {
        if (e0 >= 1) mma_stage_issue<T, L, LP>(base, static_cast<const T*>(a.Sd_in) + ((size_t)b * o + (e0 - 1)) * BS, false, lane, is_aligned16(a.Sd_in));
        else if (halo) mma_stage_issue<T, L, LP>(base, static_cast<const T*>(a.Sd_halo) + (size_t)b * BS, false, lane, is_aligned16(a.Sd_halo));
        else mma_fill_block<LP>(base, false, lane);
      }
    }
  }
  cp_async_wait_all();
  __syncwarp();
  if (roleA) {
    if (valid) mma_stage_finish<T, L, LP>(T0, true, lane);
    if (has_odd) mma_stage_finish<T, L, LP>(T1, false, lane);
    if (st3) mma_stage_finish<T, L, LP>(T3, false, lane);
  }
  if (roleB) {
    if (has_left) mma_stage_finish<T, L, LP>(T2, false, lane);
    if (st4) mma_stage_finish<T, L, LP>(T4, false, lane);
    if (stl) mma_stage_finish<T, L, LP>(base, false, lane);
  }
  __syncthreads();                                      // everything staged is visible, incl. the left neighbour's S~_d and w~

  // ---------------- Di (A), P (A), Q (B) ----------------
  if (roleA) {
    double invd[LP];
    const int r = lane < LP ? lane : 0;
    const double mine = 1.0 / T0[r * LD + r];
#pragma unroll
    for (int j = 0; j < LP; ++j) invd[j] = __shfl_sync(0xffffffffu, mine, j);
    warp_tri_inverse<LP, LA>(T0, invd, lane);               // T0 = Di
  }
  team_sync();                                          // (1) Di is ready
  double acc[NTL][NTL][2];
  if (roleA && has_odd) {
    acc_zero<LP>(acc);
    warp_gemm<LP, false, false, K_GE_N, false, KP>(acc, T1, T0, lane);         // P = F Di
    __syncwarp();
    acc_to_smem<LP>(T1, acc, 1.0, lane);
  }
  if (roleB && has_left) {
    acc_zero<LP>(acc);
    warp_gemm<LP, false, false, K_GE_N, false, KP>(acc, T2, T0, lane);         // Q = G Di
    __syncwarp();
    acc_to_smem<LP>(T2, acc, 1.0, lane);
  }
  team_sync();                                          // (2) P and Q are ready
  double wv = 0.0;
  if (roleB && do_w) {
    wv = warp_matvec<LP, true>(T0, X, lane);                               // Di^T x_e
    if (has_odd) wv -= warp_matvec<LP, true>(T1, WT, lane);                // - P^T w~_e
    if (has_left) wv -= warp_matvec<LP, true>(T2, LWT, lane);              // - Q^T w~_{e-1}
    if (lane < LP) WV[lane] = wv;
  }

  const int lr = lane >> 2, lc = lane & 3;
  if (do_sigma) {
    T* const gSd = static_cast<T*>(a.Sd_out) + (size_t)b * a.strideSd;
    T* const gSo = a.So_out != nullptr ? static_cast<T*>(a.So_out) + (size_t)b * a.strideSo : nullptr;
    const bool vSd = is_aligned16(gSd), vSo = is_aligned16(gSo);
    double accE[NTL][NTL][2];
    acc_zero<LP>(accE);
    if (roleA && valid) warp_gemm<LP, true, false, K_GE_MAX_MN, true, KP>(accE, T0, T0, lane);     // Di^T Di (lower tiles)
    team_sync();                                          // (3) Di is dead (T0 may take N1), w_{2e} is published
    // A: N1 = S~_d[e] P + S~_o[e-1] Q ;  Sigma_{2e+1,2e} = -N1  -> So_out row 2e
    if (roleA) {
      if (has_odd) {
        acc_zero<LP>(acc);
        warp_gemm<LP, false, false, K_FULL, false, KP>(acc, T3, T1, lane);
        if (has_so) warp_gemm<LP, false, false, K_FULL, false, KP>(acc, T4, T2, lane);
        acc_to_smem<LP>(T0, acc, 1.0, lane);
        T* dst = gSo + (size_t)(2 * e) * BS;
#pragma unroll
        for (int mt = 0; mt < NTL; ++mt)
#pragma unroll
          for (int nt = 0; nt < NTL; ++nt) {
            const int row = mt * 8 + lr, col = nt * 8 + 2 * lc;
            double v0 = -acc[mt][nt][0], v1 = -acc[mt][nt][1];
            if (grad) {     // gO_{2e} = 2 (gd Sigma_{2e+1,2e} - gm w~_e w_{2e}^T)
              const double wr = WT[row];
              v0 = 2.0 * (gd * v0 - gm * wr * WV[col]);
              v1 = 2.0 * (gd * v1 - gm * wr * WV[col + 1]);
            }
            frag_pair_store<T, L>(dst, row, col, v0, v1, vSo);
          }
      } else {
        mma_fill_block<LP>(T0, false, lane);
      }
    }
    // B: N2 = Q^T S~_d[e-1]^T + P^T S~_o[e-1] ;  Sigma_{2e,2e-1} = -N2  -> So_out row 2e-1 (or the halo block)
    if (roleB && has_left) {
      acc_zero<LP>(acc);
      warp_gemm<LP, true, true, K_FULL, false, KP>(acc, T2, LSD, lane);
      if (has_so) warp_gemm<LP, true, false, K_FULL, false, KP>(acc, T1, T4, lane);
    }
    team_sync();                                          // (4) S~_o[e-1] has been read by both products: T4 may take N2
    if (roleB) {
      if (has_left) {
        acc_to_smem<LP>(T4, acc, 1.0, lane);
        T* dst = e >= 1 ? (gSo != nullptr ? gSo + (size_t)(2 * e - 1) * BS : nullptr)
                        : (a.So_halo_out != nullptr ? static_cast<T*>(a.So_halo_out) + (size_t)b * BS : nullptr);
        if (dst != nullptr) {
          const bool vec = e >= 1 ? vSo : is_aligned16(a.So_halo_out);
#pragma unroll
          for (int mt = 0; mt < NTL; ++mt)
#pragma unroll
            for (int nt = 0; nt < NTL; ++nt) {
              const int row = mt * 8 + lr, col = nt * 8 + 2 * lc;
              double v0 = -acc[mt][nt][0], v1 = -acc[mt][nt][1];
              if (grad) {   // gO_{2e-1} = 2 (gd Sigma_{2e,2e-1} - gm w_{2e} w~_{e-1}^T)
                const double wr = WV[row];
                v0 = 2.0 * (gd * v0 - gm * wr * LWT[col]);
                v1 = 2.0 * (gd * v1 - gm * wr * LWT[col + 1]);
              }
              frag_pair_store<T, L>(dst, row, col, v0, v1, vec);
            }
        }
      } else {
        mma_fill_block<LP>(T4, false, lane);
      }
    }
    __syncwarp();                                         // N1 (A) / N2 (B) written by this warp are visible to its own lanes
    // Sigma_{2e,2e} = Di^T Di + P^T N1 (A) + Q^T N2^T (B; handed to A through the dead Q slot), lower tiles, mirrored through T1
    if (valid) {
      if (roleA && has_odd) warp_gemm<LP, true, false, K_FULL, true, KP>(accE, T1, T0, lane);
      if constexpr (TW == 1) {
        if (has_left) warp_gemm<LP, true, true, K_FULL, true, KP>(accE, T2, T4, lane);
      } else {
        if (roleB) {
          if (has_left) warp_gemm<LP, true, true, K_FULL, true, KP>(accE, T2, T4, lane);
          __syncwarp();                                   // Q has been read by every lane of B (A is done with it since (4))
          acc_to_smem<LP>(T2, accE, 1.0, lane);
        }
      }
    }
    team_sync();                                          // (5) B's part of Sigma_ee is in T2; P is dead for both warps
    if (roleA && valid) {
      if constexpr (TW == 2) {
#pragma unroll
        for (int mt = 0; mt < NTL; ++mt)
#pragma unroll
          for (int nt = 0; nt <= mt; ++nt) {
            const double2 q = *reinterpret_cast<const double2*>(T2 + (mt * 8 + lr) * LD + nt * 8 + 2 * lc);
            accE[mt][nt][0] += q.x;
            accE[mt][nt][1] += q.y;
          }
      }
      acc_to_smem<LP>(T1, accE, 1.0, lane);               // T1 is the mirror scratch
      __syncwarp();
      acc_mirror_from_smem<LP>(accE, T1, 1.0, lane);
      T* dst = gSd + (size_t)(2 * e) * BS;
#pragma unroll
      for (int mt = 0; mt < NTL; ++mt)
#pragma unroll
        for (int nt = 0; nt < NTL; ++nt) {
          const int row = mt * 8 + lr, col = nt * 8 + 2 * lc;
          double v0 = accE[mt][nt][0], v1 = accE[mt][nt][1];
          if (grad) {     // gR_{2e} = gd Sigma_{2e,2e} - gm w_{2e} w_{2e}^T
            const double wr = WV[row];
            v0 = gd * v0 - gm * wr * WV[col];
            v1 = gd * v1 - gm * wr * WV[col + 1];
          }
          frag_pair_store<T, L>(dst, row, col, v0, v1, vSd);
        }
    }
    // odd row copied through (B): Sigma_{2e+1,2e+1} = S~_d[e]   (gR_{2e+1} = gd S~_d[e] - gm w~_e w~_e^T)
    if (roleB && has_odd) {
      T* dst = gSd + (size_t)(2 * e + 1) * BS;
      if (grad) {
        const double* wt = WT;
        const double gdl = gd, gml = gm;
        mma_store_block_f<T, L, LP>(dst, T3, lane, vSd, [=](int r, int c, double v) { return gdl * v - gml * wt[r] * wt[c]; });
      } else {
        mma_store_block<T, L, LP>(dst, T3, lane, vSd);
      }
    }
  }
  if (roleB && do_w && lane < L) {
    T* gW = static_cast<T*>(a.w_out) + (size_t)b * a.stridew;
    const double sc = grad ? 2.0 * gm : 1.0;                // gx = 2 gm w
    if (valid) gW[(size_t)(2 * e) * L + lane] = (T)(sc * wv);
    if (has_odd) gW[(size_t)(2 * e + 1) * L + lane] = (T)(sc * WT[lane]);
  }
}

template <typename T, int L>
cudaError_t launch_mma_bwd(const LevelBwdArgs& a, cudaStream_t stream) {
  using C = MmaBwdCfg<T, L>;
  static std::atomic<unsigned char> attr_done[kMaxDevices];
  if (cudaError_t e = ensure_dynamic_smem(cr_mma_bwd_kernel<T, L>, (int)C::SMEM, attr_done); e != cudaSuccess) return e;
  const int E = (a.m + 1) / 2;
  const long long tiles = (E + C::NT - 1) / C::NT;
  const long long grid = tiles * a.batch;
  if (grid <= 0) return cudaSuccess;
  if (grid > 0x7fffffffLL) return cudaErrorInvalidValue;
  cr_mma_bwd_kernel<T, L><<<(unsigned)grid, 32 * C::W * C::TW, C::SMEM, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace crb200
