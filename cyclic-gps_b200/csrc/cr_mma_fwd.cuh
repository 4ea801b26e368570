// Large-block forward level kernel: one WARP per even node, block products on the FP64 tensor path (DMMA).
//
// Same contract as cr_level_fwd_kernel (see cr_level_fwd.cuh for the maths and the reference lines it replaces:
// cyclic_gps/cyclic_reduction.py:204-259, :412-427).  Used where a block no longer fits one thread's registers
// (automatic for fp64 ell >= 10, fp32 ell >= 17).  The triangular solves of the reference are restated as products with the explicit
// inverse Ki = K^{-1} (formed once per node by forward substitution), so that everything of order ell^3 is a dense
// product on mma.sync.m8n8k4.f64:
//   K  = chol(R_{2e})                    warp_cholesky (lane = row)            -> D output
//   Ki = K^{-1}                          warp_tri_inverse (lane = column)
//   x  = Ki y_{2e}
//   F  = O_{2e} Ki^T                     DMMA, skipping the zero half of Ki^T
//   G  = O_{2e-1}^T Ki^T                 DMMA (first operand read transposed from the row-major block)
//   O~ = -F G^T ; B = G G^T ; A = F F^T  DMMA
//   R~_e = R_{2e+1} - A_e - B_{e+1} ;  y~_e = y_{2e+1} - F x_e - G_e x_{e+1}
// A CTA is W warps = W consecutive even nodes of one series: W - 1 owned nodes plus the next even node as a
// read-only halo (its B and G x are needed by the last owned node; it skips everything else).  The only exchange
// between warps is B / G x of the right neighbour through shared memory, handed over on a named barrier per pair of
// neighbouring warps (no CTA-wide barrier: the warps of a CTA drift apart and cover each other's serial phases).
// Per node shared memory holds three padded blocks of doubles, each used several times (the halo node two):
//   S0: R_even -> K -> Ki -> R_odd (staged once Ki has been consumed)      S2: O_right -> F      S3: O_left -> G -> B
// O~ and R~ leave straight from the accumulator fragments; K, F, G leave from shared memory as coalesced rows.
#pragma once
#include "cr_level_fwd.cuh"
#include "cr_mma_common.cuh"

namespace crb200 {

template <typename T, int L>
struct MmaFwdCfg {
  using Geo = MmaGeom<L>;
  static constexpr bool ELIGIBLE = (L >= 8);
  static constexpr int LP = Geo::LP, LD = Geo::LD, BLK = Geo::BLK;
  static constexpr int VEC = 6 * LP;                          // y_even -> x | y_odd | G x | spare | column buffer (2 LP) of the Cholesky
  static constexpr int REC = 3 * BLK + VEC;                   // doubles per owned node: S0 | S2 | S3 | vectors
  static constexpr int REC_HALO = 2 * BLK + VEC;              // the halo node never forms F: S0 | S3 | vectors
  // Warps (= nodes incl. the halo) per CTA, chosen so that TWO CTAs fit on an SM: with a single CTA per SM all warps sit in the same
  // phase (load, Cholesky, products, store) and the memory pipe idles while the tensor pipe works and vice versa (LP = 32 ran
  // 2.3x off the sum of its parts); two CTAs start at different times and cover each other's phases.
  static constexpr int W = LP <= 16 ? 8 : (LP <= 24 ? 5 : 4);
  static constexpr int OWN = W - 1;
  static constexpr size_t SMEM = (size_t)(OWN * REC + REC_HALO) * sizeof(double);
  static constexpr int MIN_CTAS = (2 * (SMEM + 1024) <= 227 * 1024) ? 2 : 1;
};

template <typename T, int L>
__global__ void __launch_bounds__(32 * MmaFwdCfg<T, L>::W, MmaFwdCfg<T, L>::MIN_CTAS)
cr_mma_fwd_kernel(const LevelFwdArgs a) {
  using C = MmaFwdCfg<T, L>;
  constexpr int KP = MmaGeom<L>::KP, LA = MmaGeom<L>::LA;
  constexpr int LP = C::LP, LD = C::LD, BLK = C::BLK, OWN = C::OWN, NTL = LP / 8, BS = L * L;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool is_halo = warp == OWN;
  double* N = reinterpret_cast<double*>(smem_raw) + (size_t)warp * C::REC;       // (the halo record is the last one)
  double* S0 = N;                                          // R_even -> K -> Ki -> R_odd (staged once Ki has been consumed)
  double* S2 = N + BLK;                                    // O_right -> F          (owned nodes only)
  double* S3 = is_halo ? N + BLK : N + 2 * BLK;            // O_left -> G -> B
  double* YE = is_halo ? N + 2 * BLK : N + 3 * BLK;
  double* YO = YE + LP;
  double* V = YO + LP;
  double* CB = YE + 4 * LP;                                // column buffer of the Cholesky

  const int m = a.m;
  const int E = (m + 1) >> 1, o = m >> 1, gcnt = (m - 1) >> 1;
  const int tiles = (E + OWN - 1) / OWN;
  const int b = blockIdx.x / tiles;
  const int tile = blockIdx.x - b * tiles;
  const int e0 = tile * OWN;
  const int e = e0 + warp;
  const bool has_y = a.y != nullptr;
  const bool halo = a.O_halo != nullptr;
  const bool valid = e < E;
  const bool own = valid && warp < OWN;
  const bool do_f = own && e < o;                       // odd node e exists and is finished by this warp
  const bool has_left = valid && (e >= 1 || halo);      // link to odd node e-1 exists
  const bool need_B = has_left && (warp > 0 || (halo && e0 == 0));   // warp 0's left neighbour belongs to the previous CTA

  const T* gR = static_cast<const T*>(a.R) + (size_t)b * a.strideR;
  const T* gO = static_cast<const T*>(a.O) + (size_t)b * a.strideO;
  const T* gy = has_y ? static_cast<const T*>(a.y) + (size_t)b * a.stridey : nullptr;
  const bool vR = is_aligned16(gR), vO = is_aligned16(gO);

  // ---------------- stage in ----------------
  // vectors first (plain loads whose latency is covered by the block copies issued next)
  double ye_v = 0.0, yo_v = 0.0;
  if (has_y && lane < L) {
    if (valid) ye_v = (double)gy[(size_t)(2 * e) * L + lane];
    if (do_f) yo_v = (double)gy[(size_t)(2 * e + 1) * L + lane];
  }
  const T* srcL = has_left ? (e >= 1 ? gO + (size_t)(2 * e - 1) * BS : static_cast<const T*>(a.O_halo) + (size_t)b * BS) : nullptr;
  const bool vL = e >= 1 ? vO : is_aligned16(a.O_halo);
  if (valid) mma_stage_issue<T, L, LP>(S0, gR + (size_t)(2 * e) * BS, true, lane, vR);
  else mma_fill_block<LP>(S0, true, lane);
  if (has_left) mma_stage_issue<T, L, LP>(S3, srcL, false, lane, vL);
  else mma_fill_block<LP>(S3, false, lane);
  if (do_f) mma_stage_issue<T, L, LP>(S2, gO + (size_t)(2 * e) * BS, false, lane, vO);
  if (lane < LP) {
    YE[lane] = ye_v;
    YO[lane] = yo_v;
    V[lane] = 0.0;
  }
  cp_async_wait_all();
  __syncwarp();
  if (valid) mma_stage_finish<T, L, LP>(S0, true, lane);
  if (has_left) mma_stage_finish<T, L, LP>(S3, false, lane);
  if (do_f) mma_stage_finish<T, L, LP>(S2, false, lane);
  __syncwarp();

  // ---------------- K, Ki, x ----------------
  double ld_part = 0.0, mh_part = 0.0;
  {
    double invd[LP];
    const bool bad = warp_cholesky<LP, LA>(S0, CB, invd, lane);
    if (own) {
      if (bad && a.info != nullptr && lane == 0) {
        const long long flat = (long long)b * E + e;
        atomicMax(a.info, 0x7fffffff - (int)(flat > 0x7ffffffeLL ? 0x7ffffffeLL : flat));
      }
      if (a.logdet != nullptr && lane == 0) {
        double p = 1.0;
#pragma unroll
        for (int j = 0; j < L; ++j) p *= invd[j];
        ld_part = -log(p);
      }
      if (a.D != nullptr) mma_store_block<T, L, LP>(static_cast<T*>(a.D) + ((size_t)b * E + e) * BS, S0, lane, is_aligned16(a.D));
    }
    warp_tri_inverse<LP, LA>(S0, invd, lane);     // S0 = Ki
  }
  if (has_y) {
    const double xr = warp_matvec<LP, false>(S0, YE, lane);     // x = Ki y_even
    __syncwarp();
    if (lane < LP) YE[lane] = xr;
    __syncwarp();
    if (own) {
      mh_part = xr * xr;
      if (a.xk != nullptr && lane < L) static_cast<T*>(a.xk)[((size_t)b * E + e) * L + lane] = (T)xr;
    }
  }

  // ---------------- F, G, their products ----------------
  double acc[NTL][NTL][2];
  if (do_f) {
    acc_zero<LP>(acc);
    warp_gemm<LP, false, true, K_LE_N, false, KP>(acc, S2, S0, lane);          // F = O_right Ki^T
    __syncwarp();
    acc_to_smem<LP>(S2, acc, 1.0, lane);
    __syncwarp();
    if (a.D != nullptr) mma_store_block<T, L, LP>(static_cast<T*>(a.F) + ((size_t)b * o + e) * BS, S2, lane, is_aligned16(a.F));
  }
  if (has_left) {
    acc_zero<LP>(acc);
    warp_gemm<LP, true, true, K_LE_N, false, KP>(acc, S3, S0, lane);           // G = O_left^T Ki^T
    __syncwarp();
    acc_to_smem<LP>(S3, acc, 1.0, lane);
    __syncwarp();
    if (own) {
      if (e >= 1) {
        if (a.D != nullptr) mma_store_block<T, L, LP>(static_cast<T*>(a.G) + ((size_t)b * gcnt + (e - 1)) * BS, S3, lane, is_aligned16(a.G));
      } else if (a.G_halo != nullptr) {
        mma_store_block<T, L, LP>(static_cast<T*>(a.G_halo) + (size_t)b * BS, S3, lane, is_aligned16(a.G_halo));
      }
    }
  }
  // Ki has been consumed (x, F, G): its slot takes R_odd, which arrives while the remaining products run
  __syncwarp();
  if (do_f) mma_stage_issue<T, L, LP>(S0, gR + (size_t)(2 * e + 1) * BS, false, lane, vR);
  double ur = 0.0, vr = 0.0;
  if (has_y) {
    if (do_f) ur = warp_matvec<LP, false>(S2, YE, lane);                   // F x
    if (has_left) vr = warp_matvec<LP, false>(S3, YE, lane);               // G x
    if (lane < LP) V[lane] = vr;
  }
  if (do_f && has_left) {
    T* base = static_cast<T*>(e >= 1 ? a.On : a.On_halo);
    if (base != nullptr) {
      acc_zero<LP>(acc);
      warp_gemm<LP, false, true, K_FULL, false, KP>(acc, S2, S3, lane);        // F G^T
      T* dst = base + ((e >= 1) ? ((size_t)b * (o - 1) + (e - 1)) * BS : (size_t)b * BS);
      acc_to_global<T, L, LP>(dst, acc, -1.0, lane, is_aligned16(base));   // O~_{e-1} = -F G^T
    }
  }
  if (need_B) {
    acc_zero<LP>(acc);
    warp_gemm<LP, false, true, K_FULL, false, KP>(acc, S3, S3, lane);          // B = G G^T
    if (warp == 0) {
      // link to the virtual node -1 (chunk-partitioned series): accumulate -G G^T and -G x there
      if (a.Rh_acc != nullptr) {
        T* accp = static_cast<T*>(a.Rh_acc) + (size_t)b * BS;
        const int lr = lane >> 2, lc = lane & 3;
#pragma unroll
        for (int mt = 0; mt < NTL; ++mt)
#pragma unroll
          for (int nt = 0; nt < NTL; ++nt) {
            const int row = mt * 8 + lr, col = nt * 8 + 2 * lc;
            if (row < L && col < L) accp[row * L + col] = (T)((double)accp[row * L + col] - acc[mt][nt][0]);
            if (row < L && col + 1 < L) accp[row * L + col + 1] = (T)((double)accp[row * L + col + 1] - acc[mt][nt][1]);
          }
      }
      if (a.yh_acc != nullptr && has_y && lane < L) {
        T* accp = static_cast<T*>(a.yh_acc) + (size_t)b * L;
        accp[lane] = (T)((double)accp[lane] - vr);
      }
    } else {
      __syncwarp();                                   // G has left for global memory and is not needed any more
      acc_to_smem<LP>(S3, acc, 1.0, lane);            // B over G, read by the warp to the left after the hand-over
    }
  }
  if (warp > 0) pair_arrive(warp);                    // B and G x of this node are in shared memory (barrier id = consumer warp + 1)
  if (do_f) {
    acc_zero<LP>(acc);
    warp_gemm<LP, false, true, K_FULL, false, KP>(acc, S2, S2, lane);          // A = F F^T, kept in registers
  }
  if (do_f) {                                         // R_odd has landed in S0
    cp_async_wait_all();
    __syncwarp();
    mma_stage_finish<T, L, LP>(S0, false, lane);
    __syncwarp();
  }
  if (warp < OWN) pair_wait(warp + 1);                // B and G x of the right neighbour are visible

  // ---------------- R~_e = R_odd - A - B_{e+1},  y~_e = y_odd - F x_e - G_e x_{e+1} ----------------
  if (do_f && a.Rn != nullptr) {
    const bool next_even = (e + 1) < E;
    const bool next_halo = (warp + 1) == OWN;         // the record of the halo warp is laid out S0 | S3 | vectors
    const double* Bn = N + C::REC + (next_halo ? BLK : 2 * BLK);                  // S3 of the next warp
    const double* Vn = N + C::REC + (next_halo ? 2 * BLK : 3 * BLK) + 2 * LP;     // V of the next warp
    T* Rn = static_cast<T*>(a.Rn) + ((size_t)b * o + e) * BS;
    const bool vec = is_aligned16(a.Rn);
    const int lr = lane >> 2, lc = lane & 3;
#pragma unroll
    for (int mt = 0; mt < NTL; ++mt)
#pragma unroll
      for (int nt = 0; nt < NTL; ++nt) {
        const int row = mt * 8 + lr, col = nt * 8 + 2 * lc;
        const double2 ro = *reinterpret_cast<const double2*>(S0 + row * LD + col);
        double v0 = ro.x - acc[mt][nt][0], v1 = ro.y - acc[mt][nt][1];
        if (next_even) {
          const double2 bn = *reinterpret_cast<const double2*>(Bn + row * LD + col);
          v0 -= bn.x; v1 -= bn.y;
        }
        frag_pair_store<T, L>(Rn, row, col, v0, v1, vec);
      }
    if (has_y && a.yn != nullptr && lane < L)
      static_cast<T*>(a.yn)[((size_t)b * o + e) * L + lane] = (T)(YO[lane] - ur - (next_even ? Vn[lane] : 0.0));
  }

  // ---------------- scalars: one atomic per node ----------------
  if (a.logdet != nullptr && own && lane == 0) atomicAdd(a.logdet + acc_index(a, b, tile), ld_part);
  if (a.mahal != nullptr && has_y) {
    mh_part = warp_sum(mh_part);
    if (own && lane == 0) atomicAdd(a.mahal + acc_index(a, b, tile), mh_part);
  }
}

template <typename T, int L>
cudaError_t launch_mma_fwd(const LevelFwdArgs& a, cudaStream_t stream) {
  using C = MmaFwdCfg<T, L>;
  static std::atomic<unsigned char> attr_done[kMaxDevices];
  if (cudaError_t e = ensure_dynamic_smem(cr_mma_fwd_kernel<T, L>, (int)C::SMEM, attr_done); e != cudaSuccess) return e;
  const int E = (a.m + 1) / 2;
  const long long tiles = (E + C::OWN - 1) / C::OWN;
  const long long grid = tiles * a.batch;
  if (grid <= 0) return cudaSuccess;
  if (grid > 0x7fffffffLL) return cudaErrorInvalidValue;
  cr_mma_fwd_kernel<T, L><<<(unsigned)grid, 32 * C::W, C::SMEM, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace crb200
