// Instantiation unit: compiled once per (dtype, ell range) so the template expansions build
// in parallel.  -DCRB_T=float|double -DCRB_TN=f32|f64 -DCRB_LO=.. -DCRB_HI=..
#include "cr_level_fwd.cuh"
#include "cr_level_bwd.cuh"
#if CRB_LO <= 9   // thread-per-node kernels exist only for small blocks (CRB200_TPN_MAX_BLOCK_BYTES)
#include "cr_tpn_fwd.cuh"
#include "cr_tpn_bwd.cuh"
#include "cr_cs_fwd.cuh"
#include "cr_cs_bwd.cuh"
#define CRB_HAVE_TPN 1
#else
#define CRB_HAVE_TPN 0
#endif
#include "cr_halfsolve.cuh"
#if CRB_HI >= 8   // warp-per-node DMMA kernels: blocks that do not fit one thread's registers (and ell = 8 for comparison)
#include "cr_mma_fwd.cuh"
#include "cr_mma_bwd.cuh"
#define CRB_HAVE_MMA 1
#else
#define CRB_HAVE_MMA 0
#endif
#if !CRB_HAVE_TPN
#include "cr_tpn_common.cuh"   // MultiArgs
#endif

#define CRB_CAT_(a, b, c, d) a##_##b##_##c##_##d
#define CRB_CAT(a, b, c, d) CRB_CAT_(a, b, c, d)

namespace crb200 {

#ifdef CRB_STUB
// development builds (build.py --only ...): this (dtype, ell range) is compiled out
cudaError_t CRB_CAT(inst_fwd, CRB_TN, CRB_LO, CRB_HI)(int, const LevelFwdArgs&, cudaStream_t) { return cudaErrorNotSupported; }
cudaError_t CRB_CAT(inst_bwd, CRB_TN, CRB_LO, CRB_HI)(int, const LevelBwdArgs&, cudaStream_t) { return cudaErrorNotSupported; }
cudaError_t CRB_CAT(inst_hs, CRB_TN, CRB_LO, CRB_HI)(int, const HalfSolveArgs&, cudaStream_t) { return cudaErrorNotSupported; }
int CRB_CAT(inst_fwd_tile, CRB_TN, CRB_LO, CRB_HI)(int) { return 0; }
int CRB_CAT(inst_bwd_tile, CRB_TN, CRB_LO, CRB_HI)(int) { return 0; }
cudaError_t CRB_CAT(inst_fwd_multi, CRB_TN, CRB_LO, CRB_HI)(int, const MultiArgs<LevelFwdArgs>*, cudaStream_t) { return cudaErrorNotSupported; }
cudaError_t CRB_CAT(inst_bwd_multi, CRB_TN, CRB_LO, CRB_HI)(int, const MultiArgs<LevelBwdArgs>*, cudaStream_t) { return cudaErrorNotSupported; }
#else

#if CRB_HAVE_TPN
// lanes per node of the column-split family for (CRB_T, L); 1 = not used
template <int L>
struct CsSel {
  static constexpr int LPN = sizeof(CRB_T) == 4 ? (L == 8 ? 2 : 1) : (L == 8 ? 4 : (L == 4 ? 2 : 1));
  static constexpr bool FWD = CsFwdCfg<CRB_T, L, (LPN > 1 ? LPN : 2)>::ELIGIBLE && LPN > 1;
  static constexpr bool BWD = CsBwdCfg<CRB_T, L, (LPN > 1 ? LPN : 2)>::ELIGIBLE && LPN > 1;
};
#endif

template <int L>
struct Dispatch {
  // The warp-per-node DMMA family exists for ell >= 8.  It is the automatic choice where it measured faster than the
  // lane-per-row kernels on B200 (profiles/r2_ell_sweep.json): fp64 ell >= 10 and fp32 ell >= 17 (fp32 data has half the bytes for
  // the same fp64 arithmetic, so the cross-over sits higher).
  static constexpr bool kMma = CRB_HAVE_MMA && (L >= 8);
  static constexpr bool kMmaAuto = kMma && (sizeof(CRB_T) == 8 ? L >= 10 : L >= 17);
  static constexpr bool kSmallFamilies =
#if CRB_HAVE_TPN
      TpnFwdCfg<CRB_T, L>::ELIGIBLE || CsSel<L>::FWD;
#else
      false;
#endif
  static cudaError_t fwd(int ell, const LevelFwdArgs& a, cudaStream_t s) {
    if (ell == L) {
#if CRB_HAVE_MMA
      if constexpr (kMma) {
        if (a.variant == CRB200_MMA || (a.variant == CRB200_AUTO && !kSmallFamilies && kMmaAuto)) return launch_mma_fwd<CRB_T, L>(a, s);
      } else
#endif
      {
        if (a.variant == CRB200_MMA) return cudaErrorInvalidValue;
      }
#if CRB_HAVE_TPN
      if constexpr (CsSel<L>::FWD) {
        // auto: column-split only where the thread-per-node kernel does not exist (measured: at fp32 l=8 the
        // extra resident warps are paid for by redundant Cholesky work and half-used broadcast reads)
        if (a.variant == CRB200_COLUMN_SPLIT || (a.variant == CRB200_AUTO && !TpnFwdCfg<CRB_T, L>::ELIGIBLE))
          return launch_cs_fwd<CRB_T, L, CsSel<L>::LPN>(a, s);
      } else {
        if (a.variant == CRB200_COLUMN_SPLIT) return cudaErrorInvalidValue;
      }
      if constexpr (TpnFwdCfg<CRB_T, L>::ELIGIBLE) {
        if (a.variant != CRB200_LANE_PER_ROW) return launch_tpn_fwd<CRB_T, L>(a, s);   // incl. CRB200_COPY_ONLY
      } else
#endif
      {
        if (a.variant == CRB200_THREAD_PER_NODE) return cudaErrorInvalidValue;
      }
      return launch_level_fwd<CRB_T, L>(a, s);
    }
    return Dispatch<L + 1>::fwd(ell, a, s);
  }
  static cudaError_t bwd(int ell, const LevelBwdArgs& a, cudaStream_t s) {
    if (ell == L) {
#if CRB_HAVE_MMA
      if constexpr (kMma) {
        if (a.variant == CRB200_MMA || (a.variant == CRB200_AUTO && !kSmallFamilies && kMmaAuto)) return launch_mma_bwd<CRB_T, L>(a, s);
      } else
#endif
      {
        if (a.variant == CRB200_MMA) return cudaErrorInvalidValue;
      }
#if CRB_HAVE_TPN
      if constexpr (CsSel<L>::BWD) {
        if (a.variant == CRB200_COLUMN_SPLIT || (a.variant == CRB200_AUTO && !TpnBwdCfg<CRB_T, L>::ELIGIBLE))
          return launch_cs_bwd<CRB_T, L, CsSel<L>::LPN>(a, s);
      } else {
        if (a.variant == CRB200_COLUMN_SPLIT) return cudaErrorInvalidValue;
      }
      if constexpr (TpnBwdCfg<CRB_T, L>::ELIGIBLE) {
        if (a.variant != CRB200_LANE_PER_ROW) return launch_tpn_bwd<CRB_T, L>(a, s);
      } else
#endif
      {
        if (a.variant == CRB200_THREAD_PER_NODE) return cudaErrorInvalidValue;
      }
      return launch_level_bwd<CRB_T, L>(a, s);
    }
    return Dispatch<L + 1>::bwd(ell, a, s);
  }
  static cudaError_t hs(int ell, const HalfSolveArgs& a, cudaStream_t s) {
    if (ell == L) return launch_level_halfsolve<CRB_T, L>(a, s);
    return Dispatch<L + 1>::hs(ell, a, s);
  }
  // fused deep levels (thread-per-node family only); a == nullptr asks whether the kernel exists
  static cudaError_t fwd_multi(int ell, const MultiArgs<LevelFwdArgs>* a, cudaStream_t s) {
    if (ell != L) return Dispatch<L + 1>::fwd_multi(ell, a, s);
#if CRB_HAVE_TPN
    if constexpr (TpnFwdCfg<CRB_T, L>::ELIGIBLE) return a == nullptr ? cudaSuccess : launch_tpn_fwd_multi<CRB_T, L>(*a, s);
#endif
    return cudaErrorNotSupported;
  }
  static cudaError_t bwd_multi(int ell, const MultiArgs<LevelBwdArgs>* a, cudaStream_t s) {
    if (ell != L) return Dispatch<L + 1>::bwd_multi(ell, a, s);
#if CRB_HAVE_TPN
    if constexpr (TpnBwdCfg<CRB_T, L>::ELIGIBLE) return a == nullptr ? cudaSuccess : launch_tpn_bwd_multi<CRB_T, L>(*a, s);
#endif
    return cudaErrorNotSupported;
  }
  static int fwd_tile(int ell) {
    if (ell != L) return Dispatch<L + 1>::fwd_tile(ell);
#if CRB_HAVE_TPN
    if (TpnFwdCfg<CRB_T, L>::ELIGIBLE) return TpnFwdCfg<CRB_T, L>::OWN;
    if (CsSel<L>::FWD) return CsFwdCfg<CRB_T, L, (CsSel<L>::LPN > 1 ? CsSel<L>::LPN : 2)>::OWN;
#endif
#if CRB_HAVE_MMA
    if (kMmaAuto) return MmaFwdCfg<CRB_T, L>::OWN;
#endif
    return FwdCfg<CRB_T, L>::NG - 1;
  }
  static int bwd_tile(int ell) {
    if (ell != L) return Dispatch<L + 1>::bwd_tile(ell);
#if CRB_HAVE_TPN
    if (TpnBwdCfg<CRB_T, L>::ELIGIBLE) return TpnBwdCfg<CRB_T, L>::NT;
    if (CsSel<L>::BWD) return CsBwdCfg<CRB_T, L, (CsSel<L>::LPN > 1 ? CsSel<L>::LPN : 2)>::NT;
#endif
#if CRB_HAVE_MMA
    if (kMmaAuto) return MmaBwdCfg<CRB_T, L>::NT;
#endif
    return BwdCfg<CRB_T, L>::NG;
  }
};
template <>
struct Dispatch<CRB_HI + 1> {
  static cudaError_t fwd(int, const LevelFwdArgs&, cudaStream_t) { return cudaErrorInvalidValue; }
  static cudaError_t bwd(int, const LevelBwdArgs&, cudaStream_t) { return cudaErrorInvalidValue; }
  static cudaError_t hs(int, const HalfSolveArgs&, cudaStream_t) { return cudaErrorInvalidValue; }
  static int fwd_tile(int) { return 0; }
  static int bwd_tile(int) { return 0; }
  static cudaError_t fwd_multi(int, const MultiArgs<LevelFwdArgs>*, cudaStream_t) { return cudaErrorNotSupported; }
  static cudaError_t bwd_multi(int, const MultiArgs<LevelBwdArgs>*, cudaStream_t) { return cudaErrorNotSupported; }
};

cudaError_t CRB_CAT(inst_fwd, CRB_TN, CRB_LO, CRB_HI)(int ell, const LevelFwdArgs& a, cudaStream_t s) { return Dispatch<CRB_LO>::fwd(ell, a, s); }
cudaError_t CRB_CAT(inst_bwd, CRB_TN, CRB_LO, CRB_HI)(int ell, const LevelBwdArgs& a, cudaStream_t s) { return Dispatch<CRB_LO>::bwd(ell, a, s); }
cudaError_t CRB_CAT(inst_hs, CRB_TN, CRB_LO, CRB_HI)(int ell, const HalfSolveArgs& a, cudaStream_t s) { return Dispatch<CRB_LO>::hs(ell, a, s); }
int CRB_CAT(inst_fwd_tile, CRB_TN, CRB_LO, CRB_HI)(int ell) { return Dispatch<CRB_LO>::fwd_tile(ell); }
int CRB_CAT(inst_bwd_tile, CRB_TN, CRB_LO, CRB_HI)(int ell) { return Dispatch<CRB_LO>::bwd_tile(ell); }
cudaError_t CRB_CAT(inst_fwd_multi, CRB_TN, CRB_LO, CRB_HI)(int ell, const MultiArgs<LevelFwdArgs>* a, cudaStream_t s) { return Dispatch<CRB_LO>::fwd_multi(ell, a, s); }
cudaError_t CRB_CAT(inst_bwd_multi, CRB_TN, CRB_LO, CRB_HI)(int ell, const MultiArgs<LevelBwdArgs>* a, cudaStream_t s) { return Dispatch<CRB_LO>::bwd_multi(ell, a, s); }
#endif  // CRB_STUB

}  // namespace crb200
