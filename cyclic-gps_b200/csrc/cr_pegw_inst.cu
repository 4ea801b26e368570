// Instantiation unit of the warp-per-gap precision builder (cr_pegw.cuh), compiled once per (dtype, ell range):
// -DCRB_T=float|double -DCRB_TN=f32|f64 -DCRB_LO=.. -DCRB_HI=..   (ranges 9-12 ... 29-32)
#include "cr_pegw.cuh"

#define CRB_CAT_(a, b, c, d) a##_##b##_##c##_##d
#define CRB_CAT(a, b, c, d) CRB_CAT_(a, b, c, d)

namespace crb200 {

template <int L>
struct PegwDispatch {
  static cudaError_t fwd(int ell, const PegFwdArgs& a, cudaStream_t s) {
    if (ell == L) return launch_pegw_fwd<CRB_T, L>(a, s);
    return PegwDispatch<L + 1>::fwd(ell, a, s);
  }
  static cudaError_t bwd(int ell, const PegBwdArgs& a, cudaStream_t s) {
    if (ell == L) return launch_pegw_bwd<CRB_T, L>(a, s);
    return PegwDispatch<L + 1>::bwd(ell, a, s);
  }
};
template <>
struct PegwDispatch<CRB_HI + 1> {
  static cudaError_t fwd(int, const PegFwdArgs&, cudaStream_t) { return cudaErrorInvalidValue; }
  static cudaError_t bwd(int, const PegBwdArgs&, cudaStream_t) { return cudaErrorInvalidValue; }
};

cudaError_t CRB_CAT(inst_pegw_fwd, CRB_TN, CRB_LO, CRB_HI)(int ell, const PegFwdArgs& a, cudaStream_t s) { return PegwDispatch<CRB_LO>::fwd(ell, a, s); }
cudaError_t CRB_CAT(inst_pegw_bwd, CRB_TN, CRB_LO, CRB_HI)(int ell, const PegBwdArgs& a, cudaStream_t s) { return PegwDispatch<CRB_LO>::bwd(ell, a, s); }

}  // namespace crb200
