// Instantiations + C entry points of the precision-block builder (cr_peg.cuh, include/crb200.h).
#include "cr_peg.cuh"

namespace crb200 {
// warp-per-gap kernels for ell = 9..32 (cr_pegw.cuh), one translation unit per (dtype, ell range)
#define CRB_PEGW_DECL(TN, LO, HI)                                                               \
  cudaError_t inst_pegw_fwd_##TN##_##LO##_##HI(int, const crb200_peg_fwd_args&, cudaStream_t);  \
  cudaError_t inst_pegw_bwd_##TN##_##LO##_##HI(int, const crb200_peg_bwd_args&, cudaStream_t);
#define CRB_PEGW_RANGES(X, TN) X(TN, 9, 12) X(TN, 13, 16) X(TN, 17, 20) X(TN, 21, 24) X(TN, 25, 28) X(TN, 29, 32)
CRB_PEGW_RANGES(CRB_PEGW_DECL, f32)
CRB_PEGW_RANGES(CRB_PEGW_DECL, f64)
constexpr int kPegWideMaxEll = 32;
}  // namespace crb200

namespace {
thread_local int g_peg_last_error = 0;

#define CRB_PEGW_FWD(TN, LO, HI) if (ell >= LO && ell <= HI) return crb200::inst_pegw_fwd_##TN##_##LO##_##HI(ell, a, s);
#define CRB_PEGW_BWD(TN, LO, HI) if (ell >= LO && ell <= HI) return crb200::inst_pegw_bwd_##TN##_##LO##_##HI(ell, a, s);
cudaError_t wide_fwd(int dtype, int ell, const crb200_peg_fwd_args& a, cudaStream_t s) {
  if (dtype == CRB200_F32) { CRB_PEGW_RANGES(CRB_PEGW_FWD, f32) } else { CRB_PEGW_RANGES(CRB_PEGW_FWD, f64) }
  return cudaErrorInvalidValue;
}
cudaError_t wide_bwd(int dtype, int ell, const crb200_peg_bwd_args& a, cudaStream_t s) {
  if (dtype == CRB200_F32) { CRB_PEGW_RANGES(CRB_PEGW_BWD, f32) } else { CRB_PEGW_RANGES(CRB_PEGW_BWD, f64) }
  return cudaErrorInvalidValue;
}

template <typename T, int L>
struct PegDispatch {
  static cudaError_t fwd(int ell, const crb200_peg_fwd_args& a, cudaStream_t s) {
    if (ell == L) return crb200::launch_peg_fwd<T, L>(a, s);
    return PegDispatch<T, L + 1>::fwd(ell, a, s);
  }
  static cudaError_t bwd(int ell, const crb200_peg_bwd_args& a, cudaStream_t s) {
    if (ell == L) return crb200::launch_peg_bwd<T, L>(a, s);
    return PegDispatch<T, L + 1>::bwd(ell, a, s);
  }
};
template <typename T>
struct PegDispatch<T, crb200::kPegMaxEll + 1> {
  static cudaError_t fwd(int, const crb200_peg_fwd_args&, cudaStream_t) { return cudaErrorInvalidValue; }
  static cudaError_t bwd(int, const crb200_peg_bwd_args&, cudaStream_t) { return cudaErrorInvalidValue; }
};

int finish(cudaError_t e) {
  if (e == cudaSuccess) return CRB200_OK;
  g_peg_last_error = (int)e;
  return e == cudaErrorInvalidValue ? CRB200_EINVAL : CRB200_ECUDA;
}
}  // namespace

extern "C" {

int crb200_peg_max_ell(void) { return crb200::kPegWideMaxEll; }
int crb200_peg_sum_max_ell(void) { return crb200::kPegMaxEll; }

int crb200_peg_precision_fwd(int dtype, int ell, const crb200_peg_fwd_args* a, void* stream) {
  if (a == nullptr) return CRB200_EINVAL;
  if ((dtype != CRB200_F32 && dtype != CRB200_F64) || ell < 1 || ell > crb200::kPegWideMaxEll) return CRB200_EUNSUPPORTED;
  if (a->batch < 0 || a->n < 1 || a->R == nullptr || a->lam_re == nullptr || a->lam_im == nullptr || a->M_re == nullptr || a->M_im == nullptr)
    return CRB200_EINVAL;
  if (a->n > 1 && (a->O == nullptr || a->gaps == nullptr)) return CRB200_EINVAL;
  if (a->nterms < 0 || a->nterms > ell) return CRB200_EINVAL;
  if (a->batch == 0) return CRB200_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (ell > crb200::kPegMaxEll) return finish(wide_fwd(dtype, ell, *a, s));
  return finish(dtype == CRB200_F32 ? PegDispatch<float, 1>::fwd(ell, *a, s) : PegDispatch<double, 1>::fwd(ell, *a, s));
}

int crb200_peg_precision_bwd(int dtype, int ell, const crb200_peg_bwd_args* a, void* stream) {
  if (a == nullptr) return CRB200_EINVAL;
  if ((dtype != CRB200_F32 && dtype != CRB200_F64) || ell < 1 || ell > crb200::kPegWideMaxEll) return CRB200_EUNSUPPORTED;
  if (a->batch < 0 || a->n < 1 || (ell <= crb200::kPegMaxEll ? a->S == nullptr : a->gA == nullptr) || a->lam_re == nullptr || a->lam_im == nullptr || a->M_re == nullptr || a->M_im == nullptr)
    return CRB200_EINVAL;
  if (a->nterms < 0 || a->nterms > ell) return CRB200_EINVAL;
  if (a->n > 1 && (a->O == nullptr || a->gaps == nullptr || a->gR == nullptr || a->gO == nullptr)) return CRB200_EINVAL;
  if (a->batch == 0 || a->n < 2) return CRB200_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (ell > crb200::kPegMaxEll) return finish(wide_bwd(dtype, ell, *a, s));
  return finish(dtype == CRB200_F32 ? PegDispatch<float, 1>::bwd(ell, *a, s) : PegDispatch<double, 1>::bwd(ell, *a, s));
}

}  // extern "C"
