// Instantiations + C entry points of the precision-block builder (cr_peg.cuh, include/crb200.h).
#include "cr_peg.cuh"

namespace {
thread_local int g_peg_last_error = 0;

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

int crb200_peg_max_ell(void) { return crb200::kPegMaxEll; }

int crb200_peg_precision_fwd(int dtype, int ell, const crb200_peg_fwd_args* a, void* stream) {
  if (a == nullptr) return CRB200_EINVAL;
  if ((dtype != CRB200_F32 && dtype != CRB200_F64) || ell < 1 || ell > crb200::kPegMaxEll) return CRB200_EUNSUPPORTED;
  if (a->batch < 0 || a->n < 1 || a->R == nullptr || a->lam_re == nullptr || a->lam_im == nullptr || a->M_re == nullptr || a->M_im == nullptr)
    return CRB200_EINVAL;
  if (a->n > 1 && (a->O == nullptr || a->gaps == nullptr)) return CRB200_EINVAL;
  if (a->nterms < 0 || a->nterms > ell) return CRB200_EINVAL;
  if (a->batch == 0) return CRB200_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  return finish(dtype == CRB200_F32 ? PegDispatch<float, 1>::fwd(ell, *a, s) : PegDispatch<double, 1>::fwd(ell, *a, s));
}

int crb200_peg_precision_bwd(int dtype, int ell, const crb200_peg_bwd_args* a, void* stream) {
  if (a == nullptr) return CRB200_EINVAL;
  if ((dtype != CRB200_F32 && dtype != CRB200_F64) || ell < 1 || ell > crb200::kPegMaxEll) return CRB200_EUNSUPPORTED;
  if (a->batch < 0 || a->n < 1 || a->S == nullptr || a->lam_re == nullptr || a->lam_im == nullptr || a->M_re == nullptr || a->M_im == nullptr)
    return CRB200_EINVAL;
  if (a->nterms < 0 || a->nterms > ell) return CRB200_EINVAL;
  if (a->n > 1 && (a->O == nullptr || a->gaps == nullptr || a->gR == nullptr || a->gO == nullptr)) return CRB200_EINVAL;
  if (a->batch == 0 || a->n < 2) return CRB200_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  return finish(dtype == CRB200_F32 ? PegDispatch<float, 1>::bwd(ell, *a, s) : PegDispatch<double, 1>::bwd(ell, *a, s));
}

}  // extern "C"
