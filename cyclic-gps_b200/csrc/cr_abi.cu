// extern "C" entry points of libcrb200 (include/crb200.h): argument validation + dispatch to
// the per-(dtype, ell) template instantiations.  No allocation, no global mutable state
// except the thread-local "last CUDA error" slot.
#include <cuda_runtime.h>
#include "crb200.h"

namespace crb200 {
#define CRB_DECL(TN, LO, HI)                                                                         \
  cudaError_t inst_fwd_##TN##_##LO##_##HI(int, const crb200_fwd_args&, cudaStream_t);                \
  cudaError_t inst_bwd_##TN##_##LO##_##HI(int, const crb200_bwd_args&, cudaStream_t);                \
  cudaError_t inst_hs_##TN##_##LO##_##HI(int, const crb200_hs_args&, cudaStream_t);                  \
  int inst_fwd_tile_##TN##_##LO##_##HI(int);                                                         \
  int inst_bwd_tile_##TN##_##LO##_##HI(int);
#define CRB_RANGES(X, TN) X(TN, 1, 4) X(TN, 5, 8) X(TN, 9, 12) X(TN, 13, 16) X(TN, 17, 20) X(TN, 21, 24) X(TN, 25, 28) X(TN, 29, 32)
CRB_RANGES(CRB_DECL, f32)
CRB_RANGES(CRB_DECL, f64)
}  // namespace crb200

namespace {
thread_local int g_last_cuda_error = 0;

int finish(cudaError_t e) {
  if (e == cudaSuccess) return CRB200_OK;
  g_last_cuda_error = (int)e;
  return e == cudaErrorInvalidValue ? CRB200_EINVAL : CRB200_ECUDA;
}

#define CRB_CASE(KIND, TN, LO, HI) \
  if (ell >= LO && ell <= HI) return finish(crb200::inst_##KIND##_##TN##_##LO##_##HI(ell, *a, s));
#define CRB_FWD(TN, LO, HI) CRB_CASE(fwd, TN, LO, HI)
#define CRB_BWD(TN, LO, HI) CRB_CASE(bwd, TN, LO, HI)
#define CRB_HS(TN, LO, HI) CRB_CASE(hs, TN, LO, HI)
#define CRB_TILE_F(TN, LO, HI) if (ell >= LO && ell <= HI) return crb200::inst_fwd_tile_##TN##_##LO##_##HI(ell);
#define CRB_TILE_B(TN, LO, HI) if (ell >= LO && ell <= HI) return crb200::inst_bwd_tile_##TN##_##LO##_##HI(ell);

bool bad_common(int dtype, int ell) { return (dtype != CRB200_F32 && dtype != CRB200_F64) || ell < 1 || ell > 32; }
}  // namespace

extern "C" {

int crb200_version(void) { return 100; }
int crb200_max_ell(void) { return 32; }
int crb200_last_cuda_error(void) { return g_last_cuda_error; }

int crb200_level_fwd(int dtype, int ell, const crb200_fwd_args* a, void* stream) {
  if (a == nullptr) return CRB200_EINVAL;
  if (bad_common(dtype, ell)) return CRB200_EUNSUPPORTED;
  if (a->batch < 0 || a->m < 1 || a->R == nullptr) return CRB200_EINVAL;
  if (a->m > 1 && a->O == nullptr) return CRB200_EINVAL;
  const bool keep = a->D != nullptr;
  if (keep && ((a->m > 1 && a->F == nullptr) || (a->m > 2 && a->G == nullptr))) return CRB200_EINVAL;
  if (a->m > 1 && a->Rn == nullptr) return CRB200_EINVAL;
  if (a->m > 3 && a->On == nullptr) return CRB200_EINVAL;
  if (a->y != nullptr && a->m > 1 && a->yn == nullptr) return CRB200_EINVAL;
  if (a->O_halo != nullptr && (a->Rh_acc == nullptr || a->On_halo == nullptr || (keep && a->G_halo == nullptr))) return CRB200_EINVAL;
  if (a->batch == 0) return CRB200_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dtype == CRB200_F32) { CRB_RANGES(CRB_FWD, f32) } else { CRB_RANGES(CRB_FWD, f64) }
  return CRB200_EUNSUPPORTED;
}

int crb200_level_bwd(int dtype, int ell, const crb200_bwd_args* a, void* stream) {
  if (a == nullptr) return CRB200_EINVAL;
  if (bad_common(dtype, ell)) return CRB200_EUNSUPPORTED;
  if (a->batch < 0 || a->m < 1 || a->D == nullptr) return CRB200_EINVAL;
  if ((a->m > 1 && a->F == nullptr) || (a->m > 2 && a->G == nullptr)) return CRB200_EINVAL;
  const bool sig = a->Sd_out != nullptr, w = a->w_out != nullptr;
  if (!sig && !w) return CRB200_EINVAL;
  if (sig && a->m > 1 && (a->So_out == nullptr || a->Sd_in == nullptr)) return CRB200_EINVAL;
  if (sig && a->m > 3 && a->So_in == nullptr) return CRB200_EINVAL;
  if (w && (a->xk == nullptr || (a->m > 1 && a->w_in == nullptr))) return CRB200_EINVAL;
  if (a->G_halo != nullptr) {
    if (sig && (a->Sd_halo == nullptr || a->So_halo_out == nullptr || (a->m > 1 && a->So_halo_in == nullptr))) return CRB200_EINVAL;
    if (w && a->w_halo == nullptr) return CRB200_EINVAL;
  }
  if (a->batch == 0) return CRB200_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dtype == CRB200_F32) { CRB_RANGES(CRB_BWD, f32) } else { CRB_RANGES(CRB_BWD, f64) }
  return CRB200_EUNSUPPORTED;
}

int crb200_level_halfsolve(int dtype, int ell, const crb200_hs_args* a, void* stream) {
  if (a == nullptr) return CRB200_EINVAL;
  if (bad_common(dtype, ell)) return CRB200_EUNSUPPORTED;
  if (a->batch < 0 || a->m < 1 || a->D == nullptr || a->y == nullptr || a->xk == nullptr) return CRB200_EINVAL;
  if (a->m > 1 && (a->F == nullptr || a->yn == nullptr)) return CRB200_EINVAL;
  if (a->m > 2 && a->G == nullptr) return CRB200_EINVAL;
  if (a->batch == 0) return CRB200_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dtype == CRB200_F32) { CRB_RANGES(CRB_HS, f32) } else { CRB_RANGES(CRB_HS, f64) }
  return CRB200_EUNSUPPORTED;
}

int crb200_fwd_tile_nodes(int dtype, int ell) {
  if (bad_common(dtype, ell)) return 0;
  if (dtype == CRB200_F32) { CRB_RANGES(CRB_TILE_F, f32) } else { CRB_RANGES(CRB_TILE_F, f64) }
  return 0;
}
int crb200_bwd_tile_nodes(int dtype, int ell) {
  if (bad_common(dtype, ell)) return 0;
  if (dtype == CRB200_F32) { CRB_RANGES(CRB_TILE_B, f32) } else { CRB_RANGES(CRB_TILE_B, f64) }
  return 0;
}

}  // extern "C"
