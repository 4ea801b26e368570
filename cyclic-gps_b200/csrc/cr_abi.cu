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

inline char* adv(void* p, long long elems, int es) { return p ? static_cast<char*>(p) + elems * es : nullptr; }
inline const char* adv(const void* p, long long elems, int es) { return p ? static_cast<const char*>(p) + elems * es : nullptr; }

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

// ---- whole sweeps: the level loop in C --------------------------------------------------------
int crb200_sweep_fwd(int dtype, int ell, const crb200_sweep_fwd_args* a, void* stream) {
  if (a == nullptr) return CRB200_EINVAL;
  if (bad_common(dtype, ell)) return CRB200_EUNSUPPORTED;
  if (a->batch < 0 || a->n < 1 || a->nlevels < 1 || a->R == nullptr) return CRB200_EINVAL;
  const int es = dtype == CRB200_F32 ? 4 : 8;
  const long long bs = (long long)ell * ell, B = a->batch;
  crb200_fwd_args l{};
  l.batch = a->batch;
  l.R = a->R; l.O = a->O; l.y = a->y;
  l.strideR = a->strideR; l.strideO = a->strideO; l.stridey = a->stridey;
  l.logdet = a->logdet; l.mahal = a->mahal; l.acc_slots = a->acc_slots;
  l.Rh_acc = a->Rh_acc; l.yh_acc = a->yh_acc; l.variant = a->variant;
  const void* halo = a->O_halo;
  long long offE = 0, offO = 0, offG = 0;
  int m = a->n;
  for (int k = 0; k < a->nlevels; ++k) {
    if (m < 1) return CRB200_EINVAL;
    const long long E = (m + 1) / 2, o = m / 2, g = (m - 1) / 2;
    const int slot = k & 1;
    l.m = m;
    l.D = adv(a->D, B * bs * offE, es);
    l.F = o > 0 ? adv(a->F, B * bs * offO, es) : nullptr;
    l.G = g > 0 ? adv(a->G, B * bs * offG, es) : nullptr;
    l.xk = adv(a->X, B * ell * offE, es);
    l.Rn = o > 0 ? a->scrR[slot] : nullptr;
    l.On = o > 1 ? a->scrO[slot] : nullptr;
    l.yn = (o > 0 && a->y != nullptr) ? a->scry[slot] : nullptr;
    l.info = a->info ? a->info + k : nullptr;
    if (halo != nullptr) {
      l.O_halo = halo;
      l.G_halo = adv(a->G_halo, (long long)k * B * bs, es);
      l.On_halo = a->On_halo[slot];
    }
    const int rc = crb200_level_fwd(dtype, ell, &l, stream);
    if (rc != CRB200_OK) return rc;
    if (halo != nullptr) halo = a->On_halo[slot];
    l.R = l.Rn; l.O = l.On; l.y = l.yn;
    l.strideR = o * bs; l.strideO = (o > 1 ? o - 1 : 0) * bs; l.stridey = o * ell;
    offE += E; offO += o; offG += g;
    m = (int)o;
  }
  return CRB200_OK;
}

int crb200_sweep_halfsolve(int dtype, int ell, const crb200_sweep_hs_args* a, void* stream) {
  if (a == nullptr) return CRB200_EINVAL;
  if (bad_common(dtype, ell)) return CRB200_EUNSUPPORTED;
  if (a->batch < 0 || a->n < 1 || a->nlevels < 1 || a->D == nullptr || a->y == nullptr || a->X == nullptr) return CRB200_EINVAL;
  const int es = dtype == CRB200_F32 ? 4 : 8;
  const long long bs = (long long)ell * ell, B = a->batch;
  crb200_hs_args l{};
  l.batch = a->batch; l.mahal = a->mahal;
  l.y = a->y; l.stridey = a->stridey;
  long long offE = 0, offO = 0, offG = 0;
  int m = a->n;
  for (int k = 0; k < a->nlevels; ++k) {
    if (m < 1) return CRB200_EINVAL;
    const long long E = (m + 1) / 2, o = m / 2, g = (m - 1) / 2;
    l.m = m;
    l.D = adv(a->D, B * bs * offE, es);
    l.F = o > 0 ? adv(a->F, B * bs * offO, es) : nullptr;
    l.G = g > 0 ? adv(a->G, B * bs * offG, es) : nullptr;
    l.xk = adv(a->X, B * ell * offE, es);
    l.yn = o > 0 ? a->scry[k & 1] : nullptr;
    const int rc = crb200_level_halfsolve(dtype, ell, &l, stream);
    if (rc != CRB200_OK) return rc;
    l.y = l.yn; l.stridey = o * ell;
    offE += E; offO += o; offG += g;
    m = (int)o;
  }
  return CRB200_OK;
}

int crb200_sweep_bwd(int dtype, int ell, const crb200_sweep_bwd_args* a, void* stream) {
  if (a == nullptr) return CRB200_EINVAL;
  if (bad_common(dtype, ell)) return CRB200_EUNSUPPORTED;
  if (a->batch < 0 || a->n < 1 || a->nlevels < 1 || a->nlevels > 40 || a->D == nullptr) return CRB200_EINVAL;
  const int es = dtype == CRB200_F32 ? 4 : 8;
  const long long bs = (long long)ell * ell, B = a->batch;
  int ms[40];
  long long offE[40], offO[40], offG[40];
  {
    long long e = 0, o = 0, g = 0;
    int m = a->n;
    for (int k = 0; k < a->nlevels; ++k) {
      if (m < 1) return CRB200_EINVAL;
      ms[k] = m; offE[k] = e; offO[k] = o; offG[k] = g;
      e += (m + 1) / 2; o += m / 2; g += (m - 1) / 2;
      m /= 2;
    }
  }
  const bool sig = a->Sd_out != nullptr, w = a->w_out != nullptr;
  crb200_bwd_args l{};
  l.batch = a->batch; l.variant = a->variant;
  l.Sd_in = a->top_Sd; l.So_in = a->top_So; l.w_in = a->top_w;
  const void* so_h = a->So_halo_in;
  for (int k = a->nlevels - 1; k >= 0; --k) {
    const int m = ms[k];
    const long long o = m / 2, g = (m - 1) / 2;
    const int slot = k & 1;
    l.m = m;
    l.D = adv(a->D, B * bs * offE[k], es);
    l.F = o > 0 ? adv(a->F, B * bs * offO[k], es) : nullptr;
    l.G = g > 0 ? adv(a->G, B * bs * offG[k], es) : nullptr;
    l.xk = w ? adv(a->X, B * ell * offE[k], es) : nullptr;
    if (o <= 1) l.So_in = nullptr;
    if (k == 0) {
      l.Sd_out = a->Sd_out; l.So_out = m > 1 ? a->So_out : nullptr; l.w_out = a->w_out;
      l.strideSd = a->strideSd; l.strideSo = a->strideSo; l.stridew = a->stridew;
      l.gm = a->gm; l.gd = a->gd; l.grad_mode = a->grad_mode;
    } else {
      l.Sd_out = sig ? a->scrSd[slot] : nullptr;
      l.So_out = (sig && m > 1) ? a->scrSo[slot] : nullptr;
      l.w_out = w ? a->scrw[slot] : nullptr;
      l.strideSd = (long long)m * bs; l.strideSo = (long long)(m > 1 ? m - 1 : 0) * bs; l.stridew = (long long)m * ell;
      l.gm = nullptr; l.gd = nullptr; l.grad_mode = 0;
    }
    if (a->G_halo != nullptr) {
      l.G_halo = adv(a->G_halo, (long long)k * B * bs, es);
      l.Sd_halo = sig ? a->Sd_halo : nullptr;
      l.w_halo = w ? a->w_halo : nullptr;
      l.So_halo_in = (sig && o > 0) ? so_h : nullptr;
      l.So_halo_out = sig ? (k == 0 ? a->So_halo_out : a->So_halo[slot]) : nullptr;
    }
    const int rc = crb200_level_bwd(dtype, ell, &l, stream);
    if (rc != CRB200_OK) return rc;
    if (a->G_halo != nullptr && sig) so_h = l.So_halo_out;
    l.Sd_in = l.Sd_out; l.So_in = l.So_out; l.w_in = l.w_out;
  }
  return CRB200_OK;
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
