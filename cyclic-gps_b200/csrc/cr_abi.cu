// extern "C" entry points of libcrb200 (include/crb200.h): argument validation + dispatch to
// the per-(dtype, ell) template instantiations.  No allocation, no global mutable state
// except the thread-local "last CUDA error" slot.
#include <cuda_runtime.h>
#include <atomic>
#include "crb200.h"
#include "cr_multi.h"
#include <cstdlib>
#include <cstdint>

namespace crb200 {
#define CRB_DECL(TN, LO, HI)                                                                         \
  cudaError_t inst_fwd_##TN##_##LO##_##HI(int, const crb200_fwd_args&, cudaStream_t);                \
  cudaError_t inst_bwd_##TN##_##LO##_##HI(int, const crb200_bwd_args&, cudaStream_t);                \
  cudaError_t inst_hs_##TN##_##LO##_##HI(int, const crb200_hs_args&, cudaStream_t);                  \
  int inst_fwd_tile_##TN##_##LO##_##HI(int);                                                         \
  int inst_bwd_tile_##TN##_##LO##_##HI(int);                                                         \
  cudaError_t inst_fwd_multi_##TN##_##LO##_##HI(int, const MultiArgs<crb200_fwd_args>*, cudaStream_t); \
  cudaError_t inst_bwd_multi_##TN##_##LO##_##HI(int, const MultiArgs<crb200_bwd_args>*, cudaStream_t);
#define CRB_RANGES(X, TN) X(TN, 1, 4) X(TN, 5, 8) X(TN, 9, 12) X(TN, 13, 16) X(TN, 17, 20) X(TN, 21, 24) X(TN, 25, 28) X(TN, 29, 32)
CRB_RANGES(CRB_DECL, f32)
CRB_RANGES(CRB_DECL, f64)
}  // namespace crb200

namespace {
thread_local int g_last_cuda_error = 0;
std::atomic<long long> g_launches{0};     // diagnostic only (crb200_launch_count): kernels launched by this library

int finish(cudaError_t e) {
  if (e == cudaSuccess) return CRB200_OK;
  g_last_cuda_error = (int)e;
  return e == cudaErrorInvalidValue ? CRB200_EINVAL : CRB200_ECUDA;
}

#define CRB_CASE(KIND, TN, LO, HI) \
  if (ell >= LO && ell <= HI) { g_launches.fetch_add(1, std::memory_order_relaxed); return finish(crb200::inst_##KIND##_##TN##_##LO##_##HI(ell, *a, s)); }
#define CRB_FWD_MULTI(TN, LO, HI) if (ell >= LO && ell <= HI) return crb200::inst_fwd_multi_##TN##_##LO##_##HI(ell, ma, s);
#define CRB_BWD_MULTI(TN, LO, HI) if (ell >= LO && ell <= HI) return crb200::inst_bwd_multi_##TN##_##LO##_##HI(ell, ma, s);
#define CRB_FWD(TN, LO, HI) CRB_CASE(fwd, TN, LO, HI)
#define CRB_BWD(TN, LO, HI) CRB_CASE(bwd, TN, LO, HI)
#define CRB_HS(TN, LO, HI) CRB_CASE(hs, TN, LO, HI)
#define CRB_TILE_F(TN, LO, HI) if (ell >= LO && ell <= HI) return crb200::inst_fwd_tile_##TN##_##LO##_##HI(ell);
#define CRB_TILE_B(TN, LO, HI) if (ell >= LO && ell <= HI) return crb200::inst_bwd_tile_##TN##_##LO##_##HI(ell);

inline char* adv(void* p, long long elems, int es) { return p ? static_cast<char*>(p) + elems * es : nullptr; }
inline const char* adv(const void* p, long long elems, int es) { return p ? static_cast<const char*>(p) + elems * es : nullptr; }

bool bad_common(int dtype, int ell) { return (dtype != CRB200_F32 && dtype != CRB200_F64) || ell < 1 || ell > 32; }

// packed lower triangles: only the thread-per-node family implements them (automatic choice for every size that offers them)
int tri_stride(int dtype, int ell) { return crb200::tri_stride_elems(dtype == CRB200_F32 ? 4 : 8, ell); }
bool bad_tri(int dtype, int ell, int tri, int variant, bool halo) {
  if (tri == 0) return false;
  return tri_stride(dtype, ell) == 0 || halo || (tri & ~3) != 0 || (variant != CRB200_AUTO && variant != CRB200_THREAD_PER_NODE);
}
bool misaligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) != 0; }

// fused deep-level kernels (thread-per-node family); ma == nullptr only asks whether they exist for (dtype, ell)
cudaError_t fwd_multi(int dtype, int ell, const crb200::MultiArgs<crb200_fwd_args>* ma, cudaStream_t s) {
  if (dtype == CRB200_F32) { CRB_RANGES(CRB_FWD_MULTI, f32) } else { CRB_RANGES(CRB_FWD_MULTI, f64) }
  return cudaErrorNotSupported;
}
cudaError_t bwd_multi(int dtype, int ell, const crb200::MultiArgs<crb200_bwd_args>* ma, cudaStream_t s) {
  if (dtype == CRB200_F32) { CRB_RANGES(CRB_BWD_MULTI, f32) } else { CRB_RANGES(CRB_BWD_MULTI, f64) }
  return cudaErrorNotSupported;
}

// The fused kernel runs one CTA per series and keeps it for all deep levels, so it only pays while every series has
// its CTA resident at once (two CTAs per SM at the largest record): small batches -- single series, sub-chunk batches
// of the chunked path, boundary systems.  Measured on configs[1] (1024 series): 9.9 ms fused vs 9.3 ms per-level.
bool batch_fits_one_wave(int batch) {
  static std::atomic<int> sms[64];                            // SM count per device, 0 = not asked yet (a race only repeats the query)
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) return false;
  int v = dev < 64 ? sms[dev].load(std::memory_order_relaxed) : 0;
  if (v == 0) {
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) return false;
    if (dev < 64) sms[dev].store(v, std::memory_order_relaxed);
  }
  return batch <= 2 * v;
}

// Large batches: fuse the single-tile levels into one launch of single-warp CTAs (CRB200_FUSE_LARGE=0 keeps a launch per level;
// read once).
bool fuse_large_batches() {
  static const bool on = [] { const char* e = getenv("CRB200_FUSE_LARGE"); return e == nullptr || e[0] != '0'; }();
  return on;
}

// number of CR levels of an n-row system: floor(log2 n) + 1
int max_levels(int n) {
  int L = 0;
  for (int m = n; m >= 1; m /= 2) ++L;
  return L;
}

// First level of the fused tail of a sweep over levels 0..L-1 of an n-row system, or L when nothing is fused.
// A level joins the tail when a series has at most `max_tiles` tiles of `tile` even nodes (2 * kMultiWarps for the four-warp CTAs
// of small batches, 1 for the single-warp CTAs of large batches); the tail starts at
// level 3 or deeper (the ping-pong scratch of the caller then has room for every fused level at its own offset, so
// CTAs of different series never touch the same bytes), holds at least two and at most kMultiMax levels.
int fuse_start(int n, int L, int tile, bool odd_parity_needed, int parity, int max_tiles) {
  if (tile <= 0 || L < 5) return L;
  int ks = L, m = n;
  for (int k = 0; k < L; ++k, m /= 2) {
    const int E = (m + 1) / 2;
    if (k >= 3 && (E + tile - 1) / tile <= max_tiles) { ks = k; break; }
  }
  if (L - ks > crb200::kMultiMax) ks = L - crb200::kMultiMax;
  if (odd_parity_needed && ((ks & 1) != parity)) ++ks;
  return (L - ks >= 2) ? ks : L;
}
}  // namespace

extern "C" {

int crb200_version(void) { return 101; }
int crb200_max_ell(void) { return 32; }
int crb200_last_cuda_error(void) { return g_last_cuda_error; }
long long crb200_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

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
  if (bad_tri(dtype, ell, a->tri, a->variant, a->O_halo != nullptr)) return CRB200_EINVAL;
  if (((a->tri & 1) != 0 && misaligned16(a->R)) || ((a->tri & 2) != 0 && (misaligned16(a->D) || misaligned16(a->Rn)))) return CRB200_EINVAL;
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
  if (bad_tri(dtype, ell, a->tri, a->variant, a->G_halo != nullptr) || ((a->tri & 2) != 0 && a->grad_mode != 0)) return CRB200_EINVAL;
  if (((a->tri & 1) != 0 && (misaligned16(a->D) || misaligned16(a->Sd_in))) || ((a->tri & 2) != 0 && misaligned16(a->Sd_out))) return CRB200_EINVAL;
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
  if (a->nlevels > max_levels(a->n)) return CRB200_EINVAL;   // checked before anything is launched
  if (a->tri != 0 && (bad_tri(dtype, ell, 3, a->variant, a->O_halo != nullptr) || a->nlevels != max_levels(a->n))) return CRB200_EINVAL;
  const int es = dtype == CRB200_F32 ? 4 : 8;
  const long long bs = (long long)ell * ell, B = a->batch;
  const long long dbs = a->tri != 0 ? tri_stride(dtype, ell) : bs;      // elements per diagonal block of the reduced systems
  crb200_fwd_args l{};
  l.batch = a->batch;
  l.R = a->R; l.O = a->O; l.y = a->y;
  l.strideR = a->strideR; l.strideO = a->strideO; l.stridey = a->stridey;
  l.logdet = a->logdet; l.mahal = a->mahal; l.acc_slots = a->acc_slots;
  l.Rh_acc = a->Rh_acc; l.yh_acc = a->yh_acc; l.variant = a->variant;
  const void* halo = a->O_halo;
  long long offE = 0, offO = 0, offG = 0;
  int m = a->n;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // fused tail (automatic kernel choice only): levels ks.. run in ONE launch, one CTA per series.  Each fused level
  // writes its reduced system at its own offset of the ping-pong scratch: [0, B*m_ks) rows of slot (ks-1)&1 hold the
  // live input of level ks, and when the sweep stops early the result of the last level must sit at the base of
  // slot (L-1)&1 (that is where callers read it), which therefore has to be the OTHER slot: ks and L differ in parity.
  int ks = a->nlevels, multi_warps = crb200::kMultiWarps;
  {
    const int L = a->nlevels;
    int mlast = a->n;
    for (int k = 0; k + 1 < L; ++k) mlast /= 2;
    const bool early = mlast / 2 > 0;                       // a system is left below level L-1
    if (a->variant == CRB200_AUTO && a->batch > 0 && (batch_fits_one_wave(a->batch) || fuse_large_batches()) &&
        fwd_multi(dtype, ell, nullptr, s) == cudaSuccess) {
      multi_warps = batch_fits_one_wave(a->batch) ? crb200::kMultiWarps : 1;
      ks = fuse_start(a->n, L, crb200_fwd_tile_nodes(dtype, ell), early, (L & 1) ^ 1, multi_warps == 1 ? 1 : 2 * crb200::kMultiWarps);
    }
  }
  crb200::MultiArgs<crb200_fwd_args> multi{};
  multi.warps = multi_warps;
  long long cum[2] = {0, 0};                                // rows already claimed in each scratch slot (fused levels)
  for (int k = 0; k < a->nlevels; ++k) {
    if (m < 1) return CRB200_EINVAL;
    const long long E = (m + 1) / 2, o = m / 2, g = (m - 1) / 2;
    const int slot = k & 1;
    l.m = m;
    l.tri = a->tri != 0 ? ((k > 0 ? 1 : 0) | 2) : 0;         // level 0 reads the caller's full blocks
    l.D = adv(a->D, B * bs * offE, es);
    l.F = o > 0 ? adv(a->F, B * bs * offO, es) : nullptr;
    l.G = g > 0 ? adv(a->G, B * bs * offG, es) : nullptr;
    l.xk = adv(a->X, B * ell * offE, es);
    long long off = 0;                                      // row offset of this level's output inside its slot
    if (k >= ks) {
      if (k == ks) {
        cum[(ks - 1) & 1] = B * m;                          // live input of the first fused level
        const int sl = (a->nlevels - 1) & 1;
        int ml = m;
        for (int j = k; j + 1 < a->nlevels; ++j) ml /= 2;
        if (ml / 2 > 0) cum[sl] = B * (ml / 2);             // base of that slot is reserved for the last level's result
      }
      if (k == a->nlevels - 1) off = 0;
      else { off = cum[slot]; cum[slot] += B * o; }
    }
    l.Rn = o > 0 ? adv(a->scrR[slot], off * bs, es) : nullptr;
    l.On = o > 1 ? adv(a->scrO[slot], off * bs, es) : nullptr;
    l.yn = (o > 0 && a->y != nullptr) ? adv(a->scry[slot], off * ell, es) : nullptr;
    l.info = a->info ? a->info + k : nullptr;
    if (halo != nullptr) {
      l.O_halo = halo;
      l.G_halo = adv(a->G_halo, (long long)k * B * bs, es);
      l.On_halo = a->On_halo[slot];
    }
    if (k < ks) {
      const int rc = crb200_level_fwd(dtype, ell, &l, stream);
      if (rc != CRB200_OK) return rc;
    } else {
      multi.lv[multi.count++] = l;
    }
    if (halo != nullptr) halo = a->On_halo[slot];
    l.R = l.Rn; l.O = l.On; l.y = l.yn;
    l.strideR = o * dbs; l.strideO = (o > 1 ? o - 1 : 0) * bs; l.stridey = o * ell;
    offE += E; offO += o; offG += g;
    m = (int)o;
  }
  if (multi.count > 0) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    const int rc = finish(fwd_multi(dtype, ell, &multi, s));
    if (rc != CRB200_OK) return rc;
  }
  return CRB200_OK;
}

int crb200_sweep_halfsolve(int dtype, int ell, const crb200_sweep_hs_args* a, void* stream) {
  if (a == nullptr) return CRB200_EINVAL;
  if (bad_common(dtype, ell)) return CRB200_EUNSUPPORTED;
  if (a->batch < 0 || a->n < 1 || a->nlevels < 1 || a->D == nullptr || a->y == nullptr || a->X == nullptr) return CRB200_EINVAL;
  if (a->nlevels > max_levels(a->n)) return CRB200_EINVAL;
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
  if (a->nlevels > max_levels(a->n)) return CRB200_EINVAL;
  if (a->tri != 0 && (bad_tri(dtype, ell, 3, a->variant, a->G_halo != nullptr) || a->nlevels != max_levels(a->n) || a->top_Sd != nullptr))
    return CRB200_EINVAL;
  const int es = dtype == CRB200_F32 ? 4 : 8;
  const long long bs = (long long)ell * ell, B = a->batch;
  const long long dbs = a->tri != 0 ? tri_stride(dtype, ell) : bs;
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
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // fused tail: the deepest levels (L-1 .. ks) run in ONE launch, each writing its (Sigma_d, Sigma_o, w) at its own
  // offset of the ping-pong scratch (levels >= 3 together need less than a third of a slot)
  int ks = a->nlevels, multi_warps = crb200::kMultiWarps;
  if (a->variant == CRB200_AUTO && a->batch > 0 && (batch_fits_one_wave(a->batch) || fuse_large_batches()) &&
      bwd_multi(dtype, ell, nullptr, s) == cudaSuccess) {
    multi_warps = batch_fits_one_wave(a->batch) ? crb200::kMultiWarps : 1;
    ks = fuse_start(a->n, a->nlevels, crb200_bwd_tile_nodes(dtype, ell), false, 0, multi_warps == 1 ? 1 : 2 * crb200::kMultiWarps);
  }
  crb200::MultiArgs<crb200_bwd_args> multi{};
  multi.warps = multi_warps;
  long long cum[2] = {0, 0};
  for (int k = a->nlevels - 1; k >= 0; --k) {
    const int m = ms[k];
    const long long o = m / 2, g = (m - 1) / 2;
    const int slot = k & 1;
    l.m = m;
    l.tri = a->tri != 0 ? (1 | (k > 0 ? 2 : 0)) : 0;         // level 0 writes the caller's full blocks
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
      long long off = 0;
      if (k >= ks) { off = cum[slot]; cum[slot] += B * m; }
      l.Sd_out = sig ? adv(a->scrSd[slot], off * bs, es) : nullptr;
      l.So_out = (sig && m > 1) ? adv(a->scrSo[slot], off * bs, es) : nullptr;
      l.w_out = w ? adv(a->scrw[slot], off * ell, es) : nullptr;
      l.strideSd = (long long)m * dbs; l.strideSo = (long long)(m > 1 ? m - 1 : 0) * bs; l.stridew = (long long)m * ell;
      l.gm = nullptr; l.gd = nullptr; l.grad_mode = 0;
    }
    if (a->G_halo != nullptr) {
      l.G_halo = adv(a->G_halo, (long long)k * B * bs, es);
      l.Sd_halo = sig ? a->Sd_halo : nullptr;
      l.w_halo = w ? a->w_halo : nullptr;
      l.So_halo_in = (sig && o > 0) ? so_h : nullptr;
      l.So_halo_out = sig ? (k == 0 ? a->So_halo_out : a->So_halo[slot]) : nullptr;
    }
    if (k >= ks) {
      multi.lv[multi.count++] = l;
      if (k == ks) {
        g_launches.fetch_add(1, std::memory_order_relaxed);
        const int rc = finish(bwd_multi(dtype, ell, &multi, s));
        if (rc != CRB200_OK) return rc;
      }
    } else {
      const int rc = crb200_level_bwd(dtype, ell, &l, stream);
      if (rc != CRB200_OK) return rc;
    }
    if (a->G_halo != nullptr && sig) so_h = l.So_halo_out;
    l.Sd_in = l.Sd_out; l.So_in = l.So_out; l.w_in = l.w_out;
  }
  return CRB200_OK;
}

int crb200_tri_stride(int dtype, int ell) { return bad_common(dtype, ell) ? 0 : tri_stride(dtype, ell); }

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
