// Row-split forward level kernel: LPN lanes per even node, each lane owns RW = L / LPN rows of F, G
// and of every Schur product.
//
// Same contract and record layout as cr_tpn_fwd_kernel (cr_tpn_fwd.cuh; reference
// cyclic_gps/cyclic_reduction.py:204-259, :412-427).  The Cholesky factor K (36 elements at L = 8) and the
// half solve x are computed redundantly by every lane of a node; everything else is separable by rows:
//   F[r,:]  = O_right[r,:] K^{-T}            G[r,:] = O_left[:,r]^T K^{-T}
//   (F F^T)[r,:], (G G^T)[r,:], (-F G^T)[r,:] need the OTHER rows only as right-hand operands, which are read
//   back from the shared-memory record after a __syncwarp.
// The row offset r0 of a lane only enters shared-memory addresses, never a register index, so all lanes run
// one instruction stream.  With the same ~1.3 KB of shared memory per node in flight, an SM now holds
// LPN times more warps (see cr_cs_bwd.cuh for the motivation).
#pragma once
#include "cr_tpn_fwd.cuh"

namespace crb200 {

// record layout of the row-split kernel: [ R_even | R_odd | O_left | O_right | O~ | y_even | y_odd ]
template <typename T, int L>
struct CsFwdRec {
  static constexpr int BS = L * L;
  static constexpr int RE = 0, RO = BS, OL = 2 * BS, OR_ = 3 * BS, ON = 4 * BS, YE = 5 * BS, YO = 5 * BS + L;
  static constexpr int RAW = 5 * BS + 2 * L;
  static constexpr int NS = record_stride<T>(RAW, BS);
};

template <typename T, int L, int LPN>
struct CsFwdCfg {
  static constexpr int RW = L / LPN;
  static constexpr bool ELIGIBLE = (L % LPN == 0) && (LPN > 1) && (L <= 8) && ((RW * (int)sizeof(T)) % 16 == 0);
  static constexpr int BS = L * L;
  static constexpr int NT = 32 / LPN, OWN = NT - 1;
  using Rec = CsFwdRec<T, L>;
  static constexpr int NS = Rec::NS;
  static constexpr size_t SMEM = (size_t)NT * NS * sizeof(T);
  static constexpr int MIN_CTAS = cmin(CRB200_CS_MAX_CTAS, cmax(1, (int)((226 * 1024) / (SMEM + 1024))));
};

template <typename T, int L, int LPN>
__global__ void __launch_bounds__(32, CsFwdCfg<T, L, LPN>::MIN_CTAS)
cr_cs_fwd_kernel(const LevelFwdArgs a) {
  using Cf = CsFwdCfg<T, L, LPN>;
  using C = typename Cf::Rec;
  constexpr int BS = Cf::BS, NS = Cf::NS, NT = Cf::NT, OWN = Cf::OWN, RW = Cf::RW;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* S = reinterpret_cast<T*>(smem_raw);
  constexpr unsigned ES = sizeof(T);
  const unsigned s0 = smem_u32(S);
  const unsigned nsb = NS * ES;

  const int m = a.m;
  const int E = (m + 1) >> 1, o = m >> 1, gcnt = (m - 1) >> 1;
  const int tiles = (E + OWN - 1) / OWN;
  const int b = blockIdx.x / tiles;
  const int tile = blockIdx.x - b * tiles;
  const int e0 = tile * OWN;
  const bool has_y = a.y != nullptr;
  const bool halo = a.O_halo != nullptr;
  const int lane = threadIdx.x;

  const T* gR = static_cast<const T*>(a.R) + (size_t)b * a.strideR;
  const T* gO = static_cast<const T*>(a.O) + (size_t)b * a.strideO;
  const T* gy = has_y ? static_cast<const T*>(a.y) + (size_t)b * a.stridey : nullptr;

  // ---------------- stage in ----------------
  {
    const int r0g = 2 * e0;
    const int nR = cmin(2 * NT - 1, m - r0g);
    rec_g2s<T, BS, 2>(s0 + C::RE * ES, nsb, gR + (size_t)r0g * BS, 0, nR, is_aligned16(gR));
    const int pfirst = (r0g == 0) ? 1 : 0;
    const int nO = cmin(2 * NT - 1, m - r0g) - pfirst;
    rec_g2s<T, BS, 2>(s0 + C::OL * ES, nsb, gO + (size_t)(r0g - 1 + pfirst) * BS, pfirst, nO, is_aligned16(gO));
    if (r0g == 0 && halo)
      rec_g2s<T, BS, 2>(s0 + C::OL * ES, nsb, static_cast<const T*>(a.O_halo) + (size_t)b * BS, 0, 1, is_aligned16(a.O_halo));
    if (has_y) rec_g2s<T, L, 2>(s0 + C::YE * ES, nsb, gy + (size_t)r0g * L, 0, nR, is_aligned16(gy));
    cp_async_wait_all();
    __syncwarp();
  }

  // ---------------- per-node compute, LPN lanes per node ----------------
  const int t = lane / LPN;
  const int r0 = (lane - t * LPN) * RW;      // first row owned by this lane (addresses only)
  const bool lead = (r0 == 0);
  T* N = S + (size_t)t * NS;
  const int e = e0 + t;
  const bool valid = e < E;
  const bool own = valid && (t < OWN);
  const bool do_f = own && (e < o);
  const bool has_left = valid && (e >= 1 || halo);

  T K[L][L];
  T inv[L];
  T x[L];
  bool bad = false;
  double dprod = 1.0;
#pragma unroll
  for (int r = 0; r < L; ++r) {
    T row[L];
    lds_row<T, L>(row, N + C::RE + r * L);
#pragma unroll
    for (int c = 0; c < L; ++c) K[r][c] = valid ? row[c] : (r == c ? T(1) : T(0));
  }
#pragma unroll
  for (int c = 0; c < L; ++c) x[c] = T(0);
  if (has_y && valid) lds_row<T, L>(x, N + C::YE);
#pragma unroll
  for (int k = 0; k < L; ++k) {
    const T d = K[k][k];
    if (!(d > T(0))) bad = true;
    const T lkk = sqrt(d);
    inv[k] = T(1) / lkk;
    K[k][k] = lkk;
    dprod *= (double)lkk;
#pragma unroll
    for (int r = k + 1; r < L; ++r) K[r][k] *= inv[k];
#pragma unroll
    for (int c = k + 1; c < L; ++c)
#pragma unroll
      for (int r = c; r < L; ++r) K[r][c] = fma(-K[r][k], K[c][k], K[r][c]);
  }
  if (has_y) {
#pragma unroll
    for (int c = 0; c < L; ++c) {
      T s = x[c];
#pragma unroll
      for (int k = 0; k < c; ++k) s = fma(-x[k], K[c][k], s);
      x[c] = s * inv[c];
    }
  }
  __syncwarp();                                     // every lane of the node has read R_even and y_even
  double ld_part = 0.0, mh_part = 0.0;
  if (lead && valid) {
#pragma unroll
    for (int r = 0; r < L; ++r) {
      T row[L];
#pragma unroll
      for (int c = 0; c < L; ++c) row[c] = (c <= r) ? K[r][c] : T(0);
      sts_row<T, L>(N + C::RE + r * L, row);
    }
    if (has_y) sts_row<T, L>(N + C::YE, x);
    if (own) {
      if (a.logdet != nullptr) ld_part = log(dprod);
      if (has_y) {
#pragma unroll
        for (int c = 0; c < L; ++c) mh_part += (double)x[c] * (double)x[c];
      }
      if (bad && a.info != nullptr) {
        const long long flat = (long long)b * E + e;
        atomicMax(a.info, 0x7fffffff - (int)(flat > 0x7ffffffeLL ? 0x7ffffffeLL : flat));
      }
    }
  }

  // my rows of F = O_right K^{-T}  and of  G = O_left^T K^{-T}
  T Fm[RW][L], Gm[RW][L];
  T um[RW], vm[RW];
#pragma unroll
  for (int rr = 0; rr < RW; ++rr) {
    T f[L];
    lds_row<T, L>(f, N + C::OR_ + (r0 + rr) * L);
#pragma unroll
    for (int c = 0; c < L; ++c) {
      T s = do_f ? f[c] : T(0);
#pragma unroll
      for (int k = 0; k < c; ++k) s = fma(-f[k], K[c][k], s);
      f[c] = s * inv[c];
    }
    if (do_f) sts_row<T, L>(N + C::OR_ + (r0 + rr) * L, f);
    T ur = T(0);
#pragma unroll
    for (int c = 0; c < L; ++c) { Fm[rr][c] = f[c]; ur = fma(f[c], x[c], ur); }
    um[rr] = ur;
  }
#pragma unroll
  for (int c = 0; c < L; ++c) {
    T sl[RW];
    lds_row<T, RW>(sl, N + C::OL + c * L + r0);      // O_left[c][r0 .. r0+RW)
#pragma unroll
    for (int rr = 0; rr < RW; ++rr) Gm[rr][c] = has_left ? sl[rr] : T(0);
  }
#pragma unroll
  for (int rr = 0; rr < RW; ++rr) {
    T vr = T(0);
#pragma unroll
    for (int c = 0; c < L; ++c) {
      T s = Gm[rr][c];
#pragma unroll
      for (int k = 0; k < c; ++k) s = fma(-Gm[rr][k], K[c][k], s);
      Gm[rr][c] = s * inv[c];
      vr = fma(Gm[rr][c], x[c], vr);
    }
    vm[rr] = vr;
  }
  __syncwarp();                                     // all column-slice reads of O_left are done
  if (has_left) {
#pragma unroll
    for (int rr = 0; rr < RW; ++rr) sts_row<T, L>(N + C::OL + (r0 + rr) * L, Gm[rr]);
  }
  __syncwarp();                                     // F and G rows of every lane are visible

  // my rows of F F^T and G G^T (right-hand rows from shared memory)
  T Am[RW][L], Bm[RW][L];
#pragma unroll
  for (int c = 0; c < L; ++c) {
    T fc[L], gc[L];
    lds_row<T, L>(fc, N + C::OR_ + c * L);
    lds_row<T, L>(gc, N + C::OL + c * L);
#pragma unroll
    for (int rr = 0; rr < RW; ++rr) {
      T sa = T(0), sb = T(0);
#pragma unroll
      for (int k = 0; k < L; ++k) { sa = fma(Fm[rr][k], fc[k], sa); sb = fma(Gm[rr][k], gc[k], sb); }
      Am[rr][c] = do_f ? sa : T(0);
      Bm[rr][c] = has_left ? sb : T(0);
    }
    sched_fence();
  }
  // my rows of O~_{e-1} = -F G^T
  if (do_f && has_left) {
#pragma unroll
    for (int rr = 0; rr < RW; ++rr) {
      T on[L];
#pragma unroll
      for (int c = 0; c < L; ++c) {
        T gc[L];
        lds_row<T, L>(gc, N + C::OL + c * L);
        T s = T(0);
#pragma unroll
        for (int k = 0; k < L; ++k) s = fma(-Fm[rr][k], gc[k], s);
        on[c] = s;
      }
      sts_row<T, L>(N + C::ON + (r0 + rr) * L, on);
      sched_fence();
    }
  }
  if (halo && e0 == 0 && t == 0) {
    // link to the virtual node -1: accumulate -G G^T and -G x there (rows of this lane)
    if (a.Rh_acc != nullptr) {
      T* acc = static_cast<T*>(a.Rh_acc) + (size_t)b * BS;
#pragma unroll
      for (int rr = 0; rr < RW; ++rr)
#pragma unroll
        for (int c = 0; c < L; ++c) acc[(r0 + rr) * L + c] -= Bm[rr][c];
    }
    if (a.yh_acc != nullptr && has_y) {
      T* acc = static_cast<T*>(a.yh_acc) + (size_t)b * L;
#pragma unroll
      for (int rr = 0; rr < RW; ++rr) acc[r0 + rr] -= vm[rr];
    }
  }
  // G G^T and G x of the NEXT even node: same row block, LPN lanes further
#pragma unroll
  for (int rr = 0; rr < RW; ++rr) {
#pragma unroll
    for (int c = 0; c < L; ++c) Bm[rr][c] = __shfl_down_sync(0xffffffffu, Bm[rr][c], LPN);
    vm[rr] = __shfl_down_sync(0xffffffffu, vm[rr], LPN);
  }
  if (do_f) {
    const bool next_even = (e + 1) < E;
#pragma unroll
    for (int rr = 0; rr < RW; ++rr) {
      T row[L];
      lds_row<T, L>(row, N + C::RO + (r0 + rr) * L);
#pragma unroll
      for (int c = 0; c < L; ++c) row[c] = row[c] - Am[rr][c] - (next_even ? Bm[rr][c] : T(0));
      sts_row<T, L>(N + C::RO + (r0 + rr) * L, row);
    }
    if (has_y) {
      T yo[RW];
      lds_row<T, RW>(yo, N + C::YO + r0);
#pragma unroll
      for (int rr = 0; rr < RW; ++rr) yo[rr] = yo[rr] - um[rr] - (next_even ? vm[rr] : T(0));
      sts_row<T, RW>(N + C::YO + r0, yo);
    }
  }

  if (a.logdet != nullptr) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) ld_part += __shfl_xor_sync(0xffffffffu, ld_part, off);
    if (lane == 0) atomicAdd(a.logdet + acc_index(a, b, tile), ld_part);
  }
  if (a.mahal != nullptr && has_y) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) mh_part += __shfl_xor_sync(0xffffffffu, mh_part, off);
    if (lane == 0) atomicAdd(a.mahal + acc_index(a, b, tile), mh_part);
  }
  __syncwarp();

  // ---------------- stage out ----------------
  const int n_own = cmin(OWN, E - e0);
  const int n_odd = cmax(0, cmin(OWN, o - e0));
  if (a.D != nullptr) {
    rec_s2g<T, BS, 1>(static_cast<T*>(a.D) + ((size_t)b * E + e0) * BS, s0 + C::RE * ES, nsb, 0, n_own, is_aligned16(a.D));
    rec_s2g<T, BS, 1>(static_cast<T*>(a.F) + ((size_t)b * o + e0) * BS, s0 + C::OR_ * ES, nsb, 0, n_odd, is_aligned16(a.F));
    const int gfirst = (e0 == 0) ? 1 : 0;
    rec_s2g<T, BS, 1>(static_cast<T*>(a.G) + ((size_t)b * gcnt + (e0 + gfirst - 1)) * BS, s0 + C::OL * ES, nsb, gfirst, n_own - gfirst,
                      is_aligned16(a.G));
  }
  if (a.xk != nullptr && has_y)
    rec_s2g<T, L, 1>(static_cast<T*>(a.xk) + ((size_t)b * E + e0) * L, s0 + C::YE * ES, nsb, 0, n_own, is_aligned16(a.xk));
  if (a.Rn != nullptr && n_odd > 0) {
    rec_s2g<T, BS, 1>(static_cast<T*>(a.Rn) + ((size_t)b * o + e0) * BS, s0 + C::RO * ES, nsb, 0, n_odd, is_aligned16(a.Rn));
    if (has_y && a.yn != nullptr)
      rec_s2g<T, L, 1>(static_cast<T*>(a.yn) + ((size_t)b * o + e0) * L, s0 + C::YO * ES, nsb, 0, n_odd, is_aligned16(a.yn));
    const int ofirst = (e0 == 0) ? 1 : 0;
    const int n_on = cmax(0, cmin(e0 + n_own, o) - (e0 + ofirst));
    if (a.On != nullptr && n_on > 0)
      rec_s2g<T, BS, 1>(static_cast<T*>(a.On) + ((size_t)b * (o - 1) + (e0 + ofirst - 1)) * BS, s0 + C::ON * ES, nsb, ofirst, n_on,
                        is_aligned16(a.On));
  }
  if (halo && e0 == 0) {
    if (a.G_halo != nullptr)
      rec_s2g<T, BS, 1>(static_cast<T*>(a.G_halo) + (size_t)b * BS, s0 + C::OL * ES, nsb, 0, 1, is_aligned16(a.G_halo));
    if (a.On_halo != nullptr && o > 0)
      rec_s2g<T, BS, 1>(static_cast<T*>(a.On_halo) + (size_t)b * BS, s0 + C::ON * ES, nsb, 0, 1, is_aligned16(a.On_halo));
  }
}

template <typename T, int L, int LPN>
cudaError_t launch_cs_fwd(const LevelFwdArgs& a, cudaStream_t stream) {
  using C = CsFwdCfg<T, L, LPN>;
  static std::atomic<unsigned char> attr_done[kMaxDevices];
  if (cudaError_t e = ensure_dynamic_smem(cr_cs_fwd_kernel<T, L, LPN>, (int)C::SMEM, attr_done); e != cudaSuccess) return e;
  const int E = (a.m + 1) / 2;
  const long long tiles = (E + C::OWN - 1) / C::OWN;
  const long long grid = tiles * a.batch;
  if (grid <= 0) return cudaSuccess;
  if (grid > 0x7fffffffLL) return cudaErrorInvalidValue;
  cr_cs_fwd_kernel<T, L, LPN><<<(unsigned)grid, 32, C::SMEM, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace crb200
