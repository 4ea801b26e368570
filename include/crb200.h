/* crb200 -- C ABI of the B200-native block cyclic-reduction (CR) engine.
 *
 * Drop-in boundary for the hot path of cunningham-lab/cyclic-gps: the module-level
 * functions of cyclic_gps/cyclic_reduction.py (the reference has no FFI; its "interface" is
 * that Python module, imported at cyclic_gps/models.py:10).  Each entry below names the
 * reference code it replaces.  The library allocates nothing and takes raw DEVICE pointers plus a
 * cudaStream_t (as void*).
 *
 * Threading: every entry is re-entrant and may be called from any host thread (the autograd engine calls the
 * backward entries from its own thread).  The only process-wide state is a diagnostic launch counter and two
 * per-device caches held in atomics (kernel shared-memory opt-in done, SM count), which are idempotent under
 * races.  The CURRENT CUDA device of the calling thread must be the device that owns `stream` and the buffers
 * (the caches and the fused-tail decision are keyed by cudaGetDevice()); torch callers get this from
 * torch.cuda.device / set_device.  Argument errors (CRB200_EINVAL, CRB200_EUNSUPPORTED) are detected before
 * anything is launched.
 *
 * Conventions
 *   dtype     CRB200_F32 / CRB200_F64; all arrays of one call share it.
 *   ell       block size, 1 <= ell <= crb200_max_ell().
 *   level     a symmetric positive-definite block-tridiagonal system with m block rows:
 *             R (batch, m, ell, ell) diagonal blocks (only the lower triangle is read by the
 *             factorisation), O (batch, m-1, ell, ell) lower off-diagonal blocks
 *             (O[i] = J_{i+1,i}), y (batch, m, ell).  Row-major, contiguous per series.
 *   counts    E = ceil(m/2) even (eliminated) nodes, o = floor(m/2) odd (surviving) nodes,
 *             g = floor((m-1)/2) G links.
 *   return    0 ok; <0 = CRB200_E*.  Non-positive-definite blocks do not fail the call:
 *             they are reported through `info` (see crb200_fwd_args) and produce NaNs.
 */
#ifndef CRB200_H_
#define CRB200_H_

#ifdef __cplusplus
extern "C" {
#endif

#define CRB200_F32 0
#define CRB200_F64 1

/* kernel families: thread-per-node (sizeof(T)*ell*ell <= 400 B: fp32 ell <= 10, fp64 ell <= 7), column-split (several lanes
 * per node: fp32 ell=8, fp64 ell=4 and 8; chosen automatically for fp64 ell=8), warp-per-node with the block products on the
 * FP64 tensor path (DMMA; ell >= 8, chosen automatically where it measured faster than lane-per-row: fp64 ell >= 10, fp32 ell >= 17; fp32 data is
 * widened to fp64 on chip) and the first-generation lane-per-row kernels (any ell <= 32; kept as a cross-check) */
#define CRB200_AUTO 0
#define CRB200_LANE_PER_ROW 1
#define CRB200_THREAD_PER_NODE 2
#define CRB200_COLUMN_SPLIT 3
#define CRB200_MMA 4
#define CRB200_COPY_ONLY 99       /* profiling aid (thread-per-node kernels): stage in / out only, results are garbage */

#define CRB200_OK 0
#define CRB200_EINVAL (-1)        /* null / inconsistent argument                      */
#define CRB200_EUNSUPPORTED (-2)  /* ell or dtype outside the compiled range           */
#define CRB200_ECUDA (-3)         /* CUDA launch error; see crb200_last_cuda_error()   */

/* One CR level, forward.  Replaces decompose_step (cyclic_reduction.py:204-259) fused with
 * the per-level bodies of mahal_and_det (:412-427) and halfsolve (:318-333). */
typedef struct crb200_fwd_args {
  int batch, m;
  const void* R; const void* O; const void* y;        /* y may be NULL (factor only)              */
  long long strideR, strideO, stridey;                /* series strides of R, O, y in ELEMENTS     */
  void* D; void* F; void* G; void* xk;                /* out: K (batch,E,l,l) lower-tri, F (batch,o,l,l),
                                                         G (batch,g,l,l), x_k (batch,E,l).  D/F/G all NULL
                                                         => factors not kept; xk NULL => not stored  */
  void* Rn; void* On; void* yn;                       /* out: reduced system (batch,o,..),(batch,o-1,..),(batch,o,l) */
  double* logdet; double* mahal;                      /* in/out per-series accumulators (+= sum log diag K,
                                                         += |x_k|^2), NULL to skip; each is (batch, acc_slots) */
  int acc_slots;                                      /* >= 1 (0 = 1): a CTA adds into slot (tile mod acc_slots), which
                                                         spreads the atomics of long series over several addresses  */
  int* info;                                          /* in/out: max over failures of INT_MAX - (series*E + e);
                                                         caller zero-initialises; 0 = all blocks PD  */
  /* left halo (chunk-partitioned series): virtual surviving node -1 coupled to row 0 by O_halo */
  const void* O_halo; void* G_halo; void* On_halo; void* Rh_acc; void* yh_acc;   /* (batch,l,l)x4, (batch,l) */
  int variant;                                        /* CRB200_AUTO, or force one kernel family (tests, benchmarks) */
  int tri;                                            /* packed lower triangles (crb200_tri_stride(dtype, ell) elements per block, 0 = not
                                                         offered): bit 0 = R is packed (strideR counts packed elements), bit 1 = D and Rn
                                                         are written packed.  0 = full blocks everywhere */
} crb200_fwd_args;

/* One CR level, backward direction (deepest level first).  Replaces the per-level bodies of
 * backhalfsolve (:362-373) and inverse_blocks (:478-501); with grad_mode it also assembles
 * the gradient of gm*mahal + gd*logdet wrt (R, O, x) that torch autograd produces through
 * the reference (models.py:374-381). */
typedef struct crb200_bwd_args {
  int batch, m;
  const void* D; const void* F; const void* G; const void* xk;   /* factors of this level            */
  const void* Sd_in; const void* So_in; const void* w_in;        /* deeper level: (batch,o,..),(batch,o-1,..),(batch,o,l) */
  void* Sd_out; void* So_out; void* w_out;                       /* out: (batch,m,..),(batch,m-1,..),(batch,m,l);
                                                                    Sd_out/So_out NULL => solve only;
                                                                    w_out NULL => selected inverse only */
  long long strideSd, strideSo, stridew;                         /* series strides of the outputs (elements) */
  const double* gm; const double* gd;                            /* per-series cotangents (grad_mode)  */
  int grad_mode;                                                 /* 0: Sigma / w ; 1: gR / gO / gx     */
  const void* G_halo; const void* Sd_halo; const void* w_halo; const void* So_halo_in; void* So_halo_out;
  int variant;
  int tri;                                                       /* bit 0 = D and Sd_in are packed lower triangles, bit 1 = Sd_out is written
                                                                    packed (strideSd counts packed elements; never with grad_mode) */
} crb200_bwd_args;

/* Half solve against stored factors.  Replaces one iteration of halfsolve (:318-333):
 * x_k = D^{-1} y[0::2];  yn = y[1::2] - U x_k (Ux :40-60);  mahal += |x_k|^2 (mahal :461-467). */
typedef struct crb200_hs_args {
  int batch, m;
  const void* D; const void* F; const void* G;
  const void* y; long long stridey;
  void* xk; void* yn;
  double* mahal;
} crb200_hs_args;

/* Whole sweeps: the level loop runs inside the library (one host call, ~3 us per level instead of a
 * Python round trip).  With variant = CRB200_AUTO the deep levels of a sweep (a series has at most 8 tiles, level >= 3)
 * run as ONE fused launch where the thread-per-node family exists: one CTA per series walks those levels, writing
 * every level's intermediate result at its own offset inside the scratch buffers described below (they are large
 * enough as specified); the results the caller reads back sit where they are documented to sit.  Only batches of at
 * most two series per SM are fused (every series keeps a CTA for the whole tail).
 * Packed per-level storage: level k of a family starts at element offset
 * batch * unit * sum_{j<k} rows_j  with rows_j = E_j (D, X), o_j (F), g_j (G) and unit = l*l (l for X);
 * m_0 = n, m_{k+1} = floor(m_k / 2).  Ping-pong scratch: level k writes its result to slot [k & 1].
 *
 * Forward = decompose (:288-309) / mahal_and_det (:380-438) / the halo sweep of a sub-chunk batch.
 * After the call the system left below level nlevels-1 (if any) sits in scr*[(nlevels-1)&1] and the last
 * halo coupling in On_halo[(nlevels-1)&1]. */
typedef struct crb200_sweep_fwd_args {
  int batch, n, nlevels;
  const void* R; const void* O; const void* y;
  long long strideR, strideO, stridey;
  void* D; void* F; void* G; void* X;                 /* packed factors; D/F/G NULL => not kept; X NULL => x_k not kept */
  void* scrR[2]; void* scrO[2]; void* scry[2];        /* [0] >= batch*floor(n/2) rows, [1] >= batch*floor(n/4) rows */
  double* logdet; double* mahal; int* info;           /* info: nlevels ints, zero-initialised by the caller */
  int acc_slots;                                      /* logdet / mahal are (batch, acc_slots), see crb200_fwd_args */
  const void* O_halo; void* G_halo;                   /* G_halo: nlevels * batch blocks (NULL => not kept) */
  void* On_halo[2]; void* Rh_acc; void* yh_acc;
  int variant;
  int tri;                                            /* 1: D and the reduced diagonal blocks of the scratch are packed lower triangles (block stride
                                                         crb200_tri_stride; every level keeps the offset it has with full blocks).  Only for a full
                                                         reduction without halo whose factors go to crb200_sweep_bwd with tri = 1 and nowhere else */
} crb200_sweep_fwd_args;

/* Backward = backhalfsolve (:341-377) + inverse_blocks (:470-503) + gradient assembly, deepest level first.
 * Level k >= 1 writes slot [k & 1] (slot [1] >= batch*m_1 rows, slot [0] >= batch*m_2 rows); level 0 writes
 * Sd_out / So_out / w_out.  The final halo off-diagonal block goes to So_halo_out. */
typedef struct crb200_sweep_bwd_args {
  int batch, n, nlevels;
  const void* D; const void* F; const void* G; const void* X;
  const void* top_Sd; const void* top_So; const void* top_w;   /* solution of the system below the deepest level (or NULL) */
  void* Sd_out; void* So_out; void* w_out;
  long long strideSd, strideSo, stridew;
  void* scrSd[2]; void* scrSo[2]; void* scrw[2];
  const double* gm; const double* gd; int grad_mode;
  const void* G_halo; const void* Sd_halo; const void* w_halo; const void* So_halo_in;
  void* So_halo[2]; void* So_halo_out;
  int variant;
  int tri;                                            /* 1: D was written by crb200_sweep_fwd with tri = 1; the inner levels then also pass Sigma_d packed */
} crb200_sweep_bwd_args;

/* All levels of halfsolve (:312-338) / mahal (:461-467) against packed factors: X receives x_k of every level
 * (packed like D rows); scry[] ping-pong as in the forward sweep; mahal (batch doubles, NULL to skip). */
typedef struct crb200_sweep_hs_args {
  int batch, n, nlevels;
  const void* D; const void* F; const void* G;
  const void* y; long long stridey;
  void* X; void* scry[2];
  double* mahal;
} crb200_sweep_hs_args;

int crb200_sweep_fwd(int dtype, int ell, const crb200_sweep_fwd_args* args, void* stream);
int crb200_sweep_halfsolve(int dtype, int ell, const crb200_sweep_hs_args* args, void* stream);
int crb200_sweep_bwd(int dtype, int ell, const crb200_sweep_bwd_args* args, void* stream);

/* Precision-block builder of the LEG / PEG process.  Replaces LEGFamily.compute_PEG_precision (models.py:181-239) with the
 * posterior shift of compute_posterior_precision (:254-268) folded in, and torch autograd through them.  The caller
 * eigendecomposes G = V diag(lam) V^{-1} once on the host (as the reference's compute_eG does, model_utils.py:12-29) and
 * passes lam, M_k = V[:,k] V^{-1}[k,:] (k-major, ell*ell each) as DEVICE arrays of doubles; gaps are d_g = t_{g+1} - t_g > 0.
 * Forward writes R (batch,n,l,l) and O (batch,n-1,l,l).  Backward turns cotangents gR / gO into the cotangent gA_g of every
 * A_g = exp(c_g G), c_g = -d_g / 2, and returns the 2*l weighted sums  S[row] = sum_g E[row][g] gA_g^T  (l x l each, fp64,
 * ACCUMULATED into -- zero S first).  Rows, in order, for every m < nterms: Re e^{c lam_m}, c Re e^{c lam_m}, and if lam_im[m] != 0
 * also Im e^{c lam_m}, c Im e^{c lam_m} (conjugate pairs folded: always 2*l rows).  The caller finishes the adjoint of the matrix
 * exponential: T_m = V^{-1} (sum_g e^{c lam_m} gA_g^T) V, Z_jk = (T_j - T_k)[k][j] / (lam_j - lam_k) (the c-weighted sum where the
 * eigenvalues coincide), gG = Re(V^{-T} Z V^T).  Row 2*l of S is sum_i gR_i, the cotangent of `shift`.
 * info (may be NULL): set to 1 if some I - A A^T was not positive definite (a non-positive gap). */
typedef struct crb200_peg_fwd_args {
  int batch, n;
  const void* gaps; long long stride_gaps;             /* (batch, n-1), element type = dtype; series stride in elements */
  const double* lam_re; const double* lam_im;
  const double* M_re; const double* M_im;
  const double* shift;                                 /* (l,l) added to every diagonal block, or NULL */
  void* R; void* O; long long strideR, strideO;
  int* info;
  int nterms;                                          /* 0 = ell.  exp(cG) - I = Re sum_{k < nterms} (e^{c lam_k} - 1) M_k: a caller that folds
                                                          every conjugate pair of eigenvalues into one term (M_k doubled, partner dropped)
                                                          passes fewer than ell terms and halves the work of the expansion */
  double* logdet;                                      /* (batch) or NULL: ACCUMULATES log det of the block-tridiagonal matrix (R without shift, O)
                                                          of every series = -sum_g logdet(I - A_g A_g^T), from the Cholesky factors the kernel forms
                                                          anyway.  It is the `prior_logdet` the reference gets from a second cyclic reduction
                                                          (models.py:349-353: det(decompose(Sigma^{-1}))); zero it first */
} crb200_peg_fwd_args;

typedef struct crb200_peg_bwd_args {
  int batch, n;
  const void* gaps; long long stride_gaps;
  const double* lam_re; const double* lam_im; const double* M_re; const double* M_im;   /* as in crb200_peg_fwd_args */
  const void* O; long long strideO;                    /* the forward result */
  const void* gR; const void* gO; long long stride_gR, stride_gO;
  double* S;                                           /* (2*ell + 1, ell*ell) doubles, ACCUMULATED into: see above; the last row is sum_i gR_i */
  int nterms;                                          /* as in crb200_peg_fwd_args */
  const double* g_logdet;                              /* (batch) or NULL: cotangent of the forward's logdet output (adds 2 g B_g to gA_g) */
  double* gA;                                          /* ell > crb200_peg_sum_max_ell() only: (batch, n-1, ell, ell) doubles, receives gA_g per gap
                                                          (row-major); the weighted sums over the gaps are then the caller's GEMM and S is not touched */
} crb200_peg_bwd_args;

int crb200_peg_precision_fwd(int dtype, int ell, const crb200_peg_fwd_args* args, void* stream);
int crb200_peg_precision_bwd(int dtype, int ell, const crb200_peg_bwd_args* args, void* stream);
int crb200_peg_max_ell(void);                          /* largest ell the builder kernels exist for (32) */
int crb200_peg_sum_max_ell(void);                      /* largest ell for which the backward entry also sums over the gaps (8: thread-per-gap kernels);
                                                          above it the warp-per-gap kernels return gA per gap (field gA) */

int crb200_version(void);
int crb200_max_ell(void);
/* cudaError_t of the last failing launch on the calling thread's most recent call (0 if none). */
int crb200_last_cuda_error(void);
/* Diagnostic: number of kernels this library has launched in the process so far (bench.py's gpu_launches). */
long long crb200_launch_count(void);

int crb200_level_fwd(int dtype, int ell, const crb200_fwd_args* args, void* stream);
int crb200_level_bwd(int dtype, int ell, const crb200_bwd_args* args, void* stream);
int crb200_level_halfsolve(int dtype, int ell, const crb200_hs_args* args, void* stream);

/* Elements per packed lower triangle (ell (ell + 1) / 2 rounded up to 16 bytes) where the level kernels offer the `tri`
 * storage of the symmetric / triangular blocks (float32, ell = 8: 36 instead of 64), else 0. */
int crb200_tri_stride(int dtype, int ell);

/* Tile geometry, for the roofline bookkeeping in bench.py: even nodes owned per CTA. */
int crb200_fwd_tile_nodes(int dtype, int ell);
int crb200_bwd_tile_nodes(int dtype, int ell);

#ifdef __cplusplus
}
#endif
#endif /* CRB200_H_ */
