/* Pure C client of libcrb200 (no Python, no torch): what a non-Python host of the reference's CR path would
 * link against.  Builds a small symmetric positive-definite block-tridiagonal system, runs
 *   crb200_sweep_fwd      (decompose + mahal_and_det, cyclic_reduction.py:288-309, :380-438)
 *   crb200_sweep_bwd      (solve = backhalfsolve(halfsolve), :341-377, and inverse_blocks, :470-503)
 * on the GPU and checks log|J|, x^T J^{-1} x, J^{-1} x and the diagonal blocks of J^{-1} against a dense Cholesky
 * computed here on the CPU.  Exit code 0 = all within 1e-10 (fp64).
 *
 *   gcc -std=c99 -O2 -I include -I /usr/local/cuda/include tests/c_abi/abi_smoke.c -o abi_smoke \
 *       -L cyclic-gps_b200 -lcrb200 -L /usr/local/cuda/lib64 -lcudart -lm -Wl,-rpath,$PWD/cyclic-gps_b200
 */
#include <cuda_runtime_api.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "crb200.h"

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #call, cudaGetErrorString(e_)); return 2; } } while (0)

static double frand(unsigned* s) { *s = *s * 1664525u + 1013904223u; return ((*s >> 8) & 0xffffff) / (double)0x1000000 - 0.5; }

int main(void) {
  const int n = 77, l = 3, N = n * l, bs = l * l;
  unsigned seed = 12345u;
  /* J = block tridiagonal, diagonally dominant => SPD */
  double* R = calloc((size_t)n * bs, sizeof(double));
  double* O = calloc((size_t)(n - 1) * bs, sizeof(double));
  double* x = calloc((size_t)N, sizeof(double));
  for (int i = 0; i < n; ++i)
    for (int r = 0; r < l; ++r)
      for (int c = 0; c <= r; ++c) {
        double v = (r == c) ? 4.0 + frand(&seed) : 0.3 * frand(&seed);
        R[(size_t)i * bs + r * l + c] = v;
        R[(size_t)i * bs + c * l + r] = v;
      }
  for (int i = 0; i < (n - 1) * bs; ++i) O[i] = 0.8 * frand(&seed);
  for (int i = 0; i < N; ++i) x[i] = frand(&seed);

  /* ---- dense reference on the CPU: Cholesky of the assembled matrix ---- */
  double* J = calloc((size_t)N * N, sizeof(double));
  for (int i = 0; i < n; ++i)
    for (int r = 0; r < l; ++r)
      for (int c = 0; c < l; ++c) {
        J[(size_t)(i * l + r) * N + i * l + c] = R[(size_t)i * bs + r * l + c];
        if (i + 1 < n) {                      /* O[i] = J_{i+1,i} */
          J[(size_t)((i + 1) * l + r) * N + i * l + c] = O[(size_t)i * bs + r * l + c];
          J[(size_t)(i * l + c) * N + (i + 1) * l + r] = O[(size_t)i * bs + r * l + c];
        }
      }
  double* Lc = calloc((size_t)N * N, sizeof(double));
  double logdet = 0.0;
  for (int j = 0; j < N; ++j) {
    double d = J[(size_t)j * N + j];
    for (int k = 0; k < j; ++k) d -= Lc[(size_t)j * N + k] * Lc[(size_t)j * N + k];
    if (d <= 0) { fprintf(stderr, "test matrix not SPD\n"); return 2; }
    Lc[(size_t)j * N + j] = sqrt(d);
    logdet += 2.0 * log(Lc[(size_t)j * N + j]);
    for (int i = j + 1; i < N; ++i) {
      double s = J[(size_t)i * N + j];
      for (int k = 0; k < j; ++k) s -= Lc[(size_t)i * N + k] * Lc[(size_t)j * N + k];
      Lc[(size_t)i * N + j] = s / Lc[(size_t)j * N + j];
    }
  }
  double* w = malloc(sizeof(double) * N);   /* J^{-1} x */
  for (int i = 0; i < N; ++i) { double s = x[i]; for (int k = 0; k < i; ++k) s -= Lc[(size_t)i * N + k] * w[k]; w[i] = s / Lc[(size_t)i * N + i]; }
  double mahal = 0.0;
  for (int i = 0; i < N; ++i) mahal += w[i] * w[i];
  for (int i = N - 1; i >= 0; --i) { double s = w[i]; for (int k = i + 1; k < N; ++k) s -= Lc[(size_t)k * N + i] * w[k]; w[i] = s / Lc[(size_t)i * N + i]; }
  /* diagonal of J^{-1}: solve for unit vectors (small N) */
  double* diagInv = malloc(sizeof(double) * N);
  double* col = malloc(sizeof(double) * N);
  for (int j = 0; j < N; ++j) {
    for (int i = 0; i < N; ++i) { double s = (i == j); for (int k = 0; k < i; ++k) s -= Lc[(size_t)i * N + k] * col[k]; col[i] = s / Lc[(size_t)i * N + i]; }
    for (int i = N - 1; i >= 0; --i) { double s = col[i]; for (int k = i + 1; k < N; ++k) s -= Lc[(size_t)k * N + i] * col[k]; col[i] = s / Lc[(size_t)i * N + i]; }
    diagInv[j] = col[j];
  }

  /* ---- the same through the C ABI ---- */
  int L = 0, E_tot = 0, o_tot = 0, g_tot = 0;
  for (int m = n; m >= 1; m /= 2) { ++L; E_tot += (m + 1) / 2; o_tot += m / 2; g_tot += (m - 1) / 2; if (m == 1) break; }
  double *dR, *dO, *dx, *dD, *dF, *dG, *dX, *scr[6], *dacc, *dSd, *dSo, *dw, *bs_[6];
  int* dinfo;
  CK(cudaMalloc((void**)&dR, sizeof(double) * n * bs));
  CK(cudaMalloc((void**)&dO, sizeof(double) * (n - 1) * bs));
  CK(cudaMalloc((void**)&dx, sizeof(double) * N));
  CK(cudaMalloc((void**)&dD, sizeof(double) * E_tot * bs));
  CK(cudaMalloc((void**)&dF, sizeof(double) * (o_tot + 1) * bs));
  CK(cudaMalloc((void**)&dG, sizeof(double) * (g_tot + 1) * bs));
  CK(cudaMalloc((void**)&dX, sizeof(double) * E_tot * l));
  for (int i = 0; i < 6; ++i) { CK(cudaMalloc((void**)&scr[i], sizeof(double) * (n / 2 + 1) * bs)); CK(cudaMalloc((void**)&bs_[i], sizeof(double) * (n / 2 + 1) * bs)); }
  CK(cudaMalloc((void**)&dacc, sizeof(double) * 2));
  CK(cudaMalloc((void**)&dinfo, sizeof(int) * L));
  CK(cudaMalloc((void**)&dSd, sizeof(double) * n * bs));
  CK(cudaMalloc((void**)&dSo, sizeof(double) * n * bs));
  CK(cudaMalloc((void**)&dw, sizeof(double) * N));
  CK(cudaMemcpy(dR, R, sizeof(double) * n * bs, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dO, O, sizeof(double) * (n - 1) * bs, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dx, x, sizeof(double) * N, cudaMemcpyHostToDevice));
  CK(cudaMemset(dacc, 0, sizeof(double) * 2));
  CK(cudaMemset(dinfo, 0, sizeof(int) * L));

  crb200_sweep_fwd_args f;
  memset(&f, 0, sizeof f);
  f.batch = 1; f.n = n; f.nlevels = L;
  f.R = dR; f.O = dO; f.y = dx;
  f.strideR = (long long)n * bs; f.strideO = (long long)(n - 1) * bs; f.stridey = N;
  f.D = dD; f.F = dF; f.G = dG; f.X = dX;
  f.scrR[0] = scr[0]; f.scrR[1] = scr[1]; f.scrO[0] = scr[2]; f.scrO[1] = scr[3]; f.scry[0] = scr[4]; f.scry[1] = scr[5];
  f.logdet = dacc; f.mahal = dacc + 1; f.info = dinfo; f.acc_slots = 1;
  f.variant = CRB200_AUTO;
  int rc = crb200_sweep_fwd(CRB200_F64, l, &f, NULL);
  if (rc != CRB200_OK) { fprintf(stderr, "crb200_sweep_fwd -> %d (cuda %d)\n", rc, crb200_last_cuda_error()); return 1; }

  crb200_sweep_bwd_args b;
  memset(&b, 0, sizeof b);
  b.batch = 1; b.n = n; b.nlevels = L;
  b.D = dD; b.F = dF; b.G = dG; b.X = dX;
  b.Sd_out = dSd; b.So_out = dSo; b.w_out = dw;
  b.strideSd = (long long)n * bs; b.strideSo = (long long)(n - 1) * bs; b.stridew = N;
  b.scrSd[0] = bs_[0]; b.scrSd[1] = bs_[1]; b.scrSo[0] = bs_[2]; b.scrSo[1] = bs_[3]; b.scrw[0] = bs_[4]; b.scrw[1] = bs_[5];
  b.variant = CRB200_AUTO;
  rc = crb200_sweep_bwd(CRB200_F64, l, &b, NULL);
  if (rc != CRB200_OK) { fprintf(stderr, "crb200_sweep_bwd -> %d (cuda %d)\n", rc, crb200_last_cuda_error()); return 1; }
  CK(cudaDeviceSynchronize());

  double acc[2];
  int info[40];
  double* Sd = malloc(sizeof(double) * n * bs);
  double* wg = malloc(sizeof(double) * N);
  CK(cudaMemcpy(acc, dacc, sizeof acc, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(info, dinfo, sizeof(int) * L, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(Sd, dSd, sizeof(double) * n * bs, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(wg, dw, sizeof(double) * N, cudaMemcpyDeviceToHost));
  for (int k = 0; k < L; ++k) if (info[k] != 0) { fprintf(stderr, "non-PD block reported at level %d\n", k); return 1; }

  double e_ld = fabs(2.0 * acc[0] - logdet) / fabs(logdet), e_mh = fabs(acc[1] - mahal) / fabs(mahal), e_w = 0, wmax = 0, e_s = 0, smax = 0;
  for (int i = 0; i < N; ++i) { e_w = fmax(e_w, fabs(wg[i] - w[i])); wmax = fmax(wmax, fabs(w[i])); }
  for (int i = 0; i < n; ++i)
    for (int r = 0; r < l; ++r) { e_s = fmax(e_s, fabs(Sd[(size_t)i * bs + r * l + r] - diagInv[i * l + r])); smax = fmax(smax, fabs(diagInv[i * l + r])); }
  printf("levels %d  launches %lld  rel.err: logdet %.2e  mahal %.2e  solve %.2e  diag(J^-1) %.2e\n", L, crb200_launch_count(), e_ld, e_mh, e_w / wmax, e_s / smax);
  const int ok = e_ld < 1e-10 && e_mh < 1e-10 && e_w / wmax < 1e-10 && e_s / smax < 1e-10;
  printf(ok ? "abi_smoke ok\n" : "abi_smoke FAILED\n");
  return ok ? 0 : 1;
}
