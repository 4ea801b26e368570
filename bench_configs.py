#!/usr/bin/env python
"""Timings + parity for the BASELINE.json configs that bench.py does not cover as its headline line:

  cfg1  LEG rank 3, single series n = 1000, fp64: decompose + mahal_and_det + solve vs the reference CPU path
        and a dense Cholesky of the assembled 3000 x 3000 matrix
  cfg3  CO2-shaped training step (n = 502 with a 240-unit gap, rank 16, fp64, obs_dim 1): log_likelihood
        forward + backward and the in-sample posterior (decompose + solve + inverse_blocks)
  cfg5  kalman_timing_script sweep: n = 1e3 .. 1e7, rank in {2, 8, 32}, fp64, regular gaps, obs_dim 2:
        posterior (decompose + solve + inverse_blocks) and log-likelihood (decompose/det + mahal_and_det),
        GPU vs the reference CPU path (oracle port, bounded n) and a numpy Kalman filter (bounded n)

(cfg2 and cfg4 are `bench.py` and `bench.py --workload long`.)  One JSON document is written to --out.
GPU times: CUDA events, median of --reps after 3 warm-ups; CPU times: perf_counter, best of 2."""
import argparse
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "cyclic-gps_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402


def gpu_time(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)


def cpu_time(fn, reps=2):
    fn()
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        dt = (time.perf_counter() - t0) * 1e3
        best = dt if best is None else min(best, dt)
    return best


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-300))


def cfg1(reps):
    from cyclic_gps import cyclic_reduction as cr
    from oracle import cr_oracle as orc
    n, l = 1000, 3
    G, B, LLT = orc.leg_params(l, seed=1)
    gen = torch.Generator().manual_seed(0)
    gaps = -torch.log(torch.rand(n - 1, generator=gen, dtype=torch.float64)) + 0.01
    R, O = orc.leg_posterior_precision(gaps, G, B, LLT)
    x = torch.randn((n, l), generator=gen, dtype=torch.float64)
    Rg, Og, xg = R.cuda(), O.cuda(), x.cuda()

    def gpu():
        dec = cr.decompose(Rg, Og)
        mm, dd = cr.mahal_and_det(Rg, Og, xg)
        return dec, mm, dd, cr.solve(dec, xg)

    def cpu():
        dec = orc.factor(R, O)
        mm, dd = orc.mahal_and_logdet(R, O, x)
        return dec, mm, dd, orc.solve(dec, x)

    J = orc.assemble_dense(R, O)

    def dense():
        Lc = torch.linalg.cholesky(J)
        w = torch.cholesky_solve(x.reshape(-1, 1), Lc).reshape(n, l)
        return 2 * torch.log(torch.diagonal(Lc)).sum(), (x * w).sum(), w

    from cyclic_gps.graphs import GraphedMahalAndDet

    def loglik_grad_eager():
        Rr, Or, xr = Rg.detach().requires_grad_(True), Og.detach().requires_grad_(True), xg.detach().requires_grad_(True)
        mm, dd = cr.mahal_and_det(Rr, Or, xr)
        (mm + dd).backward()
        return mm, dd, Rr.grad

    graphed = GraphedMahalAndDet(Rg, Og, xg)

    def loglik_grad_graphed():
        return graphed(Rg, Og, xg)

    def cpu_loglik_grad():
        return orc.loglik_grads_autograd(R, O, x)

    dec, mm, dd, w = gpu()
    ld_d, mh_d, w_d = dense()
    _, mm_o, dd_o, w_o = cpu()
    mh_g, ld_g, gR_g, _, _ = loglik_grad_graphed()
    _, _, gR_e = loglik_grad_eager()
    return {"config": "cfg1: n=1000, l=3, fp64, decompose + mahal_and_det + solve",
            "gpu_ms": gpu_time(gpu, reps), "reference_cpu_ms": cpu_time(cpu), "dense_cholesky_cpu_ms": cpu_time(dense),
            "loglik_grad_gpu_eager_ms": gpu_time(loglik_grad_eager, reps),
            "loglik_grad_gpu_cuda_graph_ms": gpu_time(loglik_grad_graphed, reps),
            "loglik_grad_reference_cpu_ms": cpu_time(cpu_loglik_grad),
            "graph_vs_eager": {"mahal": rel(mh_g, mm), "logdet": rel(ld_g, dd), "gR": rel(gR_g, gR_e)},
            "cores": torch.get_num_threads(),
            "parity_vs_reference_path": {"mahal": rel(mm, mm_o), "logdet": rel(dd, dd_o), "solve": rel(w, w_o)},
            "parity_vs_dense": {"mahal": rel(mm, mh_d), "logdet": rel(dd, ld_d), "solve": rel(w, w_d)}}


def cfg3(reps):
    from cyclic_gps.models import LEGFamily
    from oracle import cr_oracle as orc
    import math
    torch.manual_seed(3)
    n, rank = 502, 16
    gaps = torch.ones(n - 1, dtype=torch.float64)
    gaps[261] = 240.0
    ts = torch.cat([torch.zeros(1, dtype=torch.float64), torch.cumsum(gaps, 0)]) / 12.0
    xs = torch.randn(n, 1, dtype=torch.float64)
    model = LEGFamily(rank=rank, obs_dim=1, train=True, data_type=torch.float64).cuda()
    tsg, xsg = ts.cuda(), xs.cuda()

    def step():
        model.zero_grad(set_to_none=True)
        ll = model.log_likelihood(tsg, xsg)
        (-ll / n).backward()
        return ll

    def posterior():
        with torch.no_grad():
            model.register_model_matrices_from_params()
            return model.compute_insample_posterior(tsg, xsg)

    ref = LEGFamily(rank=rank, obs_dim=1, train=True, data_type=torch.float64)
    ref.load_state_dict({k: v.cpu() for k, v in model.state_dict().items()})

    def ref_step():
        ref.zero_grad(set_to_none=True)
        ref.register_model_matrices_from_params()
        LLT, shift = ref._obs_terms()
        white = torch.linalg.solve(LLT, xs.T).T
        Rs, Os = ref.compute_PEG_precision(ts)
        mh, ld = orc.mahal_and_logdet(Rs + shift.unsqueeze(0), Os, white @ ref.B)
        prior = orc.logdet(orc.factor(Rs, Os))
        ll = -0.5 * ((torch.sum(white * xs) - mh) + (torch.logdet(2 * math.pi * LLT) * n + ld - prior))
        (-ll / n).backward()
        return ll

    def ref_posterior():
        with torch.no_grad():
            ref.register_model_matrices_from_params()
            dec = orc.factor(*ref.compute_posterior_precision(ts))
            return orc.solve(dec, ref.compute_v(xs)), orc.selected_inverse(dec)

    from cyclic_gps.graphs import GraphedMahalAndDet
    with torch.no_grad():
        model.register_model_matrices_from_params()
        _, shift_g = model._obs_terms()
        Kr, Ko = model._precision_blocks(tsg, shift_g)
        vg = model.compute_v(xsg)
    graphed = GraphedMahalAndDet(Kr, Ko, vg)
    from cyclic_gps.graphs import GraphedLogLikelihood
    runner = GraphedLogLikelihood(model, tsg, xsg)

    def graphed_step(move=False):
        model.zero_grad(set_to_none=True)
        ll = runner()
        (-ll / n).backward()
        if move:        # an optimiser step changes G: the eigen-constants are recomputed on the host and refreshed in place
            with torch.no_grad():
                model.R_params.add_(1e-6)
        return ll

    # the same runner with the (73-number) model on the HOST: the l x l / d x d parameter algebra then costs no launches at all
    hmodel = LEGFamily(rank=rank, obs_dim=1, train=True, data_type=torch.float64)
    hmodel.load_state_dict({k: v.cpu() for k, v in model.state_dict().items()})
    hrunner = GraphedLogLikelihood(hmodel, tsg, xsg)

    def graphed_step_host_model():
        hmodel.zero_grad(set_to_none=True)
        ll = hrunner()
        (-ll / n).backward()
        with torch.no_grad():
            hmodel.R_params.add_(1e-6)
        return ll

    ll_g = graphed_step()
    g_graphed = {k: getattr(model, k).grad.clone() for k in ("N_params", "R_params", "Lambda_params", "B")}
    ll, ll_o = step(), ref_step()
    g_err = max(rel(getattr(model, k).grad, getattr(ref, k).grad) for k in ("N_params", "R_params", "Lambda_params", "B"))
    gg_err = max(rel(g_graphed[k], getattr(ref, k).grad) for k in g_graphed)
    (mean, cov), (mean_o, (sd, so)) = posterior(), ref_posterior()
    return {"config": "cfg3: CO2-shaped, n=502 (240-unit gap), l=16, fp64, obs_dim=1",
            "gpu_train_step_ms": gpu_time(step, reps), "gpu_posterior_ms": gpu_time(posterior, reps),
            "gpu_cr_loglik_grad_cuda_graph_ms": gpu_time(lambda: graphed(Kr, Ko, vg), reps),
            "gpu_train_step_cuda_graph_ms": gpu_time(graphed_step, reps),
            "gpu_train_step_cuda_graph_moving_parameters_ms": gpu_time(lambda: graphed_step(True), reps),
            "gpu_train_step_cuda_graph_host_model_ms": gpu_time(graphed_step_host_model, reps),
            "reference_cpu_train_step_ms": cpu_time(ref_step), "reference_cpu_posterior_ms": cpu_time(ref_posterior),
            "cores": torch.get_num_threads(),
            "parity": {"loglik": rel(ll, ll_o), "param_grads": g_err, "graphed_loglik": rel(ll_g, ll_o), "graphed_param_grads": gg_err,
                       "posterior_mean": rel(mean, mean_o),
                       "posterior_cov_diag": rel(cov["Rs"], sd), "posterior_cov_off": rel(cov["Os"], so)}}


def kalman_loglik(A, Q, H, Rm, xs):
    """numpy Kalman filter log-likelihood (the comparator of cyclic_gps/kalman.py:54-60, filterpy-free)."""
    d = A.shape[0]
    x, P, ll = np.zeros((d, 1)), np.eye(d), 0.0
    for z in xs:
        x, P = A @ x, A @ P @ A.T + Q
        y = z.reshape(-1, 1) - H @ x
        S = H @ P @ H.T + Rm
        K = np.linalg.solve(S, H @ P).T
        x = x + K @ y
        P = (np.eye(d) - K @ H) @ P
        ll += float(-0.5 * (y.T @ np.linalg.solve(S, y)).item() - 0.5 * np.linalg.slogdet(2 * np.pi * S)[1])
    return ll


def cfg5(reps, max_n, cpu_budget_ms=60e3):
    from cyclic_gps import cyclic_reduction as cr
    from cyclic_gps.synth import leg_precision_blocks
    from oracle import cr_oracle as orc
    from scipy.linalg import expm
    out = []
    dev = torch.device("cuda")
    for l in (2, 8, 32):
        gen = torch.Generator().manual_seed(l)
        A0 = torch.randn((l, l), generator=gen, dtype=torch.float64)
        Rm = torch.tril((A0 - A0.T) * 0.2, diagonal=-1)
        G = torch.eye(l, dtype=torch.float64) + Rm - Rm.T + 1e-5 * torch.eye(l, dtype=torch.float64)
        B = torch.full((2, l), 0.5 / l ** 0.5, dtype=torch.float64)
        B[1] *= torch.linspace(0.5, 1.5, l, dtype=torch.float64)
        LLT = 0.55 * torch.eye(2, dtype=torch.float64)
        last_ref = last_kal = None          # (n, posterior ms, loglik ms) / (n, ms) of the largest size actually timed on the CPU
        for n in (10 ** 3, 10 ** 4, 10 ** 5, 10 ** 6, 10 ** 7):
            if n > max_n:
                continue
            if n * l * l * 8 * 14 > 150e9:
                out.append({"l": l, "n": n, "skipped": "does not fit one 180 GB B200 in fp64: blocks %.0f GB + factors %.0f GB + selected inverse %.0f GB"
                            % (2 * n * l * l * 8 / 1e9, 3 * n * l * l * 8 / 1e9, 2 * n * l * l * 8 / 1e9)})
                continue
            gaps = torch.ones((1, n - 1), dtype=torch.float64, device=dev)
            R, O = leg_precision_blocks(gaps, G.to(dev), B.to(dev), LLT.to(dev), torch.float64, chunk=max(1 << 14, (1 << 26) // (l * l)))
            R, O = R[0], O[0]
            xs = torch.randn((n, 2), dtype=torch.float64, device=dev, generator=torch.Generator(device=dev).manual_seed(n))
            v = torch.linalg.solve(LLT.to(dev), xs.T).T @ B.to(dev)

            def posterior():
                dec = cr.decompose(R, O)
                return cr.solve(dec, v), cr.inverse_blocks(dec)

            def loglik():
                return cr.mahal_and_det(R, O, v)

            row = {"l": l, "n": n, "gpu_posterior_ms": gpu_time(posterior, reps), "gpu_loglik_ms": gpu_time(loglik, reps),
                   "gpu_posterior_two_sweeps_ms": gpu_time(lambda: cr.solve_and_inverse_blocks(R, O, v), reps)}
            row["gpu_posterior_rows_per_s"] = n / (row["gpu_posterior_ms"] * 1e-3)
            est = None if last_ref is None else last_ref[1] * n / last_ref[0]
            if est is None or est <= cpu_budget_ms:
                Rc, Oc, vc = R.cpu(), O.cpu(), v.cpu()

                def ref_posterior():
                    dec = orc.factor(Rc, Oc)
                    return orc.solve(dec, vc), orc.selected_inverse(dec)

                row["reference_cpu_posterior_ms"] = cpu_time(ref_posterior, reps=1)
                row["reference_cpu_loglik_ms"] = cpu_time(lambda: orc.mahal_and_logdet(Rc, Oc, vc), reps=1)
                last_ref = (n, row["reference_cpu_posterior_ms"], row["reference_cpu_loglik_ms"])
                (w, (sd, so)), (w_o, (sd_o, so_o)) = posterior(), ref_posterior()
                mm, dd = loglik()
                mm_o, dd_o = orc.mahal_and_logdet(Rc, Oc, vc)
                row["parity"] = {"mean": rel(w, w_o), "cov_diag": rel(sd, sd_o), "cov_off": rel(so, so_o), "mahal": rel(mm, mm_o),
                                 "logdet": rel(dd, dd_o)}
            elif last_ref is not None:      # the reference scales linearly in n (kalman_timing_script.py:77,85): extrapolated, and marked so
                row["reference_cpu_posterior_ms_extrapolated"] = last_ref[1] * n / last_ref[0]
                row["reference_cpu_loglik_ms_extrapolated"] = last_ref[2] * n / last_ref[0]
                row["reference_cpu_extrapolated_from_n"] = last_ref[0]
            kest = None if last_kal is None else last_kal[1] * n / last_kal[0]
            if kest is not None and kest > cpu_budget_ms:
                row["numpy_kalman_loglik_ms_extrapolated"] = kest
                row["numpy_kalman_extrapolated_from_n"] = last_kal[0]
            else:
                Ad = expm(-0.5 * G.numpy())
                Qd = np.eye(l) - Ad @ Ad.T
                xs_c = xs.cpu().numpy()
                t0 = time.perf_counter()
                kll = kalman_loglik(Ad, Qd, B.numpy(), LLT.numpy(), xs_c)
                row["numpy_kalman_loglik_ms"] = (time.perf_counter() - t0) * 1e3
                last_kal = (n, row["numpy_kalman_loglik_ms"])
                # LEG log-likelihood from the CR quantities (models.py:301-372) vs the Kalman filter
                mm, dd = loglik()
                white = torch.linalg.solve(LLT.to(dev), xs.T).T
                shift = B.T @ torch.linalg.solve(LLT, B)
                prior = cr.det(cr.decompose(R - shift.to(dev).unsqueeze(0), O))
                ll = -0.5 * ((torch.sum(white * xs) - mm) + (torch.logdet(2 * np.pi * LLT) * n + dd - prior))
                row["loglik_vs_kalman_rel"] = abs(float(ll) - kll) / abs(kll)
            out.append(row)
            del R, O, xs, v
            torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "configs.json"))
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--max-n", type=int, default=10 ** 7)
    ap.add_argument("--only", default="1,3,5")
    ap.add_argument("--cpu-budget-s", type=float, default=60.0, help="cfg5: largest single CPU comparator run that is still timed (larger n: extrapolated)")
    args = ap.parse_args()
    if not torch.cuda.is_available():
        raise SystemExit("needs a CUDA device")
    torch.set_num_threads(os.cpu_count() or 1)
    res = {"gpu": torch.cuda.get_device_name(0), "host_cores": os.cpu_count()}
    if "1" in args.only:
        res["cfg1"] = cfg1(args.reps)
    if "3" in args.only:
        res["cfg3"] = cfg3(args.reps)
    if "5" in args.only:
        res["cfg5"] = cfg5(max(3, args.reps // 2), args.max_n, args.cpu_budget_s * 1e3)
    os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
    with open(args.out, "w") as fh:
        json.dump(res, fh, indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
