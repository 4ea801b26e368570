#!/usr/bin/env python
"""Run the reference's OWN test files, unchanged, against this package on a B200.

    python tools/run_reference_tests.py --prepare     # build container: copy the three files into .ref_scratch/
    python tools/run_reference_tests.py --run         # GPU box: pytest them with cyclic-gps_b200/ first on PYTHONPATH

`--prepare` copies tests/test_cyclic_reduction.py, tests/test_likelihood.py and tests/known_matrices_full.py (a
third-party LGPL fixture file) from /root/reference into `.ref_scratch/`, which is git-ignored (nothing of the reference
enters the history) but travels to the GPU box with the `gpurun` snapshot.  `--run` executes them with pytest; the
imports `cyclic_gps.cyclic_reduction`, `cyclic_gps.models`, `cyclic_gps.model_utils`, `cyclic_gps.data_utils` and
`cyclic_gps.kalman` resolve to this repo's package (CUDA engine, filterpy-free Kalman comparator).  The log goes to
gpurun_out/r2_reference_tests.log (copied to profiles/ by hand once it is green)."""
import argparse
import hashlib
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCRATCH = os.path.join(ROOT, ".ref_scratch")
FILES = ("test_cyclic_reduction.py", "test_likelihood.py", "known_matrices_full.py")


def prepare(ref):
    os.makedirs(SCRATCH, exist_ok=True)
    for f in FILES:
        shutil.copyfile(os.path.join(ref, "tests", f), os.path.join(SCRATCH, f))
    print("copied", FILES, "->", SCRATCH)


def run(out):
    missing = [f for f in FILES if not os.path.exists(os.path.join(SCRATCH, f))]
    if missing:
        raise SystemExit(f"{SCRATCH} lacks {missing}: run --prepare in the build container first")
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(ROOT, "cyclic-gps_b200"), SCRATCH, env.get("PYTHONPATH", "")])
    sums = {f: hashlib.sha256(open(os.path.join(SCRATCH, f), "rb").read()).hexdigest()[:16] for f in FILES}
    cmd = [sys.executable, "-m", "pytest", "-p", "no:cacheprovider", "-v", "--rootdir", SCRATCH, "-c", os.devnull,
           os.path.join(SCRATCH, "test_cyclic_reduction.py"), os.path.join(SCRATCH, "test_likelihood.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, cwd=SCRATCH)
    probe = subprocess.run([sys.executable, "-c",
                            "import cyclic_gps.cyclic_reduction as c, cyclic_gps._native as n, torch;"
                            "print('cyclic_gps from', c.__file__); print('native library', n.LIB_PATH, 'version', n.load().crb200_version());"
                            "print('device', torch.cuda.get_device_name(0))"], env=env, capture_output=True, text=True)
    os.makedirs(os.path.dirname(out), exist_ok=True)
    with open(out, "w") as fh:
        fh.write("# reference test files (unchanged, sha256/16): %s\n" % sums)
        fh.write("# command: %s\n# PYTHONPATH=%s\n" % (" ".join(cmd), env["PYTHONPATH"]))
        fh.write(probe.stdout + probe.stderr)
        fh.write(r.stdout[-20000:] + "\n" + r.stderr[-5000:])
    print(r.stdout[-3000:])
    return r.returncode


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--prepare", action="store_true")
    ap.add_argument("--run", action="store_true")
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r2_reference_tests.log"))
    a = ap.parse_args()
    if a.prepare:
        prepare(a.ref)
    if a.run:
        sys.exit(run(a.out))
