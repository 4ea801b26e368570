"""Build libcrb200.so (hand-written sm_100a kernels + C ABI) in-tree with nvcc.

    python cyclic-gps_b200/build.py [--force] [-j N]

The template instantiations are split over 16 translation units (dtype x ell range) and
compiled in parallel; objects are cached under cyclic-gps_b200/build/ keyed by a hash of
the sources and flags.  The resulting .so sits next to this file (git-ignored, but it travels
to the GPU box with the repo snapshot)."""
import argparse
import concurrent.futures as cf
import hashlib
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libcrb200.so")
BUILD = os.path.join(HERE, "build")
STAMP = os.path.join(HERE, "libcrb200.stamp")
RANGES = [(1, 4), (5, 8), (9, 12), (13, 16), (17, 20), (21, 24), (25, 28), (29, 32)]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-I" + os.path.join(ROOT, "include"), "-I" + CSRC] + os.environ.get("CRB200_NVCC_EXTRA", "").split()


def _nvcc():
    cand = os.environ.get("NVCC")
    if cand:
        return cand
    if os.path.exists("/usr/local/cuda/bin/nvcc"):
        return "/usr/local/cuda/bin/nvcc"
    return "nvcc"


def source_hash():
    h = hashlib.sha256()
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))] + [os.path.join(ROOT, "include", "crb200.h")]
    for f in files:
        h.update(os.path.basename(f).encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS[:6]).encode())
    return h.hexdigest()[:16]


def _unit_key(cmd):
    """Hash of the preprocessed translation unit (line markers stripped) + the command line:
    an object is rebuilt only when something it actually includes has changed."""
    pre = [c for c in cmd if c not in ("-c",)]
    o = pre.index("-o")
    pre = pre[:o] + pre[o + 2:] + ["-E"]
    r = subprocess.run(pre, capture_output=True, text=True)
    if r.returncode != 0:
        return None
    body = "\n".join(l for l in r.stdout.splitlines() if l.strip() and not l.startswith("#"))
    return hashlib.sha256((" ".join(cmd[1:]) + body).encode()).hexdigest()[:20]


def _compile_unit(unit, verbose=False):
    obj, cmd = unit
    key = _unit_key(cmd)
    keyfile = obj + ".key"
    if key is not None and os.path.exists(obj) and os.path.exists(keyfile) and open(keyfile).read().strip() == key:
        return ""
    out = _run(cmd)
    if key is not None:
        with open(keyfile, "w") as fh:
            fh.write(key)
    return out


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("command failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
    return r.stdout + r.stderr


def is_current():
    stamp = STAMP
    return os.path.exists(OUT) and os.path.exists(stamp) and open(stamp).read().strip() == source_hash()


def build(force=False, jobs=None, verbose=False, only=None):
    """only: e.g. "f32:5-8,f64:1-4" -> development build, every other (dtype, range) unit is a
    stub that reports cudaErrorNotSupported (never leaves a `current` stamp behind)."""
    if not force and only is None and is_current():
        return OUT
    tag = source_hash()       # taken BEFORE compiling: edits made during the build must not be stamped
    os.makedirs(BUILD, exist_ok=True)
    nvcc = _nvcc()
    units = []
    for tn, ty in (("f32", "float"), ("f64", "double")):
        for lo, hi in RANGES:
            obj = os.path.join(BUILD, f"inst_{tn}_{lo}_{hi}.o")
            stub = ["-DCRB_STUB"] if (only is not None and f"{tn}:{lo}-{hi}" not in only.split(",")) else []
            units.append((obj, [nvcc] + NVCC_FLAGS + stub + [f"-DCRB_T={ty}", f"-DCRB_TN={tn}", f"-DCRB_LO={lo}", f"-DCRB_HI={hi}",
                                                              "-c", os.path.join(CSRC, "cr_inst.cu"), "-o", obj]))
    abi_obj = os.path.join(BUILD, "cr_abi.o")
    units.append((abi_obj, [nvcc] + NVCC_FLAGS + ["-c", os.path.join(CSRC, "cr_abi.cu"), "-o", abi_obj]))
    peg_obj = os.path.join(BUILD, "cr_peg_inst.o")          # precision-block builder (cr_peg.cuh), ell = 1..8, both dtypes
    units.append((peg_obj, [nvcc] + NVCC_FLAGS + ["-c", os.path.join(CSRC, "cr_peg_inst.cu"), "-o", peg_obj]))
    for tn, ty in (("f32", "float"), ("f64", "double")):      # warp-per-gap builder (cr_pegw.cuh), ell = 9..32
        for lo, hi in RANGES[2:]:
            obj = os.path.join(BUILD, f"pegw_{tn}_{lo}_{hi}.o")
            units.append((obj, [nvcc] + NVCC_FLAGS + [f"-DCRB_T={ty}", f"-DCRB_TN={tn}", f"-DCRB_LO={lo}", f"-DCRB_HI={hi}",
                                                      "-c", os.path.join(CSRC, "cr_pegw_inst.cu"), "-o", obj]))
    jobs = jobs or min(len(units), os.cpu_count() or 4)
    # longest units (large ell) first
    units.sort(key=lambda u: -int(u[0].rsplit("_", 1)[-1].split(".")[0]) if "inst_" in u[0] else 0)
    with cf.ThreadPoolExecutor(max_workers=jobs) as ex:
        for out in ex.map((lambda u: _run(u[1])) if force else _compile_unit, units):
            if verbose and out.strip():
                print(out)
    _run([nvcc, "-shared", "-o", OUT] + [u[0] for u in units] + ["-lcudart"])
    with open(STAMP, "w") as fh:
        fh.write(tag if only is None else "partial:" + only)
    return OUT


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("-j", type=int, default=None)
    ap.add_argument("-v", action="store_true")
    ap.add_argument("--only", default=None, help='development build, e.g. "f32:5-8,f64:1-4"')
    a = ap.parse_args()
    print(build(force=a.force, jobs=a.j, verbose=a.v, only=a.only))
