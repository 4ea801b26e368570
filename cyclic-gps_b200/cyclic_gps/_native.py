"""ctypes binding of libcrb200.so (include/crb200.h) -- the C-ABI boundary of the CR engine.

The library is built in-tree by ``cyclic-gps_b200/build.py`` and found next to this package.
There is NO fallback: if the shared object is missing, or there is no CUDA device, the calls
raise.  Tensors are passed as raw device pointers (``Tensor.data_ptr()``) together with the
current CUDA stream; torch owns every buffer."""
from __future__ import annotations

import ctypes as C
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "libcrb200.so")

F32, F64 = 0, 1
OK, EINVAL, EUNSUPPORTED, ECUDA = 0, -1, -2, -3

_vp, _ll, _i = C.c_void_p, C.c_longlong, C.c_int


class FwdArgs(C.Structure):
    _fields_ = [("batch", _i), ("m", _i),
                ("R", _vp), ("O", _vp), ("y", _vp),
                ("strideR", _ll), ("strideO", _ll), ("stridey", _ll),
                ("D", _vp), ("F", _vp), ("G", _vp), ("xk", _vp),
                ("Rn", _vp), ("On", _vp), ("yn", _vp),
                ("logdet", _vp), ("mahal", _vp), ("acc_slots", _i), ("info", _vp),
                ("O_halo", _vp), ("G_halo", _vp), ("On_halo", _vp), ("Rh_acc", _vp), ("yh_acc", _vp), ("variant", _i), ("tri", _i)]


class BwdArgs(C.Structure):
    _fields_ = [("batch", _i), ("m", _i),
                ("D", _vp), ("F", _vp), ("G", _vp), ("xk", _vp),
                ("Sd_in", _vp), ("So_in", _vp), ("w_in", _vp),
                ("Sd_out", _vp), ("So_out", _vp), ("w_out", _vp),
                ("strideSd", _ll), ("strideSo", _ll), ("stridew", _ll),
                ("gm", _vp), ("gd", _vp), ("grad_mode", _i),
                ("G_halo", _vp), ("Sd_halo", _vp), ("w_halo", _vp), ("So_halo_in", _vp), ("So_halo_out", _vp), ("variant", _i), ("tri", _i)]


class HsArgs(C.Structure):
    _fields_ = [("batch", _i), ("m", _i),
                ("D", _vp), ("F", _vp), ("G", _vp),
                ("y", _vp), ("stridey", _ll),
                ("xk", _vp), ("yn", _vp), ("mahal", _vp)]


class SweepFwdArgs(C.Structure):
    _fields_ = [("batch", _i), ("n", _i), ("nlevels", _i),
                ("R", _vp), ("O", _vp), ("y", _vp),
                ("strideR", _ll), ("strideO", _ll), ("stridey", _ll),
                ("D", _vp), ("F", _vp), ("G", _vp), ("X", _vp),
                ("scrR", _vp * 2), ("scrO", _vp * 2), ("scry", _vp * 2),
                ("logdet", _vp), ("mahal", _vp), ("info", _vp), ("acc_slots", _i),
                ("O_halo", _vp), ("G_halo", _vp), ("On_halo", _vp * 2), ("Rh_acc", _vp), ("yh_acc", _vp),
                ("variant", _i), ("tri", _i)]


class SweepBwdArgs(C.Structure):
    _fields_ = [("batch", _i), ("n", _i), ("nlevels", _i),
                ("D", _vp), ("F", _vp), ("G", _vp), ("X", _vp),
                ("top_Sd", _vp), ("top_So", _vp), ("top_w", _vp),
                ("Sd_out", _vp), ("So_out", _vp), ("w_out", _vp),
                ("strideSd", _ll), ("strideSo", _ll), ("stridew", _ll),
                ("scrSd", _vp * 2), ("scrSo", _vp * 2), ("scrw", _vp * 2),
                ("gm", _vp), ("gd", _vp), ("grad_mode", _i),
                ("G_halo", _vp), ("Sd_halo", _vp), ("w_halo", _vp), ("So_halo_in", _vp),
                ("So_halo", _vp * 2), ("So_halo_out", _vp),
                ("variant", _i), ("tri", _i)]


class SweepHsArgs(C.Structure):
    _fields_ = [("batch", _i), ("n", _i), ("nlevels", _i),
                ("D", _vp), ("F", _vp), ("G", _vp),
                ("y", _vp), ("stridey", _ll),
                ("X", _vp), ("scry", _vp * 2), ("mahal", _vp)]


_dp = C.POINTER(C.c_double)


class PegFwdArgs(C.Structure):
    _fields_ = [("batch", _i), ("n", _i), ("gaps", _vp), ("stride_gaps", _ll),
                ("lam_re", _vp), ("lam_im", _vp), ("M_re", _vp), ("M_im", _vp), ("shift", _vp),
                ("R", _vp), ("O", _vp), ("strideR", _ll), ("strideO", _ll), ("info", _vp), ("nterms", _i), ("logdet", _vp)]


class PegBwdArgs(C.Structure):
    _fields_ = [("batch", _i), ("n", _i), ("gaps", _vp), ("stride_gaps", _ll),
                ("lam_re", _vp), ("lam_im", _vp), ("M_re", _vp), ("M_im", _vp),
                ("O", _vp), ("strideO", _ll), ("gR", _vp), ("gO", _vp), ("stride_gR", _ll), ("stride_gO", _ll), ("S", _vp),
                ("nterms", _i), ("g_logdet", _vp), ("gA", _vp)]


EXPORTS = ("crb200_peg_precision_fwd", "crb200_peg_precision_bwd", "crb200_peg_max_ell", "crb200_peg_sum_max_ell", "crb200_version", "crb200_max_ell", "crb200_last_cuda_error", "crb200_level_fwd",
           "crb200_level_bwd", "crb200_level_halfsolve", "crb200_sweep_fwd", "crb200_sweep_bwd", "crb200_sweep_halfsolve",
           "crb200_fwd_tile_nodes", "crb200_bwd_tile_nodes", "crb200_launch_count", "crb200_tri_stride")

_lib = None
_lock = threading.Lock()


class NativeLibraryMissing(RuntimeError):
    pass


def load():
    """Load libcrb200.so once; raises NativeLibraryMissing when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise NativeLibraryMissing(
                f"{LIB_PATH} not found: build it with `python cyclic-gps_b200/build.py` "
                "(needs nvcc; there is no CPU or PyTorch fallback for the CR engine)")
        lib = C.CDLL(LIB_PATH)
        for name in ("crb200_version", "crb200_max_ell", "crb200_last_cuda_error"):
            getattr(lib, name).restype = _i
            getattr(lib, name).argtypes = []
        lib.crb200_level_fwd.restype = _i
        lib.crb200_level_fwd.argtypes = [_i, _i, C.POINTER(FwdArgs), _vp]
        lib.crb200_level_bwd.restype = _i
        lib.crb200_level_bwd.argtypes = [_i, _i, C.POINTER(BwdArgs), _vp]
        lib.crb200_level_halfsolve.restype = _i
        lib.crb200_level_halfsolve.argtypes = [_i, _i, C.POINTER(HsArgs), _vp]
        lib.crb200_sweep_fwd.restype = _i
        lib.crb200_sweep_fwd.argtypes = [_i, _i, C.POINTER(SweepFwdArgs), _vp]
        lib.crb200_sweep_halfsolve.restype = _i
        lib.crb200_sweep_halfsolve.argtypes = [_i, _i, C.POINTER(SweepHsArgs), _vp]
        lib.crb200_sweep_bwd.restype = _i
        lib.crb200_sweep_bwd.argtypes = [_i, _i, C.POINTER(SweepBwdArgs), _vp]
        lib.crb200_launch_count.restype = C.c_longlong
        lib.crb200_launch_count.argtypes = []
        lib.crb200_peg_precision_fwd.restype = _i
        lib.crb200_peg_precision_fwd.argtypes = [_i, _i, C.POINTER(PegFwdArgs), _vp]
        lib.crb200_peg_precision_bwd.restype = _i
        lib.crb200_peg_precision_bwd.argtypes = [_i, _i, C.POINTER(PegBwdArgs), _vp]
        lib.crb200_peg_max_ell.restype = _i
        lib.crb200_peg_max_ell.argtypes = []
        lib.crb200_peg_sum_max_ell.restype = _i
        lib.crb200_peg_sum_max_ell.argtypes = []
        for name in ("crb200_fwd_tile_nodes", "crb200_bwd_tile_nodes", "crb200_tri_stride"):
            getattr(lib, name).restype = _i
            getattr(lib, name).argtypes = [_i, _i]
        _lib = lib
    return _lib


def dtype_code(dtype: torch.dtype) -> int:
    if dtype == torch.float32:
        return F32
    if dtype == torch.float64:
        return F64
    raise TypeError(f"the CR engine computes in float32 or float64, got {dtype}")


def _ptr(t):
    if t is None or isinstance(t, int):      # raw device address (engine workspaces)
        return t
    if not t.is_cuda:
        raise RuntimeError("internal error: CPU tensor handed to the native CR library")
    return t.data_ptr()


def _check(rc: int, what: str):
    if rc == OK:
        return
    lib = load()
    if rc == ECUDA:
        raise RuntimeError(f"{what}: CUDA launch failed (cudaError {lib.crb200_last_cuda_error()})")
    if rc == EUNSUPPORTED:
        raise ValueError(f"{what}: unsupported block size or dtype (1 <= ell <= {lib.crb200_max_ell()}, float32/float64)")
    raise ValueError(f"{what}: invalid argument (code {rc})")


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _fill(struct, fields):
    for k, v in fields.items():
        if isinstance(v, torch.Tensor) or v is None:
            setattr(struct, k, _ptr(v))
        elif isinstance(v, (tuple, list)):          # pair of ping-pong buffers
            arr = getattr(struct, k)
            for i, t in enumerate(v):
                arr[i] = _ptr(t)
        else:
            setattr(struct, k, v)
    return struct


# Kernel family override for tests / benchmarks: 0 auto, 1 lane-per-row, 2 thread-per-node, 3 column-split, 4 warp-per-node DMMA.
VARIANT = int(os.environ.get("CRB200_VARIANT", "0"))

# Optional launch tracer (bench.py): an object with begin(kind, dtype, ell, batch, m) -> token
# and end(token); called around every native launch on the current stream.
TRACE = None


def _traced(kind, fn, dtype, ell, a):
    tr = TRACE
    tok = tr.begin(kind, dtype, ell, a.batch, a.m) if tr is not None else None
    rc = fn(dtype_code(dtype), ell, C.byref(a), _stream())
    if tr is not None:
        tr.end(tok)
    _check(rc, "crb200_level_" + kind)


def level_fwd(dtype: torch.dtype, ell: int, **fields):
    fields.setdefault("variant", VARIANT)
    _traced("fwd", load().crb200_level_fwd, dtype, ell, _fill(FwdArgs(), fields))


def level_bwd(dtype: torch.dtype, ell: int, **fields):
    fields.setdefault("variant", VARIANT)
    _traced("bwd", load().crb200_level_bwd, dtype, ell, _fill(BwdArgs(), fields))


def level_halfsolve(dtype: torch.dtype, ell: int, **fields):
    _traced("halfsolve", load().crb200_level_halfsolve, dtype, ell, _fill(HsArgs(), fields))


def sweep_fwd(dtype: torch.dtype, ell: int, **fields):
    """All forward levels in one library call (the level loop runs in C)."""
    fields.setdefault("variant", VARIANT)
    a = _fill(SweepFwdArgs(), fields)
    tr = TRACE
    if tr is not None and getattr(tr, "enabled", True):
        # per-launch tracing needs one host call per level: bench.py switches to the level entries
        raise RuntimeError("launch tracing is only available through the per-level entries")
    _check(load().crb200_sweep_fwd(dtype_code(dtype), ell, C.byref(a), _stream()), "crb200_sweep_fwd")


def sweep_bwd(dtype: torch.dtype, ell: int, **fields):
    fields.setdefault("variant", VARIANT)
    a = _fill(SweepBwdArgs(), fields)
    _check(load().crb200_sweep_bwd(dtype_code(dtype), ell, C.byref(a), _stream()), "crb200_sweep_bwd")


def sweep_halfsolve(dtype: torch.dtype, ell: int, **fields):
    a = _fill(SweepHsArgs(), fields)
    _check(load().crb200_sweep_halfsolve(dtype_code(dtype), ell, C.byref(a), _stream()), "crb200_sweep_halfsolve")


def peg_fwd(dtype: torch.dtype, ell: int, **fields):
    """Precision blocks (R, O) of the LEG process from time gaps (crb200_peg_precision_fwd)."""
    a = _fill(PegFwdArgs(), fields)
    _check(load().crb200_peg_precision_fwd(dtype_code(dtype), ell, C.byref(a), _stream()), "crb200_peg_precision_fwd")


def peg_bwd(dtype: torch.dtype, ell: int, **fields):
    a = _fill(PegBwdArgs(), fields)
    _check(load().crb200_peg_precision_bwd(dtype_code(dtype), ell, C.byref(a), _stream()), "crb200_peg_precision_bwd")


def peg_max_ell() -> int:
    return int(load().crb200_peg_max_ell())


def peg_sum_max_ell() -> int:
    return int(load().crb200_peg_sum_max_ell())


def launch_count() -> int:
    """Kernels launched by libcrb200 in this process so far."""
    return int(load().crb200_launch_count())


def tracing() -> bool:
    tr = TRACE
    return tr is not None and getattr(tr, "enabled", True)


# Packed lower triangles for the blocks that stay inside the library (crb200.h `tri`): CRB200_TRI=0 switches them off.
TRI = os.environ.get("CRB200_TRI", "1") != "0"


def tri_stride(dtype: torch.dtype, ell: int) -> int:
    """Elements per packed lower triangle where the level kernels offer that storage (float32, ell = 8), else 0."""
    if not TRI or VARIANT not in (0, 2):
        return 0
    return int(load().crb200_tri_stride(dtype_code(dtype), ell))


def tile_nodes(dtype: torch.dtype, ell: int):
    lib = load()
    return lib.crb200_fwd_tile_nodes(dtype_code(dtype), ell), lib.crb200_bwd_tile_nodes(dtype_code(dtype), ell)
