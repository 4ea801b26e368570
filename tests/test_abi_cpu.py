"""CPU: the C-ABI shared library loads and exports every symbol include/crb200.h declares
(no compute without a GPU), and the product path fails loudly instead of falling back."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "crb200.h")).read()
    return sorted(set(re.findall(r"\b(crb200_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    from cyclic_gps import _native
    if not os.path.exists(_native.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(_native.LIB_PATH)
    names = _declared_symbols()
    assert set(names) == set(_native.EXPORTS), (names, _native.EXPORTS)
    for n in names:
        assert getattr(lib, n) is not None
    assert lib.crb200_max_ell() == 32
    assert lib.crb200_version() >= 100


def test_argument_validation_without_gpu():
    from cyclic_gps import _native
    lib = _native.load()
    a = _native.FwdArgs()
    a.batch, a.m = 1, 4
    assert lib.crb200_level_fwd(0, 8, ctypes.byref(a), None) == _native.EINVAL      # null R
    assert lib.crb200_level_fwd(0, 33, ctypes.byref(a), None) == _native.EUNSUPPORTED
    assert lib.crb200_level_fwd(5, 8, ctypes.byref(a), None) == _native.EUNSUPPORTED
    b = _native.BwdArgs()
    b.batch, b.m = 1, 4
    assert lib.crb200_level_bwd(1, 8, ctypes.byref(b), None) == _native.EINVAL
    assert lib.crb200_fwd_tile_nodes(0, 8) == 31 and lib.crb200_bwd_tile_nodes(0, 8) >= 1


def test_packed_triangle_requests_are_validated_without_gpu():
    """`tri` (packed lower triangles): offered for float32 ell = 8 only, never with a halo, a forced non-thread-per-node family,
    a partial sweep or the gradient mode of an inner-level output; all of it is rejected before anything is launched."""
    from cyclic_gps import _native
    lib = _native.load()
    assert lib.crb200_tri_stride(_native.F32, 8) == 36
    assert all(lib.crb200_tri_stride(dt, l) == 0 for dt in (_native.F32, _native.F64) for l in range(1, 33) if (dt, l) != (_native.F32, 8))
    assert lib.crb200_tri_stride(7, 8) == 0 and lib.crb200_tri_stride(_native.F32, 40) == 0
    buf = ctypes.create_string_buffer(1 << 16)                      # host memory: only its (aligned) address is looked at
    base = (ctypes.addressof(buf) + 255) & ~255
    a = _native.FwdArgs()
    a.batch, a.m = 1, 4
    for f in ("R", "O", "D", "F", "G", "Rn", "On"):
        setattr(a, f, base)
    a.tri = 2
    assert lib.crb200_level_fwd(_native.F32, 4, ctypes.byref(a), None) == _native.EINVAL        # size without packed storage
    assert lib.crb200_level_fwd(_native.F64, 8, ctypes.byref(a), None) == _native.EINVAL
    a.variant = 1                                                                                  # lane-per-row family forced
    assert lib.crb200_level_fwd(_native.F32, 8, ctypes.byref(a), None) == _native.EINVAL
    a.variant, a.tri = 0, 4                                                                        # unknown bit
    assert lib.crb200_level_fwd(_native.F32, 8, ctypes.byref(a), None) == _native.EINVAL
    a.tri, a.Rn = 2, base + 4                                                                      # packed output not 16-byte aligned
    assert lib.crb200_level_fwd(_native.F32, 8, ctypes.byref(a), None) == _native.EINVAL
    a.Rn = base
    a.O_halo = a.Rh_acc = a.On_halo = a.G_halo = base                                              # chunked (halo) sweeps keep full blocks
    assert lib.crb200_level_fwd(_native.F32, 8, ctypes.byref(a), None) == _native.EINVAL
    b = _native.BwdArgs()
    b.batch, b.m = 1, 4
    for f in ("D", "F", "G", "Sd_in", "So_in", "Sd_out", "So_out"):
        setattr(b, f, base)
    b.tri, b.grad_mode = 3, 1                                                                      # level 0 writes the caller's full gradient
    assert lib.crb200_level_bwd(_native.F32, 8, ctypes.byref(b), None) == _native.EINVAL
    s = _native.SweepFwdArgs()
    s.batch, s.n, s.nlevels, s.R, s.tri = 1, 100, 3, base, 1                                       # partial sweep: the rest system is handed out
    assert lib.crb200_sweep_fwd(_native.F32, 8, ctypes.byref(s), None) == _native.EINVAL
    t = _native.SweepBwdArgs()
    t.batch, t.n, t.nlevels, t.D, t.tri, t.top_Sd = 1, 100, 7, base, 1, base                       # a seeded descent has full blocks on top
    assert lib.crb200_sweep_bwd(_native.F32, 8, ctypes.byref(t), None) == _native.EINVAL


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful without a GPU")
def test_no_cpu_fallback():
    from cyclic_gps import cyclic_reduction as cr
    R = torch.eye(2, dtype=torch.float64).repeat(3, 1, 1)
    O = torch.zeros(2, 2, 2, dtype=torch.float64)
    with pytest.raises(RuntimeError, match="CUDA"):
        cr.decompose(R, O)
    with pytest.raises(RuntimeError, match="CUDA"):
        cr.mahal_and_det(R, O, torch.zeros(3, 2, dtype=torch.float64))


def test_helper_products_match_golden(golden):
    """UU_T / Ux / U_Tx / SigU / UtV_diags / interleave are device-agnostic tensor algebra;
    check them on the CPU against the reference's outputs (reference test :67-144)."""
    import numpy as np
    from cyclic_gps import cyclic_reduction as cr
    from helpers import assert_close
    g = golden["helpers"]
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    for p in g["cases"]:
        p = str(p)
        F, G, x, y, Sd, So = (t(g[p + k]) for k in ("F", "G", "x", "y", "Sd", "So"))
        a, b = cr.UU_T(F, G)
        assert_close(a, g[p + "UUT_d"], 1e-13)
        assert_close(b, g[p + "UUT_o"], 1e-13)
        assert_close(cr.Ux(F, G, x), g[p + "Ux"], 1e-13)
        assert_close(cr.U_Tx(F, G, y), g[p + "UTx"], 1e-13)
        a, b = cr.SigU(Sd, So, F, G)
        assert_close(a, g[p + "SigU_d"], 1e-13)
        assert_close(b, g[p + "SigU_o"], 1e-13)
        assert_close(cr.UtV_diags(F, G, a, b), g[p + "UtV"], 1e-13)
    for key in g.files:
        if key.startswith("il_") and key.endswith("_out"):
            base = key[:-4]
            assert np.array_equal(cr.interleave(t(g[base + "_a"]), t(g[base + "_b"])).numpy(), g[key])
    # the star-import contract of the reference tests: np and torch come along
    ns = {}
    exec("from cyclic_gps.cyclic_reduction import *", ns)
    for name in ("np", "torch", "decompose", "decompose_step", "halfsolve", "backhalfsolve", "solve", "det", "mahal",
                 "mahal_and_det", "inverse_blocks", "UU_T", "Ux", "U_Tx", "SigU", "UtV_diags", "interleave", "JITTER"):
        assert name in ns, name
