"""B200-native cyclic-reduction engine behind the reference's ``cyclic_gps`` Python API.

Only the hot path of cunningham-lab/cyclic-gps is provided natively: the functions of
``cyclic_gps.cyclic_reduction`` (hand-written sm_100a CUDA kernels through a C ABI), plus
the thin LEG-model glue the reference's own likelihood test imports."""
__all__ = ["cyclic_reduction"]
