// Deep levels of a sweep fused into one launch (cr_tpn_*_multi_kernel): the per-level argument blocks travel as
// ONE kernel parameter, one CTA per series: kMultiWarps warps for small batches (levels with up to 2 * kMultiWarps tiles per
// series), ONE warp for large batches (only the levels that are a single tile per series: there a launch per level is pure
// latency, and a single-warp CTA costs a series no more than its own tile).  Shared by the kernels and the ABI (host code).
#pragma once

namespace crb200 {

constexpr int kMultiMax = 12;      // levels per fused launch (12 x ~220 B of arguments < the 4 KB parameter space)
constexpr int kMultiWarps = 4;

// Packed lower triangles (crb200_*_args.tri, crb200_tri_stride): elements per packed block where the thread-per-node kernels
// offer that storage, else 0.  float32, ell = 8: 36 of 64 elements, a multiple of 16 bytes like the full block.
constexpr int tri_stride_elems(int elem_size, int ell) { return (elem_size == 4 && ell == 8) ? 36 : 0; }

template <typename Args>
struct MultiArgs {
  int count;
  int warps;        // warps per CTA: kMultiWarps or 1
  Args lv[kMultiMax];
};

}  // namespace crb200
