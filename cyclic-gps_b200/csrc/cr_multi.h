// Deep levels of a sweep fused into one launch (cr_tpn_*_multi_kernel): the per-level argument blocks travel as
// ONE kernel parameter, one CTA of kMultiWarps warps per series.  Shared by the kernels and the ABI (host code).
#pragma once

namespace crb200 {

constexpr int kMultiMax = 12;      // levels per fused launch (12 x ~220 B of arguments < the 4 KB parameter space)
constexpr int kMultiWarps = 4;

template <typename Args>
struct MultiArgs {
  int count;
  int pad_;
  Args lv[kMultiMax];
};

}  // namespace crb200
