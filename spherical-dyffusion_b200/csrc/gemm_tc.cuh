// tcgen05 / TMA tensor-core engine for the bf16 ops (placeholder traits; filled in by gemm_tc_impl.cuh).
#pragma once
#include "common.cuh"

namespace sfno {

template <class Op>
struct TcTraits {
  static constexpr bool kAvailable = false;
  static bool eligible(const Op&) { return false; }
};

template <class Op>
int launch_gemm_tc(const Op&, cudaStream_t, const char* what) {
  return fail(SFNO_ERR_UNSUPPORTED, "%s: tensor-core engine not built for this op", what);
}

}  // namespace sfno
