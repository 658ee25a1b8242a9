// Engine dispatch for the batched-GEMM ops: fp32 ops run on the CUDA-core engine, bf16 ops on the
// tcgen05/TMA engine when the problem meets its alignment rules (gemm_tc.cuh), otherwise on the
// CUDA-core engine with bf16 storage.  Both are device paths of this library; there is no host path.
#pragma once
#include "common.cuh"
#include "gemm_simt.cuh"
#include "gemm_tc.cuh"
#include "gemm_tc_ops.cuh"

namespace sfno {

// runtime switch (sfno_b200_set_option("force_simt", 1)) used by the tests to cross-check the engines
extern std::atomic<int> g_force_simt;

template <class Op>
int launch_gemm(const Op& op, cudaStream_t stream, const char* what) {
  if constexpr (TcTraits<Op>::kAvailable) {
    if (!g_force_simt.load(std::memory_order_relaxed) && TcTraits<Op>::eligible(op)) return launch_gemm_tc(op, stream, what);
  }
  return launch_gemm_simt(op, stream, what);
}

}  // namespace sfno
