// Engine dispatch for the batched-GEMM ops: fp32 ops run on the CUDA-core engine, bf16 ops on the
// tcgen05/TMA engine when the problem meets its alignment rules (gemm_tc.cuh), otherwise on the
// CUDA-core engine with bf16 storage.  Both are device paths of this library; there is no host path.
#pragma once
#include <type_traits>

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

// 1x1 convolutions and the inverse DFT carry an activation / dropout in their epilogue: the tensor-core engine gets
// instantiations with those fixed at compile time (no per-element branches), everything else takes the generic op.
// true when launch_conv / launch_idft will take the tensor-core engine (which can fuse InstanceNorm statistics)
template <class T, class TOut>
bool conv_uses_tc(const ConvArgs<T, TOut>& a) {
  if constexpr (std::is_same<T, bf16>::value) {
    using Gen = OpConv<T, TOut, -1, -1>;
    return !g_force_simt.load(std::memory_order_relaxed) && TcTraits<Gen>::eligible(Gen(a));
  }
  return false;
}
template <class T, class TOut>
bool idft_uses_tc(const IdftArgs<T, TOut>& a) {
  if constexpr (std::is_same<T, bf16>::value) {
    using Gen = OpIdft<T, TOut, -1>;
    return !g_force_simt.load(std::memory_order_relaxed) && TcTraits<Gen>::eligible(Gen(a));
  }
  return false;
}
// statistics partials per row: one per 64-column slice of every N tile
constexpr int kConvBN = 192, kIdftBN = 192;
inline int conv_stat_slices(int64_t hw) { return (int)ceil_div64(hw, kConvBN) * (kConvBN / TC_SLICE_COLS); }
inline int idft_stat_slices(int nlon) { return ceil_div(nlon, kIdftBN) * (kIdftBN / TC_SLICE_COLS); }

template <class T, class TOut>
int launch_conv(const ConvArgs<T, TOut>& a, cudaStream_t stream, const char* what) {
  using Gen = OpConv<T, TOut, -1, -1>;
  const Gen gen(a);
  if constexpr (std::is_same<T, bf16>::value) {
    if (!g_force_simt.load(std::memory_order_relaxed) && TcTraits<Gen>::eligible(gen)) {
      if (a.drop_p == 0.0f) {
        if (a.act == SFNO_ACT_GELU) return launch_gemm_tc(OpConv<T, TOut, SFNO_ACT_GELU, 0>(a), stream, what);
        if (a.act == SFNO_ACT_NONE) return launch_gemm_tc(OpConv<T, TOut, SFNO_ACT_NONE, 0>(a), stream, what);
      }
      return launch_gemm_tc(gen, stream, what);
    }
  }
  return launch_gemm_simt(gen, stream, what);
}

template <class T, class TOut>
int launch_idft(const IdftArgs<T, TOut>& a, cudaStream_t stream, const char* what) {
  using Gen = OpIdft<T, TOut, -1>;
  const Gen gen(a);
  if constexpr (std::is_same<T, bf16>::value) {
    if (!g_force_simt.load(std::memory_order_relaxed) && TcTraits<Gen>::eligible(gen)) {
      if (a.act == SFNO_ACT_GELU) return launch_gemm_tc(OpIdft<T, TOut, SFNO_ACT_GELU>(a), stream, what);
      if (a.act == SFNO_ACT_NONE) return launch_gemm_tc(OpIdft<T, TOut, SFNO_ACT_NONE>(a), stream, what);
      return launch_gemm_tc(gen, stream, what);
    }
  }
  return launch_gemm_simt(gen, stream, what);
}

}  // namespace sfno
