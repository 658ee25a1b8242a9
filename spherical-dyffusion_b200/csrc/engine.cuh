// Engine dispatch for the batched-GEMM ops: bf16 ops run on the tcgen05/TMA engine (kind::f16) when the problem meets
// its alignment rules (gemm_tc.cuh), fp32 ops on the same engine as TF32 (kind::tf32) inside a Tf32Scope
// (SFNO_PREC_TF32) and on the CUDA-core FMA engine otherwise (SFNO_PREC_F32, parity mode; also the path of shapes
// that violate TMA alignment).  All are device paths of this library; there is no host path.
#pragma once
#include <type_traits>

#include "common.cuh"
#include "gemm_simt.cuh"
#include "gemm_tc.cuh"
#include "gemm_tc_ops.cuh"

namespace sfno {

// runtime switch (sfno_b200_set_option("force_simt", 1)) used by the tests to cross-check the engines
extern std::atomic<int> g_force_simt;

// fp32-storage ops take the tensor-core engine (kind::tf32) only inside a Tf32Scope: a net / plan created with
// SFNO_PREC_TF32 opens one around its launches; SFNO_PREC_F32 keeps the CUDA-core FMA engine (parity mode).
inline int& tf32_depth() {
  static thread_local int depth = 0;
  return depth;
}
struct Tf32Scope {
  bool on;
  explicit Tf32Scope(bool enable) : on(enable) { if (on) ++tf32_depth(); }
  ~Tf32Scope() { if (on) --tf32_depth(); }
  Tf32Scope(const Tf32Scope&) = delete;
  Tf32Scope& operator=(const Tf32Scope&) = delete;
};
template <class T>
inline bool tc_allowed() {
  if (g_force_simt.load(std::memory_order_relaxed)) return false;
  return std::is_same<T, bf16>::value || tf32_depth() > 0;
}

template <class Op>
int launch_gemm(const Op& op, cudaStream_t stream, const char* what) {
  if constexpr (TcTraits<Op>::kAvailable) {
    if (tc_allowed<typename Op::InT>() && TcTraits<Op>::eligible(op)) return launch_gemm_tc(op, stream, what);
  }
  return launch_gemm_simt(op, stream, what);
}

// 1x1 convolutions and the inverse DFT carry an activation / dropout in their epilogue: the tensor-core engine gets
// instantiations with those fixed at compile time (no per-element branches), everything else takes the generic op.
// true when launch_conv / launch_idft will take the tensor-core engine (which can fuse InstanceNorm statistics)
template <class T, class TOut>
bool conv_uses_tc(const ConvArgs<T, TOut>& a) {
  using Gen = OpConv<T, TOut, -1, -1>;
  return tc_allowed<T>() && TcTraits<Gen>::eligible(Gen(a));
}
template <class T, class TOut>
bool idft_uses_tc(const IdftArgs<T, TOut>& a) {
  using Gen = OpIdft<T, TOut, -1>;
  return tc_allowed<T>() && TcTraits<Gen>::eligible(Gen(a));
}
// statistics partials per row: one per 64-column slice of every N tile
constexpr int kConvBN = SFNO_TC_CONV_BN, kIdftBN = 192;
inline int conv_stat_slices(int64_t hw) { return (int)ceil_div64(hw, kConvBN) * (kConvBN / TC_SLICE_COLS); }
inline int idft_stat_slices(int nlon) { return ceil_div(nlon, kIdftBN) * (kIdftBN / TC_SLICE_COLS); }

template <class T, class TOut>
int launch_conv(const ConvArgs<T, TOut>& a, cudaStream_t stream, const char* what) {
  using Gen = OpConv<T, TOut, -1, -1>;
  const Gen gen(a);
  if (tc_allowed<T>() && TcTraits<Gen>::eligible(gen)) {
    if (a.drop_p == 0.0f) {
      if (a.act == SFNO_ACT_GELU) return launch_gemm_tc(OpConv<T, TOut, SFNO_ACT_GELU, 0>(a), stream, what);
      if (a.act == SFNO_ACT_NONE) return launch_gemm_tc(OpConv<T, TOut, SFNO_ACT_NONE, 0>(a), stream, what);
    } else if constexpr (sizeof(T) == sizeof(TOut)) {   // the MLP with inference dropout and precomputed keep masks
      if (a.drop_mask && a.act == SFNO_ACT_GELU) return launch_gemm_tc(OpConv<T, TOut, SFNO_ACT_GELU, 2>(a), stream, what);
      if (a.drop_mask && a.act == SFNO_ACT_NONE) return launch_gemm_tc(OpConv<T, TOut, SFNO_ACT_NONE, 2>(a), stream, what);
    }
    return launch_gemm_tc(gen, stream, what);
  }
  return launch_gemm_simt(gen, stream, what);
}

template <class T, class TOut>
int launch_idft(const IdftArgs<T, TOut>& a, cudaStream_t stream, const char* what) {
  using Gen = OpIdft<T, TOut, -1>;
  const Gen gen(a);
  if (tc_allowed<T>() && TcTraits<Gen>::eligible(gen)) {
    if (a.act == SFNO_ACT_GELU) return launch_gemm_tc(OpIdft<T, TOut, SFNO_ACT_GELU>(a), stream, what);
    if (a.act == SFNO_ACT_NONE) return launch_gemm_tc(OpIdft<T, TOut, SFNO_ACT_NONE>(a), stream, what);
    return launch_gemm_tc(gen, stream, what);
  }
  return launch_gemm_simt(gen, stream, what);
}

}  // namespace sfno
