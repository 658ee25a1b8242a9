// Shared helpers for libsfno_b200: status/error plumbing, launch accounting, dtype conversion.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <nvtx3/nvToolsExt.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <string>

#include "../../include/sfno_b200.h"

namespace sfno {

// ---- error text (thread local) -------------------------------------------------------------------
std::string& last_error_ref();
int fail(int status, const char* fmt, ...);

#define SFNO_CHECK_ARG(cond, ...)                                        \
  do {                                                                   \
    if (!(cond)) return ::sfno::fail(SFNO_ERR_INVALID_ARGUMENT, __VA_ARGS__); \
  } while (0)

#define SFNO_CUDA(expr)                                                                          \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess)                                                                       \
      return ::sfno::fail(SFNO_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                          __FILE__, __LINE__);                                                   \
  } while (0)

#define SFNO_TRY(expr)         \
  do {                         \
    int _s = (expr);           \
    if (_s != SFNO_OK) return _s; \
  } while (0)

// ---- launch accounting + optional per-launch timing (sfno_b200_profile_begin/end) ---------------------
extern std::atomic<int64_t> g_launch_count;
extern std::atomic<int> g_profile_on;
void profile_mark(const char* what);  // records a CUDA event on the profiled stream after a launch
inline int post_launch(const char* what) {
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  if (g_profile_on.load(std::memory_order_relaxed)) profile_mark(what);
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(SFNO_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
  }
  return SFNO_OK;
}

// ---- optional NVTX ranges (sfno_b200_set_option("nvtx", 1) or SFNO_NVTX=1 in the environment): forward, blocks, op-level
// entry points show up as named ranges on the timeline of a tracing tool; off by default (no cost beyond one atomic load)
extern std::atomic<int> g_nvtx_on;
struct NvtxRange {
  bool on;
  explicit NvtxRange(const char* name) : on(g_nvtx_on.load(std::memory_order_relaxed) != 0) { if (on) nvtxRangePushA(name); }
  ~NvtxRange() { if (on) nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};

// ---- small math ---------------------------------------------------------------------------------------
__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline int round_up(int a, int b) { return (a + b - 1) / b * b; }
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---- storage types --------------------------------------------------------------------------------------
using bf16 = __nv_bfloat16;

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(bf16 v) { return __bfloat162float(v); }
template <class T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }

// fp32 -> nearest TF32 value (10-bit mantissa, ties away from zero), kept in fp32 storage
__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

__device__ __forceinline__ float gelu_exact(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float silu(float x) { return x / (1.0f + __expf(-x)); }
__device__ __forceinline__ float apply_act(int act, float x) {
  switch (act) {
    case SFNO_ACT_GELU: return gelu_exact(x);
    case SFNO_ACT_RELU: return fmaxf(x, 0.0f);
    case SFNO_ACT_SILU: return x / (1.0f + expf(-x));
    default: return x;
  }
}

// ---- Philox4x32-10 counter RNG (dropout masks; statistical parity only with torch's nn.Dropout) --------
template <int kRounds = 10>
__device__ __forceinline__ uint4 philox4x32(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < kRounds; ++r) {
    uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}
// uniform in [0,1) for element index `idx` of stream (seed, offset)
__device__ __forceinline__ float philox_uniform(uint64_t seed, uint64_t offset, uint64_t idx) {
  uint4 c = make_uint4((uint32_t)(idx >> 2), (uint32_t)(idx >> 34), (uint32_t)offset, (uint32_t)(offset >> 32));
  uint4 r = philox4x32<10>(c, make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  uint32_t v = (idx & 3) == 0 ? r.x : (idx & 3) == 1 ? r.y : (idx & 3) == 2 ? r.z : r.w;
  return (v >> 8) * (1.0f / 16777216.0f);
}

// Packed fp32 pairs (Blackwell FADD2 / FMUL2 / FFMA2: two IEEE fp32 operations per issued instruction).  The epilogues
// of the tensor-core engine are bound by instruction issue, not by the fp32 pipe, so the element-wise math runs on pairs.
struct f2 { unsigned long long v; };
__device__ __forceinline__ f2 f2_make(float lo, float hi) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ f2 f2_splat(float a) { return f2_make(a, a); }
__device__ __forceinline__ void f2_get(f2 a, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v)); }
__device__ __forceinline__ f2 f2_add(f2 a, f2 b) { f2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ f2 f2_mul(f2 a, f2 b) { f2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ f2 f2_fma(f2 a, f2 b, f2 c) { f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }
__device__ __forceinline__ float f2_hsum(f2 a) { float lo, hi; f2_get(a, lo, hi); return lo + hi; }

// Dropout decisions: one Philox4x32-7 block (the smallest round count of the family that passes BigCrush; masks need
// no more, and the rounds are the cost of the dropout epilogues) serves EIGHT consecutive elements (16 random bits
// each, idx8 % 8 == 0): element idx8 + j is kept iff its 16-bit value >= thr16 = round(p * 65536).  Bit j = keep.
__device__ __forceinline__ uint32_t dropout_threshold16(float p) { return (uint32_t)(p * 65536.0f + 0.5f); }
__device__ __forceinline__ float dropout_keep_scale(uint32_t thr16) { return 65536.0f / (float)(65536u - thr16); }
__device__ __forceinline__ uint32_t dropout_keep_mask8(uint64_t seed, uint64_t offset, uint64_t idx8, uint32_t thr16) {
  const uint64_t blk = idx8 >> 3;
  const uint4 c = make_uint4((uint32_t)blk, (uint32_t)(blk >> 32), (uint32_t)offset, (uint32_t)(offset >> 32));
  const uint4 r = philox4x32<7>(c, make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  uint32_t m = 0;
  m |= ((r.x & 0xffffu) >= thr16) ? 1u : 0u;   m |= ((r.x >> 16) >= thr16) ? 2u : 0u;
  m |= ((r.y & 0xffffu) >= thr16) ? 4u : 0u;   m |= ((r.y >> 16) >= thr16) ? 8u : 0u;
  m |= ((r.z & 0xffffu) >= thr16) ? 16u : 0u;  m |= ((r.z >> 16) >= thr16) ? 32u : 0u;
  m |= ((r.w & 0xffffu) >= thr16) ? 64u : 0u;  m |= ((r.w >> 16) >= thr16) ? 128u : 0u;
  return m;
}
__device__ __forceinline__ bool dropout_keep(uint64_t seed, uint64_t offset, uint64_t idx, uint32_t thr16) {
  return (dropout_keep_mask8(seed, offset, idx & ~7ull, thr16) >> (uint32_t)(idx & 7ull)) & 1u;
}

// Optional epilogue features.  The tensor-core engine instantiates its drain loop for the op's two most frequent
// feature sets (kFast0, kFast1: compile-time masks, no per-element tests) plus, if kGeneral, a run-time tested one (F < 0).
enum : int { F_RES = 1, F_STATS = 2, F_POS = 4, F_SCALE = 8 };
template <int F, int BIT>
__device__ __forceinline__ bool feat_on(bool runtime_value) {
  if constexpr (F < 0) return runtime_value;
  else return (F & BIT) != 0;
}

}  // namespace sfno
