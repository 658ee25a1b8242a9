// fp32 CUDA-core engine for the batched-GEMM ops of ops.cuh (SFNO_PREC_F32: parity mode).
// Classic 128x128x16 block tile, 256 threads, 8x8 register micro-tile, FFMA accumulate in fp32 --
// the same arithmetic class as the reference's fp32 einsum/conv path.  Thread columns are strided by 16
// so that for a fixed (i,j) sixteen lanes write sixteen consecutive N indices (the contiguous
// index of every op's output).
#pragma once
#include "common.cuh"

namespace sfno {

constexpr int SIMT_BM = 128, SIMT_BN = 128, SIMT_BK = 16, SIMT_THREADS = 256;

// per-group end of the contraction range: ops that define k_end(g) contract over [k_begin(g), k_end(g)), the others to op.K
template <class Op>
__device__ __forceinline__ auto simt_k_end(const Op& op, int g, int) -> decltype(op.k_end(g)) { return op.k_end(g); }
template <class Op>
__device__ __forceinline__ int simt_k_end(const Op& op, int, long) { return op.K; }

template <class Op>
__global__ void __launch_bounds__(SIMT_THREADS) gemm_simt_kernel(const Op op) {
  __shared__ float As[SIMT_BK][SIMT_BM + 4];
  __shared__ float Bs[SIMT_BK][SIMT_BN + 4];
  const int g = blockIdx.z;
  const int m0 = blockIdx.x * SIMT_BM, n0 = blockIdx.y * SIMT_BN;
  const int t = threadIdx.x;
  const int M = op.M, K = simt_k_end(op, g, 0);
  const int N = op.n_end(g), n_lo = op.n_begin(g);   // per-group column range actually produced
  const int m_hi = op.m_end(g), m_lo = op.m_begin(g);  // per-group row range actually produced
  if (n0 >= N || n0 + SIMT_BN <= n_lo || m0 >= m_hi || m0 + SIMT_BM <= m_lo) return;  // block-uniform
  const auto* __restrict__ A = op.A;
  const auto* __restrict__ Bm = op.Bm;

  // ---- loader bookkeeping: row offsets are hoisted out of the k loop --------------------------------
  int64_t a_base[8], b_base[8];
  int a_mm[8], b_nn[8], a_kk, b_kk;
  if (Op::A_KCONTIG) {
    a_kk = t % SIMT_BK;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      a_mm[i] = t / SIMT_BK + 16 * i;
      int m = m0 + a_mm[i];
      a_base[i] = m < M ? op.a_off(g, m) : -1;
    }
  } else {
    a_kk = t / SIMT_BM;  // 0..1, + 2*i
    a_mm[0] = t % SIMT_BM;
    int m = m0 + a_mm[0];
    a_base[0] = m < M ? op.a_off(g, m) : -1;
  }
  if (Op::B_KCONTIG) {
    b_kk = t % SIMT_BK;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      b_nn[i] = t / SIMT_BK + 16 * i;
      int n = n0 + b_nn[i];
      b_base[i] = n < N ? op.b_off(g, n) : -1;
    }
  } else {
    b_kk = t / SIMT_BN;
    b_nn[0] = t % SIMT_BN;
    int n = n0 + b_nn[0];
    b_base[0] = n < N ? op.b_off(g, n) : -1;
  }

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;

  // which 16 lanes walk the contiguous output index n: every op writes along n, but the pixel-sized N of the 1x1
  // convolutions measured faster with rows on the fast lanes (fc1 43.7 vs 69.6 ms in fp32 at ACE size)
  const int tx = Op::kSimtRowsOnFastLanes ? t / 16 : t % 16, ty = Op::kSimtRowsOnFastLanes ? t % 16 : t / 16;

  for (int k0 = op.k_begin(g); k0 < K; k0 += SIMT_BK) {
    if (Op::A_KCONTIG) {
      const int k = k0 + a_kk;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        As[a_kk][a_mm[i]] = (a_base[i] >= 0 && k < K) ? to_f32(A[a_base[i] + (int64_t)k * op.a_sk]) : 0.0f;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int kk = a_kk + 2 * i, k = k0 + kk;
        As[kk][a_mm[0]] = (a_base[0] >= 0 && k < K) ? to_f32(A[a_base[0] + (int64_t)k * op.a_sk]) : 0.0f;
      }
    }
    if (Op::B_KCONTIG) {
      const int k = k0 + b_kk;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        Bs[b_kk][b_nn[i]] = (b_base[i] >= 0 && k < K) ? to_f32(Bm[b_base[i] + (int64_t)k * op.b_sk]) : 0.0f;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int kk = b_kk + 2 * i, k = k0 + kk;
        Bs[kk][b_nn[0]] = (b_base[0] >= 0 && k < K) ? to_f32(Bm[b_base[0] + (int64_t)k * op.b_sk]) : 0.0f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SIMT_BK; ++kk) {
      float a[8], b[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = As[kk][ty + 16 * i];
#pragma unroll
      for (int j = 0; j < 8; ++j) b[j] = Bs[kk][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + ty + 16 * i;
    if (m >= m_hi || m < m_lo) continue;
    const typename Op::Row r = op.row(g, m);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + tx + 16 * j;
      if (n < N && n >= n_lo) op.store(r, g, m, n, acc[i][j]);
    }
  }
}

template <class Op>
int launch_gemm_simt(const Op& op, cudaStream_t stream, const char* what) {
  if (op.M <= 0 || op.N <= 0 || op.G <= 0) return SFNO_OK;
  dim3 grid(ceil_div(op.M, SIMT_BM), ceil_div(op.N, SIMT_BN), op.G);
  if (grid.y > 65535 || grid.z > 65535) return fail(SFNO_ERR_UNSUPPORTED, "%s: grid too large (%u,%u,%u)", what, grid.x, grid.y, grid.z);
  gemm_simt_kernel<Op><<<grid, SIMT_THREADS, 0, stream>>>(op);
  return post_launch(what);
}

}  // namespace sfno
