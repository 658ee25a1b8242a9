// GEMM "ops": every contraction on the SFNO hot path expressed as a batched GEMM
//     D[g][m][n] = sum_k A[g](m,k) * B[g](n,k)          (fp32 accumulate)
// plus a fused epilogue.  An op describes operand addressing (row offset + k stride), the problem
// size and the epilogue; the engines (gemm_simt.cuh: fp32 CUDA cores; gemm_tc.cuh: tcgen05 + TMA)
// are templated on the op.  Orientation of each GEMM is chosen so that the M index (TMEM lane /
// thread row) is the contiguous index of the OUTPUT tensor -> coalesced stores without staging.
//
// Internal tensor layouts (T = float or bf16; Kp = nlat rounded up to 8; pads are kept zero):
//   grid   x  [B][C][nlat][nlon]                      (NCHW, as the reference)
//   F, G      [mmax][B][2][C][Kp]   /  [mmax][2][B][C][Kp]     longitude-spectral, latitude contiguous
//   X, Y      [lmax][mmax][B][2][C] /  [mmax][lmax][B][2][C]   spectral, channel contiguous
#pragma once
#include "common.cuh"

namespace sfno {

// ------------------------------------------------------------------------------------------------
// forward longitude DFT (K1 of SURVEY 2.3): rows (b,c,k) x nlon -> F, fused InstanceNorm/time affine:
//   F = a[b,c] * DFT(x) + d[b,c] * 2*pi * [m==0, re]       (DFT is linear; DFT(1) = 2*pi*delta_m0)
// ------------------------------------------------------------------------------------------------
template <class T>
struct OpDft {
  static constexpr bool A_KCONTIG = true, B_KCONTIG = true;
  int G, M, N, K;  // G = B, M = C*nlat (rows (c,k) of one sample), N = 2*mmax, K = nlon
  const T* A; const T* Bm; int64_t a_sk, b_sk;
  T* f;
  const float* aff_a; const float* aff_d;  // [B*C] or nullptr
  int B, C, nlat, nlon, Kp, Wp;
  int64_t x_bstride;

  __device__ int64_t a_off(int g, int m) const { return (int64_t)g * x_bstride + (int64_t)m * nlon; }
  __device__ int64_t b_off(int, int n) const { return (int64_t)n * Wp; }
  struct Row { int64_t base; float a, d; };
  __device__ Row row(int g, int m) const {
    int c = m / nlat, k = m - c * nlat;
    Row r;
    r.base = (int64_t)g * 2 * C * Kp + (int64_t)c * Kp + k;
    r.a = aff_a ? aff_a[g * C + c] : 1.0f;
    r.d = aff_d ? aff_d[g * C + c] * 6.28318530717958647692f : 0.0f;
    return r;
  }
  __device__ void store(const Row& r, int, int, int n, float acc) const {
    int mm = n >> 1, ri = n & 1;
    float v = r.a * acc + (n == 0 ? r.d : 0.0f);
    f[(int64_t)mm * B * 2 * C * Kp + (int64_t)ri * C * Kp + r.base] = from_f32<T>(v);
  }
};

// ------------------------------------------------------------------------------------------------
// forward Legendre (K2): per m, rows (b,ri,c) x nlat -> X[l][m][b][ri][c]
// ------------------------------------------------------------------------------------------------
template <class T>
struct OpLeg {
  static constexpr bool A_KCONTIG = true, B_KCONTIG = true;
  int G, M, N, K;  // G = mmax, M = B*2*C, N = lmax, K = nlat
  const T* A; const T* Bm; int64_t a_sk, b_sk;
  T* x;
  int Kp, lmax, mmax;
  __device__ int64_t a_off(int g, int m) const { return ((int64_t)g * M + m) * Kp; }
  __device__ int64_t b_off(int g, int n) const { return ((int64_t)g * lmax + n) * Kp; }
  struct Row { int64_t base; };
  __device__ Row row(int g, int m) const { return Row{(int64_t)g * M + m}; }
  __device__ void store(const Row& r, int, int, int n, float acc) const {
    x[(int64_t)n * mmax * M + r.base] = from_f32<T>(acc);
  }
};

// ------------------------------------------------------------------------------------------------
// dhconv channel contraction (K3) as a REAL GEMM on the packed complex weight, per degree l:
//   D[(ri',o), (m,b)] = sum_(ri,c) Wp[l][(ri',o)][(ri,c)] * X[l][m][b][(ri,c)]  -> Y[m][l][b][ri'][o]
//   Wp = [[wr, -wi], [wi, wr]] (rows = output re/im, cols = input re/im)
// ------------------------------------------------------------------------------------------------
template <class T>
struct OpDhconv {
  static constexpr bool A_KCONTIG = true, B_KCONTIG = true;
  int G, M, N, K;  // G = lmax, M = 2*Cout, N = mmax*B, K = 2*Cin
  const T* A; const T* Bm; int64_t a_sk, b_sk;
  T* y;
  int B, lmax, mmax;
  __device__ int64_t a_off(int g, int m) const { return ((int64_t)g * M + m) * K; }
  __device__ int64_t b_off(int g, int n) const { return ((int64_t)g * N + n) * K; }
  struct Row { int64_t base; };
  __device__ Row row(int g, int m) const { return Row{(int64_t)g * B * M + m}; }
  __device__ void store(const Row& r, int, int, int n, float acc) const {
    int mm = n / B, b = n - mm * B;
    y[(int64_t)mm * lmax * B * M + (int64_t)b * M + r.base] = from_f32<T>(acc);
  }
};

// ------------------------------------------------------------------------------------------------
// inverse Legendre (K4), flipped so latitude k is the M index: per m,
//   D[k, (b,ri,o)] = sum_l Pt[m][k][l] * Y[m][l][(b,ri,o)]  -> G[m][ri][b][o][k]
// ------------------------------------------------------------------------------------------------
template <class T>
struct OpIleg {
  static constexpr bool A_KCONTIG = true, B_KCONTIG = false;
  int G, M, N, K;  // G = mmax, M = nlat, N = B*2*C, K = lmax
  const T* A; const T* Bm; int64_t a_sk, b_sk;
  T* g_out;
  int B, C, Kp, Lq, nlat;
  int64_t b_goff;  // Y layout [m][l][n]: b_goff = lmax*N, b_sk = N;  X layout [l][m][n]: b_goff = N, b_sk = mmax*N
  __device__ int64_t a_off(int g, int m) const { return ((int64_t)g * nlat + m) * Lq; }
  __device__ int64_t b_off(int g, int n) const { return (int64_t)g * b_goff + n; }  // + l * b_sk
  struct Row { int64_t base; };
  __device__ Row row(int g, int m) const { return Row{(int64_t)g * 2 * B * C * Kp + m}; }
  __device__ void store(const Row& r, int, int, int n, float acc) const {
    int b = n / (2 * C), rem = n - b * 2 * C;
    int ri = rem / C, o = rem - ri * C;
    g_out[r.base + (int64_t)ri * B * C * Kp + ((int64_t)b * C + o) * Kp] = from_f32<T>(acc);
  }
};

// ------------------------------------------------------------------------------------------------
// inverse longitude DFT (K5), flipped so longitude j is the M index:
//   D[j, (b,o,kp)] = sum_(m,ri) Einv[j][(m,ri)] * G[(m,ri)][(b,o,kp)]
//   epilogue: + bias[o] + add[b][o][k][j] -> act -> out[b][o][k][j]   (bias of SpectralConvS2,
//   inner-skip sum and GELU of FourierNeuralOperatorBlock.forward fused: sfnonet.py:308-311)
// ------------------------------------------------------------------------------------------------
template <class T, class TOut>
struct OpIdft {
  static constexpr bool A_KCONTIG = true, B_KCONTIG = false;
  int G, M, N, K;  // G = 1, M = nlon, N = B*C*Kp, K = 2*mmax
  const T* A; const T* Bm; int64_t a_sk, b_sk;
  TOut* out; int64_t out_bstride;
  const float* bias;                      // [C] or nullptr
  const T* add; int64_t add_bstride;      // [B][C][nlat][nlon] or nullptr
  int act;
  int C, nlat, nlon, Kp, Kq2;
  __device__ int64_t a_off(int, int m) const { return (int64_t)m * Kq2; }
  __device__ int64_t b_off(int, int n) const { return n; }  // + kk * N
  struct Row { int j; };
  __device__ Row row(int, int m) const { return Row{m}; }
  __device__ void store(const Row& r, int, int, int n, float acc) const {
    int bo = n / Kp, k = n - bo * Kp;
    if (k >= nlat) return;
    int b = bo / C, o = bo - b * C;
    int64_t pix = ((int64_t)o * nlat + k) * nlon + r.j;
    float v = acc + (bias ? bias[o] : 0.0f);
    if (add) v += to_f32(add[(int64_t)b * add_bstride + pix]);
    v = apply_act(act, v);
    out[(int64_t)b * out_bstride + pix] = from_f32<TOut>(v);
  }
};

// ------------------------------------------------------------------------------------------------
// 1x1 convolution (K6) per sample: D[p, o] = sum_c in[b][c][p] * w[(b)][o][c], pixel index p is M.
// epilogue (all optional): + bias[(b)][o] -> act -> dropout -> * branch_scale[b]
//                          + residual (optionally affine: ra[b,o]*res + rd[b,o]) + pos[o][p]
// ------------------------------------------------------------------------------------------------
template <class T, class TOut>
struct OpConv {
  static constexpr bool A_KCONTIG = false, B_KCONTIG = true;
  int G, M, N, K;  // G = batch, M = hw, N = cout, K = cin
  const T* A; const T* Bm; int64_t a_sk, b_sk;
  int64_t in_bstride;
  int64_t w_bstride; int ldw;            // w_bstride = 0 for shared weights
  const float* bias; int64_t bias_bstride;
  int act;
  float drop_p; uint64_t seed, offset;    // dropout (after act), drop_p = 0 -> off
  const float* branch_scale;              // [batch] DropPath factor (0 or 1/keep) or nullptr
  const T* res; int64_t res_bstride;      // residual [b][cout][hw] or nullptr
  const float* res_a; const float* res_d; // [batch*cout] affine on the residual or nullptr
  const T* pos;                           // [cout][hw] or nullptr
  TOut* out; int64_t out_bstride;
  __device__ int64_t a_off(int g, int m) const { return (int64_t)g * in_bstride + m; }  // + c * hw
  __device__ int64_t b_off(int g, int n) const { return (int64_t)g * w_bstride + (int64_t)n * ldw; }
  struct Row { int p; };
  __device__ Row row(int, int m) const { return Row{m}; }
  __device__ void store(const Row& r, int g, int, int n, float acc) const {
    float v = acc + (bias ? bias[(int64_t)g * bias_bstride + n] : 0.0f);
    v = apply_act(act, v);
    int64_t pix = (int64_t)n * M + r.p;
    if (drop_p > 0.0f) {
      float u = philox_uniform(seed, offset, (uint64_t)g * N * M + pix);
      v = (u >= drop_p) ? v * (1.0f / (1.0f - drop_p)) : 0.0f;
    }
    if (branch_scale) v *= branch_scale[g];
    if (res) {
      float rv = to_f32(res[(int64_t)g * res_bstride + pix]);
      if (res_a) rv = res_a[g * N + n] * rv + res_d[g * N + n];
      v += rv;
    }
    if (pos) v += to_f32(pos[pix]);
    out[(int64_t)g * out_bstride + pix] = from_f32<TOut>(v);
  }
};

}  // namespace sfno
