// Backward kernels of the op-level entry points ("autograd formulas of the custom ops"): together with the
// adjoint transforms (lib_core.cu: sfno_sht_forward_adjoint / sfno_sht_inverse_adjoint, which reuse the forward GEMM ops on
// transposed tables) they let the reference's modules -- or this package's trainable forward -- backpropagate through
// native kernels.  fp32 CUDA-core arithmetic; these are functional, not yet tuned, kernels.
#include "backward.cuh"
#include "common.cuh"
#include "engine.cuh"
#include "ops.cuh"
#include "pointwise.cuh"

namespace sfno {

// D[g][m][n] = sum_k A(g, m, k) * B(g, n, k), both operands K-contiguous, groups = (sample, K chunk): the weight gradient
// of a 1x1 convolution, gw[o][c] = sum_{b,p} gy[b][o][p] x[b][c][p], as split-K partial GEMMs.
struct OpWgrad {
  static constexpr bool A_KCONTIG = true, B_KCONTIG = true, kSimtRowsOnFastLanes = false;
  int G, M, N, K;              // G = batch * splits, M = cout, N = cin, K = pixels per chunk
  const float* A; const float* Bm; int64_t a_sk, b_sk;
  int splits; int64_t hw;
  float* part;                 // [G][M][N]
  __device__ int n_begin(int) const { return 0; }
  __device__ int n_end(int) const { return N; }
  __device__ int m_begin(int) const { return 0; }
  __device__ int m_end(int) const { return M; }
  __device__ int k_begin(int) const { return 0; }
  __device__ int64_t a_off(int g, int m) const { const int b = g / splits, s = g - b * splits; return ((int64_t)b * M + m) * hw + (int64_t)s * K; }
  __device__ int64_t b_off(int g, int n) const { const int b = g / splits, s = g - b * splits; return ((int64_t)b * N + n) * hw + (int64_t)s * K; }
  struct Row { float* out; };
  __device__ Row row(int g, int m) const { return Row{part + ((int64_t)g * M + m) * N}; }
  __device__ void store(const Row& r, int, int, int n, float acc) const { r.out[n] = acc; }
};

// The same split-K weight gradient for the tensor-core engine: operands in T (bf16 / TF32-exact fp32), K-contiguous,
// partial sums in fp32.  The K chunk of a group is a TMA dimension of the operand views ({pixel in chunk, row, chunk,
// sample}, group g -> {g % splits, g / splits}), so every group contracts over [0, K) like any other batched GEMM.
template <class T>
struct OpWgradTc : NoFeatures {
  static constexpr bool kGFastest = false;
  static constexpr bool kSimtRowsOnFastLanes = false;
  static constexpr bool kRanged = true;
  static constexpr bool A_KCONTIG = true, B_KCONTIG = true, kColContig = true, kNFastest = false;
  using OutT = float;
  using InT = T;
  __device__ bool out_tf32() const { return false; }
  __device__ int m_begin(int) const { return 0; }
  __device__ int m_end(int) const { return M; }
  __device__ int n_begin(int) const { return 0; }
  __device__ int n_end(int) const { return N; }
  __device__ int k_begin(int) const { return 0; }
  int G, M, N, K;              // G = batch * splits, M = cout, N = cin, K = pixels per chunk
  const T* A; const T* Bm; int64_t a_sk, b_sk;   // A = gy [b][cout][hw], Bm = x [b][cin][hw]
  int splits; int64_t hw;
  float* part;                 // [G][M][N]
  __device__ int64_t a_off(int g, int m) const { const int b = g / splits, s = g - b * splits; return ((int64_t)b * M + m) * hw + (int64_t)s * K; }
  __device__ int64_t b_off(int g, int n) const { const int b = g / splits, s = g - b * splits; return ((int64_t)b * N + n) * hw + (int64_t)s * K; }
  __device__ int n_store() const { return N; }
  __device__ void io_coords(int g, int row0, int col0, int (&c)[5]) const { c[0] = col0; c[1] = row0; c[2] = g; c[3] = 0; c[4] = 0; }
  struct Row { float* out; const float* res; bool valid; __device__ float stat_s() const { return 0.0f; } __device__ float stat_q() const { return 0.0f; } };
  __device__ Row row(int g, int m) const { return Row{part + ((int64_t)g * M + m) * N, nullptr, true}; }
  template <int F> __device__ Row row_f(int g, int m) const { return row(g, m); }
  __device__ void store(const Row& r, int, int, int n, float acc) const { r.out[n] = acc; }
  template <int F>
  __device__ void compute8(Row&, int, const float (&acc)[8], const float (&)[8], float (&o)[8]) const {
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = acc[i];
  }
};

template <class T>
struct TcTraits<OpWgradTc<T>> : TcTraitsBase<OpWgradTc<T>>, TcEligible<TcTraits<OpWgradTc<T>>, OpWgradTc<T>> {
  static constexpr int BN = 256;
  static constexpr uint64_t es = sizeof(T);
  static void operands(const OpWgradTc<T>& op, TmaOperand& a, TmaOperand& b) {
    const uint64_t batches = (uint64_t)(op.G / op.splits);
    a.base = op.A; a.dims[0] = op.K; a.dims[1] = op.M; a.dims[2] = op.splits; a.dims[3] = batches;
    a.strides[0] = (uint64_t)op.hw * es; a.strides[1] = (uint64_t)op.K * es; a.strides[2] = (uint64_t)op.M * op.hw * es;
    a.batched = true; a.group_lo = op.splits;
    b.base = op.Bm; b.dims[0] = op.K; b.dims[1] = op.N; b.dims[2] = op.splits; b.dims[3] = batches;
    b.strides[0] = (uint64_t)op.hw * es; b.strides[1] = (uint64_t)op.K * es; b.strides[2] = (uint64_t)op.N * op.hw * es;
    b.batched = true; b.group_lo = op.splits;
  }
  static bool extra_ok(const OpWgradTc<T>& op) { return aligned16(op.part) && op.N % 8 == 0; }
  static void io(const OpWgradTc<T>& op, TmaIo& o, TmaIo&) {
    o.base = op.part; o.es = 4; o.ok = true;
    o.dims[0] = op.N; o.dims[1] = op.M; o.dims[2] = op.G;
    o.strides[0] = (uint64_t)op.N * 4; o.strides[1] = (uint64_t)op.M * op.N * 4;
    o.box_rows[0] = 32;
  }
};

__global__ void reduce_groups_kernel(const float* __restrict__ part, int G, int64_t n, float* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float s = 0.0f;
    for (int g = 0; g < G; ++g) s += part[(int64_t)g * n + i];
    out[i] = s;
  }
}

// gb[o] = sum_{b,p} gy[b][o][p]; one block of 1024 threads per output channel, 16-byte loads where the planes allow
__global__ void __launch_bounds__(1024) bias_grad_kernel(const float* __restrict__ gy, int B, int C, int64_t hw, float* __restrict__ gb) {
  __shared__ float sh[32];
  const int o = blockIdx.x;
  float s = 0.0f;
  const bool vec = (hw & 3) == 0 && (reinterpret_cast<uintptr_t>(gy) & 15) == 0;
  for (int b = 0; b < B; ++b) {
    const float* p = gy + ((int64_t)b * C + o) * hw;
    if (vec) {
      const float4* p4 = reinterpret_cast<const float4*>(p);
      float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
      for (int64_t i = threadIdx.x; i < (hw >> 2); i += blockDim.x) {
        const float4 v = p4[i];
        s0 += v.x; s1 += v.y; s2 += v.z; s3 += v.w;
      }
      s += (s0 + s1) + (s2 + s3);
    } else {
      for (int64_t i = threadIdx.x; i < hw; i += blockDim.x) s += p[i];
    }
  }
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0f;
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (threadIdx.x == 0) gb[o] = s;
  }
}

// y[b,o,l,m] = sum_i x[b,i,l,m] w[i,o,l(,m)] (complex).  With g = dL/d(re) + i dL/d(im) of a real loss:
//   gx[b,i,l,m] = sum_o gy[b,o,l,m] conj(w[i,o,l(,m)]),   gw[i,o,l(,m)] = sum_{b(,m)} conj(x[b,i,l,m]) gy[b,o,l,m]
__global__ void contract_grad_x_kernel(int diagonal, const float2* __restrict__ gy, const float2* __restrict__ w, float2* __restrict__ gx,
                                       int B, int Cin, int Cout, int L, int M) {
  const int64_t total = (int64_t)B * Cin * L * M;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int m = (int)(idx % M);
    int64_t r = idx / M;
    const int l = (int)(r % L); r /= L;
    const int i = (int)(r % Cin);
    const int b = (int)(r / Cin);
    float re = 0.0f, im = 0.0f;
    for (int o = 0; o < Cout; ++o) {
      const float2 g = gy[(((int64_t)b * Cout + o) * L + l) * M + m];
      const float2 wv = diagonal ? w[(((int64_t)i * Cout + o) * L + l) * M + m] : w[((int64_t)i * Cout + o) * L + l];
      re = fmaf(g.x, wv.x, re); re = fmaf(g.y, wv.y, re);     // g * conj(w)
      im = fmaf(g.y, wv.x, im); im = fmaf(-g.x, wv.y, im);
    }
    gx[idx] = make_float2(re, im);
  }
}

// one warp per weight entry (i, o, l) [dhconv: reduces over b and m] or (i, o, l, m) [diagonal: reduces over b]
__global__ void contract_grad_w_kernel(int diagonal, const float2* __restrict__ x, const float2* __restrict__ gy, float2* __restrict__ gw,
                                       int B, int Cin, int Cout, int L, int M) {
  const int64_t nw = (int64_t)Cin * Cout * L * (diagonal ? M : 1);
  const int lane = threadIdx.x & 31;
  for (int64_t wi = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; wi < nw; wi += ((int64_t)gridDim.x * blockDim.x) >> 5) {
    int64_t r = wi;
    int m0 = 0;
    if (diagonal) { m0 = (int)(r % M); r /= M; }
    const int l = (int)(r % L); r /= L;
    const int o = (int)(r % Cout);
    const int i = (int)(r / Cout);
    float re = 0.0f, im = 0.0f;
    const int span = diagonal ? 1 : M;
    for (int t = lane; t < B * span; t += 32) {
      const int b = t / span, m = diagonal ? m0 : t - b * span;
      const float2 xv = x[(((int64_t)b * Cin + i) * L + l) * M + m];
      const float2 g = gy[(((int64_t)b * Cout + o) * L + l) * M + m];
      re = fmaf(xv.x, g.x, re); re = fmaf(xv.y, g.y, re);     // conj(x) * g
      im = fmaf(xv.x, g.y, im); im = fmaf(-xv.y, g.x, im);
    }
    for (int off = 16; off > 0; off >>= 1) { re += __shfl_xor_sync(0xffffffffu, re, off); im += __shfl_xor_sync(0xffffffffu, im, off); }
    if (lane == 0) gw[wi] = make_float2(re, im);
  }
}

// InstanceNorm (+ per-(b,c) affine y = xhat * A + D) backward, one block per (b, c) plane:
//   dA = sum g xhat, dD = sum g, gx = A rstd (g - dD / n - xhat dA / n)
__global__ void __launch_bounds__(512) instance_norm_backward_kernel(const float* __restrict__ x, const float* __restrict__ gout,
                                                                     const float* __restrict__ Aff, int64_t hw, float eps,
                                                                     float* __restrict__ gx, float* __restrict__ dA, float* __restrict__ dD) {
  __shared__ double sh[4][16];
  const int bc = blockIdx.x;
  const float* xp = x + (int64_t)bc * hw;
  const float* gp = gout + (int64_t)bc * hw;
  auto block_sum2 = [&](double a, double b, double& oa, double& ob) {
    for (int off = 16; off > 0; off >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, off); b += __shfl_xor_sync(0xffffffffu, b, off); }
    if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = a; sh[1][threadIdx.x >> 5] = b; }
    __syncthreads();
    a = 0.0; b = 0.0;
    for (int wv = 0; wv < (int)(blockDim.x >> 5); ++wv) { a += sh[0][wv]; b += sh[1][wv]; }
    __syncthreads();
    oa = a; ob = b;
  };
  double s = 0.0, q = 0.0;
  for (int64_t i = threadIdx.x; i < hw; i += blockDim.x) { const double v = xp[i]; s += v; q += v * v; }
  double S, Q;
  block_sum2(s, q, S, Q);
  const double mu = S / (double)hw;
  const double var = fmax(Q / (double)hw - mu * mu, 0.0);
  const float rstd = (float)(1.0 / sqrt(var + (double)eps)), muf = (float)mu;
  double sg = 0.0, sgx = 0.0;
  for (int64_t i = threadIdx.x; i < hw; i += blockDim.x) { const double g = gp[i]; sg += g; sgx += g * (double)((xp[i] - muf) * rstd); }
  double SG, SGX;
  block_sum2(sg, sgx, SG, SGX);
  const float a = Aff ? Aff[bc] : 1.0f;
  const float mg = (float)(SG / (double)hw), mgx = (float)(SGX / (double)hw);
  for (int64_t i = threadIdx.x; i < hw; i += blockDim.x) {
    const float xh = (xp[i] - muf) * rstd;
    gx[(int64_t)bc * hw + i] = a * rstd * (gp[i] - mg - xh * mgx);
  }
  if (threadIdx.x == 0) { dA[bc] = (float)SGX; dD[bc] = (float)SG; }
}

int launch_bias_grad(const float* gy, int B, int C, int64_t hw, float* gb, cudaStream_t st) {
  bias_grad_kernel<<<C, 1024, 0, st>>>(gy, B, C, hw, gb);
  return post_launch("bias_grad");
}

static int pick_splits(int64_t hw) {   // a divisor of hw, <= 32, chunks of >= 512 pixels
  int best = 1;
  for (int s = 2; s <= 32; ++s)
    if (hw % s == 0 && hw / s >= 512) best = s;
  return best;
}

// K chunks for the tensor-core weight gradient: a divisor of hw whose chunk keeps the TMA strides 16-byte aligned, with
// enough (sample, chunk, tile) units to fill the machine
static int pick_splits_tc(int64_t hw, int batch, int cin, int cout) {
  const int tiles = ceil_div(cout, 128) * ceil_div(cin, 256);
  int best = 1;
  for (int s = 1; s <= 64; ++s) {
    if (hw % s != 0 || (hw / s) % 8 != 0 || hw / s < 512) continue;
    best = s;
    if ((int64_t)batch * s * tiles >= 2 * 148) break;
  }
  return best;
}

// [rows][cols] fp32 -> transposed [cols][ld] T, zero padded (weights of the data-gradient convolution)
template <class T>
static __global__ void pack_transposed_kernel(const float* __restrict__ src, int rows, int cols, int ld, T* __restrict__ dst, int round_tf32) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)cols * ld) return;
  const int c = (int)(i / ld), r = (int)(i - (int64_t)c * ld);
  float v = r < rows ? src[(int64_t)r * cols + c] : 0.0f;
  if (round_tf32) v = tf32_rna(v);
  dst[i] = from_f32<T>(v);
}

struct ConvBwdWs { size_t gyt, xt, wt, part, total; };
static ConvBwdWs conv_bwd_ws_layout(int B, int cin, int cout, int64_t hw, int precision) {
  const size_t e = precision == SFNO_PREC_BF16 ? 2 : 4;
  const bool stage = precision != SFNO_PREC_F32;
  ConvBwdWs w{};
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 1024); return o; };
  w.gyt = take(stage ? (size_t)B * cout * hw * e : 0);
  w.xt = take(stage ? (size_t)B * cin * hw * e : 0);
  w.wt = take((size_t)cin * round_up(cout, 8) * e);
  const int splits = std::max(pick_splits(hw), pick_splits_tc(hw, B, cin, cout));
  w.part = take((size_t)B * splits * cin * cout * sizeof(float));
  w.total = off + 1024;
  return w;
}

template <class T>
static int conv1x1_backward_impl(const float* x, const float* gy, const float* wgt, float* gx, float* gw, float* gb, int B, int cin, int cout,
                                 int64_t hw, bool tf32, int precision, char* ws, cudaStream_t st) {
  const ConvBwdWs L = conv_bwd_ws_layout(B, cin, cout, hw, precision);
  const bool stage = precision != SFNO_PREC_F32;
  Tf32Scope scope(tf32);
  const T* gyt = (const T*)gy;
  const T* xt = (const T*)x;
  if (stage) {
    ConcatParts parts{};
    parts.src[0] = gy; parts.channels[0] = cout; parts.nparts = 1;
    concat_convert_kernel<T><<<dim3(256, B), 256, 0, st>>>(parts, hw, (T*)(ws + L.gyt), (int64_t)cout * hw, tf32 ? 1 : 0);
    SFNO_TRY(post_launch("convert_grad"));
    gyt = (const T*)(ws + L.gyt);
  }
  if (gx) {   // gx[b][c][p] = sum_o w[o][c] gy[b][o][p]: the forward op with the transposed weight
    const int ldw = round_up(cout, 8);
    T* wt = (T*)(ws + L.wt);
    pack_transposed_kernel<T><<<(unsigned)ceil_div64((int64_t)cin * ldw, 256), 256, 0, st>>>(wgt, cout, cin, ldw, wt, tf32 ? 1 : 0);
    SFNO_TRY(post_launch("pack_transposed"));
    ConvArgs<T, float> op{};
    op.G = B; op.M = cin; op.N = (int)hw; op.K = cout;
    op.A = wt; op.Bm = gyt; op.a_sk = 1; op.b_sk = hw;
    op.in_bstride = (int64_t)cout * hw; op.w_bstride = 0; op.ldw = ldw;
    op.bias = nullptr; op.bias_bstride = 0; op.act = SFNO_ACT_NONE;
    op.drop_p = 0.0f; op.seed = 0; op.offset = 0; op.rng_dev = nullptr; op.drop_mask = nullptr; op.branch_scale = nullptr;
    op.res = nullptr; op.res_bstride = 0; op.res_a = nullptr; op.res_d = nullptr; op.pos = nullptr;
    op.out = gx; op.out_bstride = (int64_t)cin * hw; op.stat_part = nullptr; op.round_out = 0;
    SFNO_TRY(launch_conv(op, st, "conv1x1_data_grad"));
  }
  if (gw) {
    if (stage) {
      ConcatParts parts{};
      parts.src[0] = x; parts.channels[0] = cin; parts.nparts = 1;
      concat_convert_kernel<T><<<dim3(256, B), 256, 0, st>>>(parts, hw, (T*)(ws + L.xt), (int64_t)cin * hw, tf32 ? 1 : 0);
      SFNO_TRY(post_launch("convert_input"));
      xt = (const T*)(ws + L.xt);
    }
    OpWgradTc<T> op{};
    op.splits = stage ? pick_splits_tc(hw, B, cin, cout) : pick_splits(hw);
    op.hw = hw;
    op.G = B * op.splits; op.M = cout; op.N = cin; op.K = (int)(hw / op.splits);
    op.A = gyt; op.Bm = xt; op.a_sk = 1; op.b_sk = 1; op.part = (float*)(ws + L.part);
    SFNO_TRY(launch_gemm(op, st, "conv1x1_weight_grad"));
    const int64_t n = (int64_t)cout * cin;
    reduce_groups_kernel<<<(unsigned)std::min<int64_t>(ceil_div64(n, 256), 1024), 256, 0, st>>>(op.part, op.G, n, gw);
    SFNO_TRY(post_launch("reduce_groups"));
  }
  if (gb) SFNO_TRY(launch_bias_grad(gy, B, cout, hw, gb, st));
  return SFNO_OK;
}

}  // namespace sfno

using namespace sfno;

extern "C" {

// Backward of sfno_conv1x1_ex without activation / dropout, on the engine of `precision`: data gradient = the forward op
// with the transposed weight, weight gradient = split-K GEMM over the pixels (tensor cores in bf16 / tf32, fp32 partials).
size_t sfno_conv1x1_backward_workspace_bytes(int batch, int cin, int cout, int64_t hw, int precision) {
  if (batch <= 0 || cin <= 0 || cout <= 0 || hw <= 0) return 0;
  return conv_bwd_ws_layout(batch, cin, cout, hw, precision).total;
}

int sfno_conv1x1_backward(const float* x_dev, const float* grad_y_dev, const float* weight_dev, float* grad_x_dev, float* grad_w_dev,
                          float* grad_b_dev, int batch, int cin, int cout, int64_t hw, int precision, void* workspace_dev,
                          size_t workspace_bytes, void* stream) {
  SFNO_CHECK_ARG(grad_y_dev && workspace_dev, "NULL argument");
  SFNO_CHECK_ARG(!grad_x_dev || weight_dev, "the data gradient needs the weight");
  SFNO_CHECK_ARG(!grad_w_dev || x_dev, "the weight gradient needs x");
  SFNO_CHECK_ARG(batch > 0 && cin > 0 && cout > 0 && hw > 0 && hw < (1ll << 31), "bad sizes");
  SFNO_CHECK_ARG(precision == SFNO_PREC_F32 || precision == SFNO_PREC_BF16 || precision == SFNO_PREC_TF32, "bad precision %d", precision);
  SFNO_CHECK_ARG(((uintptr_t)workspace_dev & 1023) == 0, "workspace must be 1024-byte aligned");
  if (workspace_bytes < sfno_conv1x1_backward_workspace_bytes(batch, cin, cout, hw, precision)) return fail(SFNO_ERR_WORKSPACE_TOO_SMALL, "workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  NvtxRange range("sfno_conv1x1_backward");
  if (precision == SFNO_PREC_BF16)
    return conv1x1_backward_impl<bf16>(x_dev, grad_y_dev, weight_dev, grad_x_dev, grad_w_dev, grad_b_dev, batch, cin, cout, hw, false, precision,
                                       (char*)workspace_dev, st);
  return conv1x1_backward_impl<float>(x_dev, grad_y_dev, weight_dev, grad_x_dev, grad_w_dev, grad_b_dev, batch, cin, cout, hw,
                                      precision == SFNO_PREC_TF32, precision, (char*)workspace_dev, st);
}

size_t sfno_conv1x1_weight_grad_workspace_bytes(int batch, int cin, int cout, int64_t hw) {
  if (batch <= 0 || cin <= 0 || cout <= 0 || hw <= 0) return 0;
  return (size_t)batch * pick_splits(hw) * cin * cout * sizeof(float);
}

int sfno_conv1x1_weight_grad(const float* x_dev, const float* grad_y_dev, float* grad_w_dev, float* grad_b_dev, int batch, int cin, int cout,
                             int64_t hw, void* workspace_dev, size_t workspace_bytes, void* stream) {
  SFNO_CHECK_ARG(x_dev && grad_y_dev && grad_w_dev && workspace_dev, "NULL argument");
  SFNO_CHECK_ARG(batch > 0 && cin > 0 && cout > 0 && hw > 0 && hw < (1ll << 31), "bad sizes");
  if (workspace_bytes < sfno_conv1x1_weight_grad_workspace_bytes(batch, cin, cout, hw)) return fail(SFNO_ERR_WORKSPACE_TOO_SMALL, "workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  OpWgrad op{};
  op.splits = pick_splits(hw); op.hw = hw;
  op.G = batch * op.splits; op.M = cout; op.N = cin; op.K = (int)(hw / op.splits);
  op.A = grad_y_dev; op.Bm = x_dev; op.a_sk = 1; op.b_sk = 1; op.part = (float*)workspace_dev;
  SFNO_TRY(launch_gemm_simt(op, st, "conv1x1_weight_grad"));
  const int64_t n = (int64_t)cout * cin;
  reduce_groups_kernel<<<(unsigned)std::min<int64_t>(ceil_div64(n, 256), 1024), 256, 0, st>>>(op.part, op.G, n, grad_w_dev);
  SFNO_TRY(post_launch("reduce_groups"));
  if (grad_b_dev) {
    SFNO_TRY(launch_bias_grad(grad_y_dev, batch, cout, hw, grad_b_dev, st));
  }
  return SFNO_OK;
}

int sfno_spectral_contract_backward(int operator_type, const float* x_dev, const float* weight_dev, const float* grad_out_dev, float* grad_x_dev,
                                    float* grad_w_dev, int batch, int cin, int cout, int lmax, int mmax, void* stream) {
  SFNO_CHECK_ARG(x_dev && weight_dev && grad_out_dev, "NULL argument");
  SFNO_CHECK_ARG(operator_type == SFNO_OP_DHCONV || operator_type == SFNO_OP_DIAGONAL, "bad operator_type %d", operator_type);
  SFNO_CHECK_ARG(batch > 0 && cin > 0 && cout > 0 && lmax > 0 && mmax > 0, "bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  const int diag = operator_type == SFNO_OP_DIAGONAL;
  if (grad_x_dev) {
    const int64_t total = (int64_t)batch * cin * lmax * mmax;
    contract_grad_x_kernel<<<(unsigned)std::min<int64_t>(ceil_div64(total, 256), 1 << 20), 256, 0, st>>>(
        diag, (const float2*)grad_out_dev, (const float2*)weight_dev, (float2*)grad_x_dev, batch, cin, cout, lmax, mmax);
    SFNO_TRY(post_launch("contract_grad_x"));
  }
  if (grad_w_dev) {
    const int64_t nw = (int64_t)cin * cout * lmax * (diag ? mmax : 1);
    contract_grad_w_kernel<<<(unsigned)std::min<int64_t>(ceil_div64(nw * 32, 256), 1 << 20), 256, 0, st>>>(
        diag, (const float2*)x_dev, (const float2*)grad_out_dev, (float2*)grad_w_dev, batch, cin, cout, lmax, mmax);
    SFNO_TRY(post_launch("contract_grad_w"));
  }
  return SFNO_OK;
}

int sfno_instance_norm_backward(const float* x_dev, const float* grad_out_dev, const float* affine_a_dev, float* grad_x_dev, float* grad_a_dev,
                                float* grad_d_dev, int batch, int channels, int64_t hw, float eps, void* stream) {
  SFNO_CHECK_ARG(x_dev && grad_out_dev && grad_x_dev && grad_a_dev && grad_d_dev, "NULL argument");
  SFNO_CHECK_ARG(batch > 0 && channels > 0 && hw > 0, "bad sizes");
  instance_norm_backward_kernel<<<batch * channels, 512, 0, (cudaStream_t)stream>>>(x_dev, grad_out_dev, affine_a_dev, hw, eps, grad_x_dev,
                                                                                  grad_a_dev, grad_d_dev);
  return post_launch("instance_norm_backward");
}

}  // extern "C"
