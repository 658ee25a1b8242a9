// Op-level B200-native entry points of the spectral branch (SURVEY 8b "must export"): the fused SpectralConvS2.forward
// (s2convolutions.py:158-193) and precision-selectable pieces, on the SAME tensor-core ops the whole-network executor
// uses (bf16: kind::f16, tf32: kind::tf32, fp32: CUDA-core parity engine) -- so a maintainer who swaps only the
// SpectralConvS2 / Conv2d sub-modules of the reference model gets the B200 path, not an fp32 side door.
#include <cstring>

#include "backward.cuh"
#include "common.cuh"
#include "engine.cuh"
#include "ops.cuh"
#include "pointwise.cuh"
#include "sht_plan.cuh"

struct sfno_spectral_weight {
  int operator_type = 0, cin = 0, cout = 0, lmax = 0, mmax = 0, precision = 0;
  void* wpack = nullptr;   // dhconv: T [lmax][2*cout][2*cin] (packed real form of the complex weight)
  float* wdiag = nullptr;  // diagonal: fp32 [cin][cout][lmax][mmax][2] (reference layout)
  float* bias = nullptr;   // [cout]
  bool has_bias = false;
};

namespace sfno {

struct SpecWs { size_t xt, fg, X, Y, total; };
static SpecWs spec_ws_layout(const ShtDeviceTables& f, const ShtDeviceTables& i, int B, int cin, int cout) {
  const size_t e = f.precision == SFNO_PREC_BF16 ? 2 : 4;
  const int cmax = std::max(cin, cout);
  SpecWs w{};
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 1024); return o; };
  w.xt = take((size_t)B * cin * f.nlat * f.nlon * e);
  w.fg = take((size_t)f.mmax * B * 2 * cmax * std::max(f.Kp, i.Kp) * e);
  w.X = take((size_t)f.lmax * f.mmax * B * 2 * cin * e);
  w.Y = take((size_t)f.lmax * f.mmax * B * 2 * cout * e);
  w.total = off;
  return w;
}

// analysis-shaped pair with (basis, table) = (efwd, wq): the forward transform; (einv^T, pct): the ADJOINT of the inverse
template <class T>
static int spec_forward(const ShtDeviceTables& t, const void* basis, const void* table, int B, int C, const T* x, T* F, T* X, int triangular,
                        cudaStream_t st) {
  OpDft<T> dft{};
  dft.G = B * C; dft.M = 2 * t.mmax; dft.N = t.nlat; dft.K = t.nlon;
  dft.A = (const T*)basis; dft.Bm = x; dft.a_sk = 1; dft.b_sk = 1;
  dft.f = F; dft.aff_a = nullptr; dft.aff_d = nullptr;
  dft.B = B; dft.C = C; dft.nlat = t.nlat; dft.nlon = t.nlon; dft.Kp = t.Kp; dft.Wp = t.Wp; dft.x_bstride = (int64_t)C * t.nlat * t.nlon;
  dft.a_reps = t.basis_reps; dft.round_out = 1;
  SFNO_TRY(launch_gemm(dft, st, "dft_fwd"));
  OpLeg<T> leg{};
  leg.G = t.mmax; leg.M = t.lmax; leg.N = B * 2 * C; leg.K = t.nlat;
  leg.A = (const T*)table; leg.Bm = F; leg.a_sk = 1; leg.b_sk = 1;
  leg.x = X; leg.Kp = t.Kp; leg.lmax = t.lmax; leg.mmax = t.mmax; leg.triangular = triangular; leg.round_out = 1;
  return launch_gemm(leg, st, "legendre_fwd");
}

// synthesis-shaped pair with (table, basis) = (pt, einv): the inverse transform; (wq^T, efwd^T): the ADJOINT of the forward
template <class T>
static int spec_inverse(const ShtDeviceTables& t, const void* table, const void* basis, int B, int C, const T* S, T* G, const float* bias,
                        float* out, int triangular, cudaStream_t st) {
  OpIleg<T> il{};
  il.G = t.mmax; il.M = B * 2 * C; il.N = t.nlat; il.K = t.lmax;
  il.A = S; il.Bm = (const T*)table; il.b_sk = 1;
  il.a_goff = il.M; il.a_sk = (int64_t)t.mmax * il.M;
  il.g_out = G; il.B = B; il.C = C; il.Kp = t.Kp; il.Lq = t.Lq; il.nlat = t.nlat; il.triangular = triangular; il.round_out = 1;
  SFNO_TRY(launch_gemm(il, st, "legendre_inv"));
  IdftArgs<T, float> id{};
  id.G = 1; id.M = B * C * t.Kp; id.N = t.nlon; id.K = 2 * t.mmax;
  id.A = G; id.Bm = (const T*)basis; id.a_sk = id.M; id.b_sk = 1;
  id.out = out; id.out_bstride = (int64_t)C * t.nlat * t.nlon; id.bias = bias; id.add = nullptr; id.add_bstride = 0; id.act = SFNO_ACT_NONE;
  id.C = C; id.nlat = t.nlat; id.nlon = t.nlon; id.Kp = t.Kp; id.Kq2 = t.Kq2; id.b_reps = t.basis_reps; id.stat_part = nullptr;
  id.round_out = 0;   // leaves the library in full fp32
  return launch_idft(id, st, "dft_inv");
}

template <class T>
static int spectral_conv_impl(const ShtDeviceTables& f, const ShtDeviceTables& i, const sfno_spectral_weight* w, const float* x, float* y,
                              float* residual, int B, char* ws, cudaStream_t st) {
  const SpecWs L = spec_ws_layout(f, i, B, w->cin, w->cout);
  T* xt = (T*)(ws + L.xt);
  T* FG = (T*)(ws + L.fg);
  T* X = (T*)(ws + L.X);
  T* Y = (T*)(ws + L.Y);
  const bool tf32 = f.precision == SFNO_PREC_TF32;
  Tf32Scope scope(tf32);
  const int64_t per = (int64_t)w->cin * f.nlat * f.nlon;
  const T* xin;
  if constexpr (std::is_same<T, float>::value) {
    if (tf32) {   // operands of a tf32 MMA are stored TF32-exact
      ConcatParts parts{};
      parts.src[0] = x; parts.channels[0] = w->cin; parts.nparts = 1;
      concat_convert_kernel<float><<<dim3(256, B), 256, 0, st>>>(parts, (int64_t)f.nlat * f.nlon, xt, per, 1);
      SFNO_TRY(post_launch("convert_input"));
      xin = xt;
    } else {
      xin = x;
    }
  } else {
    ConcatParts parts{};
    parts.src[0] = x; parts.channels[0] = w->cin; parts.nparts = 1;
    concat_convert_kernel<T><<<dim3(256, B), 256, 0, st>>>(parts, (int64_t)f.nlat * f.nlon, xt, per, 0);
    SFNO_TRY(post_launch("convert_input"));
    xin = xt;
  }
  const int tri = w->operator_type == SFNO_OP_DHCONV;   // dhconv never mixes wavenumbers: only degrees l >= m carry information
  SFNO_TRY(spec_forward<T>(f, f.efwd, f.wq, B, w->cin, xin, FG, X, tri, st));
  if (residual) SFNO_TRY(spec_inverse<T>(i, i.pt, i.einv, B, w->cin, X, FG, nullptr, residual, tri, st));   // s2convolutions.py:166-169
  if (w->operator_type == SFNO_OP_DHCONV) {
    OpDhconv<T> op{};
    op.G = f.lmax; op.M = f.mmax * B; op.N = 2 * w->cout; op.K = 2 * w->cin;
    op.A = X; op.Bm = (const T*)w->wpack; op.a_sk = 1; op.b_sk = 1;
    op.y = Y; op.B = B; op.lmax = f.lmax; op.mmax = f.mmax; op.triangular = 1; op.round_out = 1;
    SFNO_TRY(launch_gemm(op, st, "dhconv"));
  } else {
    const int64_t total = (int64_t)f.lmax * f.mmax * B * w->cout;
    diag_contract_internal_kernel<T><<<(unsigned)std::min<int64_t>(ceil_div64(total, 128), 1 << 20), 128, 0, st>>>(
        X, (const float2*)w->wdiag, Y, B, w->cout, f.lmax, f.mmax);
    SFNO_TRY(post_launch("diag_contract"));
  }
  return spec_inverse<T>(i, i.pt, i.einv, B, w->cout, Y, FG, w->has_bias ? w->bias : nullptr, y, tri, st);
}

// ---- backward of the fused SpectralConvS2.forward ---------------------------------------------------------------------------
// With y = iSHT(W . SHT(x)) + bias and residual = iSHT(SHT(x)) (scale_residual), the cotangents (gy, gres) give
//   GY = iSHT^T gy                       (analysis-shaped ops on the inverse plan's transposed tables)
//   gX = W^H . GY (+ iSHT^T gres)        (the forward contraction op with conjugate-transposed weights)
//   gx = SHT^T gX                        (synthesis-shaped ops on the forward plan's transposed tables)
//   gW = conj(SHT(x)) . GY summed over (m, b)   (per-degree GEMM over the (m, b) rows), gbias = sum gy
// -- the same tensor-core ops as the forward for everything but the weight gradient (fp32 CUDA-core GEMM).

// gWp[l][(ri',o)][(ri,c)] = sum_{(m,b)} GY[l][(m,b)][(ri',o)] * X[l][(m,b)][(ri,c)]: per degree, both operands with the
// GEMM row / column index contiguous (MN-major) and the contraction over the (m, b) rows.  Tensor cores in bf16 / tf32
// (fp32 result), CUDA cores in fp32.  The tensor-core engine contracts over ALL rows: the caller zero-fills the rows the
// triangular transforms do not write; the CUDA-core engine stops at the live rows (k_end).
template <class T>
struct OpDhconvWgrad : NoFeatures {
  static constexpr bool kGFastest = false;
  static constexpr bool kSimtRowsOnFastLanes = false;
  static constexpr bool kRanged = true;
  static constexpr bool A_KCONTIG = false, B_KCONTIG = false, kColContig = true, kNFastest = false;
  using OutT = float;
  using InT = T;
  __device__ bool out_tf32() const { return false; }
  int G, M, N, K;              // G = lmax, M = 2*cout, N = 2*cin, K = mmax*B rows
  const T* A; const T* Bm; int64_t a_sk, b_sk;
  int B, lmax, triangular;
  float* out;                  // [lmax][M][N]
  __device__ int n_begin(int) const { return 0; }
  __device__ int n_end(int) const { return N; }
  __device__ int m_begin(int) const { return 0; }
  __device__ int m_end(int) const { return M; }
  __device__ int k_begin(int) const { return 0; }
  __device__ int k_end(int g) const {   // rows of the live wavenumbers of degree g (what the triangular ops have written)
    const int last = lmax > 0 ? ((lmax - 1) & ~63) : 0;
    if (!triangular || g >= last) return K;
    const int e = ((g | 63) + 1) * B;
    return e < K ? e : K;
  }
  __device__ int64_t a_off(int g, int m) const { return (int64_t)g * K * M + m; }
  __device__ int64_t b_off(int g, int n) const { return (int64_t)g * K * N + n; }
  __device__ int n_store() const { return N; }
  __device__ void io_coords(int g, int row0, int col0, int (&c)[5]) const { c[0] = col0; c[1] = row0; c[2] = g; c[3] = 0; c[4] = 0; }
  struct Row { float* out; const float* res; bool valid; __device__ float stat_s() const { return 0.0f; } __device__ float stat_q() const { return 0.0f; } };
  __device__ Row row(int g, int m) const { return Row{out + ((int64_t)g * M + m) * N, nullptr, true}; }
  template <int F> __device__ Row row_f(int g, int m) const { return row(g, m); }
  __device__ void store(const Row& r, int, int, int n, float acc) const { r.out[n] = acc; }
  template <int F>
  __device__ void compute8(Row&, int, const float (&acc)[8], const float (&)[8], float (&o)[8]) const {
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = acc[i];
  }
};

template <class T>
struct TcTraits<OpDhconvWgrad<T>> : TcTraitsBase<OpDhconvWgrad<T>>, TcEligible<TcTraits<OpDhconvWgrad<T>>, OpDhconvWgrad<T>> {
  static constexpr int BN = 256;
  static constexpr uint64_t es = sizeof(T);
  static void operands(const OpDhconvWgrad<T>& op, TmaOperand& a, TmaOperand& b) {
    a.base = op.A; a.dims[0] = op.M; a.dims[1] = op.K; a.dims[2] = op.G;   // M-contiguous: {(ri',o), (m,b), l}
    a.strides[0] = (uint64_t)op.M * es; a.strides[1] = (uint64_t)op.K * op.M * es; a.batched = true;
    b.base = op.Bm; b.dims[0] = op.N; b.dims[1] = op.K; b.dims[2] = op.G;  // N-contiguous: {(ri,c), (m,b), l}
    b.strides[0] = (uint64_t)op.N * es; b.strides[1] = (uint64_t)op.K * op.N * es; b.batched = true;
  }
  static bool extra_ok(const OpDhconvWgrad<T>& op) { return aligned16(op.out) && op.N % 8 == 0 && op.M % 8 == 0; }
  static void io(const OpDhconvWgrad<T>& op, TmaIo& o, TmaIo&) {
    o.base = op.out; o.es = 4; o.ok = true;
    o.dims[0] = op.N; o.dims[1] = op.M; o.dims[2] = op.G;
    o.strides[0] = (uint64_t)op.N * 4; o.strides[1] = (uint64_t)op.M * op.N * 4;
    o.box_rows[0] = 32;
  }
};

// packed real form of the conjugate-transposed dhconv weight: rows (ri, c), cols (ri', o) = Wp^T
template <class T>
static __global__ void pack_dhconv_weight_adjoint_kernel(const float* __restrict__ w, int cin, int cout, int L, T* __restrict__ dst, int round_tf32) {
  const int64_t total = (int64_t)L * 2 * cin * 2 * cout;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int kk = (int)(i % (2 * cout));
    const int64_t r = i / (2 * cout);
    const int mm = (int)(r % (2 * cin));
    const int l = (int)(r / (2 * cin));
    const int ri_o = kk / cout, o = kk - ri_o * cout;     // contraction index (ri', o)
    const int ri_c = mm / cin, c = mm - ri_c * cin;       // output index (ri, c)
    const float* src = w + (((int64_t)c * cout + o) * L + l) * 2;
    const float wr = src[0], wi = src[1];
    float v = (ri_o == ri_c) ? wr : (ri_c == 1 ? -wi : wi);   // Wp[(ri',o)][(ri,c)] read transposed
    if (round_tf32) v = tf32_rna(v);
    dst[i] = from_f32<T>(v);
  }
}

// gw[c][o][l] = (gWp[l][(0,o)][(0,c)] + gWp[l][(1,o)][(1,c)],  gWp[l][(1,o)][(0,c)] - gWp[l][(0,o)][(1,c)])
// One block per (32 input channels, 32 degrees, output channel): reads run along c, writes along l, through a shared tile.
static __global__ void __launch_bounds__(256) unpack_dhconv_wgrad_kernel(const float* __restrict__ gwp, int cin, int cout, int L,
                                                                         float2* __restrict__ gw) {
  __shared__ float2 tile[32][33];
  const int c0 = blockIdx.x * 32, l0 = blockIdx.y * 32, o = blockIdx.z;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  const int64_t N = 2 * cin;
  for (int j = ty; j < 32; j += 8) {
    const int l = l0 + j, c = c0 + tx;
    if (l < L && c < cin) {
      const float* p = gwp + (int64_t)l * 4 * cin * cout;
      const float a = p[(int64_t)o * N + c], d = p[(int64_t)(cout + o) * N + cin + c];
      const float b1 = p[(int64_t)(cout + o) * N + c], b0 = p[(int64_t)o * N + cin + c];
      tile[j][tx] = make_float2(a + d, b1 - b0);
    }
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int c = c0 + j, l = l0 + tx;
    if (c < cin && l < L) gw[((int64_t)c * cout + o) * L + l] = tile[tx][j];
  }
}

// diagonal operator, adjoint on the internal layouts: gX[l][m][b][ri][i] = sum_o GY[l][m][b][.][o] conj(w[i][o][l][m])
template <class T>
static __global__ void diag_contract_adjoint_kernel(const T* __restrict__ GY, const float2* __restrict__ w, T* __restrict__ GX, int B, int C,
                                                    int L, int M, int round_tf32) {
  const int64_t total = (int64_t)L * M * B * C;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(idx % C);
    int64_t r = idx / C;
    const int b = (int)(r % B); r /= B;
    const int m = (int)(r % M);
    const int l = (int)(r / M);
    const T* g = GY + (((int64_t)l * M + m) * B + b) * 2 * C;
    float re = 0.0f, im = 0.0f;
    for (int o = 0; o < C; ++o) {
      const float ga = to_f32(g[o]), gb = to_f32(g[C + o]);
      const float2 wv = w[(((int64_t)i * C + o) * L + l) * M + m];
      re = fmaf(ga, wv.x, re); re = fmaf(gb, wv.y, re);
      im = fmaf(gb, wv.x, im); im = fmaf(-ga, wv.y, im);
    }
    T* x = GX + (((int64_t)l * M + m) * B + b) * 2 * C;
    x[i] = from_f32<T>(round_tf32 ? tf32_rna(re) : re);
    x[C + i] = from_f32<T>(round_tf32 ? tf32_rna(im) : im);
  }
}

// diagonal operator, weight gradient: gw[i][o][l][m] = sum_b conj(X[l][m][b][.][i]) GY[l][m][b][.][o]
template <class T>
static __global__ void diag_wgrad_kernel(const T* __restrict__ X, const T* __restrict__ GY, float2* __restrict__ gw, int B, int C, int L, int M) {
  const int64_t total = (int64_t)C * C * L * M;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int m = (int)(idx % M);
    int64_t r = idx / M;
    const int l = (int)(r % L); r /= L;
    const int o = (int)(r % C);
    const int i = (int)(r / C);
    float re = 0.0f, im = 0.0f;
    for (int b = 0; b < B; ++b) {
      const int64_t base = (((int64_t)l * M + m) * B + b) * 2 * C;
      const float xa = to_f32(X[base + i]), xb = to_f32(X[base + C + i]);
      const float ga = to_f32(GY[base + o]), gb = to_f32(GY[base + C + o]);
      re = fmaf(xa, ga, re); re = fmaf(xb, gb, re);
      im = fmaf(xa, gb, im); im = fmaf(-xb, ga, im);
    }
    gw[idx] = make_float2(re, im);
  }
}

template <class T>
static __global__ void add_inplace_kernel(T* __restrict__ a, const T* __restrict__ b, int64_t n, int round_tf32) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = to_f32(a[i]) + to_f32(b[i]);
    a[i] = from_f32<T>(round_tf32 ? tf32_rna(v) : v);
  }
}

struct SpecBwdWs { size_t xt, fg, GY, GX, X, wadj, gwp, total; };
static SpecBwdWs spec_bwd_ws_layout(const ShtDeviceTables& f, const ShtDeviceTables& i, int B, int cin, int cout, int dhconv) {
  const size_t e = f.precision == SFNO_PREC_BF16 ? 2 : 4;
  const int cmax = std::max(cin, cout);
  const size_t plane = std::max((size_t)f.nlat * f.nlon, (size_t)i.nlat * i.nlon);
  SpecBwdWs w{};
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 1024); return o; };
  w.xt = take((size_t)B * cmax * plane * e);
  w.fg = take((size_t)f.mmax * B * 2 * cmax * std::max(f.Kp, i.Kp) * e);
  w.GY = take((size_t)f.lmax * f.mmax * B * 2 * cout * e);
  w.GX = take((size_t)f.lmax * f.mmax * B * 2 * cin * e);
  w.X = take((size_t)f.lmax * f.mmax * B * 2 * cin * e);
  w.wadj = take(dhconv ? (size_t)f.lmax * 4 * cin * cout * e : 0);
  w.gwp = take(dhconv ? (size_t)f.lmax * 4 * cin * cout * sizeof(float) : 0);
  w.total = off;
  return w;
}

// fp32 field -> operand type T (rounded to TF32 in tf32 mode); returns the pointer the transforms should read
template <class T>
static int stage_field(const float* x, int B, int C, int64_t plane, bool tf32, T* xt, const T** out, cudaStream_t st) {
  if constexpr (std::is_same<T, float>::value) {
    if (!tf32) { *out = x; return SFNO_OK; }
  }
  ConcatParts parts{};
  parts.src[0] = x; parts.channels[0] = C; parts.nparts = 1;
  concat_convert_kernel<T><<<dim3(256, B), 256, 0, st>>>(parts, plane, xt, (int64_t)C * plane, tf32 ? 1 : 0);
  SFNO_TRY(post_launch("convert_input"));
  *out = xt;
  return SFNO_OK;
}

template <class T>
static int spectral_conv_backward_impl(const ShtDeviceTables& f, const ShtDeviceTables& i, const sfno_spectral_weight* w, const float* weight,
                                       const float* x, const float* gy, const float* gres, float* gx, float* gw, float* gb, int B, char* ws,
                                       cudaStream_t st) {
  const int dh = w->operator_type == SFNO_OP_DHCONV;
  const SpecBwdWs L = spec_bwd_ws_layout(f, i, B, w->cin, w->cout, dh);
  T* xt = (T*)(ws + L.xt);
  T* FG = (T*)(ws + L.fg);
  T* GY = (T*)(ws + L.GY);
  T* GX = (T*)(ws + L.GX);
  T* X = (T*)(ws + L.X);
  const bool tf32 = f.precision == SFNO_PREC_TF32;
  Tf32Scope scope(tf32);
  const int tri = dh;
  const int64_t plane_f = (int64_t)f.nlat * f.nlon, plane_i = (int64_t)i.nlat * i.nlon;
  const T* src;
  if (dh && gw) {   // the weight-gradient GEMM of the tensor-core engine contracts over every (m, b) row
    SFNO_CUDA(cudaMemsetAsync(GY, 0, (size_t)f.lmax * f.mmax * B * 2 * w->cout * sizeof(T), st));
    SFNO_CUDA(cudaMemsetAsync(X, 0, (size_t)f.lmax * f.mmax * B * 2 * w->cin * sizeof(T), st));
  }
  // GY = iSHT^T gy
  SFNO_TRY(stage_field<T>(gy, B, w->cout, plane_i, tf32, xt, &src, st));
  SFNO_TRY(spec_forward<T>(i, i.einv_t, i.pct_a, B, w->cout, src, FG, GY, tri, st));
  if (gb) SFNO_TRY(launch_bias_grad(gy, B, w->cout, plane_i, gb, st));
  if (gx) {
    if (dh) {
      T* wadj = (T*)(ws + L.wadj);
      pack_dhconv_weight_adjoint_kernel<T><<<4096, 256, 0, st>>>(weight, w->cin, w->cout, f.lmax, wadj, tf32 ? 1 : 0);
      SFNO_TRY(post_launch("pack_dhconv_weight_adjoint"));
      OpDhconv<T> op{};
      op.G = f.lmax; op.M = f.mmax * B; op.N = 2 * w->cin; op.K = 2 * w->cout;
      op.A = GY; op.Bm = wadj; op.a_sk = 1; op.b_sk = 1;
      op.y = GX; op.B = B; op.lmax = f.lmax; op.mmax = f.mmax; op.triangular = 1; op.round_out = 1;
      SFNO_TRY(launch_gemm(op, st, "dhconv_adjoint"));
    } else {
      const int64_t total = (int64_t)f.lmax * f.mmax * B * w->cin;
      diag_contract_adjoint_kernel<T><<<(unsigned)std::min<int64_t>(ceil_div64(total, 128), 1 << 20), 128, 0, st>>>(
          GY, (const float2*)weight, GX, B, w->cin, f.lmax, f.mmax, tf32 ? 1 : 0);
      SFNO_TRY(post_launch("diag_contract_adjoint"));
    }
    if (gres) {   // residual = iSHT(SHT(x)): its cotangent joins gX in the spectral domain
      SFNO_TRY(stage_field<T>(gres, B, w->cin, plane_i, tf32, xt, &src, st));
      SFNO_TRY(spec_forward<T>(i, i.einv_t, i.pct_a, B, w->cin, src, FG, X, tri, st));
      const int64_t n = (int64_t)f.lmax * f.mmax * B * 2 * w->cin;
      add_inplace_kernel<T><<<(unsigned)std::min<int64_t>(ceil_div64(n, 256), 148 * 16), 256, 0, st>>>(GX, X, n, tf32 ? 1 : 0);
      SFNO_TRY(post_launch("add_spectral"));
    }
    SFNO_TRY(spec_inverse<T>(f, f.wq_t, f.efwd_t, B, w->cin, GX, FG, nullptr, gx, tri, st));
  }
  if (gw) {
    SFNO_TRY(stage_field<T>(x, B, w->cin, plane_f, tf32, xt, &src, st));
    SFNO_TRY(spec_forward<T>(f, f.efwd, f.wq, B, w->cin, src, FG, X, tri, st));
    if (dh) {
      float* gwp = (float*)(ws + L.gwp);
      OpDhconvWgrad<T> op{};
      op.G = f.lmax; op.M = 2 * w->cout; op.N = 2 * w->cin; op.K = f.mmax * B;
      op.A = GY; op.Bm = X; op.a_sk = 2 * w->cout; op.b_sk = 2 * w->cin;
      op.B = B; op.lmax = f.lmax; op.triangular = 1; op.out = gwp;
      SFNO_TRY(launch_gemm(op, st, "dhconv_weight_grad"));
      unpack_dhconv_wgrad_kernel<<<dim3(ceil_div(w->cin, 32), ceil_div(f.lmax, 32), w->cout), 256, 0, st>>>(gwp, w->cin, w->cout, f.lmax,
                                                                                                          (float2*)gw);
      SFNO_TRY(post_launch("unpack_dhconv_wgrad"));
    } else {
      const int64_t total = (int64_t)w->cin * w->cout * f.lmax * f.mmax;
      diag_wgrad_kernel<T><<<(unsigned)std::min<int64_t>(ceil_div64(total, 256), 1 << 20), 256, 0, st>>>(X, GY, (float2*)gw, B, w->cin, f.lmax,
                                                                                                          f.mmax);
      SFNO_TRY(post_launch("diag_weight_grad"));
    }
  }
  return SFNO_OK;
}

}  // namespace sfno

using namespace sfno;

template <class T>
static int conv1x1_ex_impl(const float* x, const float* wgt, const float* bias, const float* residual, float* y, int B, int cin, int cout,
                           int64_t hw, int act, float drop_p, uint64_t seed, uint64_t offset, bool tf32, char* ws, cudaStream_t st) {
  const int ldw = round_up(cin, 8);
  T* xt = (T*)ws;
  T* wt = (T*)(ws + align_up((size_t)B * cin * hw * sizeof(T), 1024));
  T* rt = (T*)((char*)wt + align_up((size_t)cout * ldw * sizeof(T), 1024));
  Tf32Scope scope(tf32);
  ConcatParts parts{};
  parts.src[0] = x; parts.channels[0] = cin; parts.nparts = 1;
  concat_convert_kernel<T><<<dim3(256, B), 256, 0, st>>>(parts, hw, xt, (int64_t)cin * hw, tf32 ? 1 : 0);
  SFNO_TRY(post_launch("convert_input"));
  pack_rows_kernel<T><<<(unsigned)ceil_div64((int64_t)cout * ldw, 256), 256, 0, st>>>(wgt, cout, cin, ldw, wt);
  SFNO_TRY(post_launch("pack_rows"));
  if constexpr (std::is_same<T, float>::value) {
    if (tf32) {
      round_tf32_kernel<<<(unsigned)ceil_div64((int64_t)cout * ldw, 256), 256, 0, st>>>((float*)wt, (int64_t)cout * ldw);
      SFNO_TRY(post_launch("round_tf32"));
    }
  }
  ConvArgs<T, float> op{};
  op.G = B; op.M = cout; op.N = (int)hw; op.K = cin;
  op.A = wt; op.Bm = xt; op.a_sk = 1; op.b_sk = hw;
  op.in_bstride = (int64_t)cin * hw; op.w_bstride = 0; op.ldw = ldw;
  op.bias = bias; op.bias_bstride = 0; op.act = act;
  op.drop_p = drop_p; op.seed = seed; op.offset = offset; op.rng_dev = nullptr; op.branch_scale = nullptr;
  op.res = nullptr; op.res_bstride = 0; op.res_a = nullptr; op.res_d = nullptr; op.pos = nullptr;
  op.out = y; op.out_bstride = (int64_t)cout * hw; op.stat_part = nullptr; op.round_out = 0;
  if (residual) {   // the residual is an operand-typed tensor of the op: stage it in T next to the input
    ConcatParts rp{};
    rp.src[0] = residual; rp.channels[0] = cout; rp.nparts = 1;
    concat_convert_kernel<T><<<dim3(256, B), 256, 0, st>>>(rp, hw, rt, (int64_t)cout * hw, 0);
    SFNO_TRY(post_launch("convert_residual"));
    op.res = rt; op.res_bstride = (int64_t)cout * hw;
  }
  return launch_conv(op, st, "conv1x1");
}

extern "C" {

int sfno_spectral_weight_create(sfno_spectral_weight** out, int operator_type, int cin, int cout, int lmax, int mmax, int precision) {
  SFNO_CHECK_ARG(out != nullptr, "NULL argument");
  SFNO_CHECK_ARG(operator_type == SFNO_OP_DHCONV || operator_type == SFNO_OP_DIAGONAL, "bad operator_type %d", operator_type);
  SFNO_CHECK_ARG(cin > 0 && cout > 0 && lmax > 0 && mmax > 0, "bad sizes");
  SFNO_CHECK_ARG(precision == SFNO_PREC_F32 || precision == SFNO_PREC_BF16 || precision == SFNO_PREC_TF32, "bad precision %d", precision);
  if (operator_type == SFNO_OP_DIAGONAL && cin != cout) return fail(SFNO_ERR_UNSUPPORTED, "the diagonal operator needs cin == cout");
  auto* w = new sfno_spectral_weight();
  w->operator_type = operator_type; w->cin = cin; w->cout = cout; w->lmax = lmax; w->mmax = mmax; w->precision = precision;
  const size_t e = precision == SFNO_PREC_BF16 ? 2 : 4;
  cudaError_t err = cudaSuccess;
  if (operator_type == SFNO_OP_DHCONV) err = cudaMalloc(&w->wpack, (size_t)lmax * 4 * cin * cout * e);
  else err = cudaMalloc((void**)&w->wdiag, (size_t)cin * cout * lmax * mmax * 2 * sizeof(float));
  if (err == cudaSuccess) err = cudaMalloc((void**)&w->bias, (size_t)cout * sizeof(float));
  if (err != cudaSuccess) {
    cudaGetLastError();
    cudaFree(w->wpack); cudaFree(w->wdiag); cudaFree(w->bias);
    delete w;
    return fail(SFNO_ERR_CUDA, "cudaMalloc failed: %s", cudaGetErrorString(err));
  }
  *out = w;
  return SFNO_OK;
}

int sfno_spectral_weight_destroy(sfno_spectral_weight* w) {
  if (!w) return SFNO_OK;
  cudaFree(w->wpack); cudaFree(w->wdiag); cudaFree(w->bias);
  delete w;
  return SFNO_OK;
}

int sfno_spectral_weight_set(sfno_spectral_weight* w, const float* weight_dev, const float* bias_dev, void* stream) {
  SFNO_CHECK_ARG(w && weight_dev, "NULL argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (w->operator_type == SFNO_OP_DHCONV) {
    const int64_t n = (int64_t)w->lmax * 4 * w->cin * w->cout;
    if (w->precision == SFNO_PREC_BF16) pack_dhconv_weight_kernel<bf16><<<4096, 256, 0, st>>>(weight_dev, w->cin, w->cout, w->lmax, (bf16*)w->wpack);
    else pack_dhconv_weight_kernel<float><<<4096, 256, 0, st>>>(weight_dev, w->cin, w->cout, w->lmax, (float*)w->wpack);
    SFNO_TRY(post_launch("pack_dhconv_weight"));
    if (w->precision == SFNO_PREC_TF32) {
      round_tf32_kernel<<<(unsigned)std::min<int64_t>(ceil_div64(n, 256), 148 * 16), 256, 0, st>>>((float*)w->wpack, n);
      SFNO_TRY(post_launch("round_tf32"));
    }
  } else {
    SFNO_CUDA(cudaMemcpyAsync(w->wdiag, weight_dev, (size_t)w->cin * w->cout * w->lmax * w->mmax * 2 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  }
  w->has_bias = bias_dev != nullptr;
  if (bias_dev) SFNO_CUDA(cudaMemcpyAsync(w->bias, bias_dev, (size_t)w->cout * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return SFNO_OK;
}

size_t sfno_spectral_conv_workspace_bytes(const sfno_sht_plan* fwd, const sfno_sht_plan* inv, const sfno_spectral_weight* w, int batch) {
  if (!fwd || !inv || !w || batch <= 0) return 0;
  return spec_ws_layout(fwd->t, inv->t, batch, w->cin, w->cout).total;
}

int sfno_spectral_conv(const sfno_sht_plan* fwd, const sfno_sht_plan* inv, const sfno_spectral_weight* w, const float* x_dev, float* y_dev,
                       float* residual_dev, int batch, void* workspace_dev, size_t workspace_bytes, void* stream) {
  SFNO_CHECK_ARG(fwd && inv && w && x_dev && y_dev && workspace_dev, "NULL argument");
  SFNO_CHECK_ARG(batch > 0, "bad batch %d", batch);
  const ShtDeviceTables& f = fwd->t;
  const ShtDeviceTables& i = inv->t;
  if (f.precision != i.precision || f.precision != w->precision) return fail(SFNO_ERR_INVALID_ARGUMENT, "plans and weight must share one precision");
  if (f.lmax != i.lmax || f.mmax != i.mmax || f.lmax != w->lmax || f.mmax != w->mmax)
    return fail(SFNO_ERR_SHAPE_MISMATCH, "forward plan (%d,%d), inverse plan (%d,%d) and weight (%d,%d) disagree on (lmax, mmax)", f.lmax, f.mmax,
                i.lmax, i.mmax, w->lmax, w->mmax);
  SFNO_CHECK_ARG(((uintptr_t)workspace_dev & 1023) == 0, "workspace must be 1024-byte aligned");
  if (workspace_bytes < spec_ws_layout(f, i, batch, w->cin, w->cout).total) return fail(SFNO_ERR_WORKSPACE_TOO_SMALL, "workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  NvtxRange range("sfno_spectral_conv");
  return f.precision == SFNO_PREC_BF16 ? spectral_conv_impl<bf16>(f, i, w, x_dev, y_dev, residual_dev, batch, (char*)workspace_dev, st)
                                       : spectral_conv_impl<float>(f, i, w, x_dev, y_dev, residual_dev, batch, (char*)workspace_dev, st);
}

size_t sfno_spectral_conv_backward_workspace_bytes(const sfno_sht_plan* fwd, const sfno_sht_plan* inv, const sfno_spectral_weight* w, int batch) {
  if (!fwd || !inv || !w || batch <= 0) return 0;
  return spec_bwd_ws_layout(fwd->t, inv->t, batch, w->cin, w->cout, w->operator_type == SFNO_OP_DHCONV).total;
}

int sfno_spectral_conv_backward(sfno_sht_plan* fwd, sfno_sht_plan* inv, const sfno_spectral_weight* w, const float* weight_dev,
                                const float* x_dev, const float* grad_y_dev, const float* grad_residual_dev, float* grad_x_dev,
                                float* grad_weight_dev, float* grad_bias_dev, int batch, void* workspace_dev, size_t workspace_bytes,
                                void* stream) {
  SFNO_CHECK_ARG(fwd && inv && w && weight_dev && grad_y_dev && workspace_dev, "NULL argument");
  SFNO_CHECK_ARG(batch > 0, "bad batch %d", batch);
  SFNO_CHECK_ARG(!grad_weight_dev || x_dev, "the weight gradient needs x");
  SFNO_TRY(sht_tables_enable_adjoint(fwd->t));
  SFNO_TRY(sht_tables_enable_adjoint(inv->t));
  const ShtDeviceTables& f = fwd->t;
  const ShtDeviceTables& i = inv->t;
  if (f.precision != i.precision || f.precision != w->precision) return fail(SFNO_ERR_INVALID_ARGUMENT, "plans and weight must share one precision");
  if (f.lmax != i.lmax || f.mmax != i.mmax || f.lmax != w->lmax || f.mmax != w->mmax)
    return fail(SFNO_ERR_SHAPE_MISMATCH, "forward plan, inverse plan and weight disagree on (lmax, mmax)");
  SFNO_CHECK_ARG(((uintptr_t)workspace_dev & 1023) == 0, "workspace must be 1024-byte aligned");
  if (workspace_bytes < sfno_spectral_conv_backward_workspace_bytes(fwd, inv, w, batch)) return fail(SFNO_ERR_WORKSPACE_TOO_SMALL, "workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  NvtxRange range("sfno_spectral_conv_backward");
  return f.precision == SFNO_PREC_BF16
             ? spectral_conv_backward_impl<bf16>(f, i, w, weight_dev, x_dev, grad_y_dev, grad_residual_dev, grad_x_dev, grad_weight_dev,
                                                 grad_bias_dev, batch, (char*)workspace_dev, st)
             : spectral_conv_backward_impl<float>(f, i, w, weight_dev, x_dev, grad_y_dev, grad_residual_dev, grad_x_dev, grad_weight_dev,
                                                  grad_bias_dev, batch, (char*)workspace_dev, st);
}

// nn.Conv2d(cin, cout, 1) with the whole fused epilogue and a selectable engine: precision bf16 / tf32 convert the
// operands into the workspace and run the tensor-core op (fp32 result); fp32 is the CUDA-core parity engine.
size_t sfno_conv1x1_ex_workspace_bytes(int batch, int cin, int cout, int64_t hw, int precision) {
  if (batch <= 0 || cin <= 0 || cout <= 0 || hw <= 0) return 0;
  if (precision == SFNO_PREC_F32) return 1024;
  const size_t e = precision == SFNO_PREC_BF16 ? 2 : 4;
  return align_up((size_t)batch * cin * hw * e, 1024) + align_up((size_t)cout * round_up(cin, 8) * e, 1024) +
         align_up((size_t)batch * cout * hw * e, 1024) + 1024;
}

int sfno_conv1x1_ex(const float* x_dev, const float* weight_dev, const float* bias_dev, const float* residual_dev, float* y_dev, int batch,
                    int cin, int cout, int64_t hw, int activation, float dropout_p, uint64_t seed, uint64_t offset, int precision,
                    void* workspace_dev, size_t workspace_bytes, void* stream) {
  SFNO_CHECK_ARG(x_dev && weight_dev && y_dev, "NULL argument");
  SFNO_CHECK_ARG(batch > 0 && cin > 0 && cout > 0 && hw > 0 && hw < (1ll << 31), "bad sizes");
  SFNO_CHECK_ARG(dropout_p >= 0.0f && dropout_p < 1.0f, "dropout probability %f outside [0, 1)", (double)dropout_p);
  SFNO_CHECK_ARG(precision == SFNO_PREC_F32 || precision == SFNO_PREC_BF16 || precision == SFNO_PREC_TF32, "bad precision %d", precision);
  cudaStream_t st = (cudaStream_t)stream;
  if (precision == SFNO_PREC_F32) {
    ConvArgs<float, float> op{};
    op.G = batch; op.M = cout; op.N = (int)hw; op.K = cin;
    op.A = weight_dev; op.Bm = x_dev; op.a_sk = 1; op.b_sk = hw;
    op.in_bstride = (int64_t)cin * hw; op.w_bstride = 0; op.ldw = cin;
    op.bias = bias_dev; op.bias_bstride = 0; op.act = activation;
    op.drop_p = dropout_p; op.seed = seed; op.offset = offset; op.rng_dev = nullptr; op.branch_scale = nullptr;
    op.res = residual_dev; op.res_bstride = (int64_t)cout * hw; op.res_a = nullptr; op.res_d = nullptr; op.pos = nullptr;
    op.out = y_dev; op.out_bstride = (int64_t)cout * hw; op.stat_part = nullptr; op.round_out = 0;
    return launch_conv(op, st, "conv1x1");
  }
  SFNO_CHECK_ARG(workspace_dev && ((uintptr_t)workspace_dev & 1023) == 0, "workspace must be 1024-byte aligned");
  if (workspace_bytes < sfno_conv1x1_ex_workspace_bytes(batch, cin, cout, hw, precision)) return fail(SFNO_ERR_WORKSPACE_TOO_SMALL, "workspace too small");
  if (precision == SFNO_PREC_BF16)
    return conv1x1_ex_impl<bf16>(x_dev, weight_dev, bias_dev, residual_dev, y_dev, batch, cin, cout, hw, activation, dropout_p, seed, offset, false,
                                 (char*)workspace_dev, st);
  return conv1x1_ex_impl<float>(x_dev, weight_dev, bias_dev, residual_dev, y_dev, batch, cin, cout, hw, activation, dropout_p, seed, offset, true,
                                (char*)workspace_dev, st);
}

}  // extern "C"
