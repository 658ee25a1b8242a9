// GEMM "ops": every contraction on the SFNO hot path expressed as a batched GEMM
//     D[g][m][n] = sum_k A[g](m,k) * B[g](n,k)          (fp32 accumulate)
// plus a fused epilogue.  An op describes operand addressing (row offset + k stride), the problem size and the
// epilogue; the engines (gemm_simt.cuh: fp32 CUDA cores; gemm_tc.cuh: tcgen05 + TMA) are templated on the op.
//
// In the tensor-core engine one thread owns one output row m (a TMEM lane) and walks the columns n.  Every op is
// oriented so that the COLUMN index n is the contiguous index of the output tensor: the thread produces 16-byte vectors
// along its own row, every per-row quantity (bias, affine, decoded indices) lives in registers, the element-wise math
// runs on packed fp32 pairs (compute8()), and eight consecutive rows form one box of a TMA store (io_coords()).
//
// Internal tensor layouts (T = float or bf16; Kp = nlat rounded up to 8):
//   grid   x  [B][C][nlat][nlon]                      (NCHW, as the reference)
//   F, G      [mmax][B][2][C][Kp]   /  [mmax][2][B][C][Kp]     longitude-spectral, latitude contiguous
//   X, Y      [lmax][mmax][B][2][C]                             spectral, channel contiguous
#pragma once
#include "common.cuh"

namespace sfno {

// GELU for the bf16 epilogues: 0.5 x (1 + tanh(x (a + b x^2))) with (a, b) fitted to the exact erf form (max abs
// deviation 2.7e-4) and the MUFU tanh (rel. error 2^-11): together < 1/4 of a bf16 rounding step of the result, at
// 7 instructions instead of ~20 for an erf evaluation.  The fp32 path keeps the exact erff form (gelu_exact).
__device__ __forceinline__ float gelu_fast(float x) {
  const float x2 = x * x;
  const float u = x * fmaf(0.03470089f, x2, 0.80015708f);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
  const float h = 0.5f * x;
  return fmaf(h, t, h);
}
template <class T>
__device__ __forceinline__ float act_for(int act, float x) {
  if constexpr (sizeof(T) == 2) {
    if (act == SFNO_ACT_GELU) return gelu_fast(x);
  }
  return apply_act(act, x);
}
// ACT >= 0: activation fixed at compile time (bf16 tensor-core instantiations); ACT < 0: runtime value
template <class T, int ACT>
__device__ __forceinline__ float act_ct(int act_rt, float x) {
  if constexpr (ACT == SFNO_ACT_NONE) return x;
  else if constexpr (ACT == SFNO_ACT_GELU) return sizeof(T) == 2 ? gelu_fast(x) : gelu_exact(x);
  else return act_for<T>(act_rt, x);
}

// packed-pair forms of gelu_fast / act_ct (same formula, two elements per instruction)
__device__ __forceinline__ f2 gelu_fast2(f2 x) {
  const f2 x2 = f2_mul(x, x);
  const f2 u = f2_mul(x, f2_fma(f2_splat(0.03470089f), x2, f2_splat(0.80015708f)));
  float u0, u1, t0, t1;
  f2_get(u, u0, u1);
  asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(u0));
  asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(u1));
  const f2 h = f2_mul(x, f2_splat(0.5f));
  return f2_fma(h, f2_make(t0, t1), h);
}
template <class T, int ACT>
__device__ __forceinline__ f2 act_ct2(int act_rt, f2 v) {
  if constexpr (ACT == SFNO_ACT_NONE) return v;
  else if constexpr (ACT == SFNO_ACT_GELU && sizeof(T) == 2) return gelu_fast2(v);
  else {
    float lo, hi;
    f2_get(v, lo, hi);
    return f2_make(act_ct<T, ACT>(act_rt, lo), act_ct<T, ACT>(act_rt, hi));
  }
}

struct NoSplitBox {   // 32 consecutive GEMM rows never straddle two planes of the output tensor
  static constexpr bool kSplitBox = false;
  __device__ bool box_straddles(int) const { return false; }
};
struct NoFeatures : NoSplitBox {
  static constexpr int kStagingBufs = 1;   // load-bound ops: shared memory goes to the operand ring
  static constexpr int kFast0 = 0, kFast1 = 0;
  static constexpr bool kGeneral = false;
  __device__ int feat() const { return 0; }
  __device__ bool has_res() const { return false; }
  __device__ void res_coords(int, int, int, int (&)[5]) const {}
  __device__ bool wants_stats() const { return false; }
  __device__ void finish(int, int, int, float, float) const {}
};
// fp32 outputs that a later tf32 MMA consumes are rounded to TF32 (round-to-nearest) by the tensor-core epilogue: the
// MMA itself truncates, which biases every product towards zero.  Ops carry `round_out` (0 for tensors that leave the
// library: final outputs keep full fp32).  bf16 storage ignores it.

__device__ __forceinline__ void load_vec8(const bf16* p, float (&v)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x; v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void load_vec8(const float* p, float (&v)[8]) {
  const float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

// Per-group index ranges.  The Legendre tables are exactly zero for l < m, so for the zonal
// wavenumber m only degrees l >= m carry information.  All three spectral ops use the SAME predicate
//     (l, m) is live  <=>  l >= live_l0(m)        (the triangle, rounded to the 64-wide K block of the engine)
// the forward transform stores exactly the live degrees (rows), dhconv visits exactly the live wavenumbers of a degree
// (m <= l | 63) and the inverse transform contracts over exactly the live degrees.  Hence every X / Y entry that is
// ever read was written in the same forward: nothing depends on the previous contents of the workspace.
// first live degree of wavenumber m, clamped to the last 64-block of degrees so that at least one block is live
__host__ __device__ inline int live_l0(int m, int lmax) {
  const int l0 = m & ~63, last = lmax > 0 ? ((lmax - 1) & ~63) : 0;
  return l0 < last ? l0 : last;
}
struct FullRanges {
  static constexpr bool kGFastest = false;
  static constexpr bool kRanged = false;  // every tile of the M x N x G grid is live
  static constexpr bool kSimtRowsOnFastLanes = false;
  __device__ int n_begin(int) const { return 0; }
  __device__ int k_begin(int) const { return 0; }
  __device__ int m_begin(int) const { return 0; }
};

// ------------------------------------------------------------------------------------------------
// forward longitude DFT (K1 of SURVEY 2.3), one GEMM per (sample, channel): the basis rows (m,ri) are the GEMM
// rows, the latitudes k of one channel plane the columns, so the contiguous index of F (the latitude) is the
// column index and the InstanceNorm/time affine of the plane is a pair of tile-wide scalars:
//   F[m][b][ri][c][k] = a[b,c] * sum_j E[(m,ri)][j] x[b][c][k][j] + d[b,c] * 2*pi * [m==0, re]   (DFT(1) = 2*pi*delta_m0)
// ------------------------------------------------------------------------------------------------
template <class T>
struct OpDft : FullRanges, NoFeatures {
  static constexpr bool A_KCONTIG = true, B_KCONTIG = true, kColContig = true, kNFastest = false;
  using OutT = T;
  using InT = T;
  int round_out;
  __device__ bool out_tf32() const { return round_out != 0; }
  __device__ int n_end(int) const { return N; }
  __device__ int m_end(int) const { return M; }
  int G, M, N, K;  // G = B*C, M = 2*mmax, N = nlat, K = nlon
  const T* A; const T* Bm; int64_t a_sk, b_sk;   // A = basis [2*mmax][Wp], Bm = x [B][C][nlat][nlon]
  T* f;
  const float* aff_a; const float* aff_d;  // [B*C] or nullptr
  int B, C, nlat, nlon, Kp, Wp;
  int64_t x_bstride;
  int a_reps;   // replicas of the basis, [a_reps][2*mmax][Wp] (0 / 1: a single copy)

  __device__ int64_t a_off(int, int m) const { return (int64_t)m * Wp; }
  __device__ int64_t b_off(int g, int n) const {
    const int b = g / C, c = g - b * C;
    return (int64_t)b * x_bstride + ((int64_t)c * nlat + n) * nlon;
  }
  __device__ int n_store() const { return Kp; }
  // TMA view of F: {k, c, ri, b, m}; eight GEMM rows (m,ri) = 4 wavenumbers x {re, im} of one (b, c) plane
  __device__ void io_coords(int g, int row0, int col0, int (&c)[5]) const {
    const int b = g / C;
    c[0] = col0; c[1] = g - b * C; c[2] = 0; c[3] = b; c[4] = row0 >> 1;
  }
  struct Row { T* out; const T* res; bool valid; float a, d; __device__ float stat_s() const { return 0.0f; } __device__ float stat_q() const { return 0.0f; } };
  __device__ Row row(int g, int m) const {
    const int b = g / C, c = g - b * C, mm = m >> 1, ri = m & 1;
    Row r;
    r.out = f + (((int64_t)mm * B + b) * 2 + ri) * C * Kp + (int64_t)c * Kp;
    r.res = nullptr; r.valid = true;
    r.a = aff_a ? aff_a[g] : 1.0f;
    r.d = (aff_d && m == 0) ? aff_d[g] * 6.28318530717958647692f : 0.0f;
    return r;
  }
  template <int F> __device__ Row row_f(int g, int m) const { return row(g, m); }
  __device__ void store(const Row& r, int, int, int n, float acc) const { r.out[n] = from_f32<T>(fmaf(r.a, acc, r.d)); }
  template <int F>
  __device__ void compute8(Row& r, int n, const float (&acc)[8], const float (&)[8], float (&o)[8]) const {
    const f2 a2 = f2_splat(r.a);
#pragma unroll
    for (int j = 0; j < 4; ++j) f2_get(f2_mul(a2, f2_make(acc[2 * j], acc[2 * j + 1])), o[2 * j], o[2 * j + 1]);
    if (r.d != 0.0f) {  // the single (m = 0, re) row; the pad latitudes [nlat, Kp) stay zero
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] += (n + i < nlat) ? r.d : 0.0f;
    }
  }
};

// ------------------------------------------------------------------------------------------------
// forward Legendre (K2): per m, table rows l x latitude  times  F rows (b,ri,c) x latitude -> X[l][m][b][ri][c]
//   (degree l = GEMM row, the channel-contiguous index (b,ri,c) = GEMM column = contiguous output index)
// ------------------------------------------------------------------------------------------------
template <class T>
struct OpLeg : NoFeatures {
  static constexpr bool kGFastest = false;  // (group-fastest tile order measured slower: dhconv 1.01 -> 1.37 ms per forward)
  static constexpr bool kSimtRowsOnFastLanes = false;
  static constexpr bool kRanged = true;
  static constexpr bool A_KCONTIG = true, B_KCONTIG = true, kColContig = true, kNFastest = false;
  using OutT = T;
  using InT = T;
  int round_out;
  __device__ bool out_tf32() const { return round_out != 0; }
  int triangular;  // 1: store only the live degrees l >= live_l0(m)
  __device__ int m_begin(int g) const { return triangular ? live_l0(g, M) : 0; }
  __device__ int m_end(int) const { return M; }
  __device__ int n_begin(int) const { return 0; }
  __device__ int n_end(int) const { return N; }
  __device__ int k_begin(int) const { return 0; }
  int G, M, N, K;  // G = mmax, M = lmax, N = B*2*C, K = nlat
  const T* A; const T* Bm; int64_t a_sk, b_sk;   // A = wq [mmax][lmax][Kp], Bm = F [mmax][(b,ri,c)][Kp]
  T* x;
  int Kp, lmax, mmax;
  __device__ int64_t a_off(int g, int m) const { return ((int64_t)g * lmax + m) * Kp; }
  __device__ int64_t b_off(int g, int n) const { return ((int64_t)g * N + n) * Kp; }
  __device__ int n_store() const { return N; }
  // TMA view of X: {(b,ri,c), m, l}; eight GEMM rows = eight degrees of wavenumber g
  __device__ void io_coords(int g, int row0, int col0, int (&c)[5]) const { c[0] = col0; c[1] = g; c[2] = row0; c[3] = 0; c[4] = 0; }
  struct Row { T* out; const T* res; bool valid; __device__ float stat_s() const { return 0.0f; } __device__ float stat_q() const { return 0.0f; } };
  __device__ Row row(int g, int m) const { return Row{x + ((int64_t)m * mmax + g) * N, nullptr, true}; }
  template <int F> __device__ Row row_f(int g, int m) const { return row(g, m); }
  __device__ void store(const Row& r, int, int, int n, float acc) const { r.out[n] = from_f32<T>(acc); }
  template <int F>
  __device__ void compute8(Row&, int, const float (&acc)[8], const float (&)[8], float (&o)[8]) const {
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = acc[i];
  }
};

// ------------------------------------------------------------------------------------------------
// dhconv channel contraction (K3) as a REAL GEMM on the packed complex weight, per degree l:
//   D[(m,b), (ri',o)] = sum_(ri,c) X[l][m][b][(ri,c)] * Wp[l][(ri',o)][(ri,c)]  -> Y[l][m][b][ri'][o]
//   Wp = [[wr, -wi], [wi, wr]] (rows = output re/im, cols = input re/im)
// ------------------------------------------------------------------------------------------------
template <class T>
struct OpDhconv : NoFeatures {
  static constexpr bool kGFastest = false;  // (group-fastest tile order measured slower: dhconv 1.01 -> 1.37 ms per forward)
  static constexpr bool kSimtRowsOnFastLanes = false;
  static constexpr bool kRanged = true;
  static constexpr bool A_KCONTIG = true, B_KCONTIG = true, kColContig = true, kNFastest = true;
  using OutT = T;
  using InT = T;
  int round_out;
  __device__ bool out_tf32() const { return round_out != 0; }
  int triangular;  // 1: only the live wavenumbers of degree l
  __device__ int m_begin(int) const { return 0; }
  __device__ int m_end(int g) const {
    const int last = lmax > 0 ? ((lmax - 1) & ~63) : 0;  // degrees of the last block see every wavenumber (live_l0 clamp)
    if (!triangular || g >= last) return M;
    const int e = ((g | 63) + 1) * B;
    return e < M ? e : M;
  }
  __device__ int n_begin(int) const { return 0; }
  __device__ int n_end(int) const { return N; }
  __device__ int k_begin(int) const { return 0; }
  int G, M, N, K;  // G = lmax, M = mmax*B, N = 2*Cout, K = 2*Cin
  const T* A; const T* Bm; int64_t a_sk, b_sk;   // A = X [lmax][(m,b)][2Cin], Bm = Wp [lmax][2Cout][2Cin]
  T* y;
  int B, lmax, mmax;
  __device__ int64_t a_off(int g, int m) const { return ((int64_t)g * M + m) * K; }
  __device__ int64_t b_off(int g, int n) const { return ((int64_t)g * N + n) * K; }
  __device__ int n_store() const { return N; }
  // TMA view of Y: {(ri',o), (m,b), l}
  __device__ void io_coords(int g, int row0, int col0, int (&c)[5]) const { c[0] = col0; c[1] = row0; c[2] = g; c[3] = 0; c[4] = 0; }
  struct Row { T* out; const T* res; bool valid; __device__ float stat_s() const { return 0.0f; } __device__ float stat_q() const { return 0.0f; } };
  __device__ Row row(int g, int m) const { return Row{y + ((int64_t)g * M + m) * N, nullptr, true}; }
  template <int F> __device__ Row row_f(int g, int m) const { return row(g, m); }
  __device__ void store(const Row& r, int, int, int n, float acc) const { r.out[n] = from_f32<T>(acc); }
  template <int F>
  __device__ void compute8(Row&, int, const float (&acc)[8], const float (&)[8], float (&o)[8]) const {
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = acc[i];
  }
};

// ------------------------------------------------------------------------------------------------
// inverse Legendre (K4): per m, rows (b,ri,o), columns = latitude k (contiguous in G):
//   D[(b,ri,o), k] = sum_l S[m][l][(b,ri,o)] * Pt[m][k][l]  -> G[m][ri][b][o][k]
//   S is X or Y, [l][m][rows] (a_goff = rows, a_sk = mmax*rows); any (a_goff, a_sk) pair is accepted
// ------------------------------------------------------------------------------------------------
template <class T>
struct OpIleg : NoFeatures {
  static constexpr bool kGFastest = false;  // (group-fastest tile order measured slower: dhconv 1.01 -> 1.37 ms per forward)
  static constexpr bool kSimtRowsOnFastLanes = false;
  static constexpr bool kRanged = true;
  static constexpr bool A_KCONTIG = false, B_KCONTIG = true, kColContig = true, kNFastest = true;
  using OutT = T;
  using InT = T;
  int round_out;
  __device__ bool out_tf32() const { return round_out != 0; }
  int triangular;  // 1: contract only over the live degrees l >= (m & ~63)
  __device__ int n_begin(int) const { return 0; }
  __device__ int n_end(int) const { return N; }
  __device__ int k_begin(int g) const { return triangular ? live_l0(g, K) : 0; }
  __device__ int m_begin(int) const { return 0; }
  __device__ int m_end(int) const { return M; }
  int G, M, N, K;  // G = mmax, M = B*2*C, N = nlat, K = lmax
  const T* A; const T* Bm; int64_t a_sk, b_sk;
  int64_t a_goff;
  T* g_out;
  int B, C, Kp, Lq, nlat;
  __device__ int64_t a_off(int g, int m) const { return (int64_t)g * a_goff + m; }   // + l * a_sk
  __device__ int64_t b_off(int g, int n) const { return ((int64_t)g * nlat + n) * Lq; }
  __device__ int n_store() const { return Kp; }  // columns [nlat, Kp) are exact zeros (zero-filled table rows)
  // TMA view of G: {k, o, b, ri, m}; eight GEMM rows (b,ri,o) = eight channels (C % 8 == 0)
  __device__ void io_coords(int g, int row0, int col0, int (&c)[5]) const {
    const int b = row0 / (2 * C), rem = row0 - b * 2 * C, ri = rem / C;
    c[0] = col0; c[1] = rem - ri * C; c[2] = b; c[3] = ri; c[4] = g;
  }
  struct Row { T* out; const T* res; bool valid; __device__ float stat_s() const { return 0.0f; } __device__ float stat_q() const { return 0.0f; } };
  __device__ Row row(int g, int m) const {
    int b = m / (2 * C), rem = m - b * 2 * C;
    int ri = rem / C, o = rem - ri * C;
    return Row{g_out + (int64_t)g * 2 * B * C * Kp + (int64_t)ri * B * C * Kp + ((int64_t)b * C + o) * Kp, nullptr, true};
  }
  template <int F> __device__ Row row_f(int g, int m) const { return row(g, m); }
  __device__ void store(const Row& r, int, int, int n, float acc) const { r.out[n] = from_f32<T>(acc); }
  template <int F>
  __device__ void compute8(Row&, int, const float (&acc)[8], const float (&)[8], float (&o)[8]) const {
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = acc[i];
  }
};

// ------------------------------------------------------------------------------------------------
// inverse longitude DFT (K5): rows (b,o,kp), columns = longitude j (contiguous in the output):
//   D[(b,o,kp), j] = sum_(m,ri) G[(m,ri)][(b,o,kp)] * Einv[j][(m,ri)]
//   epilogue: + bias[o] + add[b][o][k][j] -> act -> out[b][o][k][j]   (bias of SpectralConvS2, inner-skip sum and
//   GELU of FourierNeuralOperatorBlock.forward fused: s2convolutions.py:188-189, sfnonet.py:308-311)
// ------------------------------------------------------------------------------------------------
template <class T, class TOut>
struct IdftArgs {
  int G, M, N, K;  // G = 1, M = B*C*Kp, N = nlon, K = 2*mmax
  const T* A; const T* Bm; int64_t a_sk, b_sk;
  TOut* out; int64_t out_bstride;
  const float* bias;                      // [C] or nullptr
  const T* add; int64_t add_bstride;      // [B][C][nlat][nlon] or nullptr
  int act;
  int C, nlat, nlon, Kp, Kq2;
  int b_reps;   // replicas of the basis, [b_reps][nlon][Kq2] (0 / 1: a single copy)
  int round_out;  // fp32 output feeds a tf32 MMA: round to TF32 in the tensor-core epilogue
  // optional fused InstanceNorm statistics of the OUTPUT: per (column slice, row) partial sum / sum of squares,
  // stat_part[(slice*2 + {0,1}) * M + row]; reduced in a fixed order by norm_affine_partials_kernel (deterministic)
  float* stat_part;
};
template <class T, class TOut, int ACT = -1>
struct OpIdft : IdftArgs<T, TOut>, FullRanges {
  static constexpr bool A_KCONTIG = false, B_KCONTIG = true, kColContig = true, kNFastest = true;
  __device__ int n_end(int) const { return this->N; }
  __device__ int m_end(int) const { return this->M; }
  using OutT = TOut;
  using InT = T;
  __device__ bool out_tf32() const { return this->round_out != 0; }
  using Args = IdftArgs<T, TOut>;
  OpIdft() = default;
  __host__ __device__ explicit OpIdft(const Args& a) : Args(a) {}
  __device__ int64_t a_off(int, int m) const { return m; }  // + kk * a_sk
  __device__ int64_t b_off(int, int n) const { return (int64_t)n * this->Kq2; }
  __device__ int n_store() const { return this->N; }
  __device__ bool has_res() const { return this->add != nullptr; }
  static constexpr int kFast0 = 0, kFast1 = F_RES | F_STATS;
  static constexpr bool kGeneral = true;
  // 2 = the TMA store of one tile reads staging buffer b while the next tile fills b^1.  Measured slower (fc1 0.22 ->
  // 0.255 ms, inverse DFT 0.20 -> 0.246 ms): the second buffer costs one operand stage (4 -> 3) and these kernels do
  // wait for operands 8-20 % of the time, which outweighs the hidden store latency.
  // fp32 in / fp32 out (tf32 mode): two buffers, one per 32-column pass, so that a residual / addend block of the whole
  // 64-column slice can be staged by TMA before the accumulator is ready.
  static constexpr int kStagingBufs = (sizeof(T) == 4 && sizeof(TOut) == 4) ? 2 : 1;
  __device__ int feat() const { return (this->add ? F_RES : 0) | (this->stat_part ? F_STATS : 0); }
  // TMA view of the grid tensor: {j, k, o, b}; eight GEMM rows (b,o,kp) = eight latitudes of one plane (Kp % 8 == 0);
  // the pad latitudes kp >= nlat are clipped by the tensor extent
  __device__ void io_coords(int, int row0, int col0, int (&c)[5]) const {
    const int bo = row0 / this->Kp, b = bo / this->C;
    c[0] = col0; c[1] = row0 - bo * this->Kp; c[2] = bo - b * this->C; c[3] = b; c[4] = 0;
  }
  __device__ void res_coords(int g, int row0, int col0, int (&c)[5]) const { io_coords(g, row0, col0, c); }
  // A plane (b, o) holds Kp GEMM rows (nlat rounded up to 8: the pad rows have no destination and are clipped by the
  // tensor extent).  The 32 rows of a warp may run past the plane; those warps store in 8-row boxes, which never do.
  static constexpr bool kSplitBox = true;
  __device__ bool box_straddles(int row0) const { return row0 % this->Kp + 32 > this->Kp; }
  struct Row {
    TOut* out; const T* res; bool valid; float bias; f2 s2, q2;   // s2 / q2: packed partial sum / sum of squares
    __device__ float stat_s() const { return f2_hsum(s2); }
    __device__ float stat_q() const { return f2_hsum(q2); }
  };
  __device__ Row row(int, int m) const {
    int bo = m / this->Kp, k = m - bo * this->Kp;
    int b = bo / this->C, o = bo - b * this->C;
    Row r;
    r.s2 = f2_splat(0.0f); r.q2 = f2_splat(0.0f);
    r.valid = k < this->nlat;
    const int64_t pix = ((int64_t)o * this->nlat + k) * this->nlon;
    r.out = this->out + (int64_t)b * this->out_bstride + pix;
    r.res = this->add ? this->add + (int64_t)b * this->add_bstride + pix : nullptr;
    r.bias = this->bias ? this->bias[o] : 0.0f;
    return r;
  }
  template <int F> __device__ Row row_f(int g, int m) const { return row(g, m); }
  __device__ void store(const Row& r, int, int, int n, float acc) const {
    if (!r.valid) return;
    float v = acc + r.bias;
    if (r.res) v += to_f32(r.res[n]);
    r.out[n] = from_f32<TOut>(act_ct<T, ACT>(this->act, v));
  }
  template <int F>
  __device__ void compute8(Row& r, int, const float (&acc)[8], const float (&res)[8], float (&o)[8]) const {
    const f2 b2 = f2_splat(r.bias);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      f2 v = f2_add(f2_make(acc[2 * j], acc[2 * j + 1]), b2);
      if (feat_on<F, F_RES>(true)) v = f2_add(v, f2_make(res[2 * j], res[2 * j + 1]));  // (res[] is zero when no addend is staged)
      v = act_ct2<T, ACT>(this->act, v);
      if (feat_on<F, F_STATS>(this->stat_part != nullptr)) { r.s2 = f2_add(r.s2, v); r.q2 = f2_fma(v, v, r.q2); }
      f2_get(v, o[2 * j], o[2 * j + 1]);
    }
  }
  __device__ bool wants_stats() const { return this->stat_part != nullptr; }
  // called once per (row, N tile) with the row's sums over that tile's columns (zeros for pad rows)
  __device__ void finish(int, int m, int slice, float s, float q) const {
    this->stat_part[((int64_t)slice * 2) * this->M + m] = s;
    this->stat_part[((int64_t)slice * 2 + 1) * this->M + m] = q;
  }
};

// ------------------------------------------------------------------------------------------------
// 1x1 convolution (K6) per sample, output channel o is the row, pixel p the (contiguous) column:
//   D[o, p] = sum_c w[(b)][o][c] * in[b][c][p]
// epilogue (all optional): + bias[(b)][o] -> act -> dropout -> * branch_scale[b]
//                          + residual (optionally affine: ra[b,o]*res + rd[b,o]) + pos[o][p]
// ------------------------------------------------------------------------------------------------
template <class T, class TOut>
struct ConvArgs {
  int G, M, N, K;  // G = batch, M = cout, N = hw, K = cin
  const T* A; const T* Bm; int64_t a_sk, b_sk;   // A = weights, Bm = input activations (b_sk = hw)
  int64_t in_bstride;
  int64_t w_bstride; int ldw;            // w_bstride = 0 for shared weights
  const float* bias; int64_t bias_bstride;
  int act;
  float drop_p; uint64_t seed, offset;    // dropout (after act), drop_p = 0 -> off
  // graph-safe Philox state: when set, the stream key is {rng_dev[0], rng_dev[1] + offset} read on the device, so a
  // captured CUDA graph draws fresh masks on every replay (the forward advances rng_dev[1] after its last kernel)
  const uint64_t* rng_dev;
  // optional precomputed keep mask, one bit per output element in the order ((g * M + m) * N + n), bit j of byte i =
  // element 8 i + j: exactly the bits dropout_keep_mask8 would produce inline.  A dedicated kernel generates them at
  // full occupancy (dropout_mask_kernel); inline, Philox costs as many issue slots as the rest of the epilogue.
  const uint8_t* drop_mask;
  const float* branch_scale;              // [batch] DropPath factor (0 or 1/keep) or nullptr
  const T* res; int64_t res_bstride;      // residual [b][cout][hw] or nullptr
  const float* res_a; const float* res_d; // [batch*cout] affine on the residual or nullptr
  const T* pos;                           // [cout][hw] or nullptr
  TOut* out; int64_t out_bstride;
  int round_out;  // fp32 output feeds a tf32 MMA: round to TF32 in the tensor-core epilogue
  // optional fused InstanceNorm statistics of the OUTPUT: stat_part[(slice*2 + {0,1}) * (G*M) + g*M + m]
  float* stat_part;
};
// ACT / DROP < 0: decided at run time (CUDA-core engine and rarely used combinations); DROP = 0: dropout off; DROP = 2:
// dropout on with the keep bits read from drop_mask -- both at compile time (the tensor-core instantiations of the MLP
// with inference dropout keep the packed GELU, carry no Philox code and lose the per-element run-time tests: the
// run-time variant cost 3.4 vs 1.65 ms per forward in fc1)
template <class T, class TOut, int ACT = -1, int DROP = -1>
struct OpConv : ConvArgs<T, TOut>, FullRanges, NoSplitBox {
  static constexpr bool kSimtRowsOnFastLanes = true;
  static constexpr bool A_KCONTIG = true, B_KCONTIG = false, kColContig = true, kNFastest = false;
  __device__ int n_end(int) const { return this->N; }
  __device__ int m_end(int) const { return this->M; }
  using OutT = TOut;
  using InT = T;
  __device__ bool out_tf32() const { return this->round_out != 0; }
  using Args = ConvArgs<T, TOut>;
  OpConv() = default;
  __host__ __device__ explicit OpConv(const Args& a) : Args(a) {}
  __device__ int64_t a_off(int g, int m) const { return (int64_t)g * this->w_bstride + (int64_t)m * this->ldw; }
  __device__ int64_t b_off(int g, int n) const { return (int64_t)g * this->in_bstride + n; }  // + c * hw
  __device__ int n_store() const { return this->N; }
  __device__ bool has_res() const { return this->res != nullptr; }
  static constexpr int kFast0 = 0, kFast1 = F_RES | F_STATS;
  static constexpr bool kGeneral = true;
  // 2 = the TMA store of one tile reads staging buffer b while the next tile fills b^1.  Measured slower (fc1 0.22 ->
  // 0.255 ms, inverse DFT 0.20 -> 0.246 ms): the second buffer costs one operand stage (4 -> 3) and these kernels do
  // wait for operands 8-20 % of the time, which outweighs the hidden store latency.
  // fp32 in / fp32 out (tf32 mode): two buffers, one per 32-column pass, so that a residual / addend block of the whole
  // 64-column slice can be staged by TMA before the accumulator is ready.
  static constexpr int kStagingBufs = (sizeof(T) == 4 && sizeof(TOut) == 4) ? 2 : 1;
  __device__ int feat() const {
    return (this->res ? F_RES : 0) | (this->stat_part ? F_STATS : 0) | (this->pos ? F_POS : 0) | (this->branch_scale ? F_SCALE : 0);
  }
  // TMA view of the activation tensor: {pixel, channel, sample}
  __device__ void io_coords(int g, int row0, int col0, int (&c)[5]) const { c[0] = col0; c[1] = row0; c[2] = g; c[3] = 0; c[4] = 0; }
  __device__ void res_coords(int g, int row0, int col0, int (&c)[5]) const {   // a residual shared by all samples has one plane
    io_coords(this->res_bstride != 0 ? g : 0, row0, col0, c);
  }
  struct Row {
    TOut* out; const T* res; bool valid; const T* pos; float bias, ra, rd, scale; uint64_t rng_base, seed, offset; f2 s2, q2;
    __device__ float stat_s() const { return f2_hsum(s2); }
    __device__ float stat_q() const { return f2_hsum(q2); }
  };
  template <int F>
  __device__ Row row_f(int g, int m) const {
    Row r;
    const int64_t off = (int64_t)m * this->N;
    r.valid = true;
    r.s2 = f2_splat(0.0f); r.q2 = f2_splat(0.0f);
    r.out = this->out + (int64_t)g * this->out_bstride + off;
    r.bias = this->bias ? this->bias[(int64_t)g * this->bias_bstride + m] : 0.0f;
    r.res = nullptr; r.ra = 1.0f; r.rd = 0.0f; r.pos = nullptr; r.scale = 1.0f; r.rng_base = 0; r.seed = 0; r.offset = 0;
    if (feat_on<F, F_RES>(this->res != nullptr)) {
      r.res = this->res + (int64_t)g * this->res_bstride + off;
      if (this->res_a) r.ra = this->res_a[g * this->M + m];
      if (this->res_d) r.rd = this->res_d[g * this->M + m];
    }
    if (feat_on<F, F_POS>(this->pos != nullptr)) r.pos = this->pos + off;
    if (feat_on<F, F_SCALE>(this->branch_scale != nullptr)) r.scale = this->branch_scale[g];
    if (dropping()) {   // the 1/keep factor of the dropout rides on the DropPath scale
      r.rng_base = ((uint64_t)g * this->M + m) * (uint64_t)this->N;
      r.scale *= dropout_keep_scale(dropout_threshold16(this->drop_p));
      if constexpr (DROP != 2) {   // inline Philox key (the mask variant never evaluates it)
        r.seed = this->rng_dev ? this->rng_dev[0] : this->seed;
        r.offset = this->offset + (this->rng_dev ? this->rng_dev[1] : 0ull);
      }
    }
    return r;
  }
  __device__ Row row(int g, int m) const { return row_f<-1>(g, m); }
  __device__ bool dropping() const {
    if constexpr (DROP == 0) return false;
    else if constexpr (DROP == 2) return true;
    else return this->drop_p > 0.0f;
  }
  __device__ void store(const Row& r, int, int, int n, float acc) const {
    float v = act_ct<T, ACT>(this->act, acc + r.bias) * r.scale;
    if (dropping() && !dropout_keep(r.seed, r.offset, r.rng_base + n, dropout_threshold16(this->drop_p))) v = 0.0f;
    if (r.res) v += fmaf(r.ra, to_f32(r.res[n]), r.rd);
    if (r.pos) v += to_f32(r.pos[n]);
    r.out[n] = from_f32<TOut>(v);
  }
  template <int F>
  __device__ void compute8(Row& r, int n, const float (&acc)[8], const float (&res)[8], float (&o)[8]) const {
    uint32_t keep = 0xffu;
    if constexpr (DROP == 2) {
      keep = (uint32_t)__ldg(this->drop_mask + ((r.rng_base + (uint64_t)n) >> 3));
    } else if (dropping()) {
      keep = this->drop_mask ? (uint32_t)__ldg(this->drop_mask + ((r.rng_base + (uint64_t)n) >> 3))
                             : dropout_keep_mask8(r.seed, r.offset, r.rng_base + n, dropout_threshold16(this->drop_p));
    }
    const f2 b2 = f2_splat(r.bias), sc2 = f2_splat(r.scale), ra2 = f2_splat(r.ra), rd2 = f2_splat(r.rd);
    float t[8];
    if (feat_on<F, F_POS>(r.pos != nullptr)) load_vec8(r.pos + n, t);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      f2 v = act_ct2<T, ACT>(this->act, f2_add(f2_make(acc[2 * j], acc[2 * j + 1]), b2));
      if (dropping())   // (r.scale carries 1/keep and the DropPath factor)
        v = f2_mul(v, f2_make(((keep >> (2 * j)) & 1u) ? r.scale : 0.0f, ((keep >> (2 * j + 1)) & 1u) ? r.scale : 0.0f));
      else if (feat_on<F, F_SCALE>(true)) v = f2_mul(v, sc2);   // (r.scale = 1 without DropPath)
      if (feat_on<F, F_RES>(r.res != nullptr)) v = f2_add(v, f2_fma(ra2, f2_make(res[2 * j], res[2 * j + 1]), rd2));
      if (feat_on<F, F_POS>(r.pos != nullptr)) v = f2_add(v, f2_make(t[2 * j], t[2 * j + 1]));
      if (feat_on<F, F_STATS>(this->stat_part != nullptr)) { r.s2 = f2_add(r.s2, v); r.q2 = f2_fma(v, v, r.q2); }
      f2_get(v, o[2 * j], o[2 * j + 1]);
    }
  }
  __device__ bool wants_stats() const { return this->stat_part != nullptr; }
  __device__ void finish(int g, int m, int slice, float s, float q) const {
    const int64_t rows = (int64_t)this->G * this->M, idx = (int64_t)g * this->M + m;
    this->stat_part[((int64_t)slice * 2) * rows + idx] = s;
    this->stat_part[((int64_t)slice * 2 + 1) * rows + idx] = q;
  }
};

}  // namespace sfno
