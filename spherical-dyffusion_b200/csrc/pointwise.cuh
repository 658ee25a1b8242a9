// Bandwidth-bound and tiny kernels around the GEMMs: InstanceNorm statistics, norm/time affine
// coefficients, weight folding, time-embedding MLPs, dtype/layout conversion.
#pragma once
#include "common.cuh"

namespace sfno {

// ---- block reductions -----------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// sum over the block; result valid in all threads.  `sh` must hold 33 floats.
__device__ __forceinline__ float block_sum(float v, float* sh) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  if (w == 0) {
    float s = lane < nw ? sh[lane] : 0.0f;
    s = warp_sum(s);
    if (lane == 0) sh[32] = s;
  }
  __syncthreads();
  return sh[32];
}

// ---- InstanceNorm statistics: one CTA per (b, c) plane ------------------------------------------------------
// nn.InstanceNorm2d semantics (sfnonet.py:641-647): biased variance over H*W, exact two-pass (mean, then centred
// sum of squares).  The plane is read from HBM ONCE: each thread keeps its 16-byte vectors in registers between the
// passes (kVecs x 16 B x blockDim covers planes up to 64 Ki bf16 / 32 Ki fp32 elements per 512 threads); larger or
// unaligned planes take the second pass from L2 / HBM.
template <class T, int kVecs>
__global__ void __launch_bounds__(512) instance_stats_kernel(const T* __restrict__ x, int64_t bstride, int C,
                                                              int64_t hw, float eps, float* __restrict__ mean_out,
                                                              float* __restrict__ rstd_out) {
  __shared__ float sh[33];
  constexpr int kPer = 16 / (int)sizeof(T);
  const int bc = blockIdx.x, b = bc / C, c = bc - b * C;
  const T* p = x + (int64_t)b * bstride + (int64_t)c * hw;
  const bool in_regs = (hw % kPer == 0) && (((uintptr_t)p & 15) == 0) && (hw <= (int64_t)kVecs * kPer * blockDim.x);
  float s = 0.0f, q = 0.0f;
  if (in_regs) {
    const int64_t nvec = hw / kPer;
    uint4 v[kVecs];
#pragma unroll
    for (int i = 0; i < kVecs; ++i) {
      const int64_t idx = (int64_t)i * blockDim.x + threadIdx.x;
      v[i] = idx < nvec ? reinterpret_cast<const uint4*>(p)[idx] : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int i = 0; i < kVecs; ++i) {
      const T* e = reinterpret_cast<const T*>(&v[i]);
#pragma unroll
      for (int k = 0; k < kPer; ++k) s += to_f32(e[k]);  // padding vectors are zeros
    }
    const float mean = block_sum(s, sh) / (float)hw;
#pragma unroll
    for (int i = 0; i < kVecs; ++i) {
      const int64_t idx = (int64_t)i * blockDim.x + threadIdx.x;
      if (idx < nvec) {
        const T* e = reinterpret_cast<const T*>(&v[i]);
#pragma unroll
        for (int k = 0; k < kPer; ++k) { const float d = to_f32(e[k]) - mean; q = fmaf(d, d, q); }
      }
    }
    const float var = block_sum(q, sh) / (float)hw;
    if (threadIdx.x == 0) { mean_out[bc] = mean; rstd_out[bc] = rsqrtf(var + eps); }
    return;
  }
  for (int64_t i = threadIdx.x; i < hw; i += blockDim.x) s += to_f32(p[i]);
  const float mean = block_sum(s, sh) / (float)hw;
  for (int64_t i = threadIdx.x; i < hw; i += blockDim.x) {
    float d = to_f32(p[i]) - mean;
    q = fmaf(d, d, q);
  }
  const float var = block_sum(q, sh) / (float)hw;
  if (threadIdx.x == 0) {
    mean_out[bc] = mean;
    rstd_out[bc] = rsqrtf(var + eps);
  }
}
template <class T>
inline void launch_instance_stats(const T* x, int64_t bstride, int B, int C, int64_t hw, float eps, float* mean, float* rstd, cudaStream_t st) {
  // bf16: 16 vectors x 8 elements x 512 threads = 65536 elements; fp32: 32 vectors x 4 x 512 = 65536 elements
  instance_stats_kernel<T, (sizeof(T) == 2 ? 16 : 32)><<<B * C, 512, 0, st>>>(x, bstride, C, hw, eps, mean, rstd);
}

// ---- per-(b,c) affine of "InstanceNorm -> time_scale_shift" (sfnonet.py:280-287,292-299) -----------------
//   y = a*x + d,  a = gamma*rstd*(1+scale),  d = (beta - gamma*mean*rstd)*(1+scale) + shift
// mean/rstd == nullptr -> no normalisation; ts == nullptr -> no time conditioning.
static __global__ void norm_affine_kernel(const float* __restrict__ mean, const float* __restrict__ rstd,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   const float* __restrict__ ts, int64_t ts_bstride, int B, int C,
                                   float* __restrict__ a_out, float* __restrict__ d_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C) return;
  const int b = i / C, c = i - b * C;
  float a = 1.0f, d = 0.0f;
  if (mean) {
    const float g = gamma ? gamma[c] : 1.0f, be = beta ? beta[c] : 0.0f;
    a = g * rstd[i];
    d = be - g * mean[i] * rstd[i];
  }
  if (ts) {
    const float sc = ts[(int64_t)b * ts_bstride + c] + 1.0f, sf = ts[(int64_t)b * ts_bstride + C + c];
    a *= sc;
    d = d * sc + sf;
  }
  a_out[i] = a;
  d_out[i] = d;
}

// Same affine, with mean / variance reduced (fixed order, fp64) from the per-slice partial sums written by the
// tensor-core epilogues: part[(slice*2 + {0,1}) * rows + row]; one (b,c) has `rows_per_bc` consecutive rows
// (1 for the convolutions, Kp latitude rows for the inverse DFT).  count = number of pixels per (b,c).
static __global__ void __launch_bounds__(256) norm_affine_partials_kernel(
    const float* __restrict__ part, int slices, int64_t rows, int rows_per_bc, float count, float eps,
    const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ ts, int64_t ts_bstride, int B,
    int C, float* __restrict__ a_out, float* __restrict__ d_out) {
  // one warp per (b,c): lanes stride over the (slice, row) partials, fp64 accumulation, fixed-order shuffle tree
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= B * C) return;
  const int b = i / C, c = i - b * C;
  double s = 0.0, q = 0.0;
  const int total = slices * rows_per_bc;
  for (int t = lane; t < total; t += 32) {
    const int sl = t / rows_per_bc, r = t - sl * rows_per_bc;
    const float* ps = part + ((int64_t)sl * 2) * rows + (int64_t)i * rows_per_bc + r;
    s += (double)ps[0];
    q += (double)ps[rows];
  }
  s = warp_sum_d(s);
  q = warp_sum_d(q);
  if (lane != 0) return;
  const double mean = s / (double)count;
  double var = q / (double)count - mean * mean;
  var = var > 0.0 ? var : 0.0;
  const float rstd = rsqrtf((float)var + eps);
  const float g = gamma ? gamma[c] : 1.0f, be = beta ? beta[c] : 0.0f;
  float a = g * rstd;
  float d = be - g * (float)mean * rstd;
  if (ts) {
    const float sc = ts[(int64_t)b * ts_bstride + c] + 1.0f, sf = ts[(int64_t)b * ts_bstride + C + c];
    a *= sc;
    d = d * sc + sf;
  }
  a_out[i] = a;
  d_out[i] = d;
}

// rows_per_bc == 1 (convolution partials, layout [slice][{s,q}][B*C]): 8 consecutive (b,c) per CTA (256 CTAs at
// B*C = 2048 -- every SM takes part), each read a 32-byte sector; 128 lane groups stride over the slices, fixed-order
// shared-memory reduction (deterministic)
constexpr int NAP_BC = 8, NAP_STREAMS = 1024 / NAP_BC;
static __global__ void __launch_bounds__(1024) norm_affine_partials_conv_kernel(
    const float* __restrict__ part, int slices, int64_t rows, float count, float eps, const float* __restrict__ gamma,
    const float* __restrict__ beta, const float* __restrict__ ts, int64_t ts_bstride, int B, int C, float* __restrict__ a_out,
    float* __restrict__ d_out) {
  __shared__ double sh_s[NAP_STREAMS][NAP_BC], sh_q[NAP_STREAMS][NAP_BC];
  const int bcl = threadIdx.x % NAP_BC, stream = threadIdx.x / NAP_BC;
  const int i = blockIdx.x * NAP_BC + bcl;
  double s = 0.0, q = 0.0;
  if (i < B * C) {
#pragma unroll 4
    for (int sl = stream; sl < slices; sl += NAP_STREAMS) {
      const float* ps = part + ((int64_t)sl * 2) * rows + i;
      s += (double)ps[0];
      q += (double)ps[rows];
    }
  }
  sh_s[stream][bcl] = s; sh_q[stream][bcl] = q;
  __syncthreads();
  if (stream != 0 || i >= B * C) return;
  for (int k = 1; k < NAP_STREAMS; ++k) { s += sh_s[k][bcl]; q += sh_q[k][bcl]; }
  const int b = i / C, c = i - b * C;
  const double mean = s / (double)count;
  double var = q / (double)count - mean * mean;
  var = var > 0.0 ? var : 0.0;
  const float rstd = rsqrtf((float)var + eps);
  const float g = gamma ? gamma[c] : 1.0f, be = beta ? beta[c] : 0.0f;
  float a = g * rstd;
  float d = be - g * (float)mean * rstd;
  if (ts) {
    const float sc = ts[(int64_t)b * ts_bstride + c] + 1.0f, sf = ts[(int64_t)b * ts_bstride + C + c];
    a *= sc;
    d = d * sc + sf;
  }
  a_out[i] = a;
  d_out[i] = d;
}

// compose a time scale/shift given as separate [B*C] arrays onto an existing affine
static __global__ void time_affine_compose_kernel(float* __restrict__ a, float* __restrict__ d, const float* __restrict__ scale,
                                           const float* __restrict__ shift, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float sc = scale[i] + 1.0f;
  a[i] *= sc;
  d[i] = d[i] * sc + shift[i];
}

// ---- fold a per-(b,c) input affine into a 1x1-conv weight: conv(a*x+d) = (W diag(a)) x + (bias + W d) -----
// one CTA per (b, o); w [cout][cin] fp32 master -> wb [B][cout][ldw] (T), bb [B][cout] fp32
template <class T>
static __global__ void __launch_bounds__(128) fold_affine_weight_kernel(const float* __restrict__ w, const float* __restrict__ bias,
                                                                 const float* __restrict__ a, const float* __restrict__ d,
                                                                 int cout, int cin, int ldw, T* __restrict__ wb,
                                                                 float* __restrict__ bb, int round_tf32 = 0) {
  __shared__ float sh[33];
  const int b = blockIdx.x / cout, o = blockIdx.x - b * cout;
  const float* wr = w + (int64_t)o * cin;
  T* dst = wb + ((int64_t)b * cout + o) * ldw;
  float s = 0.0f;
  for (int c = threadIdx.x; c < ldw; c += blockDim.x) {
    float v = 0.0f;
    if (c < cin) {
      v = wr[c] * a[b * cin + c];
      s = fmaf(wr[c], d[b * cin + c], s);
    }
    if (round_tf32) v = tf32_rna(v);   // operand of a tf32 MMA
    dst[c] = from_f32<T>(v);
  }
  s = block_sum(s, sh);
  if (threadIdx.x == 0) bb[b * cout + o] = s + (bias ? bias[o] : 0.0f);
}

// ---- y[b][n] = act_out( sum_k W[n][k] * act_in(x[b][k]) + bias[n] ), one warp per output -------------------
// time_emb_mlp (misc.py:145-147) and the per-block time_mlp (sfnonet.py:210-213).  fp32 throughout.
static __global__ void small_linear_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                    const float* __restrict__ bias, float* __restrict__ y, int B, int N, int K,
                                    int act_in, int act_out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= B * N) return;
  const int b = warp / N, n = warp - b * N;
  const float* xr = x + (int64_t)b * K;
  const float* wr = w + (int64_t)n * K;
  float s = 0.0f;
  for (int k = lane; k < K; k += 32) s = fmaf(wr[k], apply_act(act_in, xr[k]), s);
  s = warp_sum(s);
  if (lane == 0) y[warp] = apply_act(act_out, s + (bias ? bias[n] : 0.0f));
}

// SinusoidalPosEmb (misc.py:21-33): emb[b] = [sin(t*f_j), cos(t*f_j)], f_j = exp(-j*ln(1e4)/(half-1))
static __global__ void sinusoidal_kernel(const float* __restrict__ time, float scaler, float shift, int B, int dim,
                                  float* __restrict__ emb) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int half = dim / 2;
  if (i >= B * half) return;
  const int b = i / half, j = i - b * half;
  const float t = time[b] * scaler + shift;
  const float step = (float)(-9.210340371976184 / (double)(half - 1));
  const float f = expf((float)j * step);
  const float e = t * f;
  emb[(int64_t)b * dim + j] = sinf(e);
  emb[(int64_t)b * dim + half + j] = cosf(e);
}

// DropPath factor per sample (drop_path.py:5-22): floor(keep + U) / keep
static __global__ void drop_path_scale_kernel(float* __restrict__ out, int B, float drop_prob, uint64_t seed, uint64_t offset,
                                              const uint64_t* __restrict__ rng_dev) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  if (rng_dev) { seed = rng_dev[0]; offset += rng_dev[1]; }   // device-resident Philox state (graph-safe)
  const float keep = 1.0f - drop_prob;
  const float u = philox_uniform(seed, offset, (uint64_t)b);
  out[b] = floorf(keep + u) / keep;
}

static __global__ void round_tf32_kernel(float* __restrict__ p, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = tf32_rna(p[i]);
}

// device-resident Philox state {seed, offset}: one forward owns SFNO_RNG_OFFSETS_PER_FORWARD consecutive offsets
constexpr unsigned SFNO_RNG_OFFSETS_PER_FORWARD = 4096;
static __global__ void rng_advance_kernel(uint64_t* state, uint64_t by) { state[1] += by; }

// Keep masks of one dropout site, one bit per element (bit j of byte i = element 8 i + j), identical to what
// dropout_keep_mask8 yields inline for the same Philox key: one Philox4x32-7 block per byte.  n8 = elements / 8.
static __global__ void dropout_mask_kernel(uint8_t* __restrict__ mask, int64_t n8, float p, uint64_t seed, uint64_t offset,
                                           const uint64_t* __restrict__ rng_dev) {
  if (rng_dev) { seed = rng_dev[0]; offset += rng_dev[1]; }
  const uint32_t thr = dropout_threshold16(p);
  const int64_t n64 = n8 >> 3;   // eight bytes per thread and store
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n64; i += (int64_t)gridDim.x * blockDim.x) {
    unsigned long long w = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) w |= (unsigned long long)dropout_keep_mask8(seed, offset, (uint64_t)(i * 8 + j) << 3, thr) << (8 * j);
    reinterpret_cast<unsigned long long*>(mask)[i] = w;
  }
  for (int64_t i = (n64 << 3) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x)
    mask[i] = (uint8_t)dropout_keep_mask8(seed, offset, (uint64_t)i << 3, thr);
}

// position-weighted checksum of the bit patterns of `count` fp32 tensors, out[t] pre-zeroed.  grid = (tensor, chunk): the
// tensor index is the FAST block index, so the blocks of all large tensors become resident together (with the tensor as
// the slow index the eight 94 MB spectral weights of an ACE net were streamed one after the other by 64 blocks each).
// 16-byte loads, four in flight per thread: the kernel streams ~0.85 GB of parameters of the ACE net per call.
static __global__ void __launch_bounds__(512) param_fingerprint_kernel(const float* const* __restrict__ ptrs, const int64_t* __restrict__ numel,
                                                                       unsigned long long* __restrict__ out) {
  const int t = blockIdx.x;
  const uint32_t* __restrict__ p = reinterpret_cast<const uint32_t*>(ptrs[t]);
  const int64_t n = numel[t];
  const unsigned long long kMul = 0x9E3779B97F4A7C15ull;
  unsigned long long acc = 0;
  auto mix = [&](uint32_t v, int64_t i) { acc += (unsigned long long)v * ((kMul * (unsigned long long)(i + 1)) | 1ull); };
  const int64_t tid = (int64_t)blockIdx.y * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.y * blockDim.x;
  if ((int64_t)blockIdx.y * blockDim.x >= n) return;   // block-uniform: this block lies past the end of a small tensor
  int64_t done = 0;
  if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
    const uint4* __restrict__ p4 = reinterpret_cast<const uint4*>(p);
    const int64_t n4 = n >> 2;
    int64_t i = tid;
    for (; i + 3 * nth < n4; i += 4 * nth) {
      const uint4 a = p4[i], b = p4[i + nth], c = p4[i + 2 * nth], d = p4[i + 3 * nth];
      mix(a.x, 4 * i); mix(a.y, 4 * i + 1); mix(a.z, 4 * i + 2); mix(a.w, 4 * i + 3);
      mix(b.x, 4 * (i + nth)); mix(b.y, 4 * (i + nth) + 1); mix(b.z, 4 * (i + nth) + 2); mix(b.w, 4 * (i + nth) + 3);
      mix(c.x, 4 * (i + 2 * nth)); mix(c.y, 4 * (i + 2 * nth) + 1); mix(c.z, 4 * (i + 2 * nth) + 2); mix(c.w, 4 * (i + 2 * nth) + 3);
      mix(d.x, 4 * (i + 3 * nth)); mix(d.y, 4 * (i + 3 * nth) + 1); mix(d.z, 4 * (i + 3 * nth) + 2); mix(d.w, 4 * (i + 3 * nth) + 3);
    }
    for (; i < n4; i += nth) {
      const uint4 a = p4[i];
      mix(a.x, 4 * i); mix(a.y, 4 * i + 1); mix(a.z, 4 * i + 2); mix(a.w, 4 * i + 3);
    }
    done = n4 << 2;
  }
  for (int64_t i = done + tid; i < n; i += nth) mix(p[i], i);
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ unsigned long long warp_sum[16];   // one atomic per block, not per warp (integer sums: order does not matter)
  if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long s = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += warp_sum[w];
    if (s != 0) atomicAdd(out + t, s);
  }
}

// ---- conversions -----------------------------------------------------------------------------------------
// strided copy/convert of [B][C][hw] planes: dst[b*dst_bstride + c*hw + i] = src[b*src_bstride + c*hw + i]
template <class TS, class TD>
static __global__ void convert_planes_kernel(const TS* __restrict__ src, int64_t src_bstride, TD* __restrict__ dst,
                                      int64_t dst_bstride, int64_t per_sample) {
  const int b = blockIdx.y;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < per_sample; i += (int64_t)gridDim.x * blockDim.x)
    dst[(int64_t)b * dst_bstride + i] = from_f32<TD>(to_f32(src[(int64_t)b * src_bstride + i]));
}

// fp32 -> bf16, 8 elements per thread (two float4 loads, one 16-byte store); per_sample % 8 == 0, 16-byte aligned
static __global__ void convert_planes_vec8_kernel(const float* __restrict__ src, int64_t src_bstride, bf16* __restrict__ dst,
                                                  int64_t dst_bstride, int64_t per_sample) {
  const int b = blockIdx.y;
  const float4* s4 = reinterpret_cast<const float4*>(src + (int64_t)b * src_bstride);
  uint4* d4 = reinterpret_cast<uint4*>(dst + (int64_t)b * dst_bstride);
  const int64_t nvec = per_sample >> 3;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 a = s4[2 * i], c = s4[2 * i + 1];
    __nv_bfloat162 h0 = __floats2bfloat162_rn(a.x, a.y), h1 = __floats2bfloat162_rn(a.z, a.w);
    __nv_bfloat162 h2 = __floats2bfloat162_rn(c.x, c.y), h3 = __floats2bfloat162_rn(c.z, c.w);
    uint4 u;
    u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
    u.z = *reinterpret_cast<uint32_t*>(&h2); u.w = *reinterpret_cast<uint32_t*>(&h3);
    d4[i] = u;
  }
}

// Fused concat + convert (BaseModel.concat_condition_if_needed, _base_model.py:166-192): up to 3 fp32 sources
// [B][c_k][hw] are written as channels [coff_k, coff_k + c_k) of dst [B][..][hw] (dst_bstride elements per sample).
struct ConcatParts {
  const float* src[3];
  int channels[3];
  int nparts;
};
template <class T>
static __global__ void concat_convert_kernel(ConcatParts parts, int64_t hw, T* __restrict__ dst, int64_t dst_bstride, int round_tf32 = 0) {
  const int b = blockIdx.y;
  int coff = 0;
  for (int k = 0; k < parts.nparts; ++k) {
    const int64_t per = (int64_t)parts.channels[k] * hw;
    const float* s = parts.src[k] + (int64_t)b * per;
    T* d = dst + (int64_t)b * dst_bstride + (int64_t)coff * hw;
    if (sizeof(T) == 2 && per % 8 == 0 && ((((uintptr_t)s) | ((uintptr_t)d)) & 15) == 0) {
      const float4* s4 = reinterpret_cast<const float4*>(s);
      uint4* d4 = reinterpret_cast<uint4*>(d);
      for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (per >> 3); i += (int64_t)gridDim.x * blockDim.x) {
        const float4 a = s4[2 * i], c = s4[2 * i + 1];
        __nv_bfloat162 h0 = __floats2bfloat162_rn(a.x, a.y), h1 = __floats2bfloat162_rn(a.z, a.w);
        __nv_bfloat162 h2 = __floats2bfloat162_rn(c.x, c.y), h3 = __floats2bfloat162_rn(c.z, c.w);
        uint4 u;
        u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
        u.z = *reinterpret_cast<uint32_t*>(&h2); u.w = *reinterpret_cast<uint32_t*>(&h3);
        d4[i] = u;
      }
    } else if (sizeof(T) == 4 && per % 4 == 0 && ((((uintptr_t)s) | ((uintptr_t)d)) & 15) == 0) {
      const float4* s4 = reinterpret_cast<const float4*>(s);
      float4* d4 = reinterpret_cast<float4*>(d);
      for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (per >> 2); i += (int64_t)gridDim.x * blockDim.x) {
        float4 v = s4[i];
        if (round_tf32) { v.x = tf32_rna(v.x); v.y = tf32_rna(v.y); v.z = tf32_rna(v.z); v.w = tf32_rna(v.w); }
        d4[i] = v;
      }
    } else {
      for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < per; i += (int64_t)gridDim.x * blockDim.x)
        d[i] = from_f32<T>(round_tf32 ? tf32_rna(s[i]) : s[i]);
    }
    coff += parts.channels[k];
  }
}

// fp32 row-major [rows][cols] -> T [rows][ld] with zero padding (conv / linear weights)
template <class T>
static __global__ void pack_rows_kernel(const float* __restrict__ src, int rows, int cols, int ld, T* __restrict__ dst) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)rows * ld) return;
  const int r = (int)(i / ld), c = (int)(i - (int64_t)r * ld);
  dst[i] = from_f32<T>(c < cols ? src[(int64_t)r * cols + c] : 0.0f);
}

// dhconv weight [cin][cout][L][2] fp32 (s2convolutions.py:146) -> packed real form [L][2*cout][2*cin] (T):
//   rows (ri', o), cols (ri, c):  [[wr, -wi], [wi, wr]]
template <class T>
static __global__ void pack_dhconv_weight_kernel(const float* __restrict__ w, int cin, int cout, int L, T* __restrict__ dst) {
  const int64_t total = (int64_t)L * 2 * cout * 2 * cin;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int kk = (int)(i % (2 * cin));
    int64_t r = i / (2 * cin);
    int mm = (int)(r % (2 * cout));
    int l = (int)(r / (2 * cout));
    int ri_in = kk / cin, c = kk - ri_in * cin;
    int ri_out = mm / cout, o = mm - ri_out * cout;
    const float* src = w + (((int64_t)c * cout + o) * L + l) * 2;
    float wr = src[0], wi = src[1];
    float v = (ri_out == ri_in) ? wr : (ri_out == 1 ? wi : -wi);
    dst[i] = from_f32<T>(v);
  }
}

// y = a[bc]*x + d[bc] (stand-alone InstanceNorm entry point)
static __global__ void affine_apply_kernel(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ a,
                                    const float* __restrict__ d, int64_t hw) {
  const int bc = blockIdx.y;
  const float aa = a[bc], dd = d[bc];
  const float* xp = x + (int64_t)bc * hw;
  float* yp = y + (int64_t)bc * hw;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  if ((hw & 3) == 0 && ((reinterpret_cast<uintptr_t>(xp) | reinterpret_cast<uintptr_t>(yp)) & 15) == 0) {   // 16-byte path
    const float4* x4 = reinterpret_cast<const float4*>(xp);
    float4* y4 = reinterpret_cast<float4*>(yp);
    for (int64_t i = tid; i < (hw >> 2); i += nth) {
      float4 v = x4[i];
      v.x = fmaf(aa, v.x, dd); v.y = fmaf(aa, v.y, dd); v.z = fmaf(aa, v.z, dd); v.w = fmaf(aa, v.w, dd);
      y4[i] = v;
    }
  } else {
    for (int64_t i = tid; i < hw; i += nth) yp[i] = fmaf(aa, xp[i], dd);
  }
}

// reference coefficient layout <-> internal spectral layout (B = 1, C = fields):
//   coeffs [fields][lmax][mmax][2]  <->  X [lmax][mmax][2][fields]
template <class T>
static __global__ void coeffs_to_internal_kernel(const float* __restrict__ coeffs, T* __restrict__ x, int fields, int lmax, int mmax) {
  const int64_t total = (int64_t)fields * lmax * mmax * 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int f = (int)(i % fields);
    int64_t r = i / fields;
    int ri = (int)(r & 1);
    int64_t lm = r >> 1;  // l*mmax + m
    x[i] = from_f32<T>(coeffs[((int64_t)f * lmax * mmax + lm) * 2 + ri]);
  }
}
template <class T>
static __global__ void internal_to_coeffs_kernel(const T* __restrict__ x, float* __restrict__ coeffs, int fields, int lmax, int mmax) {
  const int64_t total = (int64_t)fields * lmax * mmax * 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int ri = (int)(i & 1);
    int64_t r = i >> 1;
    int64_t lm = r % ((int64_t)lmax * mmax);
    int f = (int)(r / ((int64_t)lmax * mmax));
    coeffs[i] = to_f32(x[(lm * 2 + ri) * fields + f]);
  }
}

// diagonal operator on the internal layouts: X[l][m][b][ri][c] -> Y[l][m][b][ri][o], w[i][o][l][m][2]
template <class T>
static __global__ void diag_contract_internal_kernel(const T* __restrict__ X, const float2* __restrict__ w, T* __restrict__ Y,
                                              int B, int C, int L, int M) {
  const int64_t total = (int64_t)L * M * B * C;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int o = (int)(idx % C);
    int64_t r = idx / C;
    const int b = (int)(r % B); r /= B;
    const int m = (int)(r % M);
    const int l = (int)(r / M);
    const T* xr = X + (((int64_t)l * M + m) * B + b) * 2 * C;
    float re = 0.0f, im = 0.0f;
    for (int i = 0; i < C; ++i) {
      const float xa = to_f32(xr[i]), xb = to_f32(xr[C + i]);
      const float2 wv = w[(((int64_t)i * C + o) * L + l) * M + m];
      re = fmaf(xa, wv.x, re); re = fmaf(-xb, wv.y, re);
      im = fmaf(xa, wv.y, im); im = fmaf(xb, wv.x, im);
    }
    T* yr = Y + (((int64_t)l * M + m) * B + b) * 2 * C;
    yr[o] = from_f32<T>(re);
    yr[C + o] = from_f32<T>(im);
  }
}


}  // namespace sfno
