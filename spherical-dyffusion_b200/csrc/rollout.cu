// Rollout step glue on the device.  The reference keeps a dict of named [B, H, W] fields per step, and
// between two calls of the sampler it normalises (normalizer.py:96-112), packs (packer.py:71-77), unpacks, overwrites
// one variable in a masked region (prescriber.py:68-95), denormalises and copies every step to the host
// (stepper_multistep.py:365-427).  Here the state stays packed [B][C][HW] on the GPU and each side of the sampler is
// ONE launch: normalise + pack on the way in, prescribe + denormalise on the way out.
#include "common.cuh"

namespace sfno {

// out[b][c][p] = (field_c[b][p] - mean[c]) / std[c]      field_c = fields[c], an independent [B][hw] tensor
__global__ void normalize_pack_kernel(const float* const* __restrict__ fields, const float* __restrict__ mean,
                                      const float* __restrict__ stdv, float* __restrict__ out, int C, int64_t hw, int vec_ok) {
  const int c = blockIdx.y, b = blockIdx.z;
  const float* __restrict__ src = fields[c] + (int64_t)b * hw;
  float* __restrict__ dst = out + ((int64_t)b * C + c) * hw;
  const float mu = mean ? mean[c] : 0.0f, sd = stdv ? stdv[c] : 1.0f;
  const int64_t n4 = vec_ok ? (hw >> 2) : 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(src)[i];
    reinterpret_cast<float4*>(dst)[i] = make_float4((v.x - mu) / sd, (v.y - mu) / sd, (v.z - mu) / sd, (v.w - mu) / sd);
  }
  for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = (src[i] - mu) / sd;
}

// One pass over the packed prediction gen[b][c][p] (normalised):
//   channel == prescribed:  interpolate ? mask * target + (1 - mask) * gen : (int(round(mask)) == mask_value ? target : gen)
//   written back in place (it seeds the next window) and, if denorm != nullptr, denorm = gen * std[c] + mean[c].
// target [B or 1][hw] is the normalised target of the prescribed variable, mask [B or 1][hw] the raw mask variable.
__global__ void prescribe_denormalize_kernel(float* __restrict__ gen, const float* __restrict__ target, int64_t target_bstride,
                                             const float* __restrict__ mask, int64_t mask_bstride, int prescribed, int mask_value,
                                             int interpolate, const float* __restrict__ mean, const float* __restrict__ stdv,
                                             float* __restrict__ denorm, int C, int64_t hw) {
  const int c = blockIdx.y, b = blockIdx.z;
  float* __restrict__ g = gen + ((int64_t)b * C + c) * hw;
  float* __restrict__ d = denorm ? denorm + ((int64_t)b * C + c) * hw : nullptr;
  const float mu = mean ? mean[c] : 0.0f, sd = stdv ? stdv[c] : 1.0f;
  const bool presc = c == prescribed;
  const float* __restrict__ t = presc ? target + (int64_t)b * target_bstride : nullptr;
  const float* __restrict__ m = presc ? mask + (int64_t)b * mask_bstride : nullptr;
  if (!presc && !d) return;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += (int64_t)gridDim.x * blockDim.x) {
    float v = g[i];
    if (presc) {
      const float mk = m[i];
      if (interpolate) v = mk * t[i] + (1.0f - mk) * v;
      else if ((int)rintf(mk) == mask_value) v = t[i];   // torch.round: half to even, as rintf
      g[i] = v;
    }
    if (d) d[i] = fmaf(v, sd, mu);
  }
}

}  // namespace sfno

using namespace sfno;

extern "C" {

int sfno_normalize_pack(const float* const* fields_dev, int channels, int batch, int64_t hw, const float* mean_dev,
                        const float* std_dev, float* out_dev, void* stream) {
  SFNO_CHECK_ARG(fields_dev && out_dev && channels > 0 && batch > 0 && hw > 0, "bad arguments");
  SFNO_CHECK_ARG(channels <= 65535 && batch <= 65535, "at most 65535 channels / samples");
  const int vec = 0;   // field pointers live in device memory: their alignment is unknown to the host
  const unsigned gx = (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div64(hw, 256), 64));
  normalize_pack_kernel<<<dim3(gx, (unsigned)channels, (unsigned)batch), 256, 0, (cudaStream_t)stream>>>(fields_dev, mean_dev, std_dev, out_dev,
                                                                                                        channels, hw, vec);
  return post_launch("normalize_pack");
}

int sfno_prescribe_denormalize(float* gen_norm_dev, const float* target_norm_dev, int64_t target_bstride, const float* mask_dev,
                               int64_t mask_bstride, int prescribed_channel, int mask_value, int interpolate, const float* mean_dev,
                               const float* std_dev, float* gen_denorm_dev, int channels, int batch, int64_t hw, void* stream) {
  SFNO_CHECK_ARG(gen_norm_dev && channels > 0 && batch > 0 && hw > 0, "bad arguments");
  SFNO_CHECK_ARG(channels <= 65535 && batch <= 65535, "at most 65535 channels / samples");
  SFNO_CHECK_ARG(prescribed_channel < channels, "prescribed channel %d out of range", prescribed_channel);
  SFNO_CHECK_ARG(prescribed_channel < 0 || (target_norm_dev && mask_dev), "the prescriber needs a target and a mask field");
  const unsigned gx = (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div64(hw, 256), 64));
  prescribe_denormalize_kernel<<<dim3(gx, (unsigned)channels, (unsigned)batch), 256, 0, (cudaStream_t)stream>>>(
      gen_norm_dev, target_norm_dev, target_bstride, mask_dev, mask_bstride, prescribed_channel, mask_value, interpolate, mean_dev, std_dev,
      gen_denorm_dev, channels, hw);
  return post_launch("prescribe_denormalize");
}

}  // extern "C"
