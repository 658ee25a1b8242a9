// Ensemble statistics kernels (spec: src/evaluation/metrics.py:166-175 ensemble_spread, :199-246 crps_ensemble).
// Members live on the rank that produced them; sums are all-reduced / members all-gathered by the host
// with NCCL (torch.distributed) between these calls -- nothing else crosses GPUs on this path.
#include "common.cuh"

namespace sfno {

// sums[0][i] += sum_e x[e][i];  sums[1][i] += sum_e x[e][i]^2      (bandwidth bound, float4 vectorised)
__global__ void ensemble_accumulate_kernel(const float* __restrict__ x, int E, int64_t n, float* __restrict__ sums) {
  const int64_t n4 = n >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 s = reinterpret_cast<float4*>(sums)[i];
    float4 q = reinterpret_cast<float4*>(sums + n)[i];
    for (int e = 0; e < E; ++e) {
      const float4 v = reinterpret_cast<const float4*>(x + (int64_t)e * n)[i];
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      q.x = fmaf(v.x, v.x, q.x); q.y = fmaf(v.y, v.y, q.y); q.z = fmaf(v.z, v.z, q.z); q.w = fmaf(v.w, v.w, q.w);
    }
    reinterpret_cast<float4*>(sums)[i] = s;
    reinterpret_cast<float4*>(sums + n)[i] = q;
  }
  // tail
  for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float s = sums[i], q = sums[n + i];
    for (int e = 0; e < E; ++e) { const float v = x[(int64_t)e * n + i]; s += v; q = fmaf(v, v, q); }
    sums[i] = s; sums[n + i] = q;
  }
}

__global__ void ensemble_finalize_kernel(const float* __restrict__ sums, int E, int64_t n, float* __restrict__ mean,
                                         float* __restrict__ var) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float m = sums[i] / (float)E;
    if (mean) mean[i] = m;
    if (var) var[i] = E > 1 ? fmaxf((sums[n + i] - (float)E * m * m) / (float)(E - 1), 0.0f) : 0.0f;
  }
}

// fair CRPS per grid point via the sorted form: sum_{i<j}|x_i-x_j| = sum_k (2k - E + 1) x_(k)
template <int MAXE>
__global__ void ensemble_crps_kernel(const float* __restrict__ x, const float* __restrict__ truth, int E, int64_t n,
                                     float* __restrict__ crps) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v[MAXE];
    const float y = truth[i];
    float skill = 0.0f;
#pragma unroll
    for (int e = 0; e < MAXE; ++e) {
      v[e] = e < E ? x[(int64_t)e * n + i] : 3.0e38f;
      if (e < E) skill += fabsf(v[e] - y);
    }
    // odd-even transposition sort (register resident, data independent)
#pragma unroll
    for (int pass = 0; pass < MAXE; ++pass) {
#pragma unroll
      for (int k = pass & 1; k + 1 < MAXE; k += 2) {
        const float lo = fminf(v[k], v[k + 1]), hi = fmaxf(v[k], v[k + 1]);
        v[k] = lo; v[k + 1] = hi;
      }
    }
    float spread = 0.0f;
#pragma unroll
    for (int k = 0; k < MAXE; ++k)
      if (k < E) spread = fmaf((float)(2 * k - E + 1), v[k], spread);
    crps[i] = skill / (float)E - (E > 1 ? spread / ((float)E * (float)(E - 1)) : 0.0f);
  }
}

}  // namespace sfno

using namespace sfno;

extern "C" {

int sfno_ensemble_accumulate(const float* members_dev, int members, int64_t n, float* sums_dev, void* stream) {
  SFNO_CHECK_ARG(members_dev && sums_dev && members > 0 && n > 0, "bad arguments");
  SFNO_CHECK_ARG(((uintptr_t)members_dev & 15) == 0 && ((uintptr_t)sums_dev & 15) == 0 && (n & 3) == 0, "pointers/size must be 16-byte aligned");
  ensemble_accumulate_kernel<<<(unsigned)std::min<int64_t>(ceil_div64(n / 4, 256), 148 * 16), 256, 0, (cudaStream_t)stream>>>(members_dev, members, n, sums_dev);
  return post_launch("ensemble_accumulate");
}

int sfno_ensemble_finalize(const float* sums_dev, int total_members, int64_t n, float* mean_dev, float* var_dev, void* stream) {
  SFNO_CHECK_ARG(sums_dev && total_members > 0 && n > 0, "bad arguments");
  ensemble_finalize_kernel<<<(unsigned)std::min<int64_t>(ceil_div64(n, 256), 148 * 16), 256, 0, (cudaStream_t)stream>>>(sums_dev, total_members, n, mean_dev, var_dev);
  return post_launch("ensemble_finalize");
}

int sfno_ensemble_crps(const float* members_dev, const float* truth_dev, int members, int64_t n, float* crps_dev, void* stream) {
  SFNO_CHECK_ARG(members_dev && truth_dev && crps_dev && members > 0 && n > 0, "bad arguments");
  if (members > 64) return fail(SFNO_ERR_UNSUPPORTED, "at most 64 members, got %d", members);
  const unsigned grid = (unsigned)std::min<int64_t>(ceil_div64(n, 128), 148 * 32);
  cudaStream_t st = (cudaStream_t)stream;
  if (members <= 8) ensemble_crps_kernel<8><<<grid, 128, 0, st>>>(members_dev, truth_dev, members, n, crps_dev);
  else if (members <= 16) ensemble_crps_kernel<16><<<grid, 128, 0, st>>>(members_dev, truth_dev, members, n, crps_dev);
  else if (members <= 32) ensemble_crps_kernel<32><<<grid, 128, 0, st>>>(members_dev, truth_dev, members, n, crps_dev);
  else ensemble_crps_kernel<64><<<grid, 128, 0, st>>>(members_dev, truth_dev, members, n, crps_dev);
  return post_launch("ensemble_crps");
}

}  // extern "C"
