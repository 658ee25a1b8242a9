// Ensemble statistics kernels (spec: src/evaluation/metrics.py:166-175 ensemble_spread, :199-246 crps_ensemble).
// Members live on the rank that produced them; moments are all-reduced / members all-gathered by the host
// with NCCL (torch.distributed) between these calls -- nothing else crosses GPUs on this path.
//
// Variance follows the reference's two-pass `predicted.var(dim=0)`.  Sharded form: (1) all-reduce of the raw local sums
// gives a pivot p = sum / E that is bit-identical on every rank and within a few ulp of the mean; (2) every rank
// accumulates the SHIFTED moments S1 = sum (x - p), S2 = sum (x - p)^2 of its members (small numbers: no cancellation
// however large |mean| / spread is -- surface pressure, temperature), all-reduced; (3) mean = p + S1 / E,
// var = (S2 - S1^2 / E) / (E - 1).  Raw sum-of-squares moments are never formed.
#include "common.cuh"

namespace sfno {

// sum[i] = sum_e x[e][i]   (E members of THIS rank; E may be 0)
__global__ void ensemble_local_sum_kernel(const float* __restrict__ x, int E, int64_t n, float* __restrict__ sum, int vec_ok) {
  const int64_t n4 = vec_ok ? (n >> 2) : 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 s = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    for (int e = 0; e < E; ++e) {
      const float4 v = reinterpret_cast<const float4*>(x + (int64_t)e * n)[i];
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    reinterpret_cast<float4*>(sum)[i] = s;
  }
  for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float s = 0.0f;
    for (int e = 0; e < E; ++e) s += x[(int64_t)e * n + i];
    sum[i] = s;
  }
}

// mom[0][i] = sum_e (x[e][i] - p[i]),  mom[1][i] = sum_e (x[e][i] - p[i])^2  with the pivot p = sum_global / E_total
__global__ void ensemble_shifted_moments_kernel(const float* __restrict__ x, int E, int64_t n, const float* __restrict__ sum_global,
                                                int E_total, float* __restrict__ mom, int vec_ok) {
  const float inv = 1.0f / (float)E_total;
  const int64_t n4 = vec_ok ? (n >> 2) : 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 g = reinterpret_cast<const float4*>(sum_global)[i];
    const float4 p = make_float4(g.x * inv, g.y * inv, g.z * inv, g.w * inv);
    float4 s = make_float4(0.0f, 0.0f, 0.0f, 0.0f), q = s;
    for (int e = 0; e < E; ++e) {
      const float4 v = reinterpret_cast<const float4*>(x + (int64_t)e * n)[i];
      const float d0 = v.x - p.x, d1 = v.y - p.y, d2 = v.z - p.z, d3 = v.w - p.w;
      s.x += d0; s.y += d1; s.z += d2; s.w += d3;
      q.x = fmaf(d0, d0, q.x); q.y = fmaf(d1, d1, q.y); q.z = fmaf(d2, d2, q.z); q.w = fmaf(d3, d3, q.w);
    }
    reinterpret_cast<float4*>(mom)[i] = s;
    reinterpret_cast<float4*>(mom + n)[i] = q;
  }
  for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float p = sum_global[i] * inv;
    float s = 0.0f, q = 0.0f;
    for (int e = 0; e < E; ++e) { const float d = x[(int64_t)e * n + i] - p; s += d; q = fmaf(d, d, q); }
    mom[i] = s; mom[n + i] = q;
  }
}

// mean = p + S1 / E,  var = (S2 - S1^2 / E) / (E - 1)   (S1 is tiny: the subtraction does not cancel)
__global__ void ensemble_finalize_kernel(const float* __restrict__ sum_global, const float* __restrict__ mom, int E, int64_t n,
                                         float* __restrict__ mean, float* __restrict__ var) {
  const float inv = 1.0f / (float)E;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float s1 = mom[i], s2 = mom[n + i];
    if (mean) mean[i] = fmaf(s1, inv, sum_global[i] * inv);
    if (var) var[i] = E > 1 ? fmaxf(s2 - s1 * s1 * inv, 0.0f) / (float)(E - 1) : 0.0f;
  }
}

// One pass over ALL members of a grid point held in registers: ensemble mean, unbiased two-pass variance and the fair
// CRPS via the sorted form  sum_{i<j}|x_i-x_j| = sum_k (2k - E + 1) x_(k).  truth / crps may be NULL (moments only).
// rows: optional [E] row indices into x (the all-gathered, padded member buffer of uneven shards); nullptr = 0 .. E-1
template <int MAXE>
__global__ void ensemble_stats_kernel(const float* __restrict__ x, const int* __restrict__ rows, const float* __restrict__ truth, int E,
                                      int64_t n, float* __restrict__ mean, float* __restrict__ var, float* __restrict__ crps) {
  __shared__ int srow[MAXE];
  for (int e = threadIdx.x; e < MAXE; e += blockDim.x) srow[e] = (e < E) ? (rows ? rows[e] : e) : 0;
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v[MAXE];
    const float y = truth ? truth[i] : 0.0f;
    float skill = 0.0f, s = 0.0f;
#pragma unroll
    for (int e = 0; e < MAXE; ++e) {
      v[e] = e < E ? x[(int64_t)srow[e] * n + i] : 3.0e38f;
      if (e < E) { skill += fabsf(v[e] - y); s += v[e]; }
    }
    const float mu = s / (float)E;
    if (mean) mean[i] = mu;
    if (var) {
      float q = 0.0f;
#pragma unroll
      for (int e = 0; e < MAXE; ++e)
        if (e < E) { const float d = v[e] - mu; q = fmaf(d, d, q); }
      var[i] = E > 1 ? q / (float)(E - 1) : 0.0f;
    }
    if (!crps) continue;
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

// x_out = x_s + (x_next - x_cur): the cold-sampling update of BaseDYffusion.sample_loop (dyffusion.py:519) in one pass
// (three reads, one write; x_out may alias x_s).  x_cur == nullptr: x_out = x_next + 0 * ... is not needed -- the first
// step of the loop has x_cur = x_s, i.e. x_out = x_next, which the host expresses without a kernel.
__global__ void cold_update_kernel(const float* xs, const float* __restrict__ xnext, const float* __restrict__ xcur,
                                   float* out, int64_t n, int vec_ok) {
  const int64_t n4 = vec_ok ? (n >> 2) : 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 a = reinterpret_cast<const float4*>(xs)[i], b = reinterpret_cast<const float4*>(xnext)[i],
                 c = reinterpret_cast<const float4*>(xcur)[i];
    reinterpret_cast<float4*>(out)[i] = make_float4(a.x + (b.x - c.x), a.y + (b.y - c.y), a.z + (b.z - c.z), a.w + (b.w - c.w));
  }
  for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = xs[i] + (xnext[i] - xcur[i]);
}

}  // namespace sfno

using namespace sfno;

static inline bool al16(const void* p) { return ((uintptr_t)p & 15) == 0; }
static inline unsigned stream_grid(int64_t work, int per_block) {
  return (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div64(work, per_block), 148 * 16));
}

extern "C" {

int sfno_ensemble_local_sum(const float* members_dev, int members, int64_t n, float* sum_dev, void* stream) {
  SFNO_CHECK_ARG(sum_dev && members >= 0 && n > 0 && (members == 0 || members_dev), "bad arguments");
  const int vec = al16(members_dev) && al16(sum_dev) && (n & 3) == 0;
  ensemble_local_sum_kernel<<<stream_grid(vec ? n / 4 : n, 256), 256, 0, (cudaStream_t)stream>>>(members_dev, members, n, sum_dev, vec);
  return post_launch("ensemble_local_sum");
}

int sfno_ensemble_shifted_moments(const float* members_dev, int members, int64_t n, const float* sum_global_dev, int total_members,
                                  float* moments_dev, void* stream) {
  SFNO_CHECK_ARG(sum_global_dev && moments_dev && members >= 0 && total_members > 0 && n > 0 && (members == 0 || members_dev), "bad arguments");
  const int vec = al16(members_dev) && al16(sum_global_dev) && al16(moments_dev) && (n & 3) == 0;
  ensemble_shifted_moments_kernel<<<stream_grid(vec ? n / 4 : n, 256), 256, 0, (cudaStream_t)stream>>>(members_dev, members, n, sum_global_dev,
                                                                                                total_members, moments_dev, vec);
  return post_launch("ensemble_shifted_moments");
}

int sfno_ensemble_finalize(const float* sum_global_dev, const float* moments_dev, int total_members, int64_t n, float* mean_dev,
                           float* var_dev, void* stream) {
  SFNO_CHECK_ARG(sum_global_dev && moments_dev && total_members > 0 && n > 0, "bad arguments");
  ensemble_finalize_kernel<<<stream_grid(n, 256), 256, 0, (cudaStream_t)stream>>>(sum_global_dev, moments_dev, total_members, n, mean_dev, var_dev);
  return post_launch("ensemble_finalize");
}

int sfno_ensemble_stats_rows(const float* members_dev, const int* rows_dev, const float* truth_dev, int members, int64_t n, float* mean_dev,
                             float* var_dev, float* crps_dev, void* stream);

int sfno_ensemble_stats(const float* members_dev, const float* truth_dev, int members, int64_t n, float* mean_dev, float* var_dev,
                        float* crps_dev, void* stream) {
  return sfno_ensemble_stats_rows(members_dev, nullptr, truth_dev, members, n, mean_dev, var_dev, crps_dev, stream);
}

int sfno_ensemble_stats_rows(const float* members_dev, const int* rows_dev, const float* truth_dev, int members, int64_t n, float* mean_dev,
                             float* var_dev, float* crps_dev, void* stream) {
  SFNO_CHECK_ARG(members_dev && members > 0 && n > 0, "bad arguments");
  SFNO_CHECK_ARG(crps_dev == nullptr || truth_dev != nullptr, "the CRPS needs a truth field");
  if (members > 64) return fail(SFNO_ERR_UNSUPPORTED, "at most 64 members, got %d", members);
  const unsigned grid = stream_grid(n, 128) * 2;
  cudaStream_t st = (cudaStream_t)stream;
  if (members <= 8) ensemble_stats_kernel<8><<<grid, 128, 0, st>>>(members_dev, rows_dev, truth_dev, members, n, mean_dev, var_dev, crps_dev);
  else if (members <= 16) ensemble_stats_kernel<16><<<grid, 128, 0, st>>>(members_dev, rows_dev, truth_dev, members, n, mean_dev, var_dev, crps_dev);
  else if (members <= 32) ensemble_stats_kernel<32><<<grid, 128, 0, st>>>(members_dev, rows_dev, truth_dev, members, n, mean_dev, var_dev, crps_dev);
  else ensemble_stats_kernel<64><<<grid, 128, 0, st>>>(members_dev, rows_dev, truth_dev, members, n, mean_dev, var_dev, crps_dev);
  return post_launch("ensemble_stats");
}

int sfno_ensemble_crps(const float* members_dev, const float* truth_dev, int members, int64_t n, float* crps_dev, void* stream) {
  SFNO_CHECK_ARG(truth_dev && crps_dev, "bad arguments");
  return sfno_ensemble_stats(members_dev, truth_dev, members, n, nullptr, nullptr, crps_dev, stream);
}

int sfno_cold_update(const float* x_s_dev, const float* x_next_dev, const float* x_cur_dev, float* x_out_dev, int64_t n, void* stream) {
  SFNO_CHECK_ARG(x_s_dev && x_next_dev && x_cur_dev && x_out_dev && n > 0, "bad arguments");
  const int vec = al16(x_s_dev) && al16(x_next_dev) && al16(x_cur_dev) && al16(x_out_dev) && (n & 3) == 0;
  cold_update_kernel<<<stream_grid(vec ? n / 4 : n, 256), 256, 0, (cudaStream_t)stream>>>(x_s_dev, x_next_dev, x_cur_dev, x_out_dev, n, vec);
  return post_launch("cold_update");
}

}  // extern "C"
