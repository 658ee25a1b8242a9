// Engine cross-check: runs one op of the hot path on synthetic operands through BOTH device engines (CUDA-core
// reference engine and tcgen05/TMA engine) and reports the largest difference: bf16 operands (kind::f16), or -- op kinds
// >= 100 -- fp32 operands holding TF32-exact values (kind::tf32), for which both engines compute the same products and
// differ only in the fp32 summation order.  Used by the GPU tests to localise tensor-core descriptor / pipeline bugs
// per operand layout; it is a device-vs-device check (the CUDA-core engine is the one pinned to the oracle).
#include <vector>

#include "common.cuh"
#include "engine.cuh"
#include "ops.cuh"

namespace sfno {

static __global__ void fill_bf16_kernel(bf16* p, int64_t n, uint32_t seed, float scale) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint32_t h = (uint32_t)i * 2654435761u ^ seed;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; h *= 3266489917u; h ^= h >> 16;
    p[i] = __float2bfloat16_rn(scale * ((float)(h & 0xFFFF) / 32768.0f - 1.0f));
  }
}
static __global__ void fill_f32_kernel(float* p, int64_t n, uint32_t seed, float scale) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint32_t h = (uint32_t)i * 2654435761u ^ seed;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; h *= 3266489917u; h ^= h >> 16;
    p[i] = scale * ((float)(h & 0xFFFF) / 32768.0f - 1.0f);
  }
}
static __global__ void round_tf32_selftest_kernel(float* p, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = tf32_rna(p[i]);
}
template <class T>
static __global__ void compare_kernel(const T* a, const T* b, int64_t n, unsigned int* out /*[2] float bits*/, unsigned long long* nbad) {
  float me = 0.0f, mr = 0.0f;
  unsigned long long bad = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float x = to_f32(a[i]), y = to_f32(b[i]);
    const float d = fabsf(x - y);
    if (!(d <= 3.0e38f)) { bad++; continue; }
    me = fmaxf(me, d);
    mr = fmaxf(mr, fabsf(x));
  }
  atomicMax(out, __float_as_uint(me));
  atomicMax(out + 1, __float_as_uint(mr));
  if (bad) atomicAdd(nbad, bad);
}

struct Scratch {
  std::vector<void*> ptrs;
  ~Scratch() { for (void* p : ptrs) cudaFree(p); }
  template <class T> T* get(int64_t n, bool zero = false) {
    void* p = nullptr;
    if (cudaMalloc(&p, (size_t)std::max<int64_t>(n, 1) * sizeof(T)) != cudaSuccess) return nullptr;
    ptrs.push_back(p);
    if (zero) cudaMemset(p, 0, (size_t)std::max<int64_t>(n, 1) * sizeof(T));
    return (T*)p;
  }
  bf16* rnd(int64_t n, uint32_t seed, float scale = 1.0f) {
    bf16* p = get<bf16>(n);
    if (p) fill_bf16_kernel<<<1024, 256>>>(p, n, seed, scale);
    return p;
  }
  float* rndf(int64_t n, uint32_t seed, float scale = 1.0f) {
    float* p = get<float>(n);
    if (p) fill_f32_kernel<<<256, 256>>>(p, n, seed, scale);
    return p;
  }
  // operand of element type T: bf16 values, or fp32 values that are exact in TF32
  template <class T> T* rnd_t(int64_t n, uint32_t seed, float scale = 1.0f) {
    if constexpr (std::is_same<T, bf16>::value) return rnd(n, seed, scale);
    else {
      float* p = rndf(n, seed, scale);
      if (p) round_tf32_selftest_kernel<<<256, 256>>>(p, n);
      return p;
    }
  }
};

template <class Op>
struct DefaultTc {
  int operator()(const Op& op) const { return launch_gemm_tc(op, 0, "selftest_tc"); }
};

template <class Op, class TOut, class SetOut, class TcLaunch = DefaultTc<Op>>
static int run_both(Op op, int64_t out_elems, SetOut set_out, Scratch& s, double* res, TcLaunch tc_launch = TcLaunch()) {
  TOut* o0 = s.get<TOut>(out_elems, true);
  TOut* o1 = s.get<TOut>(out_elems, true);
  unsigned int* stats = s.get<unsigned int>(2, true);
  unsigned long long* nbad = s.get<unsigned long long>(1, true);
  if (!o0 || !o1 || !stats || !nbad) return fail(SFNO_ERR_CUDA, "selftest: out of memory");
  set_out(op, o0);
  SFNO_TRY(launch_gemm_simt(op, 0, "selftest_simt"));
  set_out(op, o1);
  const bool tc_ok = TcTraits<Op>::eligible(op);   // (needs the output pointer: the epilogue's tensor maps are part of it)
  res[2] = tc_ok ? 1.0 : 0.0;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms = 0.0f;
  if (tc_ok) {
    SFNO_TRY(tc_launch(op));      // warm-up (+ correctness run)
    cudaEventRecord(e0, 0);
    for (int i = 0; i < 5; ++i) SFNO_TRY(tc_launch(op));
    cudaEventRecord(e1, 0);
  } else {
    SFNO_TRY(launch_gemm_simt(op, 0, "selftest_simt"));
  }
  compare_kernel<TOut><<<512, 256>>>(o0, o1, out_elems, stats, nbad);
  cudaError_t err = cudaDeviceSynchronize();
  if (tc_ok) { cudaEventElapsedTime(&ms, e0, e1); ms /= 5.0f; }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  if (err != cudaSuccess) return fail(SFNO_ERR_CUDA, "selftest: %s", cudaGetErrorString(err));
  unsigned int h[2];
  unsigned long long hb = 0;
  cudaMemcpy(h, stats, sizeof(h), cudaMemcpyDeviceToHost);
  cudaMemcpy(&hb, nbad, sizeof(hb), cudaMemcpyDeviceToHost);
  float fe, fr;
  memcpy(&fe, &h[0], 4); memcpy(&fr, &h[1], 4);
  res[0] = fe; res[1] = fr; res[3] = ms; res[4] = (double)hb;
  return SFNO_OK;
}

}  // namespace sfno

using namespace sfno;

template <class T>
static int selftest_impl(int op_kind, const int* d, double* res /*[5]*/) {
  Scratch s;
  const int Kr = 8;
  switch (op_kind) {
    case 0: {  // DFT: B, C, nlat, nlon, mmax
      const int B = d[0], C = d[1], nlat = d[2], nlon = d[3], mmax = d[4];
      const int Kp = round_up(nlat, Kr), Wp = round_up(nlon, Kr);
      OpDft<T> op{};
      op.G = B * C; op.M = 2 * mmax; op.N = nlat; op.K = nlon;
      op.Bm = s.template rnd_t<T>((int64_t)B * C * nlat * nlon, 1); op.A = s.template rnd_t<T>((int64_t)2 * mmax * Wp, 2, 0.1f); op.a_sk = 1; op.b_sk = 1;
      op.aff_a = s.rndf((int64_t)B * C, 3); op.aff_d = s.rndf((int64_t)B * C, 4);
      op.B = B; op.C = C; op.nlat = nlat; op.nlon = nlon; op.Kp = Kp; op.Wp = Wp; op.x_bstride = (int64_t)C * nlat * nlon;
      return run_both<OpDft<T>, T>(op, (int64_t)mmax * B * 2 * C * Kp, [](OpDft<T>& o, T* p) { o.f = p; }, s, res);
    }
    case 1: {  // LEG: B, C, nlat, lmax, mmax
      const int B = d[0], C = d[1], nlat = d[2], lmax = d[3], mmax = d[4];
      const int Kp = round_up(nlat, Kr);
      OpLeg<T> op{};
      op.G = mmax; op.M = lmax; op.N = B * 2 * C; op.K = nlat;
      op.Bm = s.template rnd_t<T>((int64_t)mmax * op.N * Kp, 5); op.A = s.template rnd_t<T>((int64_t)mmax * lmax * Kp, 6, 0.1f); op.a_sk = 1; op.b_sk = 1;
      op.Kp = Kp; op.lmax = lmax; op.mmax = mmax; op.triangular = d[5];
      return run_both<OpLeg<T>, T>(op, (int64_t)lmax * mmax * op.N, [](OpLeg<T>& o, T* p) { o.x = p; }, s, res);
    }
    case 2: {  // DHCONV: B, C, lmax, mmax
      const int B = d[0], C = d[1], lmax = d[2], mmax = d[3];
      OpDhconv<T> op{};
      op.G = lmax; op.M = mmax * B; op.N = 2 * C; op.K = 2 * C;
      op.Bm = s.template rnd_t<T>((int64_t)lmax * 4 * C * C, 7, 0.1f); op.A = s.template rnd_t<T>((int64_t)lmax * mmax * B * 2 * C, 8); op.a_sk = 1; op.b_sk = 1;
      op.B = B; op.lmax = lmax; op.mmax = mmax; op.triangular = d[4];
      return run_both<OpDhconv<T>, T>(op, (int64_t)lmax * mmax * B * 2 * C, [](OpDhconv<T>& o, T* p) { o.y = p; }, s, res);
    }
    case 3: {  // ILEG: B, C, nlat, lmax, mmax, x_layout
      const int B = d[0], C = d[1], nlat = d[2], lmax = d[3], mmax = d[4], xl = d[5] & 1;
      const int Kp = round_up(nlat, Kr), Lq = round_up(lmax, Kr);
      OpIleg<T> op{};
      op.triangular = (d[5] >> 1) & 1;
      op.G = mmax; op.M = B * 2 * C; op.N = nlat; op.K = lmax;
      op.A = s.template rnd_t<T>((int64_t)lmax * mmax * op.M, 10); op.Bm = s.template rnd_t<T>((int64_t)mmax * nlat * Lq, 9, 0.1f); op.b_sk = 1;
      if (xl) { op.a_goff = op.M; op.a_sk = (int64_t)mmax * op.M; } else { op.a_goff = (int64_t)lmax * op.M; op.a_sk = op.M; }
      op.B = B; op.C = C; op.Kp = Kp; op.Lq = Lq; op.nlat = nlat;
      return run_both<OpIleg<T>, T>(op, (int64_t)mmax * 2 * B * C * Kp, [](OpIleg<T>& o, T* p) { o.g_out = p; }, s, res);
    }
    case 4: {  // IDFT: B, C, nlat, nlon, mmax, epilogue bitmask (1 bias, 2 gelu, 4 add)
      const int B = d[0], C = d[1], nlat = d[2], nlon = d[3], mmax = d[4], epi = d[5];
      const int Kp = round_up(nlat, Kr), Kq2 = round_up(2 * mmax, Kr);
      OpIdft<T, T> op{};
      op.G = 1; op.M = B * C * Kp; op.N = nlon; op.K = 2 * mmax;
      op.A = s.template rnd_t<T>((int64_t)2 * mmax * op.M, 12); op.Bm = s.template rnd_t<T>((int64_t)nlon * Kq2, 11, 0.1f); op.a_sk = op.M; op.b_sk = 1;
      op.out_bstride = (int64_t)C * nlat * nlon;
      op.bias = (epi & 1) ? s.rndf(C, 13) : nullptr;
      op.add = (epi & 4) ? s.template rnd_t<T>((int64_t)B * C * nlat * nlon, 14) : nullptr; op.add_bstride = op.out_bstride;
      op.act = (epi & 2) ? SFNO_ACT_GELU : SFNO_ACT_NONE;
      op.C = C; op.nlat = nlat; op.nlon = nlon; op.Kp = Kp; op.Kq2 = Kq2; op.b_reps = 0;
      auto tc = [](const OpIdft<T, T>& o) {  // same compile-time specialisation as launch_idft
        const IdftArgs<T, T>& a = o;
        if (a.act == SFNO_ACT_GELU) return launch_gemm_tc(OpIdft<T, T, SFNO_ACT_GELU>(a), 0, "selftest_tc");
        return launch_gemm_tc(OpIdft<T, T, SFNO_ACT_NONE>(a), 0, "selftest_tc");
      };
      return run_both<OpIdft<T, T>, T>(op, (int64_t)B * C * nlat * nlon, [](OpIdft<T, T>& o, T* p) { o.out = p; }, s, res, tc);
    }
    case 5: case 6: {  // CONV: B, cin, cout, P, batched_w, epilogue bitmask (1 bias, 2 gelu, 4 res+affine, 8 pos, 16 dropout); 6: bf16 out
      const int B = d[0], cin = d[1], cout = d[2], P = d[3], bw = d[4], epi = d[5];
      const int ldw = round_up(cin, Kr);
      auto fill = [&](auto& op) {
        op.G = B; op.M = cout; op.N = P; op.K = cin;
        op.A = s.template rnd_t<T>((int64_t)(bw ? B : 1) * cout * ldw, 16, 0.1f); op.Bm = s.template rnd_t<T>((int64_t)B * cin * P, 15); op.a_sk = 1; op.b_sk = P;
        op.in_bstride = (int64_t)cin * P; op.w_bstride = bw ? (int64_t)cout * ldw : 0; op.ldw = ldw;
        op.bias = (epi & 1) ? s.rndf((int64_t)B * cout, 17) : nullptr; op.bias_bstride = cout;
        op.act = (epi & 2) ? SFNO_ACT_GELU : SFNO_ACT_NONE;
        op.drop_p = (epi & 16) ? 0.1f : 0.0f; op.seed = 1234; op.offset = 77; op.branch_scale = nullptr;
        op.res = (epi & 4) ? s.template rnd_t<T>((int64_t)B * cout * P, 18) : nullptr; op.res_bstride = (int64_t)cout * P;
        op.res_a = (epi & 4) ? s.rndf((int64_t)B * cout, 19) : nullptr; op.res_d = (epi & 4) ? s.rndf((int64_t)B * cout, 20) : nullptr;
        op.pos = (epi & 8) ? s.template rnd_t<T>((int64_t)cout * P, 21) : nullptr; op.out_bstride = (int64_t)cout * P;
      };
      if (op_kind == 5) {
        OpConv<T, float> op{};
        fill(op);
        auto tc = [](const OpConv<T, float>& o) {
          const ConvArgs<T, float>& a = o;
          if (a.drop_p == 0.0f && a.act == SFNO_ACT_GELU) return launch_gemm_tc(OpConv<T, float, SFNO_ACT_GELU, 0>(a), 0, "selftest_tc");
          if (a.drop_p == 0.0f && a.act == SFNO_ACT_NONE) return launch_gemm_tc(OpConv<T, float, SFNO_ACT_NONE, 0>(a), 0, "selftest_tc");
          return launch_gemm_tc(o, 0, "selftest_tc");
        };
        return run_both<OpConv<T, float>, float>(op, (int64_t)B * cout * P, [](OpConv<T, float>& o, float* p) { o.out = p; }, s, res, tc);
      }
      OpConv<T, T> op{};
      fill(op);
      auto tc = [](const OpConv<T, T>& o) {
        const ConvArgs<T, T>& a = o;
        if (a.drop_p == 0.0f && a.act == SFNO_ACT_GELU) return launch_gemm_tc(OpConv<T, T, SFNO_ACT_GELU, 0>(a), 0, "selftest_tc");
        if (a.drop_p == 0.0f && a.act == SFNO_ACT_NONE) return launch_gemm_tc(OpConv<T, T, SFNO_ACT_NONE, 0>(a), 0, "selftest_tc");
        return launch_gemm_tc(o, 0, "selftest_tc");
      };
      return run_both<OpConv<T, T>, T>(op, (int64_t)B * cout * P, [](OpConv<T, T>& o, T* p) { o.out = p; }, s, res, tc);
    }
    default:
      return fail(SFNO_ERR_INVALID_ARGUMENT, "unknown selftest op %d", op_kind);
  }
}

extern "C" int sfno_b200_selftest_gemm(int op_kind, const int* d, int nd, double* res /*[5]*/) {
  SFNO_CHECK_ARG(d && res && nd >= 6, "selftest needs 6 dims and a 5-element result");
  if (op_kind >= 100) {   // fp32 storage: CUDA-core FMA engine vs the tensor cores in kind::tf32
    Tf32Scope tf32(true);
    return selftest_impl<float>(op_kind - 100, d, res);
  }
  return selftest_impl<bf16>(op_kind, d, res);
}
