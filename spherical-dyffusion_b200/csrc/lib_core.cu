// Library plumbing, host tables entry point, SHT plan, stand-alone ops of the C ABI.
#include <cstring>
#include <vector>

#include "common.cuh"
#include "engine.cuh"
#include "ops.cuh"
#include "pointwise.cuh"
#include "sht_plan.cuh"
#include "tables.h"

namespace sfno {

std::atomic<int64_t> g_launch_count{0};
std::atomic<int> g_profile_on{0};
std::atomic<int> g_nvtx_on{[] { const char* e = getenv("SFNO_NVTX"); return (e && e[0] && e[0] != '0') ? 1 : 0; }()};

// Per-launch timing: one CUDA event after every launch on the profiled stream; durations are differences of
// consecutive events (the path is a single in-order stream, so launches execute back to back).
struct Profiler {
  cudaStream_t stream = nullptr;
  std::vector<cudaEvent_t> events;
  std::vector<std::string> names;
};
static Profiler g_prof;

void profile_mark(const char* what) {
  cudaEvent_t e;
  if (cudaEventCreate(&e) != cudaSuccess) return;
  cudaEventRecord(e, g_prof.stream);
  g_prof.events.push_back(e);
  g_prof.names.emplace_back(what);
}

std::string& last_error_ref() {
  static thread_local std::string s;
  return s;
}

int fail(int status, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  last_error_ref() = buf;
  return status;
}

// ---- table upload -------------------------------------------------------------------------------------
// round-to-nearest (ties away, as cvt.rna.tf32.f32) of an fp32 value to the 10-bit TF32 mantissa
static float round_tf32_host(float v) {
  uint32_t u;
  memcpy(&u, &v, 4);
  u = (u + 0x1000u) & 0xFFFFE000u;
  memcpy(&v, &u, 4);
  return v;
}
static thread_local bool g_upload_tf32 = false;   // tables of a SFNO_PREC_TF32 plan are stored TF32-exact

template <class T>
static int upload_padded(const std::vector<double>& src, int64_t rows, int cols, int ld, void** dst, size_t* bytes, int replicas = 1) {
  std::vector<T> host((size_t)rows * ld);
  for (int64_t r = 0; r < rows; ++r)
    for (int c = 0; c < ld; ++c) {
      float v = c < cols ? (float)src[(size_t)r * cols + c] : 0.0f;
      if (g_upload_tf32) v = round_tf32_host(v);
      if constexpr (std::is_same<T, float>::value) host[(size_t)r * ld + c] = v;
      else host[(size_t)r * ld + c] = __float2bfloat16_rn(v);
    }
  const size_t nbytes = host.size() * sizeof(T);
  SFNO_CUDA(cudaMalloc(dst, nbytes * replicas));
  for (int r = 0; r < replicas; ++r)
    SFNO_CUDA(cudaMemcpy((char*)*dst + (size_t)r * nbytes, host.data(), nbytes, cudaMemcpyHostToDevice));
  *bytes += nbytes * replicas;
  return SFNO_OK;
}

template <class T>
static int upload_all(const ShtTables& h, ShtDeviceTables& d) {
  // analysis table keeps [m][l][k]
  SFNO_TRY(upload_padded<T>(h.weights, (int64_t)d.mmax * d.lmax, d.nlat, d.Kp, &d.wq, &d.bytes));
  // synthesis table transposed to [m][k][l]
  std::vector<double> pt((size_t)d.mmax * d.nlat * d.lmax);
  for (int m = 0; m < d.mmax; ++m)
    for (int l = 0; l < d.lmax; ++l)
      for (int k = 0; k < d.nlat; ++k)
        pt[((size_t)m * d.nlat + k) * d.lmax + l] = h.pct[((size_t)m * d.lmax + l) * d.nlat + k];
  SFNO_TRY(upload_padded<T>(pt, (int64_t)d.mmax * d.nlat, d.lmax, d.Lq, &d.pt, &d.bytes));
  std::vector<double> e;
  build_dft_forward(d.nlon, d.mmax, e);
  // the DFT bases are read by every CTA of a launch at about the same time: replicas at different addresses spread
  // that broadcast over the L2 slices (each CTA reads replica blockIdx % basis_reps)
  SFNO_TRY(upload_padded<T>(e, 2 * d.mmax, d.nlon, d.Wp, &d.efwd, &d.bytes, d.basis_reps));
  build_dft_inverse(d.nlon, d.mmax, e);
  SFNO_TRY(upload_padded<T>(e, d.nlon, 2 * d.mmax, d.Kq2, &d.einv, &d.bytes, d.basis_reps));
  return SFNO_OK;
}

int sht_tables_upload(int nlat, int nlon, int lmax, int mmax, int grid, int precision, ShtDeviceTables& d) {
  ShtTables h;
  if (!build_sht_tables(nlat, nlon, lmax, mmax, grid, h))
    return fail(SFNO_ERR_INVALID_ARGUMENT, "invalid SHT geometry nlat=%d nlon=%d lmax=%d mmax=%d grid=%d", nlat, nlon, lmax, mmax, grid);
  if (mmax > nlon / 2 + 1) return fail(SFNO_ERR_INVALID_ARGUMENT, "mmax=%d exceeds nlon/2+1=%d", mmax, nlon / 2 + 1);
  d = ShtDeviceTables{};
  d.nlat = nlat; d.nlon = nlon; d.lmax = lmax; d.mmax = mmax; d.grid = grid; d.precision = precision;
  d.Kp = round_up(nlat, 8); d.Lq = round_up(lmax, 8); d.Wp = round_up(nlon, 8); d.Kq2 = round_up(2 * mmax, 8);
  g_upload_tf32 = precision == SFNO_PREC_TF32;
  int st = precision == SFNO_PREC_BF16 ? upload_all<bf16>(h, d) : upload_all<float>(h, d);
  g_upload_tf32 = false;
  if (st != SFNO_OK) sht_tables_free(d);
  return st;
}

template <class T>
static int upload_adjoint(const ShtTables& h, ShtDeviceTables& d) {
  // analysis table transposed to [m][k][l] (the layout of the synthesis table pt)
  std::vector<double> wt((size_t)d.mmax * d.nlat * d.lmax), e, et;
  for (int m = 0; m < d.mmax; ++m)
    for (int l = 0; l < d.lmax; ++l)
      for (int k = 0; k < d.nlat; ++k)
        wt[((size_t)m * d.nlat + k) * d.lmax + l] = h.weights[((size_t)m * d.lmax + l) * d.nlat + k];
  SFNO_TRY(upload_padded<T>(wt, (int64_t)d.mmax * d.nlat, d.lmax, d.Lq, &d.wq_t, &d.bytes));
  // synthesis table in the analysis layout [m][l][k]
  SFNO_TRY(upload_padded<T>(h.pct, (int64_t)d.mmax * d.lmax, d.nlat, d.Kp, &d.pct_a, &d.bytes));
  // forward basis [2mmax][nlon] -> [nlon][2mmax]
  build_dft_forward(d.nlon, d.mmax, e);
  et.assign((size_t)d.nlon * 2 * d.mmax, 0.0);
  for (int r = 0; r < 2 * d.mmax; ++r)
    for (int j = 0; j < d.nlon; ++j) et[(size_t)j * 2 * d.mmax + r] = e[(size_t)r * d.nlon + j];
  SFNO_TRY(upload_padded<T>(et, d.nlon, 2 * d.mmax, d.Kq2, &d.efwd_t, &d.bytes, d.basis_reps));
  // inverse basis [nlon][2mmax] -> [2mmax][nlon]
  build_dft_inverse(d.nlon, d.mmax, e);
  et.assign((size_t)2 * d.mmax * d.nlon, 0.0);
  for (int j = 0; j < d.nlon; ++j)
    for (int r = 0; r < 2 * d.mmax; ++r) et[(size_t)r * d.nlon + j] = e[(size_t)j * 2 * d.mmax + r];
  SFNO_TRY(upload_padded<T>(et, 2 * d.mmax, d.nlon, d.Wp, &d.einv_t, &d.bytes, d.basis_reps));
  return SFNO_OK;
}

int sht_tables_enable_adjoint(ShtDeviceTables& d) {
  if (d.wq_t) return SFNO_OK;
  ShtTables h;
  if (!build_sht_tables(d.nlat, d.nlon, d.lmax, d.mmax, d.grid, h)) return fail(SFNO_ERR_INVALID_ARGUMENT, "invalid SHT geometry");
  g_upload_tf32 = d.precision == SFNO_PREC_TF32;
  const int st = d.precision == SFNO_PREC_BF16 ? upload_adjoint<bf16>(h, d) : upload_adjoint<float>(h, d);
  g_upload_tf32 = false;
  return st;
}

void sht_tables_free(ShtDeviceTables& t) {
  cudaFree(t.wq); cudaFree(t.pt); cudaFree(t.efwd); cudaFree(t.einv);
  cudaFree(t.wq_t); cudaFree(t.efwd_t); cudaFree(t.pct_a); cudaFree(t.einv_t);
  t.wq = t.pt = t.efwd = t.einv = t.wq_t = t.efwd_t = t.pct_a = t.einv_t = nullptr;
}

// ---- stand-alone SHT through the same ops the network uses (B = 1, C = fields) -----------------------
struct ShtWs {
  size_t x_off, f_off, s_off, total;
};
static ShtWs sht_ws_layout(const ShtDeviceTables& t, int64_t fields) {
  const size_t e = t.precision == SFNO_PREC_BF16 ? 2 : 4;
  ShtWs w{};
  size_t off = 0;
  w.x_off = off; off += align_up((size_t)fields * t.nlat * t.nlon * e, 256);
  w.f_off = off; off += align_up((size_t)t.mmax * 2 * fields * t.Kp * e, 256);
  w.s_off = off; off += align_up((size_t)t.lmax * t.mmax * 2 * fields * e, 256);
  w.total = off;
  return w;
}

// analysis-shaped pair (longitude GEMM with `basis` as the A operand, then per-wavenumber GEMM with `table` as the A
// operand): the forward transform with (efwd, wq), the ADJOINT of the inverse transform with (einv^T, pct)
template <class T>
static int sht_forward_impl(const ShtDeviceTables& t, const void* basis, const void* table, const float* x, float* coeffs, int64_t fields,
                            char* ws, cudaStream_t st) {
  const ShtWs L = sht_ws_layout(t, fields);
  T* xt = (T*)(ws + L.x_off);
  T* F = (T*)(ws + L.f_off);
  T* X = (T*)(ws + L.s_off);
  const int C = (int)fields;
  const int64_t n = fields * t.nlat * t.nlon;
  const T* xin;
  if constexpr (std::is_same<T, float>::value) xin = x;
  else {
    convert_planes_kernel<float, T><<<dim3((unsigned)std::min<int64_t>(ceil_div64(n, 256), 4096), 1), 256, 0, st>>>(x, 0, xt, 0, n);
    SFNO_TRY(post_launch("convert_planes"));
    xin = xt;
  }
  OpDft<T> dft{};
  dft.G = C; dft.M = 2 * t.mmax; dft.N = t.nlat; dft.K = t.nlon;
  dft.A = (const T*)basis; dft.Bm = xin; dft.a_sk = 1; dft.b_sk = 1;
  dft.f = F; dft.aff_a = nullptr; dft.aff_d = nullptr;
  dft.B = 1; dft.C = C; dft.nlat = t.nlat; dft.nlon = t.nlon; dft.Kp = t.Kp; dft.Wp = t.Wp; dft.x_bstride = 0;
  dft.a_reps = t.basis_reps;
  dft.round_out = 1;   // F feeds the Legendre MMA
  SFNO_TRY(launch_gemm(dft, st, "dft_fwd"));
  OpLeg<T> leg{};
  leg.G = t.mmax; leg.M = t.lmax; leg.N = 2 * C; leg.K = t.nlat;
  leg.A = (const T*)table; leg.Bm = F; leg.a_sk = 1; leg.b_sk = 1;
  leg.x = X; leg.Kp = t.Kp; leg.lmax = t.lmax; leg.mmax = t.mmax; leg.triangular = 0;
  SFNO_TRY(launch_gemm(leg, st, "legendre_fwd"));
  const int64_t total = fields * t.lmax * t.mmax * 2;
  internal_to_coeffs_kernel<T><<<(unsigned)std::min<int64_t>(ceil_div64(total, 256), 8192), 256, 0, st>>>(X, coeffs, C, t.lmax, t.mmax);
  return post_launch("internal_to_coeffs");
}

// synthesis-shaped pair: the inverse transform with (pt, einv), the ADJOINT of the forward transform with (wq^T, efwd^T)
template <class T>
static int sht_inverse_impl(const ShtDeviceTables& t, const void* table, const void* basis, const float* coeffs, float* x, int64_t fields,
                            char* ws, cudaStream_t st) {
  const ShtWs L = sht_ws_layout(t, fields);
  T* xt = (T*)(ws + L.x_off);
  T* Gb = (T*)(ws + L.f_off);
  T* X = (T*)(ws + L.s_off);
  const int C = (int)fields;
  const int64_t total = fields * t.lmax * t.mmax * 2;
  coeffs_to_internal_kernel<T><<<(unsigned)std::min<int64_t>(ceil_div64(total, 256), 8192), 256, 0, st>>>(coeffs, X, C, t.lmax, t.mmax);
  SFNO_TRY(post_launch("coeffs_to_internal"));
  OpIleg<T> il{};
  il.G = t.mmax; il.M = 2 * C; il.N = t.nlat; il.K = t.lmax;
  il.A = X; il.Bm = (const T*)table; il.b_sk = 1;
  il.a_goff = il.M; il.a_sk = (int64_t)t.mmax * il.M;  // X layout [l][m][rows]
  il.g_out = Gb; il.B = 1; il.C = C; il.Kp = t.Kp; il.Lq = t.Lq; il.nlat = t.nlat; il.triangular = 0;
  il.round_out = 1;    // G feeds the inverse-DFT MMA
  SFNO_TRY(launch_gemm(il, st, "legendre_inv"));
  IdftArgs<T, float> id{};
  id.G = 1; id.M = C * t.Kp; id.N = t.nlon; id.K = 2 * t.mmax;
  id.A = Gb; id.Bm = (const T*)basis; id.a_sk = id.M; id.b_sk = 1;
  id.out = x; id.out_bstride = 0; id.bias = nullptr; id.add = nullptr; id.add_bstride = 0; id.act = SFNO_ACT_NONE;
  id.C = C; id.nlat = t.nlat; id.nlon = t.nlon; id.Kp = t.Kp; id.Kq2 = t.Kq2; id.b_reps = t.basis_reps; id.stat_part = nullptr;
  (void)xt;
  return launch_idft(id, st, "dft_inv");
}

// ---- spectral contraction in the reference layout (fp32 CUDA cores) --------------------------------------
// out[b,o,l,m] = sum_i x[b,i,l,m] * w[i,o,l(,m)]  (complex), one thread per (b,o,l,m)
__global__ void spectral_contract_kernel(int diagonal, const float2* __restrict__ x, const float2* __restrict__ w,
                                         float2* __restrict__ out, int B, int Cin, int Cout, int L, int M) {
  const int64_t total = (int64_t)B * Cout * L * M;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int m = (int)(idx % M);
    int64_t r = idx / M;
    const int l = (int)(r % L); r /= L;
    const int o = (int)(r % Cout);
    const int b = (int)(r / Cout);
    float re = 0.0f, im = 0.0f;
    for (int i = 0; i < Cin; ++i) {
      const float2 xv = x[(((int64_t)b * Cin + i) * L + l) * M + m];
      const float2 wv = diagonal ? w[(((int64_t)i * Cout + o) * L + l) * M + m] : w[((int64_t)i * Cout + o) * L + l];
      re = fmaf(xv.x, wv.x, re); re = fmaf(-xv.y, wv.y, re);
      im = fmaf(xv.x, wv.y, im); im = fmaf(xv.y, wv.x, im);
    }
    out[idx] = make_float2(re, im);
  }
}

}  // namespace sfno

using namespace sfno;

extern "C" {

int sfno_b200_abi_version(void) { return 2; }   // 2: round-2 entry points (rng state, tf32, spectral_conv, ensemble moments, step glue)

const char* sfno_b200_status_string(int status) {
  switch (status) {
    case SFNO_OK: return "ok";
    case SFNO_ERR_INVALID_ARGUMENT: return "invalid argument";
    case SFNO_ERR_CUDA: return "CUDA error";
    case SFNO_ERR_WORKSPACE_TOO_SMALL: return "workspace too small";
    case SFNO_ERR_UNSUPPORTED: return "unsupported configuration";
    case SFNO_ERR_NO_DEVICE: return "no CUDA device";
    case SFNO_ERR_UNKNOWN_PARAM: return "unknown parameter name";
    case SFNO_ERR_SHAPE_MISMATCH: return "shape mismatch";
    default: return "unknown status";
  }
}

const char* sfno_b200_last_error(void) { return last_error_ref().c_str(); }
int64_t sfno_b200_launch_count(void) { return g_launch_count.load(); }

int sfno_b200_profile_begin(void* stream) {
  if (g_profile_on.load()) return fail(SFNO_ERR_INVALID_ARGUMENT, "profile already running");
  for (cudaEvent_t e : g_prof.events) cudaEventDestroy(e);
  g_prof.events.clear();
  g_prof.names.clear();
  g_prof.stream = (cudaStream_t)stream;
  g_profile_on.store(1);
  profile_mark("begin");
  return SFNO_OK;
}

int sfno_b200_profile_end(char* names, size_t names_capacity, float* ms, int capacity) {
  if (!g_profile_on.load()) return fail(SFNO_ERR_INVALID_ARGUMENT, "profile not running");
  g_profile_on.store(0);
  SFNO_CUDA(cudaStreamSynchronize(g_prof.stream));
  const int n = (int)g_prof.events.size() - 1;
  std::string joined;
  for (int i = 0; i < n && i < capacity; ++i) {
    float t = 0.0f;
    cudaEventElapsedTime(&t, g_prof.events[i], g_prof.events[i + 1]);
    if (ms) ms[i] = t;
    joined += g_prof.names[i + 1];
    joined += '\n';
  }
  if (names && names_capacity > 0) {
    size_t len = std::min(names_capacity - 1, joined.size());
    memcpy(names, joined.data(), len);
    names[len] = 0;
  }
  for (cudaEvent_t e : g_prof.events) cudaEventDestroy(e);
  g_prof.events.clear();
  g_prof.names.clear();
  return std::min(n, capacity);
}

int sfno_sht_tables_host(int nlat, int nlon, int lmax, int mmax, int grid, double* nodes, double* quad_w,
                         double* weights, double* pct) {
  ShtTables t;
  if (!build_sht_tables(nlat, nlon, lmax, mmax, grid, t)) return fail(SFNO_ERR_INVALID_ARGUMENT, "invalid SHT geometry");
  if (nodes) memcpy(nodes, t.cost.data(), sizeof(double) * t.cost.size());
  if (quad_w) memcpy(quad_w, t.quad_w.data(), sizeof(double) * t.quad_w.size());
  if (weights) memcpy(weights, t.weights.data(), sizeof(double) * t.weights.size());
  if (pct) memcpy(pct, t.pct.data(), sizeof(double) * t.pct.size());
  return SFNO_OK;
}

int sfno_sht_plan_create(sfno_sht_plan** plan, int nlat, int nlon, int lmax, int mmax, int grid, int precision) {
  SFNO_CHECK_ARG(plan != nullptr, "plan is NULL");
  SFNO_CHECK_ARG(precision == SFNO_PREC_F32 || precision == SFNO_PREC_BF16 || precision == SFNO_PREC_TF32, "bad precision %d", precision);
  auto* p = new sfno_sht_plan();
  int st = sht_tables_upload(nlat, nlon, lmax, mmax, grid, precision, p->t);
  if (st != SFNO_OK) { delete p; return st; }
  *plan = p;
  return SFNO_OK;
}

int sfno_sht_plan_destroy(sfno_sht_plan* plan) {
  if (!plan) return SFNO_OK;
  sht_tables_free(plan->t);
  delete plan;
  return SFNO_OK;
}

size_t sfno_sht_workspace_bytes(const sfno_sht_plan* plan, int64_t fields) {
  if (!plan || fields <= 0) return 0;
  return sht_ws_layout(plan->t, fields).total;
}

int sfno_sht_forward(const sfno_sht_plan* plan, const float* x_dev, float* coeffs_dev, int64_t fields,
                     void* workspace_dev, size_t workspace_bytes, void* stream) {
  SFNO_CHECK_ARG(plan && x_dev && coeffs_dev && workspace_dev, "NULL argument");
  SFNO_CHECK_ARG(fields > 0 && fields < (1 << 24), "bad field count %lld", (long long)fields);
  if (workspace_bytes < sht_ws_layout(plan->t, fields).total) return fail(SFNO_ERR_WORKSPACE_TOO_SMALL, "workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  Tf32Scope tf32(plan->t.precision == SFNO_PREC_TF32);
  const ShtDeviceTables& t = plan->t;
  return t.precision == SFNO_PREC_BF16 ? sht_forward_impl<bf16>(t, t.efwd, t.wq, x_dev, coeffs_dev, fields, (char*)workspace_dev, st)
                                       : sht_forward_impl<float>(t, t.efwd, t.wq, x_dev, coeffs_dev, fields, (char*)workspace_dev, st);
}

int sfno_sht_inverse(const sfno_sht_plan* plan, const float* coeffs_dev, float* x_dev, int64_t fields,
                     void* workspace_dev, size_t workspace_bytes, void* stream) {
  SFNO_CHECK_ARG(plan && x_dev && coeffs_dev && workspace_dev, "NULL argument");
  SFNO_CHECK_ARG(fields > 0 && fields < (1 << 24), "bad field count %lld", (long long)fields);
  if (workspace_bytes < sht_ws_layout(plan->t, fields).total) return fail(SFNO_ERR_WORKSPACE_TOO_SMALL, "workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  Tf32Scope tf32(plan->t.precision == SFNO_PREC_TF32);
  const ShtDeviceTables& t = plan->t;
  return t.precision == SFNO_PREC_BF16 ? sht_inverse_impl<bf16>(t, t.pt, t.einv, coeffs_dev, x_dev, fields, (char*)workspace_dev, st)
                                       : sht_inverse_impl<float>(t, t.pt, t.einv, coeffs_dev, x_dev, fields, (char*)workspace_dev, st);
}

// ---- adjoint transforms (backward pass of the two custom ops) -------------------------------------------
// Both transforms are real-linear maps between [fields][nlat][nlon] and the (re, im) pairs of [fields][lmax][mmax]; their
// adjoints are the transposed maps, which have the SHAPE of the opposite transform with transposed tables.
int sfno_sht_forward_adjoint(sfno_sht_plan* plan, const float* grad_coeffs_dev, float* grad_x_dev, int64_t fields, void* workspace_dev,
                             size_t workspace_bytes, void* stream) {
  SFNO_CHECK_ARG(plan && grad_coeffs_dev && grad_x_dev && workspace_dev, "NULL argument");
  SFNO_CHECK_ARG(fields > 0 && fields < (1 << 24), "bad field count %lld", (long long)fields);
  if (workspace_bytes < sht_ws_layout(plan->t, fields).total) return fail(SFNO_ERR_WORKSPACE_TOO_SMALL, "workspace too small");
  SFNO_TRY(sht_tables_enable_adjoint(plan->t));
  cudaStream_t st = (cudaStream_t)stream;
  Tf32Scope tf32(plan->t.precision == SFNO_PREC_TF32);
  const ShtDeviceTables& t = plan->t;
  return t.precision == SFNO_PREC_BF16 ? sht_inverse_impl<bf16>(t, t.wq_t, t.efwd_t, grad_coeffs_dev, grad_x_dev, fields, (char*)workspace_dev, st)
                                       : sht_inverse_impl<float>(t, t.wq_t, t.efwd_t, grad_coeffs_dev, grad_x_dev, fields, (char*)workspace_dev, st);
}

int sfno_sht_inverse_adjoint(sfno_sht_plan* plan, const float* grad_x_dev, float* grad_coeffs_dev, int64_t fields, void* workspace_dev,
                             size_t workspace_bytes, void* stream) {
  SFNO_CHECK_ARG(plan && grad_coeffs_dev && grad_x_dev && workspace_dev, "NULL argument");
  SFNO_CHECK_ARG(fields > 0 && fields < (1 << 24), "bad field count %lld", (long long)fields);
  if (workspace_bytes < sht_ws_layout(plan->t, fields).total) return fail(SFNO_ERR_WORKSPACE_TOO_SMALL, "workspace too small");
  SFNO_TRY(sht_tables_enable_adjoint(plan->t));
  cudaStream_t st = (cudaStream_t)stream;
  Tf32Scope tf32(plan->t.precision == SFNO_PREC_TF32);
  const ShtDeviceTables& t = plan->t;
  return t.precision == SFNO_PREC_BF16 ? sht_forward_impl<bf16>(t, t.einv_t, t.pct_a, grad_x_dev, grad_coeffs_dev, fields, (char*)workspace_dev, st)
                                       : sht_forward_impl<float>(t, t.einv_t, t.pct_a, grad_x_dev, grad_coeffs_dev, fields, (char*)workspace_dev, st);
}

int sfno_spectral_contract(int operator_type, const float* x_dev, const float* weight_dev, float* out_dev, int batch,
                           int cin, int cout, int lmax, int mmax, void* stream) {
  SFNO_CHECK_ARG(x_dev && weight_dev && out_dev, "NULL argument");
  SFNO_CHECK_ARG(operator_type == SFNO_OP_DHCONV || operator_type == SFNO_OP_DIAGONAL, "bad operator_type %d", operator_type);
  SFNO_CHECK_ARG(batch > 0 && cin > 0 && cout > 0 && lmax > 0 && mmax > 0, "bad sizes");
  const int64_t total = (int64_t)batch * cout * lmax * mmax;
  spectral_contract_kernel<<<(unsigned)std::min<int64_t>(ceil_div64(total, 256), 1 << 20), 256, 0, (cudaStream_t)stream>>>(
      operator_type == SFNO_OP_DIAGONAL, (const float2*)x_dev, (const float2*)weight_dev, (float2*)out_dev, batch, cin, cout, lmax, mmax);
  return post_launch("spectral_contract");
}

size_t sfno_instance_norm_workspace_bytes(int batch, int channels) { return (size_t)4 * batch * channels * sizeof(float); }

int sfno_instance_norm(const float* x_dev, float* y_dev, const float* gamma_dev, const float* beta_dev,
                       const float* scale_dev, const float* shift_dev, int batch, int channels, int64_t hw, float eps,
                       void* workspace_dev, size_t workspace_bytes, void* stream) {
  SFNO_CHECK_ARG(x_dev && y_dev && workspace_dev, "NULL argument");
  SFNO_CHECK_ARG(batch > 0 && channels > 0 && hw > 0, "bad sizes");
  SFNO_CHECK_ARG((scale_dev == nullptr) == (shift_dev == nullptr), "scale and shift must be given together");
  if (workspace_bytes < sfno_instance_norm_workspace_bytes(batch, channels)) return fail(SFNO_ERR_WORKSPACE_TOO_SMALL, "workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int BC = batch * channels;
  float* mean = (float*)workspace_dev;
  float* rstd = mean + BC;
  float* a = rstd + BC;
  float* d = a + BC;
  launch_instance_stats<float>(x_dev, (int64_t)channels * hw, batch, channels, hw, eps, mean, rstd, st);
  SFNO_TRY(post_launch("instance_stats"));
  // scale/shift are given as separate [batch][C] arrays here: stage them as ts = [scale | shift] is not possible
  // without a copy, so the affine kernel is called with ts == nullptr and scale/shift are applied below.
  norm_affine_kernel<<<ceil_div(BC, 256), 256, 0, st>>>(mean, rstd, gamma_dev, beta_dev, nullptr, 0, batch, channels, a, d);
  SFNO_TRY(post_launch("norm_affine"));
  if (scale_dev) {
    time_affine_compose_kernel<<<ceil_div(BC, 256), 256, 0, st>>>(a, d, scale_dev, shift_dev, BC);
    SFNO_TRY(post_launch("time_affine_compose"));
  }
  affine_apply_kernel<<<dim3((unsigned)std::min<int64_t>(ceil_div64(hw, 256 * 16), 16), BC), 256, 0, st>>>(x_dev, y_dev, a, d, hw);
  return post_launch("affine_apply");
}

int sfno_conv1x1(const float* x_dev, const float* weight_dev, const float* bias_dev, const float* residual_dev,
                 float* y_dev, int batch, int cin, int cout, int64_t hw, int activation, void* stream) {
  SFNO_CHECK_ARG(x_dev && weight_dev && y_dev, "NULL argument");
  SFNO_CHECK_ARG(batch > 0 && cin > 0 && cout > 0 && hw > 0 && hw < (1ll << 31), "bad sizes");
  ConvArgs<float, float> op{};
  op.G = batch; op.M = cout; op.N = (int)hw; op.K = cin;
  op.A = weight_dev; op.Bm = x_dev; op.a_sk = 1; op.b_sk = hw;
  op.in_bstride = (int64_t)cin * hw; op.w_bstride = 0; op.ldw = cin;
  op.bias = bias_dev; op.bias_bstride = 0; op.act = activation;
  op.drop_p = 0.0f; op.seed = 0; op.offset = 0; op.branch_scale = nullptr;
  op.res = residual_dev; op.res_bstride = (int64_t)cout * hw; op.res_a = nullptr; op.res_d = nullptr; op.pos = nullptr;
  op.out = y_dev; op.out_bstride = (int64_t)cout * hw; op.stat_part = nullptr;
  return launch_conv(op, (cudaStream_t)stream, "conv1x1");
}

}  // extern "C"
