// Whole-network executor: SphericalFourierNeuralOperatorNet.forward (sfnonet.py:797-841) as a fixed
// sequence of GEMM ops + small kernels on one stream, with the reference's state_dict as input.
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"
#include "engine.cuh"
#include "ops.cuh"
#include "pointwise.cuh"
#include "sht_plan.cuh"

namespace sfno {

std::atomic<int> g_force_simt{0};

struct BlockParams {
  float *norm0_g = nullptr, *norm0_b = nullptr, *norm1_g = nullptr, *norm1_b = nullptr;  // [C]
  void* wpack = nullptr;      // T [L][2C][2C]           (dhconv, packed real form)
  float* wdiag = nullptr;     // fp32 [C][C][L][M][2]    (diagonal operator, reference layout)
  float* spec_bias = nullptr; // [C]
  float* skip_w32 = nullptr;  // [C][C] fp32 master (folded with the norm0/time affine per sample)
  void* skip_wT = nullptr;    // T [C][C]  (used as-is when the residual is materialised)
  float* skip_b = nullptr;    // [C]
  float* fc1_w32 = nullptr;   // [hid][C] fp32 master (folded with the norm1 affine per sample)
  float* fc1_b = nullptr;     // [hid]
  void* fc2_wT = nullptr;     // T [C][hid]
  float* fc2_b = nullptr;     // [C]
};

struct WsLayout {
  size_t xin, buf0, xcat, t1, res, hid, fg, X, Y;
  size_t mean, rstd, a0, d0, a1, d1, tsin, th, trepr, ts, skip_wb, skip_bb, fc1_wb, fc1_bb, dscale, stat_part, mask1, mask2;
  size_t total;
};

}  // namespace sfno

using namespace sfno;

struct sfno_net {
  sfno_net_config cfg{};
  int P = 0, C = 0, Cin = 0, Cin_p = 0, Cout = 0, Ccat = 0, Ccat_p = 0, hid = 0, tdim = 0, L = 0, M = 0, nl = 0;
  size_t esize = 4;
  ShtDeviceTables data_grid, lg_grid;
  void* pos = nullptr;
  void* enc0_w = nullptr; float* enc0_b = nullptr; void* enc1_w = nullptr;
  float *te1_w = nullptr, *te1_b = nullptr, *te3_w = nullptr, *te3_b = nullptr;
  float *tmlp_w = nullptr, *tmlp_b = nullptr;  // stacked over blocks: [nl*2C][tdim], [nl*2C]
  std::vector<BlockParams> blocks;
  void* dec0_w = nullptr; float* dec0_b = nullptr; void* dec1_w = nullptr;
  std::string names;
  std::vector<void*> allocations;
  // Option "mask_overlap" (default off): generate the dropout masks of block i on a forked stream while the spectral kernels
  // of block i run, the main stream joining before fc1.  Measured SLOWER than the in-line mask launches (interpolator forward
  // 15.17 vs 14.1-14.7 ms, window 230 vs 225 ms, profiles/r02_q_mask_overlap_ab.json): the mask blocks that fit next to a
  // persistent GEMM CTA run at a fraction of their full-occupancy speed and take issue slots from its epilogue warps.
  // Stream / events are created with the net (not inside a capture).
  int mask_overlap = 0;
  cudaStream_t side = nullptr;
  std::vector<cudaEvent_t> ev_start, ev_mask;
  int stop_after_block = -2;  // -2: run everything; -1: stop after encoder(+pos); i: stop after block i
  // bookkeeping of the last forward for debug taps
  size_t last_x_off = 0; int64_t last_x_bstride = 0; int last_batch = 0;
};

namespace sfno {

static int dev_alloc(sfno_net* n, void** p, size_t bytes, bool zero = true) {
  SFNO_CUDA(cudaMalloc(p, bytes ? bytes : 16));
  n->allocations.push_back(*p);
  if (zero) SFNO_CUDA(cudaMemset(*p, 0, bytes ? bytes : 16));
  return SFNO_OK;
}
template <class P>
static int dev_alloc_t(sfno_net* n, P** p, size_t count, size_t esize) {
  void* v = nullptr;
  SFNO_TRY(dev_alloc(n, &v, count * esize));
  *p = (P*)v;
  return SFNO_OK;
}

static WsLayout ws_layout(const sfno_net* n, int B) {
  WsLayout w{};
  size_t off = 0;
  const size_t e = n->esize;
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 1024); return o; };
  const size_t plane = (size_t)n->P;
  const int Kp = n->lg_grid.Kp;
  w.xin = take((size_t)B * n->Cin * plane * e);
  w.buf0 = take((size_t)B * n->C * plane * e);
  w.xcat = take((size_t)B * n->Ccat * plane * e);
  w.t1 = take((size_t)B * n->C * plane * e);
  w.res = take((size_t)B * n->C * plane * e);
  w.hid = take((size_t)B * std::max(n->hid, 1) * plane * e);
  w.fg = take((size_t)n->M * B * 2 * n->C * Kp * e);
  w.X = take((size_t)n->L * n->M * B * 2 * n->C * e);
  w.Y = take((size_t)n->L * n->M * B * 2 * n->C * e);
  const size_t bc = (size_t)B * n->C * sizeof(float);
  w.mean = take(bc); w.rstd = take(bc); w.a0 = take(bc); w.d0 = take(bc); w.a1 = take(bc); w.d1 = take(bc);
  w.tsin = take((size_t)B * n->C * sizeof(float));
  w.th = take((size_t)B * std::max(n->tdim, 1) * sizeof(float));
  w.trepr = take((size_t)B * std::max(n->tdim, 1) * sizeof(float));
  w.ts = take((size_t)B * n->nl * 2 * n->C * sizeof(float));
  w.skip_wb = take((size_t)B * n->C * n->C * e);
  w.skip_bb = take((size_t)B * n->C * sizeof(float));
  w.fc1_wb = take((size_t)B * std::max(n->hid, 1) * n->C * e);
  w.fc1_bb = take((size_t)B * std::max(n->hid, 1) * sizeof(float));
  w.dscale = take((size_t)B * sizeof(float));
  // keep masks of the two MLP dropout sites (one bit per element), only for nets with dropout
  const bool has_drop = n->cfg.dropout_mlp > 0.0f;
  w.mask1 = take(has_drop ? (size_t)B * std::max(n->hid, 1) * plane / 8 + 16 : 0);
  w.mask2 = take(has_drop ? (size_t)B * n->C * plane / 8 + 16 : 0);
  {  // per-slice partial sums of the fused InstanceNorm statistics (tensor-core epilogues)
    const size_t conv_slices = (size_t)conv_stat_slices(n->P);
    const size_t idft_slices = (size_t)idft_stat_slices(n->cfg.nlon);
    const size_t conv_f = conv_slices * 2 * (size_t)B * n->C, idft_f = idft_slices * 2 * (size_t)B * n->C * Kp;
    w.stat_part = take(std::max(conv_f, idft_f) * sizeof(float));
  }
  w.total = off;
  return w;
}

static thread_local bool g_pack_tf32 = false;

// operands of a tf32 MMA are stored TF32-exact (round to nearest here; the tensor core would truncate)
static int round_tf32_inplace(float* p, int64_t n, cudaStream_t st) {
  round_tf32_kernel<<<(unsigned)std::min<int64_t>(ceil_div64(n, 256), 148 * 16), 256, 0, st>>>(p, n);
  return post_launch("round_tf32");
}

template <class T>
static int pack_rows(const float* src, int rows, int cols, int ld, void* dst, cudaStream_t st) {
  const int64_t total = (int64_t)rows * ld;
  pack_rows_kernel<T><<<(unsigned)ceil_div64(total, 256), 256, 0, st>>>(src, rows, cols, ld, (T*)dst);
  SFNO_TRY(post_launch("pack_rows"));
  if constexpr (std::is_same<T, float>::value) {
    if (g_pack_tf32) return round_tf32_inplace((float*)dst, total, st);
  }
  return SFNO_OK;
}

static int copy_f32(float* dst, const float* src, int64_t n, cudaStream_t st) {
  SFNO_CUDA(cudaMemcpyAsync(dst, src, (size_t)n * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return SFNO_OK;
}

template <class T>
static int set_param_impl(sfno_net* n, const std::string& name, const float* v, int64_t numel, cudaStream_t st) {
  const int C = n->C, hid = n->hid, tdim = n->tdim, L = n->L, M = n->M;
  auto expect = [&](int64_t want) -> int {
    if (numel != want) return fail(SFNO_ERR_SHAPE_MISMATCH, "%s: expected %lld elements, got %lld", name.c_str(), (long long)want, (long long)numel);
    return SFNO_OK;
  };
  if (name == "pos_embed") { SFNO_TRY(expect((int64_t)C * n->P)); return pack_rows<T>(v, C, n->P, n->P, n->pos, st); }
  if (name == "encoder.0.weight") { SFNO_TRY(expect((int64_t)C * n->Cin)); return pack_rows<T>(v, C, n->Cin, n->Cin_p, n->enc0_w, st); }
  if (name == "encoder.0.bias") { SFNO_TRY(expect(C)); return copy_f32(n->enc0_b, v, C, st); }
  if (name == "encoder.2.weight") { SFNO_TRY(expect((int64_t)C * C)); return pack_rows<T>(v, C, C, C, n->enc1_w, st); }
  if (name == "time_emb_mlp.1.weight") { SFNO_TRY(expect((int64_t)tdim * C)); return copy_f32(n->te1_w, v, numel, st); }
  if (name == "time_emb_mlp.1.bias") { SFNO_TRY(expect(tdim)); return copy_f32(n->te1_b, v, numel, st); }
  if (name == "time_emb_mlp.3.weight") { SFNO_TRY(expect((int64_t)tdim * tdim)); return copy_f32(n->te3_w, v, numel, st); }
  if (name == "time_emb_mlp.3.bias") { SFNO_TRY(expect(tdim)); return copy_f32(n->te3_b, v, numel, st); }
  if (name == "decoder.0.weight") { SFNO_TRY(expect((int64_t)C * n->Ccat)); return pack_rows<T>(v, C, n->Ccat, n->Ccat_p, n->dec0_w, st); }
  if (name == "decoder.0.bias") { SFNO_TRY(expect(C)); return copy_f32(n->dec0_b, v, C, st); }
  if (name == "decoder.2.weight") { SFNO_TRY(expect((int64_t)n->Cout * C)); return pack_rows<T>(v, n->Cout, C, C, n->dec1_w, st); }
  if (name.rfind("blocks.", 0) == 0) {
    size_t dot = name.find('.', 7);
    if (dot == std::string::npos) return fail(SFNO_ERR_UNKNOWN_PARAM, "unknown parameter %s", name.c_str());
    int i = atoi(name.substr(7, dot - 7).c_str());
    if (i < 0 || i >= n->nl) return fail(SFNO_ERR_UNKNOWN_PARAM, "block index out of range in %s", name.c_str());
    const std::string rest = name.substr(dot + 1);
    BlockParams& b = n->blocks[i];
    if (rest == "norm0.weight") { SFNO_TRY(expect(C)); return copy_f32(b.norm0_g, v, C, st); }
    if (rest == "norm0.bias") { SFNO_TRY(expect(C)); return copy_f32(b.norm0_b, v, C, st); }
    if (rest == "norm1.weight") { SFNO_TRY(expect(C)); return copy_f32(b.norm1_g, v, C, st); }
    if (rest == "norm1.bias") { SFNO_TRY(expect(C)); return copy_f32(b.norm1_b, v, C, st); }
    if (rest == "time_mlp.1.weight") { SFNO_TRY(expect((int64_t)2 * C * tdim)); return copy_f32(n->tmlp_w + (int64_t)i * 2 * C * tdim, v, numel, st); }
    if (rest == "time_mlp.1.bias") { SFNO_TRY(expect(2 * C)); return copy_f32(n->tmlp_b + (int64_t)i * 2 * C, v, numel, st); }
    if (rest == "filter.filter.weight") {
      if (n->cfg.operator_type == SFNO_OP_DHCONV) {
        SFNO_TRY(expect((int64_t)C * C * L * 2));
        pack_dhconv_weight_kernel<T><<<4096, 256, 0, st>>>(v, C, C, L, (T*)b.wpack);
        SFNO_TRY(post_launch("pack_dhconv_weight"));
        if constexpr (std::is_same<T, float>::value) {
          if (g_pack_tf32) return round_tf32_inplace((float*)b.wpack, (int64_t)L * 4 * C * C, st);
        }
        return SFNO_OK;
      }
      SFNO_TRY(expect((int64_t)C * C * L * M * 2));
      return copy_f32(b.wdiag, v, numel, st);
    }
    if (rest == "filter.filter.bias") { SFNO_TRY(expect(C)); return copy_f32(b.spec_bias, v, C, st); }
    if (rest == "inner_skip.weight") {
      SFNO_TRY(expect((int64_t)C * C));
      SFNO_TRY(copy_f32(b.skip_w32, v, numel, st));
      return pack_rows<T>(v, C, C, C, b.skip_wT, st);
    }
    if (rest == "inner_skip.bias") { SFNO_TRY(expect(C)); return copy_f32(b.skip_b, v, C, st); }
    if (rest == "mlp.fwd.0.weight") { SFNO_TRY(expect((int64_t)hid * C)); return copy_f32(b.fc1_w32, v, numel, st); }
    if (rest == "mlp.fwd.0.bias") { SFNO_TRY(expect(hid)); return copy_f32(b.fc1_b, v, numel, st); }
    if (rest == "mlp.fwd.2.weight" || rest == "mlp.fwd.3.weight") { SFNO_TRY(expect((int64_t)C * hid)); return pack_rows<T>(v, C, hid, hid, b.fc2_wT, st); }
    if (rest == "mlp.fwd.2.bias" || rest == "mlp.fwd.3.bias") { SFNO_TRY(expect(C)); return copy_f32(b.fc2_b, v, C, st); }
  }
  return fail(SFNO_ERR_UNKNOWN_PARAM, "unknown parameter %s", name.c_str());
}

// ---- op builders ---------------------------------------------------------------------------------------------
template <class T, class TOut>
static ConvArgs<T, TOut> make_conv(int B, int P, int cin, int cout, const T* in, int64_t in_bs, const T* w, int64_t w_bs, int ldw,
                                 const float* bias, int64_t bias_bs, int act, TOut* out, int64_t out_bs) {
  ConvArgs<T, TOut> op{};
  op.G = B; op.M = cout; op.N = P; op.K = cin;
  op.A = w; op.Bm = in; op.a_sk = 1; op.b_sk = P;
  op.in_bstride = in_bs; op.w_bstride = w_bs; op.ldw = ldw;
  op.bias = bias; op.bias_bstride = bias_bs; op.act = act;
  op.drop_p = 0.0f; op.seed = 0; op.offset = 0; op.rng_dev = nullptr; op.drop_mask = nullptr; op.branch_scale = nullptr;
  op.res = nullptr; op.res_bstride = 0; op.res_a = nullptr; op.res_d = nullptr; op.pos = nullptr;
  op.out = out; op.out_bstride = out_bs; op.stat_part = nullptr;
  op.round_out = 1;   // activations stay inside the net (only the decoder output leaves it, see forward_impl)
  return op;
}

template <class T>
static int run_dft(const sfno_net* n, const ShtDeviceTables& t, int B, const T* x, int64_t x_bs, const float* a, const float* d, T* F, cudaStream_t st) {
  OpDft<T> op{};
  op.G = B * n->C; op.M = 2 * t.mmax; op.N = t.nlat; op.K = t.nlon;
  op.A = (const T*)t.efwd; op.Bm = x; op.a_sk = 1; op.b_sk = 1;
  op.f = F; op.aff_a = a; op.aff_d = d;
  op.B = B; op.C = n->C; op.nlat = t.nlat; op.nlon = t.nlon; op.Kp = t.Kp; op.Wp = t.Wp; op.x_bstride = x_bs;
  op.a_reps = t.basis_reps; op.round_out = 1;
  return launch_gemm(op, st, "dft_fwd");
}
template <class T>
static int run_leg(const sfno_net* n, const ShtDeviceTables& t, int B, const T* F, T* X, cudaStream_t st) {
  OpLeg<T> op{};
  op.G = t.mmax; op.M = t.lmax; op.N = B * 2 * n->C; op.K = t.nlat;
  op.A = (const T*)t.wq; op.Bm = F; op.a_sk = 1; op.b_sk = 1;
  op.x = X; op.Kp = t.Kp; op.lmax = t.lmax; op.mmax = t.mmax;
  op.triangular = n->cfg.operator_type == SFNO_OP_DHCONV; op.round_out = 1;
  return launch_gemm(op, st, "legendre_fwd");
}
template <class T>
static int run_ileg(const sfno_net* n, const ShtDeviceTables& t, int B, const T* S, T* G, cudaStream_t st) {
  OpIleg<T> op{};
  op.G = t.mmax; op.M = B * 2 * n->C; op.N = t.nlat; op.K = t.lmax;
  op.A = S; op.Bm = (const T*)t.pt; op.b_sk = 1;
  op.a_goff = op.M; op.a_sk = (int64_t)t.mmax * op.M;   // X and Y share the layout [l][m][(b,ri,c)]
  op.g_out = G; op.B = B; op.C = n->C; op.Kp = t.Kp; op.Lq = t.Lq; op.nlat = t.nlat;
  op.triangular = n->cfg.operator_type == SFNO_OP_DHCONV; op.round_out = 1;
  return launch_gemm(op, st, "legendre_inv");
}
// stat_part != nullptr: ask for fused output statistics; *fused reports whether the engine could provide them
template <class T>
static int run_idft(const sfno_net* n, const ShtDeviceTables& t, int B, const T* G, const float* bias, const T* add, int64_t add_bs,
                    int act, T* out, int64_t out_bs, float* stat_part, bool* fused, cudaStream_t st) {
  IdftArgs<T, T> op{};
  op.G = 1; op.M = B * n->C * t.Kp; op.N = t.nlon; op.K = 2 * t.mmax;
  op.A = G; op.Bm = (const T*)t.einv; op.a_sk = op.M; op.b_sk = 1;
  op.out = out; op.out_bstride = out_bs; op.bias = bias; op.add = add; op.add_bstride = add_bs; op.act = act;
  op.C = n->C; op.nlat = t.nlat; op.nlon = t.nlon; op.Kp = t.Kp; op.Kq2 = t.Kq2; op.b_reps = t.basis_reps;
  op.stat_part = nullptr; op.round_out = 1;
  if (fused) *fused = false;
  if (stat_part && idft_uses_tc(op)) { op.stat_part = stat_part; *fused = true; }
  return launch_idft(op, st, "dft_inv");
}

template <class T>
static int forward_impl(sfno_net* n, const ConcatParts& parts, const float* time, float* y, int B, int dropout, uint64_t seed,
                        uint64_t offset, uint64_t* rng_dev, char* ws, cudaStream_t st) {
  const sfno_net_config& cfg = n->cfg;
  const bool tf32 = std::is_same<T, float>::value && cfg.precision == SFNO_PREC_TF32;
  Tf32Scope tf32_scope(tf32);   // fp32-storage GEMMs of this forward run as kind::tf32 on the tensor cores
  const WsLayout w = ws_layout(n, B);
  const int C = n->C, P = n->P, hid = n->hid, nl = n->nl, tdim = n->tdim;
  const int64_t CP = (int64_t)C * P;
  T* xin = (T*)(ws + w.xin);
  T* buf0 = (T*)(ws + w.buf0);
  T* xcat = (T*)(ws + w.xcat);
  T* t1 = (T*)(ws + w.t1);
  T* res = (T*)(ws + w.res);
  T* hd = (T*)(ws + w.hid);
  T* FG = (T*)(ws + w.fg);
  T* X = (T*)(ws + w.X);
  T* Y = (T*)(ws + w.Y);
  float* mean = (float*)(ws + w.mean); float* rstd = (float*)(ws + w.rstd);
  float* a0 = (float*)(ws + w.a0); float* d0 = (float*)(ws + w.d0);
  float* a1 = (float*)(ws + w.a1); float* d1 = (float*)(ws + w.d1);
  float* tsin = (float*)(ws + w.tsin); float* th = (float*)(ws + w.th); float* trepr = (float*)(ws + w.trepr);
  float* ts = (float*)(ws + w.ts);
  T* skip_wb = (T*)(ws + w.skip_wb); float* skip_bb = (float*)(ws + w.skip_bb);
  T* fc1_wb = (T*)(ws + w.fc1_wb); float* fc1_bb = (float*)(ws + w.fc1_bb);
  float* dscale = (float*)(ws + w.dscale);
  const int64_t xcat_bs = (int64_t)n->Ccat * P;
  const int BC = B * C;
  // ---- input: fp32 -> T, and into the tail channels of the big-skip concat buffer (sfnonet.py:804-805,832)
  {
    // under the per-launch profile: the event gap since profile_begin is host time (wrapper, argument checks), not
    // part of the first kernel -- give it its own entry
    if (g_profile_on.load(std::memory_order_relaxed)) profile_mark("host_before_first_launch");
    dim3 grid(1184 / std::max(1, std::min(B, 8)) + 1, B);
    concat_convert_kernel<T><<<grid, 256, 0, st>>>(parts, (int64_t)P, xin, (int64_t)n->Cin * P, tf32 ? 1 : 0);
    SFNO_TRY(post_launch("convert_input"));
    if (cfg.big_skip) {
      concat_convert_kernel<T><<<grid, 256, 0, st>>>(parts, (int64_t)P, xcat + CP, xcat_bs, tf32 ? 1 : 0);
      SFNO_TRY(post_launch("convert_input_skip"));
    }
  }
  // block i reads cur and writes nxt; the last block must land in xcat (channels [0, C))
  auto buf_of = [&](int idx, T** p, int64_t* bs) {  // idx = number of blocks still to run after this tensor
    if (idx % 2 == 0) { *p = xcat; *bs = xcat_bs; } else { *p = buf0; *bs = CP; }
  };
  T* cur; int64_t cur_bs;
  buf_of(nl, &cur, &cur_bs);
  float* stat_part = (float*)(ws + w.stat_part);
  bool x_stats_fused = false;  // statistics of `cur` already sit in stat_part (written by the producing conv)
  const int conv_slices = conv_stat_slices(P);

  // ---- encoder (sfnonet.py:610-618) + position embedding (sfnonet.py:824)
  {
    auto e0 = make_conv<T, T>(B, P, n->Cin, C, xin, (int64_t)n->Cin * P, (const T*)n->enc0_w, 0, n->Cin_p, n->enc0_b, 0, cfg.activation, t1, CP);
    SFNO_TRY(launch_conv(e0, st, "encoder0"));
    auto e1 = make_conv<T, T>(B, P, C, C, t1, CP, (const T*)n->enc1_w, 0, C, nullptr, 0, SFNO_ACT_NONE, cur, cur_bs);
    // position embedding [C][P] (sfnonet.py:824) enters as a residual block with sample stride 0: the conv then runs
    // the residual + statistics fast path of the epilogue
    if (cfg.pos_embed) { e1.res = (const T*)n->pos; e1.res_bstride = 0; }
    if (cfg.instance_norm && conv_uses_tc(e1)) { e1.stat_part = stat_part; x_stats_fused = true; }
    SFNO_TRY(launch_conv(e1, st, "encoder1"));
  }
  n->last_x_off = (size_t)((char*)cur - ws); n->last_x_bstride = cur_bs; n->last_batch = B;
  // ---- time embedding (misc.py:145-147) and all per-block time MLPs (sfnonet.py:210-213) up front
  if (cfg.with_time_emb) {
    sinusoidal_kernel<<<ceil_div(B * (C / 2), 128), 128, 0, st>>>(time, cfg.time_scaler, cfg.time_shift, B, C, tsin);
    SFNO_TRY(post_launch("sinusoidal"));
    small_linear_kernel<<<ceil_div(B * tdim * 32, 256), 256, 0, st>>>(tsin, n->te1_w, n->te1_b, th, B, tdim, C, SFNO_ACT_NONE, SFNO_ACT_GELU);
    SFNO_TRY(post_launch("time_emb_fc1"));
    small_linear_kernel<<<ceil_div(B * tdim * 32, 256), 256, 0, st>>>(th, n->te3_w, n->te3_b, trepr, B, tdim, tdim, SFNO_ACT_NONE, SFNO_ACT_NONE);
    SFNO_TRY(post_launch("time_emb_fc2"));
    small_linear_kernel<<<ceil_div(B * nl * 2 * C * 32, 256), 256, 0, st>>>(trepr, n->tmlp_w, n->tmlp_b, ts, B, nl * 2 * C, tdim, SFNO_ACT_SILU, SFNO_ACT_NONE);
    SFNO_TRY(post_launch("time_mlps"));
  }
  if (n->stop_after_block == -1) return SFNO_OK;

  const int64_t ts_bs = (int64_t)nl * 2 * C;
  // forked mask generation: only for the tensor-core path with masks, never under the per-launch profile (its events
  // live on the main stream) or with a debug tap (an early return would leave the fork unjoined)
  const bool overlap_masks = n->mask_overlap && n->side != nullptr && dropout && cfg.dropout_mlp > 0.0f && tc_allowed<T>() && (P % 8) == 0 &&
                             n->stop_after_block == -2 && !g_profile_on.load(std::memory_order_relaxed);
  auto mask_grid = [](int64_t n8) { return (unsigned)std::min<int64_t>(ceil_div64(n8 / 8 + 1, 128), 148 * 4); };
  for (int i = 0; i < nl; ++i) {
    char range_name[32];
    snprintf(range_name, sizeof(range_name), "sfno block %d", i);
    NvtxRange block_range(range_name);
    const BlockParams& bp = n->blocks[i];
    const ShtDeviceTables& fwd = (i == 0) ? n->data_grid : n->lg_grid;
    const ShtDeviceTables& inv = (i == nl - 1) ? n->data_grid : n->lg_grid;
    const bool scale_residual = fwd.grid != inv.grid;  // s2convolutions.py:79-83 (nlat/nlon equal: scale_factor 1)
    T* nxt; int64_t nxt_bs;
    buf_of(nl - 1 - i, &nxt, &nxt_bs);
    const float* ts_i = cfg.with_time_emb ? ts + (int64_t)i * 2 * C : nullptr;

    if (overlap_masks) {   // fork: this block's two keep masks are generated next to its spectral kernels
      SFNO_CUDA(cudaEventRecord(n->ev_start[i], st));
      SFNO_CUDA(cudaStreamWaitEvent(n->side, n->ev_start[i], 0));
      const float pd = cfg.dropout_mlp;
      const int64_t n8a = (int64_t)B * hid * P / 8, n8b = (int64_t)B * C * P / 8;
      dropout_mask_kernel<<<mask_grid(n8a), 128, 0, n->side>>>((uint8_t*)(ws + w.mask1), n8a, pd, seed, offset + (uint64_t)i * 4 + 0, rng_dev);
      SFNO_TRY(post_launch("dropout_mask"));
      dropout_mask_kernel<<<mask_grid(n8b), 128, 0, n->side>>>((uint8_t*)(ws + w.mask2), n8b, pd, seed, offset + (uint64_t)i * 4 + 1, rng_dev);
      SFNO_TRY(post_launch("dropout_mask"));
      SFNO_CUDA(cudaEventRecord(n->ev_mask[i], n->side));
    }

    // norm0 (+ time scale/shift before the filter) as a per-(b,c) affine  (sfnonet.py:290-299)
    const bool time_before = cfg.with_time_emb && cfg.time_scale_shift_before_filter;
    if (cfg.instance_norm && x_stats_fused) {
      norm_affine_partials_conv_kernel<<<ceil_div(BC, NAP_BC), 1024, 0, st>>>(stat_part, conv_slices, (int64_t)BC, (float)P, cfg.norm_eps,
                                                                        bp.norm0_g, bp.norm0_b, time_before ? ts_i : nullptr, ts_bs, B, C, a0, d0);
      SFNO_TRY(post_launch("norm_affine0"));
    } else {
      if (cfg.instance_norm) {
        launch_instance_stats<T>(cur, cur_bs, B, C, P, cfg.norm_eps, mean, rstd, st);
        SFNO_TRY(post_launch("instance_stats0"));
      }
      norm_affine_kernel<<<ceil_div(BC, 256), 256, 0, st>>>(cfg.instance_norm ? mean : nullptr, rstd, bp.norm0_g, bp.norm0_b,
                                                            time_before ? ts_i : nullptr, ts_bs, B, C, a0, d0);
      SFNO_TRY(post_launch("norm_affine0"));
    }

    // SpectralConvS2.forward (s2convolutions.py:158-193)
    SFNO_TRY(run_dft<T>(n, fwd, B, cur, cur_bs, a0, d0, FG, st));
    SFNO_TRY(run_leg<T>(n, fwd, B, FG, X, st));
    if (scale_residual) {  // residual = inverse_transform(forward_transform(x_norm))
      SFNO_TRY(run_ileg<T>(n, inv, B, X, FG, st));
      SFNO_TRY(run_idft<T>(n, inv, B, FG, nullptr, nullptr, 0, SFNO_ACT_NONE, res, CP, nullptr, nullptr, st));
    }
    if (cfg.operator_type == SFNO_OP_DHCONV) {
      OpDhconv<T> op{};
      op.G = n->L; op.M = n->M * B; op.N = 2 * C; op.K = 2 * C;
      op.A = X; op.Bm = (const T*)bp.wpack; op.a_sk = 1; op.b_sk = 1;
      op.y = Y; op.B = B; op.lmax = n->L; op.mmax = n->M; op.triangular = 1; op.round_out = 1;
      SFNO_TRY(launch_gemm(op, st, "dhconv"));
    } else {
      const int64_t total = (int64_t)n->L * n->M * B * C;
      diag_contract_internal_kernel<T><<<(unsigned)std::min<int64_t>(ceil_div64(total, 128), 1 << 20), 128, 0, st>>>(
          X, (const float2*)bp.wdiag, Y, B, C, n->L, n->M);
      SFNO_TRY(post_launch("diag_contract"));
    }
    SFNO_TRY(run_ileg<T>(n, inv, B, Y, FG, st));

    // inner skip (sfnonet.py:308): conv1x1 of the residual; x_norm is never materialised -- its affine is
    // folded into per-sample weights.  Output lands in t1, then the inverse DFT adds itself + bias, applies GELU.
    if (scale_residual) {
      auto sk = make_conv<T, T>(B, P, C, C, res, CP, (const T*)bp.skip_wT, 0, C, bp.skip_b, 0, SFNO_ACT_NONE, t1, CP);
      SFNO_TRY(launch_conv(sk, st, "inner_skip"));
    } else {
      fold_affine_weight_kernel<T><<<B * C, 128, 0, st>>>(bp.skip_w32, bp.skip_b, a0, d0, C, C, C, skip_wb, skip_bb, tf32 ? 1 : 0);
      SFNO_TRY(post_launch("fold_skip"));
      auto sk = make_conv<T, T>(B, P, C, C, cur, cur_bs, skip_wb, (int64_t)C * C, C, skip_bb, C, SFNO_ACT_NONE, t1, CP);
      SFNO_TRY(launch_conv(sk, st, "inner_skip"));
    }
    bool t1_stats_fused = false;
    SFNO_TRY(run_idft<T>(n, inv, B, FG, bp.spec_bias, t1, CP, cfg.activation, t1, CP, cfg.instance_norm ? stat_part : nullptr,
                         &t1_stats_fused, st));

    // norm1 (+ time scale/shift after the filter) folded into fc1  (sfnonet.py:313-323)
    const bool time_after = cfg.with_time_emb && !cfg.time_scale_shift_before_filter;
    if (cfg.instance_norm && t1_stats_fused) {
      const int idft_slices = idft_stat_slices(inv.nlon);
      norm_affine_partials_kernel<<<ceil_div(BC * 32, 256), 256, 0, st>>>(stat_part, idft_slices, (int64_t)BC * inv.Kp, inv.Kp, (float)P,
                                                                     cfg.norm_eps, bp.norm1_g, bp.norm1_b, time_after ? ts_i : nullptr,
                                                                     ts_bs, B, C, a1, d1);
      SFNO_TRY(post_launch("norm_affine1"));
    } else {
      if (cfg.instance_norm) {
        launch_instance_stats<T>(t1, CP, B, C, P, cfg.norm_eps, mean, rstd, st);
        SFNO_TRY(post_launch("instance_stats1"));
      }
      norm_affine_kernel<<<ceil_div(BC, 256), 256, 0, st>>>(cfg.instance_norm ? mean : nullptr, rstd, bp.norm1_g, bp.norm1_b,
                                                            time_after ? ts_i : nullptr, ts_bs, B, C, a1, d1);
      SFNO_TRY(post_launch("norm_affine1"));
    }

    // stochastic depth factor (drop_path.py:5-22); block 0 has rate 0 (sfnonet.py:622)
    const float dp = nl > 1 ? cfg.drop_path_rate * (float)i / (float)(nl - 1) : 0.0f;
    const bool use_dp = dropout && dp > 0.0f;
    if (use_dp) {
      drop_path_scale_kernel<<<ceil_div(B, 128), 128, 0, st>>>(dscale, B, dp, seed, offset + (uint64_t)i * 4 + 2, rng_dev);
      SFNO_TRY(post_launch("drop_path_scale"));
    }
    const float pdrop = dropout ? cfg.dropout_mlp : 0.0f;

    // MLP (layers.py:73-80) + DropPath + outer skip (sfnonet.py:326-335)
    fold_affine_weight_kernel<T><<<B * hid, 128, 0, st>>>(bp.fc1_w32, bp.fc1_b, a1, d1, hid, C, C, fc1_wb, fc1_bb, tf32 ? 1 : 0);
    SFNO_TRY(post_launch("fold_fc1"));
    auto f1 = make_conv<T, T>(B, P, C, hid, t1, CP, fc1_wb, (int64_t)hid * C, C, fc1_bb, hid, cfg.activation, hd, (int64_t)hid * P);
    f1.drop_p = pdrop; f1.seed = seed; f1.offset = offset + (uint64_t)i * 4 + 0; f1.rng_dev = rng_dev;
    // dropout masks by a dedicated kernel at full occupancy (tensor-core path; P % 8 == 0 there): inline, the Philox
    // rounds doubled the time of fc1 (3.66 vs 1.64 ms per interpolator forward, profiles/r01_o_interp.json)
    const bool masks = pdrop > 0.0f && conv_uses_tc(f1) && (P % 8) == 0;
    if (overlap_masks) SFNO_CUDA(cudaStreamWaitEvent(st, n->ev_mask[i], 0));   // join (also when this block ends up without masks)
    if (masks) {
      uint8_t* m1 = (uint8_t*)(ws + w.mask1);
      if (!overlap_masks) {
        const int64_t n8 = (int64_t)B * hid * P / 8;
        dropout_mask_kernel<<<(unsigned)std::min<int64_t>(ceil_div64(n8 / 8 + 1, 256), 148 * 8), 256, 0, st>>>(m1, n8, pdrop, seed, f1.offset, rng_dev);
        SFNO_TRY(post_launch("dropout_mask"));
      }
      f1.drop_mask = m1;
    }
    SFNO_TRY(launch_conv(f1, st, "mlp_fc1"));
    auto f2 = make_conv<T, T>(B, P, hid, C, hd, (int64_t)hid * P, (const T*)bp.fc2_wT, 0, hid, bp.fc2_b, 0, SFNO_ACT_NONE, nxt, nxt_bs);
    f2.drop_p = pdrop; f2.seed = seed; f2.offset = offset + (uint64_t)i * 4 + 1; f2.rng_dev = rng_dev;
    f2.branch_scale = use_dp ? dscale : nullptr;
    if (masks) {
      uint8_t* m2 = (uint8_t*)(ws + w.mask2);
      if (!overlap_masks) {
        const int64_t n8 = (int64_t)B * C * P / 8;
        dropout_mask_kernel<<<(unsigned)std::min<int64_t>(ceil_div64(n8 / 8 + 1, 256), 148 * 8), 256, 0, st>>>(m2, n8, pdrop, seed, f2.offset, rng_dev);
        SFNO_TRY(post_launch("dropout_mask"));
      }
      f2.drop_mask = m2;
    }
    if (scale_residual) { f2.res = res; f2.res_bstride = CP; }
    else { f2.res = cur; f2.res_bstride = cur_bs; f2.res_a = a0; f2.res_d = d0; }
    x_stats_fused = false;
    if (cfg.instance_norm && i + 1 < nl && conv_uses_tc(f2)) { f2.stat_part = stat_part; x_stats_fused = true; }
    SFNO_TRY(launch_conv(f2, st, "mlp_fc2"));

    cur = nxt; cur_bs = nxt_bs;
    n->last_x_off = (size_t)((char*)cur - ws); n->last_x_bstride = cur_bs;
    if (n->stop_after_block == i) return SFNO_OK;
  }

  // ---- decoder on cat(x, residual_big) (sfnonet.py:831-837,734-744)
  {
    const int kin = cfg.big_skip ? n->Ccat : C;
    auto d0c = make_conv<T, T>(B, P, kin, C, cur, cur_bs, (const T*)n->dec0_w, 0, n->Ccat_p, n->dec0_b, 0, cfg.activation, t1, CP);
    SFNO_TRY(launch_conv(d0c, st, "decoder0"));
    auto d1c = make_conv<T, float>(B, P, C, n->Cout, t1, CP, (const T*)n->dec1_w, 0, C, nullptr, 0, SFNO_ACT_NONE, y, (int64_t)n->Cout * P);
    d1c.round_out = 0;   // the result leaves the library in full fp32
    SFNO_TRY(launch_conv(d1c, st, "decoder1"));
  }
  // device-resident Philox state: the next forward (or the next replay of a captured graph) draws from a fresh stream
  if (rng_dev && dropout) {
    rng_advance_kernel<<<1, 1, 0, st>>>(rng_dev, (uint64_t)SFNO_RNG_OFFSETS_PER_FORWARD);
    SFNO_TRY(post_launch("rng_advance"));
  }
  return SFNO_OK;
}

}  // namespace sfno

extern "C" {

int sfno_b200_set_option(const char* key, int64_t value) {
  if (!key) return fail(SFNO_ERR_INVALID_ARGUMENT, "key is NULL");
  if (strcmp(key, "force_simt") == 0) { g_force_simt.store((int)value); return SFNO_OK; }
  if (strcmp(key, "tc_debug") == 0) { g_tc_debug.store((int)value); return SFNO_OK; }
  if (strcmp(key, "nvtx") == 0) { g_nvtx_on.store(value != 0); return SFNO_OK; }
  return fail(SFNO_ERR_INVALID_ARGUMENT, "unknown option %s", key);
}

int sfno_net_create(const sfno_net_config* c, sfno_net** out) {
  SFNO_CHECK_ARG(c && out, "NULL argument");
  SFNO_CHECK_ARG(c->struct_size == (int32_t)sizeof(sfno_net_config), "sfno_net_config size mismatch (%d vs %zu)", c->struct_size, sizeof(sfno_net_config));
  SFNO_CHECK_ARG(c->precision == SFNO_PREC_F32 || c->precision == SFNO_PREC_BF16 || c->precision == SFNO_PREC_TF32, "bad precision");
  SFNO_CHECK_ARG(c->nlat >= 2 && c->nlon >= 4 && c->in_chans > 0 && c->out_chans > 0 && c->embed_dim > 0 && c->num_layers > 0, "bad sizes");
  SFNO_CHECK_ARG(c->embed_dim % 2 == 0, "embed_dim must be even");
  SFNO_CHECK_ARG(c->lmax > 0 && c->mmax > 0 && c->mmax <= c->nlon / 2 + 1, "bad lmax/mmax");
  SFNO_CHECK_ARG(c->max_batch > 0, "max_batch must be positive");
  if (c->mlp_hidden <= 0) return fail(SFNO_ERR_UNSUPPORTED, "use_mlp=False is not supported");
  if (c->operator_type != SFNO_OP_DHCONV && c->operator_type != SFNO_OP_DIAGONAL) return fail(SFNO_ERR_UNSUPPORTED, "operator_type %d", c->operator_type);
  if (c->with_time_emb && c->embed_dim < 4) return fail(SFNO_ERR_UNSUPPORTED, "time embedding needs embed_dim >= 4");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail(SFNO_ERR_NO_DEVICE, "no CUDA device"); }

  auto* n = new sfno_net();
  n->cfg = *c;
  if (c->dropout_mlp > 0.0f) {
    bool ok = cudaStreamCreateWithFlags(&n->side, cudaStreamNonBlocking) == cudaSuccess;
    n->ev_start.assign(c->num_layers, nullptr);
    n->ev_mask.assign(c->num_layers, nullptr);
    for (int i = 0; ok && i < c->num_layers; ++i)
      ok = cudaEventCreateWithFlags(&n->ev_start[i], cudaEventDisableTiming) == cudaSuccess &&
           cudaEventCreateWithFlags(&n->ev_mask[i], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) { cudaGetLastError(); if (n->side) cudaStreamDestroy(n->side); n->side = nullptr; }
  }
  n->P = c->nlat * c->nlon; n->C = c->embed_dim; n->Cin = c->in_chans; n->Cin_p = round_up(c->in_chans, 8);
  n->Cout = c->out_chans; n->Ccat = c->embed_dim + (c->big_skip ? c->in_chans : 0); n->Ccat_p = round_up(n->Ccat, 8);
  n->hid = c->mlp_hidden; n->tdim = c->with_time_emb ? c->time_dim : 0; n->L = c->lmax; n->M = c->mmax; n->nl = c->num_layers;
  n->esize = c->precision == SFNO_PREC_BF16 ? 2 : 4;
  const size_t e = n->esize;
  const int C = n->C;
  int st = sht_tables_upload(c->nlat, c->nlon, c->lmax, c->mmax, c->data_grid, c->precision, n->data_grid);
  if (st == SFNO_OK) st = sht_tables_upload(c->nlat, c->nlon, c->lmax, c->mmax, SFNO_GRID_LEGENDRE_GAUSS, c->precision, n->lg_grid);
  auto A = [&](auto** p, size_t count, size_t es) { if (st == SFNO_OK) st = dev_alloc_t(n, p, count, es); };
  std::string names;
  auto N = [&](const std::string& s) { names += s; names += '\n'; };
  if (c->pos_embed) { A(&n->pos, (size_t)C * n->P, e); N("pos_embed"); }
  A(&n->enc0_w, (size_t)C * n->Cin_p, e); N("encoder.0.weight");
  A(&n->enc0_b, C, 4); N("encoder.0.bias");
  A(&n->enc1_w, (size_t)C * C, e); N("encoder.2.weight");
  if (c->with_time_emb) {
    A(&n->te1_w, (size_t)n->tdim * C, 4); N("time_emb_mlp.1.weight");
    A(&n->te1_b, n->tdim, 4); N("time_emb_mlp.1.bias");
    A(&n->te3_w, (size_t)n->tdim * n->tdim, 4); N("time_emb_mlp.3.weight");
    A(&n->te3_b, n->tdim, 4); N("time_emb_mlp.3.bias");
    A(&n->tmlp_w, (size_t)n->nl * 2 * C * n->tdim, 4);
    A(&n->tmlp_b, (size_t)n->nl * 2 * C, 4);
  }
  n->blocks.resize(n->nl);
  const int fc2_idx = c->dropout_mlp > 0.0f ? 3 : 2;
  for (int i = 0; i < n->nl; ++i) {
    BlockParams& b = n->blocks[i];
    const std::string p = "blocks." + std::to_string(i) + ".";
    if (c->instance_norm) {
      A(&b.norm0_g, C, 4); N(p + "norm0.weight");
      A(&b.norm0_b, C, 4); N(p + "norm0.bias");
    }
    if (c->with_time_emb) { N(p + "time_mlp.1.weight"); N(p + "time_mlp.1.bias"); }
    if (c->operator_type == SFNO_OP_DHCONV) A(&b.wpack, (size_t)n->L * 4 * C * C, e);
    else A(&b.wdiag, (size_t)C * C * n->L * n->M * 2, 4);
    N(p + "filter.filter.weight");
    A(&b.spec_bias, C, 4); N(p + "filter.filter.bias");
    A(&b.skip_w32, (size_t)C * C, 4); A(&b.skip_wT, (size_t)C * C, e); N(p + "inner_skip.weight");
    A(&b.skip_b, C, 4); N(p + "inner_skip.bias");
    if (c->instance_norm) {
      A(&b.norm1_g, C, 4); N(p + "norm1.weight");
      A(&b.norm1_b, C, 4); N(p + "norm1.bias");
    }
    A(&b.fc1_w32, (size_t)n->hid * C, 4); N(p + "mlp.fwd.0.weight");
    A(&b.fc1_b, n->hid, 4); N(p + "mlp.fwd.0.bias");
    A(&b.fc2_wT, (size_t)C * n->hid, e); N(p + "mlp.fwd." + std::to_string(fc2_idx) + ".weight");
    A(&b.fc2_b, C, 4); N(p + "mlp.fwd." + std::to_string(fc2_idx) + ".bias");
  }
  A(&n->dec0_w, (size_t)C * n->Ccat_p, e); N("decoder.0.weight");
  A(&n->dec0_b, C, 4); N("decoder.0.bias");
  A(&n->dec1_w, (size_t)n->Cout * C, e); N("decoder.2.weight");
  if (st != SFNO_OK) { sfno_net_destroy(n); return st; }
  if (!names.empty()) names.pop_back();
  n->names = names;
  *out = n;
  return SFNO_OK;
}

int sfno_net_destroy(sfno_net* n) {
  if (!n) return SFNO_OK;
  for (cudaEvent_t e : n->ev_start) if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : n->ev_mask) if (e) cudaEventDestroy(e);
  if (n->side) cudaStreamDestroy(n->side);
  for (void* p : n->allocations) cudaFree(p);
  sht_tables_free(n->data_grid);
  sht_tables_free(n->lg_grid);
  delete n;
  return SFNO_OK;
}

const char* sfno_net_param_names(const sfno_net* n) { return n ? n->names.c_str() : ""; }

int sfno_net_set_param(sfno_net* n, const char* name, const float* value_dev, int64_t numel, void* stream) {
  SFNO_CHECK_ARG(n && name && value_dev, "NULL argument");
  g_pack_tf32 = n->cfg.precision == SFNO_PREC_TF32;   // packed MMA operands are stored TF32-exact
  const int st = n->cfg.precision == SFNO_PREC_BF16 ? set_param_impl<bf16>(n, name, value_dev, numel, (cudaStream_t)stream)
                                                    : set_param_impl<float>(n, name, value_dev, numel, (cudaStream_t)stream);
  g_pack_tf32 = false;
  return st;
}

int sfno_net_set_option(sfno_net* n, const char* key, int64_t value) {
  SFNO_CHECK_ARG(n && key, "NULL argument");
  if (strcmp(key, "stop_after_block") == 0) { n->stop_after_block = (int)value; return SFNO_OK; }
  if (strcmp(key, "mask_overlap") == 0) { n->mask_overlap = value != 0 && n->side != nullptr; return SFNO_OK; }
  return fail(SFNO_ERR_INVALID_ARGUMENT, "unknown option %s", key);
}

size_t sfno_net_workspace_bytes(const sfno_net* n, int batch) {
  if (!n || batch <= 0) return 0;
  return ws_layout(n, batch).total;
}

static int forward_dispatch(sfno_net* n, const ConcatParts& parts, const float* time_dev, float* y_dev, int batch, int dropout_enabled,
                            uint64_t seed, uint64_t offset, uint64_t* rng_dev, void* workspace_dev, size_t workspace_bytes, void* stream) {
  SFNO_CHECK_ARG(n && y_dev && workspace_dev, "NULL argument");
  SFNO_CHECK_ARG(batch > 0 && batch <= n->cfg.max_batch, "batch %d outside (0, max_batch=%d]", batch, n->cfg.max_batch);
  SFNO_CHECK_ARG((time_dev != nullptr) == (n->cfg.with_time_emb != 0), "time must be given iff with_time_emb");
  SFNO_CHECK_ARG(((uintptr_t)workspace_dev & 1023) == 0, "workspace must be 1024-byte aligned");
  int csum = 0;
  for (int k = 0; k < parts.nparts; ++k) {
    SFNO_CHECK_ARG(parts.src[k] != nullptr && parts.channels[k] > 0, "input part %d is empty", k);
    csum += parts.channels[k];
  }
  if (csum != n->Cin) return fail(SFNO_ERR_SHAPE_MISMATCH, "input parts have %d channels in total, the net expects %d", csum, n->Cin);
  if (workspace_bytes < ws_layout(n, batch).total) return fail(SFNO_ERR_WORKSPACE_TOO_SMALL, "workspace too small: %zu < %zu", workspace_bytes, ws_layout(n, batch).total);
  cudaStream_t st = (cudaStream_t)stream;
  NvtxRange range("sfno_net_forward");
  return n->cfg.precision == SFNO_PREC_BF16
             ? forward_impl<bf16>(n, parts, time_dev, y_dev, batch, dropout_enabled, seed, offset, rng_dev, (char*)workspace_dev, st)
             : forward_impl<float>(n, parts, time_dev, y_dev, batch, dropout_enabled, seed, offset, rng_dev, (char*)workspace_dev, st);
}

int sfno_net_forward(sfno_net* n, const float* x_dev, const float* time_dev, float* y_dev, int batch, int dropout_enabled,
                     uint64_t seed, uint64_t offset, void* workspace_dev, size_t workspace_bytes, void* stream) {
  SFNO_CHECK_ARG(n && x_dev, "NULL argument");
  ConcatParts parts{};
  parts.src[0] = x_dev; parts.channels[0] = n->Cin; parts.nparts = 1;
  return forward_dispatch(n, parts, time_dev, y_dev, batch, dropout_enabled, seed, offset, nullptr, workspace_dev, workspace_bytes, stream);
}

int sfno_net_forward_parts(sfno_net* n, const float* const* parts_dev, const int* part_channels, int nparts, const float* time_dev,
                           float* y_dev, int batch, int dropout_enabled, uint64_t seed, uint64_t offset, void* workspace_dev,
                           size_t workspace_bytes, void* stream) {
  SFNO_CHECK_ARG(n && parts_dev && part_channels, "NULL argument");
  SFNO_CHECK_ARG(nparts >= 1 && nparts <= 3, "between 1 and 3 input parts, got %d", nparts);
  ConcatParts parts{};
  parts.nparts = nparts;
  for (int k = 0; k < nparts; ++k) { parts.src[k] = parts_dev[k]; parts.channels[k] = part_channels[k]; }
  return forward_dispatch(n, parts, time_dev, y_dev, batch, dropout_enabled, seed, offset, nullptr, workspace_dev, workspace_bytes, stream);
}

int sfno_net_forward_parts_rng(sfno_net* n, const float* const* parts_dev, const int* part_channels, int nparts, const float* time_dev,
                               float* y_dev, int batch, int dropout_enabled, uint64_t* rng_state_dev, void* workspace_dev,
                               size_t workspace_bytes, void* stream) {
  SFNO_CHECK_ARG(n && parts_dev && part_channels, "NULL argument");
  SFNO_CHECK_ARG(nparts >= 1 && nparts <= 3, "between 1 and 3 input parts, got %d", nparts);
  SFNO_CHECK_ARG(rng_state_dev != nullptr || !dropout_enabled, "dropout needs the device RNG state");
  ConcatParts parts{};
  parts.nparts = nparts;
  for (int k = 0; k < nparts; ++k) { parts.src[k] = parts_dev[k]; parts.channels[k] = part_channels[k]; }
  return forward_dispatch(n, parts, time_dev, y_dev, batch, dropout_enabled, 0, 0, rng_state_dev, workspace_dev, workspace_bytes, stream);
}

int sfno_param_fingerprint(const float* const* ptrs_dev, const int64_t* numel_dev, int count, uint64_t* out_dev, void* stream) {
  SFNO_CHECK_ARG(ptrs_dev && numel_dev && out_dev && count > 0, "bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  SFNO_CUDA(cudaMemsetAsync(out_dev, 0, (size_t)count * sizeof(uint64_t), st));
  param_fingerprint_kernel<<<dim3((unsigned)count, 64), 512, 0, st>>>(ptrs_dev, numel_dev, (unsigned long long*)out_dev);
  return post_launch("param_fingerprint");
}

int64_t sfno_net_debug_tap(sfno_net* n, const char* name, float* dst_dev, int64_t capacity, void* workspace_dev, void* stream) {
  SFNO_CHECK_ARG(n && name && dst_dev && workspace_dev, "NULL argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int B = n->last_batch;
  if (B <= 0) return fail(SFNO_ERR_INVALID_ARGUMENT, "no forward has run");
  const WsLayout w = ws_layout(n, B);
  char* ws = (char*)workspace_dev;
  if (strcmp(name, "x") == 0) {
    const int64_t per = (int64_t)n->C * n->P;
    if (capacity < per * B) return fail(SFNO_ERR_INVALID_ARGUMENT, "tap buffer too small");
    dim3 grid((unsigned)std::min<int64_t>(ceil_div64(per, 256), 2048), B);
    if (n->cfg.precision == SFNO_PREC_BF16)
      convert_planes_kernel<bf16, float><<<grid, 256, 0, st>>>((const bf16*)(ws + n->last_x_off), n->last_x_bstride, dst_dev, per, per);
    else
      convert_planes_kernel<float, float><<<grid, 256, 0, st>>>((const float*)(ws + n->last_x_off), n->last_x_bstride, dst_dev, per, per);
    int s = post_launch("tap_x");
    return s == SFNO_OK ? per * B : s;
  }
  if (strcmp(name, "t_repr") == 0) {
    const int64_t cnt = (int64_t)B * n->tdim;
    if (capacity < cnt) return fail(SFNO_ERR_INVALID_ARGUMENT, "tap buffer too small");
    SFNO_CUDA(cudaMemcpyAsync(dst_dev, ws + w.trepr, (size_t)cnt * 4, cudaMemcpyDeviceToDevice, st));
    return cnt;
  }
  return fail(SFNO_ERR_INVALID_ARGUMENT, "unknown tap %s", name);
}

}  // extern "C"
