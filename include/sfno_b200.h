/*
 * sfno_b200.h -- C ABI of the B200-native SFNO forward path (libsfno_b200.so).
 *
 * The reference (Rose-STL-Lab/spherical-dyffusion) is pure Python and has no FFI for this path; the
 * entry points below are what a binding for its hot path has to call ("What a
 * C-ABI replacement must export").  Each group cites the reference interface it replaces.
 *
 * Conventions
 *   - plain C types only; every pointer named *_dev is a CUDA device pointer owned by the caller
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises
 *   - functions return 0 (SFNO_OK) or a negative sfno_status; no exceptions cross the ABI
 *   - forward entry points never allocate: scratch comes from the caller-provided workspace
 *   - handles are re-entrant per handle (one in-flight call per handle and stream)
 *   - `precision`: SFNO_PREC_F32 = fp32 storage, fp32 CUDA-core FMA accumulate (parity mode,
 *                  <= 1e-4 rel-L2 vs the fp32 reference); SFNO_PREC_BF16 = bf16 storage, tcgen05
 *                  tensor-core MMA with fp32 TMEM accumulators (throughput mode, stated bf16 bound);
 *                  SFNO_PREC_TF32 = fp32 storage, tcgen05 kind::tf32 MMA (operands rounded to TF32 by
 *                  their producers, fp32 accumulation): what the reference computes under
 *                  torch.set_float32_matmul_precision("high") (src/utilities/config_utils.py:310-313)
 */
#ifndef SFNO_B200_H_
#define SFNO_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum sfno_status {
  SFNO_OK = 0,
  SFNO_ERR_INVALID_ARGUMENT = -1,
  SFNO_ERR_CUDA = -2,
  SFNO_ERR_WORKSPACE_TOO_SMALL = -3,
  SFNO_ERR_UNSUPPORTED = -4,
  SFNO_ERR_NO_DEVICE = -5,
  SFNO_ERR_UNKNOWN_PARAM = -6,
  SFNO_ERR_SHAPE_MISMATCH = -7
} sfno_status;

enum { SFNO_GRID_LEGENDRE_GAUSS = 0, SFNO_GRID_EQUIANGULAR = 1 };
enum { SFNO_PREC_F32 = 0, SFNO_PREC_BF16 = 1, SFNO_PREC_TF32 = 2 };
enum { SFNO_OP_DHCONV = 0, SFNO_OP_DIAGONAL = 1 };
enum { SFNO_ACT_NONE = 0, SFNO_ACT_GELU = 1, SFNO_ACT_RELU = 2, SFNO_ACT_SILU = 3 };

/* ---- library ---------------------------------------------------------------------------------- */
int sfno_b200_abi_version(void);
const char* sfno_b200_status_string(int status);
/* Thread-local text of the last failure (CUDA error string, offending argument ...). */
const char* sfno_b200_last_error(void);
/* Number of kernels this library has launched since load (monotonic; used by bench.py). */
int64_t sfno_b200_launch_count(void);

/* Library-wide switches for tests: "force_simt" = 1 routes bf16 ops to the CUDA-core engine.
 * "tc_debug" = bit mask of TIMING-ONLY experiment switches of the tensor-core engine (results are wrong while any of
 * bits 0-4 is set): 1 skip A-operand loads, 2 skip B-operand loads, 4 skip global stores, 16 skip the MMAs;
 * correct-result switches: 128 = accumulate the role-wait counters read by sfno_b200_tc_counters, 256 = single 128-row
 * tiles where an op would use dual-M tiles, 512 = every CTA walks the K blocks from block 0 (no per-CTA rotation),
 * 1024 = plain stream order (no programmatic dependent launch).
 * "nvtx" = 1 (or SFNO_NVTX=1 in the environment): the network forward, its blocks and the op-level entry points push
 * named NVTX ranges for tracing tools; off by default. */
int sfno_b200_set_option(const char* key, int64_t value);

/* Measurement hook: cycles summed over the CTAs of all tensor-core launches since the last call (tc_debug bit 7):
 * out12 = {TMA producer blocked on a free stage, MMA issuer blocked on operands, MMA issuer blocked on a free
 * accumulator, epilogue warp 0 blocked on the accumulator, epilogue warp 0 blocked on the residual block,
 * CTA lifetime, epilogue warp 0 busy, number of CTAs, epilogue warp 0: per-tile set-up, drain loop, store + hand-over,
 * reserved}.  Synchronises the device and clears the counters. */
int sfno_b200_tc_counters(unsigned long long* out12);

/* Per-launch device timing for bench.py: between _begin and _end every kernel launched by this library on
 * `stream` is followed by a CUDA event; _end synchronises the stream and returns the number of launches, their
 * names ('\n'-separated) and durations in ms (difference of consecutive events on that stream). */
int sfno_b200_profile_begin(void* stream);
int sfno_b200_profile_end(char* names, size_t names_capacity, float* ms, int capacity);

/* Device-vs-device engine cross-check used by the GPU tests: runs op `op_kind` (0 DFT, 1 Legendre, 2 dhconv,
 * 3 inverse Legendre, 4 inverse DFT, 5 conv1x1) of the given dims on synthetic bf16 operands through the CUDA-core
 * engine and the tcgen05 engine; result = {max |diff|, max |ref|, tensor-core engine used (0/1), ms per launch, #non-finite}. */
int sfno_b200_selftest_gemm(int op_kind, const int* dims, int ndims, double* result);

/* ---- host-side tables (no GPU needed) -----------------------------------------------------------
 * Replaces torch_harmonics' precompute used at sfnonet.py:551-554 (quadrature.py legendre_gauss_weights /
 * clenshaw_curtiss_weights, legendre.py legpoly).  Outputs are fp64, row-major:
 *   nodes[nlat] (cos(colatitude), north -> south), quad_w[nlat],
 *   weights[mmax][lmax][nlat] = P * quad_w (analysis), pct[mmax][lmax][nlat] (synthesis).
 * Any output pointer may be NULL. */
int sfno_sht_tables_host(int nlat, int nlon, int lmax, int mmax, int grid,
                         double* nodes, double* quad_w, double* weights, double* pct);

/* ---- spherical harmonic transform pair ----------------------------------------------------------
 * Replaces torch_harmonics.RealSHT / InverseRealSHT objects built at sfnonet.py:551-554 and called at
 * s2convolutions.py:165,168,186.  The plan owns the device tables (Legendre x quadrature, DFT bases). */
typedef struct sfno_sht_plan sfno_sht_plan;
int sfno_sht_plan_create(sfno_sht_plan** plan, int nlat, int nlon, int lmax, int mmax, int grid, int precision);
int sfno_sht_plan_destroy(sfno_sht_plan* plan);
/* scratch needed by sfno_sht_forward / sfno_sht_inverse for `fields` = prod(leading dims) 2-D fields */
size_t sfno_sht_workspace_bytes(const sfno_sht_plan* plan, int64_t fields);
/* x_dev: fp32 [fields][nlat][nlon]  ->  coeffs_dev: complex64 (interleaved re,im) [fields][lmax][mmax]
 * == RealSHT.forward: 2*pi*rfft(x, norm="forward")[..., :mmax] contracted with weights[m,l,k]. */
int sfno_sht_forward(const sfno_sht_plan* plan, const float* x_dev, float* coeffs_dev, int64_t fields,
                     void* workspace_dev, size_t workspace_bytes, void* stream);
/* coeffs_dev: complex64 [fields][lmax][mmax] -> x_dev fp32 [fields][nlat][nlon]
 * == InverseRealSHT.forward: pct contraction then irfft(n=nlon, norm="forward") (Im of m=0/Nyquist dropped). */
int sfno_sht_inverse(const sfno_sht_plan* plan, const float* coeffs_dev, float* x_dev, int64_t fields,
                     void* workspace_dev, size_t workspace_bytes, void* stream);

/* ---- spectral channel contraction -----------------------------------------------------------------
 * Replaces _contract_dhconv / _contract_diagonal (contractions.py:147-169) as dispatched by
 * get_contract_fun (factorizations.py:189-225) and called at s2convolutions.py:172-179.
 *   x_dev      complex64 [batch][cin][lmax][mmax]
 *   weight_dev fp32 real view [cin][cout][lmax][2] (dhconv) or [cin][cout][lmax][mmax][2] (diagonal)
 *   out_dev    complex64 [batch][cout][lmax][mmax]
 * fp32 CUDA-core arithmetic; the packed tensor-core form is sfno_spectral_conv (and lives inside sfno_net_forward). */
int sfno_spectral_contract(int operator_type, const float* x_dev, const float* weight_dev, float* out_dev,
                           int batch, int cin, int cout, int lmax, int mmax, void* stream);

/* ---- fused spectral convolution and precision-selectable pieces (SURVEY 8b "must export") -----------
 * sfno_spectral_conv replaces SpectralConvS2.forward (src/models/sfno/s2convolutions.py:158-193) as ONE call:
 *   X = RealSHT(x);  residual = InverseRealSHT(X) if the two transforms differ (:166-169), else the caller keeps x;
 *   Y = contract(X, weight) (:172-179);  y = InverseRealSHT(Y) + bias (:186-189)
 * on the tensor-core ops of the plans' precision (SFNO_PREC_BF16: kind::f16, SFNO_PREC_TF32: kind::tf32, SFNO_PREC_F32:
 * CUDA-core parity engine).  The weight handle owns the packed form of filter.weight ([cin][cout][lmax][2] dhconv,
 * [cin][cout][lmax][mmax][2] diagonal; fp32, reference layout) and filter.bias ([cout] or NULL); call _set again after an
 * in-place update.  x_dev [batch][cin][nlat][nlon] fp32 (forward plan's grid), y_dev [batch][cout][nlat'][nlon'] fp32
 * (inverse plan's grid), residual_dev NULL or [batch][cin][nlat'][nlon']. */
typedef struct sfno_spectral_weight sfno_spectral_weight;
int sfno_spectral_weight_create(sfno_spectral_weight** weight, int operator_type, int cin, int cout, int lmax, int mmax,
                                int precision);
int sfno_spectral_weight_set(sfno_spectral_weight* weight, const float* weight_dev, const float* bias_dev, void* stream);
int sfno_spectral_weight_destroy(sfno_spectral_weight* weight);
size_t sfno_spectral_conv_workspace_bytes(const sfno_sht_plan* fwd, const sfno_sht_plan* inv,
                                          const sfno_spectral_weight* weight, int batch);
int sfno_spectral_conv(const sfno_sht_plan* fwd, const sfno_sht_plan* inv, const sfno_spectral_weight* weight,
                       const float* x_dev, float* y_dev, float* residual_dev, int batch, void* workspace_dev,
                       size_t workspace_bytes, void* stream);
/* Backward of sfno_spectral_conv (what autograd derives for s2convolutions.py:158-193): from the cotangents grad_y_dev
 * [batch][cout][nlat'][nlon'] and grad_residual_dev (NULL or [batch][cin][nlat'][nlon'], the scale_residual output) it
 * writes grad_x_dev [batch][cin][nlat][nlon] (NULL: skipped), grad_weight_dev (layout of filter.weight; NULL: skipped;
 * needs x_dev) and grad_bias_dev [cout] (NULL: skipped).  weight_dev is filter.weight itself (fp32, reference layout).
 * The transposed transforms and the conjugate-transposed contraction run on the ops of the plans' precision; the weight
 * gradient is an fp32 per-degree GEMM over the (wavenumber, sample) rows. */
size_t sfno_spectral_conv_backward_workspace_bytes(const sfno_sht_plan* fwd, const sfno_sht_plan* inv,
                                                   const sfno_spectral_weight* weight, int batch);
int sfno_spectral_conv_backward(sfno_sht_plan* fwd, sfno_sht_plan* inv, const sfno_spectral_weight* weight,
                                const float* weight_dev, const float* x_dev, const float* grad_y_dev,
                                const float* grad_residual_dev, float* grad_x_dev, float* grad_weight_dev,
                                float* grad_bias_dev, int batch, void* workspace_dev, size_t workspace_bytes, void* stream);
/* nn.Conv2d(cin, cout, 1) with the fused epilogue of the hot path -- + bias -> activation -> dropout(p; Philox stream
 * (seed, offset); layers.py:76-80, live at inference: dyffusion.py:226-235) + residual -- and a selectable engine:
 * SFNO_PREC_BF16 / SFNO_PREC_TF32 stage the operands in the workspace and run the tensor-core op, SFNO_PREC_F32 is
 * sfno_conv1x1 (no workspace needed).  Tensors as sfno_conv1x1. */
size_t sfno_conv1x1_ex_workspace_bytes(int batch, int cin, int cout, int64_t hw, int precision);
int sfno_conv1x1_ex(const float* x_dev, const float* weight_dev, const float* bias_dev, const float* residual_dev,
                    float* y_dev, int batch, int cin, int cout, int64_t hw, int activation, float dropout_p,
                    uint64_t seed, uint64_t offset, int precision, void* workspace_dev, size_t workspace_bytes,
                    void* stream);

/* ---- pointwise pieces (fp32, NCHW) -- exported so each can be parity-tested on its own ------------- */
/* nn.InstanceNorm2d(C, eps, affine=True, track_running_stats=False) (sfnonet.py:641-647), optionally
 * fused with time_scale_shift (sfnonet.py:280-287): y = norm(x)*(scale+1)+shift with scale/shift [batch][C]
 * (either may be NULL).  x_dev,y_dev: [batch][C][hw]. */
int sfno_instance_norm(const float* x_dev, float* y_dev, const float* gamma_dev, const float* beta_dev,
                       const float* scale_dev, const float* shift_dev, int batch, int channels, int64_t hw,
                       float eps, void* workspace_dev, size_t workspace_bytes, void* stream);
size_t sfno_instance_norm_workspace_bytes(int batch, int channels);
/* nn.Conv2d(cin, cout, 1) (+bias) (+activation) (+residual add): sfnonet.py:239,308,614-617,739-742, layers.py:73-75.
 * x_dev [batch][cin][hw], weight_dev [cout][cin], bias_dev [cout] or NULL, residual_dev [batch][cout][hw] or NULL. */
int sfno_conv1x1(const float* x_dev, const float* weight_dev, const float* bias_dev, const float* residual_dev,
                 float* y_dev, int batch, int cin, int cout, int64_t hw, int activation, void* stream);

/* ---- whole network -----------------------------------------------------------------------------------
 * Replaces SphericalFourierNeuralOperatorNet (sfnonet.py:340-841) for inference: ctor arguments of
 * sfnonet.py:426-466 that reach the hot path, the state_dict layout of SURVEY 8b, and forward()
 * (sfnonet.py:797-841) including concat_condition_if_needed (_base_model.py:166-192, done by the host
 * wrapper), time embedding (misc.py:132-148), time_scale_shift (sfnonet.py:280-287), SpectralConvS2.forward
 * (s2convolutions.py:158-193), MLP (layers.py:73-80), Dropout / DropPath at inference (dyffusion.py:226-235). */
typedef struct sfno_net_config {
  int32_t struct_size;         /* = sizeof(sfno_net_config), for ABI evolution */
  int32_t precision;           /* SFNO_PREC_* */
  int32_t nlat, nlon;          /* spatial_shape_in (== out; scale_factor must be 1) */
  int32_t in_chans;            /* num_input_channels + num_conditional_channels */
  int32_t out_chans;           /* num_output_channels */
  int32_t embed_dim, num_layers;
  int32_t mlp_hidden;          /* int(embed_dim * mlp_ratio); 0 = use_mlp False */
  int32_t operator_type;       /* SFNO_OP_* */
  int32_t activation;          /* SFNO_ACT_* */
  int32_t data_grid;           /* SFNO_GRID_* of trans_down / itrans_up; internal grid is Legendre-Gauss */
  int32_t lmax, mmax;          /* modes_lat, modes_lon (sfnonet.py:526-527) */
  int32_t pos_embed, big_skip; /* booleans */
  int32_t instance_norm;       /* 1 = instance_norm, 0 = none */
  int32_t with_time_emb, time_dim;            /* time_dim = embed_dim * time_dim_mult */
  int32_t time_scale_shift_before_filter;
  float   time_scaler, time_shift;            /* time_rescale (sfnonet.py:783-784); 1, 0 when off */
  float   norm_eps;                           /* 1e-6 */
  float   dropout_mlp;                        /* p of nn.Dropout inside MLP (layers.py:76-80) */
  float   drop_path_rate;                     /* linspace(0, rate, num_layers) (sfnonet.py:622) */
  int32_t max_batch;                          /* workspace / per-sample folded-weight capacity */
} sfno_net_config;

typedef struct sfno_net sfno_net;
int sfno_net_create(const sfno_net_config* config, sfno_net** net);
int sfno_net_destroy(sfno_net* net);
/* Upload one tensor of the reference state_dict (fp32, contiguous, device memory, reference shape).
 * `name` is the reference key, e.g. "blocks.3.filter.filter.weight".  The call (re)packs it into the
 * layout the kernels use; call again after an in-place update (EMA swap, load_state_dict). */
int sfno_net_set_param(sfno_net* net, const char* name, const float* value_dev, int64_t numel, void* stream);
/* Names the net expects, '\n'-separated (owned by the net). */
const char* sfno_net_param_names(const sfno_net* net);
size_t sfno_net_workspace_bytes(const sfno_net* net, int batch);
/* x_dev [batch][in_chans][nlat][nlon] fp32, time_dev [batch] fp32 (already range-checked by the caller;
 * NULL iff !with_time_emb), y_dev [batch][out_chans][nlat][nlon] fp32.
 * dropout_enabled: nn.Dropout/DropPath in "training" state (inference dropout); seed/offset key the
 * Philox stream so members differ only by their RNG stream. */
int sfno_net_forward(sfno_net* net, const float* x_dev, const float* time_dev, float* y_dev, int batch,
                     int dropout_enabled, uint64_t seed, uint64_t offset,
                     void* workspace_dev, size_t workspace_bytes, void* stream);
/* Same forward with the channel concat of BaseModel.concat_condition_if_needed (_base_model.py:166-192) fused into the
 * input conversion: parts_dev[k] is fp32 [batch][part_channels[k]][nlat][nlon] (inputs, condition, static_condition in
 * that order; 1 <= nparts <= 3, channel counts must add up to in_chans). */
int sfno_net_forward_parts(sfno_net* net, const float* const* parts_dev, const int* part_channels, int nparts,
                           const float* time_dev, float* y_dev, int batch, int dropout_enabled, uint64_t seed, uint64_t offset,
                           void* workspace_dev, size_t workspace_bytes, void* stream);
/* Same forward with a DEVICE-RESIDENT Philox state instead of host seed / offset integers: rng_state_dev[2] =
 * {seed, offset} (uint64).  Masks are keyed on the values found there when the kernels run, and the call enqueues
 * offset += 4096 after its last kernel, so a CUDA graph that captured this call draws fresh masks on every replay
 * (what torch's graph-safe generator does for nn.Dropout, dyffusion.py:226-235).  rng_state_dev may be NULL iff
 * dropout_enabled == 0. */
int sfno_net_forward_parts_rng(sfno_net* net, const float* const* parts_dev, const int* part_channels, int nparts,
                               const float* time_dev, float* y_dev, int batch, int dropout_enabled, uint64_t* rng_state_dev,
                               void* workspace_dev, size_t workspace_bytes, void* stream);
/* Per-net test switch: "stop_after_block" = -2 run everything (default), -1 stop after encoder + pos-embed,
 * i stop after block i (the output tensor is then left untouched; read the state with sfno_net_debug_tap). */
int sfno_net_set_option(sfno_net* net, const char* key, int64_t value);
/* Debug/test taps: copy an intermediate of the last forward (same workspace) as fp32 into dst_dev.
 * Known names: "x" (current activation [batch][embed][nlat][nlon]), "t_repr".  Returns the element count or a negative status. */
int64_t sfno_net_debug_tap(sfno_net* net, const char* name, float* dst_dev, int64_t capacity,
                           void* workspace_dev, void* stream);

/* ---- ensemble statistics (new; spec = src/evaluation/metrics.py:166-175,199-246) -------------------
 * Variance is the reference's two-pass `predicted.var(dim=0)` (metrics.py:166-175), never sum(x^2) - E*mean^2:
 *   1. every rank:  sfno_ensemble_local_sum        -> sum_r[n] over ITS members (members may be 0 on a rank)
 *   2. all-reduce(SUM) of sum_r (NCCL, host side)   -> sum[n]; pivot p = sum / E, bit-identical on every rank
 *   3. every rank:  sfno_ensemble_shifted_moments  -> moments_r[2][n] = { sum_e (x - p), sum_e (x - p)^2 }
 *   4. all-reduce(SUM) of moments_r                  -> moments[2][n]
 *   5. sfno_ensemble_finalize -> mean = p + S1 / E, var = (S2 - S1^2 / E) / (E - 1)   (unbiased, as torch.var)
 * members_dev [members][n] fp32 on this rank. */
int sfno_ensemble_local_sum(const float* members_dev, int members, int64_t n, float* sum_dev, void* stream);
int sfno_ensemble_shifted_moments(const float* members_dev, int members, int64_t n, const float* sum_global_dev,
                                  int total_members, float* moments_dev, void* stream);
int sfno_ensemble_finalize(const float* sum_global_dev, const float* moments_dev, int total_members, int64_t n,
                           float* mean_dev, float* var_dev, void* stream);
/* One pass over ALL members (after the all-gather): mean_dev[n], var_dev[n] (two-pass, unbiased) and the fair CRPS per
 * grid point (metrics.py:199-246): crps = mean_i|x_i - y| - sum_{i<j}|x_i-x_j| / (E*(E-1)), sorted form, no [E,E,..]
 * tensor.  Any output may be NULL; truth_dev may be NULL iff crps_dev is.  members <= 64. */
int sfno_ensemble_stats(const float* members_dev, const float* truth_dev, int members, int64_t n, float* mean_dev,
                        float* var_dev, float* crps_dev, void* stream);
int sfno_ensemble_crps(const float* members_dev, const float* truth_dev, int members, int64_t n,
                       float* crps_dev, void* stream);
/* Same on a subset of the rows of members_dev: rows_dev[members] (int32, device) are the row indices of the live members
 * in the all-gathered buffer of UNEVEN shards (25 members over 8 ranks: 4+3+...: every rank contributes max-local rows,
 * the surplus rows are padding).  The statistics are invariant under member order, so nothing is re-ordered or compacted. */
int sfno_ensemble_stats_rows(const float* members_dev, const int* rows_dev, const float* truth_dev, int members, int64_t n,
                             float* mean_dev, float* var_dev, float* crps_dev, void* stream);

/* ---- sampler glue (caller of the hot path) ---------------------------------------------
 * Cold-sampling update of BaseDYffusion.sample_loop (src/diffusion/dyffusion.py:519):
 *   x_out = x_s + (x_next - x_cur)   in one pass (x_out may alias x_s); fp32, n elements. */
int sfno_cold_update(const float* x_s_dev, const float* x_next_dev, const float* x_cur_dev, float* x_out_dev,
                     int64_t n, void* stream);

/* ---- rollout step glue -----------------------------------------------------------------
 * The state of an autoregressive rollout stays packed [batch][channels][hw] on the device; each side of the sampler is
 * one launch.
 * sfno_normalize_pack: StandardNormalizer.normalize (src/ace_inference/core/normalizer.py:96-112) + Packer.pack
 *   (packer.py:71-77): out[b][c][p] = (fields[c][b][p] - mean[c]) / std[c]; fields_dev is a DEVICE array of `channels`
 *   device pointers to independent [batch][hw] fp32 tensors (the reference's dict of named fields); mean_dev / std_dev
 *   [channels] or NULL (0 / 1). */
int sfno_normalize_pack(const float* const* fields_dev, int channels, int batch, int64_t hw, const float* mean_dev,
                        const float* std_dev, float* out_dev, void* stream);
/* sfno_prescribe_denormalize: Prescriber.__call__ (src/ace_inference/core/prescriber.py:68-95) on the packed normalised
 *   prediction, in place (the result seeds the next window, stepper_multistep.py:402-421), fused with
 *   StandardNormalizer.denormalize (normalizer.py:105-112) into gen_denorm_dev (may be NULL):
 *     channel == prescribed_channel:  interpolate ? mask * target + (1 - mask) * gen
 *                                                 : (int(round(mask)) == mask_value ? target : gen)
 *   target_norm_dev / mask_dev: [batch][hw] with the given sample strides (0 = one field shared by all samples);
 *   prescribed_channel < 0: no prescriber (denormalise only). */
int sfno_prescribe_denormalize(float* gen_norm_dev, const float* target_norm_dev, int64_t target_bstride,
                               const float* mask_dev, int64_t mask_bstride, int prescribed_channel, int mask_value,
                               int interpolate, const float* mean_dev, const float* std_dev, float* gen_denorm_dev,
                               int channels, int batch, int64_t hw, void* stream);

/* ---- backward pass of the op-level entry points -------------------------------------------
 * Autograd formulas of the custom ops: everything a training step of the reference's modules needs when their transforms,
 * contraction, 1x1 convolutions and norms run through this library.
 * Adjoint transforms: RealSHT / InverseRealSHT are real-linear maps between [fields][nlat][nlon] and the (re, im) pairs of
 * [fields][lmax][mmax]; the gradient of a real loss w.r.t. their input is the transposed map applied to the gradient
 * w.r.t. their output.  They run on the plan's engine (fp32 / tf32 / bf16) as the opposite transform's GEMM ops on
 * transposed tables (built on first use; workspace as sfno_sht_workspace_bytes). */
int sfno_sht_forward_adjoint(sfno_sht_plan* plan, const float* grad_coeffs_dev, float* grad_x_dev, int64_t fields,
                             void* workspace_dev, size_t workspace_bytes, void* stream);
int sfno_sht_inverse_adjoint(sfno_sht_plan* plan, const float* grad_x_dev, float* grad_coeffs_dev, int64_t fields,
                             void* workspace_dev, size_t workspace_bytes, void* stream);
/* _contract_dhconv / _contract_diagonal (contractions.py:147-169): grad_x = sum_o grad_out conj(w), grad_w = sum_{b(,m)}
 * conj(x) grad_out (complex64, interleaved, reference layouts as sfno_spectral_contract); either output may be NULL. */
int sfno_spectral_contract_backward(int operator_type, const float* x_dev, const float* weight_dev, const float* grad_out_dev,
                                    float* grad_x_dev, float* grad_w_dev, int batch, int cin, int cout, int lmax, int mmax,
                                    void* stream);
/* nn.Conv2d(cin, cout, 1): grad_w[cout][cin] = sum_{b,p} grad_y[b][o][p] x[b][c][p] (split-K partial GEMMs + reduction),
 * grad_b[cout] = sum_{b,p} grad_y (grad_b_dev may be NULL).  grad_x is sfno_conv1x1 of grad_y with the transposed weight. */
size_t sfno_conv1x1_weight_grad_workspace_bytes(int batch, int cin, int cout, int64_t hw);
int sfno_conv1x1_weight_grad(const float* x_dev, const float* grad_y_dev, float* grad_w_dev, float* grad_b_dev, int batch,
                             int cin, int cout, int64_t hw, void* workspace_dev, size_t workspace_bytes, void* stream);
/* nn.InstanceNorm2d (+ the per-(b,c) affine y = xhat * A + D that sfno_instance_norm fuses: A = gamma (1 + scale),
 * D = beta (1 + scale) + shift): grad_x [batch][C][hw], and per plane grad_a = sum grad_out xhat, grad_d = sum grad_out
 * ([batch][C]; the chain rule to gamma / beta / scale / shift is host-side arithmetic on these small arrays).
 * affine_a_dev [batch][C] or NULL (= 1). */
int sfno_instance_norm_backward(const float* x_dev, const float* grad_out_dev, const float* affine_a_dev, float* grad_x_dev,
                                float* grad_a_dev, float* grad_d_dev, int batch, int channels, int64_t hw, float eps,
                                void* stream);
/* Backward of sfno_conv1x1_ex (activation NONE, no dropout) on the engine of `precision`: grad_x_dev [batch][cin][hw]
 * (NULL: skipped; needs weight_dev [cout][cin]) = the forward op with the transposed weight; grad_w_dev [cout][cin]
 * (NULL: skipped; needs x_dev) = split-K GEMM over the pixels -- tensor cores with fp32 partial sums in bf16 / tf32
 * -- and grad_b_dev [cout] (NULL: skipped). */
size_t sfno_conv1x1_backward_workspace_bytes(int batch, int cin, int cout, int64_t hw, int precision);
int sfno_conv1x1_backward(const float* x_dev, const float* grad_y_dev, const float* weight_dev, float* grad_x_dev,
                          float* grad_w_dev, float* grad_b_dev, int batch, int cin, int cout, int64_t hw, int precision,
                          void* workspace_dev, size_t workspace_bytes, void* stream);

/* ---- parameter fingerprints ---------------------------------------------------------------------------
 * out_dev[i] = position-weighted 64-bit checksum of the bit patterns of tensor i (ptrs_dev[i], numel_dev[i] fp32
 * elements), one launch for all tensors.  The host wrapper compares it with the value recorded at the last
 * sfno_net_set_param to detect in-place updates that bypass autograd's version counter (`p.data.copy_`, used by the
 * reference's EMA swap src/models/modules/ema.py:54-91).  out_dev is overwritten. */
int sfno_param_fingerprint(const float* const* ptrs_dev, const int64_t* numel_dev, int count, uint64_t* out_dev,
                           void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SFNO_B200_H_ */
