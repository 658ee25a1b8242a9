// TMA operand descriptions of the ops for the tcgen05 engine (which tensor, which extents, which strides), for bf16
// storage (kind::f16) and fp32 storage consumed as TF32 (kind::tf32).
// Extents are the exact logical sizes: TMA zero-fills everything outside, so K / row tails need no padding.
#pragma once
#include "gemm_tc.cuh"
#include "ops.cuh"

namespace sfno {

template <class Op>
struct TcTraitsBase {
  static constexpr bool kAvailable = true;
  static constexpr bool kDualM = false;
  static bool use_dual(const Op&) { return true; }
  static constexpr bool kStationaryA = false;        // keep the CTA's A tile (shared basis / weights) resident in shared memory
  static bool use_stationary(const Op&) { return true; }
  static bool extra_ok(const Op&) { return true; }  // vector-store alignment rules of the epilogue, if any
  static void io(const Op&, TmaIo&, TmaIo&) {}       // TMA views of the output / residual tensors
  static bool has_residual(const Op&) { return false; }
};
inline bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

template <class Derived, class Op>
struct TcEligible {
  static bool eligible(const Op& op) {
    TmaOperand a, b;
    Derived::operands(op, a, b);
    if (!(tma_operand_ok(a) && tma_operand_ok(b) && op.M > 0 && op.N > 0 && op.K > 0 && Derived::extra_ok(op))) return false;
    // the epilogue moves tiles by TMA only: the output (and the residual, if any) must be expressible as boxes
    TmaIo o, r;
    Derived::io(op, o, r);
    return tma_io_ok(o) && (!Derived::has_residual(op) || tma_io_ok(r));
  }
};

template <class T>
struct TcTraits<OpDft<T>> : TcTraitsBase<OpDft<T>>, TcEligible<TcTraits<OpDft<T>>, OpDft<T>> {
  static constexpr int BN = 192;
  // (stationary A -- each CTA keeps its 128 basis rows resident and streams planes -- was measured: 1.74 vs 1.67 ms per
  //  forward for the streaming kernel, profiles/r02_i_stationary_ab.txt; the kernel is bound by the latency of the plane
  //  loads with 3-4 stages in flight, not by the bytes of the basis, so it stays off here)
  static constexpr bool kStationaryA = false;
  static constexpr uint64_t es = sizeof(T);
  static void operands(const OpDft<T>& op, TmaOperand& a, TmaOperand& b) {
    a.base = op.A; a.dims[0] = op.nlon; a.dims[1] = op.M;                 // basis rows (m,ri), K-contiguous, shared
    a.strides[0] = (uint64_t)op.Wp * es; a.batched = false;
    if (op.a_reps > 1) { a.dims[2] = op.a_reps; a.strides[1] = (uint64_t)op.M * op.Wp * es; a.replicas = op.a_reps; }
    b.base = op.Bm; b.dims[0] = op.nlon; b.dims[1] = op.nlat; b.dims[2] = op.C; b.dims[3] = op.B;  // {j, k, c, b}
    b.strides[0] = (uint64_t)op.nlon * es; b.strides[1] = (uint64_t)op.nlat * op.nlon * es; b.strides[2] = (uint64_t)op.x_bstride * es;
    b.batched = true; b.group_lo = op.C;
  }
  static bool extra_ok(const OpDft<T>& op) { return aligned16(op.f) && op.Kp % 8 == 0; }
  static void io(const OpDft<T>& op, TmaIo& o, TmaIo&) {
    const uint64_t kp = (uint64_t)op.Kp * es;
    o.base = op.f; o.es = (int)es; o.ok = true;
    o.dims[0] = op.Kp; o.dims[1] = op.C; o.dims[2] = 2; o.dims[3] = op.B; o.dims[4] = op.M / 2;
    o.strides[0] = kp; o.strides[1] = kp * op.C; o.strides[2] = kp * op.C * 2; o.strides[3] = kp * op.C * 2 * op.B;
    o.box_rows[0] = 1; o.box_rows[1] = 2; o.box_rows[2] = 1; o.box_rows[3] = 16;   // 32 GEMM rows = 16 wavenumbers x {re, im}
  }
};

template <class T>
struct TcTraits<OpLeg<T>> : TcTraitsBase<OpLeg<T>>, TcEligible<TcTraits<OpLeg<T>>, OpLeg<T>> {
  static constexpr uint64_t es = sizeof(T);
  // dual-M (two 128-row A tiles per B tile, single-buffered accumulators): the 180 degrees of a wavenumber fit ONE
  // 256-row tile, which halves the F traffic.  Measured (profiles/r01_j_tc_dual_ab.txt): -8 % for the triangular
  // Legendre; +30 % / +9 % / +28 % for DFT / dhconv / inverse Legendre (losing the second accumulator stage costs
  // more than the saved L2 traffic), so only this op uses it, and only with the triangular ranges.
  static constexpr bool kDualM = true;
  static bool use_dual(const OpLeg<T>& op) { return op.triangular != 0; }
  static constexpr int BN = 256;
  static void operands(const OpLeg<T>& op, TmaOperand& a, TmaOperand& b) {
    a.base = op.A; a.dims[0] = op.K; a.dims[1] = op.lmax; a.dims[2] = op.G;   // table rows l of wavenumber m
    a.strides[0] = (uint64_t)op.Kp * es; a.strides[1] = (uint64_t)op.lmax * op.Kp * es; a.batched = true;
    b.base = op.Bm; b.dims[0] = op.K; b.dims[1] = op.N; b.dims[2] = op.G;      // F rows (b,ri,c) of wavenumber m
    b.strides[0] = (uint64_t)op.Kp * es; b.strides[1] = (uint64_t)op.N * op.Kp * es; b.batched = true;
  }
  static bool extra_ok(const OpLeg<T>& op) { return aligned16(op.x) && op.N % 8 == 0; }
  static void io(const OpLeg<T>& op, TmaIo& o, TmaIo&) {
    o.base = op.x; o.es = (int)es; o.ok = true;
    o.dims[0] = op.N; o.dims[1] = op.mmax; o.dims[2] = op.lmax;
    o.strides[0] = (uint64_t)op.N * es; o.strides[1] = (uint64_t)op.mmax * op.N * es;
    o.box_rows[0] = 1; o.box_rows[1] = 32;   // 32 GEMM rows = 32 degrees of one wavenumber
  }
};

template <class T>
struct TcTraits<OpDhconv<T>> : TcTraitsBase<OpDhconv<T>>, TcEligible<TcTraits<OpDhconv<T>>, OpDhconv<T>> {
  static constexpr int BN = 256;
  static constexpr uint64_t es = sizeof(T);
  static void operands(const OpDhconv<T>& op, TmaOperand& a, TmaOperand& b) {
    a.base = op.A; a.dims[0] = op.K; a.dims[1] = op.M; a.dims[2] = op.G;       // X rows (m,b) of degree l
    a.strides[0] = (uint64_t)op.K * es; a.strides[1] = (uint64_t)op.M * op.K * es; a.batched = true;
    b.base = op.Bm; b.dims[0] = op.K; b.dims[1] = op.N; b.dims[2] = op.G;       // packed weight rows (ri',o) of degree l
    b.strides[0] = (uint64_t)op.K * es; b.strides[1] = (uint64_t)op.N * op.K * es; b.batched = true;
  }
  static bool extra_ok(const OpDhconv<T>& op) { return aligned16(op.y) && op.N % 8 == 0; }
  static void io(const OpDhconv<T>& op, TmaIo& o, TmaIo&) {
    o.base = op.y; o.es = (int)es; o.ok = true;
    o.dims[0] = op.N; o.dims[1] = op.M; o.dims[2] = op.G;
    o.strides[0] = (uint64_t)op.N * es; o.strides[1] = (uint64_t)op.M * op.N * es;
    o.box_rows[0] = 32;   // rows (m,b) of one degree; the live range of a degree ends on a multiple of 64 * B
  }
};

template <class T>
struct TcTraits<OpIleg<T>> : TcTraitsBase<OpIleg<T>>, TcEligible<TcTraits<OpIleg<T>>, OpIleg<T>> {
  static constexpr int BN = 192;
  static constexpr uint64_t es = sizeof(T);
  static void operands(const OpIleg<T>& op, TmaOperand& a, TmaOperand& b) {
    a.base = op.A; a.dims[0] = op.M; a.dims[1] = op.K; a.dims[2] = op.G;  // M-contiguous: {(b,ri,o), l, m}
    a.strides[0] = (uint64_t)op.a_sk * es; a.strides[1] = (uint64_t)op.a_goff * es; a.batched = true;
    b.base = op.Bm; b.dims[0] = op.K; b.dims[1] = op.nlat; b.dims[2] = op.G;
    b.strides[0] = (uint64_t)op.Lq * es; b.strides[1] = (uint64_t)op.nlat * op.Lq * es; b.batched = true;
  }
  static bool extra_ok(const OpIleg<T>& op) { return aligned16(op.g_out) && op.Kp % 8 == 0; }
  static void io(const OpIleg<T>& op, TmaIo& o, TmaIo&) {
    const uint64_t kp = (uint64_t)op.Kp * es;
    o.base = op.g_out; o.es = (int)es; o.ok = op.C % 8 == 0;
    o.dims[0] = op.Kp; o.dims[1] = op.C; o.dims[2] = op.B; o.dims[3] = 2; o.dims[4] = op.G;
    o.strides[0] = kp; o.strides[1] = kp * op.C; o.strides[2] = kp * op.C * op.B; o.strides[3] = kp * op.C * op.B * 2;
    if (op.C % 32 == 0) o.box_rows[0] = 32;   // 32 GEMM rows (b,ri,o) = 32 channels of one (b, ri)
  }
};

template <class T, class TOut, int ACT>
struct TcTraits<OpIdft<T, TOut, ACT>> : TcTraitsBase<OpIdft<T, TOut, ACT>>,
                                          TcEligible<TcTraits<OpIdft<T, TOut, ACT>>, OpIdft<T, TOut, ACT>> {
  static constexpr int BN = 192;
  static constexpr uint64_t ei = sizeof(T);
  static void operands(const OpIdft<T, TOut, ACT>& op, TmaOperand& a, TmaOperand& b) {
    a.base = op.A; a.dims[0] = op.M; a.dims[1] = op.K; a.dims[2] = 1;  // M-contiguous: {(b,o,kp), (m,ri)}
    a.strides[0] = (uint64_t)op.a_sk * ei; a.batched = false;
    b.base = op.Bm; b.dims[0] = op.K; b.dims[1] = op.N; b.dims[2] = 1;
    b.strides[0] = (uint64_t)op.Kq2 * ei; b.batched = false;
    if (op.b_reps > 1) { b.dims[2] = op.b_reps; b.strides[1] = (uint64_t)op.N * op.Kq2 * ei; b.replicas = op.b_reps; }
  }
  // the addend has the element type of the operands (T) and is staged through the output's staging rows: same size
  static bool extra_ok(const OpIdft<T, TOut, ACT>& op) {
    return op.nlon % 8 == 0 && op.out_bstride % 8 == 0 && aligned16(op.out) &&
           (!op.add || (sizeof(TOut) == sizeof(T) && aligned16(op.add) && op.add_bstride % 8 == 0));
  }
  static bool has_residual(const OpIdft<T, TOut, ACT>& op) { return op.add != nullptr; }
  static void io(const OpIdft<T, TOut, ACT>& op, TmaIo& o, TmaIo& r) {
    const uint64_t es = sizeof(TOut);
    const int B = op.M / (op.C * op.Kp);
    o.base = op.out; o.es = (int)es; o.ok = op.Kp % 8 == 0;
    o.dims[0] = op.nlon; o.dims[1] = op.nlat; o.dims[2] = op.C; o.dims[3] = B;
    o.strides[0] = (uint64_t)op.nlon * es; o.strides[1] = (uint64_t)op.nlat * op.nlon * es; o.strides[2] = (uint64_t)op.out_bstride * es;
    if (op.add && es == sizeof(T)) {   // loaded in 8-row boxes: a clipped 32-row LOAD would zero-fill rows of the other plane
      r = o;
      r.base = op.add; r.strides[2] = (uint64_t)op.add_bstride * es;
    }
    o.box_rows[0] = 32;   // stores: one 32-latitude box per warp; warps whose rows straddle two planes use 8-row boxes
  }
};

// engine experiment: N tile of the 1x1 convolutions (192 -> 12 epilogue warps at <= 127 registers, 4 operand stages;
// 256 -> 16 epilogue warps at <= 112 registers, 3 stages)
#ifndef SFNO_TC_CONV_BN
#define SFNO_TC_CONV_BN 192
#endif

template <class T, class TOut, int ACT, int DROP>
struct TcTraits<OpConv<T, TOut, ACT, DROP>> : TcTraitsBase<OpConv<T, TOut, ACT, DROP>>,
                                                TcEligible<TcTraits<OpConv<T, TOut, ACT, DROP>>, OpConv<T, TOut, ACT, DROP>> {
  static constexpr int BN = SFNO_TC_CONV_BN;
  // Stationary A: the CTA keeps its 128 output channels x all input channels of the weights resident and streams pixel
  // tiles.  Used for SHARED weights that fit next to a >= 3-stage ring (encoder, decoder: decoder0 0.179 -> 0.148 ms,
  // decoder1 0.082 -> 0.071 ms); with per-sample folded weights (inner skip, fc1) it measured slower than the streaming
  // kernel (fc1 2.03 vs 1.78 ms per forward, profiles/r02_i_stationary_ab.txt), fc2 (K = 512) does not fit.
  static constexpr bool kStationaryA = true;
  static bool use_stationary(const OpConv<T, TOut, ACT, DROP>& op) { return op.w_bstride == 0; }
  static constexpr uint64_t ei = sizeof(T);
  static void operands(const OpConv<T, TOut, ACT, DROP>& op, TmaOperand& a, TmaOperand& b) {
    const bool wb = op.w_bstride != 0;
    a.base = op.A; a.dims[0] = op.K; a.dims[1] = op.M; a.dims[2] = wb ? op.G : 1;  // weights [o][c], K-contiguous
    a.strides[0] = (uint64_t)op.ldw * ei; a.strides[1] = (uint64_t)op.w_bstride * ei; a.batched = wb;
    b.base = op.Bm; b.dims[0] = op.N; b.dims[1] = op.K; b.dims[2] = op.G;  // N-contiguous: {pixel, channel, sample}
    b.strides[0] = (uint64_t)op.N * ei; b.strides[1] = (uint64_t)op.in_bstride * ei; b.batched = true;
  }
  // the residual has the element type of the operands (T) and is staged through the output's staging rows: same size
  static bool extra_ok(const OpConv<T, TOut, ACT, DROP>& op) {
    return op.N % 8 == 0 && op.out_bstride % 8 == 0 && aligned16(op.out) && (op.ldw * ei) % 16 == 0 &&
           (!op.res || (sizeof(TOut) == sizeof(T) && aligned16(op.res) && op.res_bstride % 8 == 0)) && (!op.pos || aligned16(op.pos));
  }
  static bool has_residual(const OpConv<T, TOut, ACT, DROP>& op) { return op.res != nullptr; }
  static void io(const OpConv<T, TOut, ACT, DROP>& op, TmaIo& o, TmaIo& r) {
    const uint64_t es = sizeof(TOut);
    o.base = op.out; o.es = (int)es; o.ok = true;
    o.dims[0] = op.N; o.dims[1] = op.M; o.dims[2] = op.G;
    o.strides[0] = (uint64_t)op.N * es; o.strides[1] = (uint64_t)op.out_bstride * es;
    o.box_rows[0] = 32;   // 32 output channels of one sample (rows past the last channel are clipped / zero-filled)
    if (op.res && es == sizeof(T)) {
      r = o;
      r.base = op.res;
      if (op.res_bstride != 0) r.strides[1] = (uint64_t)op.res_bstride * es;
      else r.dims[2] = 1;   // one plane shared by all samples (position embedding)
    }
  }
};

// SFNO_TC_LDTM_PAIR experiment: the ops whose epilogue holds two 32-column accumulator chunks without spilling
template <> struct TcLdtmPair<OpDft<bf16>> { static constexpr bool value = true; };
template <> struct TcLdtmPair<OpIleg<bf16>> { static constexpr bool value = true; };
template <int ACT> struct TcLdtmPair<OpIdft<bf16, bf16, ACT>> { static constexpr bool value = true; };

}  // namespace sfno
