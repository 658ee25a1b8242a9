// tcgen05 / TMA / TMEM engine for the bf16 batched-GEMM ops (SFNO_PREC_BF16), sm_100a only.
//
//   persistent CTAs (one per SM), 128 x BN output tile, K in blocks of 64 bf16 (one 128-byte swizzle span)
//   warp 0      : TMA producer   (cp.async.bulk.tensor, SWIZZLE_128B, 4-D maps {inner, outer, batch, batch_hi})
//   warp 1      : MMA issuer     (tcgen05.mma.cta_group::1.kind::f16, fp32 accumulators in TMEM, 2 stages)
//   warps 2..   : epilogue       BN/64 warps per TMEM sub-partition, each draining a 64-column slice:
//                                 tcgen05.ld 32x32b.x32 -> fused op epilogue on packed fp32 pairs -> the warp's private
//                                 32 x 128-byte swizzled staging rows -> TMA store (8-row x 128-byte boxes of a <= 5-D
//                                 view of the output); residual / addend blocks arrive by TMA load the same way.
//                                 No LDG / STG in the epilogue.
//   (BN = 256: 576 threads cap the kernel at 96 registers per thread, BN = 192: 448 threads / 128 registers; moving
//    registers between roles with setmaxnreg was tried -- 640 threads, 56/104 -- and lost)
//   smem ring of kStages {A tile, B tile}, mbarrier full/empty; TMEM full/empty barriers decouple the MMA of
//   tile i+1 from the epilogue of tile i.
//
// Operands are consumed directly in the layout they have in HBM: K-contiguous operands as K-major
// SWIZZLE_128B tiles, M/N-contiguous operands as MN-major SWIZZLE_128B tiles (64-element atoms), so no
// transposition pass exists anywhere on the path.  Out-of-bounds rows / K are zero-filled by TMA.
// Launched with programmatic stream serialisation: the set-up of kernel n+1 overlaps the tail of kernel n.
#pragma once
#include <cuda.h>

#include <type_traits>

#include "common.cuh"

namespace sfno {

#ifndef SFNO_TC_WAIT_HINT_NS
#define SFNO_TC_WAIT_HINT_NS 0    // > 0: the single-thread roles wait with mbarrier.try_wait's suspend-time hint instead of nanosleep polling
#endif
#ifndef SFNO_TC_BACKOFF_NS
#define SFNO_TC_BACKOFF_NS 64   // sleep of the single-thread roles (TMA producer, MMA issuer) between barrier polls
#endif
#ifndef SFNO_TC_LDTM_PAIR
#define SFNO_TC_LDTM_PAIR 0   // 1: paired TMEM loads in the bf16 drain loop (experiment, see the drain loop)
#endif

constexpr int TC_BM = 128;
constexpr int TC_BK = 64;          // K block of the bf16 engine in elements (= one 128-byte swizzle span)
constexpr int TC_SLICE_COLS = 64;  // accumulator columns drained by one epilogue warp

// Operand element of the engine: bf16 (kind::f16) or fp32 storage consumed as TF32 (kind::tf32: the tensor core reads
// the upper 19 bits; producers round to TF32 so that nothing is truncated).  One K block is always one 128-byte swizzle
// span and one MMA always covers 32 bytes of K, so shared-memory stage sizes and MMA counts per block do not depend on
// the element; only the element counts do.
template <class TIn> struct TcElem;
template <> struct TcElem<bf16> {
  static constexpr int kBytes = 2, kBK = 64, kUmmaK = 16, kAtom = 64;   // kAtom: elements of one 128-byte MN-major atom row
  static constexpr uint32_t kFormat = 1;                                 // UMMA F16F32Format::BF16
  static constexpr CUtensorMapDataType kTmaType = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
};
template <> struct TcElem<float> {
  static constexpr int kBytes = 4, kBK = 32, kUmmaK = 8, kAtom = 32;
  static constexpr uint32_t kFormat = 2;                                 // UMMA F16F32Format::TF32
  static constexpr CUtensorMapDataType kTmaType = CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
};
constexpr int tc_epi_warps(int bn) { return 4 * (bn / TC_SLICE_COLS); }   // 4 TMEM sub-partitions x column slices
constexpr int tc_threads(int bn) { return 64 + 32 * tc_epi_warps(bn); }
constexpr int TC_WARP_TMA = 0, TC_WARP_MMA = 1;

struct TmaOperand {
  const void* base = nullptr;
  uint64_t dims[4] = {1, 1, 1, 1};   // {inner, outer, batch, batch_hi} in elements
  uint64_t strides[3] = {0, 0, 0};   // byte strides of dims[1..3]
  bool batched = false;
  int group_lo = 0;                  // > 0: group g addresses {g % group_lo, g / group_lo} in dims[2], dims[3]
  int replicas = 0;                  // > 1: a shared operand stored `replicas` times along dims[2]; CTA i reads copy i % replicas
};

// Epilogue I/O through TMA: the output (and the residual / addend block) as a 5-D tensor whose 128-byte-wide boxes
// are the swizzled staging rows of an epilogue warp: either ONE box of all 32 rows (one TMA instruction per warp and
// pass) or four boxes of 8 rows.  A warp-divergent TMA instruction is executed as a serial loop over the issuing
// lanes whose iterations wait for the previous one to release its uniform registers (measured: 14-26 % of the stall
// samples of the conv / DFT kernels, profiles/r02_c_source_stalls.md), so the 32-row form is used whenever 32
// consecutive GEMM rows map to a box of the tensor (box_rows = extents of dims[1..4], product 32 or 8); shapes that
// cannot guarantee even 8 are not eligible for this engine.
struct TmaIo {
  const void* base = nullptr;
  int es = 2;                               // element size: 2 (bf16) or 4 (fp32)
  uint64_t dims[5] = {1, 1, 1, 1, 1};
  uint64_t strides[4] = {0, 0, 0, 0};       // byte strides of dims[1..4]
  uint32_t box_rows[4] = {8, 1, 1, 1};
  bool ok = false;
};

struct TcSched {
  int m_tiles, n_tiles, groups, num_tiles, k_blocks, k16_last;  // k16_last: MMAs (32 bytes of K each) in the last k block
  int a_batched, b_batched;
  int a_glo, b_glo;  // > 0: the operand's group index splits into {g % glo, g / glo} (4-D tensor map)
  int a_rep, b_rep;  // > 1: shared operand replicated along dims[2]: this CTA reads copy blockIdx.x % rep
  int out_box32, res_box32;   // the output / residual tensor map has 32-row boxes (one TMA instruction per warp and pass)
  // stationary-A kernels: all K blocks of the CTA's A tile stay resident in shared memory, the ring holds B only
  int stat_kb, stat_stages;   // K blocks of the resident A tile (= k_blocks), B stages of the ring
  // role-wait profile (tc_debug bit7) or nullptr: cycles summed over CTAs {producer waits for a free stage, MMA waits
  // for operands, MMA waits for a free accumulator, epilogue warp 0 waits for the accumulator, epilogue warp 0 waits
  // for the residual block, CTA lifetime, epilogue warp 0 busy, CTAs}
  unsigned long long* prof;
  // experiment switch (sfno_b200_set_option("tc_debug")), WRONG results, timing only: bit0 skip A loads, bit1 skip B
  // loads, bit2 skip global stores, bit4 skip the MMAs (bits 3 and 5 -- epilogue math, TMEM loads -- were removed
  // from the drain loop once measured: profiles/r01_j_tc_dbg_sweep.txt); correct-result switches: bit7 role-wait
  // counters, bit8 single-M tiles, bit9 every CTA walks K from block 0 (no rotation)
  int dbg;
};

extern std::atomic<int> g_tc_debug;
unsigned long long* tc_prof_buffer();  // 12 device counters (allocated on first use)

// Tile walk of a persistent CTA without per-tile divisions: tile = (c * J + b) * I + a advances by a fixed step.
struct TileIter {
  int a, b, c, da, db, dc, I, J;
  __device__ void init(int tile0, int step, int I_, int J_) {
    I = I_; J = J_;
    a = tile0 % I; int r = tile0 / I; b = r % J; c = r / J;
    da = step % I; r = step / I; db = r % J; dc = r / J;
  }
  __device__ void next() {
    a += da; int carry = a >= I; a -= carry ? I : 0;
    b += db + carry; carry = b >= J; b -= carry ? J : 0;
    c += dc + carry;
  }
};

// ops that take the paired TMEM loads of the SFNO_TC_LDTM_PAIR experiment (specialised in gemm_tc_ops.cuh)
template <class Op>
struct TcLdtmPair { static constexpr bool value = false; };

template <class Op>
struct TcTraits {
  static constexpr bool kAvailable = false;
  static constexpr bool kDualM = false;
  static constexpr bool kStationaryA = false;
  static bool eligible(const Op&) { return false; }
};

// ---- PTX wrappers ----------------------------------------------------------------------------------------
namespace ptx {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// the same with a suspend-time hint (ns): the hardware may park the thread for up to that long and wakes it when the phase
// completes -- a blocked single-thread role then issues no polling instructions on its scheduler
__device__ __forceinline__ bool mbar_try_wait_hint(uint32_t bar, uint32_t parity, uint32_t hint_ns) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(hint_ns)
      : "memory");
  return ok != 0;
}
// Blocking wait with a watchdog: a pipeline bug traps (-> CUDA error) after ~2 s instead of hanging the GPU.
template <bool kBackoff = false>
__device__ __forceinline__ long long mbar_wait(uint32_t bar, uint32_t parity, bool timed = false) {
  if (!timed) {
    if (mbar_try_wait(bar, parity)) return 0;
    const long long t0 = clock64();
#if SFNO_TC_WAIT_HINT_NS > 0
    if (kBackoff) {   // single-thread roles: parked by the hardware between polls
      while (!mbar_try_wait_hint(bar, parity, SFNO_TC_WAIT_HINT_NS)) {
        if (clock64() - t0 > 4000000000LL) __trap();
      }
      return 0;
    }
#endif
    while (!mbar_try_wait(bar, parity)) {
      if (kBackoff) __nanosleep(SFNO_TC_BACKOFF_NS);  // single-thread roles with slack: do not compete with the epilogue warps for issue slots
      if (clock64() - t0 > 4000000000LL) __trap();
    }
    return 0;
  }
  // role-wait profile (tc_debug bit7): try_wait itself may block for a while, so it sits inside the timed region
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (kBackoff) __nanosleep(SFNO_TC_BACKOFF_NS);
    if (clock64() - t0 > 4000000000LL) __trap();
  }
  return clock64() - t0;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, const int (&c)[5]) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(c[4])
      : "memory");
}
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* map, uint32_t src, const int (&c)[5]) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(map), "r"(src),
               "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(c[4])
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
template <class TIn>
__device__ __forceinline__ void mma_elem(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  if constexpr (sizeof(TIn) == 2) mma_bf16(d_tmem, a_desc, b_desc, idesc, accumulate);
  else mma_tf32(d_tmem, a_desc, b_desc, idesc, accumulate);
}
// mbarrier arrive when all previously issued MMAs of this thread have completed (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, "
      "%25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
}  // namespace ptx

// ---- descriptors -----------------------------------------------------------------------------------------
// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, version 1); layout 2 = SWIZZLE_128B (16-byte chunks),
// 1 = SWIZZLE_128B_BASE32B (32-byte chunks, 4-row period): the only layout the hardware accepts for MN-major tf32 operands
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout = 2) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  d |= (uint64_t)layout << 61;
  return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): fmt x fmt -> fp32, M = 128; fmt 1 = bf16, 2 = tf32
__host__ __device__ constexpr uint32_t make_idesc(bool a_mn_major, bool b_mn_major, int n, uint32_t fmt = 1) {
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
}

constexpr int TC_STAGE_PITCH = 128;                       // bytes per staged row (one pass of an epilogue warp)
constexpr int TC_STAGING_PER_WARP = 32 * TC_STAGE_PITCH;  // 32 rows

template <int BN, bool kStaging, bool kDual = false, int kBufs = 1>
struct TcSmem {
  static constexpr int kAHalfBytes = TC_BM * TC_BK * 2;          // rows x 128 bytes: the same for bf16 (BK 64) and tf32 (BK 32)
  static constexpr int kABytes = (kDual ? 2 : 1) * kAHalfBytes;   // dual-M: two 128-row A tiles share one B tile
  static constexpr int kBBytes = BN * TC_BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  // kBufs = 2: the TMA store of one tile reads staging buffer b while the next tile fills b^1 (epilogue-bound ops);
  // kBufs = 1 leaves the shared memory to the operand ring (load-bound ops)
  static constexpr int kStagingBytes = kStaging ? kBufs * tc_epi_warps(BN) * TC_STAGING_PER_WARP : 0;
  static constexpr int kBudget = 226 * 1024 - 1024 /*alignment slack*/ - 512 /*barriers*/ - kStagingBytes;
  static constexpr int kStages = kBudget / kStageBytes > 6 ? 6 : kBudget / kStageBytes;
  static constexpr int kBarrierBytes = 512;
  static constexpr int kTotal = kStages * kStageBytes + kBarrierBytes + kStagingBytes + 1024;
  static_assert(kStages >= 2, "not enough shared memory for a pipeline");
};

// 16-byte shared-memory accessors
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void unpack_bf16x8(const uint4& u, float (&v)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x; v[2 * i + 1] = f.y;
  }
}

// kStatA: "stationary A".  Every CTA owns ONE M tile for the whole launch and keeps all K blocks of its A tile (the
// operand shared by the groups: DFT basis, convolution weights) resident in shared memory; the ring streams only B.
// The L2 -> shared-memory operand traffic of a tile drops from A + B to B (forward DFT: 245 -> 147 KB per tile,
// fc1: 160 -> 96 KB): these kernels were bound by operand delivery (profiles/r02_g_tc_dbg_sweep.txt).  A is reloaded
// only when the group changes and A is per group (per-sample folded weights: 8 times per launch).
template <class Op, int BN, bool kDual, bool kStatA = false>
__global__ void __launch_bounds__(tc_threads(BN), 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
               const __grid_constant__ CUtensorMap tma_out, const __grid_constant__ CUtensorMap tma_res,
               const __grid_constant__ CUtensorMap tma_out8, const Op op, const TcSched sc) {
  using S = TcSmem<BN, Op::kColContig, kDual, Op::kStagingBufs>;
  using E = TcElem<typename Op::InT>;
  constexpr int kBK = E::kBK;              // K block in elements (one 128-byte swizzle span)
  static_assert(!(kStatA && kDual), "stationary A and dual-M tiles are not combined");
  const int kStages = kStatA ? sc.stat_stages : S::kStages;   // (a constant for the streaming kernels)
  constexpr uint32_t kTmemCols = 512;  // two accumulator stages of up to 256 fp32 columns, or (dual-M) one stage of two
  constexpr int kBMT = kDual ? 2 * TC_BM : TC_BM;   // rows per tile
  constexpr int kAccStages = kDual ? 1 : 2;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B atoms need 1024-byte alignment
  // streaming: ring of {A, B} stages.  stationary A: [resident A: stat_kb x 16 KB][ring of B stages]
  const uint32_t ring_base = smem_base + (kStatA ? (uint32_t)sc.stat_kb * S::kAHalfBytes : 0u);
  const uint32_t kRingStage = kStatA ? (uint32_t)S::kBBytes : (uint32_t)S::kStageBytes;
  const uint32_t staging_base = ring_base + (uint32_t)kStages * kRingStage;  // 1024-byte aligned: TMA boxes of staged rows
  const uint32_t bar_base = staging_base + S::kStagingBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kStages + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 4);
  auto res_bar = [&](int w) { return bar_base + 8u * (2 * kStages + 6 + w); };  // one per epilogue warp (TMA residual loads)
  const uint32_t afull_bar = bar_base + 8u * (2 * kStages + 6 + tc_epi_warps(BN)), afree_bar = afull_bar + 8u;   // stationary A
  auto a_smem = [&](int s, int h = 0) { return kStatA ? smem_base + (uint32_t)s * S::kAHalfBytes   // s = K block of the resident tile
                                                      : ring_base + (uint32_t)s * kRingStage + (uint32_t)h * S::kAHalfBytes; };
  auto b_smem = [&](int s) { return ring_base + (uint32_t)s * kRingStage + (kStatA ? 0u : (uint32_t)S::kABytes); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool timed = sc.prof != nullptr;   // role-wait / phase timers only when somebody reads them

  if (warp == TC_WARP_TMA && lane == 0) {
    ptx::prefetch_tmap(&tma_a);
    ptx::prefetch_tmap(&tma_b);
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(tfull_bar(a), 1);
      ptx::mbar_init(tempty_bar(a), tc_epi_warps(BN));
    }
    for (int w = 0; w < tc_epi_warps(BN); ++w) ptx::mbar_init(res_bar(w), 1);
    if (kStatA) { ptx::mbar_init(afull_bar, 1); ptx::mbar_init(afree_bar, 1); }
    ptx::fence_barrier_init();
  }
  if (warp == TC_WARP_MMA) ptx::tmem_alloc(tmem_slot, kTmemCols);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  // Programmatic dependent launch: everything above (barrier init, TMEM allocation, tensor-map prefetch) may run while
  // the previous kernel of the stream drains its last tiles; nothing below touches global memory before the
  // prerequisite grid has completed and flushed.  The next kernel may start its own set-up as soon as SMs free up.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");

  // Tile order.  Consecutive tiles run concurrently on different SMs:
  //   default     : they share the operand that is re-read -- the N tiles of one M tile when the activations sit on
  //                 the A side (kNFastest), the M tiles of one N tile when they sit on the B side -- (L2 reuse);
  //   kGFastest   : they belong to different groups, so that no two CTAs pull the same per-group table / weight tile at
  //                 the same time (a broadcast read of one small region serialises on a few L2 slices).
  TileIter it;
  if (Op::kGFastest) it.init(blockIdx.x, gridDim.x, sc.groups, Op::kNFastest ? sc.n_tiles : sc.m_tiles);
  else it.init(blockIdx.x, gridDim.x, Op::kNFastest ? sc.n_tiles : sc.m_tiles, Op::kNFastest ? sc.m_tiles : sc.n_tiles);
  // stationary A: the CTA's M tile is fixed; it walks a CONTIGUOUS chunk of the (group, N tile) pairs q = g * n_tiles + nt
  // (group-major), so that a per-group A tile is reloaded at most once or twice per CTA.  The m_tiles CTAs of one chunk
  // run side by side and share its B tiles in L2.  gridDim.x is a multiple of m_tiles.
  const int sq_mt = kStatA ? (int)(blockIdx.x % (unsigned)sc.m_tiles) : 0;
  const int sq_all = kStatA ? sc.groups * sc.n_tiles : 0, sq_lanes = kStatA ? (int)(gridDim.x / (unsigned)sc.m_tiles) : 1;
  const int sq_chunk = (sq_all + sq_lanes - 1) / sq_lanes;
  const int sq_step = 1;
  int sq = kStatA ? (int)(blockIdx.x / (unsigned)sc.m_tiles) * sq_chunk : 0;
  const int sq_total = kStatA ? (sq + sq_chunk < sq_all ? sq + sq_chunk : sq_all) : 0;
  auto walk_valid = [&](int tile) { return kStatA ? sq < sq_total : tile < sc.num_tiles; };
  auto walk_next = [&]() { if (kStatA) sq += sq_step; else it.next(); };
  auto decode = [&](int& g, int& mt, int& nt) {
    if (kStatA) {
      g = sq / sc.n_tiles; nt = sq - g * sc.n_tiles; mt = sq_mt;
    } else if (Op::kGFastest) {
      g = it.a;
      if (Op::kNFastest) { nt = it.b; mt = it.c; } else { mt = it.b; nt = it.c; }
    } else {
      g = it.c;
      if (Op::kNFastest) { nt = it.a; mt = it.b; } else { mt = it.a; nt = it.b; }
    }
  };

  // tiles outside the group's live row / column ranges carry no information: all roles skip them
  auto tile_skipped = [&](int g, int mt, int nt) {
    if constexpr (!Op::kRanged) return false;
    else return nt * BN >= op.n_end(g) || (nt + 1) * BN <= op.n_begin(g) || op.m_begin(g) + mt * kBMT >= op.m_end(g);
  };

  if (warp == TC_WARP_TMA) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0, stat_g = -1;
      uint32_t phase = 0, afree_phase = 0;
      long long w_empty = 0;
      const int rep_a = sc.a_rep > 1 ? (int)(blockIdx.x % (unsigned)sc.a_rep) : 0;
      const int rep_b = sc.b_rep > 1 ? (int)(blockIdx.x % (unsigned)sc.b_rep) : 0;
      const long long t_cta = timed ? clock64() : 0;
      for (int tile = blockIdx.x; walk_valid(tile); tile += gridDim.x, walk_next()) {
        int g, mt, nt;
        decode(g, mt, nt);
        if (tile_skipped(g, mt, nt)) continue;
        const int m0 = op.m_begin(g) + mt * kBMT, n0 = nt * BN;   // M tiles start at the group's first live row
        const int halves = (kDual && m0 + TC_BM < op.m_end(g)) ? 2 : 1;
        int ga = sc.a_batched ? g : rep_a, gb = sc.b_batched ? g : rep_b, ga_hi = 0, gb_hi = 0;
        if (sc.a_glo) { ga_hi = ga / sc.a_glo; ga -= ga_hi * sc.a_glo; }
        if (sc.b_glo) { gb_hi = gb / sc.b_glo; gb -= gb_hi * sc.b_glo; }
        // every CTA walks the K blocks of its tile from its own starting block (and wraps): at any instant the CTAs
        // that share an operand (basis, table, weights) read different K slices of it, which spreads the broadcast
        // over more L2 slices.  fp32 accumulation order differs per CTA but is fixed by the static schedule.
        const int kbeg = op.k_begin(g) / kBK, nkb = sc.k_blocks - kbeg;
        const int rot = (kStatA || Op::kRanged || (sc.dbg & 512)) ? 0 : (int)(blockIdx.x % (unsigned)nkb);   // (spectral ops: measured neutral / slower)
        if constexpr (kStatA) {
          // (re)load the resident A tile: first tile of the CTA, and whenever a per-group A meets a new group
          if (stat_g == -1 || (sc.a_batched && g != stat_g)) {
            if (stat_g != -1) { w_empty += ptx::mbar_wait<true>(afree_bar, afree_phase, timed); afree_phase ^= 1u; }   // MMAs on the old tile retired
            stat_g = g;
            if (!(sc.dbg & 1)) {
              ptx::mbar_expect_tx(afull_bar, (uint32_t)sc.stat_kb * S::kAHalfBytes);
              for (int kb = 0; kb < sc.stat_kb; ++kb) {
                if (Op::A_KCONTIG) {
                  ptx::tma_load_4d(a_smem(kb), &tma_a, afull_bar, kb * kBK, m0, ga, ga_hi);
                } else {
#pragma unroll
                  for (int h = 0; h < TC_BM / E::kAtom; ++h)
                    ptx::tma_load_4d(a_smem(kb) + h * (kBK * 128), &tma_a, afull_bar, m0 + E::kAtom * h, kb * kBK, ga, ga_hi);
                }
              }
            } else {
              ptx::mbar_arrive(afull_bar);
            }
          }
        }
        for (int ik = 0; ik < nkb; ++ik) {
          const int kb = kbeg + (ik + rot < nkb ? ik + rot : ik + rot - nkb);
          w_empty += ptx::mbar_wait<true>(empty_bar(stage), phase ^ 1u, timed);
          const bool load_a = !kStatA && !(sc.dbg & 1), load_b = !(sc.dbg & 2);
          ptx::mbar_expect_tx(full_bar(stage), (load_a ? halves * S::kAHalfBytes : 0) + (load_b ? S::kBBytes : 0));
          const int k0 = kb * kBK;
          // MN-major operands arrive as atoms of E::kAtom elements (128 bytes) x kBK k-rows
          for (int hf = 0; hf < (load_a ? halves : 0); ++hf) {
            const int mh = m0 + hf * TC_BM;
            if (Op::A_KCONTIG) {
              ptx::tma_load_4d(a_smem(stage, hf), &tma_a, full_bar(stage), k0, mh, ga, ga_hi);
            } else {
#pragma unroll
              for (int h = 0; h < TC_BM / E::kAtom; ++h)
                ptx::tma_load_4d(a_smem(stage, hf) + h * (kBK * 128), &tma_a, full_bar(stage), mh + E::kAtom * h, k0, ga, ga_hi);
            }
          }
          if (!load_b) {
          } else if (Op::B_KCONTIG) {
            ptx::tma_load_4d(b_smem(stage), &tma_b, full_bar(stage), k0, n0, gb, gb_hi);
          } else {
#pragma unroll
            for (int h = 0; h < BN / E::kAtom; ++h)
              ptx::tma_load_4d(b_smem(stage) + h * (kBK * 128), &tma_b, full_bar(stage), n0 + E::kAtom * h, k0, gb, gb_hi);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
      if (sc.prof) {
        atomicAdd(sc.prof + 0, (unsigned long long)w_empty);
        atomicAdd(sc.prof + 5, (unsigned long long)(clock64() - t_cta));
        atomicAdd(sc.prof + 7, 1ull);
      }
    }
  } else if (warp == TC_WARP_MMA) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(!Op::A_KCONTIG, !Op::B_KCONTIG, BN, E::kFormat);
      // K-major: 8-row atoms of 1024 B (SBO), K advance 32 B per MMA.  MN-major: 128-byte atoms, next atom along
      // M/N after BK k-rows of 128 B (LBO), next 8 k-rows after 1024 B (SBO), K advance = kUmmaK k-rows of 128 B per MMA
      // (bf16: LBO 8192, 2048 B per MMA; tf32: LBO 4096, 1024 B per MMA).
      // MN-major tf32 operands use the 32-byte-chunk swizzle (TMA: SWIZZLE_128B_ATOM_32B): atoms of 4 k-rows (SBO 512).
      constexpr uint32_t kMnStep = (uint32_t)(E::kUmmaK * 128);
      constexpr bool kMn32 = sizeof(typename Op::InT) == 4;
      constexpr uint32_t a_lbo = Op::A_KCONTIG ? 16u : (uint32_t)(kBK * 128), a_kstep = Op::A_KCONTIG ? 32u : kMnStep;
      constexpr uint32_t b_lbo = Op::B_KCONTIG ? 16u : (uint32_t)(kBK * 128), b_kstep = Op::B_KCONTIG ? 32u : kMnStep;
      constexpr uint32_t a_sbo = (!Op::A_KCONTIG && kMn32) ? 512u : 1024u, a_layout = (!Op::A_KCONTIG && kMn32) ? 1u : 2u;
      constexpr uint32_t b_sbo = (!Op::B_KCONTIG && kMn32) ? 512u : 1024u, b_layout = (!Op::B_KCONTIG && kMn32) ? 1u : 2u;
      int stage = 0, acc = 0, stat_g = -1;
      uint32_t phase = 0, acc_phase = 0, afull_phase = 0;
      long long w_full = 0, w_tempty = 0;
      for (int tile = blockIdx.x; walk_valid(tile); tile += gridDim.x, walk_next()) {
        int g, mt, nt;
        decode(g, mt, nt);
        if (tile_skipped(g, mt, nt)) continue;
        const int kb0 = op.k_begin(g) / kBK;
        const int halves = (kDual && op.m_begin(g) + mt * kBMT + TC_BM < op.m_end(g)) ? 2 : 1;
        w_tempty += ptx::mbar_wait<true>(tempty_bar(acc), acc_phase ^ 1u, timed);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(kDual ? 0 : acc * BN);
        const int nkb = sc.k_blocks - kb0;
        const int rot = (kStatA || Op::kRanged || (sc.dbg & 512)) ? 0 : (int)(blockIdx.x % (unsigned)nkb);   // same walk as the producer
        if constexpr (kStatA) {
          if (stat_g == -1 || (sc.a_batched && g != stat_g)) {   // a new resident A tile is on its way
            stat_g = g;
            w_full += ptx::mbar_wait<true>(afull_bar, afull_phase, timed);
            afull_phase ^= 1u;
            ptx::tc_fence_after();
          }
        }
        for (int ik = 0; ik < nkb; ++ik) {
          const int kb = kb0 + (ik + rot < nkb ? ik + rot : ik + rot - nkb);
          w_full += ptx::mbar_wait<true>(full_bar(stage), phase, timed);
          ptx::tc_fence_after();
          const int nk = (kb == sc.k_blocks - 1) ? sc.k16_last : kBK / E::kUmmaK;
          for (int k = 0; k < nk; ++k) {
            const uint64_t bd = make_smem_desc(b_smem(stage) + k * b_kstep, b_lbo, b_sbo, b_layout);
            for (int hf = 0; hf < halves; ++hf) {
              const uint64_t ad = make_smem_desc(a_smem(kStatA ? kb : stage, hf) + k * a_kstep, a_lbo, a_sbo, a_layout);
              if (!(sc.dbg & 16)) ptx::mma_elem<typename Op::InT>(d_tmem + (uint32_t)(hf * BN), ad, bd, idesc, (ik != 0 || k != 0) ? 1u : 0u);
            }
          }
          ptx::mma_commit(empty_bar(stage));  // frees the smem slot when these MMAs retire
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        ptx::mma_commit(tfull_bar(acc));  // accumulator complete -> epilogue
        if (++acc == kAccStages) { acc = 0; acc_phase ^= 1u; }
        if constexpr (kStatA) {
          // the next tile of this CTA belongs to another group and A is per group: tell the producer when the MMAs that
          // read the resident tile have retired
          if (sc.a_batched && sq + sq_step < sq_total && (sq + sq_step) / sc.n_tiles != g) ptx::mma_commit(afree_bar);
        }
      }
      if (sc.prof) {
        atomicAdd(sc.prof + 1, (unsigned long long)w_full);
        atomicAdd(sc.prof + 2, (unsigned long long)w_tempty);
      }
    }
  } else {
    // ===================== epilogue =====================
    const int ew = warp - 2;              // 0 .. tc_epi_warps(BN)-1
    const int quad = warp & 3;            // TMEM sub-partition this warp may read: lanes [32*quad, 32*quad+32)
    const int part = ew >> 2;             // BN/64 warps per sub-partition, each owns a 64-column slice
    constexpr int kColsPerWarp = TC_SLICE_COLS;
    static_assert(BN % TC_SLICE_COLS == 0, "BN must be a whole number of column slices");
    int acc = 0;
    uint32_t acc_phase = 0, res_phase = 0, sbuf = 0;
    long long w_tfull = 0, w_res = 0, c_pro = 0, c_loop = 0, c_tail = 0;
    const long long t_epi = timed ? clock64() : 0;
    for (int tile = blockIdx.x; walk_valid(tile); tile += gridDim.x, walk_next()) {
      int g, mt, nt;
      decode(g, mt, nt);
      if (tile_skipped(g, mt, nt)) continue;
      const int halves = (kDual && op.m_begin(g) + mt * kBMT + TC_BM < op.m_end(g)) ? 2 : 1;
      for (int hf = 0; hf < halves; ++hf) {   // dual-M: the two 128-row halves of the tile are drained one after the other
      const bool first_half = hf == 0, last_half = hf == halves - 1;
      const long long t_tile = timed ? clock64() : 0;
      const int m_tile0 = op.m_begin(g) + mt * kBMT + hf * TC_BM;
      const int m = m_tile0 + quad * 32 + lane;
      const int n_base = nt * BN + part * kColsPerWarp;
      const bool row_ok = m < op.m_end(g);
      const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)((kDual ? hf : acc) * BN + part * kColsPerWarp);
      static_assert(Op::kColContig, "the tensor-core epilogue writes along the row: the column index must be the contiguous output index");
      // ---- drain: TMEM -> registers -> fused epilogue -> warp-private swizzled smem transpose -> coalesced 16-byte
      //      global stores (and, for residual / addend blocks, coalesced loads through the same staging rows) ----
      auto drain = [&](auto feat_c) {
        constexpr int F = decltype(feat_c)::value;   // compile-time feature mask, or < 0: tested at run time
        typename Op::Row row{};
        if (row_ok) row = op.template row_f<F>(g, m);
        using OutT = typename Op::OutT;
        constexpr int kEs = (int)sizeof(OutT);
        constexpr int kPassCols = TC_STAGE_PITCH / kEs;                 // 64 (bf16) or 32 (fp32) columns per pass
        constexpr int kPasses = (kColsPerWarp + kPassCols - 1) / kPassCols;
        constexpr bool kGuardRows = F < 0 || (F & F_POS) != 0;           // compute8 may dereference per-row pointers
        constexpr uint32_t kBufs = Op::kStagingBufs;
        uint32_t region = staging_base + ((uint32_t)ew * kBufs + sbuf) * TC_STAGING_PER_WARP;
        uint32_t my_row = region + (uint32_t)lane * TC_STAGE_PITCH;
        const uint32_t sw = (uint32_t)(lane & 7);
        const int n_end = op.n_store();
        // residual / addend blocks are staged by TMA: bf16 outputs cover the warp's 64-column slice with one staging
        // buffer; fp32 outputs need two 32-column passes, i.e. both buffers of a double-buffered staging area
        static_assert(kEs == 2 || kEs == 4, "outputs are bf16 or fp32");
        constexpr bool kResOk = (kEs == 2) || (kBufs == 2 && kPasses == 2);
        const bool has_res = kResOk && feat_on<F, F_RES>(op.has_res());
        const bool valid = row_ok && row.valid;
        // all 32 rows outside the live range, or the column slice beyond the last column: nothing to drain
        const bool warp_live = __any_sync(0xffffffffu, valid) && n_base < n_end;
        // epilogue I/O by TMA: lane 0 moves the warp's 32 staging rows as one box, or lanes 0..3 each move one 8-row x
        // 128-byte group (1024 B) of them
        const bool out32 = sc.out_box32 != 0, res32 = sc.res_box32 != 0;
        const int wrow0 = m_tile0 + quad * 32;               // first GEMM row of the warp
        const int grow0 = wrow0 + 8 * lane;                  // first GEMM row of this lane's 8-row group (lanes 0..3)
        const bool gissue = out32 ? (lane == 0 && wrow0 < op.m_end(g)) : (lane < 4 && grow0 < op.m_end(g));
        auto staging_free = [&]() {   // the TMA store that last read this staging buffer has finished reading it
          if (lane < 4) { if (kBufs == 2) ptx::bulk_wait_read1(); else ptx::bulk_wait_read(); }
          __syncwarp();
        };
        // residual / addend block: TMA load straight into the staging rows, issued BEFORE waiting for the accumulator
        // so that its latency hides behind the MMA of this tile
        if (has_res && warp_live) {
          if (kEs == 2) staging_free();
          else { if (lane < 4) ptx::bulk_wait_read(); __syncwarp(); }   // fp32: both buffers are about to be refilled
          const int res_passes = (kEs == 2 || n_end - (n_base + kPassCols) <= 0) ? 1 : 2;
          if (lane == 0) ptx::mbar_expect_tx(res_bar(ew), (uint32_t)res_passes * 4 * 1024);
          __syncwarp();
          if (res32 ? lane == 0 : lane < 4) {   // (rows outside the tensor are zero-filled and still count as bytes)
            for (int rp = 0; rp < res_passes; ++rp) {
              int c[5];
              op.res_coords(g, res32 ? wrow0 : grow0, n_base + rp * kPassCols, c);
              const uint32_t buf = staging_base + ((uint32_t)ew * kBufs + (kBufs == 2 ? (sbuf ^ (uint32_t)rp) : 0u)) * TC_STAGING_PER_WARP;
              ptx::tma_load_5d(buf + (res32 ? 0u : (uint32_t)lane * 1024u), &tma_res, res_bar(ew), c);
            }
          }
        }
        bool released = false;
        long long w0 = 0;
        if (first_half) {
          w0 = ptx::mbar_wait(tfull_bar(acc), acc_phase, timed);
          w_tfull += w0;
          ptx::tc_fence_after();
        }
        long long t_ph = 0;
        if (timed) { t_ph = clock64(); c_pro += t_ph - t_tile - w0; }
#pragma unroll 1
        for (int pass = 0; pass < (warp_live ? kPasses : 0); ++pass) {
          const int pn0 = n_base + pass * kPassCols;
          const int pcols = (kColsPerWarp - pass * kPassCols) < kPassCols ? (kColsPerWarp - pass * kPassCols) : kPassCols;
          int nvalid = n_end - pn0;
          nvalid = nvalid < pcols ? nvalid : pcols;
          if (nvalid <= 0) break;  // warp-uniform
          if (has_res && pass == 0) {   // one barrier phase covers every staged pass of the tile
            w_res += ptx::mbar_wait(res_bar(ew), res_phase, timed);
            res_phase ^= 1u;
          }
          // (without a residual the staging rows are first written by the drain loop: the wait for the previous
          //  TMA store sits behind the first TMEM load so that their latencies overlap)
          bool staging_checked = has_res;
          // 32 accumulator columns per iteration as four independent 8-column groups.  On full passes the groups
          // sit in ONE basic block, so ptxas interleaves their instruction streams: the instruction-level
          // parallelism that the few resident warps (3-4 per sub-partition) cannot supply.
          auto group8 = [&](const uint32_t (&r)[32], int q, int cofs) {
            float accv[8], resv[8], outv[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) { accv[i] = __uint_as_float(r[8 * q + i]); resv[i] = 0.0f; }
            if constexpr (kEs == 2) {
              const uint32_t slot = my_row + ((((uint32_t)cofs >> 3) ^ sw) << 4);
              if (has_res) unpack_bf16x8(lds128(slot), resv);
              op.template compute8<F>(row, pn0 + cofs, accv, resv, outv);
              sts128(slot, make_uint4(pack_bf16x2(outv[0], outv[1]), pack_bf16x2(outv[2], outv[3]),
                                      pack_bf16x2(outv[4], outv[5]), pack_bf16x2(outv[6], outv[7])));
            } else {
              const uint32_t c4 = (uint32_t)cofs >> 2;  // 16-byte chunk index (4 floats)
              const uint32_t slot0 = my_row + ((c4 ^ sw) << 4), slot1 = my_row + (((c4 + 1) ^ sw) << 4);
              if (has_res) {
                const uint4 u0 = lds128(slot0), u1 = lds128(slot1);
                resv[0] = __uint_as_float(u0.x); resv[1] = __uint_as_float(u0.y); resv[2] = __uint_as_float(u0.z); resv[3] = __uint_as_float(u0.w);
                resv[4] = __uint_as_float(u1.x); resv[5] = __uint_as_float(u1.y); resv[6] = __uint_as_float(u1.z); resv[7] = __uint_as_float(u1.w);
              }
              op.template compute8<F>(row, pn0 + cofs, accv, resv, outv);
              if (op.out_tf32()) {   // the tensor is an operand of a later tf32 MMA: round here, the MMA would truncate
#pragma unroll
                for (int i = 0; i < 8; ++i) outv[i] = tf32_rna(outv[i]);
              }
              sts128(slot0, make_uint4(__float_as_uint(outv[0]), __float_as_uint(outv[1]), __float_as_uint(outv[2]), __float_as_uint(outv[3])));
              sts128(slot1, make_uint4(__float_as_uint(outv[4]), __float_as_uint(outv[5]), __float_as_uint(outv[6]), __float_as_uint(outv[7])));
            }
          };
          const int nchunks = (nvalid + 31) >> 5;
          // Experiment (compile with -DSFNO_TC_LDTM_PAIR=1; NOT yet run on a GPU, off in the shipped library): both
          // 32-column TMEM loads of a bf16 pass are issued back to back and waited for once, so their latency is paid
          // once per pass instead of once per chunk and the accumulator stage goes back to the MMA warp before any
          // epilogue math of the tile.  Only for the ops whose drain loop keeps 64 accumulator registers without
          // spilling under the launch bound (TcLdtmPair<Op>: DFT, inverse Legendre, inverse DFT; the conv epilogues
          // spill 300-550 bytes with it, the BN = 256 kernels are capped at 96 registers).
          bool pair_done = false;
          if constexpr (SFNO_TC_LDTM_PAIR != 0 && TcLdtmPair<Op>::value) {
            if (nchunks == 2) {
              uint32_t r0[32], r1[32];
              ptx::tmem_ld32(t_row + (uint32_t)(pass * kPassCols), r0);
              ptx::tmem_ld32(t_row + (uint32_t)(pass * kPassCols + 32), r1);
              ptx::tmem_ld_wait();
              if (last_half && !released && (pass == kPasses - 1 || n_end - (pn0 + kPassCols) <= 0)) {
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(tempty_bar(acc));
                released = true;
              }
              if (!staging_checked) {
                staging_free();
                staging_checked = true;
              }
              if (!kGuardRows || valid) {
#pragma unroll
                for (int q = 0; q < 4; ++q) group8(r0, q, 8 * q);
                if (nvalid == 64) {
#pragma unroll
                  for (int q = 0; q < 4; ++q) group8(r1, q, 32 + 8 * q);
                } else {
#pragma unroll
                  for (int q = 0; q < 4; ++q)
                    if (32 + 8 * q < nvalid) group8(r1, q, 32 + 8 * q);
                }
              }
              pair_done = true;
            }
          }
#pragma unroll 1
          for (int ci = 0; ci < (pair_done ? 0 : nchunks); ++ci) {
            uint32_t r[32];
            ptx::tmem_ld32(t_row + (uint32_t)(pass * kPassCols + 32 * ci), r);
            ptx::tmem_ld_wait();
            // that was this warp's last read of the accumulator: hand the TMEM stage back to the MMA warp now, the
            // math / staging / store of the last columns need it no more
            if (last_half && !released && ci == nchunks - 1 && (pass == kPasses - 1 || n_end - (pn0 + kPassCols) <= 0)) {
              ptx::tc_fence_before();
              __syncwarp();
              if (lane == 0) ptx::mbar_arrive(tempty_bar(acc));
              released = true;
            }
            if (!staging_checked) {
              staging_free();
              staging_checked = true;
            }
            if (!kGuardRows || valid) {
              if (32 * ci + 32 <= nvalid) {
#pragma unroll
                for (int q = 0; q < 4; ++q) group8(r, q, 32 * ci + 8 * q);
              } else {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                  if (32 * ci + 8 * q < nvalid) group8(r, q, 32 * ci + 8 * q);
              }
            }
          }
          if (timed) { const long long t = clock64(); c_loop += t - t_ph; t_ph = t; }
          // staged rows -> global: one TMA store per 8-row group; rows / columns outside the tensor are clipped
          ptx::fence_proxy_async();
          __syncwarp();
          if (!(sc.dbg & 4)) {
            // a warp whose 32 rows straddle two planes of the output (Op::kSplitBox: the inverse DFT, whose planes hold
            // Kp = nlat rounded to 8 GEMM rows) falls back to four 8-row boxes through the second tensor map (TMA
            // stores do not take negative box origins, so a clipped second 32-row box is not an option)
            const bool split = Op::kSplitBox && out32 && op.box_straddles(wrow0);   // warp-uniform
            if (out32 && !split) {
              if (gissue) {
                int c[5];
                op.io_coords(g, wrow0, pn0, c);
                ptx::tma_store_5d(&tma_out, region, c);
              }
            } else if (lane < 4 && grow0 < op.m_end(g)) {
              int c[5];
              op.io_coords(g, grow0, pn0, c);
              ptx::tma_store_5d(out32 ? &tma_out8 : &tma_out, region + (uint32_t)lane * 1024u, c);
            }
          }
          if (lane < 4) ptx::bulk_commit();
          sbuf = (kBufs == 2) ? (sbuf ^ 1u) : 0u;
          region = staging_base + ((uint32_t)ew * kBufs + sbuf) * TC_STAGING_PER_WARP;
          my_row = region + (uint32_t)lane * TC_STAGE_PITCH;
        }
        // the accumulator has been drained: hand the TMEM stage back to the MMA warp before any further work
        if (last_half && !released) {   // (warps without live rows / columns never entered the drain loop)
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(tempty_bar(acc));
        }
        // fused InstanceNorm statistics: one partial per (row, N tile, column slice); no cross-warp synchronisation
        if (feat_on<F, F_STATS>(op.wants_stats()) && row_ok)
          op.finish(g, m, nt * (BN / TC_SLICE_COLS) + part, valid ? row.stat_s() : 0.0f, valid ? row.stat_q() : 0.0f);
        if (timed) c_tail += clock64() - t_ph;
      };
      if constexpr (!Op::kGeneral && Op::kFast0 == Op::kFast1) {
        drain(std::integral_constant<int, Op::kFast0>{});
      } else {
        const int feat = op.feat();
        if (feat == Op::kFast0) drain(std::integral_constant<int, Op::kFast0>{});
        else if (feat == Op::kFast1) drain(std::integral_constant<int, Op::kFast1>{});
        else drain(std::integral_constant<int, -1>{});
      }
      }  // halves
      if (++acc == kAccStages) { acc = 0; acc_phase ^= 1u; }
    }
    if (lane < 4) ptx::bulk_wait_read();  // outstanding TMA stores still read this CTA's shared memory
    if (sc.prof && ew == 0 && lane == 0) {
      atomicAdd(sc.prof + 3, (unsigned long long)w_tfull);
      atomicAdd(sc.prof + 4, (unsigned long long)w_res);
      atomicAdd(sc.prof + 6, (unsigned long long)(clock64() - t_epi - w_tfull - w_res));
      atomicAdd(sc.prof + 8, (unsigned long long)c_pro);
      atomicAdd(sc.prof + 9, (unsigned long long)c_loop);
      atomicAdd(sc.prof + 10, (unsigned long long)c_tail);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == TC_WARP_MMA) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ---- host side -----------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled get_encode_tiled();
int tc_num_sms();

template <class TIn>
inline int encode_operand(const TmaOperand& o, bool k_contig, int rows_box, CUtensorMap* map, const char* what) {
  using E = TcElem<TIn>;
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return fail(SFNO_ERR_CUDA, "%s: cuTensorMapEncodeTiled unavailable", what);
  cuuint64_t dims[4] = {o.dims[0], o.dims[1], o.dims[2], o.dims[3]};
  cuuint64_t strides[3] = {o.strides[0], o.strides[1], o.strides[2]};
  for (int i = 1; i < 3; ++i)
    if (dims[i + 1] == 1) strides[i] = strides[i - 1] * dims[i];  // extent-1 dimension: any valid multiple of 16
  cuuint32_t box[4] = {1, 1, 1, 1};
  if (k_contig) { box[0] = E::kBK; box[1] = (cuuint32_t)rows_box; }
  else { box[0] = E::kAtom; box[1] = E::kBK; }
  cuuint32_t estr[4] = {1, 1, 1, 1};
  // MN-major tf32 tiles must sit in shared memory with the 32-byte-chunk variant of the 128-byte swizzle
  const CUtensorMapSwizzle swz = (!k_contig && E::kBytes == 4) ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B;
  CUresult r = enc(map, E::kTmaType, 4, const_cast<void*>(o.base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(SFNO_ERR_CUDA, "%s: cuTensorMapEncodeTiled failed (%d): base=%p dims=(%llu,%llu,%llu,%llu) strides=(%llu,%llu,%llu) box=(%u,%u)",
                what, (int)r, o.base, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2],
                (unsigned long long)dims[3], (unsigned long long)strides[0], (unsigned long long)strides[1],
                (unsigned long long)strides[2], box[0], box[1]);
  return SFNO_OK;
}

inline int encode_io(const TmaIo& o, CUtensorMap* map, const char* what) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return fail(SFNO_ERR_CUDA, "%s: cuTensorMapEncodeTiled unavailable", what);
  cuuint64_t dims[5], strides[4];
  cuuint32_t box[5], estr[5] = {1, 1, 1, 1, 1};
  for (int i = 0; i < 5; ++i) dims[i] = o.dims[i];
  for (int i = 0; i < 4; ++i) strides[i] = o.strides[i];
  for (int i = 1; i < 4; ++i)
    if (dims[i + 1] == 1) strides[i] = strides[i - 1] * dims[i];  // extent-1 dimension: any valid multiple of 16
  box[0] = (cuuint32_t)(TC_STAGE_PITCH / o.es);
  for (int i = 0; i < 4; ++i) box[i + 1] = o.box_rows[i];
  CUresult r = enc(map, o.es == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<void*>(o.base),
                   dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(SFNO_ERR_CUDA, "%s: cuTensorMapEncodeTiled (epilogue I/O) failed (%d): base=%p dims=(%llu,%llu,%llu,%llu,%llu)", what, (int)r,
                o.base, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2],
                (unsigned long long)dims[3], (unsigned long long)dims[4]);
  return SFNO_OK;
}

inline bool tma_io_ok(const TmaIo& o) {
  if (!o.ok || !o.base || ((uintptr_t)o.base & 15) != 0) return false;
  uint32_t rows = 1;
  for (int i = 0; i < 4; ++i) {
    rows *= o.box_rows[i];
    if (o.dims[i + 1] > 1 && (o.strides[i] % 16 != 0 || o.strides[i] == 0 || o.strides[i] >= (1ull << 40))) return false;
  }
  for (int i = 0; i < 5; ++i)
    if (o.dims[i] == 0 || o.dims[i] > (1ull << 31)) return false;
  return rows == 8 || rows == 32;
}
inline int tma_io_rows(const TmaIo& o) { return (int)(o.box_rows[0] * o.box_rows[1] * o.box_rows[2] * o.box_rows[3]); }

inline bool tma_operand_ok(const TmaOperand& o) {
  if (((uintptr_t)o.base & 15) != 0) return false;
  if (o.strides[0] % 16 != 0 || o.strides[0] == 0) return false;
  for (int i = 1; i < 3; ++i)
    if (o.dims[i + 1] > 1 && (o.strides[i] % 16 != 0 || o.strides[i] == 0)) return false;
  for (int i = 0; i < 4; ++i)
    if (o.dims[i] == 0 || o.dims[i] > (1ull << 31)) return false;
  return true;
}

// shared memory of a stationary-A launch: resident A (k_blocks x 16 KB) + B ring + staging + barriers + alignment slack
constexpr int kTcMaxSmem = 226 * 1024;
template <class Op, int BN>
inline int tc_stat_smem_bytes(int k_blocks, int stages) {
  using S = TcSmem<BN, Op::kColContig, false, Op::kStagingBufs>;
  return k_blocks * S::kAHalfBytes + stages * S::kBBytes + S::kStagingBytes + S::kBarrierBytes + 1024;
}
// B stages that fit next to a resident A tile of k_blocks K blocks (0: stationary A does not fit / is not worth it)
template <class Op, int BN>
inline int tc_stat_stages(int k_blocks) {
  using S = TcSmem<BN, Op::kColContig, false, Op::kStagingBufs>;
  const int left = kTcMaxSmem - 1024 - S::kBarrierBytes - S::kStagingBytes - k_blocks * S::kAHalfBytes;
  const int st = left / S::kBBytes;
  return st >= 3 ? (st > 6 ? 6 : st) : 0;
}

template <class Op, bool kDual, bool kStatA = false>
int launch_gemm_tc_impl(const Op& op, cudaStream_t stream, const char* what, int stat_stages = 0) {
  using Tr = TcTraits<Op>;
  constexpr int BN = Tr::BN;
  static_assert(!kDual || 2 * BN <= 512, "dual-M needs both accumulators in TMEM");
  using S = TcSmem<BN, Op::kColContig, kDual, Op::kStagingBufs>;
  if (op.M <= 0 || op.N <= 0 || op.G <= 0) return SFNO_OK;
  TmaOperand a, b;
  Tr::operands(op, a, b);
  CUtensorMap ma, mb, mo, mr, mo8;
  using E = TcElem<typename Op::InT>;
  SFNO_TRY(encode_operand<typename Op::InT>(a, Op::A_KCONTIG, TC_BM, &ma, what));
  SFNO_TRY(encode_operand<typename Op::InT>(b, Op::B_KCONTIG, BN, &mb, what));
  TmaIo io_out, io_res;
  Tr::io(op, io_out, io_res);
  if (!tma_io_ok(io_out)) return fail(SFNO_ERR_UNSUPPORTED, "%s: output tensor is not expressible as TMA boxes (eligibility not checked?)", what);
  SFNO_TRY(encode_io(io_out, &mo, what));
  mo8 = mo;
  if (Op::kSplitBox && tma_io_rows(io_out) == 32) {   // 8-row boxes for the warps whose 32 rows straddle two planes
    TmaIo io8 = io_out;
    io8.box_rows[0] = 8;
    SFNO_TRY(encode_io(io8, &mo8, what));
  }
  if (Tr::has_residual(op)) {
    if (!tma_io_ok(io_res)) return fail(SFNO_ERR_UNSUPPORTED, "%s: residual tensor is not expressible as TMA boxes", what);
    SFNO_TRY(encode_io(io_res, &mr, what));
  } else {
    mr = mo;
  }
  TcSched sc;
  sc.m_tiles = ceil_div(op.M, kDual ? 2 * TC_BM : TC_BM);
  sc.n_tiles = ceil_div(op.N, BN);
  sc.groups = op.G;
  const int64_t tiles = (int64_t)sc.m_tiles * sc.n_tiles * sc.groups;
  if (tiles > (1ll << 30)) return fail(SFNO_ERR_UNSUPPORTED, "%s: too many tiles", what);
  sc.num_tiles = (int)tiles;
  sc.k_blocks = ceil_div(op.K, E::kBK);
  const int k_rem = op.K - (sc.k_blocks - 1) * E::kBK;
  sc.k16_last = ceil_div(k_rem, E::kUmmaK);
  sc.a_batched = a.batched ? 1 : 0;
  sc.b_batched = b.batched ? 1 : 0;
  sc.a_glo = a.group_lo;
  sc.b_glo = b.group_lo;
  sc.a_rep = a.replicas;
  sc.b_rep = b.replicas;
  sc.out_box32 = tma_io_rows(io_out) == 32;
  sc.res_box32 = Tr::has_residual(op) && tma_io_rows(io_res) == 32;
  sc.dbg = g_tc_debug.load(std::memory_order_relaxed);
  sc.prof = (sc.dbg & 128) ? tc_prof_buffer() : nullptr;
  sc.stat_kb = kStatA ? sc.k_blocks : 0;
  sc.stat_stages = kStatA ? stat_stages : 0;
  const int smem_bytes = kStatA ? tc_stat_smem_bytes<Op, BN>(sc.k_blocks, stat_stages) : S::kTotal;
  static std::atomic<int> attr_bytes{0};   // per-process high-water mark of the opt-in (the attribute is per function)
  auto kern = gemm_tc_kernel<Op, BN, kDual, kStatA>;
  if (attr_bytes.load(std::memory_order_acquire) < smem_bytes) {
    SFNO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kStatA ? kTcMaxSmem : smem_bytes));
    attr_bytes.store(kStatA ? kTcMaxSmem : smem_bytes, std::memory_order_release);
  }
  int grid = std::min(sc.num_tiles, tc_num_sms());
  if (kStatA) grid = (tc_num_sms() / sc.m_tiles) * sc.m_tiles;   // every CTA owns one M tile: a multiple of m_tiles
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)tc_threads(BN));
  cfg.dynamicSmemBytes = (size_t)smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // overlap this kernel's set-up with its predecessor's tail
  attr[0].val.programmaticStreamSerializationAllowed = (sc.dbg & 1024) ? 0 : 1;   // tc_debug bit10: plain stream order
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  SFNO_CUDA(cudaLaunchKernelEx(&cfg, kern, ma, mb, mo, mr, mo8, op, sc));
  return post_launch(what);
}

// dual-M (two 128-row A tiles per B tile, single-buffered accumulators) for the ops whose traits ask for it:
// less L2->SMEM operand traffic per output element; tc_debug bit8 forces the single-tile kernel (A/B comparison)
template <class Op>
int launch_gemm_tc(const Op& op, cudaStream_t stream, const char* what) {
  if constexpr (TcTraits<Op>::kStationaryA) {
    // stationary A when the whole K extent of an A tile fits next to a >= 3-stage B ring and the SMs split evenly over
    // the M tiles; tc_debug bit 12 forces the streaming kernel (A/B comparison)
    const int kb = ceil_div(op.K, TcElem<typename Op::InT>::kBK), m_tiles = ceil_div(op.M, TC_BM);
    const int st = tc_stat_stages<Op, TcTraits<Op>::BN>(kb);
    if (st > 0 && m_tiles <= tc_num_sms() / 8 && TcTraits<Op>::use_stationary(op) && !(g_tc_debug.load(std::memory_order_relaxed) & 4096))
      return launch_gemm_tc_impl<Op, false, true>(op, stream, what, st);
  }
  if constexpr (TcTraits<Op>::kDualM) {
    if (TcTraits<Op>::use_dual(op) && !(g_tc_debug.load(std::memory_order_relaxed) & 256))
      return launch_gemm_tc_impl<Op, true>(op, stream, what);
  }
  return launch_gemm_tc_impl<Op, false>(op, stream, what);
}

}  // namespace sfno
