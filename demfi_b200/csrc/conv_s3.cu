// tcgen05 implicit-GEMM convolution, third generation ("s3"): the 3xFP16 scheme of conv_h3.cu
//
//   D_main += Ah * Bh            D_corr += Ah * Bl + Al * Bh          out = D_main + D_corr / 2048
//
// with BOTH operands read by the tensor core from shared memory (SS-form tcgen05.mma): no per-tap register
// traffic at all.  ncu on conv_h3 (profiles/r1_conv_s3_profile.md) showed the L1/shared-memory data pipe at
// ~80 % and the MMA issuer waiting for its A operand 53 % of the time: every tap of every chunk went
// shared memory -> registers -> tensor memory through eight splitter warps, and the epilogue's per-thread
// 16-byte global stores (one 32-byte sector each) were another 23 % of the pipe's wavefronts.
//
// Data flow per persistent CTA (one per SM), tile = 16 rows x 8 output pixels (M = 128), N block <= 96:
//   * activations: ONE TMA box per 32-channel chunk covering the tile plus its halo {32 ch, 8+KW-1, 16+KH-1},
//     128-byte rows = pixels, 128-byte swizzle.  A source in the S16 storage format (include/demfi_b200.h) already
//     holds [32 x fp16 hi | 32 x fp16 lo] per row: the TMA tile IS the MMA operand.  An fp32 source is split ONCE per
//     halo pixel, in place, by six converter warps into the same row format (same 128 bytes, same swizzle).
//     The A operand of tap (ky, kx) is the SAME buffer behind a shifted-window descriptor: start address moved by
//     (ky * halo_w + kx) * 128 bytes, SBO = halo_w * 128 (next tile row), k-step + 32 bytes, lo half + 64 bytes.  The
//     128-byte swizzle is a function of the absolute shared-memory address bits, so unaligned window starts need no
//     base offset (verified against float64 conv2d).  The tile is 8 pixels wide precisely so that every 8-row
//     core-matrix group of the M dimension is one contiguous run of halo pixels.
//   * weights: the pre-packed [Bh rows ; Bl rows] x 32 fp16 tiles of conv_h3 (64-byte swizzle).  When the whole
//     filter bank of the N block fits beside the activation buffers (e.g. 64 -> 64 3x3: 144 KB) it is loaded once per
//     CTA and stays resident; otherwise it streams through a ring whose slots hold groups of consecutive (chunk, tap)
//     stages (one bulk copy and one barrier round trip per group).
//   * MMAs per (tap, 16 channels): Ah x [Bh;Bl] (N' = 2N: main | corr) and Al x Bh (-> corr), issued by one thread
//     whose inner loop is 4 MMAs + 32-bit adds (every instruction there is serial with the tensor pipe).
//   * the K loop is cut into segments whose partial sums are drained from tensor memory (double-buffered) and
//     added in fp32 RN by eight epilogue warps, with the gain compensation of the truncating tensor-core
//     accumulation.
//   * epilogue, planned per N block on the host: bias, operand tiles (residual, GRU h / z) TMA-loaded into the staging
//     tile, activation, result staged in the TMA 128-byte-swizzle box layout as fp32 or S16 and written with
//     cp.async.bulk.tensor stores (1-2 destinations, pixel shuffle as a strided tensor map): whole 128-byte lines
//     instead of one sector per thread.  What the planner cannot express keeps the generic per-thread epilogue of
//     common.cuh.
#include <cuda.h>
#include <cuda_fp16.h>

#include <cstring>
#include <mutex>
#include <set>
#include <utility>

#include "common.cuh"

namespace demfi {

constexpr int S3_TH = 16, S3_TW = 8, S3_BM = 128, S3_KC = 32;
constexpr int S3_CV_WARPS = 6, S3_EPI_WARPS = 8;
constexpr int S3_CV_THREADS = S3_CV_WARPS * 32;
constexpr int S3_EPI_THREADS = S3_EPI_WARPS * 32;
constexpr int S3_THREADS = (S3_CV_WARPS + S3_EPI_WARPS + 2) * 32;  // 512: 0-5 convert, 6-13 epilogue, 14 TMA, 15 MMA
constexpr int S3_MAX_NA = 6;   // halo-tile buffers (3 in general; up to 6 for 1x1 / few-tap kernels, which are HBM-latency bound)
constexpr int S3_MAX_NS = 8;   // weight ring
constexpr int S3_BOX_BYTES = S3_BM * 128;  // one 32-channel staging box
constexpr int S3_BAR_PEER = 3 * S3_MAX_NA + 4 + 2 + 2 * S3_MAX_NS;      // CTA pair: peerA[MAX_NA], peerW, peerB[MAX_NS] (leader's copies)
constexpr int S3_NBARS = S3_BAR_PEER + S3_MAX_NA + 1 + S3_MAX_NS + 1;  // (+ the second operand-tile barrier, the last one)
constexpr int S3_MAX_E = 8;    // epilogue plan entries (N blocks x sub-blocks)
constexpr int S3_MAX_O = 8;    // destination tensor maps
constexpr int S3_MAX_OL = 16;  // (entry, destination) pairs
constexpr float S3_LO_SCALE = 2048.0f;

struct S3Params {
  CUtensorMap tmap[DEMFI_MAX_SRC];  // activation sources
  // TMA epilogue, planned on the host per ENTRY = (N block, sub-block).  A sub-block is the whole N block when every segment
  // that intersects the block is alike (one result, 1-2 destinations: the common case), else 32 channels = one staging box
  // (multi-head convolutions whose heads differ in activation / format / operands inside one MMA tile: GRU z | r, the
  // dense-block "push" convolutions).  Entry e = nb * nsb + sb.
  CUtensorMap omap[S3_MAX_O];         // destinations (one per segment; one per quadrant for a pixel-shuffle segment)
  CUtensorMap rmap[DEMFI_MAX_SEG];    // first operand (residual / GRU h) of a segment
  CUtensorMap r2map[DEMFI_MAX_SEG];   // second operand (GRU z) of a segment
  int sbw, nsb;                       // sub-block width in channels, sub-blocks per full N block
  signed char e_seg[S3_MAX_E];        // representative segment of the entry (-1: nothing to store)
  signed char e_nres[S3_MAX_E];       // operands to fetch (0, 1, 2)
  signed char e_o0[S3_MAX_E], e_on[S3_MAX_E];  // destinations: ol_*[e_o0 .. e_o0 + e_on)
  int e_roff[S3_MAX_E];               // first operand: 0 = fetched into the result tile (same format), else into the second tile
  int e_rc0[S3_MAX_E];                // channel of the entry's first channel inside the operand tensors
  int e_info[S3_MAX_E];               // packed for the store phase: seg.fmt | act << 3 | nres << 6 | (operand in the second tile) << 8
  signed char ol_map[S3_MAX_OL];      // destination list: tensor map ...
  int ol_c0[S3_MAX_OL];               // ... and the channel coordinate of the entry's first channel in it
  demfi_conv_t c;
  int tiles_x, tiles_y, ntiles, n_blocks, nb_max;
  int hw, hh, halo_px;
  int a_bytes, na;
  int b_bytes, ns, resident;
  int gtaps;  // weight ring: (chunk, tap) stages per ring slot, fetched with ONE bulk copy and ONE barrier round trip
  int b_off, stg_off, bar_off, bias_off;  // byte offsets in (1024-aligned) shared memory
  int acc_stride;
  int taps, stages_per_tile, flush;
  int seg_units, nseg, seg_last, grp_units, ngrp, grp_last;  // per tile: accumulation segments / weight-ring groups, in issue units
  int unit, ustep;  // MMA issue unit: taps per unit (kernel row / column / single tap), tap-to-tap step of the A descriptor (16-byte units)
  int tma_epi, stg2_off;
  int all_full_chunks;  // every source has C % 32 == 0 (no ragged chunk)
  int all_s16;          // every source is in the S16 format: the converter warps have nothing to do
  int pair;             // CTA-pair kernel (DEMFI_CONV_TC16P)
  int offload;          // the TMA duties of the epilogue (operand fetch, stores) run on a warp of their own (all sources S16)
  int lean;             // lean epilogue: bit 0 = eligible, bit 1 = ReLU, bit 2 = one S16 operand; bit 3 = the general variant (LEAN == 2),
                        // with bit 4 = fp32 destination, bit 5 = fp32 first operand
  int pf;               // L2 prefetch of the activation chunks this many tiles ahead (0: none)
  int res_sep;          // 1: the operand tile of a lean layer has a tile of its own (fetched one tile ahead by the store warp);
                        // 2: two tiles used in turn, each holding a tile's operand and then, in place, its result (operand fetched TWO tiles ahead)
  float comp;
  int diag;
  long long* dbg;
};

namespace s3 {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
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
// non-blocking probe (try_wait may suspend the thread for a system-dependent time: wrong tool for polling two queues)
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug becomes a trapped launch (reported error), never a hung GPU.
static __device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity) {
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 8000000000ll) {
      printf("demfi conv_s3: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", (int)blockIdx.x,
             (int)threadIdx.x, bar, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (!mbar_try_wait(bar, parity)) mbar_wait_slow(bar, parity);
}
// Same bound, no call: a function call in the MMA issuer's loop makes ptxas keep the loop state in vector registers across
// the call site, and every tcgen05.mma then needs its operands moved into uniform registers (R2UR) -- see s3_issue.
__device__ __forceinline__ void mbar_wait_nocall(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 8000000000ll) __trap();
  }
}
__device__ __forceinline__ void mbar_wait_t(uint32_t bar, uint32_t parity, bool on, long long& acc) {
  if (!on) { mbar_wait(bar, parity); return; }
  const long long t = clock64();
  mbar_wait(bar, parity);
  acc += clock64() - t;
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// L2 prefetch of a box (no shared-memory destination, no barrier): the later tma_load_4d of the same box hits L2
__device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap* map, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(map),
               "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, px;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// both operands from shared memory
__device__ __forceinline__ void umma_f16_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
// descriptors passed as (low, high) 32-bit halves: only the low word (start address) ever changes
__device__ __forceinline__ void umma_f16_ss2(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                             uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accum)
      : "memory");
}
// ---- CTA pair (cta_group::2): one tcgen05.mma covers M = 256 = the 128-pixel tiles of BOTH CTAs of a cluster; every CTA
// supplies its own A rows and HALF of the B rows from the same shared-memory offsets; issued by the leader CTA only.
__device__ __forceinline__ void umma_f16_ss2_pair(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                                  uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accum)
      : "memory");
}
// completion of all earlier MMAs -> the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// Arrive on the barrier at the same offset in CTA `rank` of the cluster.  Default semantics (release at CTA scope), as the
// CUTLASS 2-SM kernels do: what is handed over is operand data written by TMA and tensor-memory reads ordered by tcgen05
// fences, not generic-proxy stores.  (.release.cluster compiles to MEMBAR.ALL.GPU + ERRBAR + CCTL.IVALL around the arrive
// and .acquire.cluster polling to an L1 invalidate per wait: ncu showed 10 % of all warp samples in those, and the hand-over
// of an accumulator from the peer's epilogue to the leader's issuer took thousands of cycles.)
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t rank) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(bar), "r"(rank));
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}
// wait for arrivals from the peer CTA: the ordinary CTA-scope wait (see mbar_arrive_remote)
__device__ __forceinline__ void mbar_wait_cluster_nocall(uint32_t bar, uint32_t parity) { mbar_wait_nocall(bar, parity); }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// K-major shared-memory matrix descriptors.  Bits: start >> 4 in [0,14), LBO >> 4 in [16,30), SBO >> 4 in [32,46),
// version 1 in [46,48), layout in [61,64).
// B (weights): rows of 32 fp16 = 64 bytes, 64-byte swizzle (layout 4), SBO = 512 B between 8-row groups.
__device__ __forceinline__ uint64_t make_desc_sw64(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (4ull << 61);
}
// A (activations): rows of 128 bytes = one halo pixel (Ah | Al), 128-byte swizzle (layout 2): the 16-byte chunks of the
// row at byte address a are XOR-ed with bits [7,10) of a -- the pattern TMA wrote and the converter kept.  SBO =
// distance between 8-row groups = one halo row.
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) |
         (2ull << 61);
}
// a = h + l / 2048 for two values: returns the packed fp16 pairs (low half = first value).  Both conversions saturate
// (cvt.rn.satfinite): a value beyond the fp16 range (|a| > 65504) becomes +-65504 + 32-ish instead of inf -- the network's
// activations are O(1..100), but an inf here would turn into NaN in the next layer's MMAs (inf - inf in the lo term).
__device__ __forceinline__ uint32_t cvt_f16x2_sat(float lo_val, float hi_val) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_val), "f"(lo_val));
  return r;
}
__device__ __forceinline__ void split2(float a0, float a1, uint32_t& hi, uint32_t& lo) {
  hi = cvt_f16x2_sat(a0, a1);
  const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hi));
  lo = cvt_f16x2_sat((a0 - hf.x) * S3_LO_SCALE, (a1 - hf.y) * S3_LO_SCALE);
}
// Transcendental epilogues (tanh / sigmoid heads, GRU gates), out of line: ONE copy of the exp / reciprocal expansions
// instead of one per unrolled call site -- inlined they were a quarter of the kernel's 7 K instructions and the
// instruction-cache pressure slowed the MMA issuer's loop by 10 % (A/B on one GPU).
// sigmoid / tanh through ex2.approx + the approximate reciprocal: absolute error ~1e-7 on outputs in [-1, 1] (the parity
// budget is 5e-4 end to end, 2e-5 per operator in the tests); the IEEE division + expf versions made the GRU convolutions
// epilogue-bound (0.69 ms in the network against 0.48 ms for the same shape with a ReLU epilogue).
__device__ __forceinline__ float sigmoid_fast(float v) { return __fdividef(1.0f, 1.0f + __expf(-v)); }
__device__ __forceinline__ float tanh_fast(float v) { return 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * v)); }
// v = accumulator + bias; h = first operand (residual / GRU h; zeros when absent); z = second operand (GRU z)
__device__ __forceinline__ float4 finish4_inl(int act, float4 v, float4 h, float4 z) {
  if (act == DEMFI_ACT_SIGMOID_MUL) {  // r * h (DeMFInet.py:846-847)
    v.x = sigmoid_fast(v.x) * h.x; v.y = sigmoid_fast(v.y) * h.y; v.z = sigmoid_fast(v.z) * h.z; v.w = sigmoid_fast(v.w) * h.w;
  } else if (act == DEMFI_ACT_GRU) {  // (1 - z) h + z tanh(q) (DeMFInet.py:847-848)
    v.x = (1.0f - z.x) * h.x + z.x * tanh_fast(v.x);
    v.y = (1.0f - z.y) * h.y + z.y * tanh_fast(v.y);
    v.z = (1.0f - z.z) * h.z + z.z * tanh_fast(v.z);
    v.w = (1.0f - z.w) * h.w + z.w * tanh_fast(v.w);
  } else {
    v.x += h.x; v.y += h.y; v.z += h.z; v.w += h.w;
    if (act == DEMFI_ACT_RELU) {
      v.x = fmaxf(v.x, 0.0f); v.y = fmaxf(v.y, 0.0f); v.z = fmaxf(v.z, 0.0f); v.w = fmaxf(v.w, 0.0f);
    } else if (act == DEMFI_ACT_TANH) {
      v.x = tanh_fast(v.x); v.y = tanh_fast(v.y); v.z = tanh_fast(v.z); v.w = tanh_fast(v.w);
    } else if (act == DEMFI_ACT_SIGMOID) {
      v.x = sigmoid_fast(v.x); v.y = sigmoid_fast(v.y); v.z = sigmoid_fast(v.z); v.w = sigmoid_fast(v.w);
    }
  }
  return v;
}
static __device__ __noinline__ float4 finish4v(int act, float4 v, float4 h, float4 z) { return finish4_inl(act, v, h, z); }
__device__ __forceinline__ float4 as_f4(uint4 u) {
  return make_float4(__uint_as_float(u.x), __uint_as_float(u.y), __uint_as_float(u.z), __uint_as_float(u.w));
}
__device__ __forceinline__ uint4 as_u4(float4 v) {
  return make_uint4(__float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z), __float_as_uint(v.w));
}
// S16 decode: eight channels = one 16-byte chunk of fp16 hi + one of fp16 lo -> two float4 (value = hi + lo / 2048)
__device__ __forceinline__ void s16_decode8(uint4 hi, uint4 lo, float4& v0, float4& v1) {
  const float k = 1.0f / S3_LO_SCALE;
  float2 h, l;
  h = __half22float2(*reinterpret_cast<const __half2*>(&hi.x)); l = __half22float2(*reinterpret_cast<const __half2*>(&lo.x));
  v0.x = fmaf(l.x, k, h.x); v0.y = fmaf(l.y, k, h.y);
  h = __half22float2(*reinterpret_cast<const __half2*>(&hi.y)); l = __half22float2(*reinterpret_cast<const __half2*>(&lo.y));
  v0.z = fmaf(l.x, k, h.x); v0.w = fmaf(l.y, k, h.y);
  h = __half22float2(*reinterpret_cast<const __half2*>(&hi.z)); l = __half22float2(*reinterpret_cast<const __half2*>(&lo.z));
  v1.x = fmaf(l.x, k, h.x); v1.y = fmaf(l.y, k, h.y);
  h = __half22float2(*reinterpret_cast<const __half2*>(&hi.w)); l = __half22float2(*reinterpret_cast<const __half2*>(&lo.w));
  v1.z = fmaf(l.x, k, h.x); v1.w = fmaf(l.y, k, h.y);
}
// Operand fetch from a staged tile (box row `base` of pixel m, sw = m & 7), channels chn..chn+7 / chn..chn+3 of the N block
__device__ __forceinline__ void op_load8(uint32_t base, int chn, uint32_t sw, bool s16, float4& v0, float4& v1) {
  if (s16) {
    const uint32_t g8 = (uint32_t)((chn & 31) >> 3);
    s16_decode8(lds128(base + ((g8 ^ sw) << 4)), lds128(base + (((g8 + 4u) ^ sw) << 4)), v0, v1);
  } else {
    const uint32_t q = (uint32_t)((chn & 31) >> 2);
    v0 = as_f4(lds128(base + ((q ^ sw) << 4)));
    v1 = as_f4(lds128(base + (((q + 1u) ^ sw) << 4)));
  }
}
__device__ __forceinline__ float4 op_load4(uint32_t base, int chn, uint32_t sw, bool s16) {
  if (s16) {  // four channels = half of an S16 chunk pair
    const uint32_t g8 = (uint32_t)((chn & 31) >> 3);
    const uint32_t half = (uint32_t)(chn & 4) << 1;  // byte offset 0 or 8 inside the 16-byte chunks
    uint2 uh, ul;
    asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(uh.x), "=r"(uh.y) : "r"(base + ((g8 ^ sw) << 4) + half));
    asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(ul.x), "=r"(ul.y) : "r"(base + (((g8 + 4u) ^ sw) << 4) + half));
    float4 v, dummy;
    s16_decode8(make_uint4(uh.x, uh.y, 0u, 0u), make_uint4(ul.x, ul.y, 0u, 0u), v, dummy);
    return v;
  }
  return as_f4(lds128(base + (((uint32_t)((chn & 31) >> 2) ^ sw) << 4)));
}
__device__ __forceinline__ void s16_encode8(float4 v0, float4 v1, uint4& hi, uint4& lo) {
  split2(v0.x, v0.y, hi.x, lo.x);
  split2(v0.z, v0.w, hi.y, lo.y);
  split2(v1.x, v1.y, hi.z, lo.z);
  split2(v1.z, v1.w, hi.w, lo.w);
}
}  // namespace s3
using namespace s3;

// One unit of issue = U consecutive taps (a kernel row; a kernel column for Nx1 kernels; one tap for U = 1), straight-line.
// a / b: low descriptor words of the unit's first tap; ustep / bstep: their steps from tap to tap.
// PAIR: cta_group::2 MMAs (M = 256 over the CTA pair; the corrections have their own accumulator columns, see the kernel).
template <int U, bool K2, bool PAIR>
__device__ __forceinline__ void s3_issue_unit(uint32_t d_main, uint32_t d_corr, uint32_t a, uint32_t a_hi, uint32_t ustep, uint32_t b,
                                              uint32_t b_hi, uint32_t bstep, uint32_t idesc_2n, uint32_t idesc_n, uint32_t accum) {
#pragma unroll
  for (int t = 0; t < U; ++t) {
    const uint32_t at = a + (uint32_t)t * ustep, bt = b + (uint32_t)t * bstep;
    if (PAIR) {
      umma_f16_ss2_pair(d_main, at, a_hi, bt, b_hi, idesc_2n, t == 0 ? accum : 1u);
      if (K2) umma_f16_ss2_pair(d_main, at + 2u, a_hi, bt + 2u, b_hi, idesc_2n, 1u);
      umma_f16_ss2_pair(d_corr, at + 4u, a_hi, bt, b_hi, idesc_n, t == 0 ? accum : 1u);  // own columns: first MMA of a segment overwrites
      if (K2) umma_f16_ss2_pair(d_corr, at + 6u, a_hi, bt + 2u, b_hi, idesc_n, 1u);
    } else {
      umma_f16_ss2(d_main, at, a_hi, bt, b_hi, idesc_2n, t == 0 ? accum : 1u);  // Ah x [Bh;Bl]  k 0..15
      if (K2) umma_f16_ss2(d_main, at + 2u, a_hi, bt + 2u, b_hi, idesc_2n, 1u);  //               k 16..31
      umma_f16_ss2(d_corr, at + 4u, a_hi, bt, b_hi, idesc_n, 1u);               // Al x Bh
      if (K2) umma_f16_ss2(d_corr, at + 6u, a_hi, bt + 2u, b_hi, idesc_n, 1u);
    }
  }
}

template <bool DBG>
__device__ __forceinline__ void mbar_wait_i(uint32_t bar, uint32_t parity, bool skip, long long& acc) {
  if (skip) return;  // diagnostics (tc_diag & 1024): the bare issue loop
  if (!DBG) { mbar_wait_nocall(bar, parity); return; }
  const long long t = clock64();
  mbar_wait_nocall(bar, parity);
  acc += clock64() - t;
}
template <bool DBG>
__device__ __forceinline__ void mbar_wait_ic(uint32_t bar, uint32_t parity, long long& acc) {  // arrivals from the peer CTA
  if (!DBG) { mbar_wait_cluster_nocall(bar, parity); return; }
  const long long t = clock64();
  mbar_wait_cluster_nocall(bar, parity);
  acc += clock64() - t;
}

// The persistent issue loop of one CTA (single thread).  Units are numbered through the tiles of the CTA; segment (accumulator
// hand-over), weight-ring group and chunk boundaries all fall on unit boundaries (s3_plan).  Barrier layout as in the kernel.
// Written for ptxas' uniform datapath: no calls, no min / max (vector-only instructions), counters that count down to a
// reload value chosen by a select -- the SASS of the loop is UTCHMMA / UTCBAR / U* instructions plus the barrier waits.
//
// MODE 0: one CTA.  MODE 1: leader of a CTA pair -- issues cta_group::2 MMAs for both CTAs, so besides its own barriers it
// waits for the peer's operands (peerA / peerB / peerW, forwarded by the peer's shadow thread) and for both CTAs' epilogues
// (the leader's tempty counts the warps of both); its commits are multicast to the barriers of both CTAs.  MODE 2: the
// shadow -- the same thread of the OTHER CTA walks the same sequence of boundaries, issues nothing, and wherever the leader
// would wait for an operand it waits for the LOCAL copy of that barrier and arrives on the leader's peer barrier.
template <int U, bool DBG, int MODE>
__device__ __forceinline__ void s3_issue(const S3Params& P, uint32_t smem_base, uint32_t tmem_base, uint32_t bars, long long& w_tempty,
                                         long long& w_ready, long long& w_peer) {
  constexpr bool PAIR = MODE != 0, LEAD = MODE == 1, SHADOW = MODE == 2;
  const demfi_conv_t& c = P.c;
  auto bar_rawfull = [&](uint32_t a) { return bars + 8u * a; };
  auto bar_cvfull = [&](uint32_t a) { return bars + 8u * ((uint32_t)S3_MAX_NA + a); };
  auto bar_aempty = [&](uint32_t a) { return bars + 8u * ((uint32_t)(2 * S3_MAX_NA) + a); };
  auto bar_tfull = [&](uint32_t a) { return bars + 8u * ((uint32_t)(3 * S3_MAX_NA) + a); };
  auto bar_tempty = [&](uint32_t a) { return bars + 8u * ((uint32_t)(3 * S3_MAX_NA + 2) + a); };
  const uint32_t bar_wfull = bars + 8u * (uint32_t)(3 * S3_MAX_NA + 4);
  auto bar_bfull = [&](uint32_t sl) { return bars + 8u * ((uint32_t)(3 * S3_MAX_NA + 6) + sl); };
  auto bar_bfree = [&](uint32_t sl) { return bars + 8u * ((uint32_t)(3 * S3_MAX_NA + 6 + S3_MAX_NS) + sl); };
  auto bar_peerA = [&](uint32_t a) { return bars + 8u * ((uint32_t)S3_BAR_PEER + a); };
  const uint32_t bar_peerW = bars + 8u * (uint32_t)(S3_BAR_PEER + S3_MAX_NA);
  auto bar_peerB = [&](uint32_t sl) { return bars + 8u * ((uint32_t)(S3_BAR_PEER + S3_MAX_NA + 1) + sl); };
  const bool bare = !PAIR && (P.diag & 1024) != 0;
  const uint32_t NA = (uint32_t)P.na, NS = (uint32_t)P.ns;
  const uint32_t idesc0 = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)((PAIR ? 2 * S3_BM : S3_BM) >> 4) << 24);  // D=f32, A=B=f16, K-major
  const uint32_t a_hi = (uint32_t)(make_desc_sw128(0u, (uint32_t)P.hw * 128u) >> 32);
  const uint32_t b_hi = (uint32_t)(make_desc_sw64(0u) >> 32);
  const uint32_t a_lo0 = ((smem_base >> 4) & 0x3FFFu) | (1u << 16);
  const uint32_t b_lo0 = ((smem_base + (uint32_t)P.b_off) >> 4) & 0x3FFFu;
  const uint32_t astep = (uint32_t)(P.a_bytes >> 4), gstep = (uint32_t)((P.gtaps * P.b_bytes) >> 4);
  const uint32_t ustep = (uint32_t)P.ustep;        // tap to tap inside a unit
  const uint32_t rstep = (uint32_t)P.hw << 3;      // unit to unit inside a chunk (next kernel row); U = 1: handled by kx
  const int KW = c.KW;
  const bool resident = P.resident != 0;
  const int upc = P.taps / U;                       // units per chunk
  const int chunks_per_tile = P.stages_per_tile / P.taps;
  // tiles of this CTA (MODE 0) / tile PAIRS of this cluster (the two CTAs take tiles 2q and 2q + 1)
  const int first = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int stride = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int count = PAIR ? (P.ntiles + 1) >> 1 : P.ntiles;
  if (first >= count) return;

  uint32_t slot = 0, sphase = 0, acc = 0, acc_phase = 0, abuf = 0, aphase = 0;
  // An activation chunk is ready: straight from TMA when EVERY source is S16 (the converter warps are idle), else after the
  // converter warps, which then see every chunk -- they convert the fp32 ones and merely pass the S16 ones on.  (Waiting on
  // the TMA barrier for the S16 chunks of a mixed convolution left the converter warps outside the flow control: lapped by
  // producer + issuer by two phases of a buffer, their parity wait never returned -- a hang seen at 1280x736 only.)
  const bool all_s16 = P.all_s16 != 0;
  auto wait_chunk = [&](uint32_t buf, uint32_t phase) {
    mbar_wait_i<DBG>(all_s16 ? bar_rawfull(buf) : bar_cvfull(buf), phase, bare, w_ready);
    if (LEAD) mbar_wait_ic<DBG>(bar_peerA(buf), phase, w_peer);
    if (SHADOW) mbar_arrive_remote(bar_peerA(buf), 0u);
  };
  auto wait_group = [&](uint32_t sl, uint32_t phase) {
    mbar_wait_i<DBG>(bar_bfull(sl), phase, bare, w_ready);
    if (LEAD) mbar_wait_ic<DBG>(bar_peerB(sl), phase, w_peer);
    if (SHADOW) mbar_arrive_remote(bar_peerB(sl), 0u);
  };
  auto wait_acc = [&](uint32_t a_, uint32_t phase) {  // both CTAs' epilogue warps arrive on the leader's barrier
    if (SHADOW) return;
    if (LEAD) mbar_wait_ic<DBG>(bar_tempty(a_), phase, w_tempty);
    else mbar_wait_i<DBG>(bar_tempty(a_), phase, bare, w_tempty);
  };
  auto commit = [&](uint32_t bar) {
    if (SHADOW || bare) return;
    if (LEAD) umma_commit_pair(bar);
    else umma_commit(bar);
  };
  if (resident && !bare) {
    mbar_wait_nocall(bar_wfull, 0);
    if (LEAD) mbar_wait_cluster_nocall(bar_peerW, 0);
    if (SHADOW) mbar_arrive_remote(bar_peerW, 0u);
  }
  // the first unit's waits
  wait_chunk(0u, 0u);
  wait_acc(acc, acc_phase ^ 1u);
  if (!resident) wait_group(slot, sphase);
  tc_fence_after();

  // Structure (measured in tools/probe_s3*: the same MMAs cost 292 clk per stage with boundary checks after every unit and
  // 235 clk with none; the floor is 224): a RUN of units up to the next boundary is issued by an inner loop that contains
  // nothing but the units and two adds; the boundaries -- chunk (activation buffer hand-over), accumulation segment
  // (accumulator hand-over), weight-ring group -- are handled between runs: commits, bookkeeping, then the waits the next run
  // needs.  With resident weights and one segment per chunk (the 64 -> 64 3x3 ResBlock convolutions) a run is a whole chunk.
  for (int tile = first; tile < count; tile += stride) {
    const bool last_tile = tile + stride >= count;
    const int nb = PAIR ? 0 : tile % P.n_blocks;
    const int N = nb == P.n_blocks - 1 ? c.cout_pad - nb * P.nb_max : P.nb_max;
    const uint32_t idesc_n = idesc0 | ((uint32_t)(N >> 3) << 17);
    const uint32_t idesc_2n = idesc0 | ((uint32_t)((2 * N) >> 3) << 17);
    // one stage of weights in THIS CTA's shared memory, in 16-byte units: 2N rows x 64 bytes; a CTA of a pair holds half
    const uint32_t bstep = PAIR ? (uint32_t)N << 2 : (uint32_t)N << 3;
    const uint32_t bunit = bstep * (uint32_t)U;
    // countdowns (in units): to the end of the chunk / the accumulation segment / the weight-ring group
    int chunk_left = upc, chunks_left = chunks_per_tile;
    int seg_left = P.nseg == 1 ? P.seg_last : P.seg_units, segs_left = P.nseg;
    int grp_left = resident ? 0x40000000 : (P.ngrp == 1 ? P.grp_last : P.grp_units), grps_left = P.ngrp;
    int si = 0, c0 = 0;
    bool k2 = c.src[0].C > 16;  // channels 16..31 of the chunk exist (else they are TMA zero fill: skip their MMAs)
    uint32_t accum = 0u;
    uint32_t a = a_lo0 + astep * abuf;
    uint32_t b = resident ? b_lo0 : b_lo0 + gstep * slot;
    uint32_t d_main = tmem_base + acc * (uint32_t)P.acc_stride;
    int kx = 0;
#pragma unroll 1
    while (chunks_left > 0) {
      int run = chunk_left;
      if (seg_left < run) run = seg_left;
      if (grp_left < run) run = grp_left;
      // corrections: accumulated onto the Ah x Bl half of the main tile (one CTA); own columns after the main tile (pair)
      const uint32_t d_corr = d_main + (PAIR ? 2u * (uint32_t)N : (uint32_t)N);
      // ---- the run: MMAs and two adds per unit ----
      if (!SHADOW) {
        if (k2) {
#pragma unroll 1
          for (int r = 0; r < run; ++r) {
            s3_issue_unit<U, true, PAIR>(d_main, d_corr, a, a_hi, ustep, b, b_hi, bstep, idesc_2n, idesc_n, accum);
            accum = 1u;
            b += bunit;
            if (U == 1) { a += 8u; if (++kx == KW) { kx = 0; a += rstep - ((uint32_t)KW << 3); } }
            else a += rstep;
          }
        } else {
#pragma unroll 1
          for (int r = 0; r < run; ++r) {
            s3_issue_unit<U, false, PAIR>(d_main, d_corr, a, a_hi, ustep, b, b_hi, bstep, idesc_2n, idesc_n, accum);
            accum = 1u;
            b += bunit;
            if (U == 1) { a += 8u; if (++kx == KW) { kx = 0; a += rstep - ((uint32_t)KW << 3); } }
            else a += rstep;
          }
        }
      }
      // ---- boundaries: commits, bookkeeping, then the waits the next run needs.  (Waiting earlier -- before the last unit
      // of the run -- was measured and is slower: with two activation buffers the next chunk is still in flight then.) ----
      chunk_left -= run; seg_left -= run; grp_left -= run;
      const bool end_chunk = chunk_left == 0, end_seg = seg_left == 0, end_group = grp_left == 0;
      const bool more = !(end_chunk && chunks_left == 1 && last_tile);  // another unit follows in this CTA
      if (end_group) {
        commit(bar_bfree(slot));
        if (++slot == NS) { slot = 0; sphase ^= 1u; }
        if (--grps_left == 0) grps_left = P.ngrp;  // (next tile)
        grp_left = grps_left == 1 ? P.grp_last : P.grp_units;
        b = b_lo0 + gstep * slot;
      }
      if (end_seg) {
        commit(bar_tfull(acc));
        acc ^= 1u;
        if (acc == 0) acc_phase ^= 1u;
        if (--segs_left == 0) segs_left = P.nseg;  // (next tile)
        seg_left = segs_left == 1 ? P.seg_last : P.seg_units;
        d_main = tmem_base + acc * (uint32_t)P.acc_stride;
        accum = 0u;
      }
      if (end_chunk) {
        commit(bar_aempty(abuf));
        if (++abuf == NA) { abuf = 0; aphase ^= 1u; }
        --chunks_left;
        chunk_left = upc;
        c0 += S3_KC;
        if (c0 >= c.src[si].C) { c0 = 0; if (++si == c.nsrc) si = 0; }  // (wraps into the next tile)
        k2 = c.src[si].C - c0 > 16;
        a = a_lo0 + astep * abuf;
        kx = 0;
      }
      if (more) {
        if (end_chunk) wait_chunk(abuf, aphase);
        if (end_seg) wait_acc(acc, acc_phase ^ 1u);
        if (end_group) wait_group(slot, sphase);
        tc_fence_after();
      }
    }
  }
}

// DBG: per-role cycle counters (tc_diag & 128).  A template parameter, not a run-time flag: the timed variants of every
// wait would otherwise sit between the hot instructions of all roles (instruction-cache footprint).
// PAIR: the CTAs of a 2-CTA cluster work on two adjacent pixel tiles with ONE stream of cta_group::2 MMAs (M = 256) issued by
// the leader; each CTA holds half of the weight rows, so the filter bank of a 64 -> 64 3x3 layer takes 72 KB instead of 144 KB
// (room for four halo buffers instead of two), the B-operand reads per SM halve (the N = 64 MMA pair becomes bound by the
// tensor pipe, 192 clk per k-step pair, instead of by shared-memory reads, 224) and a ring streams half the bytes from L2.
// LEAN: the epilogue of the common ResBlock-type layer only (one full N block of 32 / 64 channels, S16 destination, ReLU or
// no activation, at most one S16 operand added in place: s3_plan sets P.lean) -- a kernel of its own, so that neither epilogue
// pays for the other's registers and instruction-cache footprint.
// LEAN == 2: the same structure for the layers with a transcendental epilogue (tanh / sigmoid / sigmoid x h / GRU update, up to
// two S16 operands): the GRU convolutions and Ch_Reducer.  A third kernel family rather than a switch inside the ReLU one: the
// out-of-line activation call in the same store loop cost the ReLU layers 18-30 % (measured).
template <int NMAX, bool DBG, int U, bool PAIR, int LEAN>
__global__ void __launch_bounds__(S3_THREADS, 1) conv_s3_kernel(const __grid_constant__ S3Params P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const demfi_conv_t& c = P.c;
  const int NS = P.ns, NA = P.na;
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t b_base = smem_base + (uint32_t)P.b_off;
  const uint32_t bars = smem_base + (uint32_t)P.bar_off;
  auto bar_rawfull = [&](int a) { return bars + 8u * (uint32_t)a; };
  auto bar_cvfull = [&](int a) { return bars + 8u * (uint32_t)(S3_MAX_NA + a); };
  auto bar_aempty = [&](int a) { return bars + 8u * (uint32_t)(2 * S3_MAX_NA + a); };
  auto bar_tfull = [&](int a) { return bars + 8u * (uint32_t)(3 * S3_MAX_NA + a); };
  auto bar_tempty = [&](int a) { return bars + 8u * (uint32_t)(3 * S3_MAX_NA + 2 + a); };
  const uint32_t bar_wfull = bars + 8u * (uint32_t)(3 * S3_MAX_NA + 4);
  const uint32_t bar_resfull = bars + 8u * (uint32_t)(3 * S3_MAX_NA + 5);
  const uint32_t bar_resfull1 = bars + 8u * (uint32_t)(S3_NBARS - 1);  // operand of the odd tiles (P.res_sep == 2)
  auto bar_bfull = [&](int s) { return bars + 8u * (uint32_t)(3 * S3_MAX_NA + 6 + s); };
  auto bar_bfree = [&](int s) { return bars + 8u * (uint32_t)(3 * S3_MAX_NA + 6 + S3_MAX_NS + s); };
  auto bar_peer = [&](int i) { return bars + 8u * (uint32_t)(S3_BAR_PEER + i); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + (size_t)P.bar_off + 8 * S3_NBARS);
  auto n_of = [&](int nb) { return min(P.nb_max, c.cout_pad - nb * P.nb_max); };
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  // this CTA's tile sequence: tile0, tile0 + tstep, ... < tend.  A pair takes tiles (2q, 2q + 1); for an odd tile count the
  // last tile of rank 1 lies one past the end: its loads are out-of-range boxes (zero fill), its stores are skipped.
  const int tile0 = PAIR ? 2 * (int)(blockIdx.x >> 1) + (int)rank : (int)blockIdx.x;
  const int tstep = PAIR ? (int)(gridDim.x & ~1u) : (int)gridDim.x;
  const int tend = PAIR ? ((P.ntiles + 1) & ~1) : P.ntiles;

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  constexpr bool dbg = DBG;
  const int chunks_per_tile = P.stages_per_tile / P.taps;

  if (threadIdx.x == 0) {
    for (int a = 0; a < S3_MAX_NA; ++a) {
      mbar_init(bar_rawfull(a), 1);
      mbar_init(bar_cvfull(a), S3_CV_WARPS);
      mbar_init(bar_aempty(a), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull(a), 1);
      mbar_init(bar_tempty(a), PAIR ? 2 * S3_EPI_WARPS : S3_EPI_WARPS);  // pair: the leader's copy counts both CTAs' warps
    }
    if (PAIR)
      for (int i = 0; i < S3_MAX_NA + 1 + S3_MAX_NS; ++i) mbar_init(bar_peer(i), 1);
    mbar_init(bar_wfull, 1);
    mbar_init(bar_resfull, 1);
    mbar_init(bar_resfull1, 1);
    for (int s = 0; s < S3_MAX_NS; ++s) {
      mbar_init(bar_bfull(s), 1);
      mbar_init(bar_bfree(s), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == S3_CV_WARPS + S3_EPI_WARPS) {  // the producer warp owns the tensor-memory allocation (the MMA warp leaves early)
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  if (P.diag & 1024) {  // diagnostics (bare issue loop): operands = ordinary fp16 values (operand VALUES change the MMA timing)
    for (int i = threadIdx.x; i < P.stg_off / 4; i += S3_THREADS) {
      uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 97u;
      h ^= h >> 15;
      reinterpret_cast<uint32_t*>(smem)[i] = (0x3800u | (h & 0x3ffu) | ((h >> 3) & 0x8000u)) | ((0x3800u | ((h >> 10) & 0x3ffu)) << 16);
    }
    fence_async_smem();
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all();  // the peer's barriers are initialised before anything arrives on them from here
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);  // warp-uniform for ptxas (uniform registers)
  // Programmatic dependent launch: let the next kernel of the stream start its own prologue (barrier init, TMEM allocation,
  // resident-weight load) on SMs as they drain; everything here that reads or writes activations waits for the previous
  // kernel to have completed (griddepcontrol.wait), the loads of weights and bias (constants) do not.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  // ---- the TMA side of the epilogue for one tile: operand tiles in, result boxes out.  Run by thread 0 of the epilogue
  // warps, or -- when every source is S16 and the converter warps have nothing to do -- by lane 0 of converter warp 0
  // (P.offload): the issue of a bulk tensor store blocks the issuing thread for several hundred cycles and the wait for its
  // shared-memory reads for as long again, and on an epilogue warp both delay that warp's next accumulator drain, i.e. the
  // hand-over every MMA of the next tile but one waits for (measured: 1.4-2 kclk per tile on the critical warp).
  const uint32_t stg_base = smem_base + (uint32_t)P.stg_off;
  auto epi_fetch = [&](int nb, int tx0, int ty0, int n, int nboxes, uint32_t xoff, uint32_t rbar) {
    const bool pb = P.nsb > 1;
    const int e0 = nb * P.nsb, ne = pb ? nboxes : 1;
    uint32_t tx = 0;
    for (int sb = 0; sb < ne; ++sb) tx += (uint32_t)(P.e_nres[e0 + sb] * (pb ? 1 : nboxes) * S3_BOX_BYTES);
    if (tx == 0) return;
    mbar_arrive_expect_tx(rbar, tx);
    for (int sb = 0; sb < ne; ++sb) {
      const int e = e0 + sb, nr = P.e_nres[e];
      if (nr == 0) continue;
      const int sg = P.e_seg[e];
      const int b0 = pb ? sb : 0, b1 = pb ? sb + 1 : nboxes;
      for (int b = b0; b < b1; ++b) {
        tma_load_4d(stg_base + xoff + (uint32_t)(P.e_roff[e] + b * S3_BOX_BYTES), &P.rmap[sg], rbar, P.e_rc0[e] + 32 * (b - b0), tx0, ty0, n);
        if (nr > 1)
          tma_load_4d(stg_base + (uint32_t)(P.stg2_off + b * S3_BOX_BYTES), &P.r2map[sg], rbar, P.e_rc0[e] + 32 * (b - b0), tx0, ty0, n);
      }
    }
  };
  auto epi_store = [&](int nb, int tx0, int ty0, int n, int nboxes, uint32_t xoff) {
    const bool pb = P.nsb > 1;
    const int e0 = nb * P.nsb, ne = pb ? nboxes : 1;
    for (int sb = 0; sb < ne; ++sb) {
      const int e = e0 + sb;
      if (P.e_seg[e] < 0) continue;
      const int b0 = pb ? sb : 0, b1 = pb ? sb + 1 : nboxes;
      for (int j = P.e_o0[e]; j < P.e_o0[e] + P.e_on[e]; ++j)
        for (int b = b0; b < b1; ++b)
          tma_store_4d(&P.omap[P.ol_map[j]], stg_base + xoff + (uint32_t)(b * S3_BOX_BYTES), P.ol_c0[j] + 32 * (b - b0), tx0, ty0, n);
    }
    bulk_commit();
  };

  if (warp == S3_CV_WARPS + S3_EPI_WARPS + 1) {
    // ===== MMA issuer: ONE elected thread runs the whole persistent loop, and its warp takes part in nothing afterwards.
    // What sets its pace (round-2 measurements: tools/diag_timers.py, tools/mma_probe_pair.cu, ncu source page): with the
    // activation loads, the tcgen05.ld drains and the whole epilogue switched off, the thread still needed ~300 clk per
    // (tap, 32-channel) stage against 224 clk for the same MMAs in the probe.  The tensor core runs only ~100 clk ahead of
    // the issuing thread, and every instruction that goes through the memory-I/O queue (R2UR, mbarrier waits, commits) is
    // ordered behind the UTCHMMAs already issued, so the instructions after one of them are exposed.  ptxas kept the whole
    // loop state in vector registers -- 5-9 R2UR per tap -- for two reasons found by bisection on the SASS: (1) a call in
    // the loop (the out-of-line bounded wait), (2) the elected thread re-joining its warp for the block-wide barrier at the
    // end of the kernel.  With neither, the state lives in uniform registers and a unit of issue (a kernel row of taps,
    // straight-line) is UTCHMMA + uniform-datapath adds only. =====
    if (elect_one()) {
      const long long t_begin = dbg ? clock64() : 0;
      long long w_tempty = 0, w_ready = 0, w_peer = 0;
      if (!PAIR) s3_issue<U, DBG, 0>(P, smem_base, tmem_base, bars, w_tempty, w_ready, w_peer);
      else if (rank == 0) s3_issue<U, DBG, 1>(P, smem_base, tmem_base, bars, w_tempty, w_ready, w_peer);
      else s3_issue<U, DBG, 2>(P, smem_base, tmem_base, bars, w_tempty, w_ready, w_peer);
      if (P.diag & 1024) {  // bare issue loop: everything has completed when this commit arrives
        umma_commit(bar_resfull);
        mbar_wait_nocall(bar_resfull, 0);
      }
      if (dbg) {
        long long* d = P.dbg + (size_t)blockIdx.x * 16;
        d[12] = clock64() - t_begin; d[13] = w_tempty; d[14] = w_ready; d[15] = w_peer;
      }
    }
    return;  // no barrier below involves this warp: the remaining 15 warps meet at named barrier 1
  }
  if (P.diag & 1024) {
    // diagnostics: nothing but the MMA issue loop runs
  } else if (warp < S3_CV_WARPS) {
    // ===== converter: fp32 halo tile -> fp16 hi / lo planes, in place =====
    const int tid = (int)threadIdx.x;
    long long w_raw = 0, w_cv = 0;
    const long long t_begin = dbg ? clock64() : 0;
    int abuf = 0;
    uint32_t aphase = 0;
    if (P.offload && warp == 0) {
      // ===== store warp (every source S16): the TMA side of the epilogue, tile by tile; barrier 2 = "staging tile free and the
      // operand tiles requested" (this warp arrives, the epilogue threads wait), barrier 3 = "staging tile written" (they
      // arrive, this warp waits) =====
      asm volatile("griddepcontrol.wait;" ::: "memory");
      bool pending = false;
      struct TilePos { int nb, tx0, ty0, n, nboxes; bool dummy; };
      auto pos_of = [&](int tile) {
        TilePos q;
        int t = tile;
        q.nb = t % P.n_blocks;
        t /= P.n_blocks;
        q.tx0 = (t % P.tiles_x) * S3_TW;
        t /= P.tiles_x;
        q.ty0 = (t % P.tiles_y) * S3_TH;
        q.n = t / P.tiles_y;
        q.nboxes = (n_of(q.nb) + 31) >> 5;
        q.dummy = PAIR && tile >= P.ntiles;
        return q;
      };
      if (!P.res_sep) {
        for (int tile = tile0; tile < tend; tile += tstep) {
          const TilePos q = pos_of(tile);
          if (lane == 0) {
            if (pending) bulk_wait_read0();
            if (!q.dummy) epi_fetch(q.nb, q.tx0, q.ty0, q.n, q.nboxes, 0u, bar_resfull);
          }
          __syncwarp();
          asm volatile("bar.arrive 2, %0;" ::"n"(S3_EPI_THREADS + 32) : "memory");
          asm volatile("bar.sync 3, %0;" ::"n"(S3_EPI_THREADS + 32) : "memory");
          if (lane == 0 && !q.dummy) {
            epi_store(q.nb, q.tx0, q.ty0, q.n, q.nboxes, 0u);
            pending = true;
          }
        }
      } else if (P.res_sep == 2) {
        // Two tiles X0 / X1 used in turn: tile k of this CTA finds its operand in X(k & 1), adds its result IN PLACE and the
        // TMA store reads it from there.  The operand of tile k + 2 is requested as soon as the store of tile k has drained
        // X(k & 1) -- a whole tile ahead of its use (with ONE operand tile the request went out ~0.7 kclk before the store
        // loop needed it: role timers, 2.5 kclk of every 8.2 kclk tile of a ResBlock conv2 waited for it), and neither that
        // drain nor the request sits between two store loops any more: barrier 2 is released right after the store is issued.
        auto prefetch = [&](const TilePos& q) {
          if (!q.dummy)
            for (int b = 0; b < q.nboxes; ++b) tma_prefetch_4d(&P.rmap[P.e_seg[0]], P.e_rc0[0] + 32 * b, q.tx0, q.ty0, q.n);
        };
        const uint32_t x1 = (uint32_t)P.stg2_off;
        uint32_t xoff = 0;
        if (lane == 0) {
          if (tile0 < tend) {
            const TilePos q = pos_of(tile0);
            if (!q.dummy) epi_fetch(q.nb, q.tx0, q.ty0, q.n, q.nboxes, 0u, bar_resfull);
          }
          if (tile0 + tstep < tend) {
            const TilePos q = pos_of(tile0 + tstep);
            if (!q.dummy) epi_fetch(q.nb, q.tx0, q.ty0, q.n, q.nboxes, x1, bar_resfull1);
          }
          if (tile0 + 2 * tstep < tend) prefetch(pos_of(tile0 + 2 * tstep));
        }
        __syncwarp();
        if (tile0 < tend) asm volatile("bar.arrive 2, %0;" ::"n"(S3_EPI_THREADS + 32) : "memory");
        for (int tile = tile0; tile < tend; tile += tstep) {
          const TilePos q = pos_of(tile);
          asm volatile("bar.sync 3, %0;" ::"n"(S3_EPI_THREADS + 32) : "memory");
          if (lane == 0 && !q.dummy) {
            epi_store(q.nb, q.tx0, q.ty0, q.n, q.nboxes, xoff);
            pending = true;
          }
          __syncwarp();
          // the next tile works in the OTHER buffer: drained and re-filled during the previous iteration
          if (tile + tstep < tend) asm volatile("bar.arrive 2, %0;" ::"n"(S3_EPI_THREADS + 32) : "memory");
          if (lane == 0 && tile + 2 * tstep < tend) {
            if (pending) bulk_wait_read0();
            const TilePos qn = pos_of(tile + 2 * tstep);
            if (!qn.dummy) epi_fetch(qn.nb, qn.tx0, qn.ty0, qn.n, qn.nboxes, xoff, xoff ? bar_resfull1 : bar_resfull);
            if (tile + 3 * tstep < tend) prefetch(pos_of(tile + 3 * tstep));
          }
          __syncwarp();
          xoff ^= x1;
        }
      } else {
        // operand tile of its own: fetched for the NEXT tile right after this tile's store loop is through with it (barrier
        // 3), i.e. a whole TMA store + drain ahead of its use, and prefetched into L2 one tile earlier still
        auto prefetch = [&](const TilePos& q) {
          if (!q.dummy)
            for (int b = 0; b < q.nboxes; ++b) tma_prefetch_4d(&P.rmap[P.e_seg[0]], P.e_rc0[0] + 32 * b, q.tx0, q.ty0, q.n);
        };
        if (tile0 < tend && lane == 0) {
          const TilePos q = pos_of(tile0);
          if (!q.dummy) epi_fetch(q.nb, q.tx0, q.ty0, q.n, q.nboxes, 0u, bar_resfull);
          if (tile0 + tstep < tend) prefetch(pos_of(tile0 + tstep));
        }
        __syncwarp();
        if (tile0 < tend) asm volatile("bar.arrive 2, %0;" ::"n"(S3_EPI_THREADS + 32) : "memory");
        for (int tile = tile0; tile < tend; tile += tstep) {
          const TilePos q = pos_of(tile);
          const bool more = tile + tstep < tend;
          asm volatile("bar.sync 3, %0;" ::"n"(S3_EPI_THREADS + 32) : "memory");
          if (lane == 0) {
            if (!q.dummy) {
              epi_store(q.nb, q.tx0, q.ty0, q.n, q.nboxes, 0u);
              pending = true;
            }
            if (more) {
              const TilePos qn = pos_of(tile + tstep);
              if (!qn.dummy) epi_fetch(qn.nb, qn.tx0, qn.ty0, qn.n, qn.nboxes, 0u, bar_resfull);
              if (tile + 2 * tstep < tend) prefetch(pos_of(tile + 2 * tstep));
            }
            if (pending) bulk_wait_read0();
          }
          __syncwarp();
          if (more) asm volatile("bar.arrive 2, %0;" ::"n"(S3_EPI_THREADS + 32) : "memory");
        }
      }
      if (lane == 0 && pending) bulk_wait0();  // stores complete before the CTA exits
    }
    for (int tile = tile0; tile < tend && !P.all_s16; tile += tstep) {  // (every source S16: nothing to do here)
      for (int si = 0; si < c.nsrc; ++si) {
        const bool s16 = c.src[si].fmt == DEMFI_FMT_S16;  // already fp16 hi | lo rows: passed on as it landed
        for (int c0 = 0; c0 < c.src[si].C; c0 += S3_KC) {
          // every chunk passes through these warps (see s3_issue): they are part of the buffer hand-over, never lapped
          mbar_wait_t(bar_rawfull(abuf), aphase, dbg, w_raw);
          const long long t_cv = dbg ? clock64() : 0;
          if (!s16) {
            const uint32_t a_addr = smem_base + (uint32_t)(abuf * P.a_bytes);
#pragma unroll 1
            for (int p = (P.diag & 16) ? P.halo_px : tid; p < P.halo_px; p += S3_CV_THREADS) {
              // row = pixel: 32 fp32 -> [Ah 32 x fp16 | Al 32 x fp16], same 128 bytes, same 16-byte-chunk swizzle (chunk ^ (p & 7))
              const uint32_t row = a_addr + (uint32_t)p * 128u;
              const uint32_t sw = (uint32_t)p & 7u;
              uint32_t hi[16], lo[16];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const uint4 v = lds128(row + ((((uint32_t)j) ^ sw) << 4));
                split2(__uint_as_float(v.x), __uint_as_float(v.y), hi[2 * j], lo[2 * j]);
                split2(__uint_as_float(v.z), __uint_as_float(v.w), hi[2 * j + 1], lo[2 * j + 1]);
              }
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                sts128(row + ((((uint32_t)j) ^ sw) << 4), make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]));
                sts128(row + ((((uint32_t)(j + 4)) ^ sw) << 4), make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]));
              }
            }
            fence_async_smem();  // generic-proxy writes -> visible to the tensor core / TMA (async proxy)
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_cvfull(abuf));
          if (dbg) w_cv += clock64() - t_cv;
          if (++abuf == NA) { abuf = 0; aphase ^= 1u; }
        }
      }
    }
    if (dbg && threadIdx.x == 0 && !P.all_s16) {
      long long* d = P.dbg + (size_t)blockIdx.x * 16;
      d[0] = clock64() - t_begin; d[1] = w_raw; d[7] = w_cv;
    }
  } else if (warp < S3_CV_WARPS + S3_EPI_WARPS) {
    // ===== epilogue: TMEM lane = pixel row; a warp can only touch lanes 32*(warp%4)...  Group 0 takes accumulator
    // columns [0, csplit), group 1 takes [csplit, N). =====
    const int e_tid = (int)threadIdx.x - S3_CV_THREADS;
    const int grp = (warp - S3_CV_WARPS) >> 2;
    const int wq = warp & 3;
    const int m = wq * 32 + lane;
    constexpr int HMAX = (NMAX / 2 + 15) / 16 * 16;
    int acc = 0;
    uint32_t acc_phase = 0, res_phase = 0, res_phase1 = 0;
    const bool res_alt = LEAN == 1 && P.res_sep == 2;  // operand and result share one of two tiles, used in turn (store warp)
    uint32_t xoff = 0;                                  // this tile's: 0 or P.stg2_off
    long long w_tfull = 0, w_store = 0, w_ld = 0, w_arr = 0, w_s1 = 0, w_s2 = 0, w_s3 = 0, w_top = 0;
    const long long t_begin = dbg ? clock64() : 0;
    const uint32_t stg = smem_base + (uint32_t)P.stg_off;
    bool store_pending = false;
    // bias -> shared memory once per CTA: the per-tile global loads (L2 latency on a cold L1 line) stalled every warp of the
    // store phase (ncu: long-scoreboard stalls on the bias FADDs)
    const uint32_t bias_s = smem_base + (uint32_t)P.bias_off;
    for (int i = e_tid * 4; i < c.cout_pad; i += S3_EPI_THREADS * 4) {
      const float4 b = ld4(c.bias + i);
      sts128(bias_s + (uint32_t)i * 4u, make_uint4(__float_as_uint(b.x), __float_as_uint(b.y), __float_as_uint(b.z), __float_as_uint(b.w)));
    }
    asm volatile("bar.sync 4, %0;" ::"n"(S3_EPI_THREADS) : "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");  // operand tiles and destinations belong to earlier kernels
    const bool offload = P.offload != 0;
    // lean epilogue (host: s3_plan): one full N block (N == NMAX), one entry, S16 destination, activation none / ReLU, at most
    // one S16 operand in the result tile
    constexpr bool LEANOK = LEAN != 0;
    constexpr bool lean = LEAN != 0;
    const bool lean_relu = (P.lean & 2) != 0, lean_res = (P.lean & 4) != 0;
    const uint32_t lean_c0 = (uint32_t)(grp * HMAX);  // first channel of this thread (N == NMAX: csplit == HMAX)
    const uint32_t lean_bias = bias_s + lean_c0 * 4u;
    // the thread's pixel row never changes: row of its staging box, swizzle key, first 16-byte chunk
    const uint32_t lean_row0 = stg + (uint32_t)m * 128u + (lean_c0 >> 5) * (uint32_t)S3_BOX_BYTES;
    const uint32_t lean_sw = (uint32_t)m & 7u, lean_g0 = (lean_c0 & 31u) >> 3;
    const uint32_t lean_res_row0 = lean_row0 + (uint32_t)P.e_roff[0];  // the skip operand: in place, or in the operand tile of its own
    const bool per_box = P.nsb > 1;  // entries are 32-channel boxes (else one entry per N block)
    for (int tile = tile0; tile < tend; tile += tstep) {
      const bool dummy = PAIR && tile >= P.ntiles;  // (odd tile count: the pair's second tile does not exist)
      int t = tile;
      const int nb = t % P.n_blocks;
      t /= P.n_blocks;
      const int tx0 = (t % P.tiles_x) * S3_TW;
      t /= P.tiles_x;
      const int ty0 = (t % P.tiles_y) * S3_TH;
      const int n = t / P.tiles_y;
      const int N = n_of(nb), n0 = nb * P.nb_max;
      const int csplit = ((N / 2 + 15) / 16) * 16;
      const int cbeg = grp ? csplit : 0, cnum = grp ? N - csplit : csplit;
      const int oy = ty0 + (m >> 3), ox = tx0 + (m & 7);
      const bool valid = (oy < c.H) && (ox < c.W);
      const int nboxes = (N + 31) >> 5;
      const int e0 = nb * P.nsb;                                  // first entry of this N block
      const int ne = per_box ? nboxes : 1;                        // entries of this N block
      int nres_any = 0;
      if (P.tma_epi && !dummy)
        for (int sb = 0; sb < ne; ++sb) nres_any += P.e_nres[e0 + sb];
      const long long t_top = dbg ? clock64() : 0;
      if (P.tma_epi && !offload) {
        // the staging buffers are free once the previous tile's stores have read them; then fetch the operand tiles
        if (e_tid == 0) {
          if (store_pending) bulk_wait_read0();
          if (nres_any > 0) epi_fetch(nb, tx0, ty0, n, nboxes, 0u, bar_resfull);
        }
        store_pending = true;
      }
      if (dbg) w_top += clock64() - t_top;
      float sum[HMAX];
      bool first = true;
      for (int done = 0; done < P.stages_per_tile; done += P.flush) {
        const float gain = 1.0f + P.comp * (float)(2 * min(P.flush, P.stages_per_tile - done));
        mbar_wait_t(bar_tfull(acc), acc_phase, dbg, w_tfull);
        tc_fence_after();
        // accumulator columns of this thread's channels: [main | corrections] (one CTA); a pair's main tile interleaves the
        // halves held by the two CTAs -- [Ah Bh | Ah Bl] of channels 0..N/2-1, then of N/2..N-1 -- followed by Al Bh for all N
        const uint32_t tbuf = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(acc * P.acc_stride);
        const uint32_t taddr = tbuf + (uint32_t)(PAIR ? grp * N : cbeg);
        const uint32_t tcorr = tbuf + (uint32_t)(PAIR ? grp * N + (N >> 1) : N + cbeg);
        const uint32_t tcorr2 = tbuf + (uint32_t)(2 * N + cbeg);
        const long long t_ld0 = dbg ? clock64() : 0;
        if (LEANOK && lean && P.nseg == 1) {
          // full block (N == NMAX), ONE accumulation segment per tile (no partial sums live): the thread's HMAX main columns
          // and HMAX correction columns are requested together and waited for once (one tensor-memory round trip instead of
          // two; a pair adds the Al x Bh columns in a second one)
          uint32_t ra[HMAX], rb[HMAX];
#pragma unroll
          for (int col = 0; col < HMAX; col += 16) tmem_ld16_nowait(taddr + (uint32_t)col, ra + col);
#pragma unroll
          for (int col = 0; col < HMAX; col += 16) tmem_ld16_nowait(tcorr + (uint32_t)col, rb + col);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < HMAX; ++j) {
            sum[j] = fmaf(__uint_as_float(rb[j]), 1.0f / S3_LO_SCALE, __uint_as_float(ra[j]) * gain);
          }
          if (PAIR) {
#pragma unroll
            for (int col = 0; col < HMAX; col += 16) tmem_ld16_nowait(tcorr2 + (uint32_t)col, rb + col);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < HMAX; ++j) sum[j] = fmaf(__uint_as_float(rb[j]), 1.0f / S3_LO_SCALE, sum[j]);
          }
        } else
        // 32 columns at a time (the wide N = 128 blocks hold 64 columns per thread: the partial sums stay in registers, the
        // drained values pass through a 32-register window)
#pragma unroll
        for (int c32 = 0; c32 < HMAX; c32 += 32) {

          constexpr int RW = HMAX < 32 ? HMAX : 32;
          uint32_t r[RW];
#pragma unroll
          for (int col = 0; col < RW; col += 16)
            if (c32 + col < HMAX && c32 + col < cnum && !(P.diag & 512)) tmem_ld16_nowait(taddr + (uint32_t)(c32 + col), r + col);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < RW; ++j) {
            if (c32 + j < HMAX) {
              const float v = __uint_as_float(r[j]) * gain;
              sum[c32 + j] = first ? v : sum[c32 + j] + v;
            }
          }
#pragma unroll
          for (int col = 0; col < RW; col += 16)
            if (c32 + col < HMAX && c32 + col < cnum && !(P.diag & 512)) tmem_ld16_nowait(tcorr + (uint32_t)(c32 + col), r + col);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < RW; ++j)
            if (c32 + j < HMAX) sum[c32 + j] = fmaf(__uint_as_float(r[j]), 1.0f / S3_LO_SCALE, sum[c32 + j]);
          if (PAIR) {
#pragma unroll
            for (int col = 0; col < RW; col += 16)
              if (c32 + col < HMAX && c32 + col < cnum) tmem_ld16_nowait(tcorr2 + (uint32_t)(c32 + col), r + col);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < RW; ++j)
              if (c32 + j < HMAX) sum[c32 + j] = fmaf(__uint_as_float(r[j]), 1.0f / S3_LO_SCALE, sum[c32 + j]);
          }
        }
        const long long t_ld1 = dbg ? clock64() : 0;
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (PAIR && rank != 0) mbar_arrive_remote(bar_tempty(acc), 0u);
          else mbar_arrive(bar_tempty(acc));
        }
        if (dbg) { w_ld += t_ld1 - t_ld0; w_arr += clock64() - t_ld1; }
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        first = false;
      }
      const long long t_store = dbg ? clock64() : 0;
      if (P.tma_epi) {
        // ---- staged epilogue: bias, operands, activation -> swizzled box layout -> TMA store ----
        if (offload) asm volatile("bar.sync 2, %0;" ::"n"(S3_EPI_THREADS + 32) : "memory");  // the store warp released the staging tile
        if (nres_any > 0) {
          if (res_alt && xoff != 0u) {
            mbar_wait(bar_resfull1, res_phase1);
            res_phase1 ^= 1u;
          } else {
            mbar_wait(bar_resfull, res_phase);
            res_phase ^= 1u;
          }
        } else if (!offload) {
          asm volatile("bar.sync 2, %0;" ::"n"(S3_EPI_THREADS) : "memory");  // staging buffer released (thread 0 waited)
        }
        if (dbg) w_s1 += clock64() - t_store;  // waited for the staging tile / the operand tiles
        const uint32_t row = stg + (uint32_t)m * 128u;
        const uint32_t sw = (uint32_t)m & 7u;
        const float4 zero4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        // Eight columns per step.  The step's entry (the N block, or its 32-channel box) says through ONE packed word --
        // independent loads, no dependent chain across the unrolled steps -- whether the result is stored S16 (one hi chunk +
        // one lo chunk of the pixel's 128-byte group row; an operand tile in the same format is in the same place: read, then
        // overwritten in place) or fp32 (two 4-channel chunks), with which activation and operands.
        auto step8 = [&](int col, int chn, uint32_t info) {
          const int sfmt = (int)(info & 7u), act = (int)((info >> 3) & 7u), nres = (int)((info >> 6) & 3u);
          const uint32_t roff = (info & 256u) ? (uint32_t)P.stg2_off : 0u;
          const uint32_t base = row + (uint32_t)((chn >> 5) * S3_BOX_BYTES);
          const float4 b0 = as_f4(lds128(bias_s + (uint32_t)(n0 + chn) * 4u)), b1 = as_f4(lds128(bias_s + (uint32_t)(n0 + chn + 4) * 4u));
          float4 v0 = make_float4(sum[col] + b0.x, sum[col + 1] + b0.y, sum[col + 2] + b0.z, sum[col + 3] + b0.w);
          float4 v1 = make_float4(sum[col + 4] + b1.x, sum[col + 5] + b1.y, sum[col + 6] + b1.z, sum[col + 7] + b1.w);
          float4 h0 = zero4, h1 = zero4, z0 = zero4, z1 = zero4;
          if (sfmt & DEMFI_SEG_DST_S16) {
            if (nres > 0) op_load8(base + roff, chn, sw, (sfmt & DEMFI_SEG_RES_S16) != 0, h0, h1);
            if (nres > 1) op_load8(base + (uint32_t)P.stg2_off, chn, sw, (sfmt & DEMFI_SEG_RES2_S16) != 0, z0, z1);
          } else {
            if (nres > 0) {
              h0 = op_load4(base + roff, chn, sw, (sfmt & DEMFI_SEG_RES_S16) != 0);
              h1 = op_load4(base + roff, chn + 4, sw, (sfmt & DEMFI_SEG_RES_S16) != 0);
            }
            if (nres > 1) {
              z0 = op_load4(base + (uint32_t)P.stg2_off, chn, sw, (sfmt & DEMFI_SEG_RES2_S16) != 0);
              z1 = op_load4(base + (uint32_t)P.stg2_off, chn + 4, sw, (sfmt & DEMFI_SEG_RES2_S16) != 0);
            }
          }
          if (act == DEMFI_ACT_NONE || act == DEMFI_ACT_RELU) {  // the common case stays inline (a few FADD / FMNMX)
            v0.x += h0.x; v0.y += h0.y; v0.z += h0.z; v0.w += h0.w;
            v1.x += h1.x; v1.y += h1.y; v1.z += h1.z; v1.w += h1.w;
            if (act == DEMFI_ACT_RELU) {
              v0.x = fmaxf(v0.x, 0.0f); v0.y = fmaxf(v0.y, 0.0f); v0.z = fmaxf(v0.z, 0.0f); v0.w = fmaxf(v0.w, 0.0f);
              v1.x = fmaxf(v1.x, 0.0f); v1.y = fmaxf(v1.y, 0.0f); v1.z = fmaxf(v1.z, 0.0f); v1.w = fmaxf(v1.w, 0.0f);
            }
          } else {
            v0 = finish4v(act, v0, h0, z0);
            v1 = finish4v(act, v1, h1, z1);
          }
          if (sfmt & DEMFI_SEG_DST_S16) {
            const uint32_t g8 = (uint32_t)((chn & 31) >> 3);
            uint4 hi, lo;
            s16_encode8(v0, v1, hi, lo);
            sts128(base + ((g8 ^ sw) << 4), hi);
            sts128(base + (((g8 + 4u) ^ sw) << 4), lo);
          } else {
            const uint32_t q = (uint32_t)((chn & 31) >> 2);
            sts128(base + ((q ^ sw) << 4), as_u4(v0));
            sts128(base + (((q + 1u) ^ sw) << 4), as_u4(v1));
          }
        };
        if (LEAN >= 2) {
          // Lean store loop with an activation: as below, plus the operand tiles (h in place, z in the second tile) and the
          // out-of-line exp / reciprocal expansions
          const uint32_t info = (uint32_t)P.e_info[e0];
          // LEAN >= 3: the activation LEAN - 3 is a compile-time constant and its expansion inline (kernels of their own for the
          // layers of the network, s3_lean2_act_kernel).  Role timers on the run-time-activation kernel: 5.6 kclk per tile in this
          // loop against 2.0 kclk for the ReLU kernels, whatever the activation (a 1x1 with none: 5.6) -- the out-of-line call per
          // four channels (the thread's 32 partial sums live across 16 calls), not the exp / reciprocal, was the cost.
          const int act = LEAN >= 3 ? LEAN - 3 : (int)((info >> 3) & 7u);
          const int nres = (int)((info >> 6) & 3u);
          const uint32_t lean_row = lean_row0, lean_res_row = lean_res_row0;
          const uint32_t res2_row = lean_row + (uint32_t)P.stg2_off;
          // fp32 destination (the layers that feed the warps: F0 / F1, rF0 / rF1, the FAC-FB encoder output) and / or fp32 operand:
          // eight channels are two 16-byte chunks of the row instead of one hi + one lo chunk
          const bool dst_f32 = (P.lean & 16) != 0, res_f32 = (P.lean & 32) != 0;
#pragma unroll
          for (int s8 = 0; s8 < HMAX / 8; ++s8) {
            const uint32_t o_hi = (((uint32_t)s8 + lean_g0) ^ lean_sw) << 4, o_lo = (((uint32_t)s8 + lean_g0 + 4u) ^ lean_sw) << 4;
            const uint32_t f_0 = ((2u * ((uint32_t)s8 + lean_g0)) ^ lean_sw) << 4, f_1 = ((2u * ((uint32_t)s8 + lean_g0) + 1u) ^ lean_sw) << 4;
            const float4 b0 = as_f4(lds128(lean_bias + (uint32_t)s8 * 32u)), b1 = as_f4(lds128(lean_bias + (uint32_t)s8 * 32u + 16u));
            float4 v0 = make_float4(sum[8 * s8] + b0.x, sum[8 * s8 + 1] + b0.y, sum[8 * s8 + 2] + b0.z, sum[8 * s8 + 3] + b0.w);
            float4 v1 = make_float4(sum[8 * s8 + 4] + b1.x, sum[8 * s8 + 5] + b1.y, sum[8 * s8 + 6] + b1.z, sum[8 * s8 + 7] + b1.w);
            float4 h0 = make_float4(0.f, 0.f, 0.f, 0.f), h1 = h0, z0 = h0, z1 = h0;
            if (nres > 0) {
              if (res_f32) { h0 = as_f4(lds128(lean_res_row + f_0)); h1 = as_f4(lds128(lean_res_row + f_1)); }
              else s16_decode8(lds128(lean_res_row + o_hi), lds128(lean_res_row + o_lo), h0, h1);
            }
            if (nres > 1) s16_decode8(lds128(res2_row + o_hi), lds128(res2_row + o_lo), z0, z1);
            if (LEAN >= 3) {
              v0 = finish4_inl(act, v0, h0, z0);
              v1 = finish4_inl(act, v1, h1, z1);
            } else {
              v0 = finish4v(act, v0, h0, z0);
              v1 = finish4v(act, v1, h1, z1);
            }
            if (dst_f32) {
              sts128(lean_row + f_0, as_u4(v0));
              sts128(lean_row + f_1, as_u4(v1));
            } else {
              uint4 hi, lo;
              s16_encode8(v0, v1, hi, lo);
              sts128(lean_row + o_hi, hi);
              sts128(lean_row + o_lo, lo);
            }
          }
        } else if (LEANOK && lean) {
          // Lean store loop (full block, one entry, S16 destination, ReLU or none, at most one S16 operand updated in place):
          // the thread's pixel row never changes, so its eight chunk addresses of the staging box are computed once per kernel
          // and a step is two bias loads, the adds, the split and two stores -- no address arithmetic, no per-step lookups.
          const uint32_t lean_row = lean_row0 + xoff;                           // (res_alt: the tile of this turn)
          const uint32_t lean_res_row = res_alt ? lean_row : lean_res_row0;    // ... whose operand is updated in place
#pragma unroll
          for (int s8 = 0; s8 < HMAX / 8; ++s8) {
            const float4 b0 = as_f4(lds128(lean_bias + (uint32_t)s8 * 32u)), b1 = as_f4(lds128(lean_bias + (uint32_t)s8 * 32u + 16u));
            float4 v0 = make_float4(sum[8 * s8] + b0.x, sum[8 * s8 + 1] + b0.y, sum[8 * s8 + 2] + b0.z, sum[8 * s8 + 3] + b0.w);
            float4 v1 = make_float4(sum[8 * s8 + 4] + b1.x, sum[8 * s8 + 5] + b1.y, sum[8 * s8 + 6] + b1.z, sum[8 * s8 + 7] + b1.w);
            if (lean_res) {
              float4 h0, h1;
              s16_decode8(lds128(lean_res_row + ((((uint32_t)s8 + lean_g0) ^ lean_sw) << 4)),
                          lds128(lean_res_row + ((((uint32_t)s8 + lean_g0 + 4u) ^ lean_sw) << 4)), h0, h1);
              v0.x += h0.x; v0.y += h0.y; v0.z += h0.z; v0.w += h0.w;
              v1.x += h1.x; v1.y += h1.y; v1.z += h1.z; v1.w += h1.w;
            }
            if (lean_relu) {
              v0.x = fmaxf(v0.x, 0.0f); v0.y = fmaxf(v0.y, 0.0f); v0.z = fmaxf(v0.z, 0.0f); v0.w = fmaxf(v0.w, 0.0f);
              v1.x = fmaxf(v1.x, 0.0f); v1.y = fmaxf(v1.y, 0.0f); v1.z = fmaxf(v1.z, 0.0f); v1.w = fmaxf(v1.w, 0.0f);
            }
            uint4 hi, lo;
            s16_encode8(v0, v1, hi, lo);
            sts128(lean_row + ((((uint32_t)s8 + lean_g0) ^ lean_sw) << 4), hi);
            sts128(lean_row + ((((uint32_t)s8 + lean_g0 + 4u) ^ lean_sw) << 4), lo);
          }
        } else
        // one entry per N block (the common case): its packed word is read once per tile; per-box plans look it up per step
        if (!per_box) {
          const uint32_t info = (uint32_t)P.e_info[e0];
#pragma unroll
          for (int col = 0; col < HMAX; col += 8)
            if (col < cnum && !(P.diag & 32)) step8(col, cbeg + col, info);
        } else {
#pragma unroll
          for (int col = 0; col < HMAX; col += 8)
            if (col < cnum && !(P.diag & 32)) step8(col, cbeg + col, (uint32_t)P.e_info[e0 + ((cbeg + col) >> 5)]);
        }
        const long long t_s2 = dbg ? clock64() : 0;
        fence_async_smem();
        if (offload) {
          asm volatile("bar.arrive 3, %0;" ::"n"(S3_EPI_THREADS + 32) : "memory");  // the store warp takes it from here
        } else {
          asm volatile("bar.sync 3, %0;" ::"n"(S3_EPI_THREADS) : "memory");
        }
        if (dbg) w_s2 += clock64() - t_s2;
        const long long t_s3 = dbg ? clock64() : 0;
        if (!offload && e_tid == 0 && !dummy && !(P.diag & (1 | 32))) epi_store(nb, tx0, ty0, n, nboxes, 0u);
        if (dbg) w_s3 += clock64() - t_s3;
      } else if (valid && !(P.diag & 1)) {
        const int ch_lo = n0 + cbeg, ch_hi = ch_lo + cnum;
#pragma unroll 1
        for (int sgi = 0; sgi < c.nseg; ++sgi) {
          const demfi_seg_t& sg = c.seg[sgi];
          if (sg.ch0 >= ch_hi || sg.ch0 + sg.nch <= ch_lo) continue;
          const SegCursor cur = seg_cursor(c, sg, n, oy, ox);
#pragma unroll
          for (int col = 0; col < HMAX; col += 4) {
            if (col < cnum) {
              const float4 b = ld4(c.bias + ch_lo + col);
              seg_emit4(cur, ch_lo + col, make_float4(sum[col] + b.x, sum[col + 1] + b.y, sum[col + 2] + b.z, sum[col + 3] + b.w));
            }
          }
        }
      }
      if (dbg) w_store += clock64() - t_store;
      if (res_alt) xoff ^= (uint32_t)P.stg2_off;
    }
    if (P.tma_epi && !offload && e_tid == 0 && store_pending) bulk_wait0();  // stores complete before the CTA exits
    if (dbg && e_tid == 0) {
      long long* d = P.dbg + (size_t)blockIdx.x * 16;
      d[4] = clock64() - t_begin; d[5] = w_tfull; d[6] = w_store; d[10] = w_ld; d[11] = w_arr; d[2] = w_s1; d[3] = w_s2;
      if (P.all_s16) { d[1] = w_top; d[7] = w_s3; }
    }
  } else if (warp == S3_CV_WARPS + S3_EPI_WARPS) {
    // ===== TMA producer (one thread): halo tiles and, unless resident, the weight ring.  The two sequences are
    // advanced by polling so that a full weight ring never delays the next halo tile. =====
    if (elect_one()) {
      const long long t_begin = dbg ? clock64() : 0;
      const uint32_t a_tx = (uint32_t)(P.halo_px * 128);
      // packed weights of a pair: [rank][chunk][tap][N rows = Bh half ; Bl half][32 fp16] -- this CTA streams its own half
      const size_t w_rank_off = PAIR ? (size_t)rank * (size_t)P.stages_per_tile * (size_t)P.b_bytes : 0;
      if (P.resident) {
        const uint32_t b_tx = PAIR ? (uint32_t)P.b_bytes : (uint32_t)n_of(0) * 128u;
        mbar_arrive_expect_tx(bar_wfull, b_tx * (uint32_t)P.stages_per_tile);
        const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(c.wpack) + w_rank_off;
        for (int s = 0; s < P.stages_per_tile; ++s)
          bulk_load(b_base + (uint32_t)(s * P.b_bytes), wsrc + (size_t)s * b_tx, b_tx, bar_wfull);
      }
      asm volatile("griddepcontrol.wait;" ::: "memory");  // activations are the previous kernels' outputs
      // A sequence state
      int a_tile = tile0, a_src = 0, a_c0 = 0, abuf = 0;
      uint32_t aphase = 0;
      // B sequence state
      int b_tile = P.resident ? tend : tile0, b_stage = 0, slot = 0;
      uint32_t sphase = 0;
      long long w_idle = 0, t_idle = dbg ? clock64() : 0;
      while (a_tile < tend || b_tile < tend) {
        if (dbg) t_idle = clock64();
        bool issued = false;
        if (a_tile < tend && mbar_test_wait(bar_aempty(abuf), aphase ^ 1u)) {
          issued = true;
          int t = a_tile / P.n_blocks;
          const int tx0 = (t % P.tiles_x) * S3_TW - c.pad_w;
          t /= P.tiles_x;
          const int ty0 = (t % P.tiles_y) * S3_TH - c.pad_h;
          const int n = t / P.tiles_y;
          if ((P.diag & 256) && a_tile != tile0) {
            mbar_arrive(bar_rawfull(abuf));  // diagnostics: no activation traffic after the first tile (stale operands)
          } else {
            mbar_arrive_expect_tx(bar_rawfull(abuf), a_tx);
            tma_load_4d(smem_base + (uint32_t)(abuf * P.a_bytes), &P.tmap[a_src], bar_rawfull(abuf), a_c0, tx0, ty0, n);
            // L2 prefetch of the same chunk `pf` tiles ahead of this CTA's sequence: bytes in flight beyond what the halo
            // buffers hold (the HBM-bound layers -- skip-operand ResBlock convs, GRU gates, 1x1s -- sat at 3.5-3.9 TB/s with
            // one or two 23 KB chunks outstanding per SM)
            if (P.pf > 0) {
              const int pt = a_tile + P.pf * tstep;
              if (pt < P.ntiles) {
                int t2 = pt / P.n_blocks;
                const int px0 = (t2 % P.tiles_x) * S3_TW - c.pad_w;
                t2 /= P.tiles_x;
                tma_prefetch_4d(&P.tmap[a_src], a_c0, px0, (t2 % P.tiles_y) * S3_TH - c.pad_h, t2 / P.tiles_y);
              }
            }
          }
          if (++abuf == NA) { abuf = 0; aphase ^= 1u; }
          a_c0 += S3_KC;
          if (a_c0 >= c.src[a_src].C) {
            a_c0 = 0;
            if (++a_src == c.nsrc) { a_src = 0; a_tile += tstep; }
          }
        }
        if (b_tile < tend && mbar_test_wait(bar_bfree(slot), sphase ^ 1u)) {
          issued = true;
          const int nb = PAIR ? 0 : b_tile % P.n_blocks;
          const uint32_t b_tx = PAIR ? (uint32_t)P.b_bytes : (uint32_t)n_of(nb) * 128u;  // one stage: 2N rows x 64 bytes (a pair: N rows per CTA)
          const int len = min(P.gtaps, P.stages_per_tile - b_stage);
          // packed weights: [n block][chunk][tap][2*N_block rows][32 fp16]; full blocks hold nb_max channels.  The stages of a
          // group are consecutive in that order, so the group is one contiguous copy.
          const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(c.wpack) + w_rank_off +
                                (size_t)nb * (size_t)P.stages_per_tile * (size_t)P.nb_max * 128u + (size_t)b_stage * b_tx;
          mbar_arrive_expect_tx(bar_bfull(slot), b_tx * (uint32_t)len);
          bulk_load(b_base + (uint32_t)(slot * P.gtaps * P.b_bytes), wsrc, b_tx * (uint32_t)len, bar_bfull(slot));
          if (++slot == NS) { slot = 0; sphase ^= 1u; }
          b_stage += len;
          if (b_stage == P.stages_per_tile) { b_stage = 0; b_tile += tstep; }
        }
        if (dbg && !issued) w_idle += clock64() - t_idle;
      }
      if (dbg) {
        long long* d = P.dbg + (size_t)blockIdx.x * 16;
        d[8] = clock64() - t_begin; d[9] = w_idle;
      }
    }
  }

  // the epilogue warps have drained the last accumulator (hence every MMA has completed) when they arrive here
  tc_fence_before();
  asm volatile("bar.sync 1, %0;" ::"n"(S3_THREADS - 32) : "memory");
  // pair: neither CTA may retire while the other can still arrive on its barriers or its MMAs read this shared memory (the
  // issuer / shadow threads have exited by themselves; the cluster barrier counts the non-exited threads)
  if (PAIR) cluster_sync_all();
  if (warp == S3_CV_WARPS + S3_EPI_WARPS) {
    tc_fence_after();
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// Three translation units (compile time): S3_TU == 1 holds the role-timer (DBG) instantiations of the kernel, S3_TU == 2 the
// lean-epilogue ones, S3_TU == 0 everything else.  __graft_entry__.build() compiles this file three times.
#ifndef S3_TU
#define S3_TU 0
#endif
typedef void (*S3KernelFn)(S3Params);
#define S3_ROW(N_, D_, P_, L_) \
  {conv_s3_kernel<N_, D_, 1, P_, L_>, conv_s3_kernel<N_, D_, 3, P_, L_>, conv_s3_kernel<N_, D_, 5, P_, L_>, conv_s3_kernel<N_, D_, 7, P_, L_>}
S3KernelFn s3_dbg_kernel(int nidx, int uidx, int pair, int lean);
S3KernelFn s3_lean_kernel(int nidx, int uidx, int pair);
S3KernelFn s3_lean2_kernel(int uidx, int pair);  // (64 accumulator channels only)
S3KernelFn s3_lean2_act_kernel(int uidx, int pair, int act);  // activation at compile time: the network's layers; else nullptr
#if S3_TU == 1
S3KernelFn s3_dbg_kernel(int nidx, int uidx, int pair, int lean) {
  static const S3KernelFn table[4][4] = {S3_ROW(32, true, false, 0), S3_ROW(64, true, false, 0), S3_ROW(96, true, false, 0),
                                         S3_ROW(128, true, false, 0)};
  static const S3KernelFn ptable[2][4] = {S3_ROW(32, true, true, 0), S3_ROW(64, true, true, 0)};
  static const S3KernelFn ltable[2][2][4] = {{S3_ROW(32, true, false, 1), S3_ROW(64, true, false, 1)},
                                             {S3_ROW(32, true, true, 1), S3_ROW(64, true, true, 1)}};
  static const S3KernelFn l2table[2][4] = {S3_ROW(64, true, false, 2), S3_ROW(64, true, true, 2)};
  if (lean == 2) return l2table[pair][uidx];
  return lean ? ltable[pair][nidx][uidx] : pair ? ptable[nidx][uidx] : table[nidx][uidx];
}
#elif S3_TU == 2
S3KernelFn s3_lean_kernel(int nidx, int uidx, int pair) {
  static const S3KernelFn ltable[2][2][4] = {{S3_ROW(32, false, false, 1), S3_ROW(64, false, false, 1)},
                                             {S3_ROW(32, false, true, 1), S3_ROW(64, false, true, 1)}};
  return ltable[pair][nidx][uidx];
}
#elif S3_TU == 3
S3KernelFn s3_lean2_kernel(int uidx, int pair) {
  static const S3KernelFn ltable[2][4] = {S3_ROW(64, false, false, 2), S3_ROW(64, false, true, 2)};
  return ltable[pair][uidx];
}
#elif S3_TU == 4
S3KernelFn s3_lean2_act_kernel(int uidx, int pair, int act) {
  // (issue unit, pair, activation) of the layers the engine builds: GRU z / r / q (1x5 and 5x1 on pairs), Ch_Reducer (7x7, tanh),
  // the tanh feature heads of UPNet.2 / UNet dec3 and the FAC-FB encoder output (3x3 on pairs), FGAC's 1x1s (one CTA per tile)
  if (pair && uidx == 2) {
    if (act == DEMFI_ACT_SIGMOID) return conv_s3_kernel<64, false, 5, true, 3 + DEMFI_ACT_SIGMOID>;
    if (act == DEMFI_ACT_SIGMOID_MUL) return conv_s3_kernel<64, false, 5, true, 3 + DEMFI_ACT_SIGMOID_MUL>;
    if (act == DEMFI_ACT_GRU) return conv_s3_kernel<64, false, 5, true, 3 + DEMFI_ACT_GRU>;
  }
  if (pair && uidx == 3 && act == DEMFI_ACT_TANH) return conv_s3_kernel<64, false, 7, true, 3 + DEMFI_ACT_TANH>;
  if (pair && uidx == 1 && act == DEMFI_ACT_TANH) return conv_s3_kernel<64, false, 3, true, 3 + DEMFI_ACT_TANH>;
  if (pair && uidx == 1 && act == DEMFI_ACT_NONE) return conv_s3_kernel<64, false, 3, true, 3 + DEMFI_ACT_NONE>;
  if (!pair && uidx == 0 && act == DEMFI_ACT_NONE) return conv_s3_kernel<64, false, 1, false, 3 + DEMFI_ACT_NONE>;
  return nullptr;
}
#else
// ---- host --------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn s3_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}
static int s3_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}
// N blocking: the layout of the packed weights (shared with conv_h3 for DEMFI_CONV_TC16).  DEMFI_CONV_TC16W (conv_s3 only):
// 97..128 output channels stay ONE block -- an (M 128, N' 256) MMA pair per k-step is bound by the tensor pipe (192 clk for 128
// channels), two (128, 128) pairs by the operand reads from shared memory (2 x 112 clk) -- and the halo tile is fetched once.
int s3_nb_max(int kind, int cout_pad) {
  if (kind == DEMFI_CONV_TC16P) return cout_pad;  // (32 or 64: one block, halved between the CTAs of the pair)
  if (kind == DEMFI_CONV_TC16W && cout_pad > 96 && cout_pad <= 128) return cout_pad;
  return cout_pad <= 96 ? cout_pad : 64;
}
constexpr int S3_SMEM_MAX = 227 * 1024;

// TMA epilogue plan.  Per entry (an N block, or a 32-channel box of it) the segments that intersect it must all be alike (same
// channel range, activation, operands, store mode -- i.e. one result written to one or more destinations).  Pixel-shuffle
// segments need one N block per quadrant (nch / 4 == block width).
struct S3EpiPlan {
  int sbw, nsb, n_ent;
  int e_seg[S3_MAX_E], e_o0[S3_MAX_E], e_on[S3_MAX_E], e_nres[S3_MAX_E], e_mixed[S3_MAX_E], e_c0[S3_MAX_E];
  int ol_seg[S3_MAX_OL], ol_ent[S3_MAX_OL], ol_map[S3_MAX_OL];
  int n_ol;
  int o_seg[S3_MAX_O], o_q[S3_MAX_O];  // destination tensor maps: (segment, pixel-shuffle quadrant or -1)
  int n_o;
  bool any_res2;
};
static bool s3_plan_epilogue_w(const demfi_conv_t& c, int nb_max, int n_blocks, int sbw, S3EpiPlan& E) {
  const int nsb = (nb_max + sbw - 1) / sbw;
  if (n_blocks * nsb > S3_MAX_E) return false;
  E.sbw = sbw;
  E.nsb = nsb;
  E.n_ent = n_blocks * nsb;
  E.n_ol = 0;
  E.n_o = 0;
  E.any_res2 = false;
  for (int e = 0; e < E.n_ent; ++e) {
    const int nb = e / nsb, sb = e % nsb;
    const int nblk1 = nb * nb_max + (c.cout_pad - nb * nb_max < nb_max ? c.cout_pad - nb * nb_max : nb_max);
    const int n0 = nb * nb_max + sb * sbw;
    const int n1 = n0 + sbw < nblk1 ? n0 + sbw : nblk1;
    E.e_seg[e] = -1;
    E.e_o0[e] = E.n_ol;
    E.e_on[e] = 0;
    E.e_nres[e] = 0;
    E.e_mixed[e] = 0;
    E.e_c0[e] = n0;
    if (n0 >= n1) continue;  // (a sub-block beyond the end of a ragged last block)
    for (int s = 0; s < c.nseg; ++s) {
      const demfi_seg_t& g = c.seg[s];
      if (g.ch0 >= n1 || g.ch0 + g.nch <= n0) continue;
      if (E.e_seg[e] < 0) {
        E.e_seg[e] = s;
        E.e_nres[e] = g.act == DEMFI_ACT_GRU ? 2 : (g.res != nullptr ? 1 : 0);
        if (g.act == DEMFI_ACT_GRU) E.any_res2 = true;
      } else {
        const demfi_seg_t& a = c.seg[E.e_seg[e]];
        if (g.ch0 != a.ch0 || g.nch != a.nch || g.act != a.act || g.store != a.store || g.res != a.res || g.res_ld != a.res_ld ||
            g.res2 != a.res2 || g.res2_ld != a.res2_ld || g.fmt != a.fmt)
          return false;
      }
      // 32-channel entries must be whole boxes of ONE segment
      if (nsb > 1 && (g.ch0 % 32 != 0 || g.nch % 32 != 0)) return false;
      // the first operand normally shares the staging tile with the result (updated in place): that needs the same format.
      // A residual in the other format goes to the second tile instead (not available to the two-operand GRU epilogue).
      if (g.res != nullptr && ((g.fmt & DEMFI_SEG_DST_S16) != 0) != ((g.fmt & DEMFI_SEG_RES_S16) != 0)) {
        if (g.act == DEMFI_ACT_GRU) return false;
        E.e_mixed[e] = 1;
        E.any_res2 = true;
      }
      if ((g.fmt & DEMFI_SEG_DST_S16) && (g.ch0 % 32 != 0 || g.nch % 32 != 0 || (n0 - g.ch0) % 32 != 0)) return false;
      int q = -1;
      if (g.store == DEMFI_STORE_PIXEL_SHUFFLE2) {
        const int cq = g.nch / 4;
        if (cq < 1 || n0 < g.ch0 || (n0 - g.ch0) / cq != (n1 - 1 - g.ch0) / cq) return false;  // entry inside one quadrant
        if (g.res != nullptr) return false;
        q = (n0 - g.ch0) / cq;
      }
      int mi = -1;  // destination tensor map of (segment, quadrant)
      for (int j = 0; j < E.n_o; ++j)
        if (E.o_seg[j] == s && E.o_q[j] == q) mi = j;
      if (mi < 0) {
        if (E.n_o == S3_MAX_O) return false;
        mi = E.n_o++;
        E.o_seg[mi] = s;
        E.o_q[mi] = q;
      }
      if (E.n_ol == S3_MAX_OL) return false;
      E.ol_seg[E.n_ol] = s;
      E.ol_ent[E.n_ol] = e;
      E.ol_map[E.n_ol] = mi;
      ++E.n_ol;
      ++E.e_on[e];
    }
  }
  return true;
}
static bool s3_plan_epilogue(const demfi_conv_t& c, int nb_max, int n_blocks, S3EpiPlan& E) {
  if (s3_plan_epilogue_w(c, nb_max, n_blocks, nb_max, E)) return true;
  return nb_max > 32 && s3_plan_epilogue_w(c, nb_max, n_blocks, 32, E);
}

bool s3_s16_ok(const demfi_conv_t& c) {
  const int nbm = s3_nb_max(c.kind, c.cout_pad);
  static thread_local S3EpiPlan E;
  bool any_seg = false;
  for (int s = 0; s < c.nseg; ++s) any_seg = any_seg || c.seg[s].fmt != 0;
  return !any_seg || s3_plan_epilogue(c, nbm, (c.cout_pad + nbm - 1) / nbm, E);
}

// Packed weights of a CTA pair (DEMFI_CONV_TC16P): [rank][chunk][tap][N rows][32 fp16], 64-byte swizzle as in h3_pack_weights;
// rank r holds output channels r N/2 .. (r+1) N/2 - 1: rows 0..N/2-1 = fp16(w), rows N/2..N-1 = fp16((w - hi) * 2048).  The
// main MMA (N' = 2N over the pair) reads all N rows of each CTA, the correction MMA (N' = N) the first N/2 (the hi rows).
int s3_pack_weights_pair(const float* w, int Co, int Ci, int KH, int KW, const int32_t* in_map, const int32_t* src_C, int nsrc,
                         const int32_t* out_map, int cout_pad, float* out) {
  DEMFI_REQUIRE(cout_pad == 32 || cout_pad == 64, "pack_weights: DEMFI_CONV_TC16P needs cout_pad 32 or 64 (got %d)", cout_pad);
  const int taps = KH * KW, N = cout_pad, Nh = N / 2;
  int chunks = 0;
  for (int s_ = 0; s_ < nsrc; ++s_) chunks += (src_C[s_] + S3_KC - 1) / S3_KC;
  __half* o = reinterpret_cast<__half*>(out);
  for (int rank = 0; rank < 2; ++rank) {
    int chunk = 0, kbase = 0;
    for (int s_ = 0; s_ < nsrc; ++s_) {
      for (int c0 = 0; c0 < src_C[s_]; c0 += S3_KC, ++chunk) {
        for (int tap = 0; tap < taps; ++tap) {
          __half* tile = o + (((size_t)rank * chunks + chunk) * taps + tap) * (size_t)(N * S3_KC);
          for (int n = 0; n < Nh; ++n)
            for (int k = 0; k < S3_KC; ++k) {
              float v = 0.0f;
              const int cc = c0 + k;
              if (cc < src_C[s_]) {
                const int ci = in_map[kbase + cc], co = out_map[rank * Nh + n];
                if (ci >= 0 && co >= 0) v = w[((size_t)co * Ci + ci) * taps + tap];
              }
              DEMFI_REQUIRE(v > -65504.0f && v < 65504.0f, "pack_weights: weight %g outside the fp16 range", (double)v);
              const __half h = __float2half_rn(v);
              const __half l = __float2half_rn((v - __half2float(h)) * S3_LO_SCALE);
              const int rh = n, rl = Nh + n;
              tile[(size_t)rh * S3_KC + (size_t)((((k >> 3) ^ ((rh >> 1) & 3)) << 3) + (k & 7))] = h;
              tile[(size_t)rl * S3_KC + (size_t)((((k >> 3) ^ ((rl >> 1) & 3)) << 3) + (k & 7))] = l;
            }
        }
      }
      kbase += src_C[s_];
    }
  }
  return 0;
}

bool s3_supports(const demfi_conv_t& c) {
  if (c.stride != 1) return false;
  if (c.kind == DEMFI_CONV_TC16P && c.cout_pad != 32 && c.cout_pad != 64) return false;
  if (c.cout_pad % 16 != 0 || c.cout_pad < 16 || c.cout_pad > 256) return false;
  for (int s = 0; s < c.nsrc; ++s)
    if (c.src[s].up != 0) return false;
  const int hw = S3_TW + c.KW - 1, hh = S3_TH + c.KH - 1;
  if (hw > 256 || hh > 256) return false;
  const int a_bytes = (hw * hh * 128 + 1023) / 1024 * 1024;
  const int nbm = s3_nb_max(c.kind, c.cout_pad);
  const int b_bytes = (c.kind == DEMFI_CONV_TC16P ? 1 : 2) * nbm * 64;
  const int stg = ((nbm + 31) / 32) * S3_BOX_BYTES;
  return 2 * a_bytes + (nbm > 96 ? 2 : 4) * b_bytes + stg + 2048 <= S3_SMEM_MAX;
}

// Halo-tile buffers wanted.  A 1x1 kernel consumes a chunk in ~400 cycles, so the loads in flight -- not the math -- set its
// pace when the source streams from HBM: measured (tools/layer_table.py) GFF.0 (36 chunks per tile, 1.08 GB) 0.32 -> 0.28 ms and
// the FGAC 1x1s (fp32 sources) 0.31 -> 0.28 ms with 6 buffers; no gain on the LFFs (7 chunks, L2-warm S16 trunk) or the
// 1x5 / 5x1 GRU convolutions, which keep 3 (and their shared memory for the weight ring).
static int s3_want_buffers(int taps, int chunks_per_tile) {
  if (taps == 1 && (chunks_per_tile >= 16 || chunks_per_tile <= 2)) return S3_MAX_NA;
  return 3;
}

static int s3_encode(EncodeTiledFn enc, CUtensorMap* map, const float* ptr, int C_, int ld, int W, int H, int N, int bw, int bh,
                     const char* what) {
  cuuint64_t dims[4] = {(cuuint64_t)C_, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)ld * 4, (cuuint64_t)ld * 4 * W, (cuuint64_t)ld * 4 * W * H};
  cuuint32_t box[4] = {S3_KC, (cuuint32_t)bw, (cuuint32_t)bh, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DEMFI_REQUIRE(r == CUDA_SUCCESS, "conv_s3: cuTensorMapEncodeTiled failed for %s (CUresult %d)", what, (int)r);
  return 0;
}

// Everything of a launch that is pure host arithmetic (tiling, N blocking, epilogue plan, shared-memory plan, accumulation
// segments): no CUDA call, so that demfi_conv_describe() can report it on a machine without a GPU.  Returns the dynamic
// shared-memory size through *smem_out.
static int s3_plan(const demfi_conv_t& c, S3Params& P, S3EpiPlan& E, int* smem_out) {
  DEMFI_REQUIRE(s3_supports(c), "conv_s3: unsupported convolution (stride %d, cout_pad %d, %dx%d)", c.stride, c.cout_pad, c.KH, c.KW);
  DEMFI_REQUIRE(c.Hi + 2 * c.pad_h - c.KH + 1 == c.H && c.Wi + 2 * c.pad_w - c.KW + 1 == c.W, "conv_s3: inconsistent sizes");
  P.c = c;
  P.hw = S3_TW + c.KW - 1;
  P.hh = S3_TH + c.KH - 1;
  P.halo_px = P.hw * P.hh;
  P.tiles_x = (c.W + S3_TW - 1) / S3_TW;
  P.tiles_y = (c.H + S3_TH - 1) / S3_TH;
  const long long nt = (long long)P.tiles_x * P.tiles_y * c.N;
  P.nb_max = s3_nb_max(c.kind, c.cout_pad);
  P.n_blocks = (c.cout_pad + P.nb_max - 1) / P.nb_max;
  DEMFI_REQUIRE(nt > 0 && nt * P.n_blocks < (1ll << 31), "conv_s3: bad tile count");
  P.ntiles = (int)nt * P.n_blocks;
  P.a_bytes = (P.halo_px * 128 + 1023) / 1024 * 1024;
  P.pair = c.kind == DEMFI_CONV_TC16P ? 1 : 0;
  // one weight stage in a CTA's shared memory: [Bh ; Bl] = 2N rows of 64 bytes; a CTA of a pair holds half of the rows
  P.b_bytes = (P.pair ? 1 : 2) * P.nb_max * 64;
  // accumulator buffer: [main N | corrections N]; a pair keeps Al x Bh in columns of its own: [main 2N | Al Bh N]
  P.acc_stride = (P.pair ? 3 : 2) * P.nb_max;
  P.taps = c.KH * c.KW;
  // MMA issue unit: a kernel row (KW > 1), a kernel column (N x 1 kernels) or a single tap; straight-line code exists for 3, 5, 7
  P.unit = c.KW > 1 ? c.KW : c.KH;
  P.ustep = c.KW > 1 ? 8 : P.hw * 8;
  if (P.unit != 3 && P.unit != 5 && P.unit != 7) { P.unit = 1; P.ustep = 8; }
  int chunks = 0;
  for (int s = 0; s < c.nsrc; ++s) chunks += (c.src[s].C + S3_KC - 1) / S3_KC;
  P.stages_per_tile = chunks * P.taps;

  // epilogue mode
  P.tma_epi = (s3_plan_epilogue(c, P.nb_max, P.n_blocks, E) && !(get_option("tc_diag") & 2)) ? 1 : 0;
  for (int sg = 0; sg < c.nseg; ++sg)
    DEMFI_REQUIRE(c.seg[sg].fmt == 0 || P.tma_epi, "conv_s3: segment %d asks for the S16 format, which needs the TMA epilogue", sg);
  const int box_bytes_all = ((P.nb_max + 31) / 32) * S3_BOX_BYTES;
  P.all_s16 = 1;
  for (int s_ = 0; s_ < c.nsrc; ++s_)
    if (c.src[s_].fmt != DEMFI_FMT_S16) P.all_s16 = 0;
  P.offload = (P.tma_epi && P.all_s16 && !(get_option("tc_diag") & 2048)) ? 1 : 0;
  P.lean = 0;
  P.res_sep = 0;
  if (P.tma_epi && P.n_blocks == 1 && E.nsb == 1 && (c.cout_pad == 32 || c.cout_pad == 64) && E.e_seg[0] >= 0 &&
      !(get_option("tc_diag") & 4096)) {
    const demfi_seg_t& g = c.seg[E.e_seg[0]];
    const bool res_ok = E.e_nres[0] == 0 || (E.e_nres[0] == 1 && (g.fmt & DEMFI_SEG_RES_S16));
    if ((g.fmt & DEMFI_SEG_DST_S16) && g.ch0 == 0 && g.nch == c.cout_pad && res_ok && g.store == DEMFI_STORE_NHWC && !E.e_mixed[0] &&
        (g.act == DEMFI_ACT_NONE || g.act == DEMFI_ACT_RELU))
      P.lean = 1 | (g.act == DEMFI_ACT_RELU ? 2 : 0) | (E.e_nres[0] == 1 ? 4 : 0);
    // the activation variant (kernels of their own, LEAN == 2; 64 channels): tanh / sigmoid / sigmoid x h / GRU, S16 operands
    // (any destination / first-operand format: fp32 rows are written / read as plain 16-byte chunks; a first operand in the other
    // format than the destination sits in the second tile, as the planner arranged; the GRU's second operand is S16)
    const bool ops_ok = E.e_nres[0] < 2 || ((g.fmt & DEMFI_SEG_RES2_S16) && (g.fmt & DEMFI_SEG_RES_S16) && (g.fmt & DEMFI_SEG_DST_S16));
    if (P.lean == 0 && c.cout_pad == 64 && g.ch0 == 0 && g.nch == 64 && ops_ok && g.store == DEMFI_STORE_NHWC &&
        (g.act >= DEMFI_ACT_TANH || !(g.fmt & DEMFI_SEG_DST_S16) || E.e_mixed[0]) && !(get_option("tc_diag") & 16384))
      P.lean = 8 | ((g.fmt & DEMFI_SEG_DST_S16) ? 0 : 16) | ((E.e_nres[0] >= 1 && !(g.fmt & DEMFI_SEG_RES_S16)) ? 32 : 0);
    // The skip operand of a lean layer in a tile of its OWN (pairs: the halved filter bank leaves the room): its load for the
    // next tile is then issued as soon as this tile's store loop has read it, not after the result's TMA store has drained
    // the shared staging tile -- measured, the epilogue waited ~2 kclk per tile for a residual fetched that late from HBM.
    if (P.lean != 0 && E.e_nres[0] == 1 && P.offload && P.pair && !(get_option("tc_diag") & 8192)) {
      P.res_sep = 1;
      E.e_mixed[0] = 1;
      E.any_res2 = true;
      // ReLU / none kernels with an S16 operand (the ResBlock conv2): the two tiles take turns, operand and result in place
      if ((P.lean & 1) && (P.lean & 4) && !(get_option("tc_diag") & 32768)) P.res_sep = 2;
    }
  }
  // second staging tile (operands that cannot share the result tile): only as many boxes as its last user needs
  int stg2_boxes = 0;
  if (P.tma_epi && E.any_res2)
    for (int e = 0; e < E.n_ent; ++e)
      if (E.e_seg[e] >= 0 && (E.e_mixed[e] || E.e_nres[e] == 2)) {
        const int last = E.nsb > 1 ? e % E.nsb + 1 : (P.nb_max + 31) / 32;
        if (last > stg2_boxes) stg2_boxes = last;
      }
  const int stg_bytes = P.tma_epi ? box_bytes_all + stg2_boxes * S3_BOX_BYTES : 0;
  P.stg2_off = box_bytes_all;  // second operand tile, relative to the staging tile
  P.sbw = P.nb_max;
  P.nsb = 1;
  if (P.tma_epi) {
    P.sbw = E.sbw;
    P.nsb = E.nsb;
    for (int e = 0; e < E.n_ent; ++e) {
      P.e_seg[e] = (signed char)E.e_seg[e];
      P.e_o0[e] = (signed char)E.e_o0[e];
      P.e_on[e] = (signed char)E.e_on[e];
      P.e_nres[e] = (signed char)E.e_nres[e];
      P.e_roff[e] = E.e_mixed[e] ? box_bytes_all : 0;
      P.e_info[e] = E.e_seg[e] < 0 ? 0 : (c.seg[E.e_seg[e]].fmt & 7) | (c.seg[E.e_seg[e]].act << 3) | (E.e_nres[e] << 6) | (E.e_mixed[e] ? 256 : 0);
      if (E.e_seg[e] >= 0) P.e_rc0[e] = E.e_c0[e] - c.seg[E.e_seg[e]].ch0;
    }
    if (P.res_sep == 2) {  // the operand lands in the tile of its turn (the store warp adds 0 / stg2_off), not in a tile of its own
      P.e_roff[0] = 0;
      P.e_info[0] &= ~256;
    }
    for (int j = 0; j < E.n_ol; ++j) {
      const demfi_seg_t& g = c.seg[E.ol_seg[j]];
      const int n0 = E.e_c0[E.ol_ent[j]];
      P.ol_map[j] = (signed char)E.ol_map[j];
      P.ol_c0[j] = n0 - g.ch0;
      if (g.store == DEMFI_STORE_PIXEL_SHUFFLE2) P.ol_c0[j] -= ((n0 - g.ch0) / (g.nch / 4)) * (g.nch / 4);  // channel inside the quadrant
    }
  }

  // shared-memory plan: [A buffers][weights: resident bank or ring][staging][barriers]
  const int fixed = stg_bytes + 8 * S3_NBARS + 16 + 1024 + 1024;  // + barriers, tensor-memory slot, bias, alignment slack
  const int bank = P.stages_per_tile * P.b_bytes;
  P.resident = (P.n_blocks == 1 && 2 * P.a_bytes + bank + fixed <= S3_SMEM_MAX && !(get_option("tc_diag") & 4)) ? 1 : 0;
  if (P.resident) {
    P.ns = 0;
    P.gtaps = P.stages_per_tile;
    P.na = (3 * P.a_bytes + bank + fixed <= S3_SMEM_MAX) ? 3 : 2;
    int want = s3_want_buffers(P.taps, chunks);
    if (P.pair && want < 5) want = 5;  // (a pair's halved filter bank leaves room: deeper prefetch)
    while (P.na < want && (P.na + 1) * P.a_bytes + bank + fixed <= S3_SMEM_MAX) ++P.na;
  } else {
    // ring of `ns` slots, each a group of `gtaps` consecutive stages (target <= 24 KB per slot, >= 3 slots); groups are whole
    // issue units.  A unit that would need more than 32 KB per slot falls back to tap-by-tap issue.
    P.na = 3;
    if (P.unit * P.b_bytes > 32768) { P.unit = 1; P.ustep = 8; }
    const int U = P.unit;
    int g = 24576 / P.b_bytes / U * U;
    if (g < U) g = U;
    if (g > P.stages_per_tile) g = P.stages_per_tile;
    {
      const int cap = get_option("tc_stages");  // diagnostics: stages per ring slot
      if (cap >= 1 && cap < g) g = (cap + U - 1) / U * U;
    }
    int ns = 4;
    auto fits = [&](int na_, int ns_, int g_) { return na_ * P.a_bytes + ns_ * g_ * P.b_bytes + fixed <= S3_SMEM_MAX; };
    while (!fits(P.na, ns, g)) {
      if (ns > 3) --ns;
      else if (g > 2 * U) g -= U;
      else if (P.na > 2) --P.na;
      else if (ns > 2) --ns;
      else if (g > U) g -= U;
      else break;
    }
    DEMFI_REQUIRE(fits(P.na, ns, g), "conv_s3: shared-memory plan does not fit");
    int want = s3_want_buffers(P.taps, chunks);
    if (P.pair && want < 4) want = 4;
    while (P.na < want && fits(P.na + 1, ns, g)) ++P.na;
    P.ns = ns;
    P.gtaps = g;
  }
  P.b_off = P.na * P.a_bytes;
  P.stg_off = P.b_off + (P.resident ? bank : P.ns * P.gtaps * P.b_bytes);
  P.stg_off = (P.stg_off + 1023) & ~1023;
  P.bar_off = P.stg_off + stg_bytes;
  P.bias_off = (P.bar_off + 8 * S3_NBARS + 16 + 15) & ~15;
  const int smem = P.bias_off + 1024 + 1024;
  DEMFI_REQUIRE(smem <= S3_SMEM_MAX + 1024, "conv_s3: shared-memory plan (%d bytes) does not fit", smem);
  *smem_out = smem;

  P.all_full_chunks = 1;
  for (int s_ = 0; s_ < c.nsrc; ++s_)
    if (c.src[s_].C % S3_KC != 0) P.all_full_chunks = 0;
  {  // accumulation segments: whole issue units, balanced over the tile
    const int units = P.stages_per_tile / P.unit;
    int fu = get_option("tc_flush") / P.unit;
    if (fu < 1) fu = 1;
    if (get_option("tc_flush") <= 0 || fu > units) fu = units;
    const int nseg = (units + fu - 1) / fu;
    P.seg_units = (units + nseg - 1) / nseg;
    P.nseg = (units + P.seg_units - 1) / P.seg_units;
    P.seg_last = units - (P.nseg - 1) * P.seg_units;
    P.flush = P.seg_units * P.unit;
    P.grp_units = P.gtaps / P.unit;
    P.ngrp = (units + P.grp_units - 1) / P.grp_units;
    P.grp_last = units - (P.ngrp - 1) * P.grp_units;
  }
  P.pf = get_option("tc_prefetch");
  P.comp = (float)get_option("tc_comp_milli") * 1e-3f * 5.9604645e-8f;
  P.diag = get_option("tc_diag") & (1 | 16 | 32 | 64 | 128 | 256 | 512 | 1024);
  return 0;
}

// host-only description of what a DEMFI_CONV_TC16 launch on conv_s3 would do (demfi_conv_describe)
int s3_describe(const demfi_conv_t& c, int32_t* info) {
  static thread_local S3Params P;
  memset(&P, 0, sizeof(P));
  S3EpiPlan E;
  int smem = 0;
  if (s3_plan(c, P, E, &smem)) return 1;
  info[1] = P.tma_epi;
  info[2] = P.resident;
  info[3] = P.na;
  info[4] = P.ns;
  info[5] = P.gtaps;
  info[6] = P.n_blocks;
  info[7] = smem;
  info[8] = P.flush;
  info[9] = P.stages_per_tile;
  info[10] = P.nsb;  // epilogue entries per N block (1: the block is one result; else 32-channel boxes planned one by one)
  info[11] = P.nb_max;
  info[12] = P.pair;
  return 0;
}

int launch_conv_s3(const demfi_conv_t& c, cudaStream_t st) {
  EncodeTiledFn enc = s3_encode_fn();
  DEMFI_REQUIRE(enc != nullptr, "conv_s3: cuTensorMapEncodeTiled not available from the driver");
  static thread_local S3Params P;  // CUtensorMap needs 64-byte alignment; thread_local storage gives it
  memset(&P, 0, sizeof(P));
  S3EpiPlan E;
  int smem = 0;
  if (s3_plan(c, P, E, &smem)) return 1;
  for (int s = 0; s < c.nsrc; ++s)
    if (s3_encode(enc, &P.tmap[s], c.src[s].ptr, c.src[s].C, c.src[s].ld, c.Wi, c.Hi, c.N, P.hw, P.hh, "a source")) return 1;
  if (P.tma_epi) {
    bool rdone[DEMFI_MAX_SEG] = {false, false, false, false};
    for (int e = 0; e < E.n_ent; ++e) {
      const int sg = E.e_seg[e];
      if (sg < 0 || rdone[sg]) continue;
      rdone[sg] = true;
      const demfi_seg_t& g = c.seg[sg];
      if (E.e_nres[e] >= 1)
        if (s3_encode(enc, &P.rmap[sg], g.res, g.nch, g.res_ld, c.W, c.H, c.N, S3_TW, S3_TH, "an epilogue operand")) return 1;
      if (E.e_nres[e] >= 2)
        if (s3_encode(enc, &P.r2map[sg], g.res2, g.nch, g.res2_ld, c.W, c.H, c.N, S3_TW, S3_TH, "the second epilogue operand")) return 1;
    }
    for (int j = 0; j < E.n_o; ++j) {
      const demfi_seg_t& g = c.seg[E.o_seg[j]];
      if (g.store == DEMFI_STORE_PIXEL_SHUFFLE2) {
        // nn.PixelShuffle(2): accumulator channel q*cq + ch of pixel (y, x) -> pixel (2y + q/2, 2x + q%2), channel ch: the
        // quadrant is a tensor of its own with doubled pixel strides
        const int cq = g.nch / 4, q = E.o_q[j];
        const float* base = g.dst + ((size_t)(q >> 1) * (size_t)(2 * c.W) + (size_t)(q & 1)) * (size_t)g.dst_ld;
        cuuint64_t dims[4] = {(cuuint64_t)cq, (cuuint64_t)c.W, (cuuint64_t)c.H, (cuuint64_t)c.N};
        cuuint64_t strides[3] = {(cuuint64_t)g.dst_ld * 8, (cuuint64_t)g.dst_ld * 16 * c.W, (cuuint64_t)g.dst_ld * 16 * c.W * c.H};
        cuuint32_t box[4] = {S3_KC, S3_TW, S3_TH, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = enc(&P.omap[j], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        DEMFI_REQUIRE(r == CUDA_SUCCESS, "conv_s3: cuTensorMapEncodeTiled failed for a pixel-shuffle destination (CUresult %d)", (int)r);
      } else {
        if (s3_encode(enc, &P.omap[j], g.dst, g.nch, g.dst_ld, c.W, c.H, c.N, S3_TW, S3_TH, "a destination")) return 1;
      }
    }
  }
  if (P.diag & 128) {
    long long* buf = tc_debug_buffer(st);
    DEMFI_REQUIRE(buf != nullptr, "conv_s3: cannot allocate the role-timer buffer");
    P.dbg = buf;
  }
  // one kernel per (N block width, issue unit): a single instantiation of the issue loop per kernel keeps its state in
  // uniform registers (a switch over four inlined copies did not)
  static const S3KernelFn table[4][4] = {S3_ROW(32, false, false, 0), S3_ROW(64, false, false, 0), S3_ROW(96, false, false, 0),
                                         S3_ROW(128, false, false, 0)};
  static const S3KernelFn ptable[2][4] = {S3_ROW(32, false, true, 0), S3_ROW(64, false, true, 0)};
  const int nidx = P.nb_max <= 32 ? 0 : P.nb_max <= 64 ? 1 : P.nb_max <= 96 ? 2 : 3;
  S3KernelFn fn_act = nullptr;
  if ((P.lean & 8) && P.dbg == nullptr && !(get_option("tc_diag") & 65536))
    fn_act = s3_lean2_act_kernel(P.unit >> 1, P.pair, c.seg[E.e_seg[0]].act);
  const S3KernelFn fn = P.dbg != nullptr ? s3_dbg_kernel(nidx, P.unit >> 1, P.pair, (P.lean & 8) ? 2 : P.lean != 0)
                        : fn_act != nullptr ? fn_act
                        : (P.lean & 8)   ? s3_lean2_kernel(P.unit >> 1, P.pair)
                        : P.lean         ? s3_lean_kernel(nidx, P.unit >> 1, P.pair)
                        : P.pair         ? ptable[nidx][P.unit >> 1]
                                         : table[nidx][P.unit >> 1];
  {
    // cudaFuncSetAttribute applies per device: remember which devices have seen which kernel
    static std::mutex mu;
    static std::set<std::pair<int, const void*>> done;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(mu);
    if (!done.count({dev, (const void*)fn})) {
      cudaError_t e = cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, S3_SMEM_MAX);
      DEMFI_REQUIRE(e == cudaSuccess, "conv_s3: cannot raise dynamic shared memory: %s", cudaGetErrorString(e));
      done.insert({dev, (const void*)fn});
    }
  }
  int grid = P.ntiles < s3_num_sms() ? P.ntiles : s3_num_sms();
  if (get_option("tc_grid") > 0 && get_option("tc_grid") < grid) grid = get_option("tc_grid");
  if (P.pair) {  // whole clusters of two CTAs
    const int pairs = (P.ntiles + 1) / 2;
    grid = 2 * (pairs < s3_num_sms() / 2 ? pairs : s3_num_sms() / 2);
    if (get_option("tc_grid") > 1 && (get_option("tc_grid") & ~1) < grid) grid = get_option("tc_grid") & ~1;
  }
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(S3_THREADS);
    cfg.dynamicSmemBytes = (size_t)smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = get_option("tc_pdl") ? 1 : 0;
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = 2;
    attr[1].val.clusterDim.y = 1;
    attr[1].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = P.pair ? 2 : 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, fn, P);
    DEMFI_REQUIRE(e == cudaSuccess, "conv_s3 launch failed: %s", cudaGetErrorString(e));
  }
  DEMFI_LAUNCH_CHECK("conv_s3");
  return 0;
}
#endif  // S3_TU

}  // namespace demfi
