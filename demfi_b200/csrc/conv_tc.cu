// tcgen05 implicit-GEMM convolution for sm_100a: NHWC fp32 activations, fp32-parity via 3xTF32.
//
// GEMM view of a stride-1 "same" convolution: M = output pixels, N = output
// channels (cout_pad, <= 256), K = taps x input channels.  One CTA tile is 8 rows x 16 columns of
// output pixels (M = 128).  K is consumed in stages of (one tap, 32 input channels):
//   * A stage  : a TMA 4-D box {32 ch, 16 x, 8 y, 1 n} of the NHWC source at the tap's offset,
//                landed 128-byte-swizzled = the canonical K-major UMMA layout (row = pixel).
//                Out-of-image coordinates and channels past the source's C are zero-filled by
//                TMA: that IS the conv's zero padding and the ragged-K handling.
//   * B stage  : the pre-packed [cout_pad x 32] weight tile, tf32 "hi" part and fp32 residual
//                "lo" part, pre-swizzled on the host, fetched with one cp.async.bulk.
//   * split    : four warps write lo = a - (a & ~0x1fff) to a second buffer; the raw fp32 tile serves
//                as the "hi" operand because the tensor core ignores the low 13 mantissa bits
//                (verified on B200 by tests/test_conv_gpu.py::test_tc_operand_truncation_probe).
//   * MMA      : per 8-wide k-step two instructions: Ahi x [Bhi;Blo] (N' = 2N: hi*hi into the main
//                accumulator, hi*lo into the correction accumulator next to it) and Alo x Bhi into the
//                correction accumulator.  Main and correction are summed in fp32 (RN) by the epilogue:
//                the tensor core accumulates with truncation, so keeping the small terms out of the main
//                chain cuts its length 3x (measured: error grows linearly with the chain length).
// Persistent CTAs (one per SM) walk the tile list; warp roles:
//   warps 0-3 split A | warps 4-7 and 10-13 epilogue (TMEM -> regs -> fused epilogue -> HBM)
//   warp 8 TMA producer | warp 9 TMEM allocator + tcgen05.mma issuer (one elected lane)
// The accumulator pair is double-buffered in TMEM so the epilogue of tile i overlaps the MMAs of tile i+1.
// cout_pad > 128 is processed as N blocks of <= 128 channels (a tile = pixel tile x N block).
#include <cuda.h>

#include <cstring>
#include <vector>

#include "common.cuh"

namespace demfi {

constexpr int TC_TH = 8, TC_TW = 16, TC_BM = 128, TC_KC = 32;
constexpr int TC_A_BYTES = TC_BM * TC_KC * 4;  // 16 KiB
constexpr int TC_THREADS = 448;  // 14 warps: 0-3 split, 4-7 + 10-13 epilogue, 8 TMA, 9 MMA
constexpr int TC_MAX_STAGES = 6;
constexpr int TC_SMEM_BUDGET = 200 * 1024;

struct TcParams {
  CUtensorMap tmap[DEMFI_MAX_SRC];
  demfi_conv_t c;
  int tiles_x, tiles_y, ntiles;
  int n_blocks, nb_max;  // N blocking: blocks of nb_max (<=128) channels, the last one may be smaller
  int stages, buf_stride, tmem_cols;
  int mask_hi, split;
  int flush;         // K stages accumulated inside the tensor core before the partial sum is drained to registers
  int stages_per_tile;
  int diag;              // diagnostics bitmask (tc_diag): 1 no global stores, 2 interleaved (not grouped) MMA order,
                         // 4 separate lo*hi accumulator, 8/16 skip A/B loads, 32 free-running issuer (timing only)
  float comp;            // per-MMA gain correction of the truncating tensor-core accumulation (0 = off)
  int a_stages, a_base;  // A-in-TMEM variant: ring of a_stages x 64 TMEM columns (hi 32 | lo 32) starting at a_base
  long long* dbg;        // tc_diag & 128: per-CTA role timers [gridDim.x][16] (cycles), see demfi_tc_debug_read
};

// ---- PTX wrappers -----------------------------------------------------------------------
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
// Bounded wait: a protocol bug becomes a trapped launch (reported error), never a hung GPU.
static __device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity) {
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 8000000000ll) {
      printf("demfi conv_tc: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", (int)blockIdx.x,
             (int)threadIdx.x, bar, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (!mbar_try_wait(bar, parity)) mbar_wait_slow(bar, parity);
}
// mbarrier wait that adds the cycles it took to *acc when role timers are on (tc_diag & 128)
__device__ __forceinline__ void mbar_wait_t(uint32_t bar, uint32_t parity, bool on, long long& acc) {
  if (!on) { mbar_wait(bar, parity); return; }
  const long long t = clock64();
  mbar_wait(bar, parity);
  acc += clock64() - t;
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// One elected lane of a converged warp.  With elect.sync the compiler knows a single lane is active and that the
// surrounding values are warp-uniform, so tcgen05.mma / TMA operands stay in uniform registers; a plain
// `lane == 0` test made it wrap every UTCHMMA in a 6x R2UR "waterfall" loop (~230 cycles per MMA, measured).
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
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
// A operand from tensor memory (lane = pixel row, one tf32 per 32-bit column), B from shared memory
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, 128-byte-swizzle shared-memory matrix descriptor (sm_100 format): start address >> 4 in
// bits [0,14), SBO (1024 B between 8-row groups) >> 4 in [32,46), version 1 in [46,48), layout
// SWIZZLE_128B (2) in [61,64).  LBO is unused for a single 128-byte swizzle atom along K.
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

// ---- kernel -----------------------------------------------------------------------------
// ATMEM = true: the A operand (activations) goes through tensor memory: the splitter warps read the TMA-landed
// tile once (each thread its own pixel row), split in registers and tcgen05.st hi | lo into a TMEM ring; the
// MMAs take A from TMEM and only B (weights) from shared memory.  Shared-memory traffic per K stage drops from
// ~136 KB (TMA A+B, split read + 2 writes, MMA A hi+hi+lo and B reads) to ~72 KB (TMA A+B, split read, MMA B).
template <int NMAX, bool ATMEM>
__global__ void __launch_bounds__(TC_THREADS, 1) conv_tc_kernel(const __grid_constant__ TcParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const demfi_conv_t& c = P.c;
  constexpr uint32_t A_SLOTS = ATMEM ? 1u : 2u;  // shared-memory A tiles per stage (raw | lo)
  const uint32_t stage_bytes = A_SLOTS * TC_A_BYTES + 2u * (uint32_t)P.nb_max * 128u;
  auto n_of = [&](int nb) { return min(P.nb_max, c.cout_pad - nb * P.nb_max); };
  const int S = P.stages;
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bars = smem_base + (uint32_t)S * stage_bytes;  // 8-byte aligned (stage_bytes % 1024 == 0)
  auto bar_full = [&](int s) { return bars + 8u * (uint32_t)s; };
  auto bar_split = [&](int s) { return bars + 8u * (uint32_t)(S + s); };
  auto bar_empty = [&](int s) { return bars + 8u * (uint32_t)(2 * S + s); };
  auto bar_tfull = [&](int a) { return bars + 8u * (uint32_t)(3 * S + a); };
  auto bar_tempty = [&](int a) { return bars + 8u * (uint32_t)(3 * S + 2 + a); };
  auto bar_aempty = [&](int a) { return bars + 8u * (uint32_t)(3 * S + 4 + a); };  // ATMEM: TMEM A slot free
  auto bar_aready = [&](int a) { return bars + 8u * (uint32_t)(S + a); };          // ATMEM: reuses the split[] slots
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + (size_t)S * stage_bytes + 8 * (3 * S + 8));

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // provably warp-uniform
  const int lane = threadIdx.x & 31;
  const int taps = c.KH * c.KW;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_split(s), 128);
      mbar_init(bar_empty(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull(a), 1);
      mbar_init(bar_tempty(a), 256);
    }
    for (int a = 0; a < 4; ++a) mbar_init(bar_aempty(a), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 9) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)P.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // ===== A splitter =====
    if constexpr (ATMEM) {
      // thread m owns pixel row m of the tile: 128 contiguous bytes whose 16-byte groups sit at (g ^ (m & 7))
      const int m = threadIdx.x;
      const uint32_t lane_addr = ((uint32_t)((warp & 3) * 32) << 16);
      int stage = 0, sa = 0;
      uint32_t phase = 0, aphase = 0;
      const bool dbg = P.dbg != nullptr;
      long long w_full = 0, w_aempty = 0, w_st = 0;
      const long long t_begin = dbg ? clock64() : 0;
      for (int tile = blockIdx.x; tile < P.ntiles && !(P.diag & 32); tile += gridDim.x) {
        for (int st_ = 0; st_ < P.stages_per_tile; ++st_) {
          mbar_wait_t(bar_full(stage), phase, dbg, w_full);
          const uint32_t row = smem_base + (uint32_t)stage * stage_bytes + (uint32_t)m * 128u;
          uint32_t hi[32], lo[32];
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const uint4 v = lds128(row + (uint32_t)((g ^ (m & 7)) << 4));
            hi[4 * g + 0] = v.x; hi[4 * g + 1] = v.y; hi[4 * g + 2] = v.z; hi[4 * g + 3] = v.w;
          }
          if (P.split == 3) {
            const uint32_t rnd = P.mask_hi ? 0x1000u : 0u;  // tf32 round-to-nearest vs truncation
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const uint32_t h = (hi[j] + rnd) & 0xffffe000u;
              lo[j] = __float_as_uint(__uint_as_float(hi[j]) - __uint_as_float(h));
              hi[j] = h;
            }
          }
          mbar_wait_t(bar_aempty(sa), aphase ^ 1, dbg, w_aempty);
          tc_fence_after();
          const uint32_t ta = tmem_base + lane_addr + (uint32_t)(P.a_base + sa * 64);
          const long long t_st = dbg ? clock64() : 0;
          tmem_st32(ta, hi);
          if (P.split == 3) tmem_st32(ta + 32u, lo);
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(bar_aready(sa));
          if (dbg) w_st += clock64() - t_st;
          if (++stage == S) { stage = 0; phase ^= 1; }
          if (++sa == P.a_stages) { sa = 0; aphase ^= 1; }
        }
      }
      if (dbg && threadIdx.x == 0) {
        long long* d = P.dbg + (size_t)blockIdx.x * 16;
        d[0] = clock64() - t_begin; d[1] = w_full; d[2] = w_aempty; d[3] = w_st;
      }
    } else {
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x) {
      for (int s = 0; s < c.nsrc; ++s)
        for (int c0 = 0; c0 < c.src[s].C; c0 += TC_KC)
          for (int tap = 0; tap < taps; ++tap) {
            mbar_wait(bar_full(stage), phase);
            if (P.split == 3) {
              float4* a = reinterpret_cast<float4*>(smem + (size_t)stage * stage_bytes);
              float4* lo = reinterpret_cast<float4*>(smem + (size_t)stage * stage_bytes + TC_A_BYTES);
#pragma unroll
              for (int j = 0; j < TC_A_BYTES / 16 / 128; ++j) {
                const int i = threadIdx.x + j * 128;
                const float4 v = a[i];
                float4 h;
                // mask_hi: hi = tf32 round-to-nearest of a (written back in place; lo gets a random sign, so
                // the tensor core's truncation of lo is unbiased).  otherwise hi = truncation = what the tensor
                // core sees in the raw tile, nothing to write back.
                const uint32_t rnd = P.mask_hi ? 0x1000u : 0u;
                h.x = __uint_as_float((__float_as_uint(v.x) + rnd) & 0xffffe000u);
                h.y = __uint_as_float((__float_as_uint(v.y) + rnd) & 0xffffe000u);
                h.z = __uint_as_float((__float_as_uint(v.z) + rnd) & 0xffffe000u);
                h.w = __uint_as_float((__float_as_uint(v.w) + rnd) & 0xffffe000u);
                lo[i] = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
                if (P.mask_hi) a[i] = h;
              }
              fence_async_smem();
            }
            mbar_arrive(bar_split(stage));
            if (++stage == S) { stage = 0; phase ^= 1; }
          }
    }
    }
  } else if (warp < 8 || warp >= 10) {
    // ===== epilogue: 8 warps.  TMEM lane = pixel row of the tile; a warp can only touch lanes 32*(warp%4)...
    // Group 0 (warps 4-7) takes the accumulator columns [0, csplit), group 1 (warps 10-13) takes [csplit, N):
    // the epilogue was the critical role in the ncu source profile (95 % busy), so its work is halved per warp,
    // TMEM loads are issued in batches (one wait per batch) and the ReLU / linear NHWC case is inlined. =====
    const int grp = warp >= 10 ? 1 : 0;
    const int wq = warp & 3;
    const int m = wq * 32 + lane;
    constexpr int HMAX = (NMAX / 2 + 15) / 16 * 16;  // columns one group can own
    int acc = 0;
    uint32_t acc_phase = 0;
    const bool dbg = P.dbg != nullptr;
    long long w_tfull = 0, w_store = 0;
    const long long t_begin = dbg ? clock64() : 0;
    for (int tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x) {
      int t = tile;
      const int nb = t % P.n_blocks;
      t /= P.n_blocks;
      const int tx0 = (t % P.tiles_x) * TC_TW;
      t /= P.tiles_x;
      const int ty0 = (t % P.tiles_y) * TC_TH;
      const int n = t / P.tiles_y;
      const int N = n_of(nb), n0 = nb * P.nb_max;
      const int csplit = ((N / 2 + 15) / 16) * 16;
      const int cbeg = grp ? csplit : 0, cnum = grp ? N - csplit : csplit;  // this group's columns
      const int oy = ty0 + (m >> 4), ox = tx0 + (m & 15);
      const bool valid = (oy < c.H) && (ox < c.W);
      // The tensor core accumulates with truncation (a biased error that compounds over ~100 layers), so
      // the K loop is cut into segments of P.flush stages: each segment's partial sum is drained from
      // TMEM and added here in fp32 round-to-nearest while the MMAs of the next segment run.
      float sum[HMAX];
      bool first = true;
      for (int done = 0; done < P.stages_per_tile; done += P.flush) {
        const float gain = 1.0f + P.comp * (float)(4 * min(P.flush, P.stages_per_tile - done));
        mbar_wait_t(bar_tfull(acc), acc_phase, dbg, w_tfull);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(acc * P.buf_stride) + (uint32_t)cbeg;
        uint32_t r[HMAX];
#pragma unroll
        for (int col = 0; col < HMAX; col += 16)
          if (col < cnum) tmem_ld16_nowait(taddr + (uint32_t)col, r + col);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < HMAX; ++j) {
          const float v = __uint_as_float(r[j]) * gain;
          sum[j] = first ? v : sum[j] + v;
        }
        if (P.split == 3) {
#pragma unroll
          for (int col = 0; col < HMAX; col += 16)
            if (col < cnum) tmem_ld16_nowait(taddr + (uint32_t)(N + col), r + col);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < HMAX; ++j) sum[j] += __uint_as_float(r[j]);
          if (P.diag & 4) {
#pragma unroll
            for (int col = 0; col < HMAX; col += 16)
              if (col < cnum) tmem_ld16_nowait(taddr + (uint32_t)(2 * N + col), r + col);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < HMAX; ++j) sum[j] += __uint_as_float(r[j]);
          }
        }
        tc_fence_before();
        mbar_arrive(bar_tempty(acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        first = false;
      }
      const long long t_store = dbg ? clock64() : 0;
      if (valid && !(P.diag & 1)) {
        const int ch_lo = n0 + cbeg, ch_hi = ch_lo + cnum;  // absolute accumulator channels held in sum[]
#pragma unroll 1
        for (int sgi = 0; sgi < c.nseg; ++sgi) {
          const demfi_seg_t& sg = c.seg[sgi];
          if (sg.ch0 >= ch_hi || sg.ch0 + sg.nch <= ch_lo) continue;
          const SegCursor cur = seg_cursor(c, sg, n, oy, ox);
          if (sg.store == DEMFI_STORE_NHWC && (sg.act == DEMFI_ACT_RELU || sg.act == DEMFI_ACT_NONE)) {
            // fast path (ResBlocks, RDB, Mixer ...): bias (+ residual) (+ ReLU), 128-bit store, no call
            const float lo_clamp = sg.act == DEMFI_ACT_RELU ? 0.0f : -INFINITY;
#pragma unroll
            for (int col = 0; col < HMAX; col += 4) {
              const int co = ch_lo + col;
              if (col < cnum && co >= cur.lo && co < cur.hi) {
                const float4 b = ld4(c.bias + co);
                float4 v = make_float4(sum[col] + b.x, sum[col + 1] + b.y, sum[col + 2] + b.z, sum[col + 3] + b.w);
                if (cur.res != nullptr) {
                  const float4 rr = ld4(cur.res + (co - cur.lo));
                  v.x += rr.x; v.y += rr.y; v.z += rr.z; v.w += rr.w;
                }
                v.x = fmaxf(v.x, lo_clamp); v.y = fmaxf(v.y, lo_clamp); v.z = fmaxf(v.z, lo_clamp); v.w = fmaxf(v.w, lo_clamp);
                st4(cur.dst + (co - cur.lo), v);
              }
            }
          } else {
#pragma unroll
            for (int col = 0; col < HMAX; col += 4) {
              if (col < cnum) {
                const float4 b = ld4(c.bias + ch_lo + col);
                seg_emit4(cur, ch_lo + col, make_float4(sum[col] + b.x, sum[col + 1] + b.y, sum[col + 2] + b.z, sum[col + 3] + b.w));
              }
            }
          }
        }
      }
      if (dbg) w_store += clock64() - t_store;
    }
    if (dbg && warp == 4 && lane == 0) {
      long long* d = P.dbg + (size_t)blockIdx.x * 16;
      d[4] = clock64() - t_begin; d[5] = w_tfull; d[6] = w_store;
    }
  } else if (warp == 8) {
    // ===== TMA producer (one elected lane).  Per-stage work is kept minimal: no divisions, incremental pointers. =====
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      const bool do_a = !(P.diag & 8), do_b = !(P.diag & 16);  // timing diagnostics only (results are garbage)
      const size_t w_step = 2 * (size_t)c.cout_pad * TC_KC;   // floats per (chunk, tap) in the packed weights
      const bool dbg = P.dbg != nullptr;
      long long w_empty = 0;
      const long long t_begin = dbg ? clock64() : 0;
      for (int tile = blockIdx.x; tile < P.ntiles && !(P.diag & 32); tile += gridDim.x) {
        int t = tile;
        const int nb = t % P.n_blocks;
        t /= P.n_blocks;
        const int tx0 = (t % P.tiles_x) * TC_TW - c.pad_w;
        t /= P.tiles_x;
        const int ty0 = (t % P.tiles_y) * TC_TH - c.pad_h;
        const int n = t / P.tiles_y;
        const int N = n_of(nb);
        const uint32_t b_bytes = (uint32_t)N * 128u;
        const uint32_t tx_bytes = (do_a ? (uint32_t)TC_A_BYTES : 0u) + (do_b ? 2u * b_bytes : 0u);
        const float* wsrc = c.wpack + (size_t)(nb * P.nb_max) * TC_KC;
        const size_t lo_off = (size_t)c.cout_pad * TC_KC;
        for (int s = 0; s < c.nsrc; ++s) {
          const CUtensorMap* map = &P.tmap[s];
          const int Cs = c.src[s].C;
          for (int c0 = 0; c0 < Cs; c0 += TC_KC)
            for (int ky = 0; ky < c.KH; ++ky)
              for (int kx = 0; kx < c.KW; ++kx) {
                mbar_wait_t(bar_empty(stage), phase ^ 1, dbg, w_empty);
                const uint32_t sa = smem_base + (uint32_t)stage * stage_bytes;
                const uint32_t fb = bar_full(stage);
                mbar_arrive_expect_tx(fb, tx_bytes);
                if (do_a) tma_load_4d(sa, map, fb, c0, tx0 + kx, ty0 + ky, n);
                if (do_b) {  // [Bhi rows | Blo rows] land contiguously: one 2N-row K-major tile
                  bulk_load(sa + A_SLOTS * TC_A_BYTES, wsrc, b_bytes, fb);
                  bulk_load(sa + A_SLOTS * TC_A_BYTES + b_bytes, wsrc + lo_off, b_bytes, fb);
                }
                wsrc += w_step;
                if (++stage == S) { stage = 0; phase ^= 1; }
              }
        }
      }
      if (dbg) {
        long long* d = P.dbg + (size_t)blockIdx.x * 16;
        d[8] = clock64() - t_begin; d[9] = w_empty;
      }
    }
  } else {
    // ===== MMA issuer (warp 9: the whole warp walks the flattened stage loop, one elected lane issues).
    // The issuing warp's own instruction stream was the bottleneck (ncu + free-running experiments), so the
    // per-stage code is a barrier probe, a handful of uniform adds and the tcgen05 instructions. =====
    const uint32_t idesc0 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_BM >> 4) << 24);  // D=f32, A=B=tf32, K-major, M=128
    const uint64_t bdesc0 = make_desc_sw128(smem_base + A_SLOTS * TC_A_BYTES);  // B tile of stage 0
    const uint64_t adesc0 = make_desc_sw128(smem_base);                         // SS: raw/hi A tile of stage 0
    const uint64_t dstep = (uint64_t)(stage_bytes >> 4);                        // descriptor start-address step per stage
    const bool free_run = (P.diag & 32) != 0, keep_commits = (P.diag & 64) != 0;
    uint32_t stage = 0, phase = 0, sa_ = 0, aphase = 0, acc = 0, acc_phase = 0;
    const bool dbg = P.dbg != nullptr;
    long long w_tempty = 0, w_aready = 0;
    const long long t_begin = dbg ? clock64() : 0;
    for (int tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x) {
      const int N = n_of(tile % P.n_blocks);
      const uint32_t idesc_n = idesc0 | ((uint32_t)(N >> 3) << 17);         // N columns
      const uint32_t idesc_2n = idesc0 | ((uint32_t)((2 * N) >> 3) << 17);  // [hi;lo] weight tile: 2N columns
      for (int done = 0; done < P.stages_per_tile; done += P.flush) {
        const int seg_len = min(P.flush, P.stages_per_tile - done);
        mbar_wait_t(bar_tempty(acc), acc_phase ^ 1, dbg, w_tempty);  // fresh accumulator pair for this segment
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * (uint32_t)P.buf_stride;  // main [0,N), correction [N,2N)
        const uint32_t d_lohi = d_tmem + (uint32_t)((P.diag & 4) ? 2 * N : N);
        for (int i = 0; i < seg_len; ++i) {
          const uint64_t b_hi = bdesc0 + dstep * stage;
          const uint32_t first = (i == 0) ? 0u : 1u;
          if constexpr (ATMEM) {
            if (!free_run) {
              mbar_wait_t(bar_aready(sa_), aphase, dbg, w_aready);  // splitter stored A (it waited on full[stage]: B has landed too)
              tc_fence_after();
            }
            const uint32_t ta = tmem_base + (uint32_t)P.a_base + sa_ * 64u;
            if (elect_one()) {
              if (P.split == 3 && !(P.diag & 2)) {  // grouped: 4x (Ahi x [Bhi;Blo]) then 4x (Alo x Bhi)
#pragma unroll
                for (int k = 0; k < TC_KC / 8; ++k)
                  umma_tf32_ts(d_tmem, ta + (uint32_t)(k * 8), b_hi + (uint64_t)(k * 2), idesc_2n, k == 0 ? first : 1u);
#pragma unroll
                for (int k = 0; k < TC_KC / 8; ++k)
                  umma_tf32_ts(d_lohi, ta + 32u + (uint32_t)(k * 8), b_hi + (uint64_t)(k * 2), idesc_n,
                               (k == 0 && (P.diag & 4)) ? first : 1u);
              } else {
#pragma unroll
                for (int k = 0; k < TC_KC / 8; ++k) {
                  const uint64_t kk = (uint64_t)(k * 2);
                  if (P.split == 3) {
                    umma_tf32_ts(d_tmem, ta + (uint32_t)(k * 8), b_hi + kk, idesc_2n, k == 0 ? first : 1u);
                    umma_tf32_ts(d_lohi, ta + 32u + (uint32_t)(k * 8), b_hi + kk, idesc_n, (k == 0 && (P.diag & 4)) ? first : 1u);
                  } else {
                    umma_tf32_ts(d_tmem, ta + (uint32_t)(k * 8), b_hi + kk, idesc_n, k == 0 ? first : 1u);
                  }
                }
              }
              if (!free_run || keep_commits) {
                umma_commit(bar_aempty(sa_));
                umma_commit(bar_empty(stage));
              }
              if (i == seg_len - 1) umma_commit(bar_tfull(acc));
            }
            __syncwarp();
            if (++sa_ == (uint32_t)P.a_stages) { sa_ = 0; aphase ^= 1; }
          } else {
            mbar_wait(bar_split(stage), phase);
            tc_fence_after();
            const uint64_t a_hi = adesc0 + dstep * stage, a_lo = a_hi + (uint64_t)(TC_A_BYTES >> 4);
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < TC_KC / 8; ++k) {
                const uint64_t kk = (uint64_t)(k * 2);  // 32 bytes >> 4 per k-step of 8 tf32
                if (P.split == 3) {
                  umma_tf32(d_tmem, a_hi + kk, b_hi + kk, idesc_2n, k == 0 ? first : 1u);  // hi*hi | hi*lo
                  umma_tf32(d_tmem + (uint32_t)N, a_lo + kk, b_hi + kk, idesc_n, 1);       // lo*hi -> correction
                } else {
                  umma_tf32(d_tmem, a_hi + kk, b_hi + kk, idesc_n, k == 0 ? first : 1u);
                }
              }
              umma_commit(bar_empty(stage));
              if (i == seg_len - 1) umma_commit(bar_tfull(acc));
            }
            __syncwarp();
          }
          if (++stage == (uint32_t)S) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
    if (dbg && lane == 0) {
      long long* d = P.dbg + (size_t)blockIdx.x * 16;
      d[12] = clock64() - t_begin; d[13] = w_tempty; d[14] = w_aready;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)P.tmem_cols)
                 : "memory");
  }
}

// ---- host --------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
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

constexpr int TC_DBG_CTAS = 256;
static long long* g_tc_dbg = nullptr;
// role timers of the last conv_tc launch made with tc_diag & 128: [ctas][16] cycles
//   0 splitter total, 1 wait full, 2 wait A slot free, 3 tcgen05.st + arrive | 4 epilogue total, 5 wait accumulator, 6 store phase
//   8 producer total, 9 wait stage free | 12 issuer total, 13 wait accumulator free, 14 wait A ready
int tc_debug_read(long long* host, int ctas) {
  DEMFI_REQUIRE(g_tc_dbg != nullptr, "tc_debug_read: no launch was made with tc_diag & 128");
  DEMFI_REQUIRE(ctas > 0 && ctas <= TC_DBG_CTAS, "tc_debug_read: ctas out of range");
  DEMFI_REQUIRE(cudaMemcpy(host, g_tc_dbg, (size_t)ctas * 16 * sizeof(long long), cudaMemcpyDeviceToHost) == cudaSuccess, "tc_debug_read: copy failed");
  return 0;
}

long long* tc_debug_buffer(cudaStream_t st) {
  if (g_tc_dbg == nullptr && cudaMalloc(&g_tc_dbg, TC_DBG_CTAS * 16 * sizeof(long long)) != cudaSuccess) return nullptr;
  cudaMemsetAsync(g_tc_dbg, 0, TC_DBG_CTAS * 16 * sizeof(long long), st);
  return g_tc_dbg;
}

static int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

int launch_conv_tc(const demfi_conv_t& c, cudaStream_t st) {
  DEMFI_REQUIRE(c.stride == 1 && c.Hi == c.H && c.Wi == c.W, "conv_tc: only stride-1 'same' convolutions");
  DEMFI_REQUIRE(c.cout_pad % 16 == 0 && c.cout_pad >= 16 && c.cout_pad <= 256, "conv_tc: cout_pad %d not in 16..256 step 16", c.cout_pad);
  EncodeTiledFn enc = get_encode_fn();
  DEMFI_REQUIRE(enc != nullptr, "conv_tc: cuTensorMapEncodeTiled not available from the driver");
  static thread_local TcParams P;  // CUtensorMap needs 64-byte alignment; thread_local storage gives it
  memset(&P, 0, sizeof(P));
  P.c = c;
  for (int s = 0; s < c.nsrc; ++s) {
    const demfi_src_t& S = c.src[s];
    DEMFI_REQUIRE(S.up == 0, "conv_tc: up-sampled sources are not supported");
    cuuint64_t dims[4] = {(cuuint64_t)S.C, (cuuint64_t)c.W, (cuuint64_t)c.H, (cuuint64_t)c.N};
    cuuint64_t strides[3] = {(cuuint64_t)S.ld * 4, (cuuint64_t)S.ld * 4 * c.W, (cuuint64_t)S.ld * 4 * c.W * c.H};
    cuuint32_t box[4] = {TC_KC, TC_TW, TC_TH, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&P.tmap[s], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(S.ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DEMFI_REQUIRE(r == CUDA_SUCCESS, "conv_tc: cuTensorMapEncodeTiled failed for source %d (CUresult %d)", s, (int)r);
  }
  P.tiles_x = (c.W + TC_TW - 1) / TC_TW;
  P.tiles_y = (c.H + TC_TH - 1) / TC_TH;
  const long long nt = (long long)P.tiles_x * P.tiles_y * c.N;
  const bool atmem = get_option("tc_a_tmem") != 0;
  if (atmem) P.nb_max = c.cout_pad <= 96 ? c.cout_pad : 64;  // 2 accumulator pairs + the A ring must fit 512 TMEM columns
  else P.nb_max = c.cout_pad < 128 ? c.cout_pad : 128;
  P.n_blocks = (c.cout_pad + P.nb_max - 1) / P.nb_max;
  DEMFI_REQUIRE(nt > 0 && nt * P.n_blocks < (1ll << 31), "conv_tc: bad tile count");
  P.ntiles = (int)nt * P.n_blocks;
  const int stage_bytes = (atmem ? 1 : 2) * TC_A_BYTES + 2 * P.nb_max * 128;
  int stages = TC_SMEM_BUDGET / stage_bytes;
  if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
  {
    int cap = get_option("tc_stages");
    if (atmem && cap > 0 && cap < 4) cap = 4;  // the TMEM A ring variant is only validated with >= 4 smem stages
    if (cap >= 2 && cap < stages) stages = cap;
  }
  DEMFI_REQUIRE(stages >= 2, "conv_tc: not enough shared memory for two stages");
  P.stages = stages;
  if (atmem && P.a_stages > stages) P.a_stages = stages;  // the aready[] barriers reuse the split[] slots
  P.diag = get_option("tc_diag");
  if (P.diag & 128) {
    P.dbg = tc_debug_buffer(st);
    DEMFI_REQUIRE(P.dbg != nullptr, "conv_tc: cannot allocate the role-timer buffer");
  }
  P.comp = (float)get_option("tc_comp_milli") * 1e-3f * 5.9604645e-8f;  // milli-units of 2^-24 per MMA
  if (!atmem) P.diag &= (1 | 8 | 16);
  if (P.diag & 32) P.diag |= 1;  // free-running MMA issuer: timing only
  if ((P.diag & 4) && P.nb_max > 64) P.diag &= ~4;
  if (atmem) {
    P.buf_stride = ((P.diag & 4) ? 3 : 2) * P.nb_max;
    P.a_base = 2 * P.buf_stride;
    P.a_stages = (512 - P.a_base) / 64;
    if (P.a_stages > 4) P.a_stages = 4;
    P.tmem_cols = 512;
  } else {
    int buf_stride = 32;  // columns per accumulator pair (main + correction), power of two
    while (buf_stride < 2 * P.nb_max) buf_stride *= 2;
    P.buf_stride = buf_stride;
    P.tmem_cols = 2 * buf_stride;
  }
  P.mask_hi = get_option("tc_mask_hi");
  P.split = get_option("tc_split");
  P.stages_per_tile = 0;
  for (int s = 0; s < c.nsrc; ++s) P.stages_per_tile += ((c.src[s].C + TC_KC - 1) / TC_KC) * c.KH * c.KW;
  P.flush = get_option("tc_flush");
  if (P.flush <= 0 || P.flush > P.stages_per_tile) P.flush = P.stages_per_tile;
  {  // balanced segments: e.g. 18 stages with tc_flush = 8..10 -> 2 x 9 instead of 8 + 8 + 2 (one drain fewer)
    const int nseg = (P.stages_per_tile + P.flush - 1) / P.flush;
    P.flush = (P.stages_per_tile + nseg - 1) / nseg;
  }
  const int smem = stages * stage_bytes + 8 * (3 * stages + 8) + 16 + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    const void* fns[] = {(const void*)conv_tc_kernel<32, false>, (const void*)conv_tc_kernel<64, false>,
                         (const void*)conv_tc_kernel<128, false>, (const void*)conv_tc_kernel<32, true>,
                         (const void*)conv_tc_kernel<64, true>, (const void*)conv_tc_kernel<96, true>};
    for (const void* f : fns) {
      cudaError_t e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      DEMFI_REQUIRE(e == cudaSuccess, "conv_tc: cannot raise dynamic shared memory: %s", cudaGetErrorString(e));
    }
    attr_set = true;
  }
  int grid = P.ntiles < num_sms() ? P.ntiles : num_sms();
  if (get_option("tc_grid") > 0 && get_option("tc_grid") < grid) grid = get_option("tc_grid");
  if (atmem) {
    if (P.nb_max <= 32) conv_tc_kernel<32, true><<<grid, TC_THREADS, smem, st>>>(P);
    else if (P.nb_max <= 64) conv_tc_kernel<64, true><<<grid, TC_THREADS, smem, st>>>(P);
    else conv_tc_kernel<96, true><<<grid, TC_THREADS, smem, st>>>(P);
  } else {
    if (P.nb_max <= 32) conv_tc_kernel<32, false><<<grid, TC_THREADS, smem, st>>>(P);
    else if (P.nb_max <= 64) conv_tc_kernel<64, false><<<grid, TC_THREADS, smem, st>>>(P);
    else conv_tc_kernel<128, false><<<grid, TC_THREADS, smem, st>>>(P);
  }
  DEMFI_LAUNCH_CHECK("conv_tc");
  return 0;
}

// packed layout: for chunk (source-major, 32 channels) and tap: [hi tile][lo tile], each tile
// [cout_pad rows][32 floats] with the 16-byte groups of row n XOR-ed by (n & 7) (128B swizzle).
size_t tc_packed_floats(int KH, int KW, const int32_t* src_C, int nsrc, int cout_pad) {
  size_t chunks = 0;
  for (int s = 0; s < nsrc; ++s) chunks += (size_t)(src_C[s] + TC_KC - 1) / TC_KC;
  return chunks * KH * KW * 2 * (size_t)cout_pad * TC_KC;
}

int tc_pack_weights(const float* w, int Co, int Ci, int KH, int KW, const int32_t* in_map, const int32_t* src_C,
                    int nsrc, const int32_t* out_map, int cout_pad, float* out) {
  DEMFI_REQUIRE(cout_pad % 16 == 0 && cout_pad <= 256, "tc_pack_weights: cout_pad must be a multiple of 16 and <= 256");
  const int taps = KH * KW;
  const size_t tile = (size_t)cout_pad * TC_KC;
  size_t chunk = 0;
  int kbase = 0;
  for (int s = 0; s < nsrc; ++s) {
    for (int c0 = 0; c0 < src_C[s]; c0 += TC_KC, ++chunk) {
      for (int tap = 0; tap < taps; ++tap) {
        float* hi = out + (chunk * taps + tap) * 2 * tile;
        float* lo = hi + tile;
        for (int n = 0; n < cout_pad; ++n)
          for (int k = 0; k < TC_KC; ++k) {
            float v = 0.0f;
            const int c = c0 + k;
            if (c < src_C[s]) {
              const int ci = in_map[kbase + c], co = out_map[n];
              if (ci >= 0 && co >= 0) v = w[((size_t)co * Ci + ci) * taps + tap];
            }
            uint32_t bits;
            memcpy(&bits, &v, 4);
            bits = (bits + 0x1000u) & 0xffffe000u;  // tf32 round-to-nearest (ties away): residual sign is random
            float h;
            memcpy(&h, &bits, 4);
            const size_t off = (size_t)n * TC_KC + (size_t)(((k >> 2) ^ (n & 7)) << 2) + (k & 3);
            hi[off] = h;
            lo[off] = v - h;
          }
      }
    }
    kbase += src_C[s];
  }
  return 0;
}

}  // namespace demfi
