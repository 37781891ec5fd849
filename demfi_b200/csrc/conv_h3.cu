// tcgen05 implicit-GEMM convolution, second generation: kind::f16 tensor-core passes with fp32 parity
// ("3xFP16": every fp32 operand is split into two halves, a = h + l/2048, h = fp16(a), l = fp16((a-h)*2048)).
//
//   D_main += Ah * Bh            D_corr += Ah * Bl + Al * Bh          out = D_main + D_corr / 2048
//
// Products of two fp16 values are exact in the fp32 accumulator, the dropped term Al*Bl is ~2^-22 relative
// (the same as the 3xTF32 scheme of conv_tc.cu: fp16 and tf32 both carry 11 significant bits), and the
// 2048 scaling keeps the low halves in the normal fp16 range.  kind::f16 runs at twice the kind::tf32 rate
// (measured on B200 with tools/mma_probe.cu: 2046 vs 1100 dense TFLOP/s) and moves half the operand bytes.
// Operands must satisfy |a| < 65504 (fp16 range); DeMFI-Net activations are tanh/sigmoid/ReLU features of O(1).
//
// Data flow per CTA (persistent, one per SM), tile = 8 x 16 output pixels (M = 128), N block <= 96 channels:
//   * activations: ONE TMA box per 32-channel chunk covering the tile plus its halo
//     {32 ch, 16*s + KW - s, 8*s + KH - s} (s = stride), 128-byte swizzled rows = pixels.  Every tap of the
//     chunk is served from this shared-memory tile: L2 -> SM activation traffic is 1.4x the tile instead of
//     KH*KW times (the first-generation kernel was within 20 % of the L2 -> SM bandwidth).
//   * eight splitter warps split the landed halo tile ONCE, in place (a pixel's 128 bytes of fp32 become
//     Ah | Al, 2 x 32 fp16); per tap, two groups of four warps taking alternate taps copy their pixel's row at
//     the tap's offset into a ring of 32-column TMEM slots (tcgen05.st): the MMAs take A from tensor memory
//     (no shared-memory reads for A) and the conversion cost is paid per halo pixel, not per tap.
//   * weights: per (chunk, tap) a pre-packed [Bh rows ; Bl rows] x 32 fp16 tile (64-byte swizzle) by cp.async.bulk.
//   * MMAs per tap: 2 x (Ah x [Bh;Bl], N' = 2N: main | corr) + 2 x (Al x Bh -> corr), issued by ONE thread
//     whose loop is kept to a barrier probe + the tcgen05 instructions: measured (tools/mma_probe.cu, mode 4)
//     the tensor pipe only runs ~1-2 MMAs ahead of the issuing thread, so every instruction in that loop
//     is serial with the MMAs.
//   * the K loop is cut into segments whose partial sums are drained and added in fp32 RN by eight epilogue
//     warps (the tensor core accumulates with truncation), then the fused epilogue of common.cuh.
#include <cuda.h>
#include <cuda_fp16.h>

#include <cstring>

#include "common.cuh"

namespace demfi {

constexpr int H3_TH = 8, H3_TW = 16, H3_BM = 128, H3_KC = 32;
constexpr int H3_SPLIT_WARPS = 8, H3_EPI_WARPS = 8;
constexpr int H3_THREADS = (H3_SPLIT_WARPS + H3_EPI_WARPS + 2) * 32;  // 576: 0-7 split, 8-15 epilogue, 16 TMA, 17 MMA
constexpr int H3_NA = 2;      // halo-tile buffers
constexpr int H3_MAX_NS = 8;  // ring of (TMEM A slot + shared-memory weight stage)
constexpr float H3_LO_SCALE = 2048.0f;

struct H3Params {
  CUtensorMap tmap[DEMFI_MAX_SRC];
  demfi_conv_t c;
  int tiles_x, tiles_y, ntiles, n_blocks, nb_max;
  int hw, hh;      // halo tile in pixels
  int a_bytes;     // one halo buffer (multiple of 1024)
  int b_bytes;     // one weight stage = 2 * nb_max * 64
  int ns;          // ring depth (even)
  int acc_stride;  // TMEM columns per accumulator buffer: main nb_max | corr nb_max
  int a_base;      // first TMEM column of the A ring (32 columns per slot: Ah 16 | Al 16)
  int taps, stages_per_tile, flush;
  float comp;      // per-MMA gain correction of the truncating accumulation
  int diag;
  long long* dbg;
};

// ---- PTX wrappers (same conventions as conv_tc.cu) ----------------------------------------
namespace h3 {
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
      printf("demfi conv_h3: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", (int)blockIdx.x,
             (int)threadIdx.x, bar, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (!mbar_try_wait(bar, parity)) mbar_wait_slow(bar, parity);
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
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
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
// A from tensor memory (lane = pixel row, two fp16 per 32-bit column), B from shared memory
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
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
// K-major shared-memory matrix descriptor, 64-byte swizzle (rows of 32 fp16): start >> 4 in [0,14), SBO = 512 B
// between 8-row groups in [32,46), version 1 in [46,48), layout SWIZZLE_64B (4) in [61,64).
__device__ __forceinline__ uint64_t make_desc_sw64(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (4ull << 61);
}
// a = h + l / 2048 for two values: returns the packed fp16 pairs (low half = first value)
__device__ __forceinline__ void split2(float a0, float a1, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(a0, a1);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn((a0 - hf.x) * H3_LO_SCALE, (a1 - hf.y) * H3_LO_SCALE);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
}  // namespace h3
using namespace h3;

template <int NMAX>
__global__ void __launch_bounds__(H3_THREADS, 1) conv_h3_kernel(const __grid_constant__ H3Params P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const demfi_conv_t& c = P.c;
  const int NS = P.ns;
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t b_base = smem_base + (uint32_t)(H3_NA * P.a_bytes);
  const uint32_t bars = b_base + (uint32_t)(NS * P.b_bytes);
  auto bar_afull = [&](int a) { return bars + 8u * (uint32_t)a; };
  auto bar_aempty = [&](int a) { return bars + 8u * (uint32_t)(H3_NA + a); };
  auto bar_tfull = [&](int a) { return bars + 8u * (uint32_t)(2 * H3_NA + a); };
  auto bar_tempty = [&](int a) { return bars + 8u * (uint32_t)(2 * H3_NA + 2 + a); };
  auto bar_bfull = [&](int s) { return bars + 8u * (uint32_t)(2 * H3_NA + 4 + s); };
  auto bar_ready = [&](int s) { return bars + 8u * (uint32_t)(2 * H3_NA + 4 + NS + s); };
  auto bar_free = [&](int s) { return bars + 8u * (uint32_t)(2 * H3_NA + 4 + 2 * NS + s); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + (size_t)(H3_NA * P.a_bytes + NS * P.b_bytes) + 8 * (2 * H3_NA + 4 + 3 * H3_MAX_NS));
  auto n_of = [&](int nb) { return min(P.nb_max, c.cout_pad - nb * P.nb_max); };

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const bool dbg = P.dbg != nullptr;

  if (threadIdx.x == 0) {
    for (int a = 0; a < H3_NA; ++a) {
      mbar_init(bar_afull(a), 1);
      mbar_init(bar_aempty(a), H3_SPLIT_WARPS);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull(a), 1);
      mbar_init(bar_tempty(a), H3_EPI_WARPS);
    }
    for (int s = 0; s < NS; ++s) {
      mbar_init(bar_bfull(s), 1);
      mbar_init(bar_ready(s), 4);
      mbar_init(bar_free(s), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 17) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < H3_SPLIT_WARPS) {
    // ===== A splitter: thread = output pixel m of the tile; group g takes the stages q with q % 2 == g =====
    const int g = warp >> 2;
    const int m = (warp & 3) * 32 + lane;
    const int py = (m >> 4) * c.stride, px = (m & 15) * c.stride;
    const uint32_t lane_addr = ((uint32_t)((warp & 3) * 32) << 16);
    long long w_afull = 0, w_free = 0, w_st = 0, w_cv = 0;
    const long long t_begin = dbg ? clock64() : 0;
    uint32_t q = 0;  // running stage index (all tiles)
    int slot = g;    // this group's next ring slot (advances by 2; the ring depth is even)
    uint32_t sphase = 0;
    int abuf = 0;
    uint32_t aphase = 0;
    const int chunks_per_tile = P.stages_per_tile / P.taps;
    const int halo_rows = P.hw * P.hh;
    for (int tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x) {
      for (int ch = 0; ch < chunks_per_tile; ++ch) {
        mbar_wait_t(bar_afull(abuf), aphase, dbg, w_afull);
        const uint32_t a_addr = smem_base + (uint32_t)(abuf * P.a_bytes);
        // (1) split the halo tile ONCE, in place: row = pixel, 32 fp32 -> [Ah 32 x fp16 | Al 32 x fp16] (same 128 bytes,
        //     same 16-byte-group swizzle); every tap of the chunk then only copies rows to tensor memory.
        const long long t_cv = dbg ? clock64() : 0;
        for (int r = (int)threadIdx.x; r < halo_rows; r += H3_SPLIT_WARPS * 32) {
          const uint32_t row = a_addr + (uint32_t)r * 128u;
          const uint32_t sw = (uint32_t)r & 7u;
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
        asm volatile("bar.sync 1, %0;" ::"n"(H3_SPLIT_WARPS * 32) : "memory");
        if (dbg) w_cv += clock64() - t_cv;
        // (2) per tap: this thread's pixel row at the tap's offset -> one 32-column TMEM slot
        int ky = 0, kx = 0;
        for (int tap = 0; tap < P.taps; ++tap, ++q) {
          if ((int)(q & 1u) == g) {
            const uint32_t r = (uint32_t)((py + ky) * P.hw + px + kx);
            const uint32_t row = a_addr + r * 128u;
            const uint32_t sw = r & 7u;
            uint32_t v[32];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const uint4 u = lds128(row + ((((uint32_t)j) ^ sw) << 4));
              v[4 * j] = u.x; v[4 * j + 1] = u.y; v[4 * j + 2] = u.z; v[4 * j + 3] = u.w;
            }
            mbar_wait_t(bar_free(slot), sphase ^ 1u, dbg, w_free);
            tc_fence_after();
            const long long t_st = dbg ? clock64() : 0;
            tmem_st32(tmem_base + lane_addr + (uint32_t)(P.a_base + slot * 32), v);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              mbar_wait(bar_bfull(slot), sphase);  // the issuer only probes ready[]: it implies the weight tile has landed
              mbar_arrive(bar_ready(slot));
            }
            if (dbg) w_st += clock64() - t_st;
            slot += 2;
            if (slot >= NS) { slot -= NS; sphase ^= 1u; }
          }
          if (++kx == c.KW) { kx = 0; ++ky; }
        }
        // the buffer was rewritten through the generic proxy; the next TMA load writes it through the async proxy
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_aempty(abuf));
        if (++abuf == H3_NA) { abuf = 0; aphase ^= 1u; }
      }
    }
    if (dbg && threadIdx.x == 0) {
      long long* d = P.dbg + (size_t)blockIdx.x * 16;
      d[0] = clock64() - t_begin; d[1] = w_afull; d[2] = w_free; d[3] = w_st; d[7] = w_cv;
    }
  } else if (warp < H3_SPLIT_WARPS + H3_EPI_WARPS) {
    // ===== epilogue: TMEM lane = pixel row; a warp can only touch lanes 32*(warp%4)...  Group 0 (warps 8-11)
    // takes accumulator columns [0, csplit), group 1 (warps 12-15) takes [csplit, N). =====
    const int grp = (warp - H3_SPLIT_WARPS) >> 2;
    const int wq = warp & 3;
    const int m = wq * 32 + lane;
    constexpr int HMAX = (NMAX / 2 + 15) / 16 * 16;
    int acc = 0;
    uint32_t acc_phase = 0;
    long long w_tfull = 0, w_store = 0;
    const long long t_begin = dbg ? clock64() : 0;
    for (int tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x) {
      int t = tile;
      const int nb = t % P.n_blocks;
      t /= P.n_blocks;
      const int tx0 = (t % P.tiles_x) * H3_TW;
      t /= P.tiles_x;
      const int ty0 = (t % P.tiles_y) * H3_TH;
      const int n = t / P.tiles_y;
      const int N = n_of(nb), n0 = nb * P.nb_max;
      const int csplit = ((N / 2 + 15) / 16) * 16;
      const int cbeg = grp ? csplit : 0, cnum = grp ? N - csplit : csplit;
      const int oy = ty0 + (m >> 4), ox = tx0 + (m & 15);
      const bool valid = (oy < c.H) && (ox < c.W);
      float sum[HMAX];
      bool first = true;
      for (int done = 0; done < P.stages_per_tile; done += P.flush) {
        const float gain = 1.0f + P.comp * (float)(2 * min(P.flush, P.stages_per_tile - done));
        mbar_wait_t(bar_tfull(acc), acc_phase, dbg, w_tfull);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(acc * P.acc_stride) + (uint32_t)cbeg;
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
#pragma unroll
        for (int col = 0; col < HMAX; col += 16)
          if (col < cnum) tmem_ld16_nowait(taddr + (uint32_t)(N + col), r + col);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < HMAX; ++j) sum[j] = fmaf(__uint_as_float(r[j]), 1.0f / H3_LO_SCALE, sum[j]);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tempty(acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        first = false;
      }
      const long long t_store = dbg ? clock64() : 0;
      if (valid && !(P.diag & 1)) {
        const int ch_lo = n0 + cbeg, ch_hi = ch_lo + cnum;
#pragma unroll 1
        for (int sgi = 0; sgi < c.nseg; ++sgi) {
          const demfi_seg_t& sg = c.seg[sgi];
          if (sg.ch0 >= ch_hi || sg.ch0 + sg.nch <= ch_lo) continue;
          const SegCursor cur = seg_cursor(c, sg, n, oy, ox);
          if (sg.store == DEMFI_STORE_NHWC && (sg.act == DEMFI_ACT_RELU || sg.act == DEMFI_ACT_NONE)) {
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
    if (dbg && warp == H3_SPLIT_WARPS && lane == 0) {
      long long* d = P.dbg + (size_t)blockIdx.x * 16;
      d[4] = clock64() - t_begin; d[5] = w_tfull; d[6] = w_store;
    }
  } else if (warp == 16) {
    // ===== TMA producer (one thread): per chunk the halo tile, per tap the weight tile =====
    if (elect_one()) {
      long long w_aempty = 0, w_free = 0;
      const long long t_begin = dbg ? clock64() : 0;
      int slot = 0, abuf = 0;
      uint32_t sphase = 0, aphase = 0;
      const uint32_t a_tx = (uint32_t)(P.hw * P.hh * 128);
      for (int tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x) {
        int t = tile;
        const int nb = t % P.n_blocks;
        t /= P.n_blocks;
        const int tx0 = (t % P.tiles_x) * H3_TW * c.stride - c.pad_w;
        t /= P.tiles_x;
        const int ty0 = (t % P.tiles_y) * H3_TH * c.stride - c.pad_h;
        const int n = t / P.tiles_y;
        const int N = n_of(nb);
        const uint32_t b_tx = (uint32_t)N * 128u;  // 2N rows x 64 bytes
        // packed weights: [n block][chunk][tap][2*N_block rows][32 fp16]; full blocks hold nb_max channels
        const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(c.wpack) + (size_t)nb * (size_t)P.stages_per_tile * (size_t)P.nb_max * 128u;
        for (int s = 0; s < c.nsrc; ++s) {
          const CUtensorMap* map = &P.tmap[s];
          const int Cs = c.src[s].C;
          for (int c0 = 0; c0 < Cs; c0 += H3_KC) {
            mbar_wait_t(bar_aempty(abuf), aphase ^ 1u, dbg, w_aempty);
            mbar_arrive_expect_tx(bar_afull(abuf), a_tx);
            tma_load_4d(smem_base + (uint32_t)(abuf * P.a_bytes), map, bar_afull(abuf), c0, tx0, ty0, n);
            if (++abuf == H3_NA) { abuf = 0; aphase ^= 1u; }
            for (int tap = 0; tap < P.taps; ++tap) {
              mbar_wait_t(bar_free(slot), sphase ^ 1u, dbg, w_free);
              mbar_arrive_expect_tx(bar_bfull(slot), b_tx);
              bulk_load(b_base + (uint32_t)(slot * P.b_bytes), wsrc, b_tx, bar_bfull(slot));
              wsrc += b_tx;
              if (++slot == NS) { slot = 0; sphase ^= 1u; }
            }
          }
        }
      }
      if (dbg) {
        long long* d = P.dbg + (size_t)blockIdx.x * 16;
        d[8] = clock64() - t_begin; d[9] = w_free; d[10] = w_aempty;
      }
    }
  } else {
    // ===== MMA issuer: ONE thread runs the whole persistent loop =====
    if (elect_one()) {
      const uint32_t idesc0 = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(H3_BM >> 4) << 24);  // D=f32, A=B=f16, K-major, M=128
      const uint64_t bdesc0 = make_desc_sw64(b_base);
      const uint32_t bstep = (uint32_t)(P.b_bytes >> 4);
      const uint32_t a_tmem0 = tmem_base + (uint32_t)P.a_base;
      long long w_tempty = 0, w_ready = 0;
      const long long t_begin = dbg ? clock64() : 0;
      uint32_t slot = 0, sphase = 0, acc = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x) {
        const int N = n_of(tile % P.n_blocks);
        const uint32_t idesc_n = idesc0 | ((uint32_t)(N >> 3) << 17);
        const uint32_t idesc_2n = idesc0 | ((uint32_t)((2 * N) >> 3) << 17);
        for (int done = 0; done < P.stages_per_tile; done += P.flush) {
          const int seg_len = min(P.flush, P.stages_per_tile - done);
          mbar_wait_t(bar_tempty(acc), acc_phase ^ 1u, dbg, w_tempty);
          tc_fence_after();
          const uint32_t d_main = tmem_base + acc * (uint32_t)P.acc_stride;
          const uint32_t d_corr = d_main + (uint32_t)N;
          for (int i = 0; i < seg_len; ++i) {
            mbar_wait_t(bar_ready(slot), sphase, dbg, w_ready);
            tc_fence_after();
            const uint32_t ta = a_tmem0 + slot * 32u;
            const uint64_t bd = bdesc0 + (uint64_t)(bstep * slot);
            umma_f16_ts(d_main, ta, bd, idesc_2n, i == 0 ? 0u : 1u);      // Ah x [Bh;Bl]  k 0..15
            umma_f16_ts(d_main, ta + 8u, bd + 2u, idesc_2n, 1u);          //               k 16..31
            umma_f16_ts(d_corr, ta + 16u, bd, idesc_n, 1u);               // Al x Bh
            umma_f16_ts(d_corr, ta + 24u, bd + 2u, idesc_n, 1u);
            umma_commit(bar_free(slot));
            if (i == seg_len - 1) umma_commit(bar_tfull(acc));
            if (++slot == (uint32_t)NS) { slot = 0; sphase ^= 1u; }
          }
          acc ^= 1u;
          if (acc == 0) acc_phase ^= 1u;
        }
      }
      if (dbg) {
        long long* d = P.dbg + (size_t)blockIdx.x * 16;
        d[12] = clock64() - t_begin; d[13] = w_tempty; d[14] = w_ready;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 17) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// ---- host --------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn h3_encode_fn() {
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
static int h3_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}
static int h3_nb_max(int cout_pad) { return cout_pad <= 96 ? cout_pad : 64; }

bool h3_supports(const demfi_conv_t& c) {
  if (c.stride != 1 && c.stride != 2) return false;
  if (c.cout_pad % 16 != 0 || c.cout_pad < 16 || c.cout_pad > 256) return false;
  for (int s = 0; s < c.nsrc; ++s)
    if (c.src[s].up != 0) return false;
  const int hw = (H3_TW - 1) * c.stride + c.KW, hh = (H3_TH - 1) * c.stride + c.KH;
  if (hw > 256 || hh > 256) return false;
  const int a_bytes = (hw * hh * 128 + 1023) / 1024 * 1024;
  const int b_bytes = 2 * h3_nb_max(c.cout_pad) * 64;
  return H3_NA * a_bytes + 4 * b_bytes + 2048 <= 227 * 1024;
}

int launch_conv_h3(const demfi_conv_t& c, cudaStream_t st) {
  DEMFI_REQUIRE(h3_supports(c), "conv_h3: unsupported convolution (stride %d, cout_pad %d, %dx%d)", c.stride, c.cout_pad, c.KH, c.KW);
  DEMFI_REQUIRE((c.Hi + 2 * c.pad_h - c.KH) / c.stride + 1 == c.H && (c.Wi + 2 * c.pad_w - c.KW) / c.stride + 1 == c.W, "conv_h3: inconsistent sizes");
  EncodeTiledFn enc = h3_encode_fn();
  DEMFI_REQUIRE(enc != nullptr, "conv_h3: cuTensorMapEncodeTiled not available from the driver");
  static thread_local H3Params P;  // CUtensorMap needs 64-byte alignment; thread_local storage gives it
  memset(&P, 0, sizeof(P));
  P.c = c;
  P.hw = (H3_TW - 1) * c.stride + c.KW;
  P.hh = (H3_TH - 1) * c.stride + c.KH;
  for (int s = 0; s < c.nsrc; ++s) {
    const demfi_src_t& S = c.src[s];
    cuuint64_t dims[4] = {(cuuint64_t)S.C, (cuuint64_t)c.Wi, (cuuint64_t)c.Hi, (cuuint64_t)c.N};
    cuuint64_t strides[3] = {(cuuint64_t)S.ld * 4, (cuuint64_t)S.ld * 4 * c.Wi, (cuuint64_t)S.ld * 4 * c.Wi * c.Hi};
    cuuint32_t box[4] = {H3_KC, (cuuint32_t)P.hw, (cuuint32_t)P.hh, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&P.tmap[s], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(S.ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DEMFI_REQUIRE(r == CUDA_SUCCESS, "conv_h3: cuTensorMapEncodeTiled failed for source %d (CUresult %d)", s, (int)r);
  }
  P.tiles_x = (c.W + H3_TW - 1) / H3_TW;
  P.tiles_y = (c.H + H3_TH - 1) / H3_TH;
  const long long nt = (long long)P.tiles_x * P.tiles_y * c.N;
  P.nb_max = h3_nb_max(c.cout_pad);
  P.n_blocks = (c.cout_pad + P.nb_max - 1) / P.nb_max;
  DEMFI_REQUIRE(nt > 0 && nt * P.n_blocks < (1ll << 31), "conv_h3: bad tile count");
  P.ntiles = (int)nt * P.n_blocks;
  P.a_bytes = (P.hw * P.hh * 128 + 1023) / 1024 * 1024;
  P.b_bytes = 2 * P.nb_max * 64;
  P.acc_stride = 2 * P.nb_max;
  P.a_base = 2 * P.acc_stride;
  int ns = (512 - P.a_base) / 32;
  if (ns > H3_MAX_NS) ns = H3_MAX_NS;
  while (ns > 2 && H3_NA * P.a_bytes + ns * P.b_bytes + 2048 > 227 * 1024) ns -= 2;
  {
    const int cap = get_option("tc_stages");
    if (cap >= 2 && cap < ns) ns = cap & ~1;
  }
  DEMFI_REQUIRE(ns >= 2 && ns % 2 == 0, "conv_h3: ring depth %d", ns);
  P.ns = ns;
  P.taps = c.KH * c.KW;
  int chunks = 0;
  for (int s = 0; s < c.nsrc; ++s) chunks += (c.src[s].C + H3_KC - 1) / H3_KC;
  P.stages_per_tile = chunks * P.taps;
  P.flush = get_option("tc_flush");
  if (P.flush <= 0 || P.flush > P.stages_per_tile) P.flush = P.stages_per_tile;
  {  // balanced segments
    const int nseg = (P.stages_per_tile + P.flush - 1) / P.flush;
    P.flush = (P.stages_per_tile + nseg - 1) / nseg;
  }
  P.comp = (float)get_option("tc_comp_milli") * 1e-3f * 5.9604645e-8f;
  P.diag = get_option("tc_diag") & (1 | 128);
  if (P.diag & 128) {
    long long* buf = tc_debug_buffer(st);
    DEMFI_REQUIRE(buf != nullptr, "conv_h3: cannot allocate the role-timer buffer");
    P.dbg = buf;
  }
  const int smem = H3_NA * P.a_bytes + ns * P.b_bytes + 8 * (2 * H3_NA + 4 + 3 * H3_MAX_NS) + 16 + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    const void* fns[] = {(const void*)conv_h3_kernel<32>, (const void*)conv_h3_kernel<64>, (const void*)conv_h3_kernel<96>};
    for (const void* f : fns) {
      cudaError_t e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      DEMFI_REQUIRE(e == cudaSuccess, "conv_h3: cannot raise dynamic shared memory: %s", cudaGetErrorString(e));
    }
    attr_set = true;
  }
  int grid = P.ntiles < h3_num_sms() ? P.ntiles : h3_num_sms();
  if (get_option("tc_grid") > 0 && get_option("tc_grid") < grid) grid = get_option("tc_grid");
  if (P.nb_max <= 32) conv_h3_kernel<32><<<grid, H3_THREADS, smem, st>>>(P);
  else if (P.nb_max <= 64) conv_h3_kernel<64><<<grid, H3_THREADS, smem, st>>>(P);
  else conv_h3_kernel<96><<<grid, H3_THREADS, smem, st>>>(P);
  DEMFI_LAUNCH_CHECK("conv_h3");
  return 0;
}

// packed layout (fp16 pairs in a float-typed buffer): [n block][chunk (source-major, 32 channels)][tap]
// [Bh rows of the block ; Bl rows of the block][32 fp16], the 16-byte groups of row r XOR-ed by (r >> 1) & 3
// (64-byte swizzle).  Bl = fp16((w - Bh) * 2048).
size_t h3_packed_floats(int KH, int KW, const int32_t* src_C, int nsrc, int cout_pad) {
  size_t chunks = 0;
  for (int s = 0; s < nsrc; ++s) chunks += (size_t)(src_C[s] + H3_KC - 1) / H3_KC;
  return chunks * KH * KW * 2 * (size_t)cout_pad * H3_KC / 2;
}

int h3_pack_weights(const float* w, int Co, int Ci, int KH, int KW, const int32_t* in_map, const int32_t* src_C,
                    int nsrc, const int32_t* out_map, int cout_pad, float* out, int nb_max) {
  DEMFI_REQUIRE(cout_pad % 16 == 0 && cout_pad <= 256, "h3_pack_weights: cout_pad must be a multiple of 16 and <= 256");
  const int taps = KH * KW;
  const int nbm = nb_max > 0 ? nb_max : h3_nb_max(cout_pad);  // nb_max > 0: the N blocking of DEMFI_CONV_TC16W (conv_s3)
  const int n_blocks = (cout_pad + nbm - 1) / nbm;
  int chunks = 0;
  for (int s = 0; s < nsrc; ++s) chunks += (src_C[s] + H3_KC - 1) / H3_KC;
  __half* o = reinterpret_cast<__half*>(out);
  size_t base = 0;  // in halves
  for (int nb = 0; nb < n_blocks; ++nb) {
    const int N = (cout_pad - nb * nbm) < nbm ? (cout_pad - nb * nbm) : nbm;
    int chunk = 0, kbase = 0;
    for (int s = 0; s < nsrc; ++s) {
      for (int c0 = 0; c0 < src_C[s]; c0 += H3_KC, ++chunk) {
        for (int tap = 0; tap < taps; ++tap) {
          __half* tile = o + base + ((size_t)chunk * taps + tap) * (size_t)(2 * N * H3_KC);
          for (int n = 0; n < N; ++n)
            for (int k = 0; k < H3_KC; ++k) {
              float v = 0.0f;
              const int cc = c0 + k;
              if (cc < src_C[s]) {
                const int ci = in_map[kbase + cc], co = out_map[nb * nbm + n];
                if (ci >= 0 && co >= 0) v = w[((size_t)co * Ci + ci) * taps + tap];
              }
              DEMFI_REQUIRE(v > -65504.0f && v < 65504.0f, "h3_pack_weights: weight %g outside the fp16 range", (double)v);
              const __half h = __float2half_rn(v);
              const __half l = __float2half_rn((v - __half2float(h)) * H3_LO_SCALE);
              const int rh = n, rl = N + n;
              tile[(size_t)rh * H3_KC + (size_t)((((k >> 3) ^ ((rh >> 1) & 3)) << 3) + (k & 7))] = h;
              tile[(size_t)rl * H3_KC + (size_t)((((k >> 3) ^ ((rl >> 1) & 3)) << 3) + (k & 7))] = l;
            }
        }
      }
      kbase += src_C[s];
    }
    base += (size_t)chunks * taps * 2 * N * H3_KC;
  }
  return 0;
}

// The same layout written on the DEVICE from a device-resident OIHW weight (one source of src_c >= Ci channels, identity channel
// maps with zero padding): the training step repacks every weight after every optimizer step and must not wait for the GPU to do
// it on the host.  One thread per (output channel, chunk, tap, k).  Out-of-range values saturate to the fp16 limits.
__global__ void h3_pack_kernel(const float* __restrict__ w, int Co, int Ci, int taps, int cout_pad, int nbm, int chunks, __half* __restrict__ o) {
  const long long total = (long long)cout_pad * chunks * taps * H3_KC;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(idx & (H3_KC - 1));
    long long t = idx / H3_KC;
    const int tap = (int)(t % taps);
    t /= taps;
    const int chunk = (int)(t % chunks);
    const int ng = (int)(t / chunks);
    const int nb = ng / nbm, n = ng - nb * nbm;
    const int N = (cout_pad - nb * nbm) < nbm ? (cout_pad - nb * nbm) : nbm;
    __half* tile = o + (size_t)nb * chunks * taps * 2 * nbm * H3_KC + ((size_t)chunk * taps + tap) * (size_t)(2 * N * H3_KC);
    const int cc = chunk * H3_KC + k;
    float v = (cc < Ci && ng < Co) ? w[((size_t)ng * Ci + cc) * taps + tap] : 0.0f;
    v = fminf(fmaxf(v, -65504.0f), 65504.0f);
    const __half h = __float2half_rn(v);
    const __half l = __float2half_rn(fminf(fmaxf((v - __half2float(h)) * H3_LO_SCALE, -65504.0f), 65504.0f));
    const int rh = n, rl = N + n;
    tile[(size_t)rh * H3_KC + (size_t)((((k >> 3) ^ ((rh >> 1) & 3)) << 3) + (k & 7))] = h;
    tile[(size_t)rl * H3_KC + (size_t)((((k >> 3) ^ ((rl >> 1) & 3)) << 3) + (k & 7))] = l;
  }
}

int h3_pack_weights_device(const float* w, int Co, int Ci, int KH, int KW, int src_c, int cout_pad, float* out, cudaStream_t st) {
  DEMFI_REQUIRE(cout_pad % 16 == 0 && cout_pad <= 256 && Co <= cout_pad, "pack_weights_device: cout_pad must be a multiple of 16, >= Co and <= 256");
  DEMFI_REQUIRE(src_c >= Ci && src_c % 4 == 0, "pack_weights_device: src_c must be >= Ci and a multiple of 4");
  const int chunks = (src_c + H3_KC - 1) / H3_KC;
  const long long total = (long long)cout_pad * chunks * KH * KW * H3_KC;
  const int threads = 256;
  const int blocks = (int)((total + threads - 1) / threads < 4096 ? (total + threads - 1) / threads : 4096);
  h3_pack_kernel<<<blocks, threads, 0, st>>>(w, Co, Ci, KH * KW, cout_pad, h3_nb_max(cout_pad), chunks, reinterpret_cast<__half*>(out));
  DEMFI_LAUNCH_CHECK("h3_pack_kernel");
  return 0;
}

}  // namespace demfi
