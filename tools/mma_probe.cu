// tcgen05.mma cost table on the GPU it runs on (measurement tool, not a product path).
// One CTA per SM issues a long back-to-back chain of MMAs of a given kind / shape / operand form and reports
// cycles per instruction (clock64 inside the kernel, CTA 0) and the chip-wide dense rate (CUDA events).
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_probe tools/mma_probe.cu
//   run  : tools/mma_probe            (prints one JSON line per configuration)
// Configuration = (kind, TS/SS, N1, N2): per k-step one MMA with N = N1 into accumulator 1 and, when N2 > 0,
// one with N = N2 into accumulator 2 (the [hi;lo]-stacked 3-pass scheme of conv_tc.cu uses N1 = 2*Cout, N2 = Cout).
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, 0xffffffff;\n\tselp.b32 %0, 1, 0, px;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
template <int KIND>
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if constexpr (KIND == 0)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
template <int KIND>
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if constexpr (KIND == 0)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

template <int KIND, bool TS>
__global__ void __launch_bounds__(128 + 256, 1) probe(int N1, int N2, int iters, long long* cycles, int mode) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  constexpr int A_BYTES = 128 * 128, B_BYTES = 256 * 128, NBUF = 2;
  __shared__ uint64_t bar2[2];
  __shared__ volatile int stop_flag;
  // operands: small finite values (fp32 view for tf32, packed halves for f16) so the datapath toggles like real data
  for (int i = threadIdx.x; i < (mode & 2 ? 6 * 32768 : NBUF * (A_BYTES + B_BYTES)) / 4; i += blockDim.x) {
    uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 97u;
    h ^= h >> 15;
    uint32_t v;
    if (KIND == 0) v = 0x3f000000u | (h & 0x007fe000u) | ((h & 1u) << 31);
    else v = (0x3800u | (h & 0x3ffu) | ((h >> 3) & 0x8000u)) | ((0x3800u | ((h >> 10) & 0x3ffu)) << 16);
    reinterpret_cast<uint32_t*>(smem)[i] = v;
  }
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2[0])) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2[1])) : "memory");
    stop_flag = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  {  // fill the TMEM A slots (columns 448..511) with data
    uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 256u;
    if (warp < 4)
    for (int c = 0; c < 256; c += 4) {
      uint32_t h = (threadIdx.x * 64 + c) * 2246822519u;
      uint32_t v0 = KIND == 0 ? (0x3f000000u | (h & 0x007fe000u)) : (0x38003800u | (h & 0x03ff03ffu));
      asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr + c), "r"(v0), "r"(v0 ^ 0x2000u), "r"(v0 ^ 0x4000u), "r"(v0 ^ 0x6000u) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp == 0) {
    const uint32_t fmt = KIND == 0 ? 2u : 0u;
    const uint32_t idesc0 = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t id1 = idesc0 | ((uint32_t)(N1 >> 3) << 17), id2 = idesc0 | ((uint32_t)(N2 >> 3) << 17);
    const uint32_t sb = smem_u32(smem);
    long long t0 = clock64();
    if (mode & 128) {
      uint32_t junk = threadIdx.x;
      if (elect_one()) {
        for (int i = 0; i < iters; ++i) {
          for (int j = 0; j < ((mode >> 9) & 127); ++j) asm volatile("mad.lo.u32 %0, %0, 1664525, 1013904223;" : "+r"(junk));
          const uint32_t buf = (uint32_t)(i & 1) ^ (junk == 0x12345u ? 1u : 0u);
          uint64_t ad = make_desc_sw128(sb + buf * A_BYTES), bd = make_desc_sw128(sb + NBUF * A_BYTES + buf * B_BYTES);
          mma_ss<KIND>(tmem, ad, bd, id1, i ? 1u : 0u);
          mma_ss<KIND>(tmem, ad + 2u, bd + 2u, id1, 1u);
          if (!(mode & 256)) {
            mma_ss<KIND>(tmem + 256u, ad + 4u, bd, id2, i ? 1u : 0u);
            mma_ss<KIND>(tmem + 256u, ad + 6u, bd + 2u, id2, 1u);
          }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      }
    } else if (!(mode & 4)) {
      if (elect_one()) {
        for (int i = 0; i < iters; ++i) {
          const uint32_t buf = (uint32_t)(i & 1);
          uint64_t ad = make_desc_sw128(sb + buf * A_BYTES), bd = make_desc_sw128(sb + NBUF * A_BYTES + buf * B_BYTES);
          uint32_t ta = tmem + 448u + buf * 32u, d1 = tmem, d2 = tmem + 256u;
          if (mode & 1) { d1 = tmem + (uint32_t)((i / 9) & 1) * 128u; d2 = d1 + (uint32_t)N2; ta = tmem + 256u + (uint32_t)(i & 3) * 64u; }
          if (mode & 2) bd = make_desc_sw128(sb + (uint32_t)(i % 6) * 32768u + 16384u);
          // conv_s3-like operands: A = shifted window into a halo tile (start not 1024-aligned, SBO = 10 rows), B = 64-byte swizzle
          if (mode & 32) ad = (uint64_t)(((sb + buf * A_BYTES + (uint32_t)(i % 9) * 384u) >> 4) & 0x3FFF) | ((uint64_t)(1280 >> 4) << 32) | (1ull << 46) | (2ull << 61);
          if (mode & 64) bd = (uint64_t)(((sb + NBUF * A_BYTES + buf * B_BYTES) >> 4) & 0x3FFF) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (4ull << 61);
          const uint32_t ta2 = (mode & 1) ? ta + 32u : ta;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (TS) mma_ts<KIND>(d1, ta + (uint32_t)k * 8u, bd + (uint64_t)(k * 2), id1, (i | k) ? 1u : 0u);
            else mma_ss<KIND>(d1, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), id1, (i | k) ? 1u : 0u);
          }
          if (N2 > 0) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (TS) mma_ts<KIND>(d2, ta2 + (uint32_t)k * 8u, bd + (uint64_t)(k * 2), id2, (i | k) ? 1u : 0u);
              else mma_ss<KIND>(d2, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), id2, (i | k) ? 1u : 0u);
            }
          }
          if (mode & 16) {
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2[0])) : "memory");
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2[1])) : "memory");
          }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      }
    } else {
      // kernel-like issue pattern: the whole warp walks the loop, a dependent ALU chain (~64 ops) per stage, elect per stage
      uint32_t junk = threadIdx.x;
      for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 64; ++j) asm volatile("mad.lo.u32 %0, %0, 1664525, 1013904223;" : "+r"(junk));
        const uint32_t buf = (uint32_t)(i & 1) ^ (junk == 0x12345u ? 1u : 0u);
        uint64_t bd = make_desc_sw128(sb + NBUF * A_BYTES + buf * B_BYTES);
        uint32_t ta = tmem + 448u + buf * 32u, d1 = tmem, d2 = tmem + 256u;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) mma_ts<KIND>(d1, ta + (uint32_t)k * 8u, bd + (uint64_t)(k * 2), id1, (i | k) ? 1u : 0u);
          if (N2 > 0) {
#pragma unroll
            for (int k = 0; k < 4; ++k) mma_ts<KIND>(d2, ta + (uint32_t)k * 8u, bd + (uint64_t)(k * 2), id2, (i | k) ? 1u : 0u);
          }
        }
        __syncwarp();
      }
      if (elect_one())
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    __syncwarp();
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    stop_flag = 1;
  } else if (warp == 1 && (mode & 256)) {
    // second issuer (mode 256): the lo products Al x Bh into its own accumulator, same per-stage ALU chain as warp 0
    const uint32_t idesc0 = (1u << 4) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t id2 = idesc0 | ((uint32_t)(N2 >> 3) << 17);
    const uint32_t sb = smem_u32(smem);
    uint32_t junk = threadIdx.x;
    if (elect_one()) {
      for (int i = 0; i < iters; ++i) {
        for (int j = 0; j < ((mode >> 9) & 127); ++j) asm volatile("mad.lo.u32 %0, %0, 1664525, 1013904223;" : "+r"(junk));
        const uint32_t buf = (uint32_t)(i & 1) ^ (junk == 0x12345u ? 1u : 0u);
        uint64_t ad = make_desc_sw128(sb + buf * A_BYTES), bd = make_desc_sw128(sb + NBUF * A_BYTES + buf * B_BYTES);
        mma_ss<KIND>(tmem + 256u, ad + 4u, bd, id2, i ? 1u : 0u);
        mma_ss<KIND>(tmem + 256u, ad + 6u, bd + 2u, id2, 1u);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2[0])) : "memory");
    }
    __syncwarp();
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar2[0])) : "memory");
    }
  } else if (warp >= 4 && (mode & 8)) {
    // eight extra warps drain an idle accumulator region every ~3000 cycles (what the epilogue warps do)
    uint32_t r[16], sink = 0;
    const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 128u + (warp >= 8 ? 32u : 0u);
    while (!stop_flag) {
      for (int q = 0; q < 4; ++q) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                       "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr + (q & 1) * 16u) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 16; ++j) sink += r[j];
      }
      const long long t = clock64();
      while (clock64() - t < 3000) {}
    }
    if (sink == 0x7fffffffu) cycles[0] = 0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

template <int KIND, bool TS>
static void run(int N1, int N2, int sms, long long* dcyc, int mode = 0) {
  const int iters = 4096, smem = 6 * 32768 + 1024;
  cudaFuncSetAttribute(probe<KIND, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  probe<KIND, TS><<<sms, 128 + 256, smem>>>(N1, N2, 256, dcyc, mode);
  cudaEventRecord(e0);
  probe<KIND, TS><<<sms, 128 + 256, smem>>>(N1, N2, iters, dcyc, mode);
  cudaEventRecord(e1);
  cudaError_t err = cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  std::vector<long long> cyc(sms);
  cudaMemcpy(cyc.data(), dcyc, sms * sizeof(long long), cudaMemcpyDeviceToHost);
  const double kper = KIND == 0 ? 8 : 16;
  const double n_mma = (double)iters * 4 * (N2 > 0 ? 2 : 1);
  const double macs = (double)iters * 4 * 128.0 * (N1 + N2) * kper * sms;
  printf("{\"mode\": %d, \"kind\": \"%s\", \"form\": \"%s\", \"N1\": %d, \"N2\": %d, \"clk_per_mma\": %.1f, \"clk_per_kstep\": %.1f, \"ms\": %.3f, "
         "\"dense_TFLOPs\": %.1f, \"err\": \"%s\"}\n",
         mode, KIND == 0 ? "tf32" : "f16", TS ? "TS" : "SS", N1, N2, cyc[0] / n_mma, cyc[0] / ((double)iters * 4), ms,
         2 * macs / ms / 1e9, cudaGetErrorString(err));
  fflush(stdout);
}

int main(int argc, char**) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  long long* dcyc;
  cudaMalloc(&dcyc, sms * sizeof(long long));
  const int cfg[][2] = {{64, 0}, {128, 0}, {256, 0}, {32, 0}, {128, 64}, {64, 32}, {192, 96}, {256, 128}, {64, 64}, {96, 96}};
  if (argc > 3) {  // conv_s3-like stage (2 x N1 + 2 x N2 per stage) with an ALU chain of c ops per stage; single vs dual issuer
    for (int c : {0, 8, 16, 24, 32, 48})
      for (int dual : {0, 256}) run<1, false>(128, 64, sms, dcyc, 128 | dual | (c << 9));
    return 0;
  }
  if (argc > 2) {  // conv_s3-like SS operands: 32 shifted-window A, 64 SW64 B
    for (int mode : {0, 32, 64, 96}) run<1, false>(128, 64, sms, dcyc, mode);
    for (int mode : {0, 32, 64, 96}) run<1, false>(256, 128, sms, dcyc, mode);
    return 0;
  }
  if (argc > 1) {  // interference modes: 1 kernel-like TMEM layout, 2 six rotating B buffers, 4 kernel-like issue pattern,
                   // 8 concurrent tcgen05.ld warps, 16 two commits per stage
    for (int mode : {0, 1, 2, 4, 8, 16, 3, 31 - 4, 31}) run<0, true>(128, 64, sms, dcyc, mode);
    for (int mode : {0, 1, 8, 31}) run<1, true>(128, 64, sms, dcyc, mode);
    return 0;
  }
  for (auto& c : cfg) run<0, true>(c[0], c[1], sms, dcyc);
  for (auto& c : cfg) run<0, false>(c[0], c[1], sms, dcyc);
  for (auto& c : cfg) run<1, true>(c[0], c[1], sms, dcyc);
  for (auto& c : cfg) run<1, false>(c[0], c[1], sms, dcyc);
  return 0;
}
