// tcgen05.mma.cta_group::2 (CTA pair, M = 256) cost table, SS form -- companion of mma_probe.cu (measurement tool).
// A cluster of two CTAs; the leader issues per k-step one MMA with N = N1 and, when N2 > 0, one with N = N2, each CTA
// supplying its own 128 rows of A and half of the B rows from its own shared memory.  Optional extra warps stream
// shared memory (mode & 1: ld/st.shared traffic of `rate` 16-byte accesses per thread between pauses) to show how the
// operand fetch shares the shared-memory pipe.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_probe_pair tools/mma_probe_pair.cu
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
template <int CG>
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if constexpr (CG == 2)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int CG>
__global__ void __launch_bounds__(128 + 256, 1) probe(int N1, int N2, int iters, long long* cycles, int mode, int rate) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  __shared__ volatile int stop_flag;
  constexpr int A_BYTES = 128 * 128, B_BYTES = 256 * 128, NBUF = 2;
  uint32_t rank = 0;
  if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  for (int i = threadIdx.x; i < (mode & 0x100 ? 220 * 1024 : NBUF * (A_BYTES + B_BYTES) + 32768) / 4; i += blockDim.x) {
    uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 97u;
    h ^= h >> 15;
    uint32_t v = (0x3800u | (h & 0x3ffu) | ((h >> 3) & 0x8000u)) | ((0x3800u | ((h >> 10) & 0x3ffu)) << 16);
    if (mode & 0x800) v = 0u;                                             // zeros
    if (mode & 0x1000) v = (h & 0x03ff03ffu) | ((h >> 3) & 0x80008000u);  // fp16 denormals
    if (mode & 0x2000) v = (0x7c00u | (h & 0x3ffu)) | ((0x7c00u | ((h >> 10) & 0x3ffu)) << 16);  // NaN / Inf patterns
    reinterpret_cast<uint32_t*>(smem)[i] = v;
  }
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    stop_flag = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    if (CG == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) cluster_sync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (warp == 0) {
    if (rank == 0) {
      const uint32_t idesc0 = (1u << 4) | ((uint32_t)((CG == 2 ? 256 : 128) >> 4) << 24);
      const uint32_t id1 = idesc0 | ((uint32_t)(N1 >> 3) << 17), id2 = idesc0 | ((uint32_t)(N2 >> 3) << 17);
      const uint32_t sb = smem_u32(smem);
      long long t0 = clock64();
      if (elect_one()) {
        uint32_t u0 = (uint32_t)iters, u1 = (uint32_t)N1, u2 = (uint32_t)N2, u3 = (uint32_t)mode;  // uniform values (kernel parameters)
        for (int i = 0; i < iters; ++i) {
          // extra uniform-datapath work per stage: (mode >> 8) & 63 adds, as one dependent chain (mode & 64) or four chains
          const int nx = (mode >> 8) & 63;
          if (mode & 64) {
            for (int j = 0; j < nx; ++j) u0 = u0 * 5u + (uint32_t)i;
          } else {
            for (int j = 0; j < nx; j += 4) { u0 = u0 * 5u + (uint32_t)i; u1 = u1 * 3u + (uint32_t)i; u2 = u2 * 7u + (uint32_t)i; u3 = u3 * 9u + (uint32_t)i; }
          }
          const uint32_t buf = (uint32_t)(i & 1) ^ ((u0 ^ u1 ^ u2 ^ u3) == 0x12345u ? 1u : 0u);
          uint64_t ad = make_desc_sw128(sb + buf * A_BYTES), bd = make_desc_sw128(sb + NBUF * A_BYTES + buf * B_BYTES);
          // conv_s3-like: shifted-window A (unaligned start, SBO = 10 rows), 64-byte-swizzle B
          if (mode & 2) {
            ad = (uint64_t)(((sb + buf * A_BYTES + (uint32_t)(i % 9) * 384u) >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1280 >> 4) << 32) | (1ull << 46) | (2ull << 61);
            bd = (uint64_t)(((sb + NBUF * A_BYTES + buf * B_BYTES) >> 4) & 0x3FFF) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (4ull << 61);
          }
          if (mode & 0x100) {  // the kernel's resident 64 -> 64 3x3 bank: 18 stage tiles of 8 KB after 3 halo buffers of 23 KB
            const uint32_t st = (uint32_t)(i % 18), ab = (uint32_t)((i / 9) % 3), tp = (uint32_t)(i % 9);
            const uint32_t aaddr = sb + ab * 23552u + ((tp / 3u) * 10u + (tp % 3u)) * 128u, baddr = sb + 3u * 23552u + st * 8192u;
            ad = (uint64_t)((aaddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1280 >> 4) << 32) | (1ull << 46) | (2ull << 61);
            bd = (uint64_t)((baddr >> 4) & 0x3FFF) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (4ull << 61);
          }
          // accumulator layout: mode & 4 -> the conv_s3 layout (corr = second half of the main MMA's columns: every MMA
          // overlaps its predecessor's D), else disjoint regions; mode & 8 -> order m(k0) c(k0) m(k1) c(k1)
          uint32_t dm = tmem, dc = (mode & 4) ? tmem + (uint32_t)N2 : tmem + 256u;
          uint32_t first = i ? 1u : 0u;
          if (mode & 0x400) { dm = tmem + (uint32_t)((i / 9) & 1) * 128u; dc = dm + (uint32_t)N2; first = (i % 9) ? 1u : 0u; }
          if (mode & 8) {
            mma_ss<CG>(dm, ad, bd, id1, i ? 1u : 0u);
            if (N2 > 0) mma_ss<CG>(dc, ad + 4u, bd, id2, 1u);
            mma_ss<CG>(dm, ad + 2u, bd + 2u, id1, 1u);
            if (N2 > 0) mma_ss<CG>(dc, ad + 6u, bd + 2u, id2, 1u);
          } else {
            mma_ss<CG>(dm, ad, bd, id1, first);
            mma_ss<CG>(dm, ad + 2u, bd + 2u, id1, 1u);
            if (N2 > 0) {
              mma_ss<CG>(dc, ad + 4u, bd, id2, (i || (mode & 4)) ? 1u : 0u);
              mma_ss<CG>(dc, ad + 6u, bd + 2u, id2, 1u);
            }
          }
        }
        if (CG == 2)
          asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "h"((uint16_t)3) : "memory");
        else
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      }
      __syncwarp();
      uint32_t ok = 0;
      while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
      long long t1 = clock64();
      if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    } else {
      uint32_t ok = 0;  // the peer waits for the multicast commit
      while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    }
    stop_flag = 1;
  } else if (warp >= 4 && (mode & 16)) {
    // what the idle roles of conv_s3 do: spin on an mbarrier that does not complete (try_wait suspends for a while)
    __shared__ uint64_t never;
    if (threadIdx.x == 128) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&never)) : "memory");
    asm volatile("bar.sync 1, 256;" ::: "memory");
    uint32_t ok = 0;
    while (!stop_flag && !ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&never)) : "memory");
  } else if (warp == 4 && (mode & 32)) {
    // the TMA producer of conv_s3: one thread polling with test_wait (never suspends)
    __shared__ uint64_t never2;
    if (threadIdx.x == 128) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&never2)) : "memory");
      uint32_t ok = 0;
      while (!stop_flag && !ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&never2)) : "memory");
    }
  } else if (warp >= 4 && (mode & 1)) {
    // shared-memory streaming beside the MMAs: `rate` x (ld.shared.v4 + st.shared.v4) per thread, then a ~256-clock pause
    const uint32_t base = smem_u32(smem) + NBUF * (A_BYTES + B_BYTES) + (threadIdx.x - 128) * 16u;
    uint32_t x = 0, y = 0, z = 0, w = 0;
    while (!stop_flag) {
      for (int q = 0; q < rate; ++q) {
        uint32_t a, b, c, d;
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(base + (uint32_t)(q & 7) * 4096u));
        x += a; y += b; z += c; w += d;
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(base + (uint32_t)(q & 7) * 4096u), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
      }
      const long long t = clock64();
      while (clock64() - t < 256) {}
    }
    if (x == 0x7fffffffu) cycles[0] = 0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) cluster_sync();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

static int g_smem_kb = 131;
template <int CG>
static void run(int N1, int N2, int sms, long long* dcyc, int mode, int rate, int iters = 4096) {
  const int smem = g_smem_kb * 1024;
  cudaFuncSetAttribute(probe<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(sms);
  cfg.blockDim = dim3(128 + 256);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, probe<CG>, N1, N2, 256, dcyc, mode, rate);
  cudaEventRecord(e0);
  cudaLaunchKernelEx(&cfg, probe<CG>, N1, N2, iters, dcyc, mode, rate);
  cudaEventRecord(e1);
  cudaError_t err = cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  std::vector<long long> cyc(sms);
  cudaMemcpy(cyc.data(), dcyc, sms * sizeof(long long), cudaMemcpyDeviceToHost);
  const double ksteps = (double)iters * 2;
  const double macs = ksteps * 128.0 * (N1 + N2) * 16.0 * sms;  // per CTA: 128 rows x N x 16 per MMA
  printf("{\"cta_group\": %d, \"mode\": %d, \"rate\": %d, \"N1\": %d, \"N2\": %d, \"clk_per_kstep\": %.1f, \"ms\": %.3f, \"dense_TFLOPs\": %.1f, \"err\": \"%s\"}\n",
         CG, mode, rate, N1, N2, cyc[0] / ksteps, ms, 2 * macs / ms / 1e9, cudaGetErrorString(err));
  fflush(stdout);
}

int main(int argc, char** argv) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  long long* dcyc;
  cudaMalloc(&dcyc, sms * sizeof(long long));
  if (argc == 2) {  // sustained: ~3 s of back-to-back launches per configuration, for clock / power sampling from outside
    for (int cg : {1, 2}) {
      for (int rep = 0; rep < 12; ++rep) {
        if (cg == 1) run<1>(128, 64, sms, dcyc, 2, 0, 2000000);
        else run<2>(128, 64, sms, dcyc, 2, 0, 2000000);
      }
    }
    return 0;
  }
  if (argc > 5) {  // towards the kernel: dynamic shared-memory size, operand walk, accumulator hand-over, operand values
    for (int kb : {131, 160, 200, 226}) { g_smem_kb = kb; printf("smem %d KB: ", kb); run<1>(128, 64, sms, dcyc, 2, 0); }
    g_smem_kb = 131;
    for (int m : {2 | 4 | 0x400, 2 | 0x800, 2 | 0x1000, 2 | 0x2000}) run<1>(128, 64, sms, dcyc, m, 0);
    g_smem_kb = 226;
    for (int m : {2 | 0x100, 2 | 4 | 0x100 | 0x400, 2 | 4 | 0x100 | 0x400 | 0x800}) run<1>(128, 64, sms, dcyc, m, 0);
    return 0;
  }
  if (argc > 4) {  // cost of uniform-datapath instructions in the issuing thread
    for (int nx : {0, 4, 8, 16, 32}) {
      run<1>(128, 64, sms, dcyc, 2 | (nx << 8), 0);
      run<1>(128, 64, sms, dcyc, 2 | 64 | (nx << 8), 0);
    }
    return 0;
  }
  if (argc > 3) {  // spinning neighbours
    for (int m : {2, 2 | 16, 2 | 32}) { run<1>(128, 64, sms, dcyc, m, 0); run<1>(64, 32, sms, dcyc, m, 0); }
    return 0;
  }
  if (argc > 2) {  // accumulator layout / issue order study (SS, conv_s3 operands)
    for (int m : {2, 2 | 4, 2 | 8, 2 | 4 | 8}) {
      run<1>(128, 64, sms, dcyc, m, 0);
      run<1>(64, 32, sms, dcyc, m, 0);
      run<1>(192, 96, sms, dcyc, m, 0);
      run<2>(128, 64, sms, dcyc, m, 0);
    }
    return 0;
  }
  const int cfg[][2] = {{128, 64}, {64, 32}, {256, 128}, {192, 96}, {128, 0}, {64, 0}, {256, 0}, {32, 16}, {96, 48}};
  for (auto& c : cfg) run<1>(c[0], c[1], sms, dcyc, 0, 0);
  for (auto& c : cfg) run<2>(c[0], c[1], sms, dcyc, 0, 0);
  for (int mode : {2}) { run<1>(128, 64, sms, dcyc, mode, 0); run<2>(128, 64, sms, dcyc, mode, 0); }
  for (int rate : {1, 2, 4, 8, 16}) { run<1>(128, 64, sms, dcyc, 3, rate); run<2>(128, 64, sms, dcyc, 3, rate); }
  return 0;
}
