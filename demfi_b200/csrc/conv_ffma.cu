// CUDA-core (FFMA) implicit-GEMM convolution, NHWC fp32, exact fp32 arithmetic.
//
// Covers every nn.Conv call of DeMFInet.py including the shapes the tcgen05
// kernel does not take: the stride-2 4x4 UNet encoders (DeMFInet.py:575-577), the UNet decoders
// that read through nearest x2 up-sampling + concat (DeMFInet.py:591-602) and the convs with a
// handful of input or output channels (Mixer.conv_delta1 5->32, w_gen_2 64->1, Dec_last2 64->3).
//
// Tiling: one CTA = 8x16 output pixels (M=128) x BN output channels, 256 threads, K consumed in
// steps of (tap, 16 input channels).  A tile is transposed into shared memory ([k][pixel]) so
// that the inner product reads pixels with 128-bit LDS; global loads are 128-bit per pixel.
// Register prefetch + two shared-memory stages: one __syncthreads per K step.
#include "common.cuh"

namespace demfi {

constexpr int FF_BM = 128;  // pixels per CTA: 8 rows x 16 cols
constexpr int FF_TH = 8;
constexpr int FF_TW = 16;
constexpr int FF_BK = 16;
constexpr int FF_THREADS = 256;

template <int BN>
__global__ void __launch_bounds__(FF_THREADS)
conv_ffma_kernel(const __grid_constant__ demfi_conv_t p, int tiles_x, int tiles_y) {
  constexpr int NTX = BN / 4;             // threads across output channels
  constexpr int NTY = FF_THREADS / NTX;   // threads across pixels
  constexpr int TM = FF_BM / NTY;         // pixels per thread (8, 4, 2)
  constexpr int BF4 = FF_BK * BN / 4;     // float4 per B stage
  __shared__ __align__(16) float As[2][FF_BK][FF_BM + 4];
  __shared__ __align__(16) float Bs[2][FF_BK][BN];

  const int tid = threadIdx.x;
  int tile = blockIdx.x;
  const int tx0 = (tile % tiles_x) * FF_TW;
  tile /= tiles_x;
  const int ty0 = (tile % tiles_y) * FF_TH;
  const int n = tile / tiles_y;
  const int co0 = blockIdx.y * BN;

  // loader role: one pixel, two float4 (8 of the 16 channels of the step)
  const int lp = tid & (FF_BM - 1);
  const int lc = (tid >> 7) * 8;
  const int loy = ty0 + (lp >> 4), lox = tx0 + (lp & 15);
  const bool lvalid = (loy < p.H) && (lox < p.W);
  // B loader role
  const int bk = tid / (BN / 4), bc4 = tid % (BN / 4);
  const bool bactive = tid < BF4;

  // compute role
  const int ctx = tid % NTX, cty = tid / NTX;
  float acc[TM][4];
#pragma unroll
  for (int i = 0; i < TM; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }

  const int taps = p.KH * p.KW;
  int k_total = 0, nsteps = 0;
  for (int s = 0; s < p.nsrc; ++s) {
    k_total += p.src[s].C;
    nsteps += ((p.src[s].C + FF_BK - 1) / FF_BK) * taps;
  }

  // step decoder state: (source, channel chunk, tap), taps innermost
  int s_src = 0, s_c0 = 0, s_tap = 0, s_kbase = 0;
  float4 ra0, ra1, rb;

  auto load_global = [&]() {
    const demfi_src_t& S = p.src[s_src];
    const int ky = s_tap / p.KW, kx = s_tap - ky * p.KW;
    ra0 = make_float4(0.f, 0.f, 0.f, 0.f);
    ra1 = ra0;
    rb = ra0;
    if (lvalid) {
      int iy = loy * p.stride + ky - p.pad_h;
      int ix = lox * p.stride + kx - p.pad_w;
      if (iy >= 0 && iy < p.Hi && ix >= 0 && ix < p.Wi) {
        iy >>= S.up;
        ix >>= S.up;
        const int Hs = p.Hi >> S.up, Ws = p.Wi >> S.up;
        const float* src = S.ptr + (((size_t)n * Hs + iy) * Ws + ix) * (size_t)S.ld;
        const int c = s_c0 + lc;
        if (c < S.C) ra0 = __ldg(reinterpret_cast<const float4*>(src + c));
        if (c + 4 < S.C) ra1 = __ldg(reinterpret_cast<const float4*>(src + c + 4));
      }
    }
    if (bactive) {
      const int c = s_c0 + bk;
      const int co = co0 + bc4 * 4;
      if (c < S.C && co < p.cout_pad)
        rb = __ldg(reinterpret_cast<const float4*>(p.wpack + ((size_t)s_tap * k_total + s_kbase + c) * p.cout_pad + co));
    }
  };
  auto advance = [&]() {
    if (++s_tap == taps) {
      s_tap = 0;
      s_c0 += FF_BK;
      if (s_c0 >= p.src[s_src].C) {
        s_kbase += p.src[s_src].C;
        s_c0 = 0;
        ++s_src;
      }
    }
  };
  auto store_smem = [&](int buf) {
    As[buf][lc + 0][lp] = ra0.x; As[buf][lc + 1][lp] = ra0.y; As[buf][lc + 2][lp] = ra0.z; As[buf][lc + 3][lp] = ra0.w;
    As[buf][lc + 4][lp] = ra1.x; As[buf][lc + 5][lp] = ra1.y; As[buf][lc + 6][lp] = ra1.z; As[buf][lc + 7][lp] = ra1.w;
    if (bactive) *reinterpret_cast<float4*>(&Bs[buf][bk][bc4 * 4]) = rb;
  };

  load_global();
  advance();
  store_smem(0);
  __syncthreads();

  for (int it = 0; it < nsteps; ++it) {
    const int buf = it & 1;
    const bool more = (it + 1 < nsteps);
    if (more) { load_global(); advance(); }
#pragma unroll
    for (int kk = 0; kk < FF_BK; ++kk) {
      float a[TM];
      if constexpr (TM % 4 == 0) {
#pragma unroll
        for (int i = 0; i < TM; i += 4) {
          const float4 t4 = *reinterpret_cast<const float4*>(&As[buf][kk][cty * TM + i]);
          a[i] = t4.x; a[i + 1] = t4.y; a[i + 2] = t4.z; a[i + 3] = t4.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < TM; ++i) a[i] = As[buf][kk][cty * TM + i];
      }
      const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][kk][ctx * 4]);
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        acc[i][0] = fmaf(a[i], b.x, acc[i][0]);
        acc[i][1] = fmaf(a[i], b.y, acc[i][1]);
        acc[i][2] = fmaf(a[i], b.z, acc[i][2]);
        acc[i][3] = fmaf(a[i], b.w, acc[i][3]);
      }
    }
    if (more) store_smem(buf ^ 1);
    __syncthreads();
  }

  const int co = co0 + ctx * 4;
  if (co >= p.cout_pad) return;
  const float4 bias = ld4(p.bias + co);
#pragma unroll 1
  for (int s = 0; s < p.nseg; ++s) {
    if (co < p.seg[s].ch0 || co >= p.seg[s].ch0 + p.seg[s].nch) continue;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const int px = cty * TM + i;
      const int oy = ty0 + (px >> 4), ox = tx0 + (px & 15);
      if (oy < p.H && ox < p.W) {
        const SegCursor cur = seg_cursor(p, p.seg[s], n, oy, ox);
        seg_emit4(cur, co, make_float4(acc[i][0] + bias.x, acc[i][1] + bias.y, acc[i][2] + bias.z, acc[i][3] + bias.w));
      }
    }
  }
}

int launch_conv_ffma(const demfi_conv_t& c, cudaStream_t st) {
  const int tiles_x = (c.W + FF_TW - 1) / FF_TW, tiles_y = (c.H + FF_TH - 1) / FF_TH;
  const long long tiles = (long long)tiles_x * tiles_y * c.N;
  DEMFI_REQUIRE(tiles > 0 && tiles < (1ll << 31), "conv_ffma: bad tile count %lld", tiles);
  const int bn = c.cout_pad >= 64 ? 64 : (c.cout_pad > 16 ? 32 : 16);
  dim3 grid((unsigned)tiles, (unsigned)((c.cout_pad + bn - 1) / bn));
  if (bn == 64) conv_ffma_kernel<64><<<grid, FF_THREADS, 0, st>>>(c, tiles_x, tiles_y);
  else if (bn == 32) conv_ffma_kernel<32><<<grid, FF_THREADS, 0, st>>>(c, tiles_x, tiles_y);
  else conv_ffma_kernel<16><<<grid, FF_THREADS, 0, st>>>(c, tiles_x, tiles_y);
  DEMFI_LAUNCH_CHECK("conv_ffma");
  return 0;
}

}  // namespace demfi
