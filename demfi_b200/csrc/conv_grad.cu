// Backward pieces of a stride-1 convolution layer y = act(conv(x, W) + b) -- first correct CUDA path of the training row
// (the reference gets these from autograd through nn.Conv2d, main.py:443):
//   * dz = dy * act'(y)                       act_backward_kernel (ReLU / tanh / sigmoid from the stored output)
//   * dx = conv(dz, W rotated 180, Cin<->Cout) runs on the FORWARD tensor-core kernel (demfi_conv2d) with weights re-packed by
//                                             the host (demfi_b200/grad.py), so it needs nothing here
//   * dW[co,ci,ky,kx] = sum_p dz[p,co] * x[p + (ky,kx) - pad, ci],  db[co] = sum_p dz[p,co]      conv_wgrad_kernel
// conv_wgrad_kernel is the CUDA-core version (fp32 FFMA, 64 x 64 output tile per CTA and tap, pixels as the reduction
// dimension, one fp32 atomic add per output and pixel chunk): the same role conv_ffma plays for the forward -- correct first,
// and later the independent implementation the tensor-core wgrad is tested against.  Algorithmic work = the forward's MACs.
#include "common.cuh"

namespace demfi {

constexpr int WG_T = 64;     // output tile: WG_T output channels x WG_T input channels
constexpr int WG_KP = 16;    // pixels per shared-memory stage
constexpr int WG_CHUNK = 1024;  // pixels reduced by one CTA before its atomic adds (2048: 576 CTAs of a 64 -> 64 3x3 layer at 2 x 256 x 256 = 1.3 waves of the 444 resident CTAs)

struct WgradParams {
  const float* x; const float* dz; float* dw; float* db;
  int x_ld, dz_ld, Cin, Cout, N, H, W, Hi, Wi, stride, KH, KW, pad_h, pad_w;  // H, W: output (dz) plane; Hi, Wi: input (x) plane
  int co_blocks, ci_blocks;
  long long npix;
};

__global__ void __launch_bounds__(256) conv_wgrad_kernel(const WgradParams P) {
  __shared__ __align__(16) float zs[WG_KP][WG_T];
  __shared__ __align__(16) float xs[WG_KP][WG_T];
  const int tap = blockIdx.y, ky = tap / P.KW, kx = tap % P.KW;
  const int cb = blockIdx.z / P.ci_blocks, ib = blockIdx.z % P.ci_blocks;
  const int co0 = cb * WG_T, ci0 = ib * WG_T;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  const bool do_bias = (P.db != nullptr && tap == 0 && ib == 0);
  float acc[4][4] = {};
  float bsum[4] = {};
  const long long p0 = (long long)blockIdx.x * WG_CHUNK;
  const long long p1 = min(p0 + (long long)WG_CHUNK, P.npix);
  const int lc = threadIdx.x & 63, lr0 = threadIdx.x >> 6;  // loader: channel lc of rows lr0, lr0 + 4, ...
  for (long long pb = p0; pb < p1; pb += WG_KP) {
#pragma unroll
    for (int r = lr0; r < WG_KP; r += 4) {
      const long long p = pb + r;
      float zv = 0.0f, xv = 0.0f;
      if (p < p1) {
        if (co0 + lc < P.Cout) zv = P.dz[p * P.dz_ld + co0 + lc];
        const int xw = (int)(p % P.W), yh = (int)((p / P.W) % P.H);
        const long long n = p / ((long long)P.W * P.H);
        const int sy = yh * P.stride + ky - P.pad_h, sx = xw * P.stride + kx - P.pad_w;
        if (ci0 + lc < P.Cin && sy >= 0 && sy < P.Hi && sx >= 0 && sx < P.Wi)
          xv = P.x[((n * P.Hi + sy) * P.Wi + sx) * P.x_ld + ci0 + lc];
      }
      zs[r][lc] = zv;
      xs[r][lc] = xv;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < WG_KP; ++r) {
      const float4 a = *reinterpret_cast<const float4*>(&zs[r][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&xs[r][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        bsum[i] += av[i];
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = co0 + ty * 4 + i;
    if (co >= P.Cout) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ci = ci0 + tx * 4 + j;
      if (ci < P.Cin) atomicAdd(&P.dw[(((size_t)co * P.Cin + ci) * P.KH + ky) * P.KW + kx], acc[i][j]);
    }
    if (do_bias && tx == 0) atomicAdd(&P.db[co], bsum[i]);
  }
}

// ---- tensor-core wgrad (legacy mma.sync path: m16n8k8 TF32 fragments, fp32 accumulate) -------------------------------------
// Same decomposition as conv_wgrad_kernel -- one CTA per (pixel chunk, tap, 64 x 64 block of [Cout, Cin]), pixels as the reduction
// dimension, fp32 atomic adds per chunk -- with the inner product on the tensor cores.  fp32 parity through the 3xTF32 split
// (a = ah + al, ah = tf32(a), al = tf32(a - ah): D += al * bh + ah * bl + ah * bh; the dropped al * bl term is 2^-22 relative), the
// small products first.  GEMM view per CTA: D[co, ci] += sum_p A[co, p] * B[p, ci] with A[m][k] = dz[p0 + k][co0 + m] and
// B[k][n] = x[p0 + k shifted by the tap][ci0 + n], both staged in shared memory as [pixel][channel] rows of 72 floats (the fragment
// loads (k = t or t + 4, channel = g) then hit 32 distinct banks).  8 warps: warp w owns rows (w & 3) * 16 .. + 16 of the 64
// output channels and columns (w >> 2) * 32 .. + 32 of the 64 input channels = four m16n8 accumulator tiles.
// tcgen05 would need the pixel-major operands as MN-major UMMA tiles ("What comes next"); this kernel is the step from
// 9-18 TFLOP/s on the CUDA cores to the legacy tensor-core path with the launch structure unchanged.
constexpr int WM_KP = 32;   // pixels per shared-memory stage (four k8 steps)
constexpr int WM_LD = 72;   // row stride in floats: 72 mod 32 = 8

__device__ __forceinline__ uint32_t tf32_rna(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(256) conv_wgrad_mma_kernel(const WgradParams P) {
  __shared__ __align__(16) float zs[WM_KP][WM_LD];
  __shared__ __align__(16) float xs[WM_KP][WM_LD];
  const int tap = blockIdx.y, ky = tap / P.KW, kx = tap % P.KW;
  const int cb = blockIdx.z / P.ci_blocks, ib = blockIdx.z % P.ci_blocks;
  const int co0 = cb * WG_T, ci0 = ib * WG_T;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int m0 = (warp & 3) * 16, n0 = (warp >> 2) * 32;
  const bool do_bias = (P.db != nullptr && tap == 0 && ib == 0 && (warp >> 2) == 0);
  // acc: the tensor core's running sums of ONE stage (12 chained MMAs per tile); tot: their fp32 round-to-nearest sum over the
  // chunk.  The tensor core accumulates with truncation -- a gain error of ~0.27 * 2^-24 per chained MMA (profiles/r1_tc_numerics.md):
  // over the 768 MMAs of a 2048-pixel chunk that was 1.2e-5 of dW (measured), over 12 it is 2e-7.
  float acc[4][4] = {}, tot[4][4] = {};
  float bsum0 = 0.0f, bsum1 = 0.0f;   // rows m0 + g and m0 + g + 8 over this lane's k
  const long long p0 = (long long)blockIdx.x * WG_CHUNK;
  const long long p1 = min(p0 + (long long)WG_CHUNK, P.npix);
  const int lc = threadIdx.x & 63, lr0 = threadIdx.x >> 6;  // loader: channel lc of rows lr0, lr0 + 4, ... (8 rows per stage)
  const bool z_ok = co0 + lc < P.Cout, x_ok = ci0 + lc < P.Cin;
  // pixel coordinates of the loader's next row, advanced incrementally (one division per CTA, not three per row: with them the
  // index arithmetic, not the inner product, set the pace of this kernel and of the CUDA-core one alike)
  int cx, cy;
  long long cn;
  {
    const long long p = p0 + lr0;
    cx = (int)(p % P.W);
    cy = (int)((p / P.W) % P.H);
    cn = p / ((long long)P.W * P.H);
  }
  float zr[WM_KP / 4], xr[WM_KP / 4];
  auto load_stage = [&](long long pb) {  // global -> registers, rows pb + lr0 + 4 i
#pragma unroll
    for (int i = 0; i < WM_KP / 4; ++i) {
      const long long p = pb + lr0 + 4 * i;
      float zv = 0.0f, xv = 0.0f;
      if (p < p1) {
        if (z_ok) zv = P.dz[p * P.dz_ld + co0 + lc];
        const int sy = cy * P.stride + ky - P.pad_h, sx = cx * P.stride + kx - P.pad_w;
        if (x_ok && sy >= 0 && sy < P.Hi && sx >= 0 && sx < P.Wi) xv = P.x[((cn * P.Hi + sy) * P.Wi + sx) * P.x_ld + ci0 + lc];
      }
      zr[i] = zv;
      xr[i] = xv;
      cx += 4;
      while (cx >= P.W) {
        cx -= P.W;
        if (++cy == P.H) { cy = 0; ++cn; }
      }
    }
  };
  load_stage(p0);
  for (long long pb = p0; pb < p1; pb += WM_KP) {
#pragma unroll
    for (int i = 0; i < WM_KP / 4; ++i) {
      zs[lr0 + 4 * i][lc] = zr[i];
      xs[lr0 + 4 * i][lc] = xr[i];
    }
    __syncthreads();
    if (pb + WM_KP < p1) load_stage(pb + WM_KP);  // the next stage's loads are in flight during this stage's MMAs
#pragma unroll
    for (int k0 = 0; k0 < WM_KP; k0 += 8) {
      const float af[4] = {zs[k0 + t][m0 + g], zs[k0 + t][m0 + g + 8], zs[k0 + t + 4][m0 + g], zs[k0 + t + 4][m0 + g + 8]};
      uint32_t ah[4], al[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        ah[i] = tf32_rna(af[i]);
        al[i] = tf32_rna(af[i] - __uint_as_float(ah[i]));
      }
      bsum0 += af[0] + af[2];
      bsum1 += af[1] + af[3];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float b0f = xs[k0 + t][n0 + 8 * j + g], b1f = xs[k0 + t + 4][n0 + 8 * j + g];
        const uint32_t b0h = tf32_rna(b0f), b1h = tf32_rna(b1f);
        const uint32_t b0l = tf32_rna(b0f - __uint_as_float(b0h)), b1l = tf32_rna(b1f - __uint_as_float(b1h));
        mma_tf32(acc[j], al, b0h, b1h);
        mma_tf32(acc[j], ah, b0l, b1l);
        mma_tf32(acc[j], ah, b0h, b1h);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        tot[j][i] += acc[j][i];
        acc[j][i] = 0.0f;
      }
    __syncthreads();
  }
  // accumulator tile j: c0 (row g, col 2t), c1 (g, 2t + 1), c2 (g + 8, 2t), c3 (g + 8, 2t + 1)
#pragma unroll
  for (int j = 0; j < 4; ++j) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int co = co0 + m0 + g + (i >> 1) * 8, ci = ci0 + n0 + 8 * j + 2 * t + (i & 1);
      if (co < P.Cout && ci < P.Cin) atomicAdd(&P.dw[(((size_t)co * P.Cin + ci) * P.KH + ky) * P.KW + kx], tot[j][i]);
    }
  }
  if (do_bias) {  // the four lanes of a quad hold disjoint k of the same two rows
    bsum0 += __shfl_xor_sync(0xffffffffu, bsum0, 1);
    bsum0 += __shfl_xor_sync(0xffffffffu, bsum0, 2);
    bsum1 += __shfl_xor_sync(0xffffffffu, bsum1, 1);
    bsum1 += __shfl_xor_sync(0xffffffffu, bsum1, 2);
    if (t == 0) {
      if (co0 + m0 + g < P.Cout) atomicAdd(&P.db[co0 + m0 + g], bsum0);
      if (co0 + m0 + g + 8 < P.Cout) atomicAdd(&P.db[co0 + m0 + g + 8], bsum1);
    }
  }
}

// dx of a STRIDED convolution (the three 4x4 stride-2 UNet encoders, DeMFInet.py:566-571), gather form on CUDA cores:
// dx[n, yi, xi, ci] = sum over taps (ky, kx) with (yi + pad - ky) and (xi + pad - kx) divisible by the stride, and over co, of
// dz[n, (yi + pad - ky) / s, (xi + pad - kx) / s, co] * W[co, ci, ky, kx].  One thread per (input pixel, input channel); the 32
// lanes of a warp share the pixel, so the dz reads are broadcasts.  Three small layers: correct first.
__global__ void __launch_bounds__(256)
conv_dgrad_strided_kernel(const float* __restrict__ dz, int dz_ld, const float* __restrict__ w, int Cin, int Cout, int N, int H, int W,
                          int Hi, int Wi, int KH, int KW, int pad_h, int pad_w, int stride, float* __restrict__ dx, int dx_ld) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)N * Hi * Wi * Cin) return;
  const int ci = (int)(i % Cin);
  const long long p = i / Cin;
  const int xi = (int)(p % Wi), yi = (int)((p / Wi) % Hi);
  const long long n = p / ((long long)Wi * Hi);
  float acc = 0.0f;
  for (int ky = 0; ky < KH; ++ky) {
    const int ty = yi + pad_h - ky;
    if (ty < 0 || ty % stride != 0 || ty / stride >= H) continue;
    for (int kx = 0; kx < KW; ++kx) {
      const int tx = xi + pad_w - kx;
      if (tx < 0 || tx % stride != 0 || tx / stride >= W) continue;
      const float* z = dz + ((n * H + ty / stride) * W + tx / stride) * dz_ld;
      const float* wp = w + ((size_t)ci * KH + ky) * KW + kx;
      for (int co = 0; co < Cout; ++co) acc = fmaf(__ldg(z + co), __ldg(wp + (size_t)co * Cin * KH * KW), acc);
    }
  }
  dx[p * dx_ld + ci] = acc;
}

// dz = dy * act'(y), y the layer's stored OUTPUT (ReLU: y > 0; tanh: 1 - y^2; sigmoid: y (1 - y)); C channels of NHWC rows
__global__ void __launch_bounds__(256) act_backward_kernel(const float* __restrict__ dy, int dy_ld, const float* __restrict__ y, int y_ld,
                                                           long long npix, int C, int act, float* __restrict__ out, int out_ld) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix * C) return;
  const long long p = i / C;
  const int c = (int)(i % C);
  const float g = dy[p * dy_ld + c];
  float d = 1.0f;
  if (act != DEMFI_ACT_NONE) {
    const float v = y[p * y_ld + c];
    d = act == DEMFI_ACT_RELU ? (v > 0.0f ? 1.0f : 0.0f) : act == DEMFI_ACT_TANH ? 1.0f - v * v : v * (1.0f - v);
  }
  out[p * out_ld + c] = g * d;
}

}  // namespace demfi

using namespace demfi;

extern "C" {

int demfi_conv2d_wgrad(const float* x, int32_t x_ld, int32_t Cin, const float* dz, int32_t dz_ld, int32_t Cout, int32_t N,
                       int32_t H, int32_t W, int32_t KH, int32_t KW, int32_t pad_h, int32_t pad_w, int32_t stride, float* dw,
                       float* dbias, void* stream) {
  if (check_device()) return 3;
  DEMFI_REQUIRE(x && dz && dw, "conv2d_wgrad: null pointer");
  DEMFI_REQUIRE(N > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && KH > 0 && KW > 0 && KH * KW <= 65535 && x_ld >= Cin && dz_ld >= Cout &&
                    pad_h >= 0 && pad_w >= 0, "conv2d_wgrad: bad shape");
  DEMFI_REQUIRE(stride == 1 || stride == 2, "conv2d_wgrad: stride 1 or 2");
  // the input plane is the one DeMFI-Net's layers see: 'same' for stride 1, exactly stride x the output for the stride-2 encoders
  const int Hi = H * stride, Wi = W * stride;
  DEMFI_REQUIRE((Hi + 2 * pad_h - KH) / stride + 1 == H && (Wi + 2 * pad_w - KW) / stride + 1 == W,
                "conv2d_wgrad: kernel / padding / stride do not map a %d x %d input to the %d x %d output", Hi, Wi, H, W);
  WgradParams P;
  P.x = x; P.dz = dz; P.dw = dw; P.db = dbias;
  P.x_ld = x_ld; P.dz_ld = dz_ld; P.Cin = Cin; P.Cout = Cout; P.N = N; P.H = H; P.W = W; P.KH = KH; P.KW = KW;
  P.pad_h = pad_h; P.pad_w = pad_w; P.Hi = Hi; P.Wi = Wi; P.stride = stride;
  P.co_blocks = (Cout + WG_T - 1) / WG_T;
  P.ci_blocks = (Cin + WG_T - 1) / WG_T;
  P.npix = (long long)N * H * W;
  DEMFI_REQUIRE(P.co_blocks * P.ci_blocks <= 65535, "conv2d_wgrad: too many channel blocks");
  dim3 grid((unsigned)((P.npix + WG_CHUNK - 1) / WG_CHUNK), KH * KW, P.co_blocks * P.ci_blocks);
  // wgrad_kind 1 (default): the mma.sync 3xTF32 kernel; 0: the CUDA-core kernel (the independent fp32 implementation of the tests)
  if (get_option("wgrad_kind") != 0) conv_wgrad_mma_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(P);
  else conv_wgrad_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(P);
  DEMFI_LAUNCH_CHECK("conv_wgrad");
  return 0;
}

int demfi_conv2d_dgrad_strided(const float* dz, int32_t dz_ld, const float* w_oihw, int32_t Cin, int32_t Cout, int32_t N, int32_t H,
                               int32_t W, int32_t KH, int32_t KW, int32_t pad_h, int32_t pad_w, int32_t stride, float* dx,
                               int32_t dx_ld, void* stream) {
  if (check_device()) return 3;
  DEMFI_REQUIRE(dz && w_oihw && dx && N > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && KH > 0 && KW > 0 && stride >= 1 && dz_ld >= Cout &&
                    dx_ld >= Cin, "conv2d_dgrad_strided: bad arguments");
  const int Hi = H * stride, Wi = W * stride;
  DEMFI_REQUIRE((Hi + 2 * pad_h - KH) / stride + 1 == H && (Wi + 2 * pad_w - KW) / stride + 1 == W,
                "conv2d_dgrad_strided: kernel / padding / stride do not map a %d x %d input to the %d x %d output", Hi, Wi, H, W);
  const long long n = (long long)N * Hi * Wi * Cin;
  conv_dgrad_strided_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(dz, dz_ld, w_oihw, Cin, Cout, N, H, W, Hi, Wi, KH, KW,
                                                                                           pad_h, pad_w, stride, dx, dx_ld);
  DEMFI_LAUNCH_CHECK("conv_dgrad_strided");
  return 0;
}

int demfi_act_backward(const float* dy, int32_t dy_ld, const float* y, int32_t y_ld, int64_t npix, int32_t C, int32_t act,
                       float* out, int32_t out_ld, void* stream) {
  if (check_device()) return 3;
  DEMFI_REQUIRE(dy && out && npix > 0 && C > 0 && dy_ld >= C && out_ld >= C, "act_backward: bad arguments");
  DEMFI_REQUIRE(act == DEMFI_ACT_NONE || ((act == DEMFI_ACT_RELU || act == DEMFI_ACT_TANH || act == DEMFI_ACT_SIGMOID) && y && y_ld >= C),
                "act_backward: activation %d has no backward here (or y is missing)", act);
  act_backward_kernel<<<(unsigned)((npix * C + 255) / 256), 256, 0, (cudaStream_t)stream>>>(dy, dy_ld, y, y_ld, npix, C, act, out, out_ld);
  DEMFI_LAUNCH_CHECK("act_backward");
  return 0;
}

}  // extern "C"
