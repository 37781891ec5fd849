// Frame metrics of the reference's evaluation loop -- the consumer directly after the hot path:
// PSNR (utils.py:652-660) and the MATLAB-style SSIM (utils.py:663-705: 11x11 Gaussian window, sigma 1.5, 'valid' region) of a
// predicted frame against its ground truth, computed on the tensors DeMFInet.forward returned without a device->host copy of
// the images.  Arithmetic is fp64 throughout, like the reference's numpy code; the 0..255 scaling follows the call site
// main.py:763-766 (prediction: float64, rounded half-to-even; target: float32 arithmetic, not rounded).
//
// HBM-bound and tiny: 2 x 4 B read per element once (halo re-reads hit L1/L2), two doubles written per CTA.
// One CTA = one 16x16 block of window positions of one colour plane: the 26x26 halo is scaled to 0..255 into shared memory,
// filtered horizontally for the five moments (a, b, a^2, b^2, ab), then vertically; per-CTA sums are written to a partials
// array and reduced in a fixed order by a second kernel, so the result is bit-reproducible run to run.
#include "common.cuh"

namespace demfi {

constexpr int MT = 16;          // window positions per tile side
constexpr int MK = 11;          // window size (utils.py:669)
constexpr int MH = MT + MK - 1; // halo side

struct MetricParams {
  double g[MK];                 // cv2.getGaussianKernel(11, 1.5)
  double c1, c2;
  int H, W, C, target_mode;
  int tiles_x, tiles_y;
};

__device__ __forceinline__ double scale_pred(float x) {  // np.around(denorm255_np(float64 x))
  double p = ((double)x + 1.0) / 2.0;
  p = fmin(fmax(p, 0.0), 1.0) * 255.0;
  return rint(p);
}

__device__ __forceinline__ double scale_target(float x, int mode) {
  if (mode == 1) return scale_pred(x);  // second network output instead of a ground truth: same quantisation as the prediction
  float t = __fmul_rn(__fadd_rn(x, 1.0f), 0.5f);  // denorm255_np on a float32 array: (x+1)/2, clip, *255 in float32
  t = __fmul_rn(fminf(fmaxf(t, 0.0f), 1.0f), 255.0f);
  return (double)t;
}

__device__ __forceinline__ double block_sum_256(double v, double* scratch) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = 0.0;
  if (threadIdx.x == 0)
    for (int w = 0; w < 8; ++w) r += scratch[w];
  __syncthreads();
  return r;  // valid in thread 0
}

__global__ void __launch_bounds__(256) frame_metrics_tile_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                                                 MetricParams P, double* __restrict__ partials) {
  __shared__ double sa[MH][MH + 1], sb[MH][MH + 1];
  __shared__ double hm[5][MH][MT];
  __shared__ double red[8];
  const int plane = blockIdx.z;  // b * C + c
  const int ox = blockIdx.x * MT, oy = blockIdx.y * MT;
  const bool last_x = (blockIdx.x == P.tiles_x - 1), last_y = (blockIdx.y == P.tiles_y - 1);
  const float* pp = pred + (size_t)plane * P.H * P.W;
  const float* tp = target + (size_t)plane * P.H * P.W;

  double sse = 0.0;
  for (int i = threadIdx.x; i < MH * MH; i += 256) {
    const int ly = i / MH, lx = i % MH;
    const int y = oy + ly, x = ox + lx;
    double a = 0.0, b = 0.0;
    if (y < P.H && x < P.W) {
      a = scale_target(tp[(size_t)y * P.W + x], P.target_mode);  // img1 = target, img2 = output (main.py:770-771)
      b = scale_pred(pp[(size_t)y * P.W + x]);
      if ((lx < MT || last_x) && (ly < MT || last_y)) {  // every image element is owned by exactly one tile
        const double d = a - b;
        sse += d * d;
      }
    }
    sa[ly][lx] = a;
    sb[ly][lx] = b;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < MH * MT; i += 256) {
    const int ly = i / MT, lx = i % MT;
    double m0 = 0, m1 = 0, m2 = 0, m3 = 0, m4 = 0;
#pragma unroll
    for (int k = 0; k < MK; ++k) {
      const double a = sa[ly][lx + k], b = sb[ly][lx + k], g = P.g[k];
      m0 += g * a;
      m1 += g * b;
      m2 += g * (a * a);
      m3 += g * (b * b);
      m4 += g * (a * b);
    }
    hm[0][ly][lx] = m0; hm[1][ly][lx] = m1; hm[2][ly][lx] = m2; hm[3][ly][lx] = m3; hm[4][ly][lx] = m4;
  }
  __syncthreads();
  double ss = 0.0;
  {
    const int ly = threadIdx.x / MT, lx = threadIdx.x % MT;
    if (oy + ly < P.H - (MK - 1) && ox + lx < P.W - (MK - 1)) {
      double mu1 = 0, mu2 = 0, e11 = 0, e22 = 0, e12 = 0;
#pragma unroll
      for (int k = 0; k < MK; ++k) {
        const double g = P.g[k];
        mu1 += g * hm[0][ly + k][lx];
        mu2 += g * hm[1][ly + k][lx];
        e11 += g * hm[2][ly + k][lx];
        e22 += g * hm[3][ly + k][lx];
        e12 += g * hm[4][ly + k][lx];
      }
      const double mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu1_mu2 = mu1 * mu2;
      const double s1 = e11 - mu1_sq, s2 = e22 - mu2_sq, s12 = e12 - mu1_mu2;
      ss = ((2.0 * mu1_mu2 + P.c1) * (2.0 * s12 + P.c2)) / ((mu1_sq + mu2_sq + P.c1) * (s1 + s2 + P.c2));
    }
  }
  const double tsse = block_sum_256(sse, red);
  const double tss = block_sum_256(ss, red);
  if (threadIdx.x == 0) {
    const size_t blk = ((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    partials[2 * blk] = tsse;
    partials[2 * blk + 1] = tss;
  }
}

// out[b][0] = sum of squared errors over the C*H*W elements of image b, out[b][1] = sum of the SSIM map.
__global__ void __launch_bounds__(256) frame_metrics_reduce_kernel(const double* __restrict__ partials, long long per_image,
                                                                   double* __restrict__ out) {
  __shared__ double red[8];
  const double* p = partials + 2 * per_image * blockIdx.x;
  double a = 0.0, b = 0.0;
  for (long long i = threadIdx.x; i < per_image; i += 256) {
    a += p[2 * i];
    b += p[2 * i + 1];
  }
  const double ta = block_sum_256(a, red);
  const double tb = block_sum_256(b, red);
  if (threadIdx.x == 0) {
    out[2 * blockIdx.x] = ta;
    out[2 * blockIdx.x + 1] = tb;
  }
}

}  // namespace demfi

using namespace demfi;

extern "C" {

int64_t demfi_frame_metrics_workspace(int32_t B, int32_t C, int32_t H, int32_t W) {
  if (B <= 0 || C <= 0 || H < MK || W < MK) return -1;
  const int64_t tx = (W - (MK - 1) + MT - 1) / MT, ty = (H - (MK - 1) + MT - 1) / MT;
  return (int64_t)B * C * tx * ty * 2 * (int64_t)sizeof(double);
}

int demfi_frame_metrics(const float* pred, const float* target, int32_t B, int32_t C, int32_t H, int32_t W, int32_t target_mode,
                        void* workspace, int64_t workspace_bytes, double* out, void* stream) {
  if (check_device()) return 3;
  DEMFI_REQUIRE(pred && target && out && workspace, "frame_metrics: null pointer");
  DEMFI_REQUIRE(B > 0 && C > 0 && H >= MK && W >= MK, "frame_metrics: images must be at least %d x %d (the SSIM window)", MK, MK);
  DEMFI_REQUIRE(target_mode == 0 || target_mode == 1, "frame_metrics: target_mode must be 0 (ground truth) or 1 (second prediction)");
  DEMFI_REQUIRE(workspace_bytes >= demfi_frame_metrics_workspace(B, C, H, W), "frame_metrics: workspace too small");
  DEMFI_REQUIRE((long long)B * C <= 65535, "frame_metrics: too many planes");
  MetricParams P;
  double sum = 0.0;
  for (int i = 0; i < MK; ++i) {  // cv2.getGaussianKernel(11, 1.5): exp(-0.5 x^2 / sigma^2), normalised
    const double x = i - (MK - 1) * 0.5;
    P.g[i] = exp(-0.5 * x * x / (1.5 * 1.5));
    sum += P.g[i];
  }
  for (int i = 0; i < MK; ++i) P.g[i] /= sum;
  P.c1 = (0.01 * 255) * (0.01 * 255);
  P.c2 = (0.03 * 255) * (0.03 * 255);
  P.H = H; P.W = W; P.C = C; P.target_mode = target_mode;
  P.tiles_x = (W - (MK - 1) + MT - 1) / MT;
  P.tiles_y = (H - (MK - 1) + MT - 1) / MT;
  DEMFI_REQUIRE(P.tiles_y <= 65535, "frame_metrics: image too tall");
  dim3 grid(P.tiles_x, P.tiles_y, B * C);
  frame_metrics_tile_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(pred, target, P, (double*)workspace);
  DEMFI_LAUNCH_CHECK("frame_metrics_tile");
  frame_metrics_reduce_kernel<<<B, 256, 0, (cudaStream_t)stream>>>((const double*)workspace, (long long)C * P.tiles_x * P.tiles_y, out);
  DEMFI_LAUNCH_CHECK("frame_metrics_reduce");
  return 0;
}

}  // extern "C"
