// Small operators of the training step (first correct path), all HBM-bound and elementwise:
//   * demfi_l1_sum            sum |pred - target| of one tensor pair (nn.L1Loss of main.py:404-440 before the mean) and, fused,
//                             the gradient  scale * sign(pred - target)  that autograd would send back into the network
//   * demfi_adam_step         torch.optim.Adam (main.py:179-180: betas 0.9 / 0.999, eps 1e-8, optional weight decay), one tensor
//   * demfi_fgac_blend_backward   Eq.(4) out = w src + (1-w) e  (DeMFInet.py:452): dw (channel reduction), dsrc, de
//   * demfi_upsample2x_backward   nn.UpsamplingNearest2d(2): the gradient of a source pixel is the sum over its 2x2 children
#include "common.cuh"

namespace demfi {

__device__ __forceinline__ double block_sum_d256(double v, double* scratch) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = 0.0;
  if (threadIdx.x == 0)
    for (int w = 0; w < 8; ++w) r += scratch[w];
  __syncthreads();
  return r;
}

constexpr int L1_PER_BLOCK = 256 * 16;

// fixed work split and fixed summation order: the same inputs give the same bits (partials summed by l1_reduce_kernel)
__global__ void __launch_bounds__(256) l1_partial_kernel(const float* __restrict__ pred, const float* __restrict__ target, long long n,
                                                         float grad_scale, float* __restrict__ grad, double* __restrict__ partials) {
  __shared__ double red[8];
  const long long base = (long long)blockIdx.x * L1_PER_BLOCK;
  double s = 0.0;
#pragma unroll 4
  for (int k = 0; k < 16; ++k) {
    const long long i = base + (long long)k * 256 + threadIdx.x;
    if (i < n) {
      const float d = pred[i] - target[i];
      s += (double)fabsf(d);
      if (grad != nullptr) grad[i] = d > 0.0f ? grad_scale : d < 0.0f ? -grad_scale : 0.0f;  // sign(0) = 0, as torch's L1 backward
    }
  }
  const double t = block_sum_d256(s, red);
  if (threadIdx.x == 0) partials[blockIdx.x] = t;
}

__global__ void __launch_bounds__(256) l1_reduce_kernel(const double* __restrict__ partials, long long nblocks, double* __restrict__ out) {
  __shared__ double red[8];
  double s = 0.0;
  for (long long i = threadIdx.x; i < nblocks; i += 256) s += partials[i];
  const double t = block_sum_d256(s, red);
  if (threadIdx.x == 0) out[0] = t;
}

__global__ void __launch_bounds__(256) adam_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                        float* __restrict__ v, long long n, float lr, float beta1, float beta2,
                                                        float eps, float weight_decay, float bias1, float bias2_sqrt) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float gi = g[i];
  const float pi = p[i];
  if (weight_decay != 0.0f) gi = fmaf(weight_decay, pi, gi);
  const float mi = m[i] + (gi - m[i]) * (1.0f - beta1);             // exp_avg.lerp_(grad, 1 - beta1)
  const float vi = v[i] * beta2 + (1.0f - beta2) * gi * gi;         // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
  m[i] = mi;
  v[i] = vi;
  const float denom = sqrtf(vi) / bias2_sqrt + eps;
  p[i] = pi - (lr / bias1) * (mi / denom);
}

__global__ void __launch_bounds__(256)
fgac_blend_bwd_kernel(const float* __restrict__ w, int w_ld, const float* __restrict__ src, int src_ld, const float* __restrict__ e,
                      int e_ld, const float* __restrict__ gout, int gout_ld, long long npix, int C4, float* __restrict__ dw, int dw_ld,
                      float* __restrict__ dsrc, int dsrc_ld, float* __restrict__ de, int de_ld) {
  const long long p = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
  const int l = threadIdx.x & 15;
  float s = 0.0f;
  if (p < npix) {
    const float ww = __ldg(w + p * w_ld);
    for (int q = l; q < C4; q += 16) {
      const float4 g = ld4(gout + p * gout_ld + 4 * q);
      const float4 a = ld4(src + p * src_ld + 4 * q);
      const float4 b = ld4(e + p * e_ld + 4 * q);
      s += (g.x * (a.x - b.x) + g.y * (a.y - b.y)) + (g.z * (a.z - b.z) + g.w * (a.w - b.w));
      if (dsrc != nullptr) st4(dsrc + p * dsrc_ld + 4 * q, make_float4(g.x * ww, g.y * ww, g.z * ww, g.w * ww));
      const float u = 1.0f - ww;
      if (de != nullptr) st4(de + p * de_ld + 4 * q, make_float4(g.x * u, g.y * u, g.z * u, g.w * u));
    }
  }
  for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (p < npix && l == 0 && dw != nullptr) dw[p * dw_ld] = s;
}

__global__ void __launch_bounds__(256) upsample2x_bwd_kernel(const float* __restrict__ gdst, int gdst_ld, int B, int Hs, int Ws, int C4,
                                                             float* __restrict__ gsrc, int gsrc_ld) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * Hs * Ws * C4) return;
  const int q = (int)(i % C4);
  const long long p = i / C4;
  const int x = (int)(p % Ws), y = (int)((p / Ws) % Hs), n = (int)(p / ((long long)Ws * Hs));
  const size_t row = (size_t)2 * Ws;
  const size_t d0 = ((size_t)n * 2 * Hs + 2 * y) * row + 2 * x;
  const float4 a = ld4(gdst + d0 * gdst_ld + 4 * q), b = ld4(gdst + (d0 + 1) * gdst_ld + 4 * q);
  const float4 c = ld4(gdst + (d0 + row) * gdst_ld + 4 * q), d = ld4(gdst + (d0 + row + 1) * gdst_ld + 4 * q);
  st4(gsrc + p * gsrc_ld + 4 * q, make_float4((a.x + b.x) + (c.x + d.x), (a.y + b.y) + (c.y + d.y), (a.z + b.z) + (c.z + d.z),
                                              (a.w + b.w) + (c.w + d.w)));
}

}  // namespace demfi

using namespace demfi;

extern "C" {

int64_t demfi_l1_sum_workspace(int64_t n) { return n <= 0 ? -1 : ((n + L1_PER_BLOCK - 1) / L1_PER_BLOCK) * (int64_t)sizeof(double); }

int demfi_l1_sum(const float* pred, const float* target, int64_t n, float grad_scale, float* grad, void* workspace,
                 int64_t workspace_bytes, double* out, void* stream) {
  if (check_device()) return 3;
  DEMFI_REQUIRE(pred && target && out && workspace && n > 0, "l1_sum: bad arguments");
  DEMFI_REQUIRE(workspace_bytes >= demfi_l1_sum_workspace(n), "l1_sum: workspace too small");
  const long long nb = (n + L1_PER_BLOCK - 1) / L1_PER_BLOCK;
  DEMFI_REQUIRE(nb < (1ll << 31), "l1_sum: tensor too large");
  l1_partial_kernel<<<(unsigned)nb, 256, 0, (cudaStream_t)stream>>>(pred, target, n, grad_scale, grad, (double*)workspace);
  DEMFI_LAUNCH_CHECK("l1_partial");
  l1_reduce_kernel<<<1, 256, 0, (cudaStream_t)stream>>>((const double*)workspace, nb, out);
  DEMFI_LAUNCH_CHECK("l1_reduce");
  return 0;
}

int demfi_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1, float beta2,
                    float eps, float weight_decay, int32_t step, void* stream) {
  if (check_device()) return 3;
  DEMFI_REQUIRE(param && grad && exp_avg && exp_avg_sq && n > 0 && step >= 1, "adam_step: bad arguments");
  const double b1 = 1.0 - pow((double)beta1, (double)step), b2 = 1.0 - pow((double)beta2, (double)step);
  adam_step_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps,
                                                                                  weight_decay, (float)b1, (float)sqrt(b2));
  DEMFI_LAUNCH_CHECK("adam_step");
  return 0;
}

int demfi_fgac_blend_backward(const float* w, int32_t w_ld, const float* src, int32_t src_ld, const float* e, int32_t e_ld,
                              const float* gout, int32_t gout_ld, int64_t npix, int32_t C, float* dw, int32_t dw_ld, float* dsrc,
                              int32_t dsrc_ld, float* de, int32_t de_ld, void* stream) {
  if (check_device()) return 3;
  DEMFI_REQUIRE(w && src && e && gout && npix > 0 && C > 0 && C % 4 == 0 && src_ld % 4 == 0 && e_ld % 4 == 0 && gout_ld % 4 == 0 &&
                    (dsrc == nullptr || dsrc_ld % 4 == 0) && (de == nullptr || de_ld % 4 == 0), "fgac_blend_backward: bad arguments");
  fgac_blend_bwd_kernel<<<(unsigned)((npix * 16 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(w, w_ld, src, src_ld, e, e_ld, gout, gout_ld,
                                                                                               npix, C / 4, dw, dw_ld, dsrc, dsrc_ld, de, de_ld);
  DEMFI_LAUNCH_CHECK("fgac_blend_backward");
  return 0;
}

int demfi_upsample2x_backward(const float* gdst, int32_t gdst_ld, int32_t B, int32_t Hs, int32_t Ws, int32_t C, float* gsrc,
                              int32_t gsrc_ld, void* stream) {
  if (check_device()) return 3;
  DEMFI_REQUIRE(gdst && gsrc && B > 0 && Hs > 0 && Ws > 0 && C > 0 && C % 4 == 0 && gdst_ld % 4 == 0 && gsrc_ld % 4 == 0 &&
                    ((uintptr_t)gdst % 16) == 0 && ((uintptr_t)gsrc % 16) == 0, "upsample2x_backward: bad arguments");
  const long long n = (long long)B * Hs * Ws * (C / 4);
  upsample2x_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(gdst, gdst_ld, B, Hs, Ws, C / 4, gsrc, gsrc_ld);
  DEMFI_LAUNCH_CHECK("upsample2x_backward");
  return 0;
}

}  // extern "C"
