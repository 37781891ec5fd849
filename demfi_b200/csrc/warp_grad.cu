// Backward of the bilinear-sampling operators of the hot path (training row; first correct path):
//   * bwarp + Eq.(2) blend   (forward: ops.cu bwarp_blend_kernel; reference: bwarp DeMFInet.py:732-766, blend :66-71/:90-93/:146-149)
//   * FGAC sampling          (forward: ops.cu fgac_sample_kernel; reference: bilinear_sampler DeMFInet.py:499-514)
//   * complementary flow reversal (Gaussian forward splat), further down
// What the reference gets from autograd through F.grid_sample(bilinear, zeros, align_corners=True):
//   value gradient   d src[corner k] += g * w_k                    (scatter, fp32 red.global.add; zero the buffers first)
//   coordinate grad  dS/dpx = wy0 (v_ne - v_nw) + wy1 (v_se - v_sw), dS/dpy = wx0 (v_sw - v_nw) + wx1 (v_se - v_ne),
//                    out-of-image corners counting as 0 -- and d(px)/d(flow) = 1 through the normalise / un-normalise pair.
// bwarp's validity mask (warped ones < 0.999 -> 0) is piecewise constant: it multiplies the gradients, it has none itself.
// Eq.(2): out = ra A + rb B, ra = ka / (ka + kb), ka = (1-t) o, kb = t (1-o), o = sigmoid(logit):
//   d out / d logit = (A dra/do + B drb/do) o (1-o),  dra/do = ((1-t) den - ka (1-2t)) / den^2,  drb/do = (-t den - kb (1-2t)) / den^2.
// One group of 16 lanes per pixel, four channels per lane and pass (CW = 4), or one thread per pixel and scalar channels
// (CW = 1, the 3-channel pixel warp).  HBM / atomic bound: reads a, b, dout once, 8 scatter targets per pixel and channel group.
#include "common.cuh"
#include "warp.cuh"

namespace demfi {

struct CornerG {
  int idx[4];                 // clamped pixel indices y * W + x: nw, ne, sw, se
  float w[4];                 // forward weights, zero for out-of-image corners
  float in[4];                // 1 inside the image, else 0
  float wx0, wx1, wy0, wy1;
  float wsum;
};

__device__ __forceinline__ CornerG corner_grad(float px, float py, int H, int W) {
  CornerG g;
  const float fx0 = floorf(px), fy0 = floorf(py);
  g.wx0 = (fx0 + 1.0f) - px; g.wx1 = px - fx0;
  g.wy0 = (fy0 + 1.0f) - py; g.wy1 = py - fy0;
  const int x0 = (int)fminf(fmaxf(fx0, -2.0f), (float)W), y0 = (int)fminf(fmaxf(fy0, -2.0f), (float)H);
  const bool xi[2] = {x0 >= 0 && x0 < W, x0 + 1 >= 0 && x0 + 1 < W};
  const bool yi[2] = {y0 >= 0 && y0 < H, y0 + 1 >= 0 && y0 + 1 < H};
  const int xc[2] = {min(max(x0, 0), W - 1), min(max(x0 + 1, 0), W - 1)};
  const int yc[2] = {min(max(y0, 0), H - 1), min(max(y0 + 1, 0), H - 1)};
  const float wx[2] = {g.wx0, g.wx1}, wy[2] = {g.wy0, g.wy1};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int dy = k >> 1, dx = k & 1;
    const bool in = xi[dx] && yi[dy];
    g.idx[k] = yc[dy] * W + xc[dx];
    g.in[k] = in ? 1.0f : 0.0f;
    g.w[k] = in ? wx[dx] * wy[dy] : 0.0f;
  }
  g.wsum = ((g.w[0] + g.w[1]) + g.w[2]) + g.w[3];
  return g;
}

template <int CW>
__device__ __forceinline__ void load_cw(const float* p, float* v) {
  if constexpr (CW == 4) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else {
    v[0] = __ldg(p);
  }
}
template <int CW>
__device__ __forceinline__ void red_add_cw(float* p, const float* v) {
  if constexpr (CW == 4) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
  } else {
    atomicAdd(p, v[0]);
  }
}

// One sampled source of one pixel, one channel group: forward value S (out-of-image corners = 0), the scatter of
// `scale * g` into dsrc, and the two coordinate-gradient dot products with g.
template <int CW>
__device__ __forceinline__ void tap_backward(const float* __restrict__ src, float* __restrict__ dsrc, size_t ld, const CornerG& c,
                                             const float* g, float scale, float* S, float& gpx, float& gpy) {
  float v[4][CW];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    load_cw<CW>(src + (size_t)c.idx[k] * ld, v[k]);
#pragma unroll
    for (int j = 0; j < CW; ++j) v[k][j] *= c.in[k];
  }
#pragma unroll
  for (int j = 0; j < CW; ++j) {
    S[j] = ((v[0][j] * c.w[0] + v[1][j] * c.w[1]) + v[2][j] * c.w[2]) + v[3][j] * c.w[3];
    gpx += g[j] * (c.wy0 * (v[1][j] - v[0][j]) + c.wy1 * (v[3][j] - v[2][j]));
    gpy += g[j] * (c.wx0 * (v[2][j] - v[0][j]) + c.wx1 * (v[3][j] - v[1][j]));
  }
  if (dsrc != nullptr && scale != 0.0f) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (c.w[k] == 0.0f) continue;
      float d[CW];
#pragma unroll
      for (int j = 0; j < CW; ++j) d[j] = g[j] * (scale * c.w[k]);
      red_add_cw<CW>(dsrc + (size_t)c.idx[k] * ld, d);
    }
  }
}

template <int CW>
__global__ void __launch_bounds__(256)
bwarp_blend_bwd_kernel(const float* __restrict__ a, int a_ld, const float* __restrict__ b, int b_ld, const float* __restrict__ flow,
                       int flow_ld, const float* __restrict__ occ, int occ_ld, const float* __restrict__ tv,
                       const float* __restrict__ dout, int dout_ld, int B, int H, int W, int C, float* __restrict__ da, int da_ld,
                       float* __restrict__ db, int db_ld, float* __restrict__ dflow, int dflow_ld, float* __restrict__ docc,
                       int docc_ld) {
  constexpr int LPP = CW == 4 ? 16 : 1;
  const long long pix = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / LPP;
  const int lane = (int)(threadIdx.x % LPP);
  const long long npix = (long long)B * H * W;
  float gax = 0.f, gay = 0.f, gbx = 0.f, gby = 0.f, go = 0.f, o = 0.f;
  if (pix < npix) {
    const int x = (int)(pix % W), y = (int)((pix / W) % H), n = (int)(pix / ((long long)W * H));
    const float t = __ldg(tv + n);
    const float* fp = flow + pix * flow_ld;
    const float f0 = __ldg(fp), f1 = __ldg(fp + 1), f2 = __ldg(fp + 2), f3 = __ldg(fp + 3);
    o = sigmoid_f(__ldg(occ + pix * occ_ld));
    const CornerG ca = corner_grad(bwarp_coord(x, f0, W), bwarp_coord(y, f1, H), H, W);
    const CornerG cb = corner_grad(bwarp_coord(x, f2, W), bwarp_coord(y, f3, H), H, W);
    const float ma = ca.wsum < 0.999f ? 0.0f : 1.0f, mb = cb.wsum < 0.999f ? 0.0f : 1.0f;
    const float ka = (1.0f - t) * o, kb = t * (1.0f - o), den = ka + kb;
    const float ra = ka / den, rb = kb / den;
    const float dra = ((1.0f - t) * den - ka * (1.0f - 2.0f * t)) / (den * den);
    const float drb = (-t * den - kb * (1.0f - 2.0f * t)) / (den * den);
    const size_t img = (size_t)n * H * W;
    for (int ch = lane * CW; ch < C; ch += LPP * CW) {
      float g[CW], SA[CW], SB[CW];
      load_cw<CW>(dout + pix * dout_ld + ch, g);
      float px_ = 0.f, py_ = 0.f, qx_ = 0.f, qy_ = 0.f;
      tap_backward<CW>(a + img * a_ld + ch, da ? da + img * da_ld + ch : nullptr, (size_t)a_ld, ca, g, ra * ma, SA, px_, py_);
      tap_backward<CW>(b + img * b_ld + ch, db ? db + img * db_ld + ch : nullptr, (size_t)b_ld, cb, g, rb * mb, SB, qx_, qy_);
      gax += px_ * (ra * ma); gay += py_ * (ra * ma);
      gbx += qx_ * (rb * mb); gby += qy_ * (rb * mb);
#pragma unroll
      for (int j = 0; j < CW; ++j) go += g[j] * (ma * SA[j] * dra + mb * SB[j] * drb);
    }
  }
  if constexpr (LPP > 1) {
#pragma unroll
    for (int s = LPP / 2; s > 0; s >>= 1) {
      gax += __shfl_xor_sync(0xffffffffu, gax, s); gay += __shfl_xor_sync(0xffffffffu, gay, s);
      gbx += __shfl_xor_sync(0xffffffffu, gbx, s); gby += __shfl_xor_sync(0xffffffffu, gby, s);
      go += __shfl_xor_sync(0xffffffffu, go, s);
    }
  }
  if (pix < npix && lane == 0) {
    if (dflow != nullptr) {
      float* d = dflow + pix * dflow_ld;
      d[0] = gax; d[1] = gay; d[2] = gbx; d[3] = gby;
    }
    if (docc != nullptr) docc[pix * docc_ld] = go * o * (1.0f - o);
  }
}

template <int CW>
__global__ void __launch_bounds__(256)
fgac_sample_bwd_kernel(const float* __restrict__ refk, int refk_ld, const float* __restrict__ flow, int flow_ld,
                       const float* __restrict__ dout, int dout_ld, int B, int H, int W, int C, float* __restrict__ drefk,
                       int drefk_ld, float* __restrict__ dflow, int dflow_ld) {
  constexpr int LPP = CW == 4 ? 16 : 1;
  const long long pix = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / LPP;
  const int lane = (int)(threadIdx.x % LPP);
  const long long npix = (long long)B * H * W;
  float gx = 0.f, gy = 0.f;
  if (pix < npix) {
    const int n = (int)(pix / ((long long)W * H));
    const CornerG c = corner_grad(sampler_coord(__ldg(flow + pix * flow_ld), W), sampler_coord(__ldg(flow + pix * flow_ld + 1), H), H, W);
    const size_t img = (size_t)n * H * W;
    for (int ch = lane * CW; ch < C; ch += LPP * CW) {
      float g[CW], S[CW];
      load_cw<CW>(dout + pix * dout_ld + ch, g);
      tap_backward<CW>(refk + img * refk_ld + ch, drefk ? drefk + img * drefk_ld + ch : nullptr, (size_t)refk_ld, c, g, 1.0f, S, gx, gy);
    }
  }
  if constexpr (LPP > 1) {
#pragma unroll
    for (int s = LPP / 2; s > 0; s >>= 1) {
      gx += __shfl_xor_sync(0xffffffffu, gx, s);
      gy += __shfl_xor_sync(0xffffffffu, gy, s);
    }
  }
  if (pix < npix && lane == 0 && dflow != nullptr) {
    dflow[pix * dflow_ld] = gx;
    dflow[pix * dflow_ld + 1] = gy;
  }
}

// ------------------------------------------------------------------------------------------
// Backward of the complementary flow reversal (forward: ops.cu cfr_splat_kernel + cfr_finalize_kernel; reference:
// CFR_flow_t_align + fwarp + sample_one, DeMFInet.py:606-729).  The forward SCATTERS v * w_k and w_k to four targets, so the
// backward GATHERS -- no atomics:
//   1. per target pixel (cfr_finalize_bwd): from g = d/d(flow_t0, flow_t1), the forward accumulators and t, the gradients of
//      the accumulators {dA01.x, dA01.y, dn0, -, dA10.x, dA10.y, dn1, -} (same layout as acc).  With u = the combination
//      before the division and n = (1-t) n0 + t n1 > 0: du = g / n, dn = -(g . u) / n^2 (the mask of :615 is detached).
//   2. per source pixel (cfr_splat_bwd): for its four targets k (floor() taken exactly as in the forward, in-image only):
//      d/dv += w_k dA[k];  d/dw_k = v . dA[k] + dn[k];  w_k = exp(-((dy-cy_k)^2 + (dx-cx_k)^2)) => d w_k/d dx = -2 (dx-cx_k) w_k;
//      the displacement is s * v (s = t or 1-t), so d/dv += s * d/d(dx, dy).
__global__ void __launch_bounds__(256)
cfr_finalize_bwd_kernel(const float* __restrict__ acc, const float* __restrict__ tv, const float* __restrict__ gout, int gout_ld,
                        int B, int H, int W, float* __restrict__ gacc) {
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= (long long)B * H * W) return;
  const int n = (int)(pix / ((long long)W * H));
  const float t = __ldg(tv + n);
  const float4 a = __ldg((const float4*)(acc + pix * 8));
  const float4 b = __ldg((const float4*)(acc + pix * 8 + 4));
  const float* gp = gout + pix * gout_ld;
  float4 g = make_float4(__ldg(gp), __ldg(gp + 1), __ldg(gp + 2), __ldg(gp + 3));
  const float k00 = -(1.0f - t) * t, k01 = t * t, k10 = (1.0f - t) * (1.0f - t), k11 = t * (1.0f - t);
  const float ux = k00 * a.x + k01 * b.x, uy = k00 * a.y + k01 * b.y, uz = k10 * a.x - k11 * b.x, uw = k10 * a.y - k11 * b.y;
  const float norm = (1.0f - t) * a.z + t * b.z;
  float dn = 0.0f;
  if (norm > 0.0f) {
    const float rn = 1.0f / norm;
    dn = -(g.x * ux + g.y * uy + g.z * uz + g.w * uw) * rn * rn;
    g.x *= rn; g.y *= rn; g.z *= rn; g.w *= rn;
  }
  st4(gacc + pix * 8, make_float4(k00 * g.x + k10 * g.z, k00 * g.y + k10 * g.w, (1.0f - t) * dn, 0.0f));
  st4(gacc + pix * 8 + 4, make_float4(k01 * g.x - k11 * g.z, k01 * g.y - k11 * g.w, t * dn, 0.0f));
}

__device__ __forceinline__ void splat_one_bwd(const float* __restrict__ gacc, int n, int H, int W, int r, int c, float vx, float vy,
                                              float s, int slot, float& dvx, float& dvy) {
  const float dx = s * vx, dy = s * vy;
  const float fy = floorf(dy), fx = floorf(dx);
  const int iy = (int)fminf(fmaxf(fy, -(float)H - 2.f), (float)H + 2.f);
  const int ix = (int)fminf(fmaxf(fx, -(float)W - 2.f), (float)W + 2.f);
  float gvx = 0.f, gvy = 0.f, gdx = 0.f, gdy = 0.f;
#pragma unroll
  for (int oy = 0; oy < 2; ++oy)
#pragma unroll
    for (int ox = 0; ox < 2; ++ox) {
      const float cy = fy + (float)oy, cx = fx + (float)ox;
      const float wgt = expf(-((dy - cy) * (dy - cy) + (dx - cx) * (dx - cx)));
      const int tr = r + iy + oy, tc = c + ix + ox;
      if (tr >= 0 && tr < H && tc >= 0 && tc < W) {
        const float4 G = __ldg((const float4*)(gacc + (((size_t)n * H + tr) * W + tc) * 8 + slot));
        gvx += wgt * G.x;
        gvy += wgt * G.y;
        const float gw = (vx * G.x + vy * G.y + G.z) * wgt;
        gdx += gw * (-2.0f * (dx - cx));
        gdy += gw * (-2.0f * (dy - cy));
      }
    }
  dvx = gvx + s * gdx;
  dvy = gvy + s * gdy;
}

__global__ void __launch_bounds__(256)
cfr_splat_bwd_kernel(const float* __restrict__ fo, int fo_ld, const float* __restrict__ tv, const float* __restrict__ gacc, int B,
                     int H, int W, float* __restrict__ dfo, int dfo_ld) {
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= (long long)B * H * W) return;
  const int c = (int)(pix % W);
  const int r = (int)((pix / W) % H);
  const int n = (int)(pix / ((long long)W * H));
  const float t = __ldg(tv + n);
  const float4 f = __ldg((const float4*)(fo + pix * fo_ld));
  float4 d;
  splat_one_bwd(gacc, n, H, W, r, c, f.x, f.y, t, 0, d.x, d.y);
  splat_one_bwd(gacc, n, H, W, r, c, f.z, f.w, 1.0f - t, 4, d.z, d.w);
  float* o = dfo + pix * dfo_ld;
  o[0] = d.x; o[1] = d.y; o[2] = d.z; o[3] = d.w;
}

static inline bool vec_ok(int C, std::initializer_list<int> lds, std::initializer_list<const void*> ptrs) {
  if (C % 4 != 0) return false;
  for (int ld : lds)
    if (ld % 4 != 0) return false;
  for (const void* p : ptrs)
    if (p != nullptr && ((uintptr_t)p % 16) != 0) return false;
  return true;
}

}  // namespace demfi

using namespace demfi;

extern "C" {

int demfi_bwarp_blend_backward(const float* a, int32_t a_ld, const float* b, int32_t b_ld, const float* flow, int32_t flow_ld,
                               const float* occ, int32_t occ_ld, const float* t, const float* dout, int32_t dout_ld, int32_t B,
                               int32_t H, int32_t W, int32_t C, float* da, int32_t da_ld, float* db, int32_t db_ld, float* dflow,
                               int32_t dflow_ld, float* docc, int32_t docc_ld, void* stream) {
  if (check_device()) return 3;
  DEMFI_REQUIRE(a && b && flow && occ && t && dout, "bwarp_blend_backward: null input");
  DEMFI_REQUIRE(B > 0 && H > 1 && W > 1 && C > 0 && flow_ld >= 4 && (dflow == nullptr || dflow_ld >= 4), "bwarp_blend_backward: bad shape");
  const long long npix = (long long)B * H * W;
  if (vec_ok(C, {a_ld, b_ld, dout_ld, da ? da_ld : 4, db ? db_ld : 4}, {a, b, dout, da, db})) {
    bwarp_blend_bwd_kernel<4><<<(unsigned)((npix * 16 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        a, a_ld, b, b_ld, flow, flow_ld, occ, occ_ld, t, dout, dout_ld, B, H, W, C, da, da_ld, db, db_ld, dflow, dflow_ld, docc, docc_ld);
  } else {
    bwarp_blend_bwd_kernel<1><<<(unsigned)((npix + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        a, a_ld, b, b_ld, flow, flow_ld, occ, occ_ld, t, dout, dout_ld, B, H, W, C, da, da_ld, db, db_ld, dflow, dflow_ld, docc, docc_ld);
  }
  DEMFI_LAUNCH_CHECK("bwarp_blend_backward");
  return 0;
}

int demfi_fgac_sample_backward(const float* refk, int32_t refk_ld, const float* flow, int32_t flow_ld, const float* dout,
                               int32_t dout_ld, int32_t B, int32_t H, int32_t W, int32_t C, float* drefk, int32_t drefk_ld,
                               float* dflow, int32_t dflow_ld, void* stream) {
  if (check_device()) return 3;
  DEMFI_REQUIRE(refk && flow && dout, "fgac_sample_backward: null input");
  DEMFI_REQUIRE(B > 0 && H > 1 && W > 1 && C > 0 && flow_ld >= 2 && (dflow == nullptr || dflow_ld >= 2), "fgac_sample_backward: bad shape");
  const long long npix = (long long)B * H * W;
  if (vec_ok(C, {refk_ld, dout_ld, drefk ? drefk_ld : 4}, {refk, dout, drefk})) {
    fgac_sample_bwd_kernel<4><<<(unsigned)((npix * 16 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        refk, refk_ld, flow, flow_ld, dout, dout_ld, B, H, W, C, drefk, drefk_ld, dflow, dflow_ld);
  } else {
    fgac_sample_bwd_kernel<1><<<(unsigned)((npix + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        refk, refk_ld, flow, flow_ld, dout, dout_ld, B, H, W, C, drefk, drefk_ld, dflow, dflow_ld);
  }
  DEMFI_LAUNCH_CHECK("fgac_sample_backward");
  return 0;
}

int demfi_cfr_backward(const float* fo, int32_t fo_ld, const float* t, const float* acc, const float* gout, int32_t gout_ld,
                       int32_t B, int32_t H, int32_t W, float* gacc, float* dfo, int32_t dfo_ld, void* stream) {
  if (check_device()) return 3;
  DEMFI_REQUIRE(fo && t && acc && gout && gacc && dfo, "cfr_backward: null pointer");
  DEMFI_REQUIRE(B > 0 && H > 0 && W > 0 && fo_ld >= 4 && fo_ld % 4 == 0 && ((uintptr_t)fo % 16) == 0 && gout_ld >= 4 && dfo_ld >= 4 &&
                    ((uintptr_t)acc % 16) == 0 && ((uintptr_t)gacc % 16) == 0, "cfr_backward: bad shape or alignment");
  const long long npix = (long long)B * H * W;
  cfr_finalize_bwd_kernel<<<(unsigned)((npix + 255) / 256), 256, 0, (cudaStream_t)stream>>>(acc, t, gout, gout_ld, B, H, W, gacc);
  DEMFI_LAUNCH_CHECK("cfr_finalize_backward");
  cfr_splat_bwd_kernel<<<(unsigned)((npix + 255) / 256), 256, 0, (cudaStream_t)stream>>>(fo, fo_ld, t, gacc, B, H, W, dfo, dfo_ld);
  DEMFI_LAUNCH_CHECK("cfr_splat_backward");
  return 0;
}

}  // extern "C"
