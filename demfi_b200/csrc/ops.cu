// HBM-bound operators of the DeMFI-Net hot path (NHWC fp32, 128-bit accesses):
// input unpack, complementary-flow-reversal splat, bwarp + Eq.(2) blend, FGAC sampling and
// Eq.(4) blend, channel-slice copies and NCHW import/export.
#include "common.cuh"
#include "warp.cuh"

namespace demfi {

// ------------------------------------------------------------------------------------------
// bwarp + Eq.(2).  LPP lanes cooperate on one pixel, each lane owning 4 channels per pass.
// LPP = 16 (feature maps): a CTA walks an 8-row x 16-column pixel tile row by row, so the bilinear corners shared
// with the previous row (and with x-neighbours) are L1 hits instead of L2 round trips.
// LPP = 1 (3-channel pixel warp): one thread per pixel, scalar path.
// LPP = 16: ONE thread per pixel of the 16 x 8 tile first loads the flows / occlusion logit (coalesced) and does all the
// per-pixel arithmetic -- the two coordinate round trips, floor, corner weights, validity masks, sigmoid, blend factors --
// and leaves it in shared memory; the row loop is then: read 17 words (broadcast), eight independent 16-byte gathers per
// lane, blend, store.  ncu on the first version (everything inside the row loop, recomputed by each of the 16 lanes of a
// pixel, 80 registers, 36 % occupancy): DRAM traffic already ideal (511 MB read vs 501 MB algorithmic) but 2.0 TB/s --
// latency- and issue-bound (~300 instructions per lane per row).
struct PixelPlan {
  int ia[4], ib[4];   // clamped corner pixel indices (y * W + x) of the two warps: nw, ne, sw, se
  float aw[4], bw[4]; // corner weights, zero for out-of-image corners
  float ma, mb, ka, kb, rden;
};

__device__ __forceinline__ void plan_corners(const Corners& c, int H, int W, int* idx, float* w) {
  const int x0 = min(max(c.x0, 0), W - 1), x1 = min(max(c.x0 + 1, 0), W - 1);
  const int y0 = min(max(c.y0, 0), H - 1), y1 = min(max(c.y0 + 1, 0), H - 1);
  idx[0] = y0 * W + x0; idx[1] = y0 * W + x1; idx[2] = y1 * W + x0; idx[3] = y1 * W + x1;
  w[0] = c.w00; w[1] = c.w01; w[2] = c.w10; w[3] = c.w11;
}

// four corner loads (all in flight together), then the ATen accumulation order nw, ne, sw, se
__device__ __forceinline__ float4 gather4_planned(const float* base, unsigned ld, int i0, int i1, int i2, int i3, float w0, float w1,
                                                  float w2, float w3) {
  const float4 v0 = __ldg((const float4*)(base + (size_t)(unsigned)i0 * ld));
  const float4 v1 = __ldg((const float4*)(base + (size_t)(unsigned)i1 * ld));
  const float4 v2 = __ldg((const float4*)(base + (size_t)(unsigned)i2 * ld));
  const float4 v3 = __ldg((const float4*)(base + (size_t)(unsigned)i3 * ld));
  float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
  r.x += v0.x * w0; r.y += v0.y * w0; r.z += v0.z * w0; r.w += v0.w * w0;
  r.x += v1.x * w1; r.y += v1.y * w1; r.z += v1.z * w1; r.w += v1.w * w1;
  r.x += v2.x * w2; r.y += v2.y * w2; r.z += v2.z * w2; r.w += v2.w * w2;
  r.x += v3.x * w3; r.y += v3.y * w3; r.z += v3.z * w3; r.w += v3.w * w3;
  return r;
}

template <int LPP>
__global__ void __launch_bounds__(256, 4)
bwarp_blend_kernel(const float* __restrict__ a, int a_ld, const float* __restrict__ b, int b_ld,
                   const float* __restrict__ flow, int flow_ld, const float* __restrict__ occ, int occ_ld,
                   const float* __restrict__ tv, int B, int H, int W, int C, float* __restrict__ out, int out_ld,
                   float* __restrict__ occ_out, int occ_out_ld, int tiles_x, int tiles_y) {
  if constexpr (LPP == 16) {
    __shared__ PixelPlan plan[128];
    int t_ = blockIdx.x;
    const int tx = t_ % tiles_x;
    t_ /= tiles_x;
    const int ty = t_ % tiles_y;
    const int n = t_ / tiles_y;
    const int y0 = ty * 8;
    const float t = __ldg(tv + n);
    if (threadIdx.x < 128) {  // pixel (row threadIdx.x >> 4, column threadIdx.x & 15) of the tile
      const int py = y0 + ((int)threadIdx.x >> 4), px = tx * 16 + ((int)threadIdx.x & 15);
      if (py < H && px < W) {
        const long long pix = ((long long)n * H + py) * W + px;
        const float4 f = __ldg((const float4*)(flow + pix * flow_ld));
        const float o0 = sigmoid_f(__ldg(occ + pix * occ_ld));
        if (occ_out != nullptr) occ_out[pix * occ_out_ld] = o0;
        const float o1 = 1.0f - o0;
        const Corners ca = make_corners(bwarp_coord(px, f.x, W), bwarp_coord(py, f.y, H), H, W);
        const Corners cb = make_corners(bwarp_coord(px, f.z, W), bwarp_coord(py, f.w, H), H, W);
        PixelPlan& P = plan[threadIdx.x];
        plan_corners(ca, H, W, P.ia, P.aw);
        plan_corners(cb, H, W, P.ib, P.bw);
        // bwarp's validity mask: warped ones < 0.999 -> 0 (DeMFInet.py:758-766)
        P.ma = ca.wsum < 0.999f ? 0.0f : 1.0f;
        P.mb = cb.wsum < 0.999f ? 0.0f : 1.0f;
        P.ka = (1.0f - t) * o0;
        P.kb = t * o1;
        P.rden = 1.0f / (P.ka + P.kb);  // one IEEE division per pixel; the blend multiplies by it (<= 1 ulp from x / den)
      }
    }
    __syncthreads();
    const int col = (int)threadIdx.x >> 4, lane = (int)threadIdx.x & 15;
    const int x = tx * 16 + col;
    if (x >= W) return;
    const size_t img = (size_t)n * H * W;
    const float* pa = a + img * a_ld + lane * 4;
    const float* pb = b + img * b_ld + lane * 4;
    float* po = out + (img + (size_t)y0 * W + x) * out_ld + lane * 4;
#pragma unroll 1
    for (int r = 0; r < 8; ++r) {
      if (y0 + r >= H) break;
      const PixelPlan& P = plan[r * 16 + col];
      const float ma = P.ma, mb = P.mb, ka = P.ka, kb = P.kb, rden = P.rden;
      for (int ch = 0; ch < C; ch += 64) {
        if (ch + lane * 4 < C) {
          const float4 va = gather4_planned(pa + ch, (unsigned)a_ld, P.ia[0], P.ia[1], P.ia[2], P.ia[3], P.aw[0], P.aw[1], P.aw[2], P.aw[3]);
          const float4 vb = gather4_planned(pb + ch, (unsigned)b_ld, P.ib[0], P.ib[1], P.ib[2], P.ib[3], P.bw[0], P.bw[1], P.bw[2], P.bw[3]);
          float4 r4;
          r4.x = (ka * (va.x * ma) + kb * (vb.x * mb)) * rden;
          r4.y = (ka * (va.y * ma) + kb * (vb.y * mb)) * rden;
          r4.z = (ka * (va.z * ma) + kb * (vb.z * mb)) * rden;
          r4.w = (ka * (va.w * ma) + kb * (vb.w * mb)) * rden;
          st4(po + ch, r4);
        }
      }
      po += (size_t)W * out_ld;
    }
  } else {
    // LPP = 1 (3-channel pixel warp of the boosting loop): one thread per pixel, scalar channels
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)B * H * W) return;
    const int x = (int)(gid % W);
    const int y = (int)((gid / W) % H);
    const int n = (int)(gid / ((long long)W * H));
    const float t = __ldg(tv + n);
    const long long pix = gid;
    const float4 f = __ldg((const float4*)(flow + pix * flow_ld));
    const float o0 = sigmoid_f(__ldg(occ + pix * occ_ld));
    if (occ_out != nullptr) occ_out[pix * occ_out_ld] = o0;
    const float o1 = 1.0f - o0;
    const Corners ca = make_corners(bwarp_coord(x, f.x, W), bwarp_coord(y, f.y, H), H, W);
    const Corners cb = make_corners(bwarp_coord(x, f.z, W), bwarp_coord(y, f.w, H), H, W);
    const float ma = ca.wsum < 0.999f ? 0.0f : 1.0f;
    const float mb = cb.wsum < 0.999f ? 0.0f : 1.0f;
    const float ka = (1.0f - t) * o0, kb = t * o1;
    const float rden = 1.0f / (ka + kb);
    int ia[4], ib[4];
    float wa[4], wb[4];
    plan_corners(ca, H, W, ia, wa);
    plan_corners(cb, H, W, ib, wb);
    const size_t img = (size_t)n * H * W;
    const float* pa = a + img * a_ld;
    const float* pb = b + img * b_ld;
    if (C % 4 == 0) {
      for (int ch = 0; ch < C; ch += 4) {
        const float4 va = gather4_planned(pa + ch, (unsigned)a_ld, ia[0], ia[1], ia[2], ia[3], wa[0], wa[1], wa[2], wa[3]);
        const float4 vb = gather4_planned(pb + ch, (unsigned)b_ld, ib[0], ib[1], ib[2], ib[3], wb[0], wb[1], wb[2], wb[3]);
        float4 r4;
        r4.x = (ka * (va.x * ma) + kb * (vb.x * mb)) * rden;
        r4.y = (ka * (va.y * ma) + kb * (vb.y * mb)) * rden;
        r4.z = (ka * (va.z * ma) + kb * (vb.z * mb)) * rden;
        r4.w = (ka * (va.w * ma) + kb * (vb.w * mb)) * rden;
        st4(out + pix * out_ld + ch, r4);
      }
    } else {
      // scalar channels (C = 3): all 8 x C corner loads are independent; addresses computed once per corner
      const float* qa[4];
      const float* qb[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        qa[k] = pa + (size_t)(unsigned)ia[k] * (unsigned)a_ld;
        qb[k] = pb + (size_t)(unsigned)ib[k] * (unsigned)b_ld;
      }
      for (int ch = 0; ch < C; ++ch) {
        float va = 0.f, vb = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          va += __ldg(qa[k] + ch) * wa[k];
          vb += __ldg(qb[k] + ch) * wb[k];
        }
        out[pix * out_ld + ch] = (ka * (va * ma) + kb * (vb * mb)) * rden;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// FGAC sampling: absolute coordinates = flow values (DeMFInet.py:413-419, 499-508).
// A group of LPP lanes handles PPT pixels: lane k of the group does the coordinate arithmetic of pixel k (one flow load
// each, in parallel), the corner indices / weights are broadcast by shuffle, then all PPT x 4 corner gathers of a lane are
// issued before the first use (256 bytes of loads in flight per thread).
template <int LPP>
__global__ void __launch_bounds__(256)
fgac_sample_kernel(const float* __restrict__ refk, int refk_ld, const float* __restrict__ flow, int flow_ld, int B,
                   int H, int W, int C, float* __restrict__ out, int out_ld) {
  constexpr int PPT = 4;
  const int lane = (int)(threadIdx.x % LPP);
  const long long ppb = blockDim.x / LPP;  // pixels one block covers per step
  const long long p0 = (long long)blockIdx.x * ppb * PPT + threadIdx.x / LPP;
  const long long npix = (long long)B * H * W;
  const long long plane = (long long)W * H;
  int idx[4] = {0, 0, 0, 0};
  float wgt[4] = {0.f, 0.f, 0.f, 0.f};
  if (lane < PPT) {
    const long long pix = p0 + lane * ppb;
    if (pix < npix) {
      const float2 f = __ldg((const float2*)(flow + pix * flow_ld));
      // bilinear_sampler: g = 2*f/(W-1) - 1; grid_sample: ((g+1)/2)*(W-1)
      const float gx = 2.0f * f.x / (float)(W - 1) - 1.0f;
      const float gy = 2.0f * f.y / (float)(H - 1) - 1.0f;
      const Corners c = make_corners(((gx + 1.0f) / 2.0f) * (float)(W - 1), ((gy + 1.0f) / 2.0f) * (float)(H - 1), H, W);
      plan_corners(c, H, W, idx, wgt);
    }
  }
  for (int ch = lane * 4; ch < C; ch += LPP * 4) {
    float4 r[PPT];
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
      // every lane of the warp takes part in the shuffles (no early exit above)
      const int i0 = __shfl_sync(0xffffffffu, idx[0], k, LPP), i1 = __shfl_sync(0xffffffffu, idx[1], k, LPP);
      const int i2 = __shfl_sync(0xffffffffu, idx[2], k, LPP), i3 = __shfl_sync(0xffffffffu, idx[3], k, LPP);
      const float w0 = __shfl_sync(0xffffffffu, wgt[0], k, LPP), w1 = __shfl_sync(0xffffffffu, wgt[1], k, LPP);
      const float w2 = __shfl_sync(0xffffffffu, wgt[2], k, LPP), w3 = __shfl_sync(0xffffffffu, wgt[3], k, LPP);
      const long long pix = p0 + k * ppb;
      const long long n = pix < npix ? pix / plane : 0;
      r[k] = gather4_planned(refk + (size_t)n * plane * refk_ld + ch, (unsigned)refk_ld, i0, i1, i2, i3, w0, w1, w2, w3);
    }
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
      const long long pix = p0 + k * ppb;
      if (pix < npix) st4(out + pix * out_ld + ch, r[k]);
    }
  }
}

// Pure streaming (3 reads + 1 write per element): four pixels per lane group, all loads issued before the first use
// (64 bytes of loads in flight per thread instead of 32)
template <int LPP>
__global__ void __launch_bounds__(256)
fgac_blend_kernel(const float* __restrict__ w, int w_ld, const float* __restrict__ src, int src_ld,
                  const float* __restrict__ e, int e_ld, long long npix, int C, float* __restrict__ out, int out_ld) {
  constexpr int PPT = 4;  // pixels per thread, strided by the number of pixels one block covers per step
  const int lane = (int)(threadIdx.x % LPP);
  const long long ppb = blockDim.x / LPP;
  const long long p0 = (long long)blockIdx.x * ppb * PPT + threadIdx.x / LPP;
  for (int ch = lane * 4; ch < C; ch += LPP * 4) {
    float ww[PPT];
    float4 s[PPT], v[PPT];
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
      const long long pix = p0 + k * ppb;
      if (pix < npix) {
        ww[k] = __ldg(w + pix * w_ld);
        s[k] = __ldg((const float4*)(src + pix * src_ld + ch));
        v[k] = __ldg((const float4*)(e + pix * e_ld + ch));
      }
    }
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
      const long long pix = p0 + k * ppb;
      if (pix < npix) {
        float4 r;
        r.x = ww[k] * s[k].x + (1.0f - ww[k]) * v[k].x; r.y = ww[k] * s[k].y + (1.0f - ww[k]) * v[k].y;
        r.z = ww[k] * s[k].z + (1.0f - ww[k]) * v[k].z; r.w = ww[k] * s[k].w + (1.0f - ww[k]) * v[k].w;
        st4(out + pix * out_ld + ch, r);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// CFR: Gaussian forward splat.  acc layout per pixel: {a01.x, a01.y, n0, -, a10.x, a10.y, n1, -}.
__device__ __forceinline__ void splat_one(float* acc, int n, int H, int W, int r, int c, float vx, float vy, float dx,
                                          float dy, int slot) {
  const float fy = floorf(dy), fx = floorf(dx);
  const int iy = (int)fminf(fmaxf(fy, -(float)H - 2.f), (float)H + 2.f);
  const int ix = (int)fminf(fmaxf(fx, -(float)W - 2.f), (float)W + 2.f);
#pragma unroll
  for (int oy = 0; oy < 2; ++oy)
#pragma unroll
    for (int ox = 0; ox < 2; ++ox) {
      const float cy = fy + (float)oy, cx = fx + (float)ox;
      // get_gaussian_weights, DeMFInet.py:674-680 (x there is the row displacement)
      const float wgt = expf(-((dy - cy) * (dy - cy) + (dx - cx) * (dx - cx)));
      const int tr = r + iy + oy, tc = c + ix + ox;
      if (tr >= 0 && tr < H && tc >= 0 && tc < W) {
        // one 128-bit vector reduction per target (sm_90+) instead of three scalar ones: the splat is bound by L2 atomic
        // throughput (ncu: 105 us for 22.6 M scalar reductions, DRAM 7 % busy)
        float* d = acc + (((size_t)n * H + tr) * W + tc) * 8 + slot;
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d), "f"(vx * wgt), "f"(vy * wgt), "f"(wgt), "f"(0.0f)
                     : "memory");
      }
    }
}

__global__ void __launch_bounds__(256)
cfr_splat_kernel(const float* __restrict__ fo, int fo_ld, const float* __restrict__ tv, int B, int H, int W,
                 float* __restrict__ acc) {
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= (long long)B * H * W) return;
  const int c = (int)(pix % W);
  const int r = (int)((pix / W) % H);
  const int n = (int)(pix / ((long long)W * H));
  const float t = __ldg(tv + n);
  const float4 f = __ldg((const float4*)(fo + pix * fo_ld));
  // fwarp(flow_01, t*flow_01); fwarp(flow_10, (1-t)*flow_10)  (DeMFInet.py:609-612)
  splat_one(acc, n, H, W, r, c, f.x, f.y, t * f.x, t * f.y, 0);
  splat_one(acc, n, H, W, r, c, f.z, f.w, (1.0f - t) * f.z, (1.0f - t) * f.w, 4);
}

__global__ void __launch_bounds__(256)
cfr_finalize_kernel(const float* __restrict__ acc, const float* __restrict__ tv, int B, int H, int W,
                    float* __restrict__ out, int out_ld) {
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= (long long)B * H * W) return;
  const int n = (int)(pix / ((long long)W * H));
  const float t = __ldg(tv + n);
  const float4 a = __ldg((const float4*)(acc + pix * 8));
  const float4 b = __ldg((const float4*)(acc + pix * 8 + 4));
  // DeMFInet.py:614-620
  const float k00 = -(1.0f - t) * t, k01 = t * t, k10 = (1.0f - t) * (1.0f - t), k11 = t * (1.0f - t);
  float4 r;
  r.x = k00 * a.x + k01 * b.x;
  r.y = k00 * a.y + k01 * b.y;
  r.z = k10 * a.x - k11 * b.x;
  r.w = k10 * a.y - k11 * b.y;
  const float norm = (1.0f - t) * a.z + t * b.z;
  if (norm > 0.0f) { r.x /= norm; r.y /= norm; r.z /= norm; r.w /= norm; }
  st4(out + pix * out_ld, r);
}

// ------------------------------------------------------------------------------------------
// Input unpack: x [B,3,4,H,W] -> s2d [B,H/2,W/2,48], frames12 slices, mean(B0,B1) NCHW.
__global__ void __launch_bounds__(256)
pack_input_kernel(const float* __restrict__ x, int B, int H, int W, float* __restrict__ s2d, float* __restrict__ f12a,
                  int f12a_ld, float* __restrict__ f12b, int f12b_ld, float* __restrict__ mean01) {
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= (long long)B * H * W) return;
  const int xx = (int)(pix % W);
  const int yy = (int)((pix / W) % H);
  const int n = (int)(pix / ((long long)W * H));
  const size_t plane = (size_t)H * W;
  const float* xb = x + (size_t)n * 12 * plane + (size_t)yy * W + xx;
  float v[12];  // v[t*3 + c] = x[n, c, t, yy, xx]
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int t = 0; t < 4; ++t) v[t * 3 + c] = __ldg(xb + (size_t)(c * 4 + t) * plane);
  if (f12a) {
#pragma unroll
    for (int k = 0; k < 12; ++k) f12a[pix * f12a_ld + k] = v[k];
  }
  if (f12b) {
#pragma unroll
    for (int k = 0; k < 12; ++k) f12b[pix * f12b_ld + k] = v[k];
  }
  if (mean01) {
#pragma unroll
    for (int c = 0; c < 3; ++c) mean01[((size_t)n * 3 + c) * plane + (size_t)yy * W + xx] = (v[c] + v[3 + c]) / 2.0f;
  }
  if (s2d) {
    // pixel_reshuffle: out channel = cc*4 + dy*2 + dx, cc = t*3 + c  (DeMFInet.py:311-316)
    const int h2 = H >> 1, w2 = W >> 1;
    float* d = s2d + (((size_t)n * h2 + (yy >> 1)) * w2 + (xx >> 1)) * 48 + (yy & 1) * 2 + (xx & 1);
#pragma unroll
    for (int k = 0; k < 12; ++k) d[k * 4] = v[k];
  }
}

__global__ void __launch_bounds__(256)
copy_channels_kernel(const float* __restrict__ src, int src_ld, float* __restrict__ dst, int dst_ld, int nch,
                     long long npix, int act) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix * nch) return;
  const long long pix = i / nch;
  const int c = (int)(i % nch);
  dst[pix * dst_ld + c] = act_scalar(act, __ldg(src + pix * src_ld + c));
}

struct GatherParams {
  demfi_part_t part[DEMFI_MAX_PARTS];
  int nparts;
};

// one thread per (pixel, part): the parts of a pixel are copied by neighbouring threads, so the loads of all parts are in flight
// together and the destination row is completed by one warp
__global__ void __launch_bounds__(256)
gather_channels_kernel(const __grid_constant__ GatherParams G, float* __restrict__ dst, int dst_ld, long long npix) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long pix = gid / G.nparts;
  const int k = (int)(gid - pix * G.nparts);
  if (pix >= npix) return;
  const demfi_part_t& P = G.part[k];
  const float* s = P.src + pix * P.src_ld;
  float* d = dst + pix * dst_ld + P.dst_c0;
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (j < P.nch) v[j] = __ldg(s + j);
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (j < P.nch) d[j] = v[j];
  for (int j = 8; j < P.nch; ++j) d[j] = __ldg(s + j);
}

__global__ void __launch_bounds__(256)
export_nchw_kernel(const float* __restrict__ src, int src_ld, int B, int H, int W, int C, int act,
                   float* __restrict__ dst) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long plane = (long long)H * W;
  if (i >= (long long)B * C * plane) return;
  const long long p = i % plane;
  const int c = (int)((i / plane) % C);
  const long long n = i / (plane * C);
  dst[i] = act_scalar(act, __ldg(src + (n * plane + p) * src_ld + c));
}

__global__ void __launch_bounds__(256)
import_nchw_kernel(const float* __restrict__ src, int B, int H, int W, int C, float* __restrict__ dst, int dst_ld) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long plane = (long long)H * W;
  if (i >= (long long)B * C * plane) return;
  const int c = (int)(i % C);
  const long long p = (i / C) % plane;
  const long long n = i / (plane * C);
  dst[(n * plane + p) * dst_ld + c] = __ldg(src + (n * C + c) * plane + p);
}

// nearest-neighbour x2 up-sampling (nn.UpsamplingNearest2d, DeMFInet.py:573,592,597,601), 128-bit per thread
__global__ void __launch_bounds__(256)
upsample2x_kernel(const float* __restrict__ src, int src_ld, int B, int Hs, int Ws, int C4, float* __restrict__ dst,
                  int dst_ld) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int W = 2 * Ws, H = 2 * Hs;
  if (i >= (long long)B * H * W * C4) return;
  const int c4 = (int)(i % C4);
  const long long pix = i / C4;
  const int x = (int)(pix % W);
  const int y = (int)((pix / W) % H);
  const long long n = pix / ((long long)W * H);
  const float4 v = __ldg((const float4*)(src + ((n * Hs + (y >> 1)) * Ws + (x >> 1)) * src_ld) + c4);
  *((float4*)(dst + pix * dst_ld) + c4) = v;
}

static inline unsigned blocks_for(long long n, int per = 256) { return (unsigned)((n + per - 1) / per); }

// Channel mean of |a| or |a - b| per pixel (the FGAC difference / visualisation maps, DeMFInet.py:456-491): 16 lanes per
// pixel, one float4 per lane and pass, shuffle tree.  out[p] = mean_c |a[p,c] - b[p,c]|.
__global__ void __launch_bounds__(256) channel_absmean_kernel(const float* __restrict__ a, int a_ld, const float* __restrict__ b,
                                                               int b_ld, long long npix, int C4, float inv_c, float* __restrict__ out) {
  const long long p = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
  const int l = threadIdx.x & 15;
  float s = 0.0f;
  if (p < npix) {
    for (int q = l; q < C4; q += 16) {
      float4 v = ld4(a + p * a_ld + 4 * q);
      if (b != nullptr) {
        const float4 w = ld4(b + p * b_ld + 4 * q);
        v.x -= w.x; v.y -= w.y; v.z -= w.z; v.w -= w.w;
      }
      s += (fabsf(v.x) + fabsf(v.y)) + (fabsf(v.z) + fabsf(v.w));
    }
  }
  for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (p < npix && l == 0) out[p] = s * inv_c;
}

}  // namespace demfi

using namespace demfi;

// ------------------------------------------------------------------------------------------
// Pixel-wise blending of the boosting loop (PWB, DeMFInet.py:146-149), fused with the small concatenations around it:
// one thread per pixel reads the refined flows + occlusion logit (one 32-byte sector), gathers the two 3-channel frames
// through bwarp -- stored as [S0 pad | S1 pad], 4 channels each, so a bilinear corner of either is one 16-byte load and the
// two frames of a pixel share a sector --, applies Eq.(2) and writes [St(3) | sigmoid(occ)] and [flow(4)] as ONE 32-byte
// sector of the decoder's input row.  The general bwarp_blend_kernel<1> on the 36-channel interleaved row moved 3.8x the
// algorithmic bytes (sector-granular reads of 6 and writes of 4 scattered channels, ncu r1) and a separate copy kernel
// placed the flows.  Same arithmetic as bwarp_blend_kernel (validity mask, one IEEE reciprocal, ATen corner order).
__global__ void __launch_bounds__(256) pwb_kernel(const float* __restrict__ img, int img_ld, const float* __restrict__ fo, int fo_ld,
                                                   const float* __restrict__ tv, int B, int H, int W, float* __restrict__ out,
                                                   int out_ld) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)B * H * W) return;
  const int x = (int)(gid % W);
  const int y = (int)((gid / W) % H);
  const int n = (int)(gid / ((long long)W * H));
  const float t = __ldg(tv + n);
  const float4 f = __ldg((const float4*)(fo + gid * fo_ld));
  const float o0 = sigmoid_f(__ldg(fo + gid * fo_ld + 4));
  const float o1 = 1.0f - o0;
  const Corners ca = make_corners(bwarp_coord(x, f.x, W), bwarp_coord(y, f.y, H), H, W);
  const Corners cb = make_corners(bwarp_coord(x, f.z, W), bwarp_coord(y, f.w, H), H, W);
  const float ma = ca.wsum < 0.999f ? 0.0f : 1.0f;
  const float mb = cb.wsum < 0.999f ? 0.0f : 1.0f;
  const float ka = (1.0f - t) * o0, kb = t * o1;
  const float rden = 1.0f / (ka + kb);
  int ia[4], ib[4];
  float wa[4], wb[4];
  plan_corners(ca, H, W, ia, wa);
  plan_corners(cb, H, W, ib, wb);
  const float* base = img + (size_t)n * H * W * img_ld;
  const float4 va = gather4_planned(base, (unsigned)img_ld, ia[0], ia[1], ia[2], ia[3], wa[0], wa[1], wa[2], wa[3]);
  const float4 vb = gather4_planned(base + 4, (unsigned)img_ld, ib[0], ib[1], ib[2], ib[3], wb[0], wb[1], wb[2], wb[3]);
  float4 r4;
  r4.x = (ka * (va.x * ma) + kb * (vb.x * mb)) * rden;
  r4.y = (ka * (va.y * ma) + kb * (vb.y * mb)) * rden;
  r4.z = (ka * (va.z * ma) + kb * (vb.z * mb)) * rden;
  r4.w = o0;
  st4(out + gid * out_ld, r4);
  st4(out + gid * out_ld + 4, f);
}


extern "C" {

int demfi_bwarp_blend(const float* a, int32_t a_ld, const float* b, int32_t b_ld, const float* flow, int32_t flow_ld,
                      const float* occ, int32_t occ_ld, const float* t, int32_t B, int32_t H, int32_t W, int32_t C,
                      float* out, int32_t out_ld, float* occ_out, int32_t occ_out_ld, void* stream) {
  if (check_device()) return 3;
  DEMFI_REQUIRE(B > 0 && H > 1 && W > 1 && C > 0, "bwarp_blend: bad shape B=%d H=%d W=%d C=%d", B, H, W, C);
  DEMFI_REQUIRE(flow_ld % 4 == 0 && ((uintptr_t)flow % 16) == 0, "bwarp_blend: flow slice must be 16-byte aligned");
  if (C % 4 == 0)
    DEMFI_REQUIRE(a_ld % 4 == 0 && b_ld % 4 == 0 && out_ld % 4 == 0 && ((uintptr_t)a % 16) == 0 &&
                      ((uintptr_t)b % 16) == 0 && ((uintptr_t)out % 16) == 0,
                  "bwarp_blend: vector path needs 16-byte aligned slices");
  const long long npix = (long long)B * H * W;
  cudaStream_t st = (cudaStream_t)stream;
  if (C >= 64 && C % 4 == 0) {
    const int tiles_x = (W + 15) / 16, tiles_y = (H + 7) / 8;
    bwarp_blend_kernel<16><<<(unsigned)((long long)tiles_x * tiles_y * B), 256, 0, st>>>(
        a, a_ld, b, b_ld, flow, flow_ld, occ, occ_ld, t, B, H, W, C, out, out_ld, occ_out, occ_out_ld, tiles_x, tiles_y);
  } else {
    bwarp_blend_kernel<1><<<blocks_for(npix), 256, 0, st>>>(a, a_ld, b, b_ld, flow, flow_ld, occ, occ_ld, t, B, H, W, C,
                                                            out, out_ld, occ_out, occ_out_ld, 0, 0);
  }
  DEMFI_LAUNCH_CHECK("bwarp_blend");
  return 0;
}

int demfi_pwb(const float* img, int32_t img_ld, const float* fo, int32_t fo_ld, const float* t, int32_t B, int32_t H, int32_t W,
              float* out, int32_t out_ld, void* stream) {
  if (check_device()) return 3;
  DEMFI_REQUIRE(img && fo && t && out, "pwb: null argument");
  DEMFI_REQUIRE(B > 0 && H > 1 && W > 1, "pwb: bad shape B=%d H=%d W=%d", B, H, W);
  DEMFI_REQUIRE(img_ld % 4 == 0 && img_ld >= 8 && fo_ld % 4 == 0 && fo_ld >= 8 && out_ld % 4 == 0 && out_ld >= 8 &&
                    ((uintptr_t)img % 16) == 0 && ((uintptr_t)fo % 16) == 0 && ((uintptr_t)out % 16) == 0,
                "pwb: img (8 channels: S0 pad S1 pad), fo (flow 4, occlusion logit, pad 3) and out (8 channels) must be 16-byte aligned slices");
  const long long npix = (long long)B * H * W;
  pwb_kernel<<<blocks_for(npix), 256, 0, (cudaStream_t)stream>>>(img, img_ld, fo, fo_ld, t, B, H, W, out, out_ld);
  DEMFI_LAUNCH_CHECK("pwb");
  return 0;
}

int demfi_fgac_sample(const float* refk, int32_t refk_ld, const float* flow, int32_t flow_ld, int32_t B, int32_t H,
                      int32_t W, int32_t C, float* out, int32_t out_ld, void* stream) {
  if (check_device()) return 3;
  DEMFI_REQUIRE(B > 0 && H > 1 && W > 1 && C > 0 && C % 4 == 0, "fgac_sample: bad shape");
  DEMFI_REQUIRE(refk_ld % 4 == 0 && out_ld % 4 == 0 && flow_ld % 2 == 0 && ((uintptr_t)flow % 8) == 0 &&
                    ((uintptr_t)refk % 16) == 0 && ((uintptr_t)out % 16) == 0, "fgac_sample: misaligned slices");
  const long long npix = (long long)B * H * W;
  fgac_sample_kernel<16><<<blocks_for(npix * 16, 256 * 4), 256, 0, (cudaStream_t)stream>>>(refk, refk_ld, flow, flow_ld, B, H, W,
                                                                                   C, out, out_ld);
  DEMFI_LAUNCH_CHECK("fgac_sample");
  return 0;
}

int demfi_fgac_blend(const float* w, int32_t w_ld, const float* src, int32_t src_ld, const float* e, int32_t e_ld,
                     int64_t npix, int32_t C, float* out, int32_t out_ld, void* stream) {
  if (check_device()) return 3;
  DEMFI_REQUIRE(npix > 0 && C > 0 && C % 4 == 0 && src_ld % 4 == 0 && e_ld % 4 == 0 && out_ld % 4 == 0,
                "fgac_blend: bad shape");
  fgac_blend_kernel<16><<<blocks_for(npix * 16, 256 * 4), 256, 0, (cudaStream_t)stream>>>(w, w_ld, src, src_ld, e, e_ld, npix, C,
                                                                                  out, out_ld);
  DEMFI_LAUNCH_CHECK("fgac_blend");
  return 0;
}

int demfi_cfr_splat(const float* fo, int32_t fo_ld, const float* t, int32_t B, int32_t H, int32_t W, float* acc,
                    void* stream) {
  if (check_device()) return 3;
  DEMFI_REQUIRE(B > 0 && H > 0 && W > 0 && fo_ld % 4 == 0 && ((uintptr_t)fo % 16) == 0, "cfr_splat: bad arguments");
  cfr_splat_kernel<<<blocks_for((long long)B * H * W), 256, 0, (cudaStream_t)stream>>>(fo, fo_ld, t, B, H, W, acc);
  DEMFI_LAUNCH_CHECK("cfr_splat");
  return 0;
}

int demfi_cfr_finalize(const float* acc, const float* t, int32_t B, int32_t H, int32_t W, float* out, int32_t out_ld,
                       void* stream) {
  if (check_device()) return 3;
  DEMFI_REQUIRE(B > 0 && H > 0 && W > 0 && out_ld % 4 == 0 && ((uintptr_t)out % 16) == 0, "cfr_finalize: bad arguments");
  cfr_finalize_kernel<<<blocks_for((long long)B * H * W), 256, 0, (cudaStream_t)stream>>>(acc, t, B, H, W, out, out_ld);
  DEMFI_LAUNCH_CHECK("cfr_finalize");
  return 0;
}

int demfi_pack_input(const float* x, int32_t B, int32_t H, int32_t W, float* s2d, float* f12a, int32_t f12a_ld,
                     float* f12b, int32_t f12b_ld, float* mean01, void* stream) {
  if (check_device()) return 3;
  DEMFI_REQUIRE(B > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0, "pack_input: H and W must be even");
  pack_input_kernel<<<blocks_for((long long)B * H * W), 256, 0, (cudaStream_t)stream>>>(x, B, H, W, s2d, f12a, f12a_ld,
                                                                                         f12b, f12b_ld, mean01);
  DEMFI_LAUNCH_CHECK("pack_input");
  return 0;
}

int demfi_copy_channels(const float* src, int32_t src_ld, float* dst, int32_t dst_ld, int32_t nch, int64_t npix,
                        int32_t act, void* stream) {
  if (check_device()) return 3;
  DEMFI_REQUIRE(nch > 0 && npix > 0, "copy_channels: bad shape");
  copy_channels_kernel<<<blocks_for(npix * nch), 256, 0, (cudaStream_t)stream>>>(src, src_ld, dst, dst_ld, nch, npix, act);
  DEMFI_LAUNCH_CHECK("copy_channels");
  return 0;
}

int demfi_gather_channels(const demfi_part_t* parts, int32_t nparts, float* dst, int32_t dst_ld, int64_t npix, void* stream) {
  if (check_device()) return 3;
  DEMFI_REQUIRE(parts && dst && nparts >= 1 && nparts <= DEMFI_MAX_PARTS && npix > 0, "gather_channels: bad arguments");
  GatherParams G;
  G.nparts = nparts;
  for (int k = 0; k < nparts; ++k) {
    DEMFI_REQUIRE(parts[k].src && parts[k].nch > 0 && parts[k].dst_c0 >= 0 && parts[k].dst_c0 + parts[k].nch <= dst_ld,
                  "gather_channels: part %d does not fit the destination row", k);
    G.part[k] = parts[k];
  }
  gather_channels_kernel<<<blocks_for(npix * nparts), 256, 0, (cudaStream_t)stream>>>(G, dst, dst_ld, npix);
  DEMFI_LAUNCH_CHECK("gather_channels");
  return 0;
}

int demfi_channel_absmean(const float* a, int32_t a_ld, const float* b, int32_t b_ld, int64_t npix, int32_t C, float* out,
                          void* stream) {
  if (check_device()) return 3;
  DEMFI_REQUIRE(a && out && npix > 0 && C > 0 && C % 4 == 0 && a_ld % 4 == 0 && ((uintptr_t)a % 16) == 0 &&
                    (b == nullptr || (b_ld % 4 == 0 && ((uintptr_t)b % 16) == 0)), "channel_absmean: bad arguments");
  channel_absmean_kernel<<<blocks_for(npix * 16), 256, 0, (cudaStream_t)stream>>>(a, a_ld, b, b_ld, npix, C / 4, 1.0f / (float)C, out);
  DEMFI_LAUNCH_CHECK("channel_absmean");
  return 0;
}

int demfi_upsample2x(const float* src, int32_t src_ld, int32_t B, int32_t Hs, int32_t Ws, int32_t C, float* dst,
                     int32_t dst_ld, void* stream) {
  if (check_device()) return 3;
  DEMFI_REQUIRE(B > 0 && Hs > 0 && Ws > 0 && C > 0 && C % 4 == 0 && src_ld % 4 == 0 && dst_ld % 4 == 0 &&
                    ((uintptr_t)src % 16) == 0 && ((uintptr_t)dst % 16) == 0, "upsample2x: bad arguments");
  upsample2x_kernel<<<blocks_for((long long)B * Hs * Ws * C), 256, 0, (cudaStream_t)stream>>>(src, src_ld, B, Hs, Ws, C / 4, dst, dst_ld);
  DEMFI_LAUNCH_CHECK("upsample2x");
  return 0;
}

int demfi_export_nchw(const float* src, int32_t src_ld, int32_t B, int32_t H, int32_t W, int32_t C, int32_t act,
                      float* dst, void* stream) {
  if (check_device()) return 3;
  DEMFI_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0, "export_nchw: bad shape");
  export_nchw_kernel<<<blocks_for((long long)B * C * H * W), 256, 0, (cudaStream_t)stream>>>(src, src_ld, B, H, W, C, act, dst);
  DEMFI_LAUNCH_CHECK("export_nchw");
  return 0;
}

int demfi_import_nchw(const float* src, int32_t B, int32_t H, int32_t W, int32_t C, float* dst, int32_t dst_ld,
                      void* stream) {
  if (check_device()) return 3;
  DEMFI_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0, "import_nchw: bad shape");
  import_nchw_kernel<<<blocks_for((long long)B * C * H * W), 256, 0, (cudaStream_t)stream>>>(src, B, H, W, C, dst, dst_ld);
  DEMFI_LAUNCH_CHECK("import_nchw");
  return 0;
}

}  // extern "C"
