// Bilinear-sampling coordinate arithmetic shared by the forward operators (ops.cu) and their backward (warp_grad.cu).
#pragma once
#include "common.cuh"

namespace demfi {

// ------------------------------------------------------------------------------------------
// Bilinear corner setup with the reference's coordinate arithmetic.
// bwarp (DeMFInet.py:750-757): g = 2*(x + f)/max(W-1,1) - 1, then grid_sample(align_corners=True)
// un-normalises ((g+1)/2)*(W-1).  The fp32 round trip is kept so that floor() and the 0.999
// validity threshold see the same numbers the reference sees.
struct Corners {
  int x0, y0;
  float w00, w01, w10, w11;  // [dy][dx], already zeroed for out-of-image corners
  float wsum;
};

__device__ __forceinline__ Corners make_corners(float px, float py, int H, int W) {
  Corners c;
  const float fx0 = floorf(px), fy0 = floorf(py);
  // ATen grid_sampler: w_nw = (x_se - x)*(y_se - y) etc.
  const float wx0 = (fx0 + 1.0f) - px, wx1 = px - fx0;
  const float wy0 = (fy0 + 1.0f) - py, wy1 = py - fy0;
  // clamp before the int conversion so absurd flows cannot overflow
  c.x0 = (int)fminf(fmaxf(fx0, -2.0f), (float)W);
  c.y0 = (int)fminf(fmaxf(fy0, -2.0f), (float)H);
  const bool x0in = (c.x0 >= 0 && c.x0 < W), x1in = (c.x0 + 1 >= 0 && c.x0 + 1 < W);
  const bool y0in = (c.y0 >= 0 && c.y0 < H), y1in = (c.y0 + 1 >= 0 && c.y0 + 1 < H);
  c.w00 = (x0in && y0in) ? wx0 * wy0 : 0.0f;
  c.w01 = (x1in && y0in) ? wx1 * wy0 : 0.0f;
  c.w10 = (x0in && y1in) ? wx0 * wy1 : 0.0f;
  c.w11 = (x1in && y1in) ? wx1 * wy1 : 0.0f;
  c.wsum = ((c.w00 + c.w01) + c.w10) + c.w11;
  return c;
}

__device__ __forceinline__ float bwarp_coord(int i, float f, int size) {
  const float g = 2.0f * ((float)i + f) / (float)max(size - 1, 1) - 1.0f;
  return ((g + 1.0f) / 2.0f) * (float)(size - 1);
}

// FGAC / bilinear_sampler coordinate (DeMFInet.py:499-508): the flow value IS the absolute position; it takes the same
// normalise / un-normalise round trip, g = 2 f/(W-1) - 1, then ((g+1)/2)(W-1).
__device__ __forceinline__ float sampler_coord(float f, int size) {
  const float g = 2.0f * f / (float)(size - 1) - 1.0f;
  return ((g + 1.0f) / 2.0f) * (float)(size - 1);
}

}  // namespace demfi
