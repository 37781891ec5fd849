// Shared host/device helpers for the demfi_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/demfi_b200.h"

namespace demfi {

// ---- host side --------------------------------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int check_device();  // 0 if the current device is sm_100; sets the error otherwise

#define DEMFI_REQUIRE(cond, ...)          \
  do {                                    \
    if (!(cond)) {                        \
      ::demfi::set_error(__VA_ARGS__);    \
      return 1;                           \
    }                                     \
  } while (0)

#define DEMFI_LAUNCH_CHECK(name)                                                  \
  do {                                                                            \
    cudaError_t e__ = cudaGetLastError();                                         \
    if (e__ != cudaSuccess) {                                                     \
      ::demfi::set_error("%s launch failed: %s", name, cudaGetErrorString(e__));  \
      return 2;                                                                   \
    }                                                                             \
    ::demfi::count_launch();                                                      \
  } while (0)

// ---- device side ------------------------------------------------------------------------
__device__ __forceinline__ float sigmoid_f(float v) { return 1.0f / (1.0f + expf(-v)); }

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

__device__ __forceinline__ float act_scalar(int act, float v) {
  switch (act) {
    case DEMFI_ACT_RELU: return fmaxf(v, 0.0f);
    case DEMFI_ACT_TANH: return tanhf(v);
    case DEMFI_ACT_SIGMOID: return sigmoid_f(v);
    default: return v;
  }
}

// Epilogue shared by the CUDA-core and the tcgen05 convolution kernels: four consecutive
// accumulator channels [co, co+4) of output pixel (n, y, x).  `acc` already holds the bias.
// Every segment whose channel range covers `co` receives the values (ranges may overlap: that
// is how one conv output is written to two consumers' buffers).
__device__ __forceinline__ void epilogue_store4(const demfi_conv_t& p, int n, int y, int x, int co, float4 acc) {
#pragma unroll 1
  for (int s = 0; s < p.nseg; ++s) {
    const demfi_seg_t& sg = p.seg[s];
    const int c = co - sg.ch0;
    if (c < 0 || c >= sg.nch) continue;
    size_t pix;
    int cc = c;
    if (sg.store == DEMFI_STORE_PIXEL_SHUFFLE2) {
      const int cq = sg.nch >> 2;
      const int q = c / cq;
      cc = c - q * cq;
      pix = ((size_t)n * (2 * p.H) + (2 * y + (q >> 1))) * (size_t)(2 * p.W) + (2 * x + (q & 1));
    } else {
      pix = ((size_t)n * p.H + y) * (size_t)p.W + x;
    }
    float4 v = acc;
    if (sg.act == DEMFI_ACT_SIGMOID_MUL) {
      const float4 r = ld4(sg.res + pix * sg.res_ld + cc);
      v.x = sigmoid_f(v.x) * r.x; v.y = sigmoid_f(v.y) * r.y; v.z = sigmoid_f(v.z) * r.z; v.w = sigmoid_f(v.w) * r.w;
    } else if (sg.act == DEMFI_ACT_GRU) {
      const float4 h = ld4(sg.res + pix * sg.res_ld + cc);
      const float4 z = ld4(sg.res2 + pix * sg.res2_ld + cc);
      v.x = (1.0f - z.x) * h.x + z.x * tanhf(v.x);
      v.y = (1.0f - z.y) * h.y + z.y * tanhf(v.y);
      v.z = (1.0f - z.z) * h.z + z.z * tanhf(v.z);
      v.w = (1.0f - z.w) * h.w + z.w * tanhf(v.w);
    } else {
      if (sg.res != nullptr) {
        const float4 r = ld4(sg.res + pix * sg.res_ld + cc);
        v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
      }
      v.x = act_scalar(sg.act, v.x); v.y = act_scalar(sg.act, v.y);
      v.z = act_scalar(sg.act, v.z); v.w = act_scalar(sg.act, v.w);
    }
    st4(sg.dst + pix * sg.dst_ld + cc, v);
  }
}

// conv launchers (one per kernel family)
int launch_conv_ffma(const demfi_conv_t& c, cudaStream_t st);
int launch_conv_tc(const demfi_conv_t& c, cudaStream_t st);
size_t tc_packed_floats(int KH, int KW, const int32_t* src_C, int nsrc, int cout_pad);
int tc_pack_weights(const float* w, int Co, int Ci, int KH, int KW, const int32_t* in_map, const int32_t* src_C,
                    int nsrc, const int32_t* out_map, int cout_pad, float* out);
int get_option(const char* name);

}  // namespace demfi
