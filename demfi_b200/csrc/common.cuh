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

// Epilogue shared by the CUDA-core and the tcgen05 convolution kernels.
//
// seg_store4: final value of four consecutive channels -> activation/residual/GRU math -> one
// 128-bit store.  Deliberately NOT inlined: the conv kernels call it from unrolled loops and an
// inlined copy per call site blew the instruction cache (ncu: stall_no_instruction dominated the
// epilogue warps, profiles/r1_conv_tc_epilogue.md).
static __device__ __noinline__ void seg_store4(float* d, const float* r, const float* r2, int act, float4 v) {
  if (act == DEMFI_ACT_SIGMOID_MUL) {
    const float4 h = ld4(r);
    v.x = sigmoid_f(v.x) * h.x; v.y = sigmoid_f(v.y) * h.y; v.z = sigmoid_f(v.z) * h.z; v.w = sigmoid_f(v.w) * h.w;
  } else if (act == DEMFI_ACT_GRU) {
    const float4 h = ld4(r);
    const float4 z = ld4(r2);
    v.x = (1.0f - z.x) * h.x + z.x * tanhf(v.x);
    v.y = (1.0f - z.y) * h.y + z.y * tanhf(v.y);
    v.z = (1.0f - z.z) * h.z + z.z * tanhf(v.z);
    v.w = (1.0f - z.w) * h.w + z.w * tanhf(v.w);
  } else {
    if (r != nullptr) {
      const float4 h = ld4(r);
      v.x += h.x; v.y += h.y; v.z += h.z; v.w += h.w;
    }
    if (act == DEMFI_ACT_RELU) {
      v.x = fmaxf(v.x, 0.0f); v.y = fmaxf(v.y, 0.0f); v.z = fmaxf(v.z, 0.0f); v.w = fmaxf(v.w, 0.0f);
    } else if (act == DEMFI_ACT_TANH) {
      v.x = tanhf(v.x); v.y = tanhf(v.y); v.z = tanhf(v.z); v.w = tanhf(v.w);
    } else if (act == DEMFI_ACT_SIGMOID) {
      v.x = sigmoid_f(v.x); v.y = sigmoid_f(v.y); v.z = sigmoid_f(v.z); v.w = sigmoid_f(v.w);
    }
  }
  st4(d, v);
}

// Per (output pixel, segment) addressing, computed once and reused for every channel group.
struct SegCursor {
  float* dst;        // address of the segment's channel 0 at this pixel (NHWC store)
  float* dst_base;   // segment base (pixel-shuffle store addresses four other pixels)
  const float* res;  // same for the optional operands
  const float* res2;
  int lo, hi;        // accumulator-channel range [lo, hi) the segment takes
  int act, store, cq;
  int dst_ld, n, y, x, H2, W2;  // pixel-shuffle store addresses the four pixels (2y+dy, 2x+dx)
};

__device__ __forceinline__ SegCursor seg_cursor(const demfi_conv_t& p, const demfi_seg_t& sg, int n, int y, int x) {
  SegCursor c;
  const size_t pix = ((size_t)n * p.H + y) * (size_t)p.W + x;
  c.lo = sg.ch0;
  c.hi = sg.ch0 + sg.nch;
  c.act = sg.act;
  c.store = sg.store;
  c.cq = sg.nch >> 2;
  c.dst_ld = sg.dst_ld;
  c.dst = sg.dst + pix * sg.dst_ld;
  c.dst_base = sg.dst;
  c.res = sg.res ? sg.res + pix * sg.res_ld : nullptr;
  c.res2 = sg.res2 ? sg.res2 + pix * sg.res2_ld : nullptr;
  c.n = n; c.y = y; c.x = x; c.H2 = 2 * p.H; c.W2 = 2 * p.W;
  return c;
}

// channels [co, co+4) (co = absolute accumulator channel) of the cursor's pixel; v already holds the bias
__device__ __forceinline__ void seg_emit4(const SegCursor& c, int co, float4 v) {
  if (co < c.lo || co >= c.hi) return;
  const int cc = co - c.lo;
  if (c.store == DEMFI_STORE_PIXEL_SHUFFLE2) {
    const int q = cc / c.cq;
    const size_t pq = ((size_t)c.n * c.H2 + (2 * c.y + (q >> 1))) * (size_t)c.W2 + (2 * c.x + (q & 1));
    seg_store4(c.dst_base + pq * c.dst_ld + (cc - q * c.cq), nullptr, nullptr, c.act, v);
  } else {
    seg_store4(c.dst + cc, c.res ? c.res + cc : nullptr, c.res2 ? c.res2 + cc : nullptr, c.act, v);
  }
}

// conv launchers (one per kernel family)
int launch_conv_ffma(const demfi_conv_t& c, cudaStream_t st);
int tc_debug_read(long long* host, int ctas);
long long* tc_debug_buffer(cudaStream_t st);  // zeroed role-timer buffer (tc_diag & 128)
int launch_conv_h3(const demfi_conv_t& c, cudaStream_t st);
bool h3_supports(const demfi_conv_t& c);
int launch_conv_s3(const demfi_conv_t& c, cudaStream_t st);  // same packed weights as conv_h3
bool s3_supports(const demfi_conv_t& c);
bool s3_s16_ok(const demfi_conv_t& c);
int s3_describe(const demfi_conv_t& c, int32_t* info);  // host-only plan summary (demfi_conv_describe)  // the S16 requests of this convolution can be honoured (TMA epilogue plan exists)
size_t h3_packed_floats(int KH, int KW, const int32_t* src_C, int nsrc, int cout_pad);
int h3_pack_weights_device(const float* w, int Co, int Ci, int KH, int KW, int src_c, int cout_pad, float* out, cudaStream_t st);
int h3_pack_weights(const float* w, int Co, int Ci, int KH, int KW, const int32_t* in_map, const int32_t* src_C,
                    int nsrc, const int32_t* out_map, int cout_pad, float* out, int nb_max = 0);
int s3_nb_max(int kind, int cout_pad);
int s3_pack_weights_pair(const float* w, int Co, int Ci, int KH, int KW, const int32_t* in_map, const int32_t* src_C, int nsrc,
                         const int32_t* out_map, int cout_pad, float* out);
int get_option(const char* name);

}  // namespace demfi
