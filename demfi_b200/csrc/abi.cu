// Library-level entry points of the C ABI: errors, device check, weight packing, conv dispatch.
#include <atomic>
#include <cstdarg>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace demfi {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

static std::atomic<int> g_opt_mask_hi{1};
static std::atomic<int> g_opt_split{3};
static std::atomic<int> g_opt_flush{20};  // stages per accumulation segment (tools/flush_sweep.py: error unchanged up to 24)
static std::atomic<int> g_opt_atmem{1};   // A operand through tensor memory (1) or shared memory (0)
static std::atomic<int> g_opt_diag{0};
static std::atomic<int> g_opt_comp{270};
static std::atomic<int> g_opt_stages{0};  // diagnostics: cap on pipeline stages (0 = as many as fit)
static std::atomic<int> g_opt_grid{0};    // diagnostics: cap on persistent CTAs (0 = one per SM)
static std::atomic<int> g_opt_pdl{0};     // programmatic dependent launch between consecutive conv_s3 kernels (measured: 47.6 vs 46.6 ms per forward, off)
static std::atomic<int> g_opt_prefetch{0};  // conv_s3: every activation chunk is also prefetched into L2 this many tiles ahead of the CTA's sequence
static std::atomic<int> g_opt_nwide{1};     // conv_s3: Cout 97..128 as ONE N block (N' = 256 MMAs) instead of blocks of 64
static std::atomic<int> g_opt_wgrad{1};   // demfi_conv2d_wgrad: 1 = mma.sync 3xTF32 kernel, 0 = CUDA-core fp32 kernel
static std::atomic<int> g_opt_gen{3};     // DEMFI_CONV_TC16 kernel generation: 3 = conv_s3 where supported, 2 = conv_h3 only
int get_option(const char* name) {
  if (!strcmp(name, "tc_mask_hi")) return g_opt_mask_hi.load();
  if (!strcmp(name, "tc_split")) return g_opt_split.load();
  if (!strcmp(name, "tc_flush")) return g_opt_flush.load();
  if (!strcmp(name, "tc_stages")) return g_opt_stages.load();
  if (!strcmp(name, "tc_diag")) return g_opt_diag.load();
  if (!strcmp(name, "tc_comp_milli")) return g_opt_comp.load();
  if (!strcmp(name, "tc_a_tmem")) return g_opt_atmem.load();
  if (!strcmp(name, "tc_grid")) return g_opt_grid.load();
  if (!strcmp(name, "tc_gen")) return g_opt_gen.load();
  if (!strcmp(name, "tc_pdl")) return g_opt_pdl.load();
  if (!strcmp(name, "tc_prefetch")) return g_opt_prefetch.load();
  if (!strcmp(name, "tc_nwide")) return g_opt_nwide.load();
  if (!strcmp(name, "wgrad_kind")) return g_opt_wgrad.load();
  return -1;
}

// role timers of the last conv_s3 / conv_h3 launch made with tc_diag & 128: [ctas][16] cycles (tools/role_timers.py names the slots)
constexpr int TC_DBG_CTAS = 256;
static long long* g_tc_dbg = nullptr;
int tc_debug_read(long long* host, int ctas) {
  DEMFI_REQUIRE(g_tc_dbg != nullptr, "tc_debug_read: no launch was made with tc_diag & 128");
  DEMFI_REQUIRE(ctas > 0 && ctas <= TC_DBG_CTAS, "tc_debug_read: ctas out of range");
  DEMFI_REQUIRE(cudaMemcpy(host, g_tc_dbg, (size_t)ctas * 16 * sizeof(long long), cudaMemcpyDeviceToHost) == cudaSuccess, "tc_debug_read: copy failed");
  return 0;
}
long long* tc_debug_buffer(cudaStream_t st) {
  if (g_tc_dbg == nullptr && cudaMalloc(&g_tc_dbg, TC_DBG_CTAS * 16 * sizeof(long long)) != cudaSuccess) return nullptr;
  cudaMemsetAsync(g_tc_dbg, 0, TC_DBG_CTAS * 16 * sizeof(long long), st);
  return g_tc_dbg;
}

int check_device() {
  static thread_local int ok_dev = -1;
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    set_error("no CUDA device: demfi_b200 has no CPU path");
    return 3;
  }
  if (dev == ok_dev) return 0;
  cudaDeviceProp pr;
  if (cudaGetDeviceProperties(&pr, dev) != cudaSuccess) {
    set_error("cudaGetDeviceProperties failed");
    return 3;
  }
  if (pr.major != 10) {
    set_error("device %d is sm_%d%d; demfi_b200 is built for sm_100a (B200) only", dev, pr.major, pr.minor);
    return 3;
  }
  ok_dev = dev;
  return 0;
}

}  // namespace demfi

using namespace demfi;

extern "C" {

int demfi_version(void) { return DEMFI_ABI_VERSION; }
const char* demfi_last_error(void) { return g_err; }
uint64_t demfi_launch_count(void) { return g_launches.load(); }

int demfi_device_check(int device) {
  cudaDeviceProp pr;
  if (cudaGetDeviceProperties(&pr, device) != cudaSuccess) {
    set_error("device %d not available", device);
    return 3;
  }
  if (pr.major != 10) {
    set_error("device %d is sm_%d%d, need sm_100", device, pr.major, pr.minor);
    return 3;
  }
  return 0;
}

int demfi_set_option(const char* name, int32_t value) {
  DEMFI_REQUIRE(name != nullptr, "set_option: null name");
  if (!strcmp(name, "tc_mask_hi")) { g_opt_mask_hi.store(value ? 1 : 0); return 0; }
  if (!strcmp(name, "tc_split")) {
    DEMFI_REQUIRE(value == 1 || value == 3, "set_option: tc_split must be 1 or 3");
    g_opt_split.store(value);
    return 0;
  }
  if (!strcmp(name, "tc_flush")) { g_opt_flush.store(value < 0 ? 0 : value); return 0; }
  if (!strcmp(name, "tc_stages")) { g_opt_stages.store(value < 0 ? 0 : value); return 0; }
  if (!strcmp(name, "tc_diag")) { g_opt_diag.store(value); return 0; }
  if (!strcmp(name, "tc_comp_milli")) { g_opt_comp.store(value); return 0; }
  if (!strcmp(name, "tc_a_tmem")) { g_opt_atmem.store(value ? 1 : 0); return 0; }
  if (!strcmp(name, "tc_grid")) { g_opt_grid.store(value < 0 ? 0 : value); return 0; }
  if (!strcmp(name, "tc_pdl")) { g_opt_pdl.store(value ? 1 : 0); return 0; }
  if (!strcmp(name, "tc_prefetch")) { g_opt_prefetch.store(value < 0 ? 0 : value); return 0; }
  if (!strcmp(name, "tc_nwide")) { g_opt_nwide.store(value ? 1 : 0); return 0; }
  if (!strcmp(name, "wgrad_kind")) { g_opt_wgrad.store(value ? 1 : 0); return 0; }
  if (!strcmp(name, "tc_gen")) {
    DEMFI_REQUIRE(value == 2 || value == 3, "set_option: tc_gen must be 2 or 3");
    g_opt_gen.store(value);
    return 0;
  }
  set_error("set_option: unknown option '%s'", name);
  return 1;
}
int demfi_get_option(const char* name, int32_t* value) {
  DEMFI_REQUIRE(name != nullptr && value != nullptr, "get_option: null argument");
  const int v = get_option(name);
  DEMFI_REQUIRE(v >= 0, "get_option: unknown option '%s'", name);
  *value = v;
  return 0;
}

int demfi_pack_weights_device(int32_t kind, const float* w, int32_t Co, int32_t Ci, int32_t KH, int32_t KW, int32_t src_c,
                              int32_t cout_pad, float* out, void* stream) {
  DEMFI_REQUIRE(w && out, "pack_weights_device: null argument");
  DEMFI_REQUIRE(kind == DEMFI_CONV_TC16, "pack_weights_device: kind DEMFI_CONV_TC16 only (the training path's kernel)");
  DEMFI_REQUIRE(Co > 0 && Ci > 0 && KH > 0 && KW > 0, "pack_weights_device: bad shape");
  return h3_pack_weights_device(w, Co, Ci, KH, KW, src_c, cout_pad, out, static_cast<cudaStream_t>(stream));
}

size_t demfi_packed_weight_floats(int32_t kind, int32_t KH, int32_t KW, const int32_t* src_C, int32_t nsrc,
                                  int32_t cout_pad) {
  if (kind == DEMFI_CONV_TC) return 0;  // (retired, see demfi_pack_weights)
  if (kind == DEMFI_CONV_TC16 || kind == DEMFI_CONV_TC16W || kind == DEMFI_CONV_TC16P) return h3_packed_floats(KH, KW, src_C, nsrc, cout_pad);
  int k_total = 0;
  for (int s = 0; s < nsrc; ++s) k_total += src_C[s];
  return (size_t)KH * KW * k_total * cout_pad;
}

int demfi_pack_weights(int32_t kind, const float* w, int32_t Co, int32_t Ci, int32_t KH, int32_t KW,
                       const int32_t* in_map, const int32_t* src_C, int32_t nsrc, const int32_t* out_map,
                       int32_t cout_pad, float* out) {
  DEMFI_REQUIRE(w && in_map && out_map && out && src_C, "pack_weights: null argument");
  DEMFI_REQUIRE(nsrc >= 1 && nsrc <= DEMFI_MAX_SRC, "pack_weights: bad nsrc");
  int k_total = 0;
  for (int s = 0; s < nsrc; ++s) k_total += src_C[s];
  DEMFI_REQUIRE(k_total > 0 && k_total % 4 == 0 && cout_pad > 0 && cout_pad % 4 == 0, "pack_weights: bad padding");
  for (int k = 0; k < k_total; ++k) DEMFI_REQUIRE(in_map[k] >= -1 && in_map[k] < Ci, "pack_weights: in_map[%d] out of range", k);
  for (int n = 0; n < cout_pad; ++n) DEMFI_REQUIRE(out_map[n] >= -1 && out_map[n] < Co, "pack_weights: out_map[%d] out of range", n);
  DEMFI_REQUIRE(kind != DEMFI_CONV_TC, "DEMFI_CONV_TC (the first-generation 3xTF32 kernel) was retired in round 2: use DEMFI_CONV_TC16");
  if (kind == DEMFI_CONV_TC16) return h3_pack_weights(w, Co, Ci, KH, KW, in_map, src_C, nsrc, out_map, cout_pad, out);
  if (kind == DEMFI_CONV_TC16P) return s3_pack_weights_pair(w, Co, Ci, KH, KW, in_map, src_C, nsrc, out_map, cout_pad, out);
  if (kind == DEMFI_CONV_TC16W)
    return h3_pack_weights(w, Co, Ci, KH, KW, in_map, src_C, nsrc, out_map, cout_pad, out, s3_nb_max(kind, cout_pad));
  // FFMA layout: [tap][k][cout_pad]
  const int taps = KH * KW;
  for (int tap = 0; tap < taps; ++tap)
    for (int k = 0; k < k_total; ++k) {
      float* row = out + ((size_t)tap * k_total + k) * cout_pad;
      const int ci = in_map[k];
      for (int n = 0; n < cout_pad; ++n) {
        const int co = out_map[n];
        row[n] = (ci < 0 || co < 0) ? 0.0f : w[((size_t)co * Ci + ci) * taps + tap];
      }
    }
  return 0;
}

int demfi_tc_debug_read(int64_t* host, int32_t ctas) {
  DEMFI_REQUIRE(host != nullptr, "tc_debug_read: null buffer");
  return tc_debug_read(reinterpret_cast<long long*>(host), ctas);
}

int demfi_conv_describe(const demfi_conv_t* c, int32_t* info) {
  DEMFI_REQUIRE(c != nullptr && info != nullptr, "conv_describe: null argument");
  for (int i = 0; i < 16; ++i) info[i] = 0;
  if (c->kind == DEMFI_CONV_FFMA) return 0;
  DEMFI_REQUIRE(c->kind != DEMFI_CONV_TC, "conv_describe: DEMFI_CONV_TC was retired in round 2: use DEMFI_CONV_TC16");
  const bool s3_only = c->kind == DEMFI_CONV_TC16W || c->kind == DEMFI_CONV_TC16P;
  DEMFI_REQUIRE(c->kind == DEMFI_CONV_TC16 || s3_only, "conv_describe: unknown kind %d", c->kind);
  if (s3_only) DEMFI_REQUIRE(s3_supports(*c), "conv_describe: DEMFI_CONV_TC16W / TC16P need a convolution conv_s3 supports");
  if ((g_opt_gen.load() == 3 || s3_only) && s3_supports(*c)) {
    info[0] = 3;
    return s3_describe(*c, info);
  }
  info[0] = 2;
  return 0;
}

int demfi_conv2d(const demfi_conv_t* c, void* stream) {
  if (check_device()) return 3;
  DEMFI_REQUIRE(c != nullptr, "conv2d: null descriptor");
  DEMFI_REQUIRE(c->N > 0 && c->H > 0 && c->W > 0 && c->Hi > 0 && c->Wi > 0, "conv2d: bad output shape");
  DEMFI_REQUIRE(c->KH > 0 && c->KW > 0 && c->stride >= 1, "conv2d: bad kernel");
  DEMFI_REQUIRE(c->nsrc >= 1 && c->nsrc <= DEMFI_MAX_SRC && c->nseg >= 1 && c->nseg <= DEMFI_MAX_SEG, "conv2d: bad nsrc/nseg");
  DEMFI_REQUIRE(c->cout_pad > 0 && c->cout_pad % 4 == 0, "conv2d: cout_pad must be a multiple of 4");
  DEMFI_REQUIRE((c->Hi + 2 * c->pad_h - c->KH) / c->stride + 1 == c->H && (c->Wi + 2 * c->pad_w - c->KW) / c->stride + 1 == c->W,
                "conv2d: output size %dx%d inconsistent with input %dx%d", c->H, c->W, c->Hi, c->Wi);
  for (int s = 0; s < c->nsrc; ++s) {
    const demfi_src_t& S = c->src[s];
    DEMFI_REQUIRE(S.ptr && S.C > 0 && S.C % 4 == 0 && S.ld % 4 == 0 && S.ld >= S.C && ((uintptr_t)S.ptr % 16) == 0,
                  "conv2d: source %d must be a 16-byte aligned NHWC slice with C%%4==0 (C=%d ld=%d)", s, S.C, S.ld);
    DEMFI_REQUIRE(S.up == 0 || (S.up == 1 && c->Hi % 2 == 0 && c->Wi % 2 == 0), "conv2d: bad up-sampling on source %d", s);
  }
  for (int s = 0; s < c->nseg; ++s) {
    const demfi_seg_t& G = c->seg[s];
    DEMFI_REQUIRE(G.dst && G.ch0 >= 0 && G.nch > 0 && G.ch0 % 4 == 0 && G.nch % 4 == 0 && G.ch0 + G.nch <= c->cout_pad,
                  "conv2d: segment %d channel range [%d,+%d) invalid", s, G.ch0, G.nch);
    DEMFI_REQUIRE(G.dst_ld % 4 == 0 && ((uintptr_t)G.dst % 16) == 0, "conv2d: segment %d destination misaligned", s);
    DEMFI_REQUIRE(G.act >= DEMFI_ACT_NONE && G.act <= DEMFI_ACT_GRU, "conv2d: segment %d bad activation", s);
    if (G.act == DEMFI_ACT_SIGMOID_MUL || G.act == DEMFI_ACT_GRU) DEMFI_REQUIRE(G.res, "conv2d: segment %d needs res", s);
    if (G.act == DEMFI_ACT_GRU) DEMFI_REQUIRE(G.res2 && G.res2_ld % 4 == 0, "conv2d: segment %d needs res2", s);
    if (G.res) DEMFI_REQUIRE(G.res_ld % 4 == 0 && ((uintptr_t)G.res % 16) == 0, "conv2d: segment %d res misaligned", s);
    if (G.store == DEMFI_STORE_PIXEL_SHUFFLE2)
      DEMFI_REQUIRE(G.nch % 16 == 0 && G.res == nullptr, "conv2d: pixel-shuffle segment %d must have nch%%16==0 and no res", s);
  }
  DEMFI_REQUIRE(c->wpack && c->bias, "conv2d: null weights");
  {
    bool any_s16 = false;
    for (int s = 0; s < c->nsrc; ++s) {
      DEMFI_REQUIRE(c->src[s].fmt == DEMFI_FMT_F32 || c->src[s].fmt == DEMFI_FMT_S16, "conv2d: source %d has an unknown format", s);
      if (c->src[s].fmt == DEMFI_FMT_S16) {
        any_s16 = true;
        DEMFI_REQUIRE(c->src[s].C % 32 == 0 && c->src[s].up == 0, "conv2d: an S16 source needs C %% 32 == 0 and no up-sampling (source %d)", s);
      }
    }
    for (int s = 0; s < c->nseg; ++s) {
      DEMFI_REQUIRE((c->seg[s].fmt & ~7) == 0, "conv2d: segment %d has unknown format bits", s);
      if (c->seg[s].fmt != 0) {
        any_s16 = true;
        DEMFI_REQUIRE(c->seg[s].ch0 % 32 == 0 && c->seg[s].nch % 32 == 0, "conv2d: an S16 segment needs ch0 and nch multiples of 32 (segment %d)", s);
      }
    }
    if (any_s16)
      DEMFI_REQUIRE(((c->kind == DEMFI_CONV_TC16 && g_opt_gen.load() == 3) || c->kind == DEMFI_CONV_TC16W || c->kind == DEMFI_CONV_TC16P) &&
                        s3_supports(*c) && s3_s16_ok(*c),
                    "conv2d: the S16 activation format is only implemented by the conv_s3 kernel (stride 1, TMA epilogue)");
  }
  DEMFI_REQUIRE(c->kind != DEMFI_CONV_TC, "conv2d: DEMFI_CONV_TC (the first-generation 3xTF32 kernel) was retired in round 2: use DEMFI_CONV_TC16");
  if (c->kind == DEMFI_CONV_TC16) {
    if (g_opt_gen.load() == 3 && s3_supports(*c)) return launch_conv_s3(*c, (cudaStream_t)stream);
    return launch_conv_h3(*c, (cudaStream_t)stream);
  }
  if (c->kind == DEMFI_CONV_TC16W || c->kind == DEMFI_CONV_TC16P) {
    DEMFI_REQUIRE(s3_supports(*c), "conv2d: DEMFI_CONV_TC16W / TC16P need a convolution conv_s3 supports (stride 1, no up-sampled source; "
                                   "TC16P: 32 or 64 output channels)");
    return launch_conv_s3(*c, (cudaStream_t)stream);
  }
  DEMFI_REQUIRE(c->kind == DEMFI_CONV_FFMA, "conv2d: unknown kind %d", c->kind);
  return launch_conv_ffma(*c, (cudaStream_t)stream);
}

}  // extern "C"
