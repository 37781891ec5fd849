/* demfi_b200 -- C ABI of the B200-native DeMFI-Net forward / recursive-boosting hot path.
 *
 * The reference (JihyongOh/DeMFI) has no FFI layer: its hot path is the Python nn.Module
 * `DeMFInet` (DeMFInet.py:13-179) issuing ATen/cuDNN library calls.  This header is the
 * boundary a maintainer binds instead (ctypes stub: demfi_b200/_abi.py).
 * Every entry point replaces a group of reference call sites, cited per function.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no torch types.  All tensor pointers are DEVICE
 *     pointers to fp32 unless the name says `host`.
 *   - Activations are NHWC: element (n, y, x, c) at ptr[((n*H + y)*W + x)*ld + c]; `ld` is the
 *     pixel stride in floats (multiple of 4) so a tensor may be a channel slice of a wider
 *     buffer (that is how every torch.cat of the reference disappears).
 *   - Functions return 0 on success, non-zero on error; demfi_last_error() gives the message
 *     (thread-local).  Nothing here allocates device memory or synchronises the device; work
 *     is enqueued on `stream` (a cudaStream_t passed as void*).
 *   - There is no CPU fallback: every function fails if the device is not sm_100.
 */
#ifndef DEMFI_B200_H_
#define DEMFI_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DEMFI_ABI_VERSION 1
#define DEMFI_MAX_SRC 4
#define DEMFI_MAX_SEG 4

/* epilogue activation of one output segment; v = acc + bias (+ res when res != NULL and the
 * mode is not MUL/GRU) */
enum {
  DEMFI_ACT_NONE = 0,        /* out = v                                   */
  DEMFI_ACT_RELU = 1,        /* out = max(v, 0)                           */
  DEMFI_ACT_TANH = 2,        /* out = tanh(v)                             */
  DEMFI_ACT_SIGMOID = 3,     /* out = sigmoid(v)                          */
  DEMFI_ACT_SIGMOID_MUL = 4, /* out = sigmoid(acc+bias) * res    (GRU r*h, DeMFInet.py:846-847) */
  DEMFI_ACT_GRU = 5          /* out = (1-res2)*res + res2*tanh(acc+bias)  (DeMFInet.py:847-848) */
};
enum {
  DEMFI_STORE_NHWC = 0,
  DEMFI_STORE_PIXEL_SHUFFLE2 = 1 /* nn.PixelShuffle(2), DeMFInet.py:229: accumulator channel
                                    q*(nch/4)+c of pixel (y,x) -> dst pixel (2y+q/2, 2x+q%2), channel c */
};
/* Activation storage formats.  F32: plain fp32.  S16 ("split fp16"): every group of 32 channels (128 bytes of the pixel)
 * holds 32 fp16 "hi" values followed by 32 fp16 "lo" values; channel c of the group is hi[c] + lo[c] / 2048
 * (hi = fp16(v), lo = fp16((v - hi) * 2048): ~22 significant bits, |v| < 65504).  It is exactly what the tensor-core
 * kernel feeds to its kind::f16 MMAs (3xFP16 split), so a convolution whose input was written in S16 by the previous
 * convolution needs no conversion pass at all: the TMA tile is the MMA operand.  Same bytes per channel as fp32.
 * Only DEMFI_CONV_TC16 (conv_s3 kernel) reads or writes S16; every other operator takes F32. */
enum { DEMFI_FMT_F32 = 0, DEMFI_FMT_S16 = 1 };
/* demfi_seg_t.fmt bits */
enum { DEMFI_SEG_DST_S16 = 1, DEMFI_SEG_RES_S16 = 2, DEMFI_SEG_RES2_S16 = 4 };
enum {
  DEMFI_CONV_FFMA = 0, /* CUDA-core fp32 implicit GEMM (exact fp32)                                   */
  DEMFI_CONV_TC = 1,   /* RETIRED in round 2 (first-generation 3xTF32 kernel): every entry point rejects it     */
  DEMFI_CONV_TC16 = 2, /* tcgen05 kind::f16, 3xFP16 split + halo-tile activation staging; stride 1|2  */
  DEMFI_CONV_TC16W = 3, /* the same arithmetic on the conv_s3 kernel only (stride 1), with 97..128 output channels kept in
                          ONE N block (an N' = 256 MMA pair per k-step; weights packed for that blocking)              */
  DEMFI_CONV_TC16P = 4  /* the same arithmetic on CTA PAIRS (2-CTA clusters, tcgen05 cta_group::2, M = 256 = two pixel tiles
                          per MMA): 32 or 64 output channels, stride 1; each CTA holds half of the weight rows (weights
                          packed per CTA rank)                                                                          */
};

/* one input of a (virtually concatenated) convolution */
typedef struct {
  const float* ptr; /* NHWC, [N, Hi>>up, Wi>>up, ld] */
  int32_t C;        /* channels this source contributes to K (multiple of 4) */
  int32_t ld;
  int32_t up;       /* 1: source is at half the conv-input resolution and is read through
                       nearest-neighbour x2 up-sampling (nn.UpsamplingNearest2d, DeMFInet.py:573) */
  int32_t fmt;      /* DEMFI_FMT_F32 | DEMFI_FMT_S16 (S16: C and the slice offset multiples of 32 channels) */
} demfi_src_t;

/* one destination of a channel range [ch0, ch0+nch) of the accumulator */
typedef struct {
  float* dst;
  const float* res;  /* optional operand, same pixel indexing as dst (NHWC, res_ld) */
  const float* res2; /* second operand (GRU z) */
  int32_t dst_ld, res_ld, res2_ld;
  int32_t ch0, nch;  /* multiples of 4 */
  int32_t act, store;
  int32_t fmt;       /* bit mask DEMFI_SEG_DST_S16 | DEMFI_SEG_RES_S16 | DEMFI_SEG_RES2_S16 (0: everything fp32) */
} demfi_seg_t;

/* A convolution = every nn.Conv2d / nn.Conv3d([1,3,3]) call of DeMFInet.py (section 2.1 of
 * SURVEY.md lists the 27 shapes), with the torch.cat feeding it expressed as `src[]` and the
 * pointwise ops following it expressed as `seg[]`. */
typedef struct {
  int32_t N, H, W;   /* output batch / rows / cols                                   */
  int32_t Hi, Wi;    /* conv-input grid (after any up-sampling)                       */
  int32_t KH, KW, stride, pad_h, pad_w;
  int32_t nsrc, nseg;
  int32_t cout_pad;  /* accumulator channels (padded Cout, multiple of 16)             */
  int32_t kind;      /* DEMFI_CONV_FFMA | DEMFI_CONV_TC | DEMFI_CONV_TC16 | _TC16W | _TC16P */
  demfi_src_t src[DEMFI_MAX_SRC];
  demfi_seg_t seg[DEMFI_MAX_SEG];
  const float* wpack; /* device, produced by demfi_pack_weights for the same `kind`    */
  const float* bias;  /* device, cout_pad floats                                      */
} demfi_conv_t;

/* ---- library ------------------------------------------------------------------------- */
int demfi_version(void);
const char* demfi_last_error(void);
/* 0 iff `device` is compute capability 10.0 (B200).  Replaces main.py:157-159's bare
 * torch.cuda.set_device: there is no other backend to fall back to. */
int demfi_device_check(int device);

/* ---- weights ------------------------------------------------------------------------- */
/* Number of floats demfi_pack_weights writes.  src_C[nsrc] are the demfi_src_t.C values the conv
 * will be launched with (the tcgen05 layout chunks K per source). */
size_t demfi_packed_weight_floats(int32_t kind, int32_t KH, int32_t KW, const int32_t* src_C, int32_t nsrc,
                                  int32_t cout_pad);
/* HOST -> HOST repack of one nn.Conv weight (state_dict layout [Co, Ci, KH, KW], fp32) into
 * the layout the kernels stream.  in_map[k] (k < k_total = sum(src_C)) is the reference input
 * channel that internal channel k carries, or -1 for padding; out_map[n] (n < cout_pad) likewise
 * for output channels.  kind FFMA: [tap][k][cout_pad] fp32.  kind TC: per (32-channel chunk, tap) a
 * [cout_pad x 32] K-major tile, 128-byte-swizzled, stored twice: tf32 "hi" part and fp32
 * residual "lo" part (3xTF32 split).  kind TC16: per (N block, 32-channel chunk, tap) a
 * [2*N x 32] fp16 K-major tile, 64-byte-swizzled: rows 0..N-1 hold h = fp16(w), rows N..2N-1 hold
 * l = fp16((w - h) * 2048) (3xFP16 split; the returned count is in floats, two fp16 per float). */
int demfi_pack_weights(int32_t kind, const float* w_oihw_host, int32_t Co, int32_t Ci, int32_t KH, int32_t KW,
                       const int32_t* in_map, const int32_t* src_C, int32_t nsrc, const int32_t* out_map,
                       int32_t cout_pad, float* out_host);

/* DEVICE -> DEVICE form of demfi_pack_weights for kind DEMFI_CONV_TC16 and ONE source of src_c >= Ci channels (identity channel
 * maps; input channels Ci..src_c-1 and output channels Co..cout_pad-1 are zero padding), stream-ordered: the training step
 * (main.py:443-445: backward, optimizer step, next forward) repacks every weight once per step and must not wait for the GPU
 * to do it on the host.  Same bytes as demfi_pack_weights for in-range weights; values beyond the fp16 range saturate. */
int demfi_pack_weights_device(int32_t kind, const float* w_oihw_dev, int32_t Co, int32_t Ci, int32_t KH, int32_t KW,
                              int32_t src_c, int32_t cout_pad, float* out_dev, void* stream);

/* ---- convolution (DeMFInet.py: every nn.Conv call; see demfi_conv_t) -------------------- */
int demfi_conv2d(const demfi_conv_t* conv, void* stream);

/* Host-only (no GPU needed): what demfi_conv2d would do with this descriptor.  info[0] = kernel (0 conv_ffma, 1 conv_tc,
 * 2 conv_h3, 3 conv_s3); for conv_s3 also info[1] = TMA-store epilogue (else the generic per-thread one), [2] = weights
 * resident in shared memory, [3] = halo-tile buffers, [4] = weight-ring slots, [5] = (chunk, tap) stages per slot,
 * [6] = N blocks, [7] = dynamic shared memory in bytes, [8] = stages per accumulation segment, [9] = stages per tile,
 * [10] = epilogue entries per N block (1: one result per block; > 1: every 32-channel box has its own activation / operands /
 * format / destinations), [11] = N block width, [12] = 1 for the CTA-pair kernel (DEMFI_CONV_TC16P).
 * Lets a plan be checked for silent fall-backs to slower paths without launching anything. */
int demfi_conv_describe(const demfi_conv_t* conv, int32_t info[16]);

/* ---- memory-bound operators ------------------------------------------------------------ */
/* Input unpack.  x is the caller's [B,3,4,H,W] NCHW-T tensor (DeMFInet.py:51-55).  Writes
 *  - s2d:   [B,H/2,W/2,48] pixel_reshuffle(cat(B0,B1,B-1,B2), 2)   (DeMFInet.py:234-235,290-316)
 *  - f12a/f12b: the 12 frame channels (order t*3+c) into two NHWC slices (Mixer / D2 inputs,
 *    DeMFInet.py:119,151-155); either may be NULL
 *  - mean01: [B,3,H,W] NCHW = mean(x[:,:,0:2], dim=2)               (DeMFInet.py:178) */
int demfi_pack_input(const float* x, int32_t B, int32_t H, int32_t W, float* s2d, float* f12a, int32_t f12a_ld,
                     float* f12b, int32_t f12b_ld, float* mean01, void* stream);

/* Complementary flow reversal (CFR_flow_t_align + fwarp + sample_one, DeMFInet.py:606-729).
 * fo: NHWC [B,H,W,fo_ld], channels 0-1 flow_01, 2-3 flow_10.  acc: [B,H,W,8] scratch, must be
 * zero on entry to splat (the splat adds with fp32 red.global.add).  t: device [B]. */
int demfi_cfr_splat(const float* fo, int32_t fo_ld, const float* t, int32_t B, int32_t H, int32_t W, float* acc,
                    void* stream);
/* out: 4 channels (flow_t0.x, flow_t0.y, flow_t1.x, flow_t1.y) at out[pix*out_ld]. */
int demfi_cfr_finalize(const float* acc, const float* t, int32_t B, int32_t H, int32_t W, float* out,
                       int32_t out_ld, void* stream);

/* bwarp + Eq.(2) blend (bwarp DeMFInet.py:732-766; blend :66-71, :90-93, :146-149):
 *   out = ((1-t)*o*bwarp(a, flow[0:2]) + t*(1-o)*bwarp(b, flow[2:4])) / ((1-t)*o + t*(1-o)),
 *   o = sigmoid(occ_logit).  C channels (multiple of 4 for the vector path; C=3 uses the
 *   scalar path).  a, b, out: NHWC with their own ld; flow: 4 channels at flow[pix*flow_ld];
 *   occ: 1 channel at occ[pix*occ_ld]; t: device [B].
 *   occ_out (optional): receives sigmoid(occ_logit) at occ_out[pix*occ_out_ld]. */
int demfi_bwarp_blend(const float* a, int32_t a_ld, const float* b, int32_t b_ld, const float* flow, int32_t flow_ld,
                      const float* occ, int32_t occ_ld, const float* t, int32_t B, int32_t H, int32_t W, int32_t C,
                      float* out, int32_t out_ld, float* occ_out, int32_t occ_out_ld, void* stream);

/* Pixel-wise blending of the boosting loop (PWB, DeMFInet.py:146-149) fused with the concatenations around it
 * (DeMFInet.py:151-155): img = [S0' (3) pad | S1' (3) pad] (8 channels per pixel), fo = the iteration's refined
 * [flow_t0 (2), flow_t1 (2), occlusion logit, pad 3]; writes out = [St (3), sigmoid(occ) | flow_t0, flow_t1] (8 channels)
 * -- the same Eq.(2) arithmetic as demfi_bwarp_blend with C = 3, with every access a whole 16 / 32-byte unit. */
int demfi_pwb(const float* img, int32_t img_ld, const float* fo, int32_t fo_ld, const float* t, int32_t B, int32_t H, int32_t W,
              float* out, int32_t out_ld, void* stream);

/* FGAC sampling (FGAC.forward step (i), DeMFInet.py:403-419 + bilinear_sampler :499-514) with
 * rr = sr = 0: out(y,x,:) = bilinear(ref_k, at absolute position (flow.x, flow.y)), zeros
 * outside.  The correlation/softmax that follows in the reference runs over one element and is
 * identically 1 (DeMFInet.py:438-443), so it is not materialised. */
int demfi_fgac_sample(const float* refk, int32_t refk_ld, const float* flow, int32_t flow_ld, int32_t B, int32_t H,
                      int32_t W, int32_t C, float* out, int32_t out_ld, void* stream);
/* Eq.(4), DeMFInet.py:452: out = w*src + (1-w)*e, w one channel (already sigmoid-ed). */
int demfi_fgac_blend(const float* w, int32_t w_ld, const float* src, int32_t src_ld, const float* e, int32_t e_ld,
                     int64_t npix, int32_t C, float* out, int32_t out_ld, void* stream);

/* Channel-slice copy between NHWC buffers: dst[p*dst_ld + j] = act(src[p*src_ld + j]), j < nch.
 * Replaces the small torch.cat assemblies (DeMFInet.py:117-123, 151-155). act: NONE or SIGMOID. */
int demfi_copy_channels(const float* src, int32_t src_ld, float* dst, int32_t dst_ld, int32_t nch, int64_t npix,
                        int32_t act, void* stream);
/* Several channel-slice copies into ONE destination buffer in one pass over its pixels (the torch.cat assemblies of
 * ref_list / Agg3, DeMFInet.py:117-123, 151-155): for every part k, dst[p*dst_ld + dst_c0[k] + j] = src[k][p*src_ld[k] + j],
 * j < nch[k].  At most DEMFI_MAX_PARTS parts; all parts index the same npix pixels. */
#define DEMFI_MAX_PARTS 8
typedef struct {
  const float* src;
  int32_t src_ld, nch, dst_c0, reserved;
} demfi_part_t;
int demfi_gather_channels(const demfi_part_t* parts, int32_t nparts, float* dst, int32_t dst_ld, int64_t npix, void* stream);
/* Channel mean of absolute values per pixel: out[p] = mean_{c<C} |a[p*a_ld + c] - (b ? b[p*b_ld + c] : 0)|, out a dense
 * [npix] map.  The reduction half of the FGAC difference map (DeMFInet.py:456-462) and of the four visualisation maps
 * (:465-491); the per-sample min-max normalisation that follows is done by the host mirror.  C % 4 == 0. */
int demfi_channel_absmean(const float* a, int32_t a_ld, const float* b, int32_t b_ld, int64_t npix, int32_t C, float* out,
                          void* stream);
/* nn.UpsamplingNearest2d(scale_factor=2) (DeMFInet.py:573): src [B,Hs,Ws,C of src_ld] -> dst [B,2Hs,2Ws,C of dst_ld].
 * Materialises the UNet decoder inputs so that dec1-3 run on the tensor-core conv. */
int demfi_upsample2x(const float* src, int32_t src_ld, int32_t B, int32_t Hs, int32_t Ws, int32_t C, float* dst,
                     int32_t dst_ld, void* stream);
/* NHWC slice -> NCHW tensor [B,C,H,W] (the tensors DeMFInet.forward returns, DeMFInet.py:170-179). */
int demfi_export_nchw(const float* src, int32_t src_ld, int32_t B, int32_t H, int32_t W, int32_t C, int32_t act,
                      float* dst, void* stream);
/* NCHW [B,C,H,W] -> NHWC slice (test helper and teacher-forced entry). */
int demfi_import_nchw(const float* src, int32_t B, int32_t H, int32_t W, int32_t C, float* dst, int32_t dst_ld,
                      void* stream);

/* ---- backward of a convolution layer (training row; first correct path) ------------------------ */
/* The reference obtains these from autograd through nn.Conv2d (main.py:443).  For y = act(conv(x, W) + b), stride 1, 'same':
 *   dz = dy * act'(y)                    demfi_act_backward (act: NONE / RELU / TANH / SIGMOID, from the stored output y)
 *   dx = conv(dz, W rotated by 180 degrees with Cin and Cout exchanged): the forward kernel, demfi_conv2d, on re-packed weights
 *   dW, db                               demfi_conv2d_wgrad: dw[co][ci][ky][kx] += sum_p dz[p][co] * x[p + (ky,kx) - pad][ci]
 *                                        (OIHW fp32, the layout of the parameter), dbias[co] += sum_p dz[p][co] (may be NULL).
 * H, W are the OUTPUT plane (that of dz); x is [N, H*stride, W*stride, x_ld] (stride 1: 'same'; stride 2: the UNet encoders).
 * x, dz: NHWC with row strides x_ld / dz_ld.  Both gradients are ACCUMULATED (+=) with fp32 atomics: zero them first. */
int demfi_conv2d_wgrad(const float* x, int32_t x_ld, int32_t Cin, const float* dz, int32_t dz_ld, int32_t Cout, int32_t N,
                       int32_t H, int32_t W, int32_t KH, int32_t KW, int32_t pad_h, int32_t pad_w, int32_t stride, float* dw,
                       float* dbias, void* stream);
/* dx of the stride-2 encoders (Refine_Module.enc1-3, 4x4 / stride 2 / pad 1), CUDA cores, gather form: dz [N,H,W,dz_ld] is the
 * gradient at the conv's pre-activation output, w_oihw the parameter itself ([Cout,Cin,KH,KW] fp32 on the device), dx
 * [N, H*stride, W*stride, dx_ld] is written. */
int demfi_conv2d_dgrad_strided(const float* dz, int32_t dz_ld, const float* w_oihw, int32_t Cin, int32_t Cout, int32_t N, int32_t H,
                               int32_t W, int32_t KH, int32_t KW, int32_t pad_h, int32_t pad_w, int32_t stride, float* dx,
                               int32_t dx_ld, void* stream);
int demfi_act_backward(const float* dy, int32_t dy_ld, const float* y, int32_t y_ld, int64_t npix, int32_t C, int32_t act,
                       float* out, int32_t out_ld, void* stream);

/* Backward of demfi_bwarp_blend (same a, b, flow, occ, t as the forward call; dout = gradient of its output): da, db are
 * ACCUMULATED with fp32 atomics at the four bilinear corners (zero them first; either may be NULL), dflow (4 channels:
 * d/d flow of the a-warp x, y, then of the b-warp) and docc (1 channel, gradient of the occlusion LOGIT) are written.  The
 * validity mask of bwarp is piecewise constant and only gates the gradients (autograd through DeMFInet.py:757-766). */
int demfi_bwarp_blend_backward(const float* a, int32_t a_ld, const float* b, int32_t b_ld, const float* flow, int32_t flow_ld,
                               const float* occ, int32_t occ_ld, const float* t, const float* dout, int32_t dout_ld, int32_t B,
                               int32_t H, int32_t W, int32_t C, float* da, int32_t da_ld, float* db, int32_t db_ld, float* dflow,
                               int32_t dflow_ld, float* docc, int32_t docc_ld, void* stream);
/* Backward of demfi_fgac_sample: drefk accumulated (atomics, may be NULL), dflow (2 channels) written. */
int demfi_fgac_sample_backward(const float* refk, int32_t refk_ld, const float* flow, int32_t flow_ld, const float* dout,
                               int32_t dout_ld, int32_t B, int32_t H, int32_t W, int32_t C, float* drefk, int32_t drefk_ld,
                               float* dflow, int32_t dflow_ld, void* stream);

/* Backward of demfi_cfr_splat + demfi_cfr_finalize: fo, t as in the forward; acc = the accumulators the forward splat left
 * ([B,H,W,8], unchanged by finalize); gout = gradient of finalize's 4-channel output (flow_t0, flow_t1); gacc = scratch
 * [B,H,W,8] (written); dfo = gradient w.r.t. the 4 channels (flow_01, flow_10) of fo (written).  Gathers only, no atomics;
 * floor() is taken exactly as in the forward and has no gradient (autograd through DeMFInet.py:606-729). */
int demfi_cfr_backward(const float* fo, int32_t fo_ld, const float* t, const float* acc, const float* gout, int32_t gout_ld,
                       int32_t B, int32_t H, int32_t W, float* gacc, float* dfo, int32_t dfo_ld, void* stream);

/* Eq.(4) backward (forward: demfi_fgac_blend, out = w*src + (1-w)*e): dw[p] = sum_c gout*(src - e) (written, 1 channel),
 * dsrc = gout*w, de = gout*(1-w) (written; any of the three may be NULL). */
int demfi_fgac_blend_backward(const float* w, int32_t w_ld, const float* src, int32_t src_ld, const float* e, int32_t e_ld,
                              const float* gout, int32_t gout_ld, int64_t npix, int32_t C, float* dw, int32_t dw_ld, float* dsrc,
                              int32_t dsrc_ld, float* de, int32_t de_ld, void* stream);
/* nn.UpsamplingNearest2d(2) backward: gsrc[n,y,x,:] = sum of gdst over the 2x2 children (written). */
int demfi_upsample2x_backward(const float* gdst, int32_t gdst_ld, int32_t B, int32_t Hs, int32_t Ws, int32_t C, float* gsrc,
                              int32_t gsrc_ld, void* stream);
/* One term of the reconstruction losses (nn.L1Loss, main.py:404-440): out[0] (device double) = sum_i |pred[i] - target[i]|
 * over n contiguous elements, fixed summation order; when grad != NULL also grad[i] = grad_scale * sign(pred[i] - target[i])
 * (what autograd sends back for  grad_scale * n * mean|.|;  the caller folds lambda / 3 / n into grad_scale).
 * workspace: at least demfi_l1_sum_workspace(n) bytes of device memory. */
int64_t demfi_l1_sum_workspace(int64_t n);
int demfi_l1_sum(const float* pred, const float* target, int64_t n, float grad_scale, float* grad, void* workspace,
                 int64_t workspace_bytes, double* out, void* stream);
/* torch.optim.Adam (main.py:179-180), one parameter tensor, in place: g' = grad + weight_decay*param; exp_avg, exp_avg_sq
 * updated; param -= lr / (1 - beta1^step) * exp_avg / (sqrt(exp_avg_sq) / sqrt(1 - beta2^step) + eps).  step counts from 1. */
int demfi_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1, float beta2,
                    float eps, float weight_decay, int32_t step, void* stream);

/* ---- evaluation metrics (the consumer right after the hot path) ------------------------- */
/* PSNR / SSIM sums of predicted frames against their targets as the reference's evaluation loop computes them
 * (main.py:763-771 with utils.py:652-705, 718-721): pred, target are NCHW [B,C,H,W] fp32 in [-1,1] on the device (what
 * DeMFInet.forward returns / the ground-truth batch).  Both are scaled to 0..255 inside the kernel: the prediction in fp64 and
 * rounded half-to-even (np.around of the float64 array), the target in fp32 arithmetic and NOT rounded when target_mode = 0
 * (a ground truth, main.py:765-766) or exactly like the prediction when target_mode = 1 (a second network output, the
 * "PSNR(ours, reference)" parity figure).  out (device, 2*B doubles): out[2b] = sum of squared errors over the C*H*W elements
 * of image b, out[2b+1] = sum of the SSIM map over the C*(H-10)*(W-10) window positions; the host divides and takes the log
 * (demfi_b200/metrics.py).  fp64 arithmetic, fixed summation order (bit-reproducible).  H, W >= 11.  workspace: device memory
 * of at least demfi_frame_metrics_workspace(B, C, H, W) bytes (returns -1 for an invalid shape). */
int64_t demfi_frame_metrics_workspace(int32_t B, int32_t C, int32_t H, int32_t W);
int demfi_frame_metrics(const float* pred, const float* target, int32_t B, int32_t C, int32_t H, int32_t W, int32_t target_mode,
                        void* workspace, int64_t workspace_bytes, double* out, void* stream);

/* ---- introspection ---------------------------------------------------------------------- */
/* Number of kernel launches issued by this library since load (bench.py's gpu_launches). */
uint64_t demfi_launch_count(void);
/* Runtime options (diagnostics / measurement): "tc_mask_hi" (1, default: the activation tile is split as
 * hi = tf32 round-to-nearest, written back to shared memory, lo = a - hi; 0: hi = truncation, i.e. the raw
 * fp32 tile is fed and the tensor core's own truncation is relied on -- measured identical operand
 * behaviour on B200, slightly more biased residual), "tc_split" (3: 3xTF32, fp32-parity mode, default; 1: single-pass TF32, NOT parity
 * grade, for measurement only), "tc_flush" (K stages of 32 channels accumulated inside the tensor core
 * before the partial sum is drained and added in fp32 round-to-nearest; default 10 (segments are balanced: 18 stages -> 2 x 9), 0 = whole K), "tc_comp_milli" (gain correction of the truncating tensor-core accumulation in
 * units of 1e-3 * 2^-24 per chained MMA; default 270 = the measured -0.27*2^-24 bias per MMA on B200),
 * "tc_a_tmem" (1: activation operand through tensor memory, 0: through shared memory), "tc_stages" /
 * "tc_grid" (caps on pipeline depth / persistent CTAs, 0 = auto), "tc_diag" (timing diagnostics bitmask),
 * "tc_gen" (DEMFI_CONV_TC16 kernel: 3 = conv_s3 where it applies (default), 2 = conv_h3 only), "tc_pdl" (1: programmatic
 * dependent launch between consecutive conv_s3 kernels; default 0 -- measured slightly slower on the full forward),
 * "tc_prefetch" (conv_s3: every activation chunk is also prefetched into L2 this many tiles ahead; default 0 -- measured: no gain),
 * "wgrad_kind" (demfi_conv2d_wgrad: 1 = mma.sync 3xTF32 kernel (default), 0 = CUDA-core fp32 kernel).  "tc_diag" bits used by
 * tools/ and tests/: 128 role timers, 32768 one operand tile for the skip operand of a lean layer (instead of two used in turn),
 * 65536 lean activation kernels with the activation read at run time (instead of the compile-time ones).
 * Returns non-zero for an unknown option. */
int demfi_set_option(const char* name, int32_t value);
int demfi_get_option(const char* name, int32_t* value);
/* Role timers of the last tensor-core conv launched with "tc_diag" & 128 (measurement only): host[ctas][16]
 * cycle counts per persistent CTA -- 0 splitter total, 1 its wait for the TMA stage, 2 its wait for a free A slot,
 * 3 tcgen05.st + arrive; 4 epilogue total, 5 its wait for an accumulator, 6 store phase; 8 producer total,
 * 9 its wait for a free stage; 12 MMA issuer total, 13 its wait for a drained accumulator, 14 its wait for A.
 * Synchronises the device. */
int demfi_tc_debug_read(int64_t* host, int32_t ctas);

#ifdef __cplusplus
}
#endif
#endif /* DEMFI_B200_H_ */
