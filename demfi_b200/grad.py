"""Differentiable convolution layer through the C ABI -- the first slice of the training row.

`conv2d(x, weight, bias, act, stride)` computes `act(F.conv2d(x, weight, bias, ...))` for the layer shapes of DeMFI-Net (stride 1
with 'same' padding: every nn.Conv2d / nn.Conv3d[1,k,k] of `DeMFInet.py`; stride 2, 4x4, padding 1: the three UNet encoders,
whose dx runs on the CUDA-core `demfi_conv2d_dgrad_strided`) and is differentiable:

  forward   demfi_conv2d            tcgen05 kernel, 3xFP16 split (fp32 parity), fused bias + activation
  dz        demfi_act_backward      dy * act'(y) from the stored output
  dx        demfi_conv2d            the SAME forward kernel on the weights rotated by 180 degrees with Cin/Cout exchanged
  dW, db    demfi_conv2d_wgrad      CUDA-core fp32 kernel (first correct path; accumulates with fp32 atomics)

What the reference gets from autograd through `nn.Conv2d` (main.py:443).  NCHW tensors at the boundary (imported / exported
with the ABI's layout kernels), NHWC inside.  No CPU or ATen fallback: CPU tensors raise.  `train_net.forward_train` builds the
whole network's training graph from this function.  Weights are packed on the device (demfi_pack_weights_device, stream-ordered;
DEMFI_GRAD_HOST_PACK=1: on the host, as the first version did) once per change of the parameter (its version counter).
"""
from __future__ import annotations

import os
import weakref
from typing import Tuple

import numpy as np
import torch

from . import _abi as A

ACTS = {"none": A.ACT_NONE, "relu": A.ACT_RELU, "tanh": A.ACT_TANH, "sigmoid": A.ACT_SIGMOID}
_ru = lambda v, m: (v + m - 1) // m * m
_MAX_COUT = 256    # demfi_conv2d: cout_pad <= 256


def _stream(dev):
    return torch.cuda.current_stream(dev).cuda_stream


def _pack(w: np.ndarray, b: np.ndarray, src_c: int, dev) -> Tuple[torch.Tensor, torch.Tensor, int]:
    """OIHW fp32 -> the tensor-core kernel's packed hi/lo layout for ONE source of `src_c` (>= Cin, zero-weight padding) channels"""
    lib = A.lib()
    Co, Ci, KH, KW = w.shape
    cout_pad = _ru(Co, 16)
    sC = (A.i32 * 1)(src_c)
    in_map = (A.i32 * src_c)(*(list(range(Ci)) + [-1] * (src_c - Ci)))
    out_map = (A.i32 * cout_pad)(*(list(range(Co)) + [-1] * (cout_pad - Co)))
    n = lib.demfi_packed_weight_floats(A.CONV_TC16, KH, KW, sC, 1, cout_pad)
    packed = np.empty(n, dtype=np.float32)
    w = np.ascontiguousarray(w, dtype=np.float32)
    A.check(lib.demfi_pack_weights(A.CONV_TC16, w.ctypes.data, Co, Ci, KH, KW, in_map, sC, 1, out_map, cout_pad,
                                   packed.ctypes.data), "demfi_pack_weights")
    bias = np.zeros(cout_pad, dtype=np.float32)
    bias[:Co] = b
    return torch.from_numpy(packed).to(dev), torch.from_numpy(bias).to(dev), cout_pad


def _pack_dev(w: torch.Tensor, b, src_c: int, dev) -> Tuple[torch.Tensor, torch.Tensor, int]:
    """the same packed layout made ON the device from a device-resident OIHW weight (demfi_pack_weights_device): stream-ordered,
    no host round trip -- every weight is repacked once per optimizer step, and a `.cpu()` per layer made the host wait for the
    GPU ~400 times per training step"""
    lib = A.lib()
    Co, Ci, KH, KW = w.shape
    cout_pad = _ru(Co, 16)
    sC = (A.i32 * 1)(src_c)
    n = lib.demfi_packed_weight_floats(A.CONV_TC16, KH, KW, sC, 1, cout_pad)
    w = w.detach().to(torch.float32).contiguous()
    packed = torch.empty(n, dtype=torch.float32, device=dev)
    A.check(lib.demfi_pack_weights_device(A.CONV_TC16, w.data_ptr(), Co, Ci, KH, KW, src_c, cout_pad, packed.data_ptr(), _stream(dev)),
            "demfi_pack_weights_device")
    bias = torch.zeros(cout_pad, dtype=torch.float32, device=dev)
    if b is not None:
        bias[:Co] = b.detach()
    return packed, bias, cout_pad


def _host_pack() -> bool:
    return os.environ.get("DEMFI_GRAD_HOST_PACK", "0") == "1"


def _cache_on() -> bool:
    return os.environ.get("DEMFI_GRAD_PACK_CACHE", "1") == "1"


def _cached(weight: torch.Tensor, key: tuple, bias, make):
    """Packed weights live ON the parameter they were made from (an attribute of the tensor object, or of its base for a
    view such as the [:, :, 0] slice of a Conv3d weight): they die with it and are rebuilt when its version counter moves
    (load_state_dict, torch optimizers and train.Adam all bump it) -- packed once per optimizer step instead of at every
    call.  (A first version keyed a global table on (data_ptr, version, shape): a freed parameter's address is handed to the
    next model, and its stale packed weights with it -- wrong gradients in the second model of a process.)"""
    if not _cache_on():
        return make()
    holder = weight._base if weight._base is not None else weight
    table = holder.__dict__.setdefault("_demfi_packed", {})
    key = key + (weight.data_ptr() - holder.data_ptr(), tuple(weight.shape), tuple(weight.stride()))
    ent = table.get(key)
    bver = None if bias is None else bias._version
    if ent is not None and ent[0] == weight._version and ent[1] == bver and (bias is None or ent[2]() is bias):
        return ent[3]
    out = make()
    table[key] = (weight._version, bver, None if bias is None else weakref.ref(bias), out)
    return out


def _pack_forward(weight: torch.Tensor, bias, src_c: int, dev):
    """packed forward weights of a layer (cached on the parameter, see _cached)"""
    def make():
        if weight.is_cuda and not _host_pack():  # (host tensors: only the host-side tests of the packing helpers get here)
            return _pack_dev(weight, bias, src_c, dev)
        b = bias.detach().cpu().numpy() if bias is not None else np.zeros(weight.shape[0], dtype=np.float32)
        return _pack(weight.detach().cpu().numpy(), b, src_c, dev)
    return _cached(weight, ("f", src_c, str(dev)), bias, make)


def _pack_rotated(weight: torch.Tensor, c0: int, c1: int, src_c: int, dev):
    """packed weights of the dx convolution for input channels [c0, c1): w_t[ci, co, ky, kx] = W[co, ci, KH-1-ky, KW-1-kx]"""
    def make():
        w_rot = weight.detach().flip(2, 3).transpose(0, 1)[c0:c1].contiguous()
        if w_rot.is_cuda and not _host_pack():
            return _pack_dev(w_rot, None, src_c, dev)
        return _pack(w_rot.cpu().numpy(), np.zeros(c1 - c0, dtype=np.float32), src_c, dev)
    return _cached(weight, ("r", c0, c1, src_c, str(dev)), None, make)


def _conv_nhwc(x: torch.Tensor, C_: int, wdev: torch.Tensor, bdev: torch.Tensor, cout_pad: int, k: Tuple[int, int], act: int,
               stride: int = 1):
    """x: [N,Hi,Wi,ld] fp32 NHWC of which the first C_ channels are read; returns [N,Hi/stride,Wi/stride,cout_pad].
    stride 1: 'same' padding; stride 2: the 4x4 / pad 1 encoders of the UNet."""
    N, Hi, Wi, ld = x.shape
    H, W = Hi // stride, Wi // stride
    pad = (k[0] // 2, k[1] // 2) if stride == 1 else (1, 1)
    y = torch.empty(N, H, W, cout_pad, dtype=torch.float32, device=x.device)
    d = A.Conv()
    d.N, d.H, d.W, d.Hi, d.Wi = N, H, W, Hi, Wi
    d.KH, d.KW, d.stride, d.pad_h, d.pad_w = k[0], k[1], stride, pad[0], pad[1]
    d.nsrc, d.nseg, d.cout_pad, d.kind = 1, 1, cout_pad, A.CONV_TC16
    d.src[0].ptr, d.src[0].C, d.src[0].ld = x.data_ptr(), C_, ld
    s = d.seg[0]
    s.dst, s.dst_ld, s.ch0, s.nch, s.act, s.store = y.data_ptr(), cout_pad, 0, cout_pad, act, A.STORE_NHWC
    d.wpack, d.bias = wdev.data_ptr(), bdev.data_ptr()
    A.check(A.lib().demfi_conv2d(d, _stream(x.device)), "demfi_conv2d")
    return y


def _to_nhwc(t: torch.Tensor, ld: int) -> torch.Tensor:
    N, C_, H, W = t.shape
    out = torch.zeros(N, H, W, ld, dtype=torch.float32, device=t.device) if ld != C_ else \
        torch.empty(N, H, W, ld, dtype=torch.float32, device=t.device)
    src = t.contiguous()      # held until the launch below has been issued (a temporary would go back to the allocator first)
    A.check(A.lib().demfi_import_nchw(src.data_ptr(), N, H, W, C_, out.data_ptr(), ld, _stream(t.device)), "import_nchw")
    return out


def _to_nchw(buf: torch.Tensor, C_: int) -> torch.Tensor:
    N, H, W, ld = buf.shape
    out = torch.empty(N, C_, H, W, dtype=torch.float32, device=buf.device)
    A.check(A.lib().demfi_export_nchw(buf.data_ptr(), ld, N, H, W, C_, A.ACT_NONE, out.data_ptr(), _stream(buf.device)), "export_nchw")
    return out


class _Conv2d(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, act: int, stride: int = 1):
        if not x.is_cuda:
            raise A.DemfiError("demfi_b200.grad.conv2d runs on the GPU only (no CPU fallback)")
        Co, Ci, KH, KW = weight.shape
        if x.shape[1] != Ci or (stride == 1 and (KH % 2 == 0 or KW % 2 == 0)) or \
                (stride == 2 and ((KH, KW) != (4, 4) or x.shape[2] % 2 or x.shape[3] % 2)) or stride not in (1, 2):
            raise ValueError("conv2d: the layer shapes of DeMFI-Net only -- stride 1 with an odd 'same' kernel, or the UNet "
                             "encoders (4x4, stride 2, padding 1, even input size)")
        dev = x.device
        with torch.cuda.device(dev):
            cin_pad = _ru(Ci, 8)
            xb = _to_nhwc(x.detach().float(), cin_pad)
            wdev, bdev, cout_pad = _pack_forward(weight, bias, cin_pad, dev)
            yb = _conv_nhwc(xb, cin_pad, wdev, bdev, cout_pad, (KH, KW), act, stride)
            y = _to_nchw(yb, Co)
        ctx.save_for_backward(weight)
        ctx.xb, ctx.yb, ctx.act, ctx.has_bias, ctx.shape, ctx.stride = xb, yb, act, bias is not None, tuple(x.shape), stride
        return y

    @staticmethod
    def backward(ctx, dy):
        (weight,) = ctx.saved_tensors
        xb, yb, act = ctx.xb, ctx.yb, ctx.act
        Co, Ci, KH, KW = weight.shape
        N, _, Hi, Wi = ctx.shape
        stride = ctx.stride
        H, W = Hi // stride, Wi // stride                     # the output plane: that of dy, y and dz
        pad = (KH // 2, KW // 2) if stride == 1 else (1, 1)
        dev = dy.device
        lib = A.lib()
        with torch.cuda.device(dev):
            st = _stream(dev)
            co_pad = _ru(Co, 8)
            dyb = _to_nhwc(dy.detach().float(), co_pad)
            dz = torch.zeros_like(dyb)
            A.check(lib.demfi_act_backward(dyb.data_ptr(), co_pad, yb.data_ptr(), yb.shape[3], N * H * W, Co, act, dz.data_ptr(),
                                           co_pad, st), "demfi_act_backward")
            dx = dw = db = None
            if ctx.needs_input_grad[0] and stride != 1:
                dxb = torch.empty(N, Hi, Wi, _ru(Ci, 4), dtype=torch.float32, device=dev)
                wc = weight.detach().float().contiguous()
                A.check(lib.demfi_conv2d_dgrad_strided(dz.data_ptr(), co_pad, wc.data_ptr(), Ci, Co, N, H, W, KH, KW, pad[0], pad[1],
                                                       stride, dxb.data_ptr(), dxb.shape[3], st), "demfi_conv2d_dgrad_strided")
                dx = _to_nchw(dxb, Ci)
            elif ctx.needs_input_grad[0]:
                # dx = conv(dz, W^T rotated): w_t[ci, co, ky, kx] = W[co, ci, KH-1-ky, KW-1-kx]; the forward kernel produces at
                # most 256 output channels per launch, so wide inputs (GFF.0: 1152, UNet dec1: 384) go in slices of Cin
                parts = []
                for c0 in range(0, Ci, _MAX_COUT):
                    c1 = min(Ci, c0 + _MAX_COUT)
                    wdev, bdev, ci_pad = _pack_rotated(weight, c0, c1, co_pad, dev)
                    parts.append(_to_nchw(_conv_nhwc(dz, co_pad, wdev, bdev, ci_pad, (KH, KW), A.ACT_NONE), c1 - c0))
                dx = parts[0] if len(parts) == 1 else torch.cat(parts, 1)
            if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
                dw = torch.zeros(Co, Ci, KH, KW, dtype=torch.float32, device=dev)
                db = torch.zeros(Co, dtype=torch.float32, device=dev) if ctx.has_bias else None
                A.check(lib.demfi_conv2d_wgrad(xb.data_ptr(), xb.shape[3], Ci, dz.data_ptr(), co_pad, Co, N, H, W, KH, KW,
                                               pad[0], pad[1], stride, dw.data_ptr(), db.data_ptr() if db is not None else None, st),
                        "demfi_conv2d_wgrad")
        return dx, dw, db, None, None


def conv2d(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor = None, act: str = "none", stride: int = 1) -> torch.Tensor:
    """act(conv2d(x, weight, bias)): stride 1 with 'same' padding (odd KH, KW), or stride 2 with a 4x4 kernel and padding 1 (the
    UNet encoders); x [N,Cin,H,W], weight [Cout,Cin,KH,KW] on a B200."""
    return _Conv2d.apply(x, weight, bias, ACTS[act], stride)


# ---------------------------------------------------------------------------------------------- a first differentiable stack
def decoder_d1(model, feats: torch.Tensor) -> torch.Tensor:
    """D1 of DeMFI-Net (`DeMFInet.py:95-102`: Dec_first -> 5 x ResidualBlock_noBN_3D -> Dec_last1 -> Dec_last2, all
    Conv3d[1,3,3], i.e. the same 2-D stack applied to every frame) as a differentiable function of its input and of the
    module's OWN parameters, every convolution through `conv2d` above.  feats: [F*B, 64, H, W] (the frames batched).
    The first multi-layer slice of the training row: forward + backward + `train.Adam` reproduce torch autograd on the same
    stack (tests/test_grad_gpu.py)."""
    w2 = lambda conv: conv.weight.squeeze(2)                      # [Co,Ci,1,3,3] -> [Co,Ci,3,3] (a view: gradients flow back)
    x = conv2d(feats, w2(model.Dec_first), model.Dec_first.bias, "relu")
    for blk in model.Decoder_res:
        y = conv2d(x, w2(blk.conv1), blk.conv1.bias, "relu")
        x = x + conv2d(y, w2(blk.conv2), blk.conv2.bias, "none")  # ResidualBlock_noBN_3D, DeMFInet.py:517-540
    x = conv2d(x, w2(model.Dec_last1), model.Dec_last1.bias, "relu")
    return conv2d(x, w2(model.Dec_last2), model.Dec_last2.bias, "none")
