"""CUDA-event timing of the backward kernels of the training slice (BASELINE config 4 shapes: 256x256 crops, 2 per GPU):
the CUDA-core wgrad, the dx convolution on the tcgen05 kernel, and the warp / splat backward operators."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from demfi_b200 import _abi as A, grad

DEV = torch.device("cuda:0")


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    lib = A.lib()
    st = torch.cuda.current_stream(DEV).cuda_stream
    N, H, W = 2, 256, 256
    for ci, co, k in ((64, 64, (3, 3)), (192, 64, (7, 7)), (128, 128, (1, 5)), (192, 32, (3, 3))):
        x = torch.randn(N, H, W, ci, device=DEV)
        dz = torch.randn(N, H, W, co, device=DEV)
        dw = torch.zeros(co, ci, *k, device=DEV)
        db = torch.zeros(co, device=DEV)
        wg = lambda: A.check(lib.demfi_conv2d_wgrad(x.data_ptr(), ci, ci, dz.data_ptr(), co, co, N, H, W, k[0], k[1], k[0] // 2,
                                                    k[1] // 2, 1, dw.data_ptr(), db.data_ptr(), st), "wgrad")
        A.set_option("wgrad_kind", 0)
        ms = timed(wg)
        A.set_option("wgrad_kind", 1)
        ms_mma = timed(wg)
        flop = 2.0 * N * H * W * ci * co * k[0] * k[1]
        # dx through the forward tensor-core kernel
        wt = torch.randn(ci, co, *k) / (co * k[0] * k[1]) ** 0.5           # already "transposed": maps co -> ci
        wdev, bdev, cpad = grad._pack(wt.numpy(), torch.zeros(ci).numpy(), co, DEV)
        ms_dx = timed(lambda: grad._conv_nhwc(dz, co, wdev, bdev, cpad, k, A.ACT_NONE))
        print(json.dumps({"layer": f"{ci}->{co} {k[0]}x{k[1]} @ {N}x{H}x{W}", "GFLOP": round(flop / 1e9, 2),
                          "wgrad_cuda_core_ms": round(ms, 3), "wgrad_cuda_core_TFLOPs": round(flop / ms / 1e9, 2),
                          "wgrad_mma_3xtf32_ms": round(ms_mma, 3), "wgrad_mma_TFLOPs": round(flop / ms_mma / 1e9, 2),
                          "dx_tcgen05_ms": round(ms_dx, 3), "dx_TFLOPs": round(flop / ms_dx / 1e9, 1)}), flush=True)
    # warp / splat backward (HBM- and atomic-bound): algorithmic bytes = a, b, dout read + da, db read-modify-write
    C = 64
    a, b, g = (torch.randn(N, H, W, C, device=DEV) for _ in range(3))
    fl = torch.randn(N, H, W, 4, device=DEV) * 2
    occ = torch.randn(N, H, W, 1, device=DEV)
    t = torch.tensor([0.3, 0.7], device=DEV)
    da, db_ = torch.zeros_like(a), torch.zeros_like(b)
    dfl, doc = torch.zeros(N, H, W, 4, device=DEV), torch.zeros(N, H, W, 1, device=DEV)
    ms = timed(lambda: A.check(lib.demfi_bwarp_blend_backward(a.data_ptr(), C, b.data_ptr(), C, fl.data_ptr(), 4, occ.data_ptr(), 1, t.data_ptr(),
                                                              g.data_ptr(), C, N, H, W, C, da.data_ptr(), C, db_.data_ptr(), C, dfl.data_ptr(), 4,
                                                              doc.data_ptr(), 1, st), "bwarp_bwd"))
    byt = N * H * W * (3 * C + 2 * 2 * C + 10) * 4
    print(json.dumps({"op": "bwarp_blend_backward C=64", "ms": round(ms, 4), "algorithmic_GBps": round(byt / ms / 1e6, 1)}), flush=True)
    acc = torch.zeros(N, H, W, 8, device=DEV)
    fo = torch.randn(N, H, W, 8, device=DEV) * 2
    A.check(lib.demfi_cfr_splat(fo.data_ptr(), 8, t.data_ptr(), N, H, W, acc.data_ptr(), st), "splat")
    gacc, dfo = torch.zeros(N, H, W, 8, device=DEV), torch.zeros(N, H, W, 4, device=DEV)
    ms = timed(lambda: A.check(lib.demfi_cfr_backward(fo.data_ptr(), 8, t.data_ptr(), acc.data_ptr(), dfl.data_ptr(), 4, N, H, W,
                                                      gacc.data_ptr(), dfo.data_ptr(), 4, st), "cfr_bwd"))
    print(json.dumps({"op": "cfr_backward", "ms": round(ms, 4), "algorithmic_GBps": round(N * H * W * (4 + 8 + 4 + 8 + 8 * 4 * 2 + 4) * 4 / ms / 1e6, 1)}), flush=True)


if __name__ == "__main__":
    main()
