"""Timing of the HBM-bound operators at the north-star resolution (736x1280) through the C ABI, CUDA events, inputs
larger than L2 rotated between iterations.  Reports algorithmic GB/s (every operand once)."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from demfi_b200 import _abi as A

DEV = torch.device("cuda:0")
H, W = 736, 1280


def timeit(fn, iters=6, warm=2):
    for _ in range(warm):
        fn()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    return min(ev[i].elapsed_time(ev[i + 1]) for i in range(iters))


def smooth_flow(n, h, w, ch, amp=3.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    low = torch.randn(n, ch, h // 32 + 2, w // 32 + 2, generator=g) * amp
    f = torch.nn.functional.interpolate(low, size=(h, w), mode="bilinear", align_corners=False)
    return f.permute(0, 2, 3, 1).contiguous().to(DEV)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    lib = A.lib()
    st = torch.cuda.current_stream(DEV).cuda_stream
    npix = H * W
    t = torch.tensor([0.375], device=DEV)
    rows = []

    def report(name, ms, nbytes):
        rows.append({"op": name, "ms": round(ms, 4), "GB/s": round(nbytes / ms / 1e6, 1), "algorithmic_MB": round(nbytes / 1e6, 1)})
        print(json.dumps(rows[-1]), flush=True)

    if a.only in ("", "bwarp64"):
        fa, fb = torch.randn(1, H, W, 64, device=DEV), torch.randn(1, H, W, 64, device=DEV)
        flow, occ = smooth_flow(1, H, W, 4), torch.randn(1, H, W, 4, device=DEV)
        out = torch.empty(1, H, W, 64, device=DEV)
        fn = lambda: A.check(lib.demfi_bwarp_blend(fa.data_ptr(), 64, fb.data_ptr(), 64, flow.data_ptr(), 4, occ.data_ptr(), 4, t.data_ptr(),
                                                   1, H, W, 64, out.data_ptr(), 64, None, 0, st), "bwarp")
        report("bwarp_blend C=64", timeit(fn), npix * 4 * (64 + 64 + 4 + 1 + 64))
    if a.only in ("", "bwarp3"):
        buf = torch.randn(1, H, W, 36, device=DEV)
        flow, occ = smooth_flow(1, H, W, 8), None
        fn = lambda: A.check(lib.demfi_bwarp_blend(buf.data_ptr(), 36, buf.data_ptr() + 12, 36, flow.data_ptr(), 8, flow.data_ptr() + 16, 8, t.data_ptr(),
                                                   1, H, W, 3, buf.data_ptr() + 24, 36, buf.data_ptr() + 88, 36, st), "pwb")
        report("bwarp_blend C=3 (PWB)", timeit(fn), npix * 4 * (3 + 3 + 4 + 1 + 3 + 1))
    if a.only in ("", "fgac"):
        rk = torch.randn(1, H, W, 64, device=DEV)
        flow = smooth_flow(1, H, W, 8)
        out = torch.empty(1, H, W, 64, device=DEV)
        fn = lambda: A.check(lib.demfi_fgac_sample(rk.data_ptr(), 64, flow.data_ptr(), 8, 1, H, W, 64, out.data_ptr(), 64, st), "fgac_sample")
        report("fgac_sample C=64", timeit(fn), npix * 4 * (64 + 2 + 64))
        se = torch.randn(1, H, W, 128, device=DEV)
        wl = torch.rand(1, H, W, 4, device=DEV)
        agg = torch.empty(1, H, W, 204, device=DEV)
        fn = lambda: A.check(lib.demfi_fgac_blend(wl.data_ptr(), 4, se.data_ptr(), 128, se.data_ptr() + 256, 128, npix, 64, agg.data_ptr(), 204, st), "fgac_blend")
        report("fgac_blend C=64", timeit(fn), npix * 4 * (1 + 64 + 64 + 64))
    if a.only in ("", "cfr"):
        fo = smooth_flow(1, H, W, 8)
        acc = torch.zeros(1, H, W, 8, device=DEV)
        out = torch.empty(1, H, W, 204, device=DEV)
        def fn():
            acc.zero_()
            A.check(lib.demfi_cfr_splat(fo.data_ptr(), 8, t.data_ptr(), 1, H, W, acc.data_ptr(), st), "splat")
            A.check(lib.demfi_cfr_finalize(acc.data_ptr(), t.data_ptr(), 1, H, W, out.data_ptr() + 4 * 192, 204, st), "finalize")
        report("cfr zero+splat+finalize", timeit(fn), npix * 4 * (8 + 4 + 12 + 8 + 4))
    if a.only in ("", "copy"):
        src = torch.randn(1, H, W, 8, device=DEV)
        dst = torch.empty(1, H, W, 36, device=DEV)
        fn = lambda: A.check(lib.demfi_copy_channels(src.data_ptr(), 8, dst.data_ptr() + 40, 36, 4, npix, 0, st), "copy")
        report("copy_channels 4ch (ld 8 -> ld 36)", timeit(fn), npix * 4 * 8)
        sp = torch.randn(3, H, W, 4, device=DEV)
        o = torch.empty(3, 3, H, W, device=DEV)
        fn = lambda: A.check(lib.demfi_export_nchw(sp.data_ptr(), 4, 3, H, W, 3, 0, o.data_ptr(), st), "export")
        report("export_nchw 3ch x3 frames", timeit(fn), 3 * npix * 4 * 6)
        x = torch.randn(1, 3, 4, H, W, device=DEV)
        s2d = torch.empty(1, H // 2, W // 2, 48, device=DEV)
        r = torch.empty(1, H, W, 32, device=DEV)
        a3 = torch.empty(1, H, W, 36, device=DEV)
        m = torch.empty(1, 3, H, W, device=DEV)
        fn = lambda: A.check(lib.demfi_pack_input(x.data_ptr(), 1, H, W, s2d.data_ptr(), r.data_ptr() + 36, 32, a3.data_ptr() + 92, 36, m.data_ptr(), st), "pack")
        report("pack_input", timeit(fn), npix * 4 * (12 + 12 * 3 + 3))


if __name__ == "__main__":
    main()
