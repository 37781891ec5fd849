"""Race hunt: run the same tcgen05 conv many times and compare every output bit-for-bit with the first."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from demfi_b200 import _abi as A
from tools.bench_conv import make_conv

def run(tag, iters, **opts):
    for k, v in opts.items():
        A.set_option(k, v)
    d, keep = make_conv(A.CONV_TC, 1, 736, 1280, [64], 64, (3, 3))
    out = keep[-1]
    lib = A.lib()
    st = torch.cuda.current_stream().cuda_stream
    A.check(lib.demfi_conv2d(C.byref(d), st), "conv")
    torch.cuda.synchronize()
    ref = out.clone()
    bad = 0
    for i in range(iters):
        out.zero_()
        A.check(lib.demfi_conv2d(C.byref(d), st), "conv")
        if i % 10 == 9:
            torch.cuda.synchronize()
            if not torch.equal(out, ref):
                bad += 1
    torch.cuda.synchronize()
    print(json.dumps({"cfg": tag, "iters": iters, "mismatching_checks": bad, "max_abs": float((out - ref).abs().max())}), flush=True)

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 300
run("ts flush2", iters, tc_a_tmem=1, tc_flush=2)
run("ts flush0", iters, tc_a_tmem=1, tc_flush=0)
run("ss flush2", iters, tc_a_tmem=0, tc_flush=2)
run("ss flush0", iters, tc_a_tmem=0, tc_flush=0)
