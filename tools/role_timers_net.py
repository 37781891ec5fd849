"""Role timers (tc_diag & 128: clock64 cycles per warp role, medians over the CTAs) of the convolutions AS THE ENGINE LAUNCHES THEM at
736x1280, N_tst = 3: operands, formats and epilogue plans of the real layers, one line per distinct layer shape.  kclk per tile."""
import ctypes as C, json, os, re, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from demfi_b200 import _abi as A, synth
from demfi_b200.engine import Engine
from tools.role_timers import NAMES

dev = torch.device("cuda:0")
eng = Engine(synth.make_state_dict(0), 1, 736, 1280, dev)
x = synth.make_frames(736, 1280, 0).to(dev)
t = torch.tensor([[0.375]], device=dev)
eng.forward(x, t, 3)
torch.cuda.synchronize()
st = torch.cuda.current_stream(dev).cuda_stream
only = sys.argv[1] if len(sys.argv) > 1 else ""
seen = set()
lib = A.lib()
for ops in (eng.ops_prefix_ff, eng.ops_stage1, eng._iter_ops(0, True)):
    for op in ops:
        if op[0] != "conv" or op[3] not in (A.CONV_TC16, A.CONV_TC16P, A.CONV_TC16W):
            continue
        key = re.sub(r"\.\d+\.", ".N.", op[2])
        if key in seen or (only and not re.search(only, op[2])):
            continue
        seen.add(key)
        d = op[1]
        if d.stride != 1:
            continue
        info = (A.i32 * 16)()
        lib.demfi_conv_describe(C.byref(d), info)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        eng._run_one(op, st)
        ev[0].record()
        eng._run_one(op, st)
        ev[1].record()
        A.set_option("tc_diag", 128)
        eng._run_one(op, st)
        torch.cuda.synchronize()
        A.set_option("tc_diag", 0)
        buf = np.zeros((148, 16), dtype=np.int64)
        A.check(lib.demfi_tc_debug_read(buf.ctypes.data_as(C.POINTER(C.c_int64)), 148), "debug_read")
        pair = op[3] == A.CONV_TC16P
        med = np.median(buf[0::2], axis=0) if pair else np.median(buf, axis=0)
        peer = np.median(buf[1::2], axis=0) if pair else None
        tiles = d.N * ((d.H + 15) // 16) * ((d.W + 7) // 8) * info[6]
        per_cta = tiles / (148.0 if not pair else 148.0)
        row = {"layer": op[2], "k": [d.KH, d.KW], "N": d.N, "HW": [d.H, d.W], "srcC": [d.src[i].C for i in range(d.nsrc)],
               "src_s16": [int(d.src[i].fmt == A.FMT_S16) for i in range(d.nsrc)], "co": d.cout_pad, "pair": int(pair),
               "resident": info[2], "na": info[3], "ns": info[4], "nblk": info[6], "stages": info[9],
               "ms": round(ev[0].elapsed_time(ev[1]), 3), "tiles_per_cta": round(per_cta, 1),
               "kclk_per_tile": {NAMES[i]: round(float(med[i]) / 1e3 / per_cta, 2) for i in NAMES}}
        if pair:
            row["peer_kclk_per_tile"] = {NAMES[i]: round(float(peer[i]) / 1e3 / per_cta, 2) for i in (4, 5, 6, 10, 1, 2)}
        print(json.dumps(row), flush=True)
