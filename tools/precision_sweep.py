"""End-to-end error of the tcgen05 path vs the committed reference goldens as a function of the
accumulation-segment length (tc_flush), next to the CUDA-core path."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from demfi_b200 import synth, _abi as A
from demfi_b200.engine import Engine
sys.path.insert(0, ROOT)
from oracle import demfi_oracle as O  # checker only

dev = torch.device("cuda:0")
sd = synth.make_state_dict(0)
meta = json.load(open(os.path.join(ROOT, "tests/golden/meta.json")))
for case in ("c32x32_n1_noise", "c64x96_n3", "c256x256_n1"):
    cfg = meta["cases"][case]["cfg"]
    gold = dict(np.load(os.path.join(ROOT, "tests/golden", case + ".npz")))
    x = synth.make_frames(cfg["h"], cfg["w"], 0, cfg["batch"], cfg["smooth"]).to(dev)
    t = torch.tensor(cfg["t"]).reshape(-1, 1).to(dev)
    for kind, flush, mh, comp in (("ffma", 0, 1, 0), ("tc16", 0, 1, 0), ("tc16", 2, 1, 0), ("tc16", 0, 1, 270), ("tc16", 16, 1, 270), ("tc16", 8, 1, 270),
                                  ("tc16", 4, 1, 270), ("tc16", 2, 1, 270), ("tc16", 2, 1, 300), ("tc16", 1, 1, 300), ("tc16", 8, 1, 240)):
        A.set_option("tc_flush", flush)
        A.set_option("tc_mask_hi", mh)
        A.set_option("tc_comp_milli", comp)
        eng = Engine(sd, cfg["batch"], cfg["h"], cfg["w"], dev, conv_kind=kind)
        got = O.flatten_outputs(eng.forward(x, t, cfg["n"]))
        torch.cuda.synchronize()
        errs = {k: float((got[k].cpu() - torch.from_numpy(g)).abs().max()) for k, g in gold.items() if k in got}
        fr = {k: float(((got[k].cpu() - torch.from_numpy(g)).abs() > 5e-4).float().mean()) for k, g in gold.items() if k in got}
        print(json.dumps({"case": case, "kind": kind, "flush": flush, "rn_split": mh, "comp": comp, "worst": max(errs.values()),
                          "worst_S": max(v for k, v in errs.items() if k.startswith("S")),
                          "worst_flow": max(v for k, v in errs.items() if k.startswith("flow")),
                          "frac_gt_5e-4": max(fr.values())}), flush=True)
        del eng
A.set_option("tc_flush", 10)
A.set_option("tc_mask_hi", 1)
A.set_option("tc_comp_milli", 270)
