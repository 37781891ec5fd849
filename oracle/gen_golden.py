"""ORACLE TOOLING -- runs only in the build container (needs /root/reference).

Generates the golden vectors under tests/golden/ by executing the UNMODIFIED reference
module `/root/reference/DeMFInet.py` (class `DeMFInet`, `DeMFInet.py:13-179`) on the seeded
synthetic inputs / weights of `demfi_b200/synth.py`.  The reference repo ships no fixtures
of its own, so these files are what pins `oracle/demfi_oracle.py`.

    python oracle/gen_golden.py            # rewrites tests/golden/*.npz and *.json

The GPU box has no /root/reference: tests only ever read the committed files.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from demfi_b200 import synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

# name -> (H, W, N, batch, t list, smooth inputs, which outputs to keep (None = all))
CASES = {
    "c64x96_n3": dict(h=64, w=96, n=3, batch=1, t=[0.375], smooth=True, keep=None, inter=True),
    "c48x40_n2_b2": dict(h=48, w=40, n=2, batch=2, t=[0.25, 0.875], smooth=True, keep=None, inter=False),
    "c32x32_n1_noise": dict(h=32, w=32, n=1, batch=1, t=[0.5], smooth=False, keep=None, inter=False),
    "c256x256_n1": dict(h=256, w=256, n=1, batch=1, t=[0.5], smooth=True,
                        keep=["Stp", "St_final0", "flow1", "occ1"], inter=False),
}


def load_reference():
    sys.path.insert(0, REF)
    import DeMFInet as ref_mod  # the reference file, unmodified
    sys.path.remove(REF)
    return ref_mod


def name_outputs(res):
    s1, sf, fl, oc, tb = res[:5]
    d = {"S0p": s1[0], "S1p": s1[1], "Stp": s1[2], "two_blurry": tb}
    for i, tri in enumerate(sf):
        d[f"S0_final{i}"], d[f"S1_final{i}"], d[f"St_final{i}"] = tri
    for i, f in enumerate(fl):
        d[f"flow{i}"] = f
    for i, o in enumerate(oc):
        d[f"occ{i}"] = o
    return d


def run_case(net, cfg, hooks=False):
    x = synth.make_frames(cfg["h"], cfg["w"], seed=0, batch=cfg["batch"], smooth=cfg["smooth"])
    t = torch.tensor(cfg["t"], dtype=torch.float32).reshape(cfg["batch"], 1)
    inter = {}
    handles = []
    if hooks:
        def grab(name, sl=None):
            def fn(_m, _i, o):
                o = o[0] if isinstance(o, (tuple, list)) else o
                inter[name] = o.detach().clone()
            return fn
        handles.append(net.FF_RDB_Module.register_forward_hook(
            lambda m, i, o: inter.update(flow_01=o[2].clone(), flow_10=o[3].clone(),
                                         occ_logit_ff=o[4].clone(), F0_c8=o[0][:, ::8].clone())))
        handles.append(net.FAC_FB_Module.register_forward_hook(
            lambda m, i, o: inter.update(aF0_c8=o[0][:, ::8].clone(), aF1_c8=o[1][:, ::8].clone())))
        handles.append(net.Ch_Reducer.register_forward_hook(
            lambda m, i, o: inter.update(F_rec0_c8=torch.tanh(o)[:, ::8].clone())))
    with torch.no_grad():
        res = net(x, t, cfg["n"])
    for h in handles:
        h.remove()
    d = name_outputs(res)
    d.update(inter)
    return {k: v.detach().numpy().astype(np.float32) for k, v in d.items()}


def main():
    os.makedirs(GOLD, exist_ok=True)
    ref_mod = load_reference()
    args = synth.default_args()
    torch.manual_seed(0)
    net = ref_mod.DeMFInet(args).eval()
    ref_sd = net.state_dict()
    with open(os.path.join(GOLD, "state_dict_keys.json"), "w") as f:
        json.dump({k: list(v.shape) for k, v in ref_sd.items()}, f, indent=0)
    sd = synth.make_state_dict(seed=0)
    net.load_state_dict(sd, strict=True)

    meta = {"torch": torch.__version__, "threads": torch.get_num_threads(), "cases": {}}
    for name, cfg in CASES.items():
        out = run_case(net, cfg, hooks=cfg["inter"])
        if cfg["keep"] is not None:
            out = {k: v for k, v in out.items() if k in cfg["keep"]}
        np.savez(os.path.join(GOLD, name + ".npz"), **out)
        # the reference's own noise floor: same forward with 1 thread
        nt = torch.get_num_threads()
        torch.set_num_threads(1)
        out1 = run_case(net, cfg, hooks=False)
        torch.set_num_threads(nt)
        noise = {k: float(np.abs(out1[k] - v).max()) for k, v in out.items() if k in out1}
        meta["cases"][name] = {"cfg": {k: v for k, v in cfg.items() if k != "keep"},
                               "self_noise_max_abs_1_vs_%d_threads" % nt: noise,
                               "max_abs_flow": float(max(np.abs(v).max() for k, v in out.items() if k.startswith("flow")))}
        print(name, "saved; self-noise max", max(noise.values()), "max|flow|", meta["cases"][name]["max_abs_flow"])
    with open(os.path.join(GOLD, "meta.json"), "w") as f:
        json.dump(meta, f, indent=1)


if __name__ == "__main__":
    main()
