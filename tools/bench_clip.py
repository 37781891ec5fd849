"""Clip-level throughput of the folder runner: 1280x720 PNG frames in, PNG frames out, x8 MFI, N_tst=3.
Reports interpolated frames/s including decode, H2D, the network (prefix reused across the 7 t of a pair), D2H and encode."""
import json, os, shutil, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cv2, numpy as np, torch
from demfi_b200 import synth
from demfi_b200.DeMFInet import DeMFInet
from demfi_b200.clip import FolderRunner

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 8
root = tempfile.mkdtemp(prefix="demfi_clip_")
os.makedirs(os.path.join(root, "clip"))
base = synth.make_frames(720 + 16, 1280 + 32, seed=1)[0, :, 0]  # [3,H+16,W+32]
for i in range(frames):
    crop = base[:, (i * 2) % 16:(i * 2) % 16 + 720, (i * 4) % 32:(i * 4) % 32 + 1280]
    cv2.imwrite(os.path.join(root, "clip", f"{i:05d}.png"), ((crop.permute(1, 2, 0).numpy() + 1) * 127.5).clip(0, 255).astype(np.uint8))
dev = torch.device("cuda:0")
net = DeMFInet(synth.default_args()).to(dev).eval()
net.load_state_dict(synth.make_state_dict(0), strict=True)
net.final_only = True  # the runner only writes the last iteration's frames
run = FolderRunner(net, multiple=8, num_update=3, io_threads=16)
run.run(root)  # warm-up (engine build, weight packing, page cache)
shutil.rmtree(os.path.join(root, "clip_sharply_interpolated_x8"))
torch.cuda.synchronize()
t0 = time.perf_counter()
stats = run.run(root)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(json.dumps({"clip_frames": frames, "pairs": stats["pairs"], "interpolated": stats["interpolated"], "deblurred": stats["deblurred"],
                  "seconds": round(dt, 3), "interpolated_frames_per_sec_incl_png_io": round(stats["interpolated"] / dt, 2),
                  "output_frames_per_sec": round((stats["interpolated"] + stats["deblurred"]) / dt, 2)}))
shutil.rmtree(root)
