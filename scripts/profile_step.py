"""Runs a few tracking steps (cfg2 by default) so ncu can capture individual kernels:
    ncu --set full --clock-control none --import-source on -k regex:mlp_tc_kernel --launch-skip 24 \
        --launch-count 12 -o gpurun_out/prof python scripts/profile_step.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from captra_b200 import track  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "bottle"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
cfg = track.make_cfg(workload)
trk = track.Tracker(cfg).to(dev).eval()
b = track.synthetic_track_batch(B, workload, seed=0)
pts, mean = torch.from_numpy(b["points"]).to(dev), torch.from_numpy(b["points_mean"]).to(dev)
pose = {k: torch.from_numpy(v).to(dev) for k, v in b["pose"].items()}
for _ in range(steps):
    out = trk.step(pts, mean, pose)
torch.cuda.synchronize()
print("ok", float(out["scale"].sum()))
