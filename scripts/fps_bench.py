"""FPS kernel timing at the shapes that matter (cfg2 sa1 / sa2, cfg5 level 1, the crop's 20480 -> 4096), CUDA events,
median of 5.  CAPTRA_FPS_CLUSTER=0 -> streaming kernel above 8192 points; CAPTRA_FPS_SYNC=0 -> every warp fences."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from captra_b200 import fused_ops, synthetic  # noqa: E402

dev = torch.device("cuda:0")
out = {"env": {k: os.environ.get(k) for k in ("CAPTRA_FPS_CLUSTER", "CAPTRA_FPS_SYNC")}}
for B, N, M in ((32, 4096, 512), (32, 512, 128), (1, 20480, 4096), (64, 16384, 4096), (8, 16384, 4096), (1, 8192, 4096), (1, 32768, 4096)):
    x = torch.from_numpy(synthetic.batch_uniform(B, N, seed=1)).to(dev)
    fused_ops.fps_gather(x, M)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fused_ops.fps_gather(x, M)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    us = float(np.median(ts))
    out["fps[B=%d,%d->%d]" % (B, N, M)] = {"us": round(us, 1), "us_per_round": round(us / (M - 1), 3)}
print(json.dumps(out))
