"""Workload for the per-op ncu captures (profiles/r02_ncu_ops_*): every named kernel of the hot path once, at the shapes
bench.py times, between cudaProfilerStart/Stop (run under `ncu --profile-from-start off`).

    ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/r02_ops python scripts/ncu_ops.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from captra_b200 import _lib, data_crop, fused_ops, synthetic, track  # noqa: E402
from captra_b200.pointnet_lib import pointnet2_utils as futils  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
which = sys.argv[1] if len(sys.argv) > 1 else "all"

trk = track.Tracker(track.make_cfg("bottle", str(dev)), seed=0).to(dev).eval()
b = track.synthetic_track_batch(32, "bottle", seed=0)
pts, mean = torch.from_numpy(b["points"]).to(dev), torch.from_numpy(b["points_mean"]).to(dev)
pose = {k: torch.from_numpy(v).to(dev) for k, v in b["pose"].items()}
gt = {k: torch.from_numpy(np.asarray(v, dtype=np.float32)).to(dev) for k, v in b["gt"].items()}
for _ in range(2):
    new = trk.step(pts, mean, pose)
    trk.eval_sums(gt, new)

x = torch.from_numpy(synthetic.batch_surface_box(32, 4096, seed=0)[0]).to(dev)
idx = futils.furthest_point_sample(x, 512)
ctr = torch.gather(x, 1, idx.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()
feats = torch.randn(32, 128, 4096, device=dev)
gidx = futils.ball_query(0.2, 128, x, ctr)
big = torch.from_numpy(np.stack([synthetic.surface_box(16384, np.random.default_rng(i))[0] for i in range(64)])).to(dev)
_, bctr = fused_ops.fps_gather(big, 4096)
M3 = torch.randn(1 << 20, 3, 3, device=dev)
R3 = torch.empty_like(M3)
depth, mask, c, K = synthetic.depth_scene(seed=1, obj_radius=0.25, obj_depth=0.6)
d_t, m_t = torch.from_numpy(depth).to(dev), torch.from_numpy(mask).to(dev)
g = torch.randn(32, 128, 512, 16, device=dev)
gi = gidx[:, :, :16].contiguous()


def standalone():
    futils.grouping_operation(feats, gidx)
    futils.gather_operation(feats, idx)
    d, i3 = futils.three_nn(x, ctr)
    futils.three_interpolate(torch.randn(32, 128, 512, device=dev), i3, torch.softmax(-d, -1).contiguous())
    for r, k in ((0.05, 32), (0.2, 128)):
        futils.ball_query(r, k, x, ctr)
    fused_ops.ball_query_group(0.05, 64, big, bctr, big.transpose(1, 2).contiguous())
    fused_ops.fps_gather(big[:8], 1024)                      # cluster variant (16384 points)
    _lib.call("rot3", _lib.load().captra_procrustes_rot3, M3.shape[0], M3.data_ptr(), R3.data_ptr(), _lib.stream_ptr(dev), device=dev)
    data_crop.crop_ball_from_depth_image(d_t, m_t, c + np.array([0, 0, 0.1]), 0.3, cam_intrinsics=K, num_points=1024)
    out = torch.zeros(32, 128, 4096, device=dev)
    from captra_b200 import pointnet2_cuda as P
    P.group_points_grad_wrapper(32, 128, 4096, 512, 16, g, gi, out)


standalone()
torch.cuda.synchronize()
torch.cuda.profiler.start()
if which in ("all", "frame"):
    new = trk.step(pts, mean, pose)
    trk.eval_sums(gt, new)
if which in ("all", "ops"):
    standalone()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
