"""Small invocations of the round-2 kernels for compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool memcheck  python scripts/sanitize_ops.py
    compute-sanitizer --tool racecheck python scripts/sanitize_ops.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from captra_b200 import data_crop, frame_ops, fused_ops, synthetic, track  # noqa: E402
from captra_b200 import pointnet2_cuda as P  # noqa: E402

dev = torch.device("cuda:0")
gen = torch.Generator().manual_seed(0)
x = torch.from_numpy(synthetic.batch_surface_box(3, 1000, seed=1)[0]).to(dev)
new_xyz, idxs = fused_ops.fps_ball_query(x, 64, [0.1, 0.3], [8, 16])
big = torch.from_numpy(synthetic.batch_uniform(2, 9000, seed=2)).to(dev)
fused_ops.fps_gather(big, 40)                                            # cluster variant
fused_ops.ball_query_group(0.08, 16, big, big[:, :100].contiguous(), big.transpose(1, 2).contiguous())
B, Pn, N = 3, 2, 500
pose = {"rotation": torch.eye(3).repeat(B, Pn, 1, 1).to(dev), "translation": torch.randn(B, Pn, 3, 1, generator=gen).to(dev),
        "scale": (torch.rand(B, Pn, generator=gen) + 0.2).to(dev)}
pts, mean = (torch.randn(B, 3, N, generator=gen) * 0.2).to(dev), torch.randn(B, 3, 1, generator=gen).to(dev)
frame_ops.canonicalize(pts, mean, pose["rotation"].reshape(-1, 3, 3), pose["translation"].reshape(-1, 3, 1), pose["scale"].reshape(-1), parts=Pn,
                       want_cm=True, want_dup=True)
labels, nocs, seg = frame_ops.coord_head_post(torch.randn(B * N, 3, generator=gen).to(dev), torch.randn(B * N, 6, generator=gen).to(dev), B, N)
rot = frame_ops.rot_head_post([torch.randn(B, N, 6, generator=gen).to(dev) for _ in range(Pn)], labels, pose["rotation"], False)
frame_ops.part_fit_track(labels, nocs.reshape(B, Pn, 3, N), pts, mean, rot, False, pose["scale"], pose["translation"])
frame_ops.track_eval(pose, {"rotation": rot, "translation": pose["translation"] + 0.01, "scale": pose["scale"]}, False,
                     pred={"seg": seg, "nocs": nocs, "labels": labels}, gt_labels=labels, gt_nocs=nocs[:, :3].contiguous())
depth, mask, c, K = synthetic.depth_scene(seed=3, obj_radius=0.05, obj_depth=1.2, height=240, width=320,
                                          intrinsics=((295.5, 0, 161.2), (0, 295.1, 122.0), (0, 0, 1)))
data_crop.crop_ball_from_depth_image(torch.from_numpy(depth).to(dev), torch.from_numpy(mask).to(dev), c + np.array([0, 0, 0.02]), 0.06,
                                     cam_intrinsics=K, num_points=256)
g = torch.randn(2, 8, 32, 4, generator=gen).to(dev)
gi = torch.randint(0, 50, (2, 32, 4), generator=gen, dtype=torch.int32).to(dev)
out = torch.zeros(2, 8, 50, device=dev)
P.group_points_grad_wrapper(2, 8, 50, 32, 4, g, gi, out)
torch.cuda.synchronize()
print("sanitize workload done")
