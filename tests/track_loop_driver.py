"""Runs in its OWN process (tests/test_dropin_gpu.py::test_reference_tracking_loop): the reference's own tracking loop --
network/models/model.py EvalTrackModel.set_data / forward / compute_loss (model.py:309-593), unmodified from
oracle/_ref/pyref -- on the GPU on top of the drop-in (captra_b200.install_dropin), for a short synthetic trajectory
batch, next to this package's Tracker + frame_ops.track_eval on the same weights and frames.  Prints one JSON line:
per-frame pose differences and the eval / loss means of both sides."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
PYREF = os.path.join(ROOT, "oracle", "_ref", "pyref")
category = sys.argv[1] if len(sys.argv) > 1 else "bottle"
device = sys.argv[2] if len(sys.argv) > 2 else "cuda:0"

import numpy as np  # noqa: E402
import torch  # noqa: E402

import ref_stubs  # noqa: E402

ref_stubs.install()
sys.path[:0] = [os.path.join(PYREF, "network", "models"), os.path.join(PYREF, "pose_utils"), os.path.join(PYREF, "datasets", "nocs_data"),
                os.path.join(PYREF, "datasets"), PYREF]
import captra_b200  # noqa: E402

on_gpu = device.startswith("cuda")
if on_gpu:
    captra_b200.install_dropin(pose=True, mirror_pointnet_lib=False)
import model as RM   # noqa: E402  -- the REFERENCE's network/models/model.py

assert RM.__file__.startswith(PYREF)
from captra_b200 import track   # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
dev = torch.device(device)
B, T, N = 3, 3, 4096
cfg = track.make_cfg(category, device=device)
P = cfg["num_parts"]
cfg.update({"num_joints": P - 1, "loss_weight": {}, "pose_perturb": {"type": "normal", "s": 0.02, "t": 0.03, "r": 5.0},
            "init_frame": {"gt": True}, "data_radius": 0.6, "batch_size": B,
            "track_cfg": {"gt_label": False, "nocs2d_label": False, "nocs2d_path": None}})

# a short trajectory batch in the reference's data format (EvalTrackModel.set_data, model.py:376-384): frame 0 carries
# the initial pose, frames 1.. are tracked
frames, ours_frames = [], []
for t in range(T):
    b = track.synthetic_track_batch(B, category, n=N, seed=20)          # a static scene: the same clouds in every frame, re-tracked from the previous estimate
    gt = {k: np.asarray(v, dtype=np.float32) for k, v in b["gt"].items()}
    labels = np.random.default_rng(t).integers(0, P + cfg["obj"]["extra_dims"], size=(B, N))
    nocs = (np.random.default_rng(100 + t).random((B, 3, N)) - 0.5).astype(np.float32)
    meta = {"nocs2camera": [{k: torch.from_numpy(np.ascontiguousarray(v[:, p])) for k, v in gt.items()} for p in range(P)],
            "points_mean": torch.from_numpy(b["points_mean"]), "path": ["a/b/%d/%04d.pkl" % (i, t) for i in range(B)],
            "nocs_corners": torch.zeros(B, P, 2, 3)}
    frames.append({"points": torch.from_numpy(b["points"]), "labels": torch.from_numpy(labels), "nocs": torch.from_numpy(nocs), "meta": meta})
    ours_frames.append((b, gt, labels, nocs))

ref = RM.EvalTrackModel(cfg)
track.init_weights(ref.npcs_net, 0)
track.init_weights(ref.net, 1)
ref = ref.to(dev).eval()
ref.set_data(frames)
with torch.device(dev):            # torch 2.x: networks.py:127-128 indexes a CPU eye with CUDA labels (torch 1.6 allowed it)
    ref.forward()
    ref.compute_loss(test=True, per_instance=False, eval_iou=False)
ref_poses = ref.pred_dict["poses"]
ref_loss = ref.loss_dict
out = {"category": category, "device": device, "frames": T, "clouds": B,
       "ref_avg_pred": {k: float(v) for k, v in ref_loss["avg_pred"].items()},
       "ref_avg_seg": float(ref_loss["avg_seg"]), "ref_avg_nocs": float(ref_loss["avg_nocs"])}

if on_gpu:
    from captra_b200 import frame_ops
    trk = track.Tracker(cfg, seed=0).to(dev).eval()
    pose = {k: torch.from_numpy(v).to(dev) for k, v in ours_frames[0][1].items()}          # init_frame.gt: the first pose is the ground truth
    sums = torch.zeros(5 * P + 5, device=dev)
    diffs = []
    for t in range(1, T):
        b, gt, labels, nocs = ours_frames[t]
        pose, pred = trk.step(torch.from_numpy(b["points"]).to(dev), torch.from_numpy(b["points_mean"]).to(dev), pose, want_pred=True)
        frame_ops.track_eval({k: torch.from_numpy(v).to(dev) for k, v in gt.items()}, pose, cfg["obj_sym"], pred=pred,
                             gt_labels=torch.from_numpy(labels).to(dev), gt_nocs=torch.from_numpy(nocs).to(dev), out=sums, accumulate=True)
        diffs.append({k: float((pose[k] - ref_poses[t][k]).abs().max()) for k in ("rotation", "translation", "scale")})
    out["pose_max_abs_diff_per_frame"] = diffs
    out["ours"] = frame_ops.eval_means(sums, P)
print(json.dumps(out))
