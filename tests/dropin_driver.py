"""Runs in its OWN process (tests/test_dropin_gpu.py): the reference's own Python for the hot path, copied unmodified
to oracle/_ref/pyref by oracle/Makefile, executed on the GPU on top of the drop-in.  Prints one JSON line.

  mode a   INTEGRATION route A: `import pointnet2_cuda` -> captra_b200.pointnet2_cuda (ctypes -> C ABI)
  mode b   INTEGRATION route B: `import pointnet2_cuda` -> oracle/_ref/routeb/pointnet2_cuda.so, the reference's own
           C++ wrappers + pybind module linked against libcaptra_ops.so
Checks: (1) the reference's pointnet2_utils.py Functions (furthest_point_sample, ball_query, grouping_operation,
gather_operation, three_nn, three_interpolate) against the CPU oracle, bit-exact; (2) the reference's own CoordNet and
PartCanonNet (networks.py) on CUDA tensors -- their pointnet_utils.py takes its CUDA branch, and (mode a) pose_utils
resolves to the device pose fit -- against this package's fused Tracker.step on the same weights and inputs.
"""
import importlib.util
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PYREF = os.path.join(ROOT, "oracle", "_ref", "pyref")
mode = sys.argv[1]

import numpy as np  # noqa: E402
import torch  # noqa: E402

sys.path[:0] = [os.path.join(PYREF, "network", "models"), os.path.join(PYREF, "pose_utils"), PYREF]
import captra_b200  # noqa: E402

if mode == "b":
    spec = importlib.util.spec_from_file_location("pointnet2_cuda", os.path.join(ROOT, "oracle", "_ref", "routeb", "pointnet2_cuda.so"))
    ext = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ext)
    sys.modules["pointnet2_cuda"] = ext
else:
    captra_b200.install_dropin(pose=True, mirror_pointnet_lib=False)

from pointnet_lib import pointnet2_utils as futils   # noqa: E402  -- the REFERENCE's file
import networks as RN   # noqa: E402  -- the REFERENCE's networks.py
import pointnet_utils as RPU   # noqa: E402

assert futils.__file__.startswith(PYREF) and RN.__file__.startswith(PYREF) and RPU.CUDA
from captra_b200 import synthetic, track   # noqa: E402
from oracle import cpu_ref   # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
dev = torch.device("cuda:0")
out = {"mode": mode, "pointnet2_cuda": sys.modules["pointnet2_cuda"].__name__ + " @ " + getattr(sys.modules["pointnet2_cuda"], "__file__", "?")}

# ---- (1) the reference's autograd Functions over the drop-in, vs the CPU oracle
pts = synthetic.batch_surface_box(2, 4096, seed=3)[0]
x = torch.from_numpy(pts).to(dev)
idx = futils.furthest_point_sample(x, 512)
out["fps_exact"] = bool(np.array_equal(idx.cpu().numpy(), cpu_ref.furthest_point_sample(pts, 512)))
ctr = futils.gather_operation(x.transpose(1, 2).contiguous(), idx).transpose(1, 2).contiguous()
out["gather_exact"] = bool(np.array_equal(ctr.cpu().numpy(), np.take_along_axis(pts, idx.cpu().numpy().astype(np.int64)[..., None], 1)))
bq = futils.ball_query(0.1, 64, x, ctr)
want_bq = cpu_ref.ball_query(0.1, 64, pts, ctr.cpu().numpy())
out["ball_query_exact"] = bool(np.array_equal(bq.cpu().numpy(), want_bq))
feats = torch.randn(2, 16, 4096, device=dev)
g = futils.grouping_operation(feats, bq)
out["group_exact"] = bool(np.array_equal(g.cpu().numpy(), cpu_ref.grouping_operation(feats.cpu().numpy(), want_bq)))
d, i3 = futils.three_nn(x, ctr)
wd, wi = cpu_ref.three_nn(pts, ctr.cpu().numpy())
out["three_nn_exact"] = bool(np.array_equal(i3.cpu().numpy(), wi) and np.array_equal(d.cpu().numpy(), wd))
w = cpu_ref.interp_weights(wd)
f2 = torch.randn(2, 16, 512, device=dev)
it = futils.three_interpolate(f2, i3, torch.from_numpy(w).to(dev))
out["three_interpolate_exact"] = bool(np.array_equal(it.cpu().numpy(), cpu_ref.three_interpolate(f2.cpu().numpy(), wi, w)))

# ---- (2) the reference's own networks on the GPU over the drop-in vs the fused Tracker
res = {}
for category in ("bottle", "laptop"):
    cfg = track.make_cfg(category, device=str(dev))
    P = cfg["num_parts"]
    trk = track.Tracker(cfg, seed=0).to(dev).eval()
    npcs_net = track.init_weights(RN.CoordNet(cfg), 0).to(dev).eval()
    net = track.init_weights(RN.PartCanonNet(cfg), 1).to(dev).eval()
    b = track.synthetic_track_batch(2, category, n=4096, seed=0)
    p, m = torch.from_numpy(b["points"]).to(dev), torch.from_numpy(b["points_mean"]).to(dev)
    pose = {k: torch.from_numpy(v).to(dev) for k, v in b["pose"].items()}
    # torch 1.6 (the reference's pin) let a CPU tensor be indexed by CUDA indices (networks.py:127-128 builds `eye_mat`
    # on the CPU and indexes it with CUDA labels); torch 2.x does not, so factory calls default to the GPU here --
    # an environment setting, the reference's code is untouched
    with torch.no_grad(), torch.device(dev):          # model.py:454-476 with the reference's modules
        canon = {k: pose[k][:, trk.root] for k in ("rotation", "translation", "scale")}
        pred = npcs_net({"points": p, "points_mean": m, "canon_pose": canon})
        labels = torch.max(pred["seg"], dim=-2)[1]
        ref = net({"points": p, "points_mean": m, "state": {"part": pose}, "pred_labels": labels,
                   "pred_nocs": pred["nocs"].reshape(2, P, 3, -1)}, test_mode=True)["part"]
    ours, opred = trk.step(p, m, pose, want_pred=True)
    res[category] = {
        "labels_equal": bool((labels == opred["labels"]).all()),
        "nocs_max_abs": float((pred["nocs"] - opred["nocs"]).abs().max()),
        "rotation_max_abs": float((ref["rotation"] - ours["rotation"]).abs().max()),
        "scale_max_abs": float((ref["scale"] - ours["scale"]).abs().max()),
        "translation_max_abs": float((ref["translation"] - ours["translation"]).abs().max()),
        "procrustes_module": sys.modules["pose_utils.procrustes"].__name__ if "pose_utils.procrustes" in sys.modules else None,
    }
out["frame"] = res

# ---- (3) training drop-in: the reference's CoordNet in TRAINING mode (BatchNorm batch statistics, autograd through the
# drop-in ops and the differentiable scale / translation fit, networks.py:54-108) vs this package's mirror in training mode
from captra_b200 import networks as ON   # noqa: E402
tr = {}
for category in ("bottle", "laptop"):
    cfg = track.make_cfg(category, device=str(dev))
    P = cfg["num_parts"]
    ref_net = track.init_weights(RN.CoordNet(cfg), 5).to(dev).train()
    our_net = track.init_weights(ON.CoordNet(cfg), 5).to(dev).train()
    b = track.synthetic_track_batch(2, category, n=1024, seed=3)
    gen = torch.Generator().manual_seed(1)
    inp = {"points": torch.from_numpy(b["points"]).to(dev), "points_mean": torch.from_numpy(b["points_mean"]).to(dev),
           "labels": torch.randint(0, P + cfg["obj"]["extra_dims"], (2, 1024), generator=gen).to(dev),
           "gt_part": {k: torch.from_numpy(np.asarray(v, dtype=np.float32)).to(dev) for k, v in b["gt"].items()},
           "init_part": {k: torch.from_numpy(v).to(dev) for k, v in b["pose"].items()}}
    inp["canon_pose"] = {k: inp["init_part"][k][:, 0] for k in ("rotation", "translation", "scale")}
    grads = []
    outs = []
    for net in (ref_net, our_net, ref_net):      # the reference twice: its own run-to-run spread (cuDNN algorithm choice, atomics)
        with torch.device(dev):
            pred = net(dict(inp))
        loss = (pred["nocs"] ** 2).mean() + pred["seg"][:, 0].mean() + pred["part"]["scale"].sum() + pred["part"]["translation"].abs().sum()
        net.zero_grad()
        loss.backward()
        grads.append({k: p.grad.detach().clone() for k, p in net.named_parameters() if p.grad is not None})
        outs.append({"loss": float(loss), "scale": pred["part"]["scale"].detach(), "translation": pred["part"]["translation"].detach(),
                     "nocs": pred["nocs"].detach()})
    assert grads[0].keys() == grads[1].keys() and len(grads[0]) > 50
    # relative to the largest gradient of the model: a conv bias in front of a training-mode BatchNorm has an analytically
    # zero gradient (pure rounding noise on both sides), so a per-tensor ratio would be meaningless there
    gmax = max(float(g.abs().max()) for g in grads[0].values())
    rel = max(float((grads[0][k] - grads[1][k]).abs().max()) for k in grads[0]) / gmax
    self_rel = max(float((grads[0][k] - grads[2][k]).abs().max()) for k in grads[0]) / gmax
    tr[category] = {"loss_ref": outs[0]["loss"], "loss_ours": outs[1]["loss"], "params_with_grad": len(grads[0]),
                    "nocs_max_abs": float((outs[0]["nocs"] - outs[1]["nocs"]).abs().max()),
                    "scale_max_abs": float((outs[0]["scale"] - outs[1]["scale"]).abs().max()),
                    "translation_max_abs": float((outs[0]["translation"] - outs[1]["translation"]).abs().max()),
                    "grad_max_rel": rel, "grad_ref_vs_ref_rel": self_rel,
                    "nocs_ref_vs_ref_max_abs": float((outs[0]["nocs"] - outs[2]["nocs"]).abs().max())}
out["train"] = tr
print(json.dumps(out))
