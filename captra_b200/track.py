"""Per-frame tracking step (the loop body of EvalTrackModel.forward, model.py:409-478) on top of
the B200 kernels, plus the synthetic workloads of BASELINE.json used by bench.py and the tests.

One "frame" = CoordNet forward on B clouds + argmax labels + PartCanonNet forward on B*P
canonicalised clouds + fused pose fit -> the new 9-DoF part poses for B trajectories.
"""
import os

import numpy as np
import torch

from . import synthetic
from .networks import CoordNet, PartCanonNet


def default_pointnet_cfg():
    """configs/pointnet_config/pointnet2_camera.yml:1-44 (the unused lstm*/flow1 keys dropped)."""
    return {
        "sa1": {"npoint": 512, "radius_list": [0.05, 0.1, 0.2], "nsample_list": [32, 64, 128],
                "mlp_list": [[32, 32, 64], [64, 64, 128], [64, 96, 128]]},
        "sa2": {"npoint": 128, "radius_list": [0.2, 0.4], "nsample_list": [64, 128],
                "mlp_list": [[128, 128, 256], [128, 196, 256]]},
        "sa3": {"mlp": [256, 512, 1024]},
        "fp3": {"mlp": [256, 256]}, "fp2": {"mlp": [256, 128]}, "fp1": {"mlp": [128, 128]},
    }


CATEGORIES = {
    # the six NOCS-REAL275 categories, configs/obj_config/obj_info_nocs.yml:7-127 (one rigid part each, extra_dims 1
    # = one background class; `sym` per category :9,25,40,79,95,111) -- one checkpoint pair per category,
    # scripts/track/nocs/1_bottle.sh ... 6_mug.sh
    "bottle": dict(num_parts=1, sym=True, tree=[-1], extra_dims=1),
    "bowl": dict(num_parts=1, sym=True, tree=[-1], extra_dims=1),
    "camera": dict(num_parts=1, sym=False, tree=[-1], extra_dims=1),
    "can": dict(num_parts=1, sym=True, tree=[-1], extra_dims=1),
    "nocs_laptop": dict(num_parts=1, sym=False, tree=[-1], extra_dims=1),
    "mug": dict(num_parts=1, sym=False, tree=[-1], extra_dims=1),
    # SAPIEN articulated laptop, obj_info_sapien.yml:52-65 (two parts, no background class)
    "laptop": dict(num_parts=2, sym=False, tree=[-1, 0], extra_dims=0),
}
NOCS_CATEGORIES = ("bottle", "bowl", "camera", "can", "nocs_laptop", "mug")     # obj_category 1..6


def make_cfg(category="bottle", device="cuda:0"):
    c = CATEGORIES[category]
    return {
        "pointnet": {"camera": default_pointnet_cfg()}, "device": device,
        "network": {"type": "rot_coord_track", "backbone_out_dim": 128, "nocs_head_dims": [128]},
        "num_parts": c["num_parts"], "obj_sym": c["sym"], "obj_tree": c["tree"],
        "obj": {"extra_dims": c["extra_dims"]}, "num_points": 4096,
    }


def init_weights(module, seed=0):
    """Reference init (trainer.py:18-37,111 with weight_init: xavier -- xavier_normal_, gain sqrt(2), zero bias, on
    every Conv*/Linear*; norm layers keep their defaults) + randomised BatchNorm running statistics so the
    folded-BN path is exercised.  Every tensor is drawn from its own generator seeded by
    (seed, state-dict key), so any module with the same state-dict keys and shapes -- the reference's own
    CoordNet / PartCanonNet (tests/golden/make_golden.py) or this package's mirrors -- gets bit-identical
    values, independent of module registration order.  Deterministic for a given torch build."""
    import zlib
    mods = dict(module.named_modules())

    def gen(key):
        return torch.Generator().manual_seed((zlib.crc32(key.encode()) + 7919 * seed) % (2 ** 31))
    with torch.no_grad():
        for name, m in mods.items():
            pre = name + "." if name else ""
            if isinstance(m, (torch.nn.Conv1d, torch.nn.Conv2d, torch.nn.Linear)):
                fan_in = m.weight.shape[1] * (m.weight[0, 0].numel() if m.weight.dim() > 2 else 1)
                fan_out = m.weight.shape[0] * (m.weight[0, 0].numel() if m.weight.dim() > 2 else 1)
                std = (2.0 ** 0.5) * (2.0 / (fan_in + fan_out)) ** 0.5
                m.weight.copy_(torch.randn(m.weight.shape, generator=gen(pre + "weight")) * std)
                if m.bias is not None:
                    m.bias.zero_()
            elif isinstance(m, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d)):
                m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=gen(pre + "running_mean")) * 0.1)
                m.running_var.copy_(torch.rand(m.running_var.shape, generator=gen(pre + "running_var")) + 0.5)
    return module


def make_trained_like(coordnet, num_parts):
    """Synthetic stand-in for a TRAINED CoordNet (there are no checkpoints offline).  Random weights give NOCS
    predictions uncorrelated with the cloud, and the scale fit then is a small difference of large sums: any
    1e-5 feature noise shows up as 1e-3 on the scale.  A trained CoordNet predicts NOCS ~ canonical coordinates
    and a segmentation that follows the geometry, so route the canonicalised xyz (skip connection of fp1,
    backbones.py:67) through to the heads: channel i carries relu(x_i), channel 3+i relu(-x_i); the NOCS head
    outputs sigmoid(4 x_i + 0.05 * (random deep features)) - 0.5; the segmentation splits the parts along x
    (P > 1) and calls |x| > 0.3 background (extra class).  Works on the reference's CoordNet and on this
    package's mirror alike (same attribute names)."""
    bb = coordnet.backbone
    with torch.no_grad():
        def passthrough(conv, bn, first=False):
            w = conv.weight
            w[:6] = 0
            if first:
                for i in range(3):
                    w[i, i] = 1.0
                    w[3 + i, i] = -1.0
            else:
                for i in range(6):
                    w[i, i] = 1.0
            conv.bias[:6] = 0
            if bn is not None:
                bn.weight[:6] = 1.0
                bn.bias[:6] = 0
                bn.running_mean[:6] = 0
                bn.running_var[:6] = 1.0 - bn.eps
        passthrough(bb.fp1.mlp_convs[0], bb.fp1.mlp_bns[0], first=True)
        passthrough(bb.fp1.mlp_convs[1], bb.fp1.mlp_bns[1])
        passthrough(bb.conv1, bb.bn1)
        passthrough(coordnet.nocs_head[0], coordnet.nocs_head[1])
        last = coordnet.nocs_head[3]
        last.weight.mul_(0.05)
        last.bias.zero_()
        for p in range(num_parts):
            for i in range(3):
                last.weight[3 * p + i, :6] = 0
                last.weight[3 * p + i, i] = 4.0
                last.weight[3 * p + i, 3 + i] = -4.0
        seg = coordnet.seg_head[0]
        seg.weight.mul_(0.05)
        seg.bias.zero_()
        seg.weight[:, :6] = 0
        for p in range(num_parts):                    # parts: slabs along x
            c = (p + 0.5) / num_parts - 0.5           # slab centre in [-0.5, 0.5]
            seg.weight[p, 0], seg.weight[p, 3] = 16.0 * c, -16.0 * c
            seg.bias[p] = -8.0 * c * c
        if seg.weight.shape[0] > num_parts:           # background: |x| > 0.3
            seg.weight[num_parts, 0] = seg.weight[num_parts, 3] = 8.0
            seg.bias[num_parts] = -2.4
    return coordnet


def state_dict_digest(module):
    """sha256 over the float tensors of a state dict in key order (pins init_weights across machines)."""
    import hashlib
    h = hashlib.sha256()
    for k, v in sorted(module.state_dict().items()):
        if v.is_floating_point():
            h.update(k.encode())
            h.update(v.detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


class Tracker(torch.nn.Module):
    """npcs_net + net of EvalTrackModel (model.py:312-313)."""

    def __init__(self, cfg, seed=0):
        super().__init__()
        self.cfg = cfg
        self.npcs_net = init_weights(CoordNet(cfg), seed)
        self.net = init_weights(PartCanonNet(cfg), seed + 1)
        self.num_parts = cfg["num_parts"]
        self.root = [p for p in range(self.num_parts) if cfg["obj_tree"][p] == -1][0]
        self._side = {}

    # The rotation network's encoder and heads do not depend on the CoordNet's output (only the final masked mean needs
    # its labels), and each chain has phases that leave most of the GPU idle (FPS: one CTA per cloud; the group-all /
    # fp3 layers: 32 tiles for 148 SMs).  CAPTRA_TWO_STREAM=0 runs the two networks back to back (A/B knob).
    TWO_STREAM = os.environ.get("CAPTRA_TWO_STREAM", "1") != "0"

    def _side_stream(self, key):
        """key: a device, or (device, name) for the helper streams."""
        if key not in self._side:
            self._side[key] = torch.cuda.Stream(key[0] if isinstance(key, tuple) else key)
        return self._side[key]

    def _step_two_stream(self, points, points_mean, last_pose, canon):
        from . import frame_ops
        from .pointnet_utils import SharedGeom
        P, B = self.num_parts, points.shape[0]
        main = torch.cuda.current_stream(points.device)
        side = self._side_stream(points.device)
        aux_a, aux_b = self._side_stream((points.device, "aux_a")), self._side_stream((points.device, "aux_b"))
        geom = SharedGeom(aux_a)                    # the CoordNet's geometry (shared with the RotationNet for a rigid object)
        geom_b = geom if P == 1 else SharedGeom(aux_b)
        for st in (side, aux_a, aux_b):
            st.wait_stream(main)                    # the frame's inputs
        # CoordNet on the current stream; every coordinate-only result it produces carries an event
        pred = self.npcs_net({"points": points, "points_mean": points_mean, "canon_pose": canon, "geom": geom})
        with torch.cuda.stream(side):
            if P == 1:
                xyz_pm = geom["xyz_pm"]             # waits (on the device) for the CoordNet's canonicalisation only
            else:
                flat = {k: last_pose[k].reshape((-1,) + last_pose[k].shape[2:]) for k in ("rotation", "translation", "scale")}
                xyz_pm = frame_ops.canonicalize(points, points_mean, flat["rotation"], flat["translation"], flat["scale"], parts=P)[0]
            raws = self.net.regress_net.forward_heads(xyz_pm, B, geom=geom_b)
        for st in (side, aux_a, aux_b):
            main.wait_stream(st)
        return pred, raws

    @torch.no_grad()
    def step(self, points, points_mean, last_pose, want_pred=False):
        """model.py:454-476.  points [B,3,N] (mean-subtracted), points_mean [B,3,1],
        last_pose {'rotation' [B,P,3,3], 'translation' [B,P,3,1], 'scale' [B,P]} -> new pose dict
        (with want_pred: (pose, {'seg','nocs','labels','points'}), the CoordNet predictions of this frame)."""
        canon = {k: last_pose[k][:, self.root] for k in ("rotation", "translation", "scale")}
        # With one rigid part both networks see the cloud canonicalised by the same pose
        # (networks.py:38-41 vs :184-187), so FPS picks, ball-query lists and 3-NN weights -- functions
        # of the coordinates only -- are computed once and shared (bit-identical results).
        extra = {}
        fused = points.is_cuda and not self.training
        if fused and self.TWO_STREAM:
            pred, extra["rot_raws"] = self._step_two_stream(points, points_mean, last_pose, canon)
            geom = None
        else:
            geom = {} if self.num_parts == 1 else None
            pred = self.npcs_net({"points": points, "points_mean": points_mean, "canon_pose": canon, "geom": geom})
        B = points.shape[0]
        pred_npcs = pred["nocs"].reshape(B, self.num_parts, 3, -1)
        pred_labels = pred["labels"] if "labels" in pred else torch.max(pred["seg"], dim=-2)[1]
        out = self.net(dict({"points": points, "points_mean": points_mean, "state": {"part": last_pose},
                             "pred_labels": pred_labels, "pred_nocs": pred_npcs, "geom": geom}, **extra), test_mode=True)
        if want_pred:
            return out["part"], {"seg": pred["seg"], "nocs": pred["nocs"], "labels": pred_labels, "points": pred["points"]}
        return out["part"]

    def eval_sums(self, gt, pose, **kw):
        """Per-frame pose-error sums of model.py:523-526 (eval_part_full, yaxis_only = sym) on the device:
        [1, 5P + 5], additive over trajectories and ranks (frame_ops.track_eval / eval_means)."""
        from . import frame_ops
        return frame_ops.track_eval(gt, pose, self.cfg["obj_sym"], **kw).view(1, -1)


class MixedTracker(torch.nn.Module):
    """BASELINE cfg4: one batch holding trajectories of several NOCS categories, each category with its own pair of
    networks (the reference tracks one category per checkpoint: scripts/track/nocs/1_bottle.sh ... 6_mug.sh,
    --obj_category=k).  The batch is kept GROUPED BY CATEGORY (group_by_category gives the permutation), so each
    category's clouds are one contiguous slice: no gather / scatter, every group runs its own launch set on its own
    weights, and the result is identical -- bit for bit -- to running each category's Tracker alone on its slice."""

    def __init__(self, categories, device="cuda:0", seed=0):
        """categories: the category NAME of every cloud of the (already grouped) batch, e.g. from group_by_category."""
        super().__init__()
        self.spans = []                       # (category, start, stop), contiguous
        for i, c in enumerate(categories):
            if self.spans and self.spans[-1][0] == c:
                self.spans[-1][2] = i + 1
            else:
                assert all(c != s[0] for s in self.spans), "clouds must be grouped by category (use group_by_category)"
                self.spans.append([c, i, i + 1])
        parts = {CATEGORIES[c]["num_parts"] for c, _, _ in self.spans}
        assert len(parts) == 1, "categories of one batch must have the same number of parts (pose tensors are [B,P,...])"
        self.num_parts = parts.pop()
        # per-category weights: seed offset by the category's NOCS id so the six weight sets differ
        self.trackers = torch.nn.ModuleDict(
            {c: Tracker(make_cfg(c, device=str(device)), seed=seed + 10 * (list(CATEGORIES).index(c) + 1)) for c, _, _ in self.spans})

    # Each category's slice is an independent launch set on its own weights: run them on separate streams, so one
    # category's latency-bound phases (FPS: one CTA per cloud) overlap the others' tensor work.  Matters most when a
    # rank holds only a handful of clouds per category (cfg4 on 8 GPUs: 5-6).  CAPTRA_CATEGORY_STREAMS=0: back to back.
    CATEGORY_STREAMS = os.environ.get("CAPTRA_CATEGORY_STREAMS", "1") != "0"

    @torch.no_grad()
    def step(self, points, points_mean, last_pose):
        outs = []
        concurrent = self.CATEGORY_STREAMS and points.is_cuda and len(self.spans) > 1
        if concurrent:
            main = torch.cuda.current_stream(points.device)
            if not hasattr(self, "_cat_streams"):
                self._cat_streams = [torch.cuda.Stream(points.device) for _ in self.spans]
        for i, (c, a, b) in enumerate(self.spans):
            args = (points[a:b], points_mean[a:b], {k: v[a:b] for k, v in last_pose.items()})
            if concurrent:
                st = self._cat_streams[i]
                st.wait_stream(main)
                with torch.cuda.stream(st):
                    outs.append(self.trackers[c].step(*args))
            else:
                outs.append(self.trackers[c].step(*args))
        if concurrent:
            for st in self._cat_streams:
                main.wait_stream(st)
        if len(outs) == 1:
            return outs[0]
        return {k: torch.cat([o[k] for o in outs], dim=0) for k in outs[0]}

    def eval_sums(self, gt, pose, out=None, accumulate=False):
        """One row of sums per category, in span order: [ncat, 5P + 5]."""
        if out is None:
            out = torch.zeros(len(self.spans), 5 * self.num_parts + 5, dtype=torch.float32, device=pose["scale"].device)
        for i, (c, a, b) in enumerate(self.spans):
            self.trackers[c].eval_sums({k: v[a:b] for k, v in gt.items()}, {k: v[a:b] for k, v in pose.items()},
                                       out=out[i], accumulate=accumulate)
        return out


def group_by_category(category_ids, names=NOCS_CATEGORIES):
    """cfg4 assigns category = global trajectory index mod 6 (shard.category_of).  Returns (order, names_sorted):
    `order` is the stable permutation that groups a shard's trajectories by category, names_sorted the category name
    of every trajectory after that permutation (MixedTracker's constructor argument)."""
    order = sorted(range(len(category_ids)), key=lambda i: (category_ids[i], i))
    return order, [names[category_ids[i]] for i in order]


class GraphedStep:
    """Tracker.step captured once into a CUDA graph (static shapes): ~300 kernel launches per frame
    become one cudaGraphLaunch, which removes the launch-bound gaps between the many small kernels.
    Inputs are copied into the graph's static buffers; the returned pose tensors are the graph's
    static outputs (clone them to keep a frame's result across calls)."""

    def __init__(self, tracker, points, points_mean, pose, warmup=2, gt=None, check_every=0):
        """gt (optional part-pose dict): the graph then also evaluates tracker.eval_sums(gt, new pose) -- the
        end-of-frame loss / metric reduction -- into self.sums (static), with no extra host launches.
        check_every = k > 0: every k-th replay synchronises and reads the fp16x3 saturation flag (mlp.f16_overflowed);
        a saturated operand means the result is not fp32-accurate and raises (re-run with CAPTRA_MLP_IMPL=1).  The
        flag is always checked once here, after the warm-up frames."""
        from . import _lib, mlp
        self.check_every, self._replays = int(check_every), 0
        if mlp.DEFAULT_IMPL == 2:
            mlp.f16_overflowed(reset=True)
        self.inp = {"points": points.clone(), "points_mean": points_mean.clone(),
                    "pose": {k: v.clone() for k, v in pose.items()}}
        self.gt = {k: v.clone().float() for k, v in gt.items()} if gt is not None else None
        self.sums = None
        if gt is not None:
            ncat = len(tracker.spans) if hasattr(tracker, "spans") else 1
            self.sums = torch.zeros(ncat, 5 * tracker.num_parts + 5, dtype=torch.float32, device=points.device)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):            # packs weights, sets kernel attributes, warms the allocator
                tracker.step(self.inp["points"], self.inp["points_mean"], self.inp["pose"])
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self._check_overflow()
        self.graph = torch.cuda.CUDAGraph()
        n0 = _lib.launch_count()
        with torch.cuda.graph(self.graph):
            self.out = tracker.step(self.inp["points"], self.inp["points_mean"], self.inp["pose"])
            if self.gt is not None:      # accumulated over the frames replayed since the last self.sums.zero_()
                self.sums = tracker.eval_sums(self.gt, self.out, out=self.sums, accumulate=True)
        self.launches_per_replay = _lib.launch_count() - n0     # this library's kernels inside the graph

    @staticmethod
    def _check_overflow():
        from . import _lib, mlp
        if mlp.DEFAULT_IMPL == 2 and mlp.f16_overflowed(reset=True):
            raise _lib.CaptraError("an fp16x3 operand saturated (|activation| >= 65504): poses are not fp32-accurate for this "
                                   "model / input; run with CAPTRA_MLP_IMPL=1 (3xTF32, full exponent range)")

    def __call__(self, points, points_mean, pose, gt=None):
        self._replays += 1
        if self.check_every and self._replays % self.check_every == 0:
            torch.cuda.synchronize()
            self._check_overflow()
        if gt is not None:
            for k, v in self.gt.items():
                v.copy_(gt[k], non_blocking=True)
        self.inp["points"].copy_(points, non_blocking=True)
        self.inp["points_mean"].copy_(points_mean, non_blocking=True)
        for k, v in self.inp["pose"].items():
            v.copy_(pose[k], non_blocking=True)
        self.graph.replay()
        return self.out


def synthetic_track_batch(b, category="bottle", n=4096, seed=0):
    """Host-side (numpy, float32) inputs of one frame for b trajectories: mean-subtracted camera
    points [b,3,n], their mean [b,3,1] and a perturbed initial pose per part (pose_perturb of
    config_track.yml:46-50: r=5 deg, t=0.03, s=0.02)."""
    c = CATEGORIES[category]
    P = c["num_parts"]
    case = synthetic.pose_fit_case(b, P, n, seed=seed, sym=False)
    cam = case["cam"]                                  # [b,n,3]
    mean = cam.mean(1, keepdims=True)
    points = np.ascontiguousarray(np.swapaxes(cam - mean, 1, 2)).astype(np.float32)
    R = case["R"].copy()
    t, s = case["t"].copy(), case["s"].copy()
    for bi in range(b):
        rng = np.random.default_rng([seed + 1, bi])    # per trajectory: cloud i is the same whatever the batch size
        for pi in range(P):
            axis = rng.normal(size=3)
            axis /= np.linalg.norm(axis)
            ang = np.deg2rad(5.0) * rng.normal()
            K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
            dR = np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K
            R[bi, pi] = R[bi, pi] @ dR
            t[bi, pi] += rng.normal(scale=0.03, size=3)
            s[bi, pi] += rng.normal(scale=0.02)
    pose = {"rotation": R.astype(np.float32), "translation": t[..., None].astype(np.float32), "scale": s.astype(np.float32)}
    return {"points": points, "points_mean": np.swapaxes(mean, 1, 2).astype(np.float32), "pose": pose,
            "gt": {"rotation": case["R"], "translation": case["t"][..., None], "scale": case["s"]}}
