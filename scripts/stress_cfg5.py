"""BASELINE.json configs[4] "dense stress": B=64 clouds x 16384 points, nsample K=64, 4 SA levels
(4096/1024/256/64 centroids, radii .05/.1/.2/.4) -- the ball_query + group HBM-roofline run.
Prints one JSON line per level and a total: algorithmic bytes / CUDA-event time.

Three variants per level:
  query        ball_query alone                      (12 B (N+M) + 4 B M K bytes)
  group        group_points alone on C channels     (4 B C N + 4 B M K + 4 B C M K bytes)
  query+group  the two back to back (what the metric names)
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from captra_b200 import fused_ops, synthetic  # noqa: E402
from captra_b200.pointnet_lib import pointnet2_utils as futils  # noqa: E402

PEAK = 6556.2
if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")):
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]

dev = torch.device("cuda:0")
B, K = int(os.environ.get("STRESS_B", 64)), 64
kind = os.environ.get("STRESS_CLOUD", "surface")
if kind == "uniform":
    pts = synthetic.batch_uniform(B, 16384, seed=0)
else:
    pts = np.stack([synthetic.surface_box(16384, np.random.default_rng(i))[0] for i in range(B)])
xyz = torch.from_numpy(pts).to(dev)
levels = [(4096, 0.05, 3), (1024, 0.1, 128), (256, 0.2, 256), (64, 0.4, 256)]   # (centroids, radius, channels grouped)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)) * 1e-3


total_bytes, total_t = 0, 0.0
cur = xyz
for li, (M, r, C) in enumerate(levels):
    N = cur.shape[1]
    _, ctr = fused_ops.fps_gather(cur, M)
    feats = torch.randn(B, C, N, device=dev)
    idx = futils.ball_query(r, K, cur, ctr)
    t_q = timed(lambda: futils.ball_query(r, K, cur, ctr))
    t_g = timed(lambda: futils.grouping_operation(feats, idx))
    t_qg = timed(lambda: futils.grouping_operation(feats, futils.ball_query(r, K, cur, ctr)))
    by_q = 12 * B * (N + M) + 4 * B * M * K
    by_g = 4 * B * C * N + 4 * B * M * K + 4 * B * C * M * K
    by_qg = 12 * B * (N + M) + 4 * B * C * N + 4 * B * M * K + 4 * B * C * M * K
    hits = float((idx != idx[..., :1]).float().mean())
    print(json.dumps({"level": li + 1, "N": N, "M": M, "K": K, "radius": r, "C": C, "cloud": kind,
                      "query_us": t_q * 1e6, "query_gbs": by_q / t_q / 1e9,
                      "group_us": t_g * 1e6, "group_gbs": by_g / t_g / 1e9, "group_frac": by_g / t_g / 1e9 / PEAK,
                      "query_group_us": t_qg * 1e6, "query_group_gbs": by_qg / t_qg / 1e9,
                      "query_group_frac": by_qg / t_qg / 1e9 / PEAK, "distinct_frac": hits}))
    total_bytes += by_qg
    total_t += t_qg
    cur = ctr
print(json.dumps({"total": True, "alg_bytes": total_bytes, "us": total_t * 1e6, "gbs": total_bytes / total_t / 1e9,
                  "frac_of_hbm_peak": total_bytes / total_t / 1e9 / PEAK, "peak_gbs": PEAK, "B": B}))
