"""Timing probes for the fused tcgen05 MLP kernel: runs the two dominant SA scales of cfg2 with the
CAPTRA_TC_DBG knobs (1 no A stores, 2 no W copies, 4 no MMAs, 8 no last epilogue; results are
garbage, only the time matters) to see which role bounds the tile time."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from captra_b200 import fused_ops, synthetic  # noqa: E402
from captra_b200.mlp import PackedMLP  # noqa: E402

dev = torch.device("cuda:0")
B = 32
gen = torch.Generator().manual_seed(0)


def mk(cin, couts):
    ws, bs, last = [], [], cin
    for c in couts:
        ws.append((torch.randn(c, last, generator=gen) / last ** 0.5).to(dev))
        bs.append((0.1 * torch.randn(c, generator=gen)).to(dev))
        last = c
    return PackedMLP(ws, bs, impl=int(os.environ.get('PROBE_IMPL', '1')))


cases = []
pts = torch.from_numpy(synthetic.batch_surface_box(B, 4096, seed=0)[0]).to(dev)
_, c1 = fused_ops.fps_gather(pts, 512)
i1 = fused_ops.ball_query_multi([0.2], [128], pts, c1)[0]
cases.append(("sa1 K=128 6->64-96-128", mk(6, [64, 96, 128]), pts, c1, pts.contiguous(), i1))
_, c2 = fused_ops.fps_gather(c1, 128)
i2 = fused_ops.ball_query_multi([0.4], [128], c1, c2)[0]
f2 = torch.randn(B, 512, 320, generator=gen).to(dev)
cases.append(("sa2 K=128 323->128-196-256", mk(323, [128, 196, 256]), c1, c2, f2, i2))
i3 = fused_ops.ball_query_multi([0.2], [64], c1, c2)[0]
cases.append(("sa2 K=64 323->128-128-256", mk(323, [128, 128, 256]), c1, c2, f2, i3))

for name, mlp, xyz, ctr, feats, idx in cases:
    out = torch.empty(B, ctr.shape[1], mlp.cout, device=dev)
    for dbg in (0,):
        os.environ["CAPTRA_TC_DBG"] = str(dbg)
        for _ in range(2):
            mlp.sa_max(xyz, ctr, feats, idx, out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            mlp.sa_max(xyz, ctr, feats, idx, out)
        e1.record()
        torch.cuda.synchronize()
        print("%-30s dbg=%2d  %8.1f us" % (name, dbg, 1e3 * e0.elapsed_time(e1) / 5))
os.environ["CAPTRA_TC_DBG"] = "0"

# phase timestamps of CTA 0's first tiles (dbg bit 32)
import ctypes
from captra_b200 import _lib
L = _lib.load()
for name, mlp, xyz, ctr, feats, idx in cases:
    out = torch.empty(B, ctr.shape[1], mlp.cout, device=dev)
    for dbg in (32,):
        buf = (ctypes.c_longlong * 512)()
        L.captra_debug_tc_timestamps(buf, 255)
        os.environ["CAPTRA_TC_DBG"] = str(dbg)
        mlp.sa_max(xyz, ctr, feats, idx, out)
        n = L.captra_debug_tc_timestamps(buf, 255)
        ts = [(buf[2 * i], buf[2 * i + 1]) for i in range(n)]
        print(name, "dbg", dbg, "stamps", n)
        line = []
        for i in range(1, min(n, 19)):
            line.append("%d->%d:%d" % (ts[i - 1][1], ts[i][1], ts[i][0] - ts[i - 1][0]))
        print("  " + "  ".join(line))
os.environ["CAPTRA_TC_DBG"] = "0"
