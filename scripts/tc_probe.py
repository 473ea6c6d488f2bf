"""Timing probe for the fused tcgen05 MLP kernel: the SA scales and dense-row shapes of cfg2, each
timed alone (PROBE_IMPL = 1 | 2; PROBE_STAMPS=1 prints the clock64 phase stamps of a build made with
CAPTRA_TC_DBG bit 32); PROBE_DBG=0,2,4,... times the knock-out knobs.  The knobs and stamps exist only in the probes
build of the library (-DCAPTRA_TC_PROBES, see csrc/mlp_tc.cu): select it with CAPTRA_LIB_PATH; the product library ignores them."""
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
    return PackedMLP(ws, bs, impl=int(os.environ.get('PROBE_IMPL', '2')))


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

for r, k in ((0.1, 64), (0.05, 32)):
    ii = fused_ops.ball_query_multi([r], [k], pts, c1)[0]
    cases.append(("sa1 K=%d 6->%s" % (k, "32-32-64" if k == 32 else "64-64-128"), mk(6, [32, 32, 64] if k == 32 else [64, 64, 128]), pts, c1, pts.contiguous(), ii))


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / n


# projected layer 0 (captra_sa_mlp_max_pre): the sa2 scales as the tracker runs them
for k, idxk, couts in ((128, i2, [196, 256]), (64, i3, [128, 256])):
    tail = mk(128, couts)
    P = torch.randn(B * 512, 256, generator=gen).to(dev)
    tab = torch.randn(4, 128, generator=gen).to(dev) * 0.1
    cases.append(("sa2pre K=%d 128->%s" % (k, "-".join(map(str, couts))), tail, c1, c2, (P[:, :128], tab), idxk))

DBGS = [int(v) for v in os.environ.get("PROBE_DBG", "0").split(",")]
if os.environ.get("PROBE_CASES"):   # e.g. PROBE_CASES=0,1 keeps only those SA cases (ncu captures)
    cases = [cases[int(i)] for i in os.environ["PROBE_CASES"].split(",")]
N_TIMED = int(os.environ.get("PROBE_N", "5"))


def run_case(mlp, xyz, ctr, feats, idx, out):
    if isinstance(feats, tuple):
        mlp.sa_max_pre(xyz, ctr, feats[0], feats[1], idx, out)
    else:
        mlp.sa_max(xyz, ctr, feats, idx, out)


for name, mlp, xyz, ctr, feats, idx in cases:
    out = torch.empty(B, ctr.shape[1], mlp.cout, device=dev)
    rows = B * ctr.shape[1] * idx.shape[2]
    fl = 2.0 * rows * sum(a * b for a, b in zip([mlp.cin] + mlp.couts[:-1], mlp.couts))
    for dbg in DBGS:   # CAPTRA_TC_DBG knobs: results are garbage, only the time matters
        os.environ["CAPTRA_TC_DBG"] = str(dbg)
        us = timeit(lambda: run_case(mlp, xyz, ctr, feats, idx, out), N_TIMED)
        print("%-30s dbg=%2d %8.1f us  %6.1f TFLOP/s (algorithmic)" % (name, dbg, us, fl / us * 1e-6))
os.environ["CAPTRA_TC_DBG"] = "0"

if os.environ.get("PROBE_DENSE", "1") == "0":
    sys.exit(0)
# dense rows: the RotationRegressor head layers (GroupNorm affine on load) and fp1 + conv1
R = B * 4096
x512 = torch.randn(R, 512, generator=gen).to(dev)
sc, sh = torch.rand(B, 512, generator=gen).to(dev) + 0.5, torch.randn(B, 512, generator=gen).to(dev)
for cin, cout in ((512, 512), (512, 256)):
    m = PackedMLP([(torch.randn(cout, cin, generator=gen) / cin ** 0.5).to(dev)], [torch.zeros(cout).to(dev)], relu_last=False,
                  impl=int(os.environ.get('PROBE_IMPL', '2')))
    y = torch.empty(R, cout, device=dev)
    us = timeit(lambda: m.rows_affine(x512, sc, sh, 4096, out=y))
    print("%-30s %8.1f us  %6.1f TFLOP/s" % ("head %d->%d affine" % (cin, cout), us, 2.0 * R * cin * cout / us * 1e-6))
x128 = torch.randn(R, 128, generator=gen).to(dev)
for couts in ([512], [128, 128, 128]):
    m = mk(128, couts)
    us = timeit(lambda: m.rows(x128))
    fl = 2.0 * R * sum(a * b for a, b in zip([128] + couts[:-1], couts))
    print("%-30s %8.1f us  %6.1f TFLOP/s" % ("rows 128->%s" % "-".join(map(str, couts)), us, fl / us * 1e-6))

def stamps(name, fn):
    import ctypes
    from captra_b200 import _lib
    L = _lib.load()
    buf = (ctypes.c_longlong * 512)()
    L.captra_debug_tc_timestamps(buf, 255)
    os.environ["CAPTRA_TC_DBG"] = "32"
    fn()
    os.environ["CAPTRA_TC_DBG"] = "0"
    n = L.captra_debug_tc_timestamps(buf, 255)
    ts = [(buf[2 * i], buf[2 * i + 1]) for i in range(n)]
    print(name, "stamps", n)
    print("  " + "  ".join("%d->%d:%d" % (ts[i - 1][1], ts[i][1], ts[i][0] - ts[i - 1][0]) for i in range(1, min(n, 25))))


if os.environ.get("PROBE_STAMPS"):
    m = PackedMLP([(torch.randn(512, 512, generator=gen) / 512 ** 0.5).to(dev)], [torch.zeros(512).to(dev)], relu_last=False,
                  impl=int(os.environ.get('PROBE_IMPL', '2')))
    y = torch.empty(R, 512, device=dev)
    stamps("head 512->512 affine", lambda: m.rows_affine(x512, sc, sh, 4096, out=y))
    m = mk(128, [512])
    stamps("rows 128->512", lambda: m.rows(x128))
    m = mk(128, [128, 128, 128])
    stamps("rows 128->128-128-128", lambda: m.rows(x128))
    import ctypes
    from captra_b200 import _lib
    L = _lib.load()
    for name, mlp, xyz, ctr, feats, idx in cases:
        out = torch.empty(B, ctr.shape[1], mlp.cout, device=dev)
        buf = (ctypes.c_longlong * 512)()
        L.captra_debug_tc_timestamps(buf, 255)
        os.environ["CAPTRA_TC_DBG"] = "32"
        run_case(mlp, xyz, ctr, feats, idx, out)
        os.environ["CAPTRA_TC_DBG"] = "0"
        n = L.captra_debug_tc_timestamps(buf, 255)
        ts = [(buf[2 * i], buf[2 * i + 1]) for i in range(n)]
        print(name, "stamps", n)
        print("  " + "  ".join("%d->%d:%d" % (ts[i - 1][1], ts[i][1], ts[i][0] - ts[i - 1][0]) for i in range(1, min(n, 19))))
