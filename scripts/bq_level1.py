"""BASELINE cfg5 level 1 (B clouds x 16384 points, 4096 centroids, r = 0.05, K = 64, xyz grouped) three times: the workload
scripts/ncu_source.sh and the ncu launch lists profile.  python scripts/bq_level1.py [B] [surface|uniform]"""
import numpy as np, torch, sys
sys.path.insert(0, '.')
from captra_b200 import fused_ops, synthetic
dev = torch.device('cuda:0')
B = 64 if len(sys.argv) < 2 else int(sys.argv[1])
kind = 'surface' if len(sys.argv) < 3 else sys.argv[2]
if kind == 'surface':
    pts = np.stack([synthetic.surface_box(16384, np.random.default_rng(i))[0] for i in range(B)])
else:
    pts = synthetic.batch_uniform(B, 16384, seed=0)
cur = torch.from_numpy(pts).to(dev)
_, ctr = fused_ops.fps_gather(cur, 4096)
feats = cur.transpose(1, 2).contiguous()
for _ in range(3):
    fused_ops.ball_query_group(0.05, 64, cur, ctr, feats)
torch.cuda.synchronize()
