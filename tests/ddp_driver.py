"""Runs under torchrun with 2 ranks (tests/test_ddp_gpu.py): the mirror CoordNet in TRAINING mode wrapped in
torch DistributedDataParallel (NCCL), two SGD steps on per-rank shards of a synthetic batch -- the training-side use of
the drop-in ops (gradient all-reduce over NVLink).  Rank 0 prints one JSON line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from torch.nn.parallel import DistributedDataParallel as DDP  # noqa: E402

from captra_b200 import networks, track  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False

cfg = track.make_cfg("laptop", device=str(dev))
P = cfg["num_parts"]
net = track.init_weights(networks.CoordNet(cfg), 7).to(dev).train()
ddp = DDP(net, device_ids=[local])
opt = torch.optim.SGD(ddp.parameters(), lr=1e-3)
losses = []
for step in range(2):
    b = track.synthetic_track_batch(2, "laptop", n=1024, seed=100 * rank + step)       # every rank its own trajectories
    gen = torch.Generator().manual_seed(rank)
    inp = {"points": torch.from_numpy(b["points"]).to(dev), "points_mean": torch.from_numpy(b["points_mean"]).to(dev),
           "labels": torch.randint(0, P, (2, 1024), generator=gen).to(dev),
           "gt_part": {k: torch.from_numpy(np.asarray(v, dtype=np.float32)).to(dev) for k, v in b["gt"].items()},
           "init_part": {k: torch.from_numpy(v).to(dev) for k, v in b["pose"].items()}}
    inp["canon_pose"] = {k: inp["init_part"][k][:, 0] for k in ("rotation", "translation", "scale")}
    pred = ddp(inp)
    loss = (pred["nocs"] ** 2).mean() + pred["seg"][:, 0].mean() + pred["part"]["scale"].sum() + pred["part"]["translation"].abs().sum()
    opt.zero_grad()
    loss.backward()
    opt.step()
    losses.append(float(loss.detach()))
# after the gradient all-reduce and the identical optimiser step the replicas must hold identical parameters
flat = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
digest = torch.stack([flat.double().sum(), flat.double().abs().sum(), flat.double().pow(2).sum()])
gathered = [torch.zeros_like(digest) for _ in range(world)]
dist.all_gather(gathered, digest)
same = all(torch.equal(g, gathered[0]) for g in gathered)
grads = sum(1 for p in net.parameters() if p.grad is not None)
if rank == 0:
    print(json.dumps({"world": world, "replicas_identical": bool(same), "losses_rank0": losses, "params_with_grad": grads,
                      "finite": bool(torch.isfinite(flat).all())}))
dist.destroy_process_group()
