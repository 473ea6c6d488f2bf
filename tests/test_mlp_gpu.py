"""GPU parity for the fused per-point MLP kernels and the PointNet++ modules built on them
(SURVEY section 8 rows a7, a8, a14).  Floating point: compared against plain torch fp32
(TF32 off) at rtol 1e-4 / atol 1e-5 -- the reference's own conv->BN->ReLU sequence differs from
a BN-folded GEMM by ~1e-6 relative -- and against the REFERENCE's module outputs in
tests/golden/backbone.npz (real reference modules, see tests/golden/make_golden.py)."""
import json
import os

import numpy as np
import pytest
import torch

from captra_b200 import synthetic

pytestmark = pytest.mark.gpu
TOL = dict(rtol=1e-4, atol=1e-5)
IMPLS = [0, 1, 2]


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False


def _rand_mlp(cin, couts, gen, cuda):
    ws, bs, last = [], [], cin
    for c in couts:
        ws.append((torch.randn(c, last, generator=gen) / last ** 0.5).to(cuda))
        bs.append((0.1 * torch.randn(c, generator=gen)).to(cuda))
        last = c
    return ws, bs


def _torch_mlp(x, ws, bs, relu_last=True):
    for i, (w, b) in enumerate(zip(ws, bs)):
        x = torch.addmm(b, x, w.t())
        if i < len(ws) - 1 or relu_last:
            x = torch.relu(x)
    return x


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("rows,ca,cb,couts", [
    (1000, 131, 0, [128, 128, 128]), (4096, 6, 128, [128, 128, 20]), (777, 3, 0, [32, 32, 64]),
    (512, 320, 256, [256, 128]), (130, 515, 0, [256, 512, 1024]), (64, 5, 7, [9]), (1, 16, 0, [16, 16, 16, 16])])
def test_point_mlp_rows(impl, rows, ca, cb, couts, cuda):
    from captra_b200.mlp import PackedMLP
    gen = torch.Generator().manual_seed(rows + ca)
    ws, bs = _rand_mlp(ca + cb, couts, gen, cuda)
    A = torch.randn(rows, ca + 5, generator=gen).to(cuda)[:, :ca]        # row stride != ca
    Bm = torch.randn(rows, cb, generator=gen).to(cuda) if cb else None
    x = torch.cat([A, Bm], 1) if cb else A
    for relu_last in (True, False):
        mlp = PackedMLP(ws, bs, relu_last=relu_last, impl=impl)
        got = mlp.rows(A, Bm)
        torch.testing.assert_close(got, _torch_mlp(x, ws, bs, relu_last), **TOL)


@pytest.mark.parametrize("impl", IMPLS)
def test_point_mlp_group_max_broadcast_and_offset(impl, cuda):
    from captra_b200.mlp import PackedMLP
    gen = torch.Generator().manual_seed(0)
    # group-all SA (pointnet_utils.py:319-343): max over the 128 points of each cloud
    B, N, D = 3, 128, 37
    ws, bs = _rand_mlp(3 + D, [64, 96, 200], gen, cuda)
    xyz, f = torch.randn(B * N, 3, generator=gen).to(cuda), torch.randn(B * N, D, generator=gen).to(cuda)
    got = PackedMLP(ws, bs, impl=impl).rows(xyz, f, group=N)
    want = _torch_mlp(torch.cat([xyz, f], 1), ws, bs).view(B, N, -1).max(1)[0]
    torch.testing.assert_close(got, want, **TOL)
    # groups smaller than a tile and not a power of two
    for g in (8, 24, 32, 100, 192):
        x = torch.randn(5 * g, 20, generator=gen).to(cuda)
        ws2, bs2 = _rand_mlp(20, [48, 70], gen, cuda)
        got = PackedMLP(ws2, bs2, impl=impl).rows(x, None, group=g)
        torch.testing.assert_close(got, _torch_mlp(x, ws2, bs2).view(5, g, -1).max(1)[0], **TOL)
    # FP3 (pointnet_utils.py:281-282): one coarse row repeated over the N points of its cloud
    ws3, bs3 = _rand_mlp(16 + 40, [32, 32], gen, cuda)
    p1, p2 = torch.randn(B * N, 16, generator=gen).to(cuda), torch.randn(B, 40, generator=gen).to(cuda)
    got = PackedMLP(ws3, bs3, impl=impl).rows(p1, p2, bcast_rows=N)
    want = _torch_mlp(torch.cat([p1, p2.repeat_interleave(N, 0)], 1), ws3, bs3)
    torch.testing.assert_close(got, want, **TOL)
    # write into a wider buffer at a column offset
    buf = torch.full((B * N, 100), -7.0, device=cuda)
    PackedMLP(ws3, bs3, impl=impl).rows(p1, p2, bcast_rows=N, out=buf, col_off=60)
    torch.testing.assert_close(buf[:, 60:92], want, **TOL)
    assert (buf[:, :60] == -7).all() and (buf[:, 92:] == -7).all()


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("N,S,K,cfeat,couts", [(4096, 512, 32, 3, [32, 32, 64]), (4096, 512, 128, 0, [64, 96, 128]),
                                                (512, 128, 64, 320, [128, 128, 256]), (300, 50, 48, 7, [16, 24]),
                                                (512, 128, 128, 320, [128, 196, 256])])
def test_sa_mlp_max_vs_torch(impl, N, S, K, cfeat, couts, cuda, oracle):
    from captra_b200.mlp import PackedMLP
    B = 2
    gen = torch.Generator().manual_seed(N + K)
    pts = synthetic.batch_surface_box(B, N, seed=K)[0]
    ctr_idx = oracle.furthest_point_sample(pts, S)
    ctr = np.stack([pts[b, ctr_idx[b]] for b in range(B)])
    idx = oracle.ball_query(0.25, K, pts, ctr)
    xyz, new_xyz, gidx = torch.from_numpy(pts).to(cuda), torch.from_numpy(ctr).to(cuda), torch.from_numpy(idx).to(cuda)
    feats = torch.randn(B, N, cfeat, generator=gen).to(cuda) if cfeat else None
    ws, bs = _rand_mlp(cfeat + 3, couts, gen, cuda)
    out = torch.full((B, S, couts[-1] + 8), -3.0, device=cuda)
    PackedMLP(ws, bs, impl=impl).sa_max(xyz, new_xyz, feats, gidx, out, col_off=8)
    bi = torch.arange(B, device=cuda).view(B, 1, 1)
    g_xyz = xyz[bi, gidx.long()] - new_xyz.unsqueeze(2)                    # [B,S,K,3]
    rows = torch.cat([feats[bi, gidx.long()], g_xyz], -1) if cfeat else g_xyz
    want = _torch_mlp(rows.reshape(-1, cfeat + 3), ws, bs).view(B, S, K, -1).max(2)[0]
    torch.testing.assert_close(out[..., 8:], want, **TOL)
    assert (out[..., :8] == -3).all()


@pytest.mark.parametrize("impl", [1, 2])
@pytest.mark.parametrize("off,ld,ca,cb", [(0, 64, 64, 0), (4, 72, 64, 0), (1, 67, 64, 0), (0, 64, 64, 64), (0, 40, 40, 24), (8, 136, 128, 6)])
def test_layer0_loader_alignment_paths(impl, off, ld, ca, cb, cuda):
    """The coalesced layer-0 loader of the tensor-core kernel picks per 8-channel unit between one 32-byte
    load, two 16-byte loads and scalar loads by the address it sees: rows that are 32-byte aligned, only
    16-byte aligned (column offset 4), unaligned (odd row stride), and units that straddle / lie in the
    second input segment must all give the same numbers."""
    from captra_b200.mlp import PackedMLP
    gen = torch.Generator().manual_seed(off + ld + ca + cb)
    rows = 3 * 128 + 37                                                    # a partial last tile
    big = torch.randn(rows, off + ld, generator=gen).to(cuda)
    A = big[:, off:off + ca]
    Bm = torch.randn(rows, cb, generator=gen).to(cuda) if cb else None
    x = torch.cat([A, Bm], 1) if cb else A
    for couts in ([64, 64], [256]):
        ws, bs = _rand_mlp(ca + cb, couts, gen, cuda)
        got = PackedMLP(ws, bs, impl=impl).rows(A, Bm)
        torch.testing.assert_close(got, _torch_mlp(x.contiguous(), ws, bs), **TOL)


@pytest.mark.parametrize("impl", [1, 2])
@pytest.mark.parametrize("S,K,cfeat", [(50, 32, 3), (50, 32, 0), (25, 64, 16), (13, 32, 323)])
def test_sa_metadata_ring_partial_tiles(impl, S, K, cfeat, cuda, oracle):
    """SA launches whose row count is not a multiple of the 128-row tile and that walk many tiles per CTA
    slot: the TMA warp's two-tile metadata ring must stay in step with the producers, rows past the
    end must contribute nothing, for the in-ring (<= 8 input channels) and the gathered layer 0."""
    from captra_b200.mlp import PackedMLP
    B, N = 1, 700
    gen = torch.Generator().manual_seed(S * K + cfeat)
    pts = synthetic.batch_surface_box(B, N, seed=S)[0]
    ctr_idx = oracle.furthest_point_sample(pts, S)
    ctr = np.stack([pts[b, ctr_idx[b]] for b in range(B)])
    idx = oracle.ball_query(0.3, K, pts, ctr)
    xyz, new_xyz, gidx = torch.from_numpy(pts).to(cuda), torch.from_numpy(ctr).to(cuda), torch.from_numpy(idx).to(cuda)
    feats = torch.randn(B, N, cfeat, generator=gen).to(cuda) if cfeat else None
    ws, bs = _rand_mlp(cfeat + 3, [32, 48, 64], gen, cuda)
    out = torch.empty(B, S, 64, device=cuda)
    PackedMLP(ws, bs, impl=impl).sa_max(xyz, new_xyz, feats, gidx, out)
    bi = torch.arange(B, device=cuda).view(B, 1, 1)
    g_xyz = xyz[bi, gidx.long()] - new_xyz.unsqueeze(2)
    rows = torch.cat([feats[bi, gidx.long()], g_xyz], -1) if cfeat else g_xyz
    want = _torch_mlp(rows.reshape(-1, cfeat + 3), ws, bs).view(B, S, K, -1).max(2)[0]
    torch.testing.assert_close(out, want, **TOL)


@pytest.mark.parametrize("impl", [1, 2])
@pytest.mark.parametrize("N,S,K,cfeat,couts", [(512, 128, 128, 320, [128, 196, 256]), (512, 128, 64, 320, [128, 128, 256]),
                                                (600, 37, 32, 45, [40, 24, 50]), (256, 64, 64, 64, [72, 256])])
def test_sa_mlp_max_projected_layer0(impl, N, S, K, cfeat, couts, cuda, oracle):
    """captra_sa_mlp_max_pre: layer 0 applied per point (W_f f) and completed in the loader
    (+ W_x (x - c) + b, ReLU) must equal the plain formulation on the concatenated rows."""
    from captra_b200.mlp import PackedMLP
    B = 2
    gen = torch.Generator().manual_seed(N + K + cfeat)
    pts = synthetic.batch_surface_box(B, N, seed=K)[0]
    ctr_idx = oracle.furthest_point_sample(pts, S)
    ctr = np.stack([pts[b, ctr_idx[b]] for b in range(B)])
    idx = oracle.ball_query(0.25, K, pts, ctr)
    xyz, new_xyz, gidx = torch.from_numpy(pts).to(cuda), torch.from_numpy(ctr).to(cuda), torch.from_numpy(idx).to(cuda)
    feats = torch.randn(B, N, cfeat, generator=gen).to(cuda)
    ws, bs = _rand_mlp(cfeat + 3, couts, gen, cuda)
    c1 = couts[0]
    proj = PackedMLP([ws[0][:, :cfeat].contiguous()], [torch.zeros(c1, device=cuda)], relu_last=False, impl=impl)
    P = torch.full((B * N, c1 + 24), 9.0, device=cuda)                        # a column block of a wider buffer
    proj.rows(feats.reshape(B * N, cfeat), out=P, col_off=8)
    tab = torch.cat([ws[0][:, cfeat:].t(), bs[0][None]], 0).contiguous()
    out = torch.full((B, S, couts[-1] + 8), -3.0, device=cuda)
    PackedMLP(ws[1:], bs[1:], impl=impl).sa_max_pre(xyz, new_xyz, P[:, 8:8 + c1], tab, gidx, out, col_off=8)
    bi = torch.arange(B, device=cuda).view(B, 1, 1)
    g_xyz = xyz[bi, gidx.long()] - new_xyz.unsqueeze(2)
    rows = torch.cat([feats[bi, gidx.long()], g_xyz], -1)
    want = _torch_mlp(rows.reshape(-1, cfeat + 3), ws, bs).view(B, S, K, -1).max(2)[0]
    torch.testing.assert_close(out[..., 8:], want, **TOL)
    assert (out[..., :8] == -3).all()


def _golden_backbone(tag, cuda):
    from captra_b200.backbones import PointNet2Msg
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "backbone.npz"))
    net_cfg = json.loads(bytes(g["net_cfg_json"]).decode())
    cfg = {"pointnet": {"camera": net_cfg}, "device": "cuda:0"}
    net = PointNet2Msg(cfg, out_dim=20, net_type="camera", use_xyz_feat=(tag == "coord"))
    sd = {k[len(tag) + 4:]: torch.from_numpy(g[k]) for k in g.files if k.startswith(tag + "/sd/")}
    missing, unexpected = net.load_state_dict(sd, strict=True)
    return net.to(cuda).eval(), torch.from_numpy(g[tag + "/input"]).to(cuda), torch.from_numpy(g[tag + "/output"]).to(cuda)


@pytest.fixture(params=IMPLS)
def impl(request, monkeypatch):
    from captra_b200 import mlp
    monkeypatch.setattr(mlp, "DEFAULT_IMPL", request.param)
    return request.param


@pytest.mark.parametrize("tag", ["coord", "rot"])
def test_backbone_matches_reference_modules(tag, impl, cuda):
    """The reference's own PointNet2Msg (torch Conv/BN, eval) produced tests/golden/backbone.npz;
    same weights through state_dict -> fused kernels must reproduce it."""
    net, x, want = _golden_backbone(tag, cuda)
    with torch.no_grad():
        got = net(x)
    assert got.shape == want.shape
    torch.testing.assert_close(got, want, rtol=1e-4, atol=2e-5)
    # unfused composition (the training-mode graph) in eval mode agrees too, and is differentiable
    xg = x.clone().requires_grad_(True)
    y = net(xg)
    torch.testing.assert_close(y, want, rtol=1e-4, atol=2e-5)
    y.sum().backward()
    assert torch.isfinite(xg.grad).all() and xg.grad.abs().sum() > 0


def test_backbone_repacks_after_weight_update(cuda):
    net, x, want = _golden_backbone("coord", cuda)
    with torch.no_grad():
        y0 = net(x).clone()
        net.conv1.weight.mul_(2.0)
        net.bn1.running_mean.add_(0.5)
        y1 = net(x)
    assert not torch.allclose(y0, y1)
    xg = x.clone().requires_grad_(True)
    torch.testing.assert_close(net(xg).detach(), y1, rtol=1e-4, atol=2e-5)


def test_backbone_full_size_shapes(impl, cuda):
    """BASELINE cfg2 shapes (pointnet2_camera.yml): B=4 here, 4096 points, both nets."""
    from captra_b200.backbones import PointNet2Msg
    from captra_b200.track import default_pointnet_cfg
    cfg = {"pointnet": {"camera": default_pointnet_cfg()}, "device": "cuda:0"}
    pts = torch.from_numpy(synthetic.batch_surface_box(4, 4096, seed=0)[0]).to(cuda).transpose(1, 2).contiguous()
    for use_xyz in (True, False):
        torch.manual_seed(0)
        net = PointNet2Msg(cfg, 128, use_xyz_feat=use_xyz).to(cuda).eval()
        with torch.no_grad():
            fused = net(pts)
        ref = net(pts.clone().requires_grad_(True)).detach()   # unfused torch composition
        assert fused.shape == (4, 128, 4096)
        torch.testing.assert_close(fused, ref, rtol=2e-4, atol=2e-5)


def test_rotation_head_groupnorm_fused_vs_torch(cuda):
    """RotationRegressor head (blocks.py:146-193: conv1d + GroupNorm(C/2) + ReLU x3 + conv1d): the fused
    path (tcgen05 GEMMs, GroupNorm folded into the consumer's operand load) vs the torch modules."""
    from captra_b200.networks import MLPConv1d
    torch.manual_seed(0)
    head = MLPConv1d(128, [512, 512, 256, 6]).to(cuda).eval()
    for m in head.modules():
        if isinstance(m, torch.nn.GroupNorm):
            m.weight.data.uniform_(0.5, 1.5)
            m.bias.data.normal_(0, 0.2)
    B, N = 3, 4096
    feat = torch.randn(B, 128, N, device=cuda)
    from captra_b200 import mlp
    with torch.no_grad():
        want = head(feat)                                            # [B, 6, N]
        for impl in (1, 2):
            mlp.DEFAULT_IMPL, old = impl, mlp.DEFAULT_IMPL
            head.__dict__.pop("_cache", None)            # re-pack for this impl
            try:
                got = head.forward_pm(feat.transpose(1, 2).contiguous()).transpose(1, 2)
            finally:
                mlp.DEFAULT_IMPL = old
            head.__dict__.pop("_cache", None)
            torch.testing.assert_close(got, want, rtol=1e-3, atol=2e-4)  # split-precision GEMMs through 4 layers + GroupNorm


def test_group_norm_affine_kernel(cuda):
    from captra_b200.mlp import group_norm_affine
    torch.manual_seed(1)
    for (B, N, C) in ((2, 4096, 512), (3, 100, 256), (1, 777, 64)):
        gn = torch.nn.GroupNorm(C // 2, C).to(cuda)
        gn.weight.data.uniform_(0.5, 1.5)
        gn.bias.data.normal_(0, 0.2)
        y = (torch.randn(B * N, C, device=cuda) * 2 + 0.5)
        scale, shift = group_norm_affine(y, B, N, gn)
        got = y.view(B, N, C) * scale.view(B, 1, C) + shift.view(B, 1, C)
        want = gn(y.view(B, N, C).transpose(1, 2)).transpose(1, 2)
        torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("impl", [1, 2])
@pytest.mark.parametrize("clouds,npts,cin,cout,affine", [(3, 256, 128, 512, False), (2, 384, 512, 256, True), (5, 128, 64, 72, True)])
def test_group_norm_statistics_fused_in_epilogue(impl, clouds, npts, cin, cout, affine, cuda):
    """captra_point_mlp_gnstats + captra_group_norm_finalize: the per-(cloud, channel) GroupNorm affine taken from
    the producing GEMM's epilogue must equal the one computed by the separate statistics pass over its output."""
    from captra_b200.mlp import PackedMLP, group_norm_affine, group_norm_finalize
    gen = torch.Generator().manual_seed(clouds * npts + cout)
    R = clouds * npts
    x = torch.randn(R, cin, generator=gen).to(cuda)
    ws, bs = _rand_mlp(cin, [cout], gen, cuda)
    gn = torch.nn.GroupNorm(cout // 2, cout).to(cuda)
    with torch.no_grad():
        gn.weight.copy_(torch.rand(cout, generator=gen) + 0.5)
        gn.bias.copy_(torch.randn(cout, generator=gen) * 0.1)
    mlp = PackedMLP(ws, bs, relu_last=False, impl=impl)
    sc = (torch.rand(clouds, cin, generator=gen) + 0.5).to(cuda) if affine else None
    sh = (torch.randn(clouds, cin, generator=gen) * 0.2).to(cuda) if affine else None
    y, stats = mlp.rows_stats(x, sc, sh, npts)
    want_y = mlp.rows_affine(x, sc, sh, npts) if affine else mlp.rows(x)
    torch.testing.assert_close(y, want_y, rtol=0, atol=0)                 # same kernel, same numbers
    s1, t1 = group_norm_finalize(stats, clouds, npts, gn)
    s0, t0 = group_norm_affine(y, clouds, npts, gn)
    torch.testing.assert_close(s1, s0, rtol=2e-5, atol=1e-6)
    torch.testing.assert_close(t1, t0, rtol=2e-5, atol=2e-6)
    ref = gn(want_y.view(clouds, npts, cout).transpose(1, 2)).transpose(1, 2).reshape(R, cout)
    torch.testing.assert_close(y * s1.repeat_interleave(npts, 0) + t1.repeat_interleave(npts, 0), ref, rtol=1e-4, atol=1e-4)
