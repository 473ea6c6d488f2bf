"""CPU: the host-side algebra behind the projected SA layer 0 (captra_sa_mlp_max_pre).  The first conv of an SA
scale acts on [grouped features | grouped_xyz - centroid] (pointnet_utils.py:239-246 of the reference); splitting its
BN-folded weight into a per-point part and a per-(centroid, sample) part must reproduce the reference's
conv -> BN -> ReLU on the concatenated rows."""
import torch

from captra_b200.mlp import fold_conv_bn
from captra_b200.pointnet_utils import split_first_layer


def test_split_first_layer_equals_conv_bn_relu_on_concatenated_rows():
    gen = torch.Generator().manual_seed(0)
    D, c1, B, N, S, K = 37, 24, 2, 50, 7, 5
    conv = torch.nn.Conv2d(D + 3, c1, 1)
    bn = torch.nn.BatchNorm2d(c1).eval()
    with torch.no_grad():
        conv.weight.copy_(torch.randn(conv.weight.shape, generator=gen) * 0.3)
        conv.bias.copy_(torch.randn(c1, generator=gen) * 0.1)
        bn.weight.copy_(torch.rand(c1, generator=gen) + 0.5)
        bn.bias.copy_(torch.randn(c1, generator=gen) * 0.1)
        bn.running_mean.copy_(torch.randn(c1, generator=gen) * 0.1)
        bn.running_var.copy_(torch.rand(c1, generator=gen) + 0.5)
    feats = torch.randn(B, N, D, generator=gen)
    xyz = torch.randn(B, N, 3, generator=gen)
    ctr = torch.randn(B, S, 3, generator=gen)
    idx = torch.randint(0, N, (B, S, K), generator=gen)
    bi = torch.arange(B).view(B, 1, 1)
    g_f = feats[bi, idx]                                   # [B,S,K,D]
    g_d = xyz[bi, idx] - ctr[:, :, None, :]                # [B,S,K,3]
    with torch.no_grad():
        # the reference's formulation: channels first [B, D+3, S, K], features then centred coordinates
        rows = torch.cat([g_f, g_d], -1).permute(0, 3, 1, 2)
        want = torch.relu(bn(conv(rows))).permute(0, 2, 3, 1)
        W0, b0 = fold_conv_bn(conv, bn)
        Wf, tab = split_first_layer(W0, b0, D)
        assert Wf.shape == (c1, D) and tab.shape == (4, c1)
        P = feats @ Wf.t()                                  # once per point
        got = torch.relu(P[bi, idx] + g_d[..., 0:1] * tab[0] + g_d[..., 1:2] * tab[1] + g_d[..., 2:3] * tab[2] + tab[3])
    torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-5)
