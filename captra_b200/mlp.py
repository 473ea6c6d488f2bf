"""Host side of the fused per-point MLP kernels (csrc/mlp_simt.cu, csrc/mlp_tc.cu).

`PackedMLP` folds conv bias + eval-mode BatchNorm into (W', b'), hands them to captra_mlp_pack
once, and exposes the two execution calls of include/captra_ops.h section 3.  Internal tensors
are POINT-major ([rows, channels], channels contiguous): that is the natural A-operand layout
 and makes every gather a contiguous row read.
"""
import ctypes
import os

import torch

from . import _lib

MAX_LAYERS = 4
# 0 = exact fp32 CUDA-core kernel, 1 = tcgen05 3xTF32 kernel, 2 = tcgen05 fp16x3 kernel (default: the
# same hi/lo split arithmetic on fp16 operands -- 22 mantissa bits, half the operand bytes, K=16 per
# MMA; operands saturate at 65504 and f16_overflowed() reports if that ever happened).
# CAPTRA_MLP_IMPL overrides.
DEFAULT_IMPL = int(os.environ.get("CAPTRA_MLP_IMPL", "2"))


def f16_overflowed(reset=True):
    """True if an fp16x3 kernel met an operand >= 65504 since the last reset (device sync).  Such a
    result is not fp32-accurate: re-run with CAPTRA_MLP_IMPL=1 (3xTF32, full exponent range)."""
    rc = _lib.load().captra_f16_overflow_flag(1 if reset else 0)
    if rc > 0:
        raise _lib.CaptraError("f16_overflow_flag: " + _lib.load().captra_last_error().decode())
    return rc < 0


class MlpDesc(ctypes.Structure):
    """captra_mlp_desc (include/captra_ops.h)."""
    _fields_ = [("nlayers", ctypes.c_int), ("cin", ctypes.c_int), ("cout", ctypes.c_int * MAX_LAYERS),
                ("w", ctypes.c_void_p * MAX_LAYERS), ("bias", ctypes.c_void_p * MAX_LAYERS),
                ("relu_last", ctypes.c_int)]


def fold_conv_bn(conv, bn=None):
    """(W [cout,cin], b [cout]) of conv (kernel size 1) followed by eval-mode BatchNorm:
    y = gamma * (Wx + b - mean) / sqrt(var + eps) + beta."""
    W = conv.weight.detach().reshape(conv.out_channels, conv.in_channels).to(torch.float32)
    b = conv.bias.detach().to(torch.float32) if conv.bias is not None else torch.zeros(conv.out_channels, device=W.device)
    if bn is not None:
        g = bn.weight.detach() if bn.weight is not None else torch.ones_like(bn.running_var)
        beta = bn.bias.detach() if bn.bias is not None else torch.zeros_like(bn.running_var)
        scale = g / torch.sqrt(bn.running_var.detach() + bn.eps)
        W = W * scale[:, None]
        b = (b - bn.running_mean.detach()) * scale + beta
    return W.contiguous(), b.contiguous()


class PackedMLP:
    """A chain of up to 4 (linear + bias [+ ReLU]) layers.  Weights are packed lazily per kernel
    implementation; `impl` is the preferred one (1 = tcgen05 3xTF32) and the exact-fp32 CUDA-core
    kernel (0) serves the shapes the tensor-core kernel does not cover (layers wider than 256,
    nsample not in {32,64,128}, grouped max on dense rows).  Both are kernels of this library --
    this is shape dispatch, not a fallback off the GPU."""

    TC_GROUPS = (32, 64, 128)

    def __init__(self, weights, biases, relu_last=True, impl=None):
        assert 1 <= len(weights) <= MAX_LAYERS and len(weights) == len(biases)
        self.impl = DEFAULT_IMPL if impl is None else impl
        self.device = weights[0].device
        if self.device.type != "cuda":
            raise _lib.CaptraError("PackedMLP: weights must live on a CUDA device (no CPU path)")
        self.cin = int(weights[0].shape[1])
        self.couts = [int(w.shape[0]) for w in weights]
        self.relu_last = bool(relu_last)
        self._w = [w.detach().to(torch.float32).contiguous() for w in weights]
        self._b = [b.detach().to(torch.float32).contiguous() for b in biases]
        for l, w in enumerate(self._w):
            assert w.shape[1] == (self.cin if l == 0 else self.couts[l - 1]), "layer %d: cin mismatch" % l
        d = MlpDesc()
        d.nlayers, d.cin, d.relu_last = len(weights), self.cin, 1 if relu_last else 0
        for l in range(len(weights)):
            d.cout[l], d.w[l], d.bias[l] = self.couts[l], self._w[l].data_ptr(), self._b[l].data_ptr()
        self.desc = d
        self._packs = {}
        self._tc_ok = None
        # Chains the fused tensor-core kernel cannot hold on chip (a layer wider than 256 columns, or
        # activations + weight stages beyond 227 KB of shared memory: sa3, fp3, fp2) run layer by layer
        # on the same tcgen05 kernel, a wide layer split into 256-column chunks over grid.y; the
        # [rows, cout] intermediates are a few MB and stay in L2.
        self._layers = None
        if self.impl in (1, 2) and len(weights) > 1 and not self._tc_supported():
            n = len(weights)
            self._layers = [PackedMLP([self._w[l]], [self._b[l]], relu_last=(l < n - 1 or self.relu_last), impl=self.impl)
                            for l in range(n)]
        else:
            self._pack(self._pick(self.impl))

    def _tc_supported(self):
        if self._tc_ok is None:
            self._tc_ok = _lib.load().captra_mlp_pack_bytes(ctypes.byref(self.desc), self.impl if self.impl in (1, 2) else 1) >= 0
        return self._tc_ok

    def _pick(self, impl, group=0, sa=False):
        if impl in (1, 2) and self._tc_supported():
            if sa and (group not in self.TC_GROUPS or not self.relu_last):
                return 0
            if not sa and group and (group not in self.TC_GROUPS or not self.relu_last):
                return 0
            return impl
        return 0

    def _pack(self, impl):
        if impl not in self._packs:
            L = _lib.load()
            nbytes = L.captra_mlp_pack_bytes(ctypes.byref(self.desc), impl)
            if nbytes < 0:
                raise _lib.CaptraError("mlp_pack_bytes: " + L.captra_last_error().decode())
            buf = torch.empty((nbytes + 3) // 4, dtype=torch.float32, device=self.device)
            _lib.check(L.captra_mlp_pack(ctypes.byref(self.desc), impl, buf.data_ptr(),
                                         _lib.stream_ptr(self.device)), "mlp_pack")
            self._packs[impl] = buf
        return self._packs[impl]

    @property
    def cout(self):
        return self.couts[-1]

    def sa_max(self, xyz, new_xyz, feats, idx, out, col_off=0):
        """Fused SA scale.  xyz [B,N,3], new_xyz [B,S,3], feats [B,N,cfeat] or None,
        idx [B,S,K] int32, out [B,S,ldo] (written at columns col_off : col_off+cout)."""
        B, N, _ = xyz.shape
        S, K = idx.shape[1], idx.shape[2]
        cfeat = 0 if feats is None else feats.shape[2]
        f32, i32 = torch.float32, torch.int32
        impl = 0 if self._layers is not None else self._pick(self.impl, group=K, sa=True)
        _lib.call("sa_mlp_max[B=%d,N=%d,S=%d,K=%d,C=%d->%s,impl=%d]" % (B, N, S, K, cfeat + 3, "-".join(map(str, self.couts)), impl),
                  _lib.load().captra_sa_mlp_max, B, N, S, K, cfeat, _lib.ptr(xyz, f32, "xyz"), _lib.ptr(new_xyz, f32, "new_xyz"),
                  _lib.ptr(feats, f32, "feats") if cfeat else None, _lib.ptr(idx, i32, "idx"),
                  ctypes.byref(self.desc), self._pack(impl).data_ptr(), _lib.ptr(out, f32, "out"),
                  out.shape[-1], col_off, impl, _lib.stream_ptr(xyz.device), device=xyz.device)
        return out

    def sa_max_pre(self, xyz, new_xyz, pre, tab, idx, out, col_off=0):
        """Fused SA scale whose layer 0 was projected per point (captra_sa_mlp_max_pre): `self` holds layers
        1.. of the scale; pre [B*N, cpre] is a column block (row stride pre.stride(0)) of W_f * features,
        tab [4, cpre] = the coordinate columns of the folded first-layer weight and its bias."""
        B, N, _ = xyz.shape
        S, K = idx.shape[1], idx.shape[2]
        cpre = pre.shape[1]
        f32, i32 = torch.float32, torch.int32
        impl = self._pick(self.impl, group=K, sa=True)
        if impl not in (1, 2) or self._layers is not None:
            raise _lib.CaptraError("sa_max_pre: the projected layer 0 needs the fused tcgen05 path")
        if not (pre.is_cuda and pre.dtype == f32 and pre.dim() == 2 and pre.stride(1) == 1 and pre.shape[0] == B * N):
            raise _lib.CaptraError("sa_max_pre: pre must be a CUDA fp32 [B*N, cpre] block with contiguous channels")
        _lib.call("sa_mlp_max_pre[B=%d,N=%d,S=%d,K=%d,C=%d->%s,impl=%d]" % (B, N, S, K, cpre, "-".join(map(str, self.couts)), impl),
                  _lib.load().captra_sa_mlp_max_pre, B, N, S, K, cpre, _lib.ptr(xyz, f32, "xyz"), _lib.ptr(new_xyz, f32, "new_xyz"),
                  pre.data_ptr(), pre.stride(0), _lib.ptr(tab, f32, "tab"), _lib.ptr(idx, i32, "idx"),
                  ctypes.byref(self.desc), self._pack(impl).data_ptr(), _lib.ptr(out, f32, "out"),
                  out.shape[-1], col_off, impl, _lib.stream_ptr(xyz.device), device=xyz.device)
        return out

    def rows(self, segA, segB=None, bcast_rows=0, group=0, out=None, col_off=0):
        """Pointwise MLP.  segA [R,ca] / segB [R,cb] (or [R/bcast_rows,cb]) point-major row blocks
        (last dim contiguous, arbitrary row stride); returns out [R or R/group, cout]."""
        f32 = torch.float32

        def seg(t, name):
            if t is None or t.shape[-1] == 0:
                return None, 0, 0
            if not (t.is_cuda and t.dtype == f32 and t.dim() == 2 and t.stride(1) == 1):
                raise _lib.CaptraError("%s must be a CUDA fp32 [rows, ch] tensor with contiguous channels" % name)
            return t.data_ptr(), t.stride(0), t.shape[1]
        if self._layers is not None:
            x = self._layers[0].rows(segA, segB, bcast_rows=bcast_rows)
            for layer in self._layers[1:-1]:
                x = layer.rows(x)
            return self._layers[-1].rows(x, group=group, out=out, col_off=col_off)
        pa, lda, ca = seg(segA, "segA")
        pb, ldb, cb = seg(segB, "segB")
        R = segA.shape[0] if segA is not None else (segB.shape[0] * max(bcast_rows, 1))
        nout = R // group if group else R
        if out is None:
            out = torch.empty(nout, self.cout, dtype=f32, device=self.device)
        impl = self._pick(self.impl, group=group, sa=False)
        _lib.call("point_mlp[R=%d,C=%d->%s,g=%d,impl=%d]" % (R, self.cin, "-".join(map(str, self.couts)), group, impl),
                  _lib.load().captra_point_mlp, R, pa, lda, ca, pb, ldb, cb, bcast_rows, ctypes.byref(self.desc),
                  self._pack(impl).data_ptr(), _lib.ptr(out, f32, "out"), out.shape[-1], col_off, group, impl,
                  _lib.stream_ptr(self.device), device=self.device)
        return out

    def rows_affine(self, x, scale, shift, rows_per_cloud, out=None, col_off=0):
        """Single-segment pointwise MLP whose input is normalised on load:
        row r of cloud b = r // rows_per_cloud enters as relu(x[r] * scale[b] + shift[b])
        (GroupNorm + ReLU of the producer layer, see group_norm_affine).  tcgen05 path only."""
        f32 = torch.float32
        impl = self._pick(self.impl)
        if impl not in (1, 2) or self._layers is not None:
            raise _lib.CaptraError("rows_affine needs a chain the fused tcgen05 kernel supports")
        R, cin = x.shape
        if out is None:
            out = torch.empty(R, self.cout, dtype=f32, device=self.device)
        _lib.call("point_mlp[R=%d,C=%d->%s,g=0,impl=%d,affine]" % (R, self.cin, "-".join(map(str, self.couts)), impl),
                  _lib.load().captra_point_mlp_affine, R, _lib.ptr(x, f32, "x"), x.stride(0), cin,
                  _lib.ptr(scale, f32, "scale"), _lib.ptr(shift, f32, "shift"), rows_per_cloud,
                  ctypes.byref(self.desc), self._pack(impl).data_ptr(), _lib.ptr(out, f32, "out"), out.shape[-1], col_off, impl,
                  _lib.stream_ptr(self.device), device=self.device)
        return out


    def rows_stats(self, x, scale, shift, rows_per_cloud, out=None):
        """rows_affine (scale / shift may be None: plain input) that also returns the per-32-row-block column
        sums / sums of squares of its OUTPUT (captra_point_mlp_gnstats), which group_norm_finalize turns into the
        next GroupNorm's affine -- the activation is not read a second time for its statistics."""
        f32 = torch.float32
        impl = self._pick(self.impl)
        if impl not in (1, 2) or self._layers is not None:
            raise _lib.CaptraError("rows_stats needs a chain the fused tcgen05 kernel supports")
        R, cin = x.shape
        if out is None:
            out = torch.empty(R, self.cout, dtype=f32, device=self.device)
        stats = torch.empty(((R + 127) // 128) * 4, 2, self.cout, dtype=f32, device=self.device)
        _lib.call("point_mlp[R=%d,C=%d->%s,g=0,impl=%d,%sstats]" % (R, self.cin, "-".join(map(str, self.couts)), impl,
                                                                   "affine," if scale is not None else ""),
                  _lib.load().captra_point_mlp_gnstats, R, _lib.ptr(x, f32, "x"), x.stride(0), cin,
                  _lib.ptr(scale, f32, "scale") if scale is not None else None,
                  _lib.ptr(shift, f32, "shift") if shift is not None else None, rows_per_cloud,
                  ctypes.byref(self.desc), self._pack(impl).data_ptr(), _lib.ptr(out, f32, "out"), out.shape[-1], 0,
                  stats.data_ptr(), impl, _lib.stream_ptr(self.device), device=self.device)
        return out, stats


def group_norm_finalize(stats, clouds, npts, gn):
    """Per-(cloud, channel) scale/shift of `gn` from the block statistics written by PackedMLP.rows_stats."""
    C = stats.shape[2]
    scale = torch.empty(clouds, C, dtype=torch.float32, device=stats.device)
    shift = torch.empty_like(scale)
    f32 = torch.float32
    _lib.call("group_norm_finalize[B=%d,n=%d,C=%d]" % (clouds, npts, C), _lib.load().captra_group_norm_finalize,
              clouds, npts, C, C // gn.num_groups, stats.data_ptr(),
              gn.weight.detach().contiguous().data_ptr() if gn.weight is not None else None,
              gn.bias.detach().contiguous().data_ptr() if gn.bias is not None else None,
              float(gn.eps), scale.data_ptr(), shift.data_ptr(), _lib.stream_ptr(stats.device), device=stats.device)
    return scale, shift


def group_norm_affine(y, clouds, npts, gn):
    """Per-(cloud, channel) scale/shift equivalent to `gn` (torch.nn.GroupNorm) applied to the pre-norm
    activation y [clouds*npts, C] (point-major): GroupNorm(y) == y * scale + shift."""
    C = y.shape[1]
    cpg = C // gn.num_groups
    scale = torch.empty(clouds, C, dtype=torch.float32, device=y.device)
    shift = torch.empty_like(scale)
    g = gn.weight.detach().contiguous() if gn.weight is not None else None
    b = gn.bias.detach().contiguous() if gn.bias is not None else None
    _lib.call("group_norm_affine[B=%d,n=%d,C=%d]" % (clouds, npts, C), _lib.load().captra_group_norm_affine,
              clouds, npts, C, cpg, _lib.ptr(y, torch.float32, "y"), y.stride(0),
              g.data_ptr() if g is not None else None, b.data_ptr() if b is not None else None, float(gn.eps),
              scale.data_ptr(), shift.data_ptr(), _lib.stream_ptr(y.device), device=y.device)
    return scale, shift
