"""DGCNN encoder with the reference's module surface (reference models/dgcnn.py:17-127), backed by
libr3dfs.so.  Parameter / buffer names are the reference's (`edge_convs.{i}.layer.{0,1,3,4}.*`,
`conv.layer.{0,1,3,4}.*`) so its checkpoints load unchanged.  Stand-alone use is eval mode (BatchNorm
folded into the kernels' per-channel affine); training with batch statistics runs per episode through
`MPTI_SelfAtten.forward(train=True)` (r3dfsseg_b200/train.py), which reads these parameters.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from ..ops import get_edge_feature, get_graph_feature, knn  # noqa: F401  (re-exported)


def _no_training(mod: nn.Module) -> None:
    if mod.training:
        raise NotImplementedError(
            "r3dfsseg_b200: this module has no stand-alone training-mode forward; the training "
            "step runs whole episodes through MPTI_SelfAtten.forward(..., train=True) — call "
            ".eval() to use it on its own (there is deliberately no PyTorch fallback)")


class _PointwiseStack(nn.Module):
    """1x1 conv (+BN) (+LeakyReLU 0.2) stack; `rank` 1 -> Conv1d/BatchNorm1d, 2 -> Conv2d/BatchNorm2d."""

    def __init__(self, rank, in_feat, layer_dims, batch_norm=True, relu=True, bias=False):
        super().__init__()
        conv = nn.Conv1d if rank == 1 else nn.Conv2d
        norm = nn.BatchNorm1d if rank == 1 else nn.BatchNorm2d
        self.layer_dims = layer_dims
        mods = []
        width = in_feat
        for out_dim in layer_dims:
            mods.append(conv(width, out_dim, kernel_size=1, bias=bias))
            if batch_norm:
                mods.append(norm(out_dim))
            if relu:
                mods.append(nn.LeakyReLU(0.2))
            width = out_dim
        self.layer = nn.Sequential(*mods)

    def _stages(self):
        """[(conv, bn or None, act code)] in order."""
        out, mods, i = [], list(self.layer), 0
        while i < len(mods):
            cv, bn, act = mods[i], None, ops.ACT_NONE
            i += 1
            if i < len(mods) and isinstance(mods[i], nn.modules.batchnorm._BatchNorm):
                bn = mods[i]
                i += 1
            if i < len(mods) and isinstance(mods[i], nn.LeakyReLU):
                act = ops.ACT_LRELU
                i += 1
            out.append((cv, bn, act))
        return out

    def forward(self, x):
        _no_training(self)
        lead, spatial = x.shape[0], x.shape[2:]
        h = x.reshape(lead, x.shape[1], -1).transpose(1, 2).reshape(-1, x.shape[1])  # point-major
        for cv, bn, act in self._stages():
            w = cv.weight.reshape(cv.weight.shape[0], -1)
            if bn is not None:
                s, t = ops.fold_bn(bn, cv.bias)
            else:
                s = None
                t = cv.bias.detach().float().contiguous() if cv.bias is not None else None
            h = ops.op.linear(h.contiguous(), w, s, t, act)
        return h.reshape(lead, -1, h.shape[1]).transpose(1, 2).reshape(lead, h.shape[1], *spatial)


class conv2d(_PointwiseStack):
    """reference models/dgcnn.py:45-61"""

    def __init__(self, in_feat, layer_dims, batch_norm=True, relu=True, bias=False):
        super().__init__(2, in_feat, layer_dims, batch_norm, relu, bias)


class conv1d(_PointwiseStack):
    """reference models/dgcnn.py:64-80"""

    def __init__(self, in_feat, layer_dims, batch_norm=True, relu=True, bias=False):
        super().__init__(1, in_feat, layer_dims, batch_norm, relu, bias)


class DGCNN(nn.Module):
    """Stacked EdgeConv blocks + point MLP (reference models/dgcnn.py:83-127).
    forward(x (B, nfeat, N)) -> (first EdgeConv output (B, 64, N), MLP output (B, mlp_widths[-1], N)),
    or (list of EdgeConv outputs, MLP output) with return_edgeconvs=True."""

    def __init__(self, edgeconv_widths, mlp_widths, nfeat, k=20, return_edgeconvs=False):
        super().__init__()
        self.n_edgeconv = len(edgeconv_widths)
        self.k = k
        self.return_edgeconvs = return_edgeconvs
        self.edge_convs = nn.ModuleList()
        width = nfeat
        for widths in edgeconv_widths:
            self.edge_convs.append(conv2d(2 * width, widths))
            width = widths[-1]
        self.conv = conv1d(sum(w[-1] for w in edgeconv_widths), mlp_widths)

    def forward(self, x):
        _no_training(self)
        outs = []
        for blk in self.edge_convs:
            stages = blk._stages()
            if len(stages) != 2 or stages[0][0].weight.shape[0] != 64 or stages[1][0].weight.shape[0] != 64:
                raise NotImplementedError("fused EdgeConv is built for width-[64, 64] blocks")
            (c1, b1, _), (c2, b2, _) = stages
            s1, t1 = ops.fold_bn(b1)
            s2, t2 = ops.fold_bn(b2)
            x = ops.op.edgeconv(x, c1.weight, s1, t1, c2.weight, s2, t2, self.k).transpose(1, 2)
            outs.append(x)
        out = self.conv(torch.cat(outs, dim=1))
        if self.return_edgeconvs:
            return outs, out
        return outs[0], out
