"""Multi-prototype transductive inference with the reference's module surface
(reference models/mpti.py:18-40 BaseLearner, :45-85 and :414-577 MPTI_SelfAtten).

`forward(support_x, support_y, query_x, query_y, ...)` keeps the reference's signature and return
arity; in eval mode the whole episode — features, multi-scale degree-based noise suppression,
FPS multi-prototypes, k-NN Gaussian affinity graph, label propagation, loss — is ONE call into
libr3dfs.so (`r3dfs_mpti_forward`), with no host synchronisation inside.  `forward_episodes`
runs a batch of independent episodes through the same call.  State-dict keys are the reference's.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from .attention import SelfAttention
from .dgcnn import DGCNN


class BaseLearner(nn.Module):
    """Conv1d(bias) + BatchNorm1d stack, ReLU between layers (reference models/mpti.py:18-40)."""

    def __init__(self, in_channels, params):
        super().__init__()
        self.num_convs = len(params)
        self.convs = nn.ModuleList()
        width = in_channels
        for out_dim in params:
            self.convs.append(nn.Sequential(nn.Conv1d(width, out_dim, 1), nn.BatchNorm1d(out_dim)))
            width = out_dim

    def forward(self, x):
        if self.training:
            raise NotImplementedError("r3dfsseg_b200: stand-alone call in training mode; use "
                                      "forward(..., train=True) for the training step or .eval()")
        B, _, N = x.shape
        h = x.transpose(1, 2).reshape(B * N, -1)
        for i, seq in enumerate(self.convs):
            s, t = ops.fold_bn(seq[1], seq[0].bias)
            act = ops.ACT_RELU if i != self.num_convs - 1 else ops.ACT_NONE
            h = ops.op.linear(h.contiguous(), seq[0].weight.reshape(seq[0].weight.shape[0], -1), s, t, act)
        return h.reshape(B, N, -1).transpose(1, 2)


class MPTI_SelfAtten(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.n_way = args.n_way
        self.k_shot = args.k_shot
        self.in_channels = args.pc_in_dim
        self.n_points = args.pc_npts
        self.use_attention = args.use_attention
        self.n_subprototypes = args.n_subprototypes
        self.k_connect = args.k_connect
        self.sigma = args.sigma
        self.n_classes = self.n_way + 1
        if not self.use_attention:
            raise NotImplementedError("only use_attention=True is built (the reference's optimiser "
                                      "has no other branch, models/mpti_learner.py:26-32)")
        self.encoder = DGCNN(args.edgeconv_widths, args.dgcnn_mlp_widths, args.pc_in_dim,
                             k=args.dgcnn_k)
        self.base_learner = BaseLearner(args.dgcnn_mlp_widths[-1], args.base_widths)
        self.att_learner = SelfAttention(args.dgcnn_mlp_widths[-1], args.output_dim)
        self.feat_dim = args.edgeconv_widths[0][-1] + args.output_dim + args.base_widths[-1]
        self.shot_seed = getattr(args, "shot_seed", 1)
        self.proj = nn.Linear(self.feat_dim, 128)  # way-contrast head (train only)
        # label-propagation solver settings (the reference inverts the dense system instead)
        self.lp_alpha = 0.99
        self.cg_tol = float(getattr(args, "cg_tol", 1e-6))
        self.cg_max_iter = int(getattr(args, "cg_max_iter", 200))
        self._packed = None
        self._packed_sig = None
        self._last_diag = None
        self._flat_state = None       # training: parameters as views of one flat buffer
        self._train_ws = None
        self._train_step = 0

    # ---------------------------------------------------------------------------------------
    def _weights(self) -> ops.PackedWeights:
        if torch.compiler.is_compiling():
            # tracing (torch.compile / export): the packed tensors are graph inputs, built beforehand
            if self._packed is None:
                raise RuntimeError("call model.pack_weights() once before tracing the module")
            return self._packed
        sig = tuple((t.data_ptr(), t._version) for t in self.state_dict(keep_vars=True).values())
        if self._packed is None or sig != self._packed_sig:
            self._packed = ops.PackedWeights(self)
            self._packed_sig = sig
        return self._packed

    def pack_weights(self) -> ops.PackedWeights:
        """Folds the BatchNorm statistics into per-channel scale/shift and packs the 31 weight
        tensors the kernels read (done lazily in eager mode; call it once before torch.compile)."""
        self._packed = None
        return self._weights()

    def _cfg(self, n_query: int, mdns: bool):
        return ops.make_cfg(self.n_way, self.k_shot, n_query, self.n_points, self.n_subprototypes,
                            self.k_connect, self.sigma, self.lp_alpha, mdns, self.cg_max_iter,
                            self.cg_tol)

    def getFeatures(self, x):
        """(B, C_in, N) -> (B, 192, N)   (reference models/mpti.py:579-589)"""
        if self.training:
            raise NotImplementedError("r3dfsseg_b200: stand-alone call in training mode; use "
                                      "forward(..., train=True) for the training step or .eval()")
        pw = self._weights()
        return ops.op.features(x, pw.tensors, self.in_channels, int(self.encoder.k)).transpose(1, 2)

    # ---------------------------------------------------------------------------------------
    def forward_episodes(self, support_x, support_y, query_x, query_y=None, eval=True,
                         want_diag=False, workspace=None, support_feat=None, query_feat=None,
                         stage_events=None):
        """Batch of E independent episodes (leading dim E on every tensor).
        Returns dict(logits (E, n_query, N, n_way+1), loss (E), pred (E, n_query, N)).
        With `support_feat` (E, n_way*k_shot*N, 192) / `query_feat` (E, n_query*N, 192) given,
        the encoder is skipped and only the graph half runs on those features."""
        if self.training:
            raise NotImplementedError("r3dfsseg_b200: stand-alone call in training mode; use "
                                      "forward(..., train=True) for the training step or .eval()")
        if support_feat is None and stage_events is None and query_y is not None:
            # the registered op r3dfs::mpti_forward (traceable: torch.compile, fake tensors)
            pw = self._weights()
            r = ops.op.mpti_forward(pw.tensors, self.in_channels, int(self.encoder.k), support_x,
                                    support_y, query_x, query_y, self.n_subprototypes,
                                    self.k_connect, float(self.sigma), self.lp_alpha, bool(eval),
                                    self.cg_max_iter, self.cg_tol, workspace)
            out = {"logits": r[0], "loss": r[1], "pred": r[2]}
            if want_diag:
                out["diag"] = {"proto_count": r[3], "clean_flag": r[4], "cg_iters": r[5],
                               "cg_resid": r[6]}
            return out
        cfg = self._cfg(query_x.shape[1], mdns=bool(eval))
        return ops.mpti_forward(self._weights(), cfg, support_x, support_y, query_x, query_y,
                                want_diag=want_diag, workspace=workspace,
                                support_feat=support_feat, query_feat=query_feat,
                                stage_events=stage_events)

    def forward(self, support_x, support_y, query_x, query_y, gt_support_y=None, gt_query_y=None,
                train=False, logger=None, step=None, path=None, sampled_classes=None,
                bg_pcd_x=None, bg_pcd_y=None, support_c=None, support_flag=None, pcd_1024=None,
                label_1024=None, pcd_cutout=None, label_cutout=None, eval=False):
        """Same contract as reference models/mpti.py:414-577.
        train=False: returns (query_pred (n_query, n_way+1, N), lp_loss).
        train=True (module in .train() mode): returns the reference's 7-tuple (query_pred, lp_loss,
        contrast_loss, query_acc_LP, query_acc_original, clean_ratio_LP_avg,
        clean_ratio_original_avg); lp_loss / contrast_loss carry gradients to the parameters through
        r3dfs_mpti_train_backward.  The two clean-ratio entries are logging-only diagnostics in the
        reference (:514-552); both come from r3dfs_mpti_train_clean_ratio (NaN without gt_support_y)."""
        if train:
            if not self.training:
                raise RuntimeError("forward(train=True) needs the module in .train() mode "
                                   "(reference models/mpti_learner.py:61)")
            if support_flag is None:
                raise ValueError("forward(train=True) needs support_flag (way-contrast labels)")
            from ..train import train_episode
            query_pred, lp_loss, contrast = train_episode(self, support_x, support_y, query_x,
                                                          query_y, support_flag)
            with torch.no_grad():
                pl = query_pred.argmax(1)
                gq = gt_query_y if gt_query_y is not None else query_y
                denom = float(self.n_way * self.n_points)
                acc_lp = (pl == gq).sum().float() / denom
                acc_orig = (query_y == gq).sum().float() / denom
                ratio_lp = ratio_orig = torch.full((), float("nan"), device=query_pred.device)
                if gt_support_y is not None:  # models/mpti.py:514-552 (logging-only diagnostics)
                    from ..train import clean_ratios
                    ratio_lp, ratio_orig = clean_ratios(self, support_y, gt_support_y)
            return query_pred, lp_loss, contrast, acc_lp, acc_orig, ratio_lp, ratio_orig
        if self.training:
            raise RuntimeError("call forward(..., train=True) in training mode, or .eval() first")
        sx = support_x.reshape(self.n_way, self.k_shot, self.in_channels, self.n_points) \
            if support_x.dim() != 4 else support_x
        out = self.forward_episodes(sx.unsqueeze(0), support_y.unsqueeze(0), query_x.unsqueeze(0),
                                    query_y.unsqueeze(0), eval=eval, want_diag=True)
        self._last_diag = out["diag"]
        query_pred = out["logits"][0].transpose(1, 2)
        return query_pred, out["loss"][0]

    # the reference sets these as plain attributes inside forward; reading them here syncs
    @property
    def num_prototypes(self) -> Optional[int]:
        return None if self._last_diag is None else int(self._last_diag["proto_count"][0].sum())

    @property
    def num_nodes(self) -> Optional[int]:
        p = self.num_prototypes
        return None if p is None else p + self.n_way * self.n_points

    def computeCrossEntropyLoss(self, query_logits, query_labels):
        return F.cross_entropy(query_logits, query_labels)
