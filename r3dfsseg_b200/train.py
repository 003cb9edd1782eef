"""Meta-training step on libr3dfs.so (reference models/mpti_learner.py:16-79).

The reference trains one episode per step: `MPTI_SelfAtten.forward(train=True)` -> `loss = lp_loss +
0.1 * contrastive_loss` -> `loss.backward()` -> `Adam.step()` -> `StepLR.step()`.  Here the forward
and the backward are each ONE call into the C ABI (`r3dfs_mpti_train_forward/backward`); autograd
only sees a single `torch.autograd.Function` whose inputs are the model's parameters, so the
reference's training loop (`loss.backward(); optimizer.step()`) runs unchanged.  All trainable
tensors are views into one flat fp32 buffer (the reference's `named_parameters()` order), which is
what NCCL all-reduces for data-parallel episodes and what the fused Adam kernel updates.
PyTorch supplies memory, streams, autograd bookkeeping and torch.distributed; no torch kernel
computes anything on this path.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import _lib, ops
from ._lib import check

PARAM_NAMES: List[str] = []
for _i in range(3):
    _p = f"encoder.edge_convs.{_i}.layer."
    PARAM_NAMES += [_p + "0.weight", _p + "1.weight", _p + "1.bias", _p + "3.weight",
                    _p + "4.weight", _p + "4.bias"]
PARAM_NAMES += ["encoder.conv.layer.0.weight", "encoder.conv.layer.1.weight",
                "encoder.conv.layer.1.bias", "encoder.conv.layer.3.weight",
                "encoder.conv.layer.4.weight", "encoder.conv.layer.4.bias"]
for _i in range(2):
    _p = f"base_learner.convs.{_i}."
    PARAM_NAMES += [_p + "0.weight", _p + "0.bias", _p + "1.weight", _p + "1.bias"]
PARAM_NAMES += ["att_learner.q_map.weight", "att_learner.k_map.weight", "att_learner.v_map.weight",
                "proj.weight", "proj.bias"]
assert len(PARAM_NAMES) == _lib.N_PARAMS

BN_PREFIXES: List[str] = []
for _i in range(3):
    BN_PREFIXES += [f"encoder.edge_convs.{_i}.layer.1", f"encoder.edge_convs.{_i}.layer.4"]
BN_PREFIXES += ["encoder.conv.layer.1", "encoder.conv.layer.4", "base_learner.convs.0.1",
                "base_learner.convs.1.1"]
assert len(BN_PREFIXES) == _lib.N_BN


def param_layout(in_dim: int):
    """-> (offsets list of N_PARAMS + 1, offset of the first non-encoder tensor)."""
    off = (C.c_int64 * (_lib.N_PARAMS + 1))()
    g0 = _lib.lib().r3dfs_train_param_layout(int(in_dim), off)
    return list(off), int(g0)


def bn_layout():
    off = (C.c_int64 * (_lib.N_BN + 1))()
    _lib.lib().r3dfs_train_bn_layout(off)
    return list(off)


class FlatState:
    """The model's trainable tensors and BatchNorm running statistics re-homed as views into two
    flat device buffers (values preserved).  Safe to rebuild at any time."""

    def __init__(self, model: nn.Module):
        named = dict(model.named_parameters())
        missing = [n for n in PARAM_NAMES if n not in named]
        extra = [n for n in named if n not in PARAM_NAMES]
        if missing or extra:
            raise RuntimeError(f"unexpected parameter set: missing {missing}, extra {extra}")
        self.in_dim = int(named[PARAM_NAMES[0]].shape[1] // 2)
        self.offsets, self.group0 = param_layout(self.in_dim)
        dev = named[PARAM_NAMES[0]].device
        ops._need_cuda(named[PARAM_NAMES[0]])
        self.params = [named[n] for n in PARAM_NAMES]
        total = self.offsets[-1]
        self.flat = torch.empty(total, dtype=torch.float32, device=dev)
        for i, (p, o) in enumerate(zip(self.params, self.offsets)):
            if p.numel() != self.offsets[i + 1] - o:
                raise RuntimeError("parameter shapes differ from the reference defaults")
            view = self.flat[o:o + p.numel()].view(p.shape)
            view.copy_(p.data)
            p.data = view
        mods = dict(model.named_modules())
        self.bn = [mods[n] for n in BN_PREFIXES]
        self.bn_offsets = bn_layout()
        self.running = torch.empty(2 * self.bn_offsets[-1], dtype=torch.float32, device=dev)
        for bn, o in zip(self.bn, self.bn_offsets):
            Cn = bn.num_features
            m, v = self.running[2 * o:2 * o + Cn], self.running[2 * o + Cn:2 * o + 2 * Cn]
            m.copy_(bn.running_mean)
            v.copy_(bn.running_var)
            bn.running_mean.data = m
            bn.running_var.data = v

    def intact(self) -> bool:
        base = self.flat.data_ptr()
        for p, o in zip(self.params, self.offsets):
            if p.data_ptr() != base + 4 * o or p.dtype != torch.float32:
                return False
        rb = self.running.data_ptr()
        for bn, o in zip(self.bn, self.bn_offsets):
            if bn.running_mean.data_ptr() != rb + 8 * o:
                return False
        return True


def flat_state(model: nn.Module) -> FlatState:
    fs = getattr(model, "_flat_state", None)
    if fs is None or not fs.intact():
        fs = FlatState(model)
        model._flat_state = fs
    return fs


def dropout_mask(seed: int, shape, p: float, device) -> torch.Tensor:
    """Counter-based 0/1 keep mask (uint8) generated on the device."""
    mask = torch.empty(shape, dtype=torch.uint8, device=device)
    with torch.cuda.device(device):
        check(_lib.lib().r3dfs_dropout_mask(C.c_uint64(seed & (2 ** 64 - 1)), mask.numel(), float(p),
                                            ops._p(mask), ops._stream()), "r3dfs_dropout_mask")
    return mask


class _TrainEpisode(torch.autograd.Function):
    """(parameters) -> (query logits, lp_loss, contrast_loss); gradients come from ONE C-ABI call."""

    @staticmethod
    def forward(ctx, model, fs, cfg, sx, sy, flag, qx, qy, p_drop, keep_s, keep_q, update_running,
                *params):
        dev = sx.device
        L = _lib.lib()
        nq, N, nc = cfg.n_query, cfg.n_points, cfg.n_way + 1
        need = L.r3dfs_mpti_train_workspace(C.byref(cfg), fs.in_dim, int(model.encoder.k))
        if need == 0:
            raise _lib.R3dfsError("episode configuration not supported by the training path")
        ws = getattr(model, "_train_ws", None)
        if ws is None or ws.numel() < need or ws.device != dev:
            ws = torch.empty(need, dtype=torch.uint8, device=dev)
            model._train_ws = ws
        logits = torch.empty((nq, N, nc), dtype=torch.float32, device=dev)
        losses = torch.zeros(2, dtype=torch.float32, device=dev)
        iters = torch.zeros(1, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            check(L.r3dfs_mpti_train_forward(
                C.byref(cfg), fs.in_dim, int(model.encoder.k), ops._p(fs.flat),
                ops._p(fs.running if update_running else None),
                ops._p(sx), sx.stride(1), sx.stride(2), sx.stride(3), ops._p(sy), ops._p(flag),
                ops._p(qx), qx.stride(0), qx.stride(1), qx.stride(2), ops._p(qy),
                float(p_drop), ops._p(keep_s), ops._p(keep_q), ops._p(logits), ops._p(losses),
                ops._p(iters), ops._p(ws), ws.numel(), ops._stream()), "r3dfs_mpti_train_forward")
        ctx.model, ctx.fs, ctx.cfg = model, fs, cfg
        # every saved activation lives in the ONE shared workspace: a later training forward
        # overwrites it, so this graph's backward is only valid while the stamp still matches
        model._train_gen = int(getattr(model, "_train_gen", 0)) + 1
        ctx.gen = model._train_gen
        ctx.saved = (sy, flag, qy, keep_s, keep_q, ws)
        ctx.p_drop = float(p_drop)
        ctx.mark_non_differentiable(logits)
        model._last_cg_iters = iters
        return logits, losses[0], losses[1]

    @staticmethod
    def backward(ctx, _g_logits, g_lp, g_ct):
        model, fs, cfg = ctx.model, ctx.fs, ctx.cfg
        sy, flag, qy, keep_s, keep_q, ws = ctx.saved
        if ws is not model._train_ws or ctx.gen != model._train_gen:
            raise RuntimeError(
                "backward() of a training episode after another training forward of the same model: "
                "the saved activations live in one shared workspace and have been overwritten. "
                "Call backward() (or accumulate .grad) before the next forward(train=True).")
        # the two loss weights are host scalars of the C call (mpti_learner.py:66: 1 and 0.1).
        # Reading them waits for the forward on this stream; the learner's own train() step passes
        # them as python floats through `backward_weights` and skips the read-back.
        bw = getattr(model, "_backward_weights", None)
        if bw is not None:
            w_lp, w_ct = bw
        else:
            w_lp = 0.0 if g_lp is None else float(g_lp)
            w_ct = 0.0 if g_ct is None else float(g_ct)
        grads = torch.empty_like(fs.flat)
        with torch.cuda.device(grads.device):
            check(_lib.lib().r3dfs_mpti_train_backward(
                C.byref(cfg), fs.in_dim, int(model.encoder.k), ops._p(fs.flat), ops._p(sy),
                ops._p(flag), ops._p(qy), ctx.p_drop, ops._p(keep_s), ops._p(keep_q), w_lp, w_ct,
                ops._p(grads), ops._p(ws), ws.numel(), ops._stream()), "r3dfs_mpti_train_backward")
        model._last_grad_flat = grads
        out = [grads[o:o + p.numel()].view(p.shape) for p, o in zip(fs.params, fs.offsets)]
        return (None,) * 12 + tuple(out)


def train_episode(model, support_x, support_y, query_x, query_y, support_flag,
                  dropout_p: Optional[float] = None, keep_support: Optional[torch.Tensor] = None,
                  keep_query: Optional[torch.Tensor] = None, update_running: bool = True):
    """Training forward of one episode -> (query_pred (n_query, n_way+1, N), lp_loss, contrast_loss),
    differentiable wrt the model's parameters.  Attention dropout: pass explicit keep masks, or
    leave them None to draw them on the device from the model's step counter (p from the module)."""
    dev = ops._need_cuda(support_x, support_y, query_x, query_y, support_flag)
    fs = flat_state(model)
    n_way, k_shot, N = model.n_way, model.k_shot, model.n_points
    sx = ops._f32(support_x).reshape(n_way, k_shot, model.in_channels, N)
    if sx.stride(0) != k_shot * sx.stride(1):
        sx = sx.contiguous()
    qx = ops._f32(query_x)
    sy = support_y.to(torch.int32).contiguous()
    qy = query_y.to(torch.int64).contiguous()
    flag = support_flag.to(device=dev, dtype=torch.int32).contiguous()
    p = float(model.att_learner.dropout.p) if dropout_p is None else float(dropout_p)
    if p > 0.0 and keep_support is None:
        seed = int(torch.initial_seed()) + 7919 * int(getattr(model, "_train_step", 0))
        keep_support = dropout_mask(2 * seed, (n_way * k_shot, N, N), p, dev)
        keep_query = dropout_mask(2 * seed + 1, (qx.shape[0], N, N), p, dev)
    if p == 0.0:
        keep_support = keep_query = None
    model._train_step = int(getattr(model, "_train_step", 0)) + 1
    model._packed = None  # running statistics change: the folded eval weights are stale
    cfg = model._cfg(qx.shape[0], mdns=False)
    model._last_n_query = int(qx.shape[0])
    logits, lp, ct = _TrainEpisode.apply(model, fs, cfg, sx, sy, flag, qx, qy, p, keep_support,
                                         keep_query, update_running, *fs.params)
    if update_running:
        for bn in fs.bn:  # two getFeatures calls per episode (models/mpti.py:434-436)
            bn.num_batches_tracked += 2
    return logits.transpose(1, 2), lp, ct


def clean_ratios(model, support_y: torch.Tensor, gt_support_y: torch.Tensor):
    """(clean_ratio_LP_avg, clean_ratio_original_avg) of the model's last training forward
    (reference models/mpti.py:514-552), as 0-d device tensors."""
    fs = flat_state(model)
    ws = model._train_ws
    dev = ws.device
    cfg = model._cfg(model._last_n_query, mdns=False)
    sy = support_y.to(device=dev, dtype=torch.int32).contiguous()
    gy = gt_support_y.to(device=dev, dtype=torch.int32).contiguous()
    out = torch.empty(2, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(_lib.lib().r3dfs_mpti_train_clean_ratio(
            C.byref(cfg), fs.in_dim, int(model.encoder.k), ops._p(sy), ops._p(gy), ops._p(out),
            ops._p(ws), ws.numel(), ops._stream()), "r3dfs_mpti_train_clean_ratio")
    return out[0], out[1]


def export_decisions(model) -> Dict[str, object]:
    """Discrete decisions of the model's last training forward (r3dfs_mpti_train_export), as CPU
    tensors in the reference's own numbering — what a teacher-forced parity run needs:
      knn_support / knn_query: 3 x (B, N, k) int64;  assign: [bg, way 0, ...] local prototype ids;
      cassign: {(way, shot): ...};  I: (n_nodes, k_connect) int64 graph neighbours with nodes
      numbered [prototypes of bg, way 0, ... | query points]."""
    fs = flat_state(model)
    ws = model._train_ws
    dev = ws.device
    n_way, k_shot, N, k = model.n_way, model.k_shot, model.n_points, int(model.encoder.k)
    nq = model._last_n_query
    cfg = model._cfg(nq, mdns=False)
    Cn, S, kc = n_way * k_shot, n_way + 1, model.k_connect
    slot = model.n_subprototypes + 1
    ppad = (S * slot + 63) // 64 * 64
    nn_ = ppad + nq * N
    i32 = dict(dtype=torch.int32, device=dev)
    t = dict(knn_support=[torch.empty((Cn, N, k), **i32) for _ in range(3)],
             knn_query=[torch.empty((nq, N, k), **i32) for _ in range(3)],
             set_off=torch.empty(S, **i32), set_n=torch.empty(S, **i32),
             proto_cnt=torch.empty(S, **i32), assign=torch.empty(Cn * N, **i32),
             cloud_fg_off=torch.empty(Cn, **i32), fg_cnt=torch.empty(Cn, **i32),
             cproto_cnt=torch.empty(Cn, **i32), cassign=torch.empty(Cn * N, **i32),
             nbr=torch.empty((nn_, kc), **i32),
             valid=torch.empty(nn_, dtype=torch.uint8, device=dev))
    ex = _lib.TrainExport()
    for i in range(3):
        ex.knn_support[i] = t["knn_support"][i].data_ptr()
        ex.knn_query[i] = t["knn_query"][i].data_ptr()
    for name in ("set_off", "set_n", "proto_cnt", "assign", "cloud_fg_off", "fg_cnt", "cproto_cnt",
                 "cassign", "nbr", "valid"):
        setattr(ex, name, t[name].data_ptr())
    with torch.cuda.device(dev):
        check(_lib.lib().r3dfs_mpti_train_export(C.byref(cfg), fs.in_dim, k, C.byref(ex), ops._p(ws),
                                                 ws.numel(), ops._stream()),
              "r3dfs_mpti_train_export")
    c = {n: ([x.cpu() for x in v] if isinstance(v, list) else v.cpu()) for n, v in t.items()}
    out: Dict[str, object] = {"knn_support": [x.long() for x in c["knn_support"]],
                              "knn_query": [x.long() for x in c["knn_query"]]}
    off, num = c["set_off"].tolist(), c["set_n"].tolist()
    out["assign"] = [c["assign"][o:o + m].long() for o, m in zip(off, num)]
    foff, fnum = c["cloud_fg_off"].tolist(), c["fg_cnt"].tolist()
    out["cassign"] = {(ci // k_shot, ci % k_shot): c["cassign"][foff[ci]:foff[ci] + fnum[ci]].long()
                      for ci in range(Cn)}
    # node slots -> compact numbering [prototypes in set order | query points]
    pc = c["proto_cnt"].tolist()
    remap = torch.full((nn_,), -1, dtype=torch.long)
    run = 0
    for s_ in range(S):
        remap[s_ * slot:s_ * slot + pc[s_]] = torch.arange(run, run + pc[s_])
        run += pc[s_]
    remap[ppad:] = torch.arange(run, run + nq * N)
    rows = torch.nonzero(c["valid"].bool()).squeeze(1)
    out["I"] = remap[c["nbr"].long()[rows]]
    out["proto_cnt"] = pc
    return out


def all_reduce_gradients(flat_grad: torch.Tensor) -> float:
    """Data-parallel episodes (one per rank): sum-all-reduce the ONE flat gradient bucket in place
    (NCCL over NVLink on GPUs) and return the factor that turns the sum into the mean — the Adam
    kernel applies it, so no separate scaling pass runs.  1.0 when not distributed."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM)
        return 1.0 / dist.get_world_size()
    return 1.0


def _world_size() -> int:
    import torch.distributed as dist
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def broadcast_tensors(tensors, src: int = 0) -> None:
    import torch.distributed as dist
    for t in tensors:
        dist.broadcast(t, src)


def average_tensor(t: torch.Tensor) -> None:
    """In-place mean over the ranks (NCCL: one AVG all-reduce; gloo has no AVG: sum, then divide)."""
    import torch.distributed as dist
    if t.is_cuda:
        dist.all_reduce(t, op=dist.ReduceOp.AVG)
    else:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        t /= dist.get_world_size()


def broadcast_state(model: nn.Module, src: int = 0) -> None:
    """Data-parallel replicas must start from the same weights and BatchNorm statistics: broadcast
    the flat parameter buffer and the flat running-statistics buffer of rank `src` (no-op when not
    distributed).  Called when the optimiser is built."""
    if _world_size() > 1:
        fs = flat_state(model)
        broadcast_tensors((fs.flat, fs.running), src)
        model._packed = None


def average_running_stats(model: nn.Module) -> None:
    """Every rank updates its BatchNorm running statistics from its own episode; the replicas keep
    ONE set by averaging them (one NCCL AVG all-reduce of the 2 x 1 856-float buffer), so that eval
    and checkpoints do not depend on the rank.  The reference has no notion of replicas."""
    if _world_size() > 1:
        average_tensor(flat_state(model).running)


class FusedAdam(torch.optim.Optimizer):
    """torch.optim.Adam as configured at reference models/mpti_learner.py:26-32 (encoder lr 1e-4,
    everything else args.lr; betas (0.9, 0.999), eps 1e-8, no weight decay) as one kernel over the
    flat parameter buffer, with the data-parallel gradient mean folded in."""

    def __init__(self, model: nn.Module, lr: float = 1e-3, encoder_lr: float = 1e-4,
                 betas=(0.9, 0.999), eps: float = 1e-8):
        self.model = model
        fs = flat_state(model)
        enc = [p for n, p in zip(PARAM_NAMES, fs.params) if n.startswith("encoder.")]
        rest = [p for n, p in zip(PARAM_NAMES, fs.params) if not n.startswith("encoder.")]
        super().__init__([{"params": enc, "lr": encoder_lr}, {"params": rest, "lr": lr}],
                         dict(lr=lr, betas=betas, eps=eps))
        self.exp_avg = torch.zeros_like(fs.flat)
        self.exp_avg_sq = torch.zeros_like(fs.flat)
        self.step_count = 0
        self.sync_bn_every = 1
        broadcast_state(model)

    def _flat_grad(self, fs: FlatState) -> Optional[torch.Tensor]:
        if all(p.grad is None for p in fs.params):
            return None  # torch.optim.Adam skips parameters without a gradient
        g = getattr(self.model, "_last_grad_flat", None)
        if g is not None and all(
                p.grad is not None and p.grad.data_ptr() == g.data_ptr() + 4 * o
                for p, o in zip(fs.params, fs.offsets)):
            return g
        parts = [(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1)
                 for p in fs.params]
        return torch.cat(parts)

    # ---- checkpoints in the reference optimizer's layout (torch.optim.Adam over the four groups
    # of models/mpti_learner.py:26-32), so that a run resumes on either side ---------------------
    def state_dict(self):
        from .checkpoint import adam_state_to_reference
        fs = flat_state(self.model)
        return adam_state_to_reference(
            self.exp_avg, self.exp_avg_sq, self.step_count, PARAM_NAMES, [p.shape for p in fs.params],
            fs.offsets, (self.param_groups[0]["lr"], self.param_groups[1]["lr"]),
            self.defaults["betas"], self.defaults["eps"])

    def load_state_dict(self, state_dict):
        from .checkpoint import adam_state_from_reference
        fs = flat_state(self.model)
        m, v, step, lrs = adam_state_from_reference(state_dict, PARAM_NAMES,
                                                    [p.shape for p in fs.params], fs.offsets, fs.flat)
        self.exp_avg, self.exp_avg_sq, self.step_count = m, v, step
        self.param_groups[0]["lr"], self.param_groups[1]["lr"] = lrs

    @torch.no_grad()
    def step(self, closure=None):
        fs = flat_state(self.model)
        if self.exp_avg.data_ptr() == 0 or self.exp_avg.numel() != fs.flat.numel():
            raise RuntimeError("optimizer state does not match the model")
        g = self._flat_grad(fs)
        if g is None:
            if _world_size() > 1:
                raise RuntimeError("FusedAdam.step() without gradients on this rank would leave the "
                                   "other ranks waiting in the gradient all-reduce")
            return None
        scale = all_reduce_gradients(g)
        if self.sync_bn_every and (self.step_count + 1) % self.sync_bn_every == 0:
            average_running_stats(self.model)
        self.step_count += 1
        self.model._packed = None  # folded eval weights are stale after the update
        b1, b2 = self.defaults["betas"]
        with torch.cuda.device(g.device):
            check(_lib.lib().r3dfs_adam_step(
                ops._p(fs.flat), ops._p(g), ops._p(self.exp_avg), ops._p(self.exp_avg_sq),
                fs.flat.numel(), fs.group0, float(self.param_groups[0]["lr"]),
                float(self.param_groups[1]["lr"]), float(b1), float(b2), float(self.defaults["eps"]),
                self.step_count, float(scale), ops._stream()), "r3dfs_adam_step")
        return None


class MPTILearner_V3:
    """reference models/mpti_learner.py:16-102, same constructor / train / test contract."""

    def __init__(self, args, mode: str = "train", model: Optional[nn.Module] = None):
        from .models import MPTI_SelfAtten
        self.model = model if model is not None else MPTI_SelfAtten(args)
        self.model.cuda()
        from .checkpoint import load_model_checkpoint, load_pretrain_checkpoint
        model_ckpt = getattr(args, "model_checkpoint_path", None)
        pretrain_ckpt = getattr(args, "pretrain_checkpoint_path", None)
        if mode == "train":
            self.optimizer = FusedAdam(self.model, lr=args.lr)
            self.lr_scheduler = torch.optim.lr_scheduler.StepLR(
                self.optimizer, step_size=args.step_size, gamma=args.gamma)
            # reference models/mpti_learner.py:37-43; the only difference: with neither path given
            # the weights stay as constructed (the reference insists on a pre-trained encoder)
            if model_ckpt is None:
                if pretrain_ckpt is not None:
                    load_pretrain_checkpoint(self.model, pretrain_ckpt)
            else:
                load_model_checkpoint(self.model, model_ckpt, optimizer=self.optimizer, mode="train")
        elif mode == "test":
            if model_ckpt is not None:
                load_model_checkpoint(self.model, model_ckpt, mode="test")
        else:
            raise ValueError("Wrong GraphLearner mode (%s)! Option:train/test" % mode)

    def train(self, data, logger=None):
        [support_x, support_y, query_x, query_y, support_c, query_c, gt_support_y, gt_query_y,
         bg_pcd_x, bg_pcd_y, support_flag] = data
        self.model.train()
        out = self.model(support_x, support_y, query_x, query_y, gt_support_y=gt_support_y,
                         gt_query_y=gt_query_y, train=True, logger=logger, bg_pcd_x=bg_pcd_x,
                         bg_pcd_y=bg_pcd_y, support_c=support_c, support_flag=support_flag)
        query_logits, lp_loss, contrastive_loss = out[0], out[1], out[2]
        loss = lp_loss + 0.1 * contrastive_loss
        self.optimizer.zero_grad()
        self.model._backward_weights = (1.0, 0.1)  # d loss / d (lp_loss, contrastive_loss), known here
        try:
            loss.backward()
        finally:
            self.model._backward_weights = None
        self.optimizer.step()
        self.lr_scheduler.step()
        query_pred = query_logits.argmax(dim=1)
        correct = torch.eq(query_pred, query_y).sum().item()
        accuracy = correct / (query_y.shape[0] * query_y.shape[1])
        return (loss, lp_loss, contrastive_loss, accuracy) + tuple(out[3:])

    def test(self, data, sampled_classes=None, step=None, path=None, eval=False):
        [support_x, support_y, query_x, query_y, _, _, gt_support_y] = data
        self.model.eval()
        with torch.no_grad():
            logits, loss = self.model(support_x, support_y, query_x, query_y,
                                      gt_support_y=gt_support_y, sampled_classes=sampled_classes,
                                      step=step, path=path, support_flag=None, eval=eval)
            pred = logits.argmax(dim=1)
            correct = torch.eq(pred, query_y).sum().item()
            accuracy = correct / (query_y.shape[0] * query_y.shape[1])
        return pred, loss, accuracy
