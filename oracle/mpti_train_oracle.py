"""TEST INFRASTRUCTURE — the parity oracle of the meta-training step.  Not product code: only
tests/ and bench scripts' cpu_baseline legs may import this module.

A CPU (torch, fp32, autograd) restatement of the reference's TRAINING forward
(`MPTI_SelfAtten.forward(train=True)`, reference models/mpti.py:414-577 with
`per_way_contrast_loss` :226-313) and of one optimiser step of `MPTILearner_V3.train`
(reference models/mpti_learner.py:50-79).  Differences from the eval restatement in
mpti_oracle.py: BatchNorm uses batch statistics (support and query clouds are two separate
`getFeatures` calls, models/mpti.py:434-436, so they are normalised separately and the running
statistics are updated twice), attention dropout is applied with a caller-supplied keep mask
(the reference draws it from torch's global RNG, which cannot be shared with a CUDA kernel),
MDNS is off (`train == True` -> `pl_support_y = None`, :482), and the way-contrast loss is added.

Pinned against the reference's own modules (imported unmodified under oracle/ref_shims.py, with
`att_learner.dropout.p = 0`) by tests/test_oracle_golden.py::test_train_oracle_vs_reference and the
committed tests/golden/golden_train.pt (oracle/make_golden_train.py).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F

from . import mpti_oracle as O

EPS_BN = 1e-5
BN_MOMENTUM = 0.1


def _bn_train(x, P, prefix, running: Optional[Dict[str, torch.Tensor]]):
    """nn.BatchNorm{1,2}d in training mode: batch statistics; running stats updated in `running`
    (momentum 0.1, unbiased variance) when given."""
    rm = rv = None
    if running is not None:
        rm, rv = running[prefix + ".running_mean"], running[prefix + ".running_var"]
        running[prefix + ".num_batches_tracked"] += 1
    return F.batch_norm(x, rm, rv, P[prefix + ".weight"], P[prefix + ".bias"], True, BN_MOMENTUM,
                        EPS_BN)


def edgeconv_block_train(x, P, prefix, k, running, idx=None):
    """models/dgcnn.py:115-118, conv2d stack :45-61, BatchNorm2d over (B, N, k).
    idx: optional forced neighbour lists (B, N, k) (teacher-forced parity runs)."""
    if idx is None:
        idx = O.knn(x.detach(), k)
    e = O.get_edge_feature(x, k, idx)
    h = F.leaky_relu(_bn_train(F.conv2d(e, P[prefix + ".layer.0.weight"]), P, prefix + ".layer.1",
                               running), 0.2)
    h = F.leaky_relu(_bn_train(F.conv2d(h, P[prefix + ".layer.3.weight"]), P, prefix + ".layer.4",
                               running), 0.2)
    return h.max(dim=-1)[0]


def get_features_train(x, P, running=None, keep_mask=None, dropout_p=0.0, dgcnn_k=20, knn_idx=None):
    """models/mpti.py:579-589 in training mode.  keep_mask: (B, N, N) 0/1 attention-dropout keep
    mask (None = no dropout); kept entries are scaled by 1 / (1 - dropout_p)."""
    outs = []
    h = x
    for i in range(3):
        h = edgeconv_block_train(h, P, f"encoder.edge_convs.{i}", dgcnn_k, running,
                                 None if knn_idx is None else knn_idx[i])
        outs.append(h)
    h = torch.cat(outs, dim=1)
    h = F.leaky_relu(_bn_train(F.conv1d(h, P["encoder.conv.layer.0.weight"]), P,
                               "encoder.conv.layer.1", running), 0.2)
    l2 = F.leaky_relu(_bn_train(F.conv1d(h, P["encoder.conv.layer.3.weight"]), P,
                                "encoder.conv.layer.4", running), 0.2)
    # BaseLearner (models/mpti.py:35-40)
    b = F.relu(_bn_train(F.conv1d(l2, P["base_learner.convs.0.0.weight"],
                                  P["base_learner.convs.0.0.bias"]), P, "base_learner.convs.0.1",
                         running))
    b = _bn_train(F.conv1d(b, P["base_learner.convs.1.0.weight"], P["base_learner.convs.1.0.bias"]),
                  P, "base_learner.convs.1.1", running)
    # SelfAttention (models/attention.py:39-48) with dropout on the attention map
    q = F.conv1d(l2, P["att_learner.q_map.weight"])
    kk = F.conv1d(l2, P["att_learner.k_map.weight"])
    v = F.conv1d(l2, P["att_learner.v_map.weight"])
    attn = F.softmax(torch.matmul(q.transpose(1, 2) / (q.shape[1] ** 0.5), kk), dim=-1)
    if keep_mask is not None:
        attn = attn * keep_mask.to(attn.dtype) / (1.0 - dropout_p)
    a = torch.matmul(attn, v.transpose(1, 2)).transpose(1, 2)
    return torch.cat((outs[0], a, b), dim=1)


def multi_prototypes_train(feat, k, assign=None):
    """models/mpti.py:597-634 with gradients flowing through the member means only.
    assign: optional forced cluster assignment (n,) (teacher-forced parity runs)."""
    n = feat.shape[0]
    if assign is not None:
        assign = assign.long()
        m = int(assign.max()) + 1
        protos = torch.stack([feat[torch.nonzero(assign == i).squeeze(1)].mean(0) for i in range(m)])
        return protos, assign, m, None
    if k / n < 1:
        fd = feat.detach()
        seeds = O.fps(fd, O.fps_count(n, k)).unique()
        m = len(seeds)
        far = fd[seeds]
        dist = torch.stack([torch.norm(fd - far[j] + 1e-6, 2.0, 1) for j in range(m)], dim=1)
        assign = torch.argmin(dist, dim=1)
        protos = torch.stack([feat[torch.nonzero(assign == i).squeeze(1)].mean(0) for i in range(m)])
        return protos, assign, m, seeds
    return feat, torch.arange(n), n, torch.arange(n)


def per_way_contrast_loss(support_feat, support_y, support_flag, P, n_way, k_shot, fps_k=4,
                          temp=0.1, cassign=None):
    """models/mpti.py:226-313.  cassign: optional {(way, shot): forced assignment}."""
    clean = bool(support_flag[0, 0] * k_shot == torch.sum(support_flag[0]))

    def shot_protos(way, k):
        fg = support_y[way, k] == 1
        f = support_feat[way, k][:, fg].transpose(1, 0)
        p, _, _, _ = multi_prototypes_train(f, fps_k, None if cassign is None else cassign[(way, k)])
        return F.normalize(F.linear(p, P["proj.weight"], P["proj.bias"]), p=2, dim=1)

    total = []
    for way in range(n_way):
        feats, labels = [], []
        for k in range(k_shot):
            z = shot_protos(way, k)
            feats.append(z)
            labels.append(torch.zeros(z.shape[0]) + float(support_flag[way, k]))
        if clean:
            other = way + 1 if way < n_way - 1 else 0
            for k in range(2):
                z = shot_protos(other, k)
                feats.append(z)
                labels.append(torch.zeros(z.shape[0]) - 1.0)
        f = torch.cat(feats, 0)
        lab = torch.cat(labels, 0)
        lmask = 1.0 - torch.eye(lab.shape[0])
        gmask = torch.eq(lab[:, None], lab[None, :]).float() * lmask
        logits = torch.matmul(f, f.t()) / temp
        exp_logits = torch.exp(logits) * lmask
        log_prob = logits - torch.log(exp_logits.sum(1, keepdim=True))
        mlpp = (gmask * log_prob).sum(1) / gmask.sum(1)
        total.append((-mlpp).mean())
    return sum(total) / len(total)


def forward_train(P: Dict[str, torch.Tensor], support_x, support_y, query_x, query_y, support_flag,
                  running=None, keep_mask_support=None, keep_mask_query=None, dropout_p=0.0,
                  n_subprototypes=100, k_connect=200, sigma=1.0, dgcnn_k=20, keep=False,
                  forced=None):
    """models/mpti.py:414-577 with train=True -> dict(query_pred, lp_loss, contrast_loss).
    forced: optional dict of discrete decisions taken from another implementation (they carry no
    gradient): knn_support / knn_query (3 x (B, N, k)), assign (list per set: bg, way 0, ...),
    cassign ({(way, shot): (n_fg,)}), I ((n, k_connect) graph neighbours)."""
    fz = forced or {}
    n_way, k_shot = support_y.shape[:2]
    N = support_y.shape[-1]
    n_cls = n_way + 1
    sx = support_x.reshape(n_way * k_shot, -1, N)
    sf = get_features_train(sx, P, running, keep_mask_support, dropout_p, dgcnn_k,
                            fz.get("knn_support"))
    D = sf.shape[1]
    support_feat = sf.reshape(n_way, k_shot, D, N)
    qf = get_features_train(query_x, P, running, keep_mask_query, dropout_p, dgcnn_k,
                            fz.get("knn_query"))
    query_feat = qf.transpose(1, 2).contiguous().view(-1, D)
    contrast = per_way_contrast_loss(support_feat, support_y, support_flag, P, n_way, k_shot,
                                     cassign=fz.get("cassign"))
    protos, labels = [], []
    for i in range(n_way):
        f = support_feat[i].transpose(1, 2).contiguous().view(-1, D)
        f = f[torch.nonzero(support_y[i].reshape(-1)).squeeze(1)]
        p, _, _, _ = multi_prototypes_train(f, n_subprototypes,
                                            fz["assign"][i + 1] if "assign" in fz else None)
        lab = torch.zeros(p.shape[0], n_cls)
        lab[:, i + 1] = 1
        protos.append(p)
        labels.append(lab)
    fb = support_feat.transpose(2, 3).contiguous().view(-1, D)
    fb = fb[torch.nonzero(torch.logical_not(support_y).reshape(-1)).squeeze(1)]
    pb, _, _, _ = multi_prototypes_train(fb, n_subprototypes, fz["assign"][0] if "assign" in fz else None)
    lb = torch.zeros(pb.shape[0], n_cls)
    lb[:, 0] = 1
    prototypes = torch.cat([pb] + protos, 0)
    Pn = prototypes.shape[0]
    Y = torch.zeros(Pn + query_feat.shape[0], n_cls)
    Y[:Pn] = torch.cat([lb] + labels, 0)
    node_feat = torch.cat((prototypes, query_feat), 0)
    if "I" in fz:
        I = fz["I"].long()
    else:
        I, _ = O.knn_graph_exact(node_feat.detach(), k_connect)
    A, _, _ = O.affinity_dense(node_feat, k_connect, sigma, I)
    Z = O.label_propagate_dense(A, Y)
    query_pred = Z[Pn:].view(-1, N, n_cls).transpose(1, 2)
    lp_loss = F.cross_entropy(query_pred, query_y)
    out = dict(query_pred=query_pred, lp_loss=lp_loss, contrast_loss=contrast, num_prototypes=Pn)
    if keep:
        out.update(support_feat=support_feat, query_feat=query_feat, node_feat=node_feat, Z=Z, I=I)
    return out


ENCODER_LR = 1e-4


def param_groups(P: Dict[str, torch.Tensor], lr: float):
    """models/mpti_learner.py:26-32: encoder at 1e-4, everything else at args.lr."""
    enc = [v for k, v in P.items() if k.startswith("encoder.")]
    rest = [v for k, v in P.items() if not k.startswith("encoder.")]
    return [{"params": enc, "lr": ENCODER_LR}, {"params": rest, "lr": lr}]


def split_state_dict(sd: Dict[str, torch.Tensor]):
    """-> (trainable parameters (requires_grad clones), BN running buffers (clones))."""
    P, running = {}, {}
    for k, v in sd.items():
        if k.endswith("running_mean") or k.endswith("running_var") or k.endswith("num_batches_tracked"):
            running[k] = v.clone()
        else:
            P[k] = v.clone().requires_grad_(True)
    return P, running
