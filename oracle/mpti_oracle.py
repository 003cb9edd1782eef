"""TEST INFRASTRUCTURE — the parity oracle.  Not product code: only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this module.

A CPU (torch, fp32) restatement of the reference's episode hot path, written from the reference's
algorithm and following its computational form (materialised (B,N,N) distance matrices, the
(B,2C,N,k) edge tensor, dense (n,n) affinity, dense torch.inverse) so that timing it is a fair
stand-in for "the reference's CPU path".  Each function cites the reference lines it follows.

Pinning: the reference has no tests or golden vectors, and three of its dependencies
(faiss, torch_cluster, torch<=1.8 pairwise_distance) are absent/unpinned, so their semantics are
fixed by oracle/ref_shims.py.  This restatement is pinned against the reference's own modules
imported unmodified under those shims: tests/test_oracle_golden.py compares it with tests/golden/*.pt,
which oracle/make_golden.py / make_golden_train.py generate from the reference (build container only).

Third-party arithmetic restated here (not under /root/reference, unpinned there):
  * faiss.IndexFlatL2.search — exact squared-L2 k-NN, ascending, query itself in column 0;
  * torch_cluster.fps — start index 0, dist = min(dist, |x - x_last|^2), argmax (first maximum),
    m = ceil(fp32(n) * fp32(ratio)) samples;
  * F.pairwise_distance (torch 1.8) — norm(x1 - x2 + 1e-6, p, dim=1).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F

EPS_BN = 1e-5


# ------------------------------------------------------------------------------------------------
# DGCNN  (reference models/dgcnn.py)
# ------------------------------------------------------------------------------------------------
def knn(x: torch.Tensor, k: int) -> torch.Tensor:
    """models/dgcnn.py:17-23"""
    inner = -2 * torch.matmul(x.transpose(2, 1), x)
    xx = torch.sum(x ** 2, dim=1, keepdim=True)
    pd = -xx - inner - xx.transpose(2, 1)
    return pd.topk(k=k, dim=-1)[1]


def knn_scores(x: torch.Tensor) -> torch.Tensor:
    """The ranking key of `knn` in float64 (for tie-aware comparisons)."""
    x = x.double()
    inner = -2 * torch.matmul(x.transpose(2, 1), x)
    xx = torch.sum(x ** 2, dim=1, keepdim=True)
    return -xx - inner - xx.transpose(2, 1)


def get_edge_feature(x: torch.Tensor, K: int = 20, idx: Optional[torch.Tensor] = None) -> torch.Tensor:
    """models/dgcnn.py:26-42"""
    B, C, N = x.shape
    if idx is None:
        idx = knn(x, K)
    central = x.unsqueeze(-1).expand(-1, -1, -1, K)
    gidx = idx.unsqueeze(1).expand(-1, C, -1, -1).reshape(B, C, N * K)
    nbr = torch.gather(x, 2, gidx).view(B, C, N, K)
    return torch.cat((nbr - central, central), dim=1)


def _bn(x, sd, prefix):
    return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"],
                        sd[prefix + ".weight"], sd[prefix + ".bias"], False, 0.0, EPS_BN)


def edgeconv_block(x, sd, prefix, k, idx=None):
    """models/dgcnn.py:115-118 with the conv2d stack of :45-61 (eval BN)."""
    e = get_edge_feature(x, k, idx)
    h = F.leaky_relu(_bn(F.conv2d(e, sd[prefix + ".layer.0.weight"]), sd, prefix + ".layer.1"), 0.2)
    h = F.leaky_relu(_bn(F.conv2d(h, sd[prefix + ".layer.3.weight"]), sd, prefix + ".layer.4"), 0.2)
    return h.max(dim=-1)[0]


def dgcnn_forward(x, sd, prefix="encoder.", k=20, n_edgeconv=3):
    """models/dgcnn.py:113-127 -> (edgeconv_outputs[0], out)"""
    outs = []
    for i in range(n_edgeconv):
        x = edgeconv_block(x, sd, f"{prefix}edge_convs.{i}", k)
        outs.append(x)
    h = torch.cat(outs, dim=1)
    h = F.leaky_relu(_bn(F.conv1d(h, sd[prefix + "conv.layer.0.weight"]), sd, prefix + "conv.layer.1"), 0.2)
    h = F.leaky_relu(_bn(F.conv1d(h, sd[prefix + "conv.layer.3.weight"]), sd, prefix + "conv.layer.4"), 0.2)
    return outs[0], h


def base_learner(x, sd, prefix="base_learner."):
    """models/mpti.py:35-40"""
    n = len({k.split(".")[2] for k in sd if k.startswith(prefix + "convs.")})
    for i in range(n):
        p = f"{prefix}convs.{i}"
        x = _bn(F.conv1d(x, sd[p + ".0.weight"], sd[p + ".0.bias"]), sd, p + ".1")
        if i != n - 1:
            x = F.relu(x)
    return x


def self_attention(x, sd, prefix="att_learner."):
    """models/attention.py:39-48 (eval: dropout is the identity)"""
    q = F.conv1d(x, sd[prefix + "q_map.weight"])
    k = F.conv1d(x, sd[prefix + "k_map.weight"])
    v = F.conv1d(x, sd[prefix + "v_map.weight"])
    temperature = q.shape[1] ** 0.5
    attn = F.softmax(torch.matmul(q.transpose(1, 2) / temperature, k), dim=-1)
    return torch.matmul(attn, v.transpose(1, 2)).transpose(1, 2)


def get_features(x, sd, dgcnn_k=20):
    """models/mpti.py:579-589 (use_attention=True)"""
    l1, l2 = dgcnn_forward(x, sd, k=dgcnn_k)
    return torch.cat((l1, self_attention(l2, sd), base_learner(l2, sd)), dim=1)


# ------------------------------------------------------------------------------------------------
# multi-prototypes  (reference models/mpti.py:597-715)
# ------------------------------------------------------------------------------------------------
def fps_count(n: int, k: int) -> int:
    ratio = torch.tensor(k / n, dtype=torch.float32)
    return int(torch.ceil(torch.tensor(float(n), dtype=torch.float32) * ratio).item())


def fps(feat: torch.Tensor, m: int) -> torch.Tensor:
    """torch_cluster.fps(random_start=False): picks in selection order."""
    out = torch.empty(m, dtype=torch.long)
    out[0] = 0
    dist = (feat - feat[0]).pow(2).sum(1)
    for i in range(1, m):
        a = int(dist.argmax())
        out[i] = a
        dist = torch.min(dist, (feat - feat[a]).pow(2).sum(1))
    return out


def pairwise_distance_t18(x1, x2, p=2.0, eps=1e-6):
    return torch.norm(x1 - x2 + eps, p, 1)


def multi_prototypes(feat: torch.Tensor, k: int):
    """models/mpti.py:597-634 -> (prototypes, assignments, num, seed_index)"""
    n = feat.shape[0]
    if k / n < 1:
        seeds = fps(feat, fps_count(n, k)).unique()
        m = len(seeds)
        far = feat[seeds]
        # same values as the reference's (n,192,m) broadcast, evaluated seed by seed to bound memory
        dist = torch.stack([torch.norm(feat - far[j] + 1e-6, 2.0, 1) for j in range(m)], dim=1)
        assign = torch.argmin(dist, dim=1)
        protos = feat.new_zeros((m, feat.shape[1]))
        for i in range(m):
            protos[i] = feat[torch.nonzero(assign == i).squeeze(1)].mean(0)
        return protos, assign, m, seeds
    return feat, torch.arange(n), n, torch.arange(n)


# ------------------------------------------------------------------------------------------------
# multi-scale degree-based noise suppression  (reference models/mpti.py:87-223, 316-371)
# ------------------------------------------------------------------------------------------------
def grid_sampling(spatial, feat, n_x, n_y, n_z):
    """models/mpti.py:316-371"""
    mins = [torch.min(spatial[:, a]) for a in range(3)]
    maxs = [torch.max(spatial[:, a]) for a in range(3)]
    ns = [n_x, n_y, n_z]
    ds = [(maxs[a] - mins[a]) / ns[a] for a in range(3)]
    starts = [[mins[a] + i * ds[a] for i in range(ns[a])] for a in range(3)]
    seeds = []
    assign = torch.zeros(spatial.shape[0], dtype=torch.long)
    count = 0
    for x in starts[0]:
        xm = (spatial[:, 0] >= x) * (spatial[:, 0] <= x + ds[0])
        for y in starts[1]:
            ym = (spatial[:, 1] >= y) * (spatial[:, 1] <= y + ds[1])
            for z in starts[2]:
                zm = (spatial[:, 2] >= z) * (spatial[:, 2] <= z + ds[2])
                mask = xm * ym * zm
                if torch.sum(mask) > 0:
                    seeds.append(torch.mean(feat[mask], dim=0, keepdim=True))
                    assign[mask] = count
                    count += 1
    return torch.cat(seeds, dim=0), assign, count


def mdns_flags_one_scale(support_feat, support_y, support_x, n_x, n_y, n_z, internals=None):
    """models/mpti.py:87-176 -> flag (n_way, k_shot).  `internals` (a list) receives, per way, the
    degree vector and the per-shot grid_sampling results (tests of the CUDA kernels' internals)."""
    n_way, k_shot = support_y.shape[:2]
    flag = torch.zeros((n_way, k_shot))
    for way in range(n_way):
        seeds, lens = [], []
        grids = []
        for k in range(k_shot):
            fg = support_y[way, k] == 1
            f = support_feat[way, k][:, fg].transpose(1, 0)
            sp = support_x[way, k][:, fg].transpose(1, 0)
            s, a, n = grid_sampling(sp, f, n_x, n_y, n_z)
            seeds.append(s)
            lens.append(n)
            grids.append((s, a, n))
        sn = F.normalize(torch.cat(seeds, dim=0), p=2, dim=1)
        cos = torch.mm(sn, sn.t()) * (1.0 - torch.eye(sn.shape[0], dtype=sn.dtype))
        if n_x == 1 and n_y == 1 and n_z == 1:
            cos = cos.pow(3)
        deg = cos.sum(1)
        if internals is not None:
            internals.append(dict(degree=deg, grids=grids))
        mask = deg > deg.mean()
        c = 0
        for k in range(k_shot):
            flag[way, k] = 1.0 if torch.mean(mask[c:c + lens[k]].float()) > 0.5 else 0.0
            c += lens[k]
    return flag


def mdns_multi_scale(support_feat, support_y, support_x):
    """models/mpti.py:178-223 -> (pl_support_y list, clean_flag (n_way, k_shot))"""
    flags = [mdns_flags_one_scale(support_feat, support_y, support_x, *s) for s in ((1, 1, 1), (2, 2, 1))]
    total = torch.stack(flags, 0).mean(0)
    n_way, k_shot = support_y.shape[:2]
    clean = torch.ones((n_way, k_shot))
    pl = []
    for way in range(n_way):
        parts = []
        for k in range(k_shot):
            y = support_y[way, k][support_y[way, k] > 0]
            if total[way, k] < 0.5:
                y = torch.zeros_like(y)
                clean[way, k] = 0
            parts.append(y)
        wy = torch.cat(parts)
        if torch.sum(wy) == 0:
            wy = torch.ones_like(wy)
            clean[way] = 1
        pl.append(wy)
    return pl, clean


# ------------------------------------------------------------------------------------------------
# affinity + label propagation  (reference models/mpti.py:717-776)
# ------------------------------------------------------------------------------------------------
def knn_graph_exact(X: torch.Tensor, k: int):
    """faiss.IndexFlatL2 search of k+1 with the query itself in column 0, column 0 dropped
    (models/mpti.py:733-736).  float64 distances, ties -> lowest index.  Returns (I, d2 float64)."""
    Xd = X.double()
    sq = (Xd * Xd).sum(1)
    d2 = sq[:, None] + sq[None, :] - 2.0 * (Xd @ Xd.t())
    d2.fill_diagonal_(-1.0)
    order = torch.argsort(d2, dim=1, stable=True)[:, :k + 1]
    return order[:, 1:], d2


def knn_graph_fp32_topk(X: torch.Tensor, k: int) -> torch.Tensor:
    """Speed stand-in for the faiss search in the TIMING arm of bench.py only (cpu_baseline /
    --impl reference): FP32 Gram form + topk, what faiss.IndexFlatL2 does with BLAS, instead of the
    parity definition above (FP64 distances + a full stable argsort of n^2 keys)."""
    sq = (X * X).sum(1)
    d2 = sq[:, None] + sq[None, :] - 2.0 * (X @ X.t())
    d2.fill_diagonal_(-1.0)
    return d2.topk(k + 1, dim=1, largest=False)[1][:, 1:]


def affinity_dense(node_feat: torch.Tensor, k: int, sigma: float, I: Optional[torch.Tensor] = None):
    """models/mpti.py:717-756 -> dense A (n, n)"""
    n, D = node_feat.shape
    if I is None:
        I, _ = knn_graph_exact(node_feat, k)
    sim = node_feat.new_empty((n, k))
    for s in range(0, n, 512):  # same values as the reference's (n,k,D) gather, in row chunks
        nb = node_feat[I[s:s + 512]]                                   # (c, k, D)
        dist = torch.norm(node_feat[s:s + 512, None, :] - nb + 1e-6, 2.0, 2)
        sim[s:s + 512] = torch.exp(-0.5 * (dist / sigma) ** 2)
    A = node_feat.new_zeros((n, n)).scatter_(1, I, sim)
    A = A + A.t()
    A = A * (1 - torch.eye(n, dtype=A.dtype))
    return A, I, sim


def label_propagate_dense(A: torch.Tensor, Y: torch.Tensor, alpha: float = 0.99, dtype=None):
    """models/mpti.py:758-776 (dense inverse).  dtype=float64 gives the high-precision answer;
    default: the dtype of A."""
    eps = np.finfo(float).eps
    dtype = A.dtype if dtype is None else dtype
    A = A.to(dtype)
    D = A.sum(1)
    Dsi = torch.diag_embed(torch.sqrt(1.0 / (D + eps)))
    S = Dsi @ A @ Dsi
    n = A.shape[0]
    return torch.inverse(torch.eye(n, dtype=dtype) - alpha * S + eps) @ Y.to(dtype)


# ------------------------------------------------------------------------------------------------
# the episode  (reference models/mpti.py:414-577, train=False)
# ------------------------------------------------------------------------------------------------
def forward_episode(sd: Dict[str, torch.Tensor], support_x, support_y, query_x, query_y,
                    n_subprototypes=100, k_connect=200, sigma=1.0, dgcnn_k=20, eval_mdns=True,
                    keep: bool = False, support_feat=None, query_feat=None,
                    timing_knn: bool = False) -> Dict[str, object]:
    """`support_feat` (n_way*k_shot, D, N) / `query_feat` (n_query, D, N), when given, replace the
    getFeatures calls (used to check the graph half on features produced elsewhere).
    timing_knn: FP32 topk stand-in for faiss (bench.py's CPU timing arm; never in parity tests)."""
    n_way, k_shot = support_y.shape[:2]
    N = support_y.shape[-1]
    n_cls = n_way + 1
    sx = support_x.reshape(n_way * k_shot, -1, N)
    if support_feat is None:
        support_feat = get_features(sx, sd, dgcnn_k)
    D = support_feat.shape[1]
    support_feat = support_feat.reshape(n_way, k_shot, D, N)
    if query_feat is None:
        query_feat = get_features(query_x, sd, dgcnn_k)
    query_feat = query_feat.transpose(1, 2).contiguous().view(-1, D)
    out: Dict[str, object] = {}
    pl, clean = None, None
    if eval_mdns:
        pl, clean = mdns_multi_scale(support_feat, support_y, support_x.reshape(n_way, k_shot, -1, N))
    # foreground prototypes (models/mpti.py:636-688)
    protos, labels, counts, seed_idx, sets = [], [], [], [], []
    for i in range(n_way):
        f = support_feat[i].transpose(1, 2).contiguous().view(-1, D)
        f = f[torch.nonzero(support_y[i].reshape(-1)).squeeze(1)]
        if pl is not None:
            f = f[pl[i] == 1]
        p, _, m, s = multi_prototypes(f, n_subprototypes)
        lab = p.new_zeros(p.shape[0], n_cls)
        lab[:, i + 1] = 1
        protos.append(p); labels.append(lab); counts.append(m); seed_idx.append(s); sets.append(f)
    # background prototypes (models/mpti.py:690-715)
    fb = support_feat.transpose(2, 3).contiguous().view(-1, D)
    fb = fb[torch.nonzero(torch.logical_not(support_y).reshape(-1)).squeeze(1)]
    pb, _, mb, sb = multi_prototypes(fb, n_subprototypes)
    lb = pb.new_zeros(pb.shape[0], n_cls)
    lb[:, 0] = 1
    prototypes = torch.cat([pb] + protos, 0)
    proto_labels = torch.cat([lb] + labels, 0)
    P = prototypes.shape[0]
    n = P + query_feat.shape[0]
    Y = prototypes.new_zeros(n, n_cls)
    Y[:P] = proto_labels
    node_feat = torch.cat((prototypes, query_feat), 0)
    A, I, sim = affinity_dense(node_feat, k_connect, sigma,
                               I=knn_graph_fp32_topk(node_feat, k_connect) if timing_knn else None)
    Z = label_propagate_dense(A, Y)
    query_pred = Z[P:].view(-1, N, n_cls).transpose(1, 2)
    loss = F.cross_entropy(query_pred, query_y) if query_y is not None else None
    out.update(query_pred=query_pred, loss=loss, pred=query_pred.argmax(1), clean_flag=clean,
               proto_count=[mb] + counts, num_prototypes=P)
    if keep:
        out.update(support_feat=support_feat, query_feat=query_feat, node_feat=node_feat, A=A, I=I,
                   sim=sim, Y=Y, Z=Z, seed_idx=[sb] + seed_idx, sets=[fb] + sets)
    return out


def forward_episode_fp64(sd, support_x, support_y, query_x, query_y, **kw):
    """The same episode with every floating-point tensor (weights, clouds, features, graph, solve)
    in float64: the ADJUDICATOR of free-running parity.  Two FP32 implementations of this path
    (the reference and the CUDA library) legitimately differ where an FP32 distance tie flips a
    kNN / FPS / argmin decision; how far each of them is from this run is the yardstick
    (tests/test_gpu_parity.py::test_free_running_parity_fp64_adjudicated)."""
    sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
    for name in ("support_feat", "query_feat"):
        if kw.get(name) is not None:
            kw[name] = kw[name].double()
    return forward_episode(sd64, support_x.double(), support_y, query_x.double(), query_y, **kw)


# ------------------------------------------------------------------------------------------------
# evaluate_metric  (reference eval_noise.py:23-72), vectorised
# ------------------------------------------------------------------------------------------------
def confusion_counts(pred_list: List[np.ndarray], gt_list: List[np.ndarray],
                     label2class_list: List[np.ndarray], test_classes: List[int]) -> np.ndarray:
    """(3, n_slots) int64: gt / predicted / true-positive counts per test-class slot."""
    n_slots = len(test_classes) + 1
    out = np.zeros((3, n_slots), dtype=np.int64)
    for pred, gt, l2c in zip(pred_list, gt_list, label2class_list):
        slot = np.array([0] + [test_classes.index(int(c)) + 1 for c in l2c], dtype=np.int64)
        g, p = slot[np.asarray(gt).reshape(-1)], slot[np.asarray(pred).reshape(-1)]
        np.add.at(out[0], g, 1)
        np.add.at(out[1], p, 1)
        hit = np.asarray(gt).reshape(-1) == np.asarray(pred).reshape(-1)
        np.add.at(out[2], g[hit], 1)
    return out


def mean_iou(counters: np.ndarray) -> float:
    """eval_noise.py:64-70: IoU per slot, mean over the foreground slots."""
    gt, pos, tp = counters.astype(np.float64)
    iou = tp / (gt + pos - tp)
    return float(np.mean(iou[1:]))
