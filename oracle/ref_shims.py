"""TEST INFRASTRUCTURE — not product code.

Imports the reference's own modules, unmodified, from /root/reference under a small
compatibility shim layer so that they run on CPU with the torch in this image.  This only
works in the build container (where /root/reference is mounted); it is used to
(1) pin oracle/mpti_oracle.py and (2) generate the golden vectors under tests/golden/
(see oracle/make_golden.py).  Nothing here travels to the GPU box at run time.

Shims — these definitions ARE the pinned semantics of the three unpinned
third-party dependencies of `models/mpti.py`:

* `faiss.IndexFlatL2(d).add(X)/.search(X, k)` (call site `models/mpti.py:733-735`):
  exact squared-L2 top-k, ascending, computed in float64, the query itself forced to
  column 0 (the reference drops column 0 blindly at `:736`), ties -> lowest index.
* `torch_cluster.fps(src, None, ratio, random_start=False)` (call site `models/mpti.py:613`):
  upstream CPU algorithm — start at index 0, `dist = min(dist, sum((y - y[last])**2, 1))`
  in the tensor's dtype, `argmax` (first maximum), `m = ceil(float32(n) * float32(ratio))`
  picks, returned in selection order.
* `torch_scatter`: empty stub (only used by the never-called `Check_Proto_Cleanness`).
* `F.pairwise_distance`: torch<=1.8 semantics `norm(x1 - x2 + eps, p, dim=1)` (README pins
  pytorch 1.8; current torch reduces the last dim and the reference then crashes at
  `models/mpti.py:751`).
* `Tensor.cuda()` / `Module.cuda()` -> identity for CPU runs.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

REFERENCE_ROOT = "/root/reference"


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "mpti.py"))


class _IndexFlatL2:
    def __init__(self, d):
        self.d = d
        self.X = None

    def add(self, X):
        self.X = np.asarray(X, dtype=np.float32)

    def search(self, Q, k):
        X = torch.from_numpy(self.X).double()
        Qt = torch.from_numpy(np.asarray(Q, dtype=np.float32)).double()
        d2 = (Qt * Qt).sum(1, keepdim=True) + (X * X).sum(1)[None, :] - 2.0 * (Qt @ X.T)
        same = Qt.shape == X.shape and bool(torch.equal(Qt, X))
        if same:
            d2.fill_diagonal_(-1.0)  # self first
        order = torch.argsort(d2, dim=1, stable=True)[:, :k]
        D = torch.gather(d2, 1, order).clamp_min(0).float().numpy()
        return D, order.numpy().astype(np.int64)


def fps_shim(src, batch=None, ratio=0.5, random_start=True):
    assert batch is None and not random_start
    n = src.shape[0]
    r = torch.tensor(ratio, dtype=src.dtype)
    m = int(torch.ceil(torch.tensor(float(n), dtype=src.dtype) * r).item())
    out = torch.empty(m, dtype=torch.long)
    out[0] = 0
    y = src.detach()
    dist = (y - y[0]).pow(2).sum(1)
    for i in range(1, m):
        a = int(dist.argmax())
        out[i] = a
        dist = torch.min(dist, (y - y[a]).pow(2).sum(1))
    return out


def old_pairwise_distance(x1, x2, p=2.0, eps=1e-6, keepdim=False):
    return torch.norm(x1 - x2 + eps, p, 1, keepdim)


_loaded = None


def load_reference():
    """Returns a namespace with the reference's `dgcnn`, `attention`, `mpti`,
    `mpti_learner`, `protonet` modules and `evaluate_metric`."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise RuntimeError("reference tree not mounted at /root/reference")
    faiss = types.ModuleType("faiss")
    faiss.IndexFlatL2 = _IndexFlatL2
    tc = types.ModuleType("torch_cluster")
    tc.fps = fps_shim
    ts = types.ModuleType("torch_scatter")
    ts.scatter_mean = ts.scatter_add = ts.scatter_max = None
    sys.modules.setdefault("faiss", faiss)
    sys.modules.setdefault("torch_cluster", tc)
    sys.modules.setdefault("torch_scatter", ts)
    F.pairwise_distance = old_pairwise_distance
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    # the reference package is called `models`; make sure ours never shadows it
    for name in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
        del sys.modules[name]
    sys.path.insert(0, REFERENCE_ROOT)
    sys.dont_write_bytecode = True
    try:
        import models.dgcnn as dgcnn
        import models.attention as attention
        import models.mpti as mpti
        import models.mpti_learner as mpti_learner
        import models.protonet as protonet
    finally:
        sys.path.remove(REFERENCE_ROOT)
    ns = types.SimpleNamespace(dgcnn=dgcnn, attention=attention, mpti=mpti,
                               mpti_learner=mpti_learner, protonet=protonet)
    # evaluate_metric lives in a script with heavy imports (h5py...): exec just that function
    src = open(os.path.join(REFERENCE_ROOT, "eval_noise.py")).read()
    start = src.index("def evaluate_metric")
    end = src.index("def test_few_shot")
    g = {"np": np}
    exec(compile(src[start:end], "eval_noise.py[evaluate_metric]", "exec"), g)
    ns.evaluate_metric = g["evaluate_metric"]
    _loaded = ns
    return ns


class QuietLogger:
    def cprint(self, *_a, **_k):
        pass


@contextlib.contextmanager
def quiet():
    """The reference prints from inside forward (`models/mpti.py:453,462,667,672`)."""
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        yield
