"""Small end-to-end run for compute-sanitizer (memcheck): one 2-way 1-shot episode + the
stand-alone ops at small sizes.  Usage: compute-sanitizer --tool memcheck python scripts/sanitizer_smoke.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from r3dfsseg_b200 import ops  # noqa: E402
from r3dfsseg_b200.episodes import default_args, make_episode  # noqa: E402
from r3dfsseg_b200.models import MPTI_SelfAtten  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
dev = "cuda:0"
sd = torch.load(os.path.join(ROOT, "tests", "golden", "weights_fixture.pt"))
g = torch.Generator().manual_seed(0)
for C, N, k, impl in ((9, 300, 20, 2), (64, 257, 20, 2), (70, 130, 5, 1), (9, 1100, 20, 2),
                      (64, 1030, 20, 2)):  # the last two take the two-pass tensor-core kernel
    x = torch.randn((2, C, N), generator=g).to(dev)
    idx = ops.knn(x, k, impl=impl)
    e = ops.get_edge_feature(x, k, idx)
    assert e.shape == (2, 2 * C, N, k)
feat = torch.randn((3000, 192), generator=g).to(dev) * 0.2
off = torch.tensor([0, 1000, 1050], dtype=torch.int32, device=dev)
n = torch.tensor([1000, 50, 1950], dtype=torch.int32, device=dev)
ops.multi_prototypes(feat, off, n, 100)
m = MPTI_SelfAtten(default_args(2, 1))
m.load_state_dict(sd)
m = m.to(dev).eval()
ep = make_episode(0, 2, 1)
pred, loss = m(ep.support_x.to(dev), ep.support_y.to(dev), ep.query_x.to(dev), ep.query_y.to(dev),
               eval=True)
torch.cuda.synchronize()
# point-major get_edge_feature (row-gather kernel) and one training step
xpm = torch.randn((2, 1100, 64), generator=g).to(dev)
e = ops.get_edge_feature(xpm.transpose(1, 2), 20, ops.knn(xpm.transpose(1, 2), 20))
if os.environ.get("SMOKE_TRAIN", "1") == "1":
    from r3dfsseg_b200 import train as T
    m5 = MPTI_SelfAtten(default_args(2, 2))
    m5.load_state_dict(sd)
    m5 = m5.to(dev).train()
    ep2 = make_episode(1, 2, 2)
    qp, lp, ct = T.train_episode(m5, ep2.support_x.to(dev), ep2.support_y.to(dev), ep2.query_x.to(dev),
                                 ep2.query_y.to(dev), ep2.support_flag.to(dev))
    (lp + 0.1 * ct).backward()
    torch.cuda.synchronize()
    print("train smoke ok", float(lp), float(ct))
print("sanitizer smoke ok", float(loss), tuple(pred.shape))
