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
for C, N, k, impl in ((9, 300, 20, 2), (64, 257, 20, 2), (70, 130, 5, 1)):
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
print("sanitizer smoke ok", float(loss), tuple(pred.shape))
