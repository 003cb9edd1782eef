"""Selection kernels of the affinity graph: warp-per-row (default) vs block-per-row
(R3DFS_SELECT_BLOCK=1) on the same inputs; run once per variant, then `cmp` to compare."""
import sys
import torch
from r3dfsseg_b200 import ops

mode = sys.argv[1]
out = sys.argv[2]
G, n, D, k = 25, 2368, 192, 200
g = torch.Generator().manual_seed(0)
centers = torch.randn((6, D), generator=g) * 0.15
feat = centers[torch.randint(0, 6, (G, n), generator=g)] + torch.randn((G, n, D), generator=g) * 0.06
valid = torch.ones((G, n), dtype=torch.uint8)
valid[:, 300:320] = 0
valid[3, 1000:] = 0
valid[3, 1000:1100] = 1
feat, valid = feat.cuda(), valid.cuda()
if mode == "run":
    for _ in range(3):
        nbr, sim = ops.affinity_knn(feat, valid, k, 1.0)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(5):
        nbr, sim = ops.affinity_knn(feat, valid, k, 1.0)
    t1.record()
    torch.cuda.synchronize()
    print("affinity ms", t0.elapsed_time(t1) / 5)
    torch.save({"nbr": nbr.cpu(), "sim": sim.cpu(), "valid": valid.cpu()}, out)
else:
    a, b = torch.load(sys.argv[2]), torch.load(sys.argv[3])
    v = a["valid"].bool()
    same = (a["nbr"] == b["nbr"])[v]
    print("rows", int(v.sum()), "identical rows", int(same.all(-1).sum()),
          "sim equal", bool(torch.equal(a["sim"][v], b["sim"][v])))
    bad = torch.nonzero(~same.all(-1)).flatten()[:5]
    for r in bad.tolist():
        ra, rb = a["nbr"][v][r], b["nbr"][v][r]
        d = torch.nonzero(ra != rb).flatten()
        print("row", r, "first diff at", int(d[0]), ra[d[0]:d[0] + 6].tolist(), rb[d[0]:d[0] + 6].tolist())
