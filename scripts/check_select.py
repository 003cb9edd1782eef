"""Selection / in-edge kernels of the affinity graph on the same inputs under different A/B switches.
    python scripts/check_select.py run OUT.pt [n] [G]     (one process per switch setting)
    python scripts/check_select.py cmp A.pt B.pt
Graph 1 has fewer valid nodes than k (the filler path), graph 2 a block of invalid nodes."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

mode = sys.argv[1]
if mode == "run":
    from r3dfsseg_b200 import ops
    out = sys.argv[2]
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 2368
    G = int(sys.argv[4]) if len(sys.argv) > 4 else 25
    D, k = 192, 200
    g = torch.Generator().manual_seed(0)
    centers = torch.randn((6, D), generator=g) * 0.15
    feat = centers[torch.randint(0, 6, (G, n), generator=g)] + torch.randn((G, n, D), generator=g) * 0.06
    feat[0, 5] = feat[0, 900]          # exact duplicates -> ties at distance 0 and at the cut
    feat[0, 6] = feat[0, 900]
    valid = torch.ones((G, n), dtype=torch.uint8)
    valid[:, 300:320] = 0
    if G > 1:
        valid[1] = 0
        valid[1, 10:160] = 1           # 150 valid nodes < k
    feat, valid = feat.cuda(), valid.cuda()
    for _ in range(3):
        nbr, sim = ops.affinity_knn(feat, valid, k, 1.0)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(5):
        nbr, sim = ops.affinity_knn(feat, valid, k, 1.0)
    t1.record()
    torch.cuda.synchronize()
    Y = torch.zeros((G, n, 3), device="cuda")
    Y[:, :100, 0] = 1
    Y[:, 100:200, 1] = 1
    Y[:, 200:300, 2] = 1
    Z, iters, resid = ops.label_propagate(nbr, sim, valid, Y)
    print("affinity ms", t0.elapsed_time(t1) / 5, "cg iters", iters.tolist()[:4])
    torch.save({"nbr": nbr.cpu(), "sim": sim.cpu(), "valid": valid.cpu(), "Z": Z.cpu()}, out)
else:
    a, b = torch.load(sys.argv[2]), torch.load(sys.argv[3])
    v = a["valid"].bool()
    same = (a["nbr"] == b["nbr"])[v]
    za, zb = torch.nan_to_num(a["Z"], nan=-1.0), torch.nan_to_num(b["Z"], nan=-1.0)
    ok = bool(same.all()) and torch.equal(a["sim"][v], b["sim"][v]) and torch.equal(za, zb)
    print("rows", int(v.sum()), "identical rows", int(same.all(-1).sum()),
          "sim equal", bool(torch.equal(a["sim"][v], b["sim"][v])), "Z equal", torch.equal(za, zb))
    sys.exit(0 if ok else 1)
