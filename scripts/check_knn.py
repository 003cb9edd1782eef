"""k-NN (k = 20) of the same clouds under different A/B switches of the measurement build.
    python scripts/check_knn.py run OUT.pt      (one process per switch setting)
    python scripts/check_knn.py cmp A.pt B.pt
Shapes: C = 9 and C = 64 at N = 2048, ragged tiles (N = 1100, 1501), k < 20, exact duplicate points
(ties), one large cloud."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

mode = sys.argv[1]
if mode == "run":
    from r3dfsseg_b200 import ops
    res = {}
    for C, N, k, B in [(9, 2048, 20, 12), (64, 2048, 20, 12), (9, 1100, 20, 3), (64, 1030, 13, 3),
                       (33, 1501, 20, 2), (64, 1024, 1, 2), (64, 8192, 20, 1), (6, 4000, 7, 2)]:
        g = torch.Generator().manual_seed(C * 1000 + N)
        x = torch.rand((B, C, N), generator=g) if C <= 9 else torch.randn((B, C, N), generator=g)
        x[0, :, 7] = x[0, :, 500]            # duplicates: ties at distance 0 and inside the lists
        x[0, :, 8] = x[0, :, 500]
        x = x.cuda()
        ms = None
        for _ in range(2):
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            idx = ops.knn(x, k, impl=2)
            t1.record()
            torch.cuda.synchronize()
            ms = t0.elapsed_time(t1)
        res[f"{C}_{N}_{k}"] = idx.cpu()
        print(f"C={C} N={N} k={k} B={B}: {ms:.3f} ms", flush=True)
    torch.save(res, sys.argv[2])
else:
    a, b = torch.load(sys.argv[2]), torch.load(sys.argv[3])
    # same neighbour sets; the order inside a run of exactly tied keys follows the tile geometry
    bad = [key for key in a if not torch.equal(a[key].sort(-1)[0], b[key].sort(-1)[0])]
    swaps = {key: int((a[key] != b[key]).any(-1).sum()) for key in a}
    print("rows whose tie order differs:", swaps)
    for key in bad:
        d = (a[key] != b[key]).any(-1)
        bi, ri = d.nonzero()[0].tolist()
        print(key, "rows differing:", int(d.sum()), "of", d.numel(), "first:", bi, ri,
              a[key][bi, ri].tolist(), b[key][bi, ri].tolist())
    print("differing shapes:", bad)
    sys.exit(1 if bad else 0)
