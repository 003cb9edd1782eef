"""Linear layers (impl 3 = TMA-fed tensor-core kernel) under different A/B switches of the measurement build.
    python scripts/check_linear.py run OUT.pt      (one process per switch setting)
    python scripts/check_linear.py cmp A.pt B.pt
Shapes: the encoder's layers, a ragged M / N / K (K a multiple of 4 only), every activation."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

mode = sys.argv[1]
if mode == "run":
    from r3dfsseg_b200 import ops
    res = {}
    for M, K, N, act in [(4096, 192, 512, 2), (4096, 512, 256, 2), (3000, 256, 128, 1), (2500, 128, 64, 0),
                         (1111, 256, 192, 0), (777, 20, 70, 2), (130, 64, 128, 0), (5000, 36, 33, 1)]:
        g = torch.Generator().manual_seed(M + K + N)
        x = torch.randn((M, K), generator=g).cuda()
        w = (torch.randn((N, K), generator=g) / K ** 0.5).cuda()
        s = (torch.rand((N,), generator=g) + 0.5).cuda()
        t = torch.randn((N,), generator=g).cuda()
        res[f"{M}_{K}_{N}_{act}"] = ops.linear(x, w, s, t, act, impl=3).cpu()
    torch.save(res, sys.argv[2])
else:
    a, b = torch.load(sys.argv[2]), torch.load(sys.argv[3])
    bad = [key for key in a if not torch.equal(a[key], b[key])]
    for key in bad:
        print(key, "max abs diff", float((a[key] - b[key]).abs().max()))
    print("differing cases:", bad)
    sys.exit(1 if bad else 0)
