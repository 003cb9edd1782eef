"""SelfAttention under different A/B switches of the measurement build, and its time at the bench's size.
    python scripts/check_attention.py run OUT.pt      (one process per switch setting)
    python scripts/check_attention.py cmp A.pt B.pt
Cases: no first sweep (small activations), approximate first sweep (large), ragged N, N < 128."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

mode = sys.argv[1]
if mode == "run":
    from r3dfsseg_b200 import ops
    res = {}
    for scale, N, B in [(0.1, 2048, 3), (2.5, 2048, 3), (0.5, 1000, 2), (4.0, 333, 2), (1.0, 100, 2),
                        (3.0, 4096, 1)]:
        g = torch.Generator().manual_seed(int(scale * 100) + N)
        x = (torch.randn((B, N, 256), generator=g) * scale).cuda()
        wqkv = (torch.randn((192, 256), generator=g) / 16).cuda()
        res[f"{scale}_{N}"] = ops.attention(x, wqkv).cpu()
    torch.save(res, sys.argv[2])
    for scale in (0.1, 2.5):
        g = torch.Generator().manual_seed(1)
        x = (torch.randn((300, 2048, 256), generator=g) * scale).cuda()
        wqkv = (torch.randn((192, 256), generator=g) / 16).cuda()
        for _ in range(2):
            ops.attention(x, wqkv)
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(3):
            ops.attention(x, wqkv)
        t1.record()
        torch.cuda.synchronize()
        print(f"scale {scale}: qkv projection + attention of 300 clouds x 2048: {t0.elapsed_time(t1) / 3:.3f} ms", flush=True)
else:
    a, b = torch.load(sys.argv[2]), torch.load(sys.argv[3])
    bad = [key for key in a if not torch.equal(a[key], b[key])]
    for key in bad:
        print(key, "max abs diff", float((a[key] - b[key]).abs().max()), "of", float(a[key].abs().max()))
    print("differing cases:", bad)
    sys.exit(1 if bad else 0)
