"""BASELINE.json configs[1]: DGCNN EdgeConv microbench — knn, get_edge_feature (materialising) and
the fused EdgeConv block at batch 64 x {2048, 8192} points, k = 20, one B200.
CUDA events, 10 warm-up + median of 30; prints one JSON line per (op, C, N)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from r3dfsseg_b200 import ops  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.isfile(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}


def timeit(fn, warm=10, rep=30):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(rep):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    dev = "cuda:0"
    sd = torch.load(os.path.join(ROOT, "tests", "golden", "weights_fixture.pt"))
    B, k = 64, 20
    g = torch.Generator().manual_seed(0)
    for N in (2048, 8192):
        for C, blk in ((9, 0), (64, 1)):
            x = (torch.rand((B, C, N), generator=g) if C == 9 else torch.randn((B, C, N), generator=g)).to(dev)
            p = f"encoder.edge_convs.{blk}"

            def fold(pref):
                s = sd[pref + ".weight"] / torch.sqrt(sd[pref + ".running_var"] + 1e-5)
                return s.to(dev), (sd[pref + ".bias"] - sd[pref + ".running_mean"] * s).to(dev)
            s1, t1 = fold(p + ".layer.1")
            s2, t2 = fold(p + ".layer.4")
            w1, w2 = sd[p + ".layer.0.weight"].to(dev), sd[p + ".layer.3.weight"].to(dev)
            idx = ops.knn(x, k)
            recs = []
            for impl, name in ((2, "knn_tcgen05"), (1, "knn_fp32_simt")):
                ms = timeit(lambda: ops.knn(x, k, impl=impl))
                fl = 2.0 * B * N * N * C + 3.0 * B * N * N
                recs.append(dict(op=name, ms=ms, tflops=fl / ms / 1e9, bound="tensor",
                                 frac=fl / ms / 1e9 / PEAKS["bf16_tflops"]))
            ms = timeit(lambda: ops.get_edge_feature(x, k, idx))
            by = 4.0 * B * C * N + 8.0 * B * N * k + 4.0 * B * 2 * C * N * k
            recs.append(dict(op="get_edge_feature", ms=ms, gbs=by / ms / 1e6, bound="hbm",
                             frac=by / ms / 1e6 / PEAKS["hbm_gbs"]))
            ms = timeit(lambda: ops.edgeconv(x, w1, s1, t1, w2, s2, t2, k))
            fl = 2.0 * B * N * N * C + 2.0 * (2 * C) * 64 * B * N * k + 2.0 * 64 * 64 * B * N * k
            recs.append(dict(op="edgeconv_fused", ms=ms, tflops=fl / ms / 1e9, bound="tensor",
                             frac=fl / ms / 1e9 / PEAKS["bf16_tflops"],
                             compulsory_mb=4.0 * B * N * (C + 64) / 1e6))
            for r in recs:
                r.update(B=B, C=C, N=N, k=k)
                print(json.dumps({kk: (round(v, 5) if isinstance(v, float) else v) for kk, v in r.items()}))
            del x, idx
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
