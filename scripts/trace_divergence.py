#!/usr/bin/env python
"""Where does a free-running CUDA episode leave the FP64 adjudicator?  (GPU box; TEST tooling.)
For one parity case: kNN set flips per EdgeConv layer (teacher-forced on the FP64 layer inputs), feature
deviations, MDNS flags, FPS seed sets and prototype counts — each for CUDA-vs-FP64 and for the
FP32 oracle-vs-FP64, so that a label disagreement can be attributed to the first discrete decision
that flipped."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from oracle import mpti_oracle as O  # noqa: E402
from r3dfsseg_b200 import ops  # noqa: E402
from r3dfsseg_b200.episodes import default_args, make_episode  # noqa: E402
from r3dfsseg_b200.models import MPTI_SelfAtten  # noqa: E402


def set_flips(a, b):
    """rows whose neighbour SETS differ"""
    a, b = a.sort(-1).values, b.sort(-1).values
    return int((a != b).any(-1).sum())


def main():
    name = sys.argv[1]
    dev = torch.device("cuda", 0)
    torch.set_num_threads(os.cpu_count())
    sd = torch.load(os.path.join(ROOT, "tests", "golden", "weights_fixture.pt"))
    sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
    c = torch.load(os.path.join(ROOT, "tests", "golden", "golden_parity.pt"))[name]
    n_way, k_shot = c["n_way"], c["k_shot"]
    ep = make_episode(c["seed"], n_way, k_shot, dataset=c["dataset"], noise_ratio=c["noise_ratio"])
    m = MPTI_SelfAtten(default_args(n_way, k_shot))
    m.load_state_dict(sd)
    m = m.to(dev).eval()
    X = torch.cat([ep.support_x.reshape(n_way * k_shot, 9, -1), ep.query_x], 0)
    rep = {"case": name}
    with torch.no_grad():
        # kNN flips per layer on IDENTICAL inputs (the FP64 run's layer inputs, rounded to FP32)
        x64 = X.double()
        for layer in range(3):
            key64 = O.knn_scores(x64.float())                      # FP64 ranking of the FP32 inputs
            idx64 = key64.topk(20, dim=-1)[1]
            idx32 = O.knn(x64.float(), 20)
            idxc = ops.knn(x64.float().to(dev), 20).cpu()
            rep[f"knn{layer}_rows"] = int(idx64.shape[0] * idx64.shape[1])
            rep[f"knn{layer}_flips_ref32"] = set_flips(idx32, idx64)
            rep[f"knn{layer}_flips_cuda"] = set_flips(idxc, idx64)
            x64 = O.edgeconv_block(x64, sd64, f"encoder.edge_convs.{layer}", 20)
        f64 = O.get_features(X.double(), sd64)
        f32 = O.get_features(X, sd)
        fc = m.getFeatures(X.to(dev)).cpu()
        for tag, f in (("ref32", f32), ("cuda", fc)):
            d = (f.double() - f64).abs().amax(1)                   # per point
            rep[f"feat_pts_off_{tag}"] = [int((d > t).sum()) for t in (1e-4, 1e-3, 1e-2, 1e-1)]
        ns = n_way * k_shot
        outs = {}
        for tag, f, s in (("fp64", f64, sd64), ("ref32", f32, sd), ("cuda", fc, sd)):
            fn = O.forward_episode
            kw = dict(eval_mdns=c["eval"], support_feat=f[:ns], query_feat=f[ns:], keep=True)
            sx = ep.support_x.double() if tag == "fp64" else ep.support_x
            qx = ep.query_x.double() if tag == "fp64" else ep.query_x
            outs[tag] = fn(s, sx, ep.support_y, qx, ep.query_y, **kw)
        for tag in ("ref32", "cuda"):
            o, g = outs[tag], outs["fp64"]
            rep[f"clean_equal_{tag}"] = None if not c["eval"] else bool(
                torch.equal(o["clean_flag"], g["clean_flag"]))
            rep[f"proto_count_{tag}"] = o["proto_count"]
            rep[f"seed_sets_equal_{tag}"] = [bool(torch.equal(a, b)) for a, b in
                                             zip(o["seed_idx"], g["seed_idx"])]
            rep[f"seeds_differ_{tag}"] = [int(len(set(a.tolist()) ^ set(b.tolist())) // 2)
                                          if len(a) == len(b) else -1
                                          for a, b in zip(o["seed_idx"], g["seed_idx"])]
            rep[f"labels_{tag}"] = float((o["pred"] == g["pred"]).float().mean())
        rep["proto_count_fp64"] = outs["fp64"]["proto_count"]
        pred, _ = m(ep.support_x.to(dev), ep.support_y.to(dev), ep.query_x.to(dev),
                    ep.query_y.to(dev), gt_support_y=ep.gt_support_y.to(dev), eval=c["eval"])
        rep["labels_cuda_full_vs_fp64"] = float(
            (pred.argmax(1).cpu() == outs["fp64"]["pred"]).float().mean())
        rep["labels_cuda_full_vs_cudafeat_oracle"] = float(
            (pred.argmax(1).cpu() == outs["cuda"]["pred"]).float().mean())
    print(json.dumps(rep))


if __name__ == "__main__":
    main()
