#!/usr/bin/env python
"""Free-running parity report (GPU box): for every case of tests/golden/golden_parity.pt, how far
the CUDA episode and the reference's FP32 run are from the FP64 adjudicator, and from each other.
One JSON line per case; the bars of tests/test_gpu_parity.py::test_free_running_parity_* were set
from this report."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from r3dfsseg_b200.episodes import default_args, make_episode  # noqa: E402
from r3dfsseg_b200.models import MPTI_SelfAtten  # noqa: E402


def stats(a, ref):
    rel = ((a - ref).abs() / ref.abs().max()).reshape(-1)
    return dict(labels=float((a.argmax(1) == ref.argmax(1)).float().mean()),
                median=float(rel.median()), p999=float(torch.quantile(rel, 0.999)),
                max=float(rel.max()))


def main():
    dev = torch.device("cuda", 0)
    sd = torch.load(os.path.join(ROOT, "tests", "golden", "weights_fixture.pt"))
    gold = torch.load(os.path.join(ROOT, "tests", "golden", "golden_parity.pt"))
    models = {}
    for name, c in gold.items():
        key = (c["n_way"], c["k_shot"])
        if key not in models:
            m = MPTI_SelfAtten(default_args(*key))
            m.load_state_dict(sd)
            models[key] = m.to(dev).eval()
        m = models[key]
        ep = make_episode(c["seed"], c["n_way"], c["k_shot"], dataset=c["dataset"],
                          noise_ratio=c["noise_ratio"])
        with torch.no_grad():
            pred, loss = m(ep.support_x.to(dev), ep.support_y.to(dev), ep.query_x.to(dev),
                           ep.query_y.to(dev), gt_support_y=ep.gt_support_y.to(dev), eval=c["eval"])
        pred = pred.cpu()
        row = dict(name=name, cuda_vs_fp64=stats(pred, c["query_pred_fp64"]),
                   ref32_vs_fp64=stats(c["query_pred"], c["query_pred_fp64"]),
                   cuda_vs_ref32=stats(pred, c["query_pred"]), loss=float(loss),
                   loss_ref32=float(c["loss"]), loss_fp64=float(c["loss_fp64"]),
                   acc=float((pred.argmax(1) == ep.query_y).float().mean()),
                   clean_equal=(None if not c["eval"] else bool(torch.equal(
                       m._last_diag["clean_flag"][0].cpu(), c["clean_flag_fp64"]))))
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
