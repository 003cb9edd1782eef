#!/usr/bin/env python
"""TEST INFRASTRUCTURE (GPU box).  Meta-trains the weights fixture with the repo's own training path
(`MPTILearner_V3.train`, reference models/mpti_learner.py:50-79) on synthetic noisy episodes, so
that the golden episodes are evaluated with class margins that are not degenerate (SURVEY.md §8(d)
"Weights fixture": "a few hundred meta-training steps so labels are non-trivial").

    python scripts/train_fixture.py --steps 600 --out gpurun_out/weights_trained.pt

Starts from tests/golden/weights_init.pt (the BN-calibrated initialisation), trains 2-way 5-shot
episodes with the reference's train-mode noise ratios [0, 0.2, 0.4] (README.md:51), and writes
  * the trained state_dict (reference key names),
  * a JSON log (loss / accuracy curve, held-out eval accuracy before and after),
  * `check`: losses of ONE extra training forward (dropout off, running statistics frozen) on a fixed
    episode, which `oracle/mpti_train_oracle.py` re-computes on the CPU from the same weights
    (tests/test_oracle_golden.py::test_trained_fixture_loss_matches_oracle).
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

CHECK_SEED = 777


def eval_accuracy(model, seeds, n_way, k_shot, noise_ratio, dataset):
    from r3dfsseg_b200.episodes import make_episode
    model.eval()
    acc, iou_num, iou_den = [], 0.0, 0.0
    with torch.no_grad():
        for s in seeds:
            ep = make_episode(s, n_way, k_shot, noise_ratio=noise_ratio, dataset=dataset)
            pred, _ = model(ep.support_x.cuda(), ep.support_y.cuda(), ep.query_x.cuda(),
                            ep.query_y.cuda(), gt_support_y=ep.gt_support_y.cuda(), eval=True)
            lab = pred.argmax(1).cpu()
            acc.append(float((lab == ep.query_y).float().mean()))
            for c in range(1, n_way + 1):
                tp = float(((lab == c) & (ep.query_y == c)).sum())
                un = float(((lab == c) | (ep.query_y == c)).sum())
                iou_num += tp
                iou_den += un
    return sum(acc) / len(acc), iou_num / max(iou_den, 1.0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=600)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "weights_trained.pt"))
    args = ap.parse_args()
    from r3dfsseg_b200 import train as T
    from r3dfsseg_b200.episodes import default_args, make_episode
    from r3dfsseg_b200.models import MPTI_SelfAtten

    torch.manual_seed(0)
    dev = torch.device("cuda", 0)
    sd0 = torch.load(os.path.join(ROOT, "tests", "golden", "weights_init.pt"))
    margs = default_args(2, 5)
    model = MPTI_SelfAtten(margs)
    model.load_state_dict(sd0)
    model = model.to(dev)
    log = {"steps": args.steps, "curve": []}
    log["eval_before"] = {
        "s3dis_2way_clean": eval_accuracy(model, range(0, 8), 2, 5, 0.0, "s3dis"),
        "s3dis_2way_ood40": eval_accuracy(model, range(0, 8), 2, 5, 0.4, "s3dis")}
    learner = T.MPTILearner_V3(margs, mode="train", model=model)
    t0 = time.time()
    run = [0.0, 0.0, 0.0, 0.0]
    for step in range(args.steps):
        ep = make_episode(20000 + step, 2, 5, noise_ratio=[0.0, 0.2, 0.4])
        z = torch.zeros_like(ep.support_y)
        data = [ep.support_x.to(dev), ep.support_y.to(dev), ep.query_x.to(dev), ep.query_y.to(dev),
                z.to(dev), torch.zeros(ep.query_y.shape, dtype=torch.int32, device=dev),
                ep.gt_support_y.to(dev), ep.query_y.to(dev), None, None, ep.support_flag.to(dev)]
        out = learner.train(data)
        run[0] += float(out[0]); run[1] += float(out[1]); run[2] += float(out[2]); run[3] += out[3]
        if (step + 1) % 25 == 0:
            row = [step + 1] + [x / 25 for x in run]
            log["curve"].append(row)
            print("step %d loss %.4f lp %.4f contrast %.4f acc %.4f" % tuple(row), flush=True)
            run = [0.0, 0.0, 0.0, 0.0]
    log["train_seconds"] = time.time() - t0
    log["eval_after"] = {
        "s3dis_2way_clean": eval_accuracy(model, range(0, 8), 2, 5, 0.0, "s3dis"),
        "s3dis_2way_ood40": eval_accuracy(model, range(0, 8), 2, 5, 0.4, "s3dis")}
    # one more training forward on a fixed episode, dropout off, running statistics untouched
    ep = make_episode(CHECK_SEED, 2, 5, noise_ratio=0.2)
    model.train()
    qp, lp, ct = T.train_episode(model, ep.support_x.to(dev), ep.support_y.to(dev),
                                 ep.query_x.to(dev), ep.query_y.to(dev), ep.support_flag.to(dev),
                                 dropout_p=0.0, update_running=False)
    log["check"] = {"seed": CHECK_SEED, "noise_ratio": 0.2, "lp_loss": float(lp),
                    "contrast_loss": float(ct)}
    model.eval()
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    torch.save(sd, args.out)
    with open(os.path.splitext(args.out)[0] + ".json", "w") as f:
        json.dump(log, f)
    print(json.dumps({k: log[k] for k in ("eval_before", "eval_after", "check", "train_seconds")}))


if __name__ == "__main__":
    main()
