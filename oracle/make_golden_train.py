"""TEST INFRASTRUCTURE.  Generates tests/golden/golden_train.pt by running the REFERENCE's own
training forward + backward (MPTI_SelfAtten.forward(train=True) imported unmodified from
/root/reference under oracle/ref_shims.py; attention dropout probability set to 0 on the built
module so the run is deterministic) on seeded synthetic episodes.  Build container only.
Run:  python -m oracle.make_golden_train

Stored per case: lp_loss, contrast_loss, per-parameter gradient norms of
loss = lp_loss + 0.1 * contrast_loss (models/mpti_learner.py:66), a strided sample of every
gradient tensor, and the BatchNorm running statistics after the step.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shims  # noqa: E402
from r3dfsseg_b200.episodes import default_args, make_episode  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
GRAD_SAMPLE_STRIDE = 29

TRAIN_CASES = [
    # name, seed, n_way, k_shot, dataset, noise_ratio
    ("train_s3dis_2way_5shot_clean", 11, 2, 5, "s3dis", 0.0),
    ("train_s3dis_2way_5shot_noisy", 12, 2, 5, "s3dis", 0.4),
]


def run_reference(ref, sd, ep, n_way, k_shot):
    m = ref.mpti.MPTI_SelfAtten(default_args(n_way, k_shot))
    m.load_state_dict(sd)
    m.train()
    m.att_learner.dropout.p = 0.0
    with ref_shims.quiet():
        out = m(ep.support_x, ep.support_y, ep.query_x, ep.query_y, gt_support_y=ep.gt_support_y,
                gt_query_y=ep.query_y, train=True, logger=ref_shims.QuietLogger(),
                support_flag=ep.support_flag)
    query_pred, lp_loss, contrast = out[0], out[1], out[2]
    run_reference.diag = tuple(float(x) for x in out[3:7])  # acc_LP, acc_orig, clean_LP, clean_orig
    loss = lp_loss + 0.1 * contrast
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in m.named_parameters()}
    buffers = {k: v.detach().clone() for k, v in m.named_buffers()}
    return query_pred.detach(), lp_loss.detach(), contrast.detach(), grads, buffers


def main():
    ref = ref_shims.load_reference()
    torch.set_num_threads(os.cpu_count())
    sd = torch.load(os.path.join(GOLD, "weights_fixture.pt"))
    out = {}
    for name, seed, n_way, k_shot, ds, noise in TRAIN_CASES:
        ep = make_episode(seed, n_way, k_shot, dataset=ds, noise_ratio=noise)
        qp, lp, ct, grads, buffers = run_reference(ref, sd, ep, n_way, k_shot)
        out[name] = dict(
            seed=seed, n_way=n_way, k_shot=k_shot, dataset=ds, noise_ratio=noise,
            lp_loss=lp.clone(), contrast_loss=ct.clone(), diagnostics=run_reference.diag,
            query_pred_sub=qp[:, :, ::16].contiguous().clone(),
            grad_norm={k: g.norm().clone() for k, g in grads.items()},
            grad_sample={k: g.reshape(-1)[::GRAD_SAMPLE_STRIDE].clone() for k, g in grads.items()},
            running={k: v for k, v in buffers.items()})
        print(name, "lp", float(lp), "contrast", float(ct), "total grad norm",
              float(torch.sqrt(sum(g.pow(2).sum() for g in grads.values()))))
    torch.save(out, os.path.join(GOLD, "golden_train.pt"))
    print("golden_train.pt written")


if __name__ == "__main__":
    main()
