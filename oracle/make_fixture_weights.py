"""TEST INFRASTRUCTURE.  Builds tests/golden/weights_init.pt — the initialisation that scripts/train_fixture.py meta-trains
into tests/golden/weights_fixture.pt, the "checkpoint" shared by
the oracle, the reference-under-shims and the CUDA path (SURVEY.md §8d "Weights fixture").

`torch.manual_seed(0)` + the reference's own `MPTI_SelfAtten(args)` constructor, then
  1. q/k maps scaled so the attention logits are not flat (std ~1.5),
  2. the three 64-channel output groups scaled to std ~0.18 so that Gaussian affinities
     with sigma=1 are informative (raw BN-calibrated features give d^2 ~ 200 -> exp
     underflow; raw init gives d^2 ~ 0 -> every affinity 1),
  3. a BN-calibration pass (train-mode getFeatures, cumulative-average momentum) over
     48 synthetic clouds so running statistics are not the (0, 1) initial values.
Needs /root/reference (build container only).  Run:  python -m oracle.make_fixture_weights
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_shims  # noqa: E402
from r3dfsseg_b200.episodes import default_args, make_episode  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                   "tests", "golden", "weights_init.pt")


def calibration_clouds():
    xs = []
    for s in range(1000, 1004):
        ep = make_episode(s, 2, 5)
        xs.append(ep.support_x.reshape(10, 9, 2048))
        xs.append(ep.query_x)
    return torch.cat(xs, 0)


def calibrate(m, X):
    bns = [mod for mod in m.modules()
           if isinstance(mod, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d))]
    for mod in bns:
        mod.reset_running_stats()
        mod.momentum = None
    m.train()
    p = m.att_learner.dropout.p
    m.att_learner.dropout.p = 0.0
    with torch.no_grad():
        for i in range(0, X.shape[0], 12):
            m.getFeatures(X[i:i + 12])
    m.att_learner.dropout.p = p
    for mod in bns:
        mod.momentum = 0.1
    m.eval()


def main():
    ref = ref_shims.load_reference()
    torch.manual_seed(0)
    m = ref.mpti.MPTI_SelfAtten(default_args(2, 5))
    X = calibration_clouds()
    target = 0.18
    with torch.no_grad():
        bn = m.encoder.edge_convs[0].layer[4]
        bn.weight.mul_(target / 0.8)
        bn.bias.mul_(target / 0.8)
    calibrate(m, X)
    with torch.no_grad():
        _, f2 = m.encoder(X[:12])
        q = m.att_learner.q_map(f2)
        k = m.att_learner.k_map(f2)
        logits = torch.matmul(q.transpose(1, 2) / m.att_learner.temperature, k)
        a = (1.5 / logits.std().item()) ** 0.5
        m.att_learner.q_map.weight.mul_(a)
        m.att_learner.k_map.weight.mul_(a)
        f = m.getFeatures(X[:12])
        s_att = f[:, 64:128].std(dim=2).mean().item()
        s_base = f[:, 128:].std(dim=2).mean().item()
        m.att_learner.v_map.weight.mul_(target / s_att)
        last = m.base_learner.convs[-1][1]
        last.weight.mul_(target / s_base)
        last.bias.mul_(target / s_base)
    calibrate(m, X)
    with torch.no_grad():
        f = m.getFeatures(X[:12])
    print("per-point std by group:", f[:, :64].std(dim=2).mean().item(),
          f[:, 64:128].std(dim=2).mean().item(), f[:, 128:].std(dim=2).mean().item())
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    torch.save(sd, OUT)
    print("wrote", OUT, sum(v.numel() for v in sd.values()), "values")


if __name__ == "__main__":
    main()
