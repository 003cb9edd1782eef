"""GPU debug: training forward/backward vs the teacher-forced CPU oracle."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import mpti_train_oracle as TO
from r3dfsseg_b200.episodes import default_args, make_episode
from r3dfsseg_b200.models import MPTI_SelfAtten
from r3dfsseg_b200 import train as T

torch.set_num_threads(os.cpu_count())
sd = torch.load(os.path.join(ROOT, "tests/golden/weights_fixture.pt"))
for seed, noise, p_drop in ((11, 0.0, 0.0), (12, 0.4, 0.1)):
    ep = make_episode(seed, 2, 5, dataset="s3dis", noise_ratio=noise)
    m = MPTI_SelfAtten(default_args(2, 5))
    m.load_state_dict(sd)
    m = m.cuda().train()
    ks = kq = None
    if p_drop > 0:
        ks = T.dropout_mask(1234, (10, 2048, 2048), p_drop, "cuda")
        kq = T.dropout_mask(99, (2, 2048, 2048), p_drop, "cuda")
        print("keep fraction", float(ks.float().mean()), float(kq.float().mean()))
    qp, lp, ct = T.train_episode(m, ep.support_x.cuda(), ep.support_y.cuda(), ep.query_x.cuda(), ep.query_y.cuda(),
                                 ep.support_flag.cuda(), dropout_p=p_drop, keep_support=ks, keep_query=kq)
    (lp + 0.1 * ct).backward()
    torch.cuda.synchronize()
    forced = T.export_decisions(m)
    P, running = TO.split_state_dict(sd)
    t0 = time.time()
    out = TO.forward_train(P, ep.support_x, ep.support_y, ep.query_x, ep.query_y, ep.support_flag, running=running,
                           keep_mask_support=None if ks is None else ks.cpu(), keep_mask_query=None if kq is None else kq.cpu(),
                           dropout_p=p_drop, forced=forced)
    (out["lp_loss"] + 0.1 * out["contrast_loss"]).backward()
    print("oracle s", time.time() - t0)
    print("seed", seed, "lp", float(lp), float(out["lp_loss"]), "ct", float(ct), float(out["contrast_loss"]),
          "logit err", float((qp.detach().cpu() - out["query_pred"]).abs().max() / out["query_pred"].abs().max()))
    named = dict(m.named_parameters())
    for k in T.PARAM_NAMES:
        gr = named[k].grad.detach().cpu()
        ref = P[k].grad
        err = float((gr - ref).abs().max()) / (float(ref.abs().max()) + 1e-12)
        print(f"  {k:45s} norm {float(gr.norm()):.4e} ref {float(ref.norm()):.4e} max relerr {err:.2e}")
    for k, v in m.named_buffers():
        if v.dtype.is_floating_point:
            e = float((v.cpu() - running[k]).abs().max())
            if e > 1e-5:
                print("  running", k, e)
        else:
            assert int(v) == int(running[k]), (k, int(v), int(running[k]))
