#!/usr/bin/env python
"""torchrun helper of tests/test_gpu_train.py::test_replicas_stay_identical_nccl_world2."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from r3dfsseg_b200 import train as T
    from r3dfsseg_b200.episodes import default_args, make_episode
    from r3dfsseg_b200.models import MPTI_SelfAtten
    sd = torch.load(os.path.join(ROOT, "tests", "golden", "weights_fixture.pt"))
    m = MPTI_SelfAtten(default_args(2, 5))
    m.load_state_dict(sd)
    m = m.to(dev).train()
    with torch.no_grad():   # ranks start apart on purpose
        for p in m.parameters():
            p.add_(0.01 * (rank + 1))
        for b in m.buffers():
            if b.dtype.is_floating_point:
                b.add_(0.1 * rank)
    learner = T.MPTILearner_V3(default_args(2, 5), mode="train", model=m)
    for step in range(3):
        ep = make_episode(900 + 10 * rank + step, 2, 5, noise_ratio=0.2)
        z = torch.zeros_like(ep.support_y)
        data = [ep.support_x.to(dev), ep.support_y.to(dev), ep.query_x.to(dev), ep.query_y.to(dev),
                z.to(dev), torch.zeros(ep.query_y.shape, dtype=torch.int32, device=dev),
                ep.gt_support_y.to(dev), ep.query_y.to(dev), None, None, ep.support_flag.to(dev)]
        learner.train(data)
    fs = T.flat_state(m)
    both = torch.cat([fs.flat, fs.running])
    gathered = [torch.empty_like(both) for _ in range(dist.get_world_size())]
    dist.all_gather(gathered, both)
    equal = all(torch.equal(gathered[0], g) for g in gathered[1:])
    moved = bool(fs.flat.abs().sum() > 0)
    if rank == 0:
        print("REPLICAS_EQUAL" if equal and moved else "REPLICAS_DIFFER", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if equal else 1)


if __name__ == "__main__":
    main()
