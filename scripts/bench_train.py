#!/usr/bin/env python
"""Meta-training step benchmark (BASELINE.json configs[4]): 2-way 5-shot episodic training step with
the way-contrast loss, one episode per GPU per step, NCCL gradient all-reduce of the ONE flat
1.5 MB bucket, fused Adam.

    python scripts/bench_train.py [--steps K] [--warmup W] [--cpu-steps C]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P scripts/bench_train.py --gpus N

One JSON line (rank 0): value = training steps/s x episodes per step (episodes/s trained, whole job),
device-timed max over ranks, with the per-phase split (forward / backward / all-reduce / Adam) and the
CPU oracle's step (reference path under torch autograd + torch.optim.Adam) on the host cores.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def cpu_step_seconds(n_steps, threads):
    from oracle import mpti_train_oracle as TO
    from r3dfsseg_b200.episodes import make_episode
    torch.set_num_threads(threads)
    sd = torch.load(os.path.join(ROOT, "tests", "golden", "weights_fixture.pt"))
    P, running = TO.split_state_dict(sd)
    opt = torch.optim.Adam(TO.param_groups(P, 1e-3), lr=1e-3)
    t0 = time.perf_counter()
    for s in range(n_steps):
        ep = make_episode(500 + s, 2, 5, noise_ratio=0.2)
        out = TO.forward_train(P, ep.support_x, ep.support_y, ep.query_x, ep.query_y, ep.support_flag,
                               running=running)
        opt.zero_grad()
        (out["lp_loss"] + 0.1 * out["contrast_loss"]).backward()
        opt.step()
    return (time.perf_counter() - t0) / n_steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--cpu-steps", type=int, default=2)
    ap.add_argument("--episodes", type=int, default=8, help="distinct synthetic episodes cycled per rank")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.all_reduce(torch.zeros(1, device=dev))
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    line = measure(dev, dist, rank, world, args.steps, args.warmup, args.episodes, args.cpu_steps)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def measure(dev, dist, rank, world, steps, warmup, n_episodes=8, cpu_steps=0):
    """One timed run of the training step on this rank (all ranks call it); returns the JSON line
    (a dict) on every rank.  Used by main() and by bench.py (`extra.train_step`)."""
    import types
    args = types.SimpleNamespace(steps=steps, warmup=warmup, episodes=n_episodes,
                                 cpu_steps=cpu_steps)
    from r3dfsseg_b200 import _lib, train as T
    from r3dfsseg_b200.episodes import default_args, make_episode
    from r3dfsseg_b200.models import MPTI_SelfAtten

    sd = torch.load(os.path.join(ROOT, "tests", "golden", "weights_fixture.pt"))
    margs = default_args(2, 5)
    model = MPTI_SelfAtten(margs)
    model.load_state_dict(sd)
    model = model.to(dev).train()
    opt = T.FusedAdam(model, lr=margs.lr)
    eps = []
    for i in range(args.episodes):  # noisy-shot count drawn from [0, 0.2, 0.4] * k_shot (README.md:51)
        ep = make_episode(7000 + 100 * rank + i, 2, 5, noise_ratio=(0.0, 0.2, 0.4)[i % 3])
        eps.append([t.to(dev) for t in (ep.support_x, ep.support_y, ep.query_x, ep.query_y,
                                        ep.support_flag)])
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

    def one_step(i, ev=None):
        sx, sy, qx, qy, flag = eps[i % len(eps)]
        if ev: ev[0].record()
        qp, lp, ct = T.train_episode(model, sx, sy, qx, qy, flag)
        loss = lp + 0.1 * ct
        if ev: ev[1].record()
        opt.zero_grad()
        loss.backward()
        if ev: ev[2].record()
        opt.step()
        if ev: ev[3].record()
        return loss

    def sync_all():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    for i in range(args.warmup):
        one_step(i)
    sync_all()
    L = _lib.lib()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    launches0 = L.r3dfs_launch_count()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    sync_all()
    t0.record()
    for s in range(args.steps):
        flush.zero_()
        loss = one_step(args.warmup + s, evs[s])
    t1.record()
    sync_all()
    launches = L.r3dfs_launch_count() - launches0
    total_ms = t0.elapsed_time(t1)
    fwd = sum(e[0].elapsed_time(e[1]) for e in evs) / args.steps
    bwd = sum(e[1].elapsed_time(e[2]) for e in evs) / args.steps
    upd = sum(e[2].elapsed_time(e[3]) for e in evs) / args.steps
    if dist is not None:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t[0])
    # the all-reduce alone, on the same bucket
    ar_us = None
    if dist is not None:
        g = torch.zeros_like(T.flat_state(model).flat)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(5):
            dist.all_reduce(g)
        sync_all()
        a.record()
        for _ in range(20):
            dist.all_reduce(g)
        b.record()
        torch.cuda.synchronize()
        ar_us = a.elapsed_time(b) / 20 * 1e3
    cpu = None
    if rank == 0 and world == 1 and args.cpu_steps > 0:
        threads = os.cpu_count() or 1
        sec = cpu_step_seconds(args.cpu_steps, threads)
        cpu = {"value": 1.0 / sec, "unit": "steps/s", "cores": threads, "kind": "port",
               "sample": f"{args.cpu_steps} steps (forward + autograd backward + torch Adam) of "
                         f"oracle/mpti_train_oracle.py, torch fp32, {threads} threads"}
    if True:
        line = {"metric": "MPTI meta-training episodes/s (2-way 5-shot, way-contrast loss)",
                "value": world * args.steps / (total_ms * 1e-3), "unit": "episodes/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": "2-way 5-shot episodic meta-training step, one episode per GPU "
                                       "per step, lp + 0.1 * way-contrast, attention dropout 0.1, Adam "
                                       "(BASELINE.json configs[4])",
                           "l2": "512 MiB flush write between steps"},
                "phases_ms": {"forward": round(fwd, 3), "backward": round(bwd, 3),
                              "allreduce_plus_adam": round(upd, 3)},
                "allreduce_us_1p5MB": ar_us, "gpu_launches": int(launches),
                "final_loss": float(loss.detach()), "cpu_baseline": cpu}
    return line


if __name__ == "__main__":
    main()
