"""Sharded, batched evaluation driver: the reference's `test_few_shot` + `evaluate_metric`
(reference eval_noise.py:23-113) for many episodes per call and many GPUs.

Episodes are independent (eval_noise.py:85-106), so episode i belongs to rank i % world; every rank
runs its shard through `MPTI_SelfAtten.forward_episodes` in batches, accumulates the three
evaluate_metric counter rows (ground truth / predicted / true positive per test-class slot,
eval_noise.py:35-37) on the device, and ONE sum all-reduce of (3 x n_slots) int64 + the loss sum
ends the run.  Nothing else crosses GPUs.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch


def shard_indices(n_items: int, rank: int, world: int) -> List[int]:
    """Static round-robin sharding of independent episodes."""
    if not (0 <= rank < world):
        raise ValueError("rank must be in [0, world)")
    return list(range(rank, n_items, world))


def class_slots(sampled_classes: Sequence[int], test_classes: Sequence[int]) -> List[int]:
    """Episode-local label l (1-based) -> slot test_classes.index(sampled_classes[l-1]) + 1
    (reference eval_noise.py:48-59); slot 0 is the background."""
    tc = list(test_classes)
    return [tc.index(int(c)) + 1 for c in sampled_classes]


def iou_from_counters(counters) -> Dict[str, object]:
    """IoU_c = TP / (GT + Pred - TP); mean over the foreground slots (eval_noise.py:64-70)."""
    c = counters.detach().cpu().numpy().astype(np.float64) if torch.is_tensor(counters) \
        else np.asarray(counters, dtype=np.float64)
    gt, pos, tp = c
    with np.errstate(divide="ignore", invalid="ignore"):
        iou = tp / (gt + pos - tp)
    return {"iou": iou, "mean_iou": float(np.mean(iou[1:]))}


def all_reduce_eval_state(counters: torch.Tensor, loss_sum: torch.Tensor, n_episodes: torch.Tensor):
    """Sum the per-rank partial results in place.  No-op without an initialised process group
    (single GPU).  Works with any backend (nccl on GPUs, gloo in the CPU tests)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(counters, op=dist.ReduceOp.SUM)
        dist.all_reduce(loss_sum, op=dist.ReduceOp.SUM)
        dist.all_reduce(n_episodes, op=dist.ReduceOp.SUM)
    return counters, loss_sum, n_episodes


class EpisodeEvaluator:
    """Runs a list of episodes (r3dfsseg_b200.episodes.Episode-like objects) through the model.

    model         r3dfsseg_b200.models.MPTI_SelfAtten on a CUDA device, eval mode
    test_classes  the fold's test class ids (defines the counter slots)
    batch         episodes per C-ABI call
    """

    def __init__(self, model, test_classes: Sequence[int], batch: int = 16, eval_mdns: bool = True):
        self.model = model
        self.test_classes = list(test_classes)
        self.batch = int(batch)
        self.eval_mdns = eval_mdns
        self.device = next(model.parameters()).device

    def _stage(self, eps):
        """Pinned, point-major batch (the reference's .h5 layout, loader.py:1687-1721) -> device."""
        sx = torch.stack([e.support_x.transpose(2, 3) for e in eps]).pin_memory()
        sy = torch.stack([e.support_y for e in eps]).pin_memory()
        qx = torch.stack([e.query_x.transpose(1, 2) for e in eps]).pin_memory()
        qy = torch.stack([e.query_y for e in eps]).pin_memory()
        slot = torch.tensor([class_slots(e.sampled_classes, self.test_classes) for e in eps],
                            dtype=torch.int32).pin_memory()
        dev = self.device
        return (sx.to(dev, non_blocking=True).transpose(3, 4), sy.to(dev, non_blocking=True),
                qx.to(dev, non_blocking=True).transpose(2, 3), qy.to(dev, non_blocking=True),
                slot.to(dev, non_blocking=True))

    def run(self, episodes: Sequence, rank: int = 0, world: int = 1, logger=None,
            log_every: int = 50) -> Dict[str, object]:
        """logger: optional object with `.cprint(str)` (the reference's utils.logger); it gets the
        reference's progress line every `log_every` episodes (eval_noise.py:94-95) and the
        per-class IoU printout at the end (:64-68)."""
        from . import ops
        mine = [episodes[i] for i in shard_indices(len(episodes), rank, world)]
        n_slots = len(self.test_classes) + 1
        counters = torch.zeros((3, n_slots), dtype=torch.int64, device=self.device)
        loss_sum = torch.zeros((), dtype=torch.float64, device=self.device)
        for s in range(0, len(mine), self.batch):
            sx, sy, qx, qy, slot = self._stage(mine[s:s + self.batch])
            out = self.model.forward_episodes(sx, sy, qx, qy, eval=self.eval_mdns)
            ops.confusion_accumulate(out["pred"], qy, slot, counters)
            loss_sum += out["loss"].double().sum()
            done = min(s + self.batch, len(mine))
            if logger is not None and done // log_every > s // log_every:
                from datetime import datetime
                logger.cprint("[Eval] Iter: %d | Loss: %.4f | %s" % (
                    done, float(out["loss"][-1]), str(datetime.now())))
        n = torch.tensor(float(len(mine)), dtype=torch.float64, device=self.device)
        all_reduce_eval_state(counters, loss_sum, n)
        res = iou_from_counters(counters)
        res.update(counters=counters.cpu(), mean_loss=float(loss_sum / n), n_episodes=int(n))
        if logger is not None and rank == 0:
            for c in range(n_slots):
                logger.cprint("class %d: iou %f" % (c, res["iou"][c]))
            logger.cprint("mean IoU: %f\n" % res["mean_iou"])
        return res


class _ItemEpisode:
    """Adapter: (data list, sampled_classes) as the reference's test loader yields them."""

    def __init__(self, data, sampled_classes):
        self.support_x, self.support_y, self.query_x, self.query_y = data[0], data[1], data[2], data[3]
        self.sampled_classes = np.asarray(sampled_classes).reshape(-1)


def test_few_shot(test_loader, learner, logger, test_classes, path=None, eval=False, batch: int = 16):
    """Drop-in for reference eval_noise.py:75-113: same arguments, returns (mean_loss, mean_IoU).
    `test_loader` yields (data, sampled_classes) per episode exactly as the reference's DataLoader
    (batch_size 1, `batch_test_task_collate_test`); episodes are run `batch` at a time through one
    C-ABI call each and — under torch.distributed — sharded over the ranks."""
    import torch.distributed as dist
    rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    episodes = [_ItemEpisode(data, classes) for data, classes in test_loader]
    learner.model.eval()
    ev = EpisodeEvaluator(learner.model, test_classes, batch=batch, eval_mdns=bool(eval))
    res = ev.run(episodes, rank, world, logger=logger)
    return res["mean_loss"], res["mean_iou"]
