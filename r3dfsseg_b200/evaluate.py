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

    def run(self, episodes: Sequence, rank: int = 0, world: int = 1) -> Dict[str, object]:
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
        n = torch.tensor(float(len(mine)), dtype=torch.float64, device=self.device)
        all_reduce_eval_state(counters, loss_sum, n)
        res = iou_from_counters(counters)
        res.update(counters=counters.cpu(), mean_loss=float(loss_sum / n), n_episodes=int(n))
        return res
