"""Checkpoint files in the reference's format (reference utils/checkpoint_util.py:10-75) so that a
run can be resumed on either side:

  pre-training:   `{'params': encoder.state_dict()}`                      (load_pretrain_checkpoint)
  meta-training:  `<dir>/checkpoint.tar` = `{'iteration', 'IoU', 'model_state_dict',
                  'optimizer_state_dict'}`                                (load/save_model_checkpoint)

`optimizer_state_dict` is torch.optim.Adam's over the reference's four parameter groups (encoder,
base_learner, att_learner, proj — models/mpti_learner.py:26-32).  The fused optimizer of this repo
keeps ONE flat first/second-moment buffer; `adam_state_to_reference` / `adam_state_from_reference`
convert between the two layouts (pure tensor code, no device work).
"""
from __future__ import annotations

import os
from typing import Dict, List, Sequence, Tuple

import torch

GROUP_PREFIXES = ("encoder.", "base_learner.", "att_learner.", "proj.")


def _group_index_lists(names: Sequence[str]) -> List[List[int]]:
    groups: List[List[int]] = [[] for _ in GROUP_PREFIXES]
    for i, n in enumerate(names):
        for g, pre in enumerate(GROUP_PREFIXES):
            if n.startswith(pre):
                groups[g].append(i)
                break
        else:
            raise ValueError(f"parameter {n} belongs to none of the reference's optimizer groups")
    flat = [i for g in groups for i in g]
    if flat != list(range(len(names))):
        raise ValueError("parameter order differs from the reference's optimizer order")
    return groups


def adam_state_to_reference(exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor, step: int,
                            names: Sequence[str], shapes: Sequence[torch.Size],
                            offsets: Sequence[int], lrs: Sequence[float], betas=(0.9, 0.999),
                            eps: float = 1e-8) -> Dict:
    """Flat moment buffers -> `torch.optim.Adam.state_dict()` of the reference's optimizer.
    lrs: (encoder lr, lr of the other three groups)."""
    groups = _group_index_lists(names)
    state = {}
    if step > 0:
        for i, (shape, o) in enumerate(zip(shapes, offsets)):
            n = int(torch.Size(shape).numel())
            state[i] = {"step": torch.tensor(float(step)),
                        "exp_avg": exp_avg[o:o + n].detach().clone().view(shape).cpu(),
                        "exp_avg_sq": exp_avg_sq[o:o + n].detach().clone().view(shape).cpu()}
    pgs = []
    for g, idx in enumerate(groups):
        pgs.append({"lr": float(lrs[0] if g == 0 else lrs[1]), "betas": tuple(betas), "eps": eps,
                    "weight_decay": 0, "amsgrad": False, "params": idx})
    return {"state": state, "param_groups": pgs}


def adam_state_from_reference(sd: Dict, names: Sequence[str], shapes: Sequence[torch.Size],
                              offsets: Sequence[int], like: torch.Tensor
                              ) -> Tuple[torch.Tensor, torch.Tensor, int, Tuple[float, float]]:
    """`torch.optim.Adam.state_dict()` of the reference's optimizer -> (exp_avg, exp_avg_sq, step,
    (encoder lr, other lr)) in the flat layout; `like` gives dtype/device/size of the flat buffer."""
    groups = _group_index_lists(names)
    pgs = sd["param_groups"]
    if [len(p["params"]) for p in pgs] != [len(g) for g in groups]:
        raise ValueError("optimizer state was saved for a different parameter grouping")
    m, v = torch.zeros_like(like), torch.zeros_like(like)
    steps = set()
    # saved indices follow the saved groups' order; map position -> our parameter index
    order = [i for p in pgs for i in p["params"]]
    for pos, saved_idx in enumerate(order):
        st = sd["state"].get(saved_idx)
        if st is None:
            continue
        shape, o = shapes[pos], offsets[pos]
        n = int(torch.Size(shape).numel())
        if tuple(st["exp_avg"].shape) != tuple(shape):
            raise ValueError(f"optimizer state of {names[pos]} has shape {tuple(st['exp_avg'].shape)}")
        m[o:o + n] = st["exp_avg"].reshape(-1).to(m)
        v[o:o + n] = st["exp_avg_sq"].reshape(-1).to(v)
        steps.add(int(float(st["step"])))
    if len(steps) > 1:
        raise ValueError(f"parameters were stepped a different number of times: {sorted(steps)}")
    return m, v, (steps.pop() if steps else 0), (float(pgs[0]["lr"]), float(pgs[1]["lr"]))


# ---- files --------------------------------------------------------------------------------------
# File layouts are the reference's (utils/checkpoint_util.py, mpti_train_noise.py:137-144): the
# pre-training file is {'params': encoder state_dict without the 'encoder.' prefix}; a run directory
# holds `checkpoint.tar` = {iteration, IoU, loss, model_state_dict, optimizer_state_dict}.
class CheckpointError(ValueError):
    pass


def read_checkpoint(run_dir: str) -> dict:
    """`<run_dir>/checkpoint.tar` as a dict, on the CPU, with its required entries verified."""
    path = os.path.join(run_dir, "checkpoint.tar")
    if not os.path.isfile(path):
        raise CheckpointError(f"no checkpoint.tar under {run_dir!r}")
    ck = torch.load(path, map_location="cpu")
    missing = [k for k in ("iteration", "IoU", "model_state_dict") if k not in ck]
    if missing:
        raise CheckpointError(f"{path}: entries {missing} are missing")
    return ck


def restore(model, ck: dict, optimizer=None) -> dict:
    """Weights (non-strict, like the reference: older files lack `proj.*`) and, when an optimizer is
    given and the file carries its state, the Adam moments.  Copies in place, so parameters that are
    views into the flat training buffers stay views."""
    model.load_state_dict(ck["model_state_dict"], strict=False)
    had_opt = False
    if optimizer is not None and ck.get("optimizer_state_dict") is not None:
        optimizer.load_state_dict(ck["optimizer_state_dict"])
        had_opt = True
    return {"iteration": int(ck["iteration"]), "iou": float(ck["IoU"]), "optimizer_restored": had_opt}


def load_encoder_pretraining(model, path: str) -> int:
    """Copies every tensor of the pre-training file that has an `encoder.<name>` twin in `model`;
    returns how many were copied."""
    if not path or not os.path.isfile(path):
        raise CheckpointError(f"pre-training file not found: {path!r}")
    blob = torch.load(path, map_location="cpu")
    if "params" not in blob:
        raise CheckpointError(f"{path}: no 'params' entry")
    own = model.state_dict()
    n = 0
    with torch.no_grad():
        for name, value in blob["params"].items():
            dst = own.get("encoder." + name)
            if dst is not None and dst.shape == value.shape:
                dst.copy_(value)
                n += 1
    return n


# the reference's entry-point names, for scripts written against utils/checkpoint_util.py
def load_pretrain_checkpoint(model, pretrain_checkpoint_path):
    load_encoder_pretraining(model, pretrain_checkpoint_path)
    return model


def load_model_checkpoint(model, model_checkpoint_path, optimizer=None, mode="test"):
    info = restore(model, read_checkpoint(model_checkpoint_path),
                   optimizer if mode == "train" else None)
    print("checkpoint of iteration %d (IoU %.4f) loaded%s" % (
        info["iteration"], info["iou"],
        "" if mode != "train" else (", optimizer state restored" if info["optimizer_restored"]
                                    else ", optimizer starts fresh")))
    return model if mode == "test" else (model, optimizer)


def save_model_checkpoint(model, optimizer, output_path, iteration, iou, loss=0.0):
    """What the reference's training script writes at its best validation IoU
    (mpti_train_noise.py:137-144)."""
    os.makedirs(output_path, exist_ok=True)
    torch.save(dict(iteration=int(iteration), IoU=float(iou), loss=float(loss),
                    model_state_dict={k: v.detach().cpu().clone() for k, v in model.state_dict().items()},
                    optimizer_state_dict=optimizer.state_dict()),
               os.path.join(output_path, "checkpoint.tar"))
