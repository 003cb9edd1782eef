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


# ---- files (reference utils/checkpoint_util.py) --------------------------------------------------
def load_pretrain_checkpoint(model, pretrain_checkpoint_path):
    """Encoder weights of the pre-training stage into `model.encoder` (:10-23)."""
    if pretrain_checkpoint_path is None:
        raise ValueError("Pretrained checkpoint must be given.")
    params = torch.load(pretrain_checkpoint_path, map_location="cpu")["params"]
    own = model.state_dict()
    picked = {"encoder." + k: v for k, v in params.items() if "encoder." + k in own}
    own.update(picked)
    model.load_state_dict(own)     # copies in place: views into the flat buffers stay intact
    return model


def load_model_checkpoint(model, model_checkpoint_path, optimizer=None, mode="test"):
    """`<dir>/checkpoint.tar` (:26-45).  mode 'test' -> model; 'train' -> (model, optimizer)."""
    try:
        ck = torch.load(os.path.join(model_checkpoint_path, "checkpoint.tar"), map_location="cpu")
        start_iter, start_iou = ck["iteration"], ck["IoU"]
    except Exception:
        raise ValueError("Model checkpoint file must be correctly given (%s)." % model_checkpoint_path)
    model.load_state_dict(ck["model_state_dict"], strict=False)
    if mode == "test":
        print("Load model checkpoint at Iteration %d (IoU %f)..." % (start_iter, start_iou))
        return model
    try:
        optimizer.load_state_dict(ck["optimizer_state_dict"])
    except Exception:
        print("Checkpoint does not include optimizer state dict...")
    print("Resume from checkpoint at Iteration %d (IoU %f)..." % (start_iter, start_iou))
    return model, optimizer


def save_model_checkpoint(model, optimizer, output_path, iteration, iou, loss=0.0):
    """What the reference's training script writes at its best validation IoU
    (mpti_train_noise.py:137-144)."""
    os.makedirs(output_path, exist_ok=True)
    torch.save(dict(iteration=int(iteration), IoU=float(iou), loss=float(loss),
                    model_state_dict={k: v.detach().cpu().clone() for k, v in model.state_dict().items()},
                    optimizer_state_dict=optimizer.state_dict()),
               os.path.join(output_path, "checkpoint.tar"))
