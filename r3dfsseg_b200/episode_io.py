"""Episode files and staging: the step right before the hot path.

The reference stores every pre-sampled test episode as one `.h5` file with eight datasets
(reference dataloaders/loader.py:1687-1721) and collates it with two transposes that leave the
clouds point-major in memory (`batch_test_task_collate_test`, :1676-1684).  Here:

* `write_episode` / `read_episode` keep that schema (same dataset names, dtypes and shapes).  The
  container is HDF5 when `h5py` is importable (it is not in this image) and a NumPy `.npz` with the
  same keys otherwise — the arrays are identical either way;
* `collate_test` is the reference's collate: tensors as `MPTILearner_V3.test` expects them;
* `stage_batch` packs a list of episodes into ONE pinned, point-major host buffer per tensor, the
  layout `r3dfs_mpti_forward` consumes directly (E episodes per call, no per-episode H2D copies,
  no stride fix-ups on the device).
"""
from __future__ import annotations

import os
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch

try:  # pragma: no cover - h5py is absent from the build image
    import h5py as _h5
except ImportError:  # noqa: D401
    _h5 = None

SCHEMA: Dict[str, str] = {
    "support_ptclouds": "float32",   # (n_way, k_shot, N, 9)
    "support_masks": "int32",        # (n_way, k_shot, N)
    "query_ptclouds": "float32",     # (n_way * n_queries, N, 9)
    "query_labels": "int64",         # (n_way * n_queries, N)
    "sampled_classes": "int32",      # (n_way,)
    "support_clusters": "int32",     # (n_way, k_shot, N)  segment labels (unused by MPTI)
    "query_clusters": "int32",       # (n_way * n_queries, N)
    "gt_support_masks": "int32",     # (n_way, k_shot, N)
}
ORDER = list(SCHEMA)


def episode_arrays(ep) -> Tuple[np.ndarray, ...]:
    """r3dfsseg_b200.episodes.Episode -> the 8-tuple the reference's write_episode takes."""
    sx = ep.support_x.transpose(2, 3).contiguous().numpy()
    qx = ep.query_x.transpose(1, 2).contiguous().numpy()
    zs = np.zeros(ep.support_y.shape, np.int32)
    zq = np.zeros(tuple(ep.query_y.shape), np.int32)
    return (sx, ep.support_y.numpy(), qx, ep.query_y.numpy(),
            np.asarray(ep.sampled_classes, np.int32), zs, zq, ep.gt_support_y.numpy())


def write_episode(out_filename: str, data: Sequence[np.ndarray]) -> str:
    """reference dataloaders/loader.py:1687-1706.  Returns the path written (the extension is
    switched to .npz when HDF5 is unavailable)."""
    arrays = {k: np.asarray(v, dtype=SCHEMA[k]) for k, v in zip(ORDER, data)}
    if _h5 is not None and out_filename.endswith(".h5"):
        with _h5.File(out_filename, "w") as f:
            for k, v in arrays.items():
                f.create_dataset(k, data=v, dtype=SCHEMA[k])
        return out_filename
    path = os.path.splitext(out_filename)[0] + ".npz"
    np.savez(path, **arrays)
    return path


def read_episode(file_name: str) -> Tuple[np.ndarray, ...]:
    """reference dataloaders/loader.py:1709-1721: the 8-tuple in ORDER."""
    if file_name.endswith(".h5"):
        if _h5 is None:
            raise RuntimeError("reading .h5 episodes needs h5py (not installed); use the .npz twin")
        with _h5.File(file_name, "r") as f:
            return tuple(f[k][:] for k in ORDER)
    with np.load(file_name) as f:
        return tuple(f[k] for k in ORDER)


def collate_test(item: Sequence[np.ndarray]):
    """`batch_test_task_collate_test` (reference dataloaders/loader.py:1676-1684) for one episode:
    -> (data list of 7 tensors, sampled_classes).  The transposes are views: memory stays
    point-major, which is what the kernels read."""
    sx, sy, qx, qy, classes, sc, qc, gy = item
    data = [torch.from_numpy(sx).transpose(2, 3), torch.from_numpy(sy),
            torch.from_numpy(qx).transpose(1, 2), torch.from_numpy(qy.astype(np.int64)),
            torch.from_numpy(sc), torch.from_numpy(qc), torch.from_numpy(gy)]
    return data, classes


class EpisodeFolder:
    """The reference's `MyTestDataset` (dataloaders/loader.py:1640-1660) over a directory of
    episode files; iterating yields what its DataLoader yields: (data, sampled_classes)."""

    def __init__(self, folder: str):
        names = sorted(n for n in os.listdir(folder) if n.endswith((".h5", ".npz")))
        self.file_names = [os.path.join(folder, n) for n in names]

    def __len__(self):
        return len(self.file_names)

    def __getitem__(self, index):
        return read_episode(self.file_names[index])

    def item(self, index):
        """(data, sampled_classes) of episode `index` — what iterating yields; thread-safe, so the
        evaluation driver reads several files at once."""
        return collate_test(self[index])

    def __iter__(self):
        for i in range(len(self)):
            yield collate_test(self[i])


def stage_batch(items: Sequence[Sequence[np.ndarray]], pin: bool = True):
    """E episode 8-tuples -> pinned host tensors (support (E, n_way, k_shot, N, 9), masks,
    query (E, n_q, N, 9), labels, classes (E, n_way)), ready for one non-blocking H2D copy and one
    `forward_episodes` call (pass `x.transpose(-1, -2)` views to keep the reference's (.., 9, N)
    argument convention without moving memory)."""
    def cat(i, dtype):
        t = torch.from_numpy(np.stack([np.asarray(it[i]) for it in items]).astype(dtype, copy=False))
        return t.pin_memory() if pin and torch.cuda.is_available() else t
    return (cat(0, np.float32), cat(1, np.int32), cat(2, np.float32), cat(3, np.int64),
            torch.from_numpy(np.stack([np.asarray(it[4], np.int32) for it in items])))
