"""Episode files and staging: the step right before the hot path.

The reference stores every pre-sampled test episode as one `.h5` file with eight datasets
(reference dataloaders/loader.py:1687-1721) and collates it with two transposes that leave the
clouds point-major in memory (`batch_test_task_collate_test`, :1676-1684).  Here:

* `write_episode` / `read_episode` keep that schema (same dataset names, dtypes and shapes).  The
  container is HDF5 when `h5py` is importable (it is not in this image) and a NumPy `.npz` with the
  same keys otherwise — the arrays are identical either way; `.r3ep` is the same eight datasets as
  one flat, 64-byte-aligned file that reader threads `readinto()` pinned staging memory directly
  (`convert_folder` converts a folder once);
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
    if out_filename.endswith(RAW_EXT):
        return write_episode_raw(out_filename, data)
    arrays = {k: np.asarray(v, dtype=SCHEMA[k]) for k, v in zip(ORDER, data)}
    if _h5 is not None and out_filename.endswith(".h5"):
        with _h5.File(out_filename, "w") as f:
            for k, v in arrays.items():
                f.create_dataset(k, data=v, dtype=SCHEMA[k])
        return out_filename
    path = os.path.splitext(out_filename)[0] + ".npz"
    np.savez(path, **arrays)
    return path


# ---- raw container -------------------------------------------------------------------------------
# The same eight datasets (names, dtypes, shapes of the reference's schema) as ONE flat file:
#   8-byte magic | uint32 header length | JSON header {name: [dtype, shape, byte offset]} | arrays,
# each 64-byte aligned.  A reader thread `readinto()`s an array straight into the pinned staging
# buffer the H2D copy starts from — no parsing, no intermediate copy (the .h5 / .npz containers
# need both).  `convert_folder` turns a folder of .h5 / .npz episodes into this form once.
RAW_MAGIC = b"R3EP0001"
RAW_EXT = ".r3ep"


def write_episode_raw(out_filename: str, data: Sequence[np.ndarray]) -> str:
    import json
    arrays = {k: np.ascontiguousarray(np.asarray(v, dtype=SCHEMA[k])) for k, v in zip(ORDER, data)}
    path = os.path.splitext(out_filename)[0] + RAW_EXT
    # two passes: offsets depend on the header length
    header, blob_len = {}, 0
    for guess in range(2):
        hdr = json.dumps(header).encode() if header else b"{}" + b" " * 1024
        start = (8 + 4 + len(hdr) + 63) // 64 * 64
        off, header = start, {}
        for k, v in arrays.items():
            header[k] = [SCHEMA[k], list(v.shape), off]
            off = (off + v.nbytes + 63) // 64 * 64
        blob_len = off
        hdr2 = json.dumps(header).encode()
        if (8 + 4 + len(hdr2) + 63) // 64 * 64 == start:
            break
    hdr = json.dumps(header).encode()
    with open(path, "wb") as f:
        f.write(RAW_MAGIC)
        f.write(np.uint32(len(hdr)).tobytes())
        f.write(hdr)
        for k, v in arrays.items():
            f.seek(header[k][2])
            f.write(v.tobytes())
        f.truncate(blob_len)
    return path


def _raw_header(f) -> Dict[str, list]:
    import json
    if f.read(8) != RAW_MAGIC:
        raise ValueError("not an R3EP episode file")
    n = int(np.frombuffer(f.read(4), dtype=np.uint32)[0])
    return json.loads(f.read(n).decode())


def read_episode_raw(file_name: str) -> Tuple[np.ndarray, ...]:
    with open(file_name, "rb") as f:
        hdr = _raw_header(f)
        out = []
        for k in ORDER:
            dt, shape, off = hdr[k]
            f.seek(off)
            a = np.empty(shape, dtype=dt)
            f.readinto(memoryview(a).cast("B"))
            out.append(a)
    return tuple(out)


def read_episode_raw_into(file_name: str, dst: Dict[str, np.ndarray]) -> np.ndarray:
    """Reads the datasets named in `dst` straight into those (C-contiguous, writable) arrays —
    e.g. rows of pinned staging tensors — and returns sampled_classes."""
    with open(file_name, "rb") as f:
        hdr = _raw_header(f)
        for k, a in dst.items():
            dt, shape, off = hdr[k]
            if a.dtype != np.dtype(dt) or list(a.shape) != list(shape):
                raise ValueError(f"{file_name}: {k} is {dt}{shape}, staging buffer is "
                                 f"{a.dtype}{list(a.shape)}")
            f.seek(off)
            f.readinto(memoryview(a).cast("B"))
        dt, shape, off = hdr["sampled_classes"]
        f.seek(off)
        return np.frombuffer(f.read(int(np.prod(shape)) * np.dtype(dt).itemsize), dtype=dt).copy()


def convert_folder(src: str, dst: str) -> int:
    """.h5 / .npz episodes of `src` -> raw episode files in `dst`; returns the number converted."""
    os.makedirs(dst, exist_ok=True)
    n = 0
    for name in sorted(os.listdir(src)):
        if name.endswith((".h5", ".npz")):
            write_episode_raw(os.path.join(dst, name), read_episode(os.path.join(src, name)))
            n += 1
    return n


def read_episode(file_name: str) -> Tuple[np.ndarray, ...]:
    """reference dataloaders/loader.py:1709-1721: the 8-tuple in ORDER."""
    if file_name.endswith(RAW_EXT):
        return read_episode_raw(file_name)
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
        names = sorted(n for n in os.listdir(folder) if n.endswith((".h5", ".npz", RAW_EXT)))
        self.file_names = [os.path.join(folder, n) for n in names]

    def __len__(self):
        return len(self.file_names)

    def __getitem__(self, index):
        return read_episode(self.file_names[index])

    def item(self, index):
        """(data, sampled_classes) of episode `index` — what iterating yields; thread-safe, so the
        evaluation driver reads several files at once."""
        return collate_test(self[index])

    def read_into(self, index, dst: Dict[str, np.ndarray]) -> np.ndarray:
        """Episode `index` into staging arrays keyed by dataset name (point-major, the on-disk
        layout); returns sampled_classes.  Raw files are read in place, the others via a copy."""
        name = self.file_names[index]
        if name.endswith(RAW_EXT):
            return read_episode_raw_into(name, dst)
        item = dict(zip(ORDER, read_episode(name)))
        for k, a in dst.items():
            np.copyto(a, item[k].astype(a.dtype, copy=False).reshape(a.shape))
        return np.asarray(item["sampled_classes"], np.int32)

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
