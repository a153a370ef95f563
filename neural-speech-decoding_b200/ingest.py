"""Ingestion of recorded EEG for the decoder: the collector's CSV windows and continuous streams.

The reference stores every 5-second window as text -- ``np.savetxt(f, data, delimiter=",", fmt="%.7f")`` of a
``[625, 8]`` array, one file per window, label = file-name prefix (``Neural_decoding_data_collector.py:129-139``) --
and reads it back with ``np.loadtxt`` one file at a time.  Here the raw bytes of all files are read into ONE pinned
host buffer, copied to the GPU once, and parsed there (``na_csv_parse_f32``: bit-identical to ``np.loadtxt(...,
dtype=np.float32)``); a continuous ``[n_samples, C]`` recording is cut into (overlapping) windows by kernel K1
(``streaming_process.py:35-58`` emits non-overlapping 5-second windows: ``hop = window``).  There is no CPU fallback.
"""
from __future__ import annotations

from pathlib import Path
from typing import Iterable, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import _lib, ops

PathLike = Union[str, Path]


def read_files_pinned(paths: Sequence[PathLike]) -> Tuple[torch.Tensor, torch.Tensor]:
    """Raw bytes of ``paths`` back to back in one pinned uint8 tensor + int64 offsets ``[len(paths) + 1]``."""
    paths = [Path(p) for p in paths]
    if not paths:
        raise ValueError("read_files_pinned: no files")
    sizes = [p.stat().st_size for p in paths]
    offsets = np.zeros(len(paths) + 1, dtype=np.int64)
    np.cumsum(sizes, out=offsets[1:])
    pin = torch.cuda.is_available()
    buf = torch.empty((int(offsets[-1]),), dtype=torch.uint8, pin_memory=pin)
    view = buf.numpy()
    for p, a, b in zip(paths, offsets[:-1], offsets[1:]):
        with open(p, "rb") as f:
            got = f.readinto(memoryview(view[a:b]))
        if got != b - a:
            raise IOError(f"{p}: short read ({got} of {b - a} bytes)")
    return buf, torch.from_numpy(offsets)


def parse_csv_bytes(text: torch.Tensor, offsets: torch.Tensor, rows: int = 625, cols: int = 8,
                    names: Optional[Sequence[str]] = None) -> torch.Tensor:
    """``text`` uint8 (CUDA), ``offsets`` int64 ``[N + 1]`` (CUDA) -> fp32 ``[N, rows, cols]`` on the same device.
    Raises ``ValueError`` (numpy's error for a malformed file) if a file does not hold exactly rows x cols plain
    decimal fields."""
    ops._require_cuda(text, offsets)
    if text.dtype != torch.uint8 or offsets.dtype != torch.int64 or offsets.dim() != 1 or offsets.numel() < 2:
        raise RuntimeError("parse_csv_bytes: text must be uint8, offsets int64 [N + 1]")
    n = offsets.numel() - 1
    sizes = (offsets[1:] - offsets[:-1])
    max_bytes = int(sizes.max().item())
    if int(sizes.min().item()) < 0 or int(offsets[-1].item()) > text.numel():
        raise RuntimeError("parse_csv_bytes: offsets are not a partition of text")
    out = torch.empty((n, rows, cols), dtype=torch.float32, device=text.device)
    status = torch.empty((n, 2), dtype=torch.int32, device=text.device)
    with torch.cuda.device(text.device):
        _lib.call("na_csv_parse_f32", text.data_ptr(), offsets.data_ptr(), out.data_ptr(), status.data_ptr(), n, rows * cols,
                  max(1, max_bytes), ops._stream())
    st = status.cpu().numpy()
    bad = np.nonzero((st[:, 0] != rows * cols) | (st[:, 1] != 0))[0]
    if bad.size:
        i = int(bad[0])
        who = names[i] if names is not None else f"file {i}"
        raise ValueError(f"{who}: expected {rows}x{cols} = {rows * cols} plain decimal fields, found {int(st[i, 0])} "
                         f"({int(st[i, 1])} unparsable); {bad.size} bad file(s) in total")
    return out


def load_csv_windows(paths: Sequence[PathLike], device: Union[str, torch.device] = "cuda", rows: int = 625,
                     cols: int = 8) -> torch.Tensor:
    """The reference's ``np.stack([np.loadtxt(p, delimiter=",", dtype=np.float32) for p in paths])`` as one H2D copy of
    the raw bytes + one GPU kernel.  Returns fp32 ``[len(paths), rows, cols]`` on ``device``."""
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("load_csv_windows: the parser runs on the GPU (no CPU fallback); pass a cuda device")
    buf, offsets = read_files_pinned(paths)
    text = buf.to(device, non_blocking=True)
    return parse_csv_bytes(text, offsets.to(device), rows, cols, [str(p) for p in paths])


def load_labelled_dir(path: PathLike, classes: Sequence[str], device: Union[str, torch.device] = "cuda",
                      rows: int = 625, cols: int = 8) -> Tuple[torch.Tensor, torch.Tensor, List[str]]:
    """All ``<class>_*.csv`` windows under ``path`` whose prefix is in ``classes`` (label = index in ``classes``,
    ``CLASS_NAMES`` order of ``lstm_eeg_model.py:11``); other prefixes (e.g. yes/no for the 3-class model) are skipped."""
    files = sorted(Path(path).glob("*.csv"))
    keep = [(f, classes.index(f.name.split("_")[0])) for f in files if f.name.split("_")[0] in classes]
    if not keep:
        raise FileNotFoundError(f"no *.csv with a prefix in {list(classes)} under {path}")
    X = load_csv_windows([f for f, _ in keep], device, rows, cols)
    y = torch.tensor([l for _, l in keep], dtype=torch.int64, device=X.device)
    return X, y, [f.name for f, _ in keep]


def windows_from_stream(samples: torch.Tensor, window: int = 625, hop: Optional[int] = None,
                        zscore: bool = False) -> torch.Tensor:
    """Continuous recording ``[n_samples, C]`` (CUDA) -> windows ``[B, window, C]`` starting every ``hop`` samples
    (default ``hop = window``: the non-overlapping 5-second windows of ``streaming_process.py:35-58``); optional
    per-window per-channel z-score (``Frontend/app.py:166-170``).  One pass of kernel K1."""
    if samples.dim() != 2:
        raise ValueError(f"windows_from_stream expects [n_samples, channels], got {tuple(samples.shape)}")
    hop = window if hop is None else int(hop)
    if hop < 1:
        raise ValueError("hop must be >= 1")
    return ops.window_zscore(samples, int(window), hop, bool(zscore), False, ops.NA_F32)
