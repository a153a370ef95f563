"""Trainer for the decoder -- the replacement of the reference's missing ``DeepLearning/lstm_trainer.ipynb``
(``.MISSING_LARGE_BLOBS``; SURVEY F2: optimiser, LR, epochs, split and label map of the original are unknown,
so this recipe is ours: mean cross-entropy, Adam, filename-prefix labels).

    python -m neural_speech_decoding_b200.trainer --data /path/to/EEG_data_collection --out model.pth \
        [--classes food,water,backgroundnoise] [--epochs 30] [--batch 64] [--lr 1e-3] [--bf16]
    torchrun --nproc-per-node 8 -m neural_speech_decoding_b200.trainer ...        # data-parallel, NCCL

Data: ``*.csv`` windows of ``[625, 8]`` (``%.7f``, no header; Neural_decoding_data_collector.py:129-139), label =
file-name prefix; or a ``.npz`` with ``X [N,T,C]`` and ``prefix [N]`` (tests/golden/eeg_windows.npz).  The output
is a plain 16-key ``state_dict`` that the reference's ``SimplePredictor`` loads with ``strict=True``.
"""
from __future__ import annotations

import argparse
import os
from pathlib import Path
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch

from .dp import DataParallelTrainer, shard_batch
from .optim import FusedAdam
from .lstm_eeg_model import EEG_LSTM

DEFAULT_CLASSES = ("food", "water", "backgroundnoise")      # CLASS_NAMES order, lstm_eeg_model.py:11


def load_windows(path: str, classes: Sequence[str]) -> Tuple[np.ndarray, np.ndarray]:
    """-> (X [N,T,C] float32, y [N] int64); files whose prefix is not in ``classes`` are skipped."""
    p = Path(path)
    idx = {c: i for i, c in enumerate(classes)}
    if p.suffix == ".npz":
        d = np.load(p)
        X, prefix = d["X"], [str(s) for s in d["prefix"]]
    else:
        files = sorted(p.glob("*.csv"))
        if not files:
            raise FileNotFoundError(f"no *.csv under {p}")
        X = np.stack([np.loadtxt(f, delimiter=",", dtype=np.float32) for f in files])
        prefix = [f.name.split("_")[0] for f in files]
    keep = [i for i, s in enumerate(prefix) if s in idx]
    return np.ascontiguousarray(X[keep], dtype=np.float32), np.array([idx[prefix[i]] for i in keep], dtype=np.int64)


def split_indices(n: int, val_frac: float, seed: int) -> Tuple[np.ndarray, np.ndarray]:
    perm = np.random.default_rng(seed).permutation(n)
    n_val = int(round(n * val_frac))
    return perm[n_val:], perm[:n_val]


@torch.no_grad()
def evaluate(model: EEG_LSTM, X: torch.Tensor, y: torch.Tensor, num_classes: int) -> Dict[str, object]:
    was_training = model.training
    model.eval()
    logits, _ = model.decode(X, want_probs=False)
    pred = logits.argmax(1)
    conf = torch.zeros((num_classes, num_classes), dtype=torch.int64)
    for t, p in zip(y.cpu().tolist(), pred.cpu().tolist()):
        conf[t, p] += 1
    model.train(was_training)
    return {"loss": torch.nn.functional.cross_entropy(logits, y).item(), "acc": (pred == y).float().mean().item(),
            "confusion": conf.tolist()}


def train(X: np.ndarray, y: np.ndarray, num_classes: int, epochs: int = 30, batch: int = 64, lr: float = 1e-3,
          val_frac: float = 0.2, seed: int = 0, bf16: bool = False, dropout: float = 0.60, device=None,
          log=print) -> Tuple[EEG_LSTM, List[Dict[str, object]]]:
    """Single- or multi-process (torch.distributed already initialised) training loop."""
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    device = device or torch.device("cuda", torch.cuda.current_device())
    torch.manual_seed(seed)                              # identical init on every rank
    model = EEG_LSTM(input_size=X.shape[2], num_classes=num_classes, dropout=dropout).to(device)
    if bf16:
        model.compute_dtype = torch.bfloat16
    # the init above is shared; the train-mode noise (RReLU slopes, dropout masks, the in-kernel dropout seed drawn
    # from the CPU generator) must differ between ranks, or the batch shards would see perfectly correlated noise
    torch.manual_seed(seed + 7919 * (rank + 1))
    tr_idx, va_idx = split_indices(len(X), val_frac, seed)
    Xd = X.to(device) if isinstance(X, torch.Tensor) else torch.from_numpy(X).to(device)       # GPU-ingested or numpy
    yd = y.to(device) if isinstance(y, torch.Tensor) else torch.from_numpy(y).to(device)
    # one-launch Adam on the GPU; plain torch Adam only where the parameters are not on a CUDA device (host-logic tests)
    opt = FusedAdam(model.parameters(), lr=lr) if torch.device(device).type == "cuda" else torch.optim.Adam(model.parameters(), lr=lr)
    trainer = DataParallelTrainer(model, opt, world_size=world)
    history = []
    for ep in range(epochs):
        model.train()
        order = np.random.default_rng(seed + 1 + ep).permutation(tr_idx)     # same order on every rank
        tot, nb = 0.0, 0
        for s in range(0, len(order), batch):
            gb = order[s:s + batch]
            mine = gb[shard_batch(len(gb), rank, world)]
            if len(mine) == 0:                          # keep the collective in step
                mine = gb[:1]
                scale = 0.0
            else:
                scale = 1.0
            ix = torch.from_numpy(mine).to(device)
            xb, yb = Xd[ix], yd[ix]
            loss = trainer.step([(xb, yb)], global_batch=len(gb) / scale if scale else float("inf"))
            tot += loss.item()
            nb += 1
        rec = {"epoch": ep + 1, "train_loss": tot / max(nb, 1)}
        if len(va_idx):
            rec.update({"val_" + k: v for k, v in evaluate(model, Xd[torch.from_numpy(va_idx).to(device)],
                                                         yd[torch.from_numpy(va_idx).to(device)], num_classes).items()})
        history.append(rec)
        if rank == 0 and log:
            log({k: (round(v, 4) if isinstance(v, float) else v) for k, v in rec.items() if k != "val_confusion"})
    return model, history


def main(argv=None) -> None:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--data", required=True)
    ap.add_argument("--out", default="lstm_classifier.pth")
    ap.add_argument("--classes", default=",".join(DEFAULT_CLASSES))
    ap.add_argument("--epochs", type=int, default=30)
    ap.add_argument("--batch", type=int, default=64, help="global batch")
    ap.add_argument("--lr", type=float, default=1e-3)
    ap.add_argument("--val-frac", type=float, default=0.2)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--bf16", action="store_true", help="tensor-core tier (tcgen05) for forward and backward")
    args = ap.parse_args(argv)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        import torch.distributed as dist
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        from .dp import bind_to_gpu_numa
        bind_to_gpu_numa(local)                     # host buffers next to this rank's GPU
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    classes = [c.strip() for c in args.classes.split(",") if c.strip()]
    if Path(args.data).is_dir():
        # CSV directory: one H2D copy of the raw bytes + the GPU parser (ingest.py), not one np.loadtxt per file
        from .ingest import load_labelled_dir
        X, y, _ = load_labelled_dir(args.data, classes, torch.device("cuda", torch.cuda.current_device()))
    else:
        X, y = load_windows(args.data, classes)
    model, _ = train(X, y, len(classes), args.epochs, args.batch, args.lr, args.val_frac, args.seed, args.bf16)
    if int(os.environ.get("RANK", "0")) == 0:
        torch.save({k: v.detach().cpu() for k, v in model.state_dict().items()}, args.out)
        print(f"saved {args.out} ({len(classes)} classes: {classes})")
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
