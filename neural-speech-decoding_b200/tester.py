"""Drop-in for ``Neuro-Alpha-App/Utilities/tester.py``: ``run_trials`` / ``TrialResult`` with the
reference's signature and its 10-trial probability averaging (tester.py:30-110), plus the batched
sibling ``run_trials_batched`` that the reference does not have (B independent sessions at once).

The acquisition side (``StreamingProcess``: BrainFlow + serial hardware) is out of scope and is
imported from the host application, exactly as the reference does.  The averaging itself --
fp32 zeros, ``+=`` in arrival order, one division (tester.py:54,89,97) -- runs in the K5 kernel,
which reproduces that rounding bit for bit.
"""
from __future__ import annotations

import time
from dataclasses import dataclass
from multiprocessing import Queue
from pathlib import Path
from typing import Optional

import numpy as np
import torch

from . import ops
from .lstm_eeg_model import SimplePredictor

DEFAULT_SERIAL = "/dev/cu.usbserial-FTB6SPL3"                      # tester.py:17
DEFAULT_MODEL = str(Path(__file__).resolve().parent / "LSTM_Model" / "lstm_classifier_Water_Food_Bg_Noise.pth")

StreamingProcess = None   # resolved lazily from the host app; tests / apps may assign a producer class


@dataclass
class TrialResult:            # tester.py:23-27
    trials: int
    avg_probs: Optional[np.ndarray]
    avg_chunk: Optional[np.ndarray] = None


def _producer_class():
    if StreamingProcess is not None:
        return StreamingProcess
    for mod in ("Utilities.streaming_process", "streaming_process"):
        try:
            return __import__(mod, fromlist=["StreamingProcess"]).StreamingProcess
        except ImportError:
            continue
    raise ImportError("StreamingProcess (reference Utilities/streaming_process.py) is not importable; "
                      "assign neural_speech_decoding_b200.tester.StreamingProcess to a producer class")


def device_trial_mean(stack: np.ndarray) -> np.ndarray:
    """[R, ...] float32 host array -> mean over trials on the GPU (K5), returned on the host."""
    t = torch.from_numpy(np.ascontiguousarray(stack, dtype=np.float32)).to(ops.compute_device(torch.device("cpu")))
    return ops.trial_mean(t).cpu().numpy()


def run_trials(trials: int = 10, serial_port: str = DEFAULT_SERIAL, num_channels: int = 8,
               window_seconds: float = 5.0, model_path: str = DEFAULT_MODEL, verbose: bool = True) -> TrialResult:
    """Collect ``trials`` windows from the producer, classify each, return the averaged
    probabilities and the averaged chunk.  Same contract as tester.py:30-110."""
    q = Queue(maxsize=8)
    producer = _producer_class()(serial_port=serial_port, num_channels=num_channels,
                                 window_seconds=window_seconds, out_queue=q)
    producer.start()
    producer.recording_flag.value = True

    predictor = None
    probs_seen, chunks_seen = [], []
    try:
        while len(probs_seen) < trials:
            if not producer.is_alive():
                raise RuntimeError("Producer exited unexpectedly")
            try:
                item = q.get(timeout=6.5)
            except Exception:
                if verbose:
                    print("Waiting for chunk...", flush=True)
                continue
            chunk = np.asarray(item["data"])
            if predictor is None:
                predictor = SimplePredictor(pth_path=model_path, sr=item["sr"], channel_order=item.get("channels"),
                                            input_size=num_channels, hidden_size=48, num_layers=2, num_classes=3,
                                            dropout=0.60, device="cpu", tailoring_lambda=1.25e-29,
                                            class_names=["Food", "Water", "None"])
            probs, label = predictor.predict(chunk)
            probs_seen.append(probs)
            chunks_seen.append(chunk)
            if verbose:
                stamp = time.strftime("%H:%M:%S")
                print(f"[Trial {len(probs_seen):02d} @ {stamp}] pred={label} probs={np.round(probs, 3)}")

        collected = len(probs_seen)
        avg_probs = device_trial_mean(np.stack(probs_seen)) if collected else None
        avg_chunk = device_trial_mean(np.stack(chunks_seen)) if collected else None
        if verbose:
            if avg_probs is not None:
                print(f"\nAveraged over {collected} trials: {np.round(avg_probs, 3)}")
                print(f"Averaged chunk shape: {avg_chunk.shape}")
            else:
                print("No trials completed; no average available.")
        return TrialResult(trials=collected, avg_probs=avg_probs, avg_chunk=avg_chunk)
    finally:
        producer.recording_flag.value = False
        producer.stop()
        producer.join(timeout=5.0)


def run_trials_batched(windows, model, return_device: bool = False, chunk_trials: int = 0):
    """B sessions x R trials in one go.

    ``windows``: ``[R,B,T,C]`` float32 -- a CUDA tensor, a CPU tensor (pinned or not) or a numpy array.
    Every trial is one forward of the B windows; class probabilities are averaged over the R trials
    in trial order with run_trials' rounding (K5).  Host inputs are copied in groups of ``chunk_trials``
    whole trials (default: enough trials for ~8k windows) on a side stream, overlapped with the decode.
    Returns ``avg_probs [B,K]`` (numpy unless ``return_device``).
    """
    if isinstance(windows, np.ndarray):
        windows = torch.from_numpy(np.ascontiguousarray(windows, dtype=np.float32))
    if windows.dim() != 4:
        raise ValueError("run_trials_batched expects windows of shape [R,B,T,C]")
    R, B, T, C = windows.shape
    dev = next(model.parameters()).device
    if dev.type != "cuda":
        raise RuntimeError("run_trials_batched: the model must live on a CUDA device (no CPU fallback)")
    m = model
    with torch.inference_mode():
        NC = m.fc[3].out_features
        probs_all = torch.empty((R, B, NC), dtype=torch.float32, device=dev)
        if windows.is_cuda:
            flat = windows.reshape(R * B, T, C)
            _, p = m.decode(flat, want_probs=True)
            probs_all = p.reshape(R, B, NC)
        else:
            # Host windows: stream them in groups of whole trials, H2D on a side stream (two device buffers),
            # so the copy of group g+1 overlaps the decode of group g.  A decode launch costs one full
            # 625-step round (~1.8 ms on B200) however few windows it holds, so groups are sized to carry
            # at least ~8k windows: then the pipeline is bound by the PCIe copy, not by launch rounds.
            g = chunk_trials if chunk_trials > 0 else max(1, min(R, -(-8192 // max(B, 1))))
            # group plan: full groups of g trials, but the LAST group is a single trial -- nothing overlaps the decode of
            # the last group (the pipeline's tail), and a short launch costs only the dependent-chain latency
            starts = list(range(0, R, g))
            if chunk_trials <= 0 and R > 1 and (R - starts[-1]) > 1:
                starts.append(R - 1)
            bounds = list(zip(starts, starts[1:] + [R]))
            main = torch.cuda.current_stream(dev)
            side = torch.cuda.Stream(device=dev)
            bufs = [torch.empty((g, B, T, C), dtype=torch.float32, device=dev) for _ in range(2)]
            ready = [torch.cuda.Event() for _ in range(2)]
            freed = [torch.cuda.Event() for _ in range(2)]
            side.wait_stream(main)
            for i, (r0, r1) in enumerate(bounds):
                k, n = i & 1, r1 - r0
                with torch.cuda.stream(side):
                    if i >= 2:
                        side.wait_event(freed[k])
                    bufs[k][:n].copy_(windows[r0:r0 + n], non_blocking=True)
                    ready[k].record(side)
                main.wait_event(ready[k])
                _, p = m.decode(bufs[k][:n].reshape(n * B, T, C), want_probs=True)
                probs_all[r0:r0 + n].copy_(p.reshape(n, B, NC))
                freed[k].record(main)
        avg = ops.trial_mean(probs_all)
    return avg if return_device else avg.cpu().numpy()
