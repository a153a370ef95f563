#!/usr/bin/env python
"""Generate tests/golden/* by running the REAL reference (TEST INFRASTRUCTURE).

Runs only in the build container, where ``/root/reference`` is mounted read-only;
the GPU box has no reference tree, so everything the ``-m gpu`` tests need is
written here as small fixtures.  Nothing is copied from the reference's sources:
the reference modules are imported (or, for ``normalize_eeg``, the single function
is exec'd out of ``Frontend/app.py`` because that file imports streamlit at the top)
and only their inputs/outputs are stored.

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py
"""
from __future__ import annotations

import ast
import os
import sys
import threading
import types
from pathlib import Path

import numpy as np
import torch

REF = Path(os.environ.get("NA_REFERENCE", "/root/reference"))
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"
sys.dont_write_bytecode = True
sys.path.insert(0, str(REF / "Neuro-Alpha-App"))

PREFIXES = ["food", "water", "backgroundnoise", "yes", "no"]          # file-name prefixes
THREE = {"food": 0, "water": 1, "backgroundnoise": 2}                 # CLASS_NAMES order, lstm_eeg_model.py:11
FIVE = {"yes": 0, "no": 1, "water": 2, "food": 3, "backgroundnoise": 4}   # north_star order


def stub_brainflow():
    """brainflow is not installed; tester.py only needs the names to import."""
    bf = types.ModuleType("brainflow")
    bs = types.ModuleType("brainflow.board_shim")

    class _Ids:
        class NEUROPAWN_KNIGHT_BOARD:
            value = 0

    bs.BoardShim = type("BoardShim", (), {})
    bs.BrainFlowInputParams = type("BrainFlowInputParams", (), {})
    bs.BoardIds = _Ids
    bf.board_shim = bs
    sys.modules.setdefault("brainflow", bf)
    sys.modules.setdefault("brainflow.board_shim", bs)


def load_windows():
    files = sorted(p for p in (REF / "EEG_data_collection").glob("*.csv"))
    X = np.stack([np.loadtxt(p, delimiter=",", dtype=np.float32) for p in files])
    names = [p.name for p in files]
    prefix = [n.split("_")[0] for n in names]
    return X, names, prefix


def csv_bytes_fixture():
    """Raw bytes of a few of the reference's own CSV windows + what np.loadtxt (the reference's reader) makes of
    them: the pin for the GPU CSV parser (na_csv_parse_f32)."""
    files = sorted(p for p in (REF / "EEG_data_collection").glob("*.csv"))
    pick = [files[i] for i in (0, 57, 133, 210, 323)]
    raw = [p.read_bytes() for p in pick]
    offsets = np.zeros(len(raw) + 1, dtype=np.int64)
    np.cumsum([len(r) for r in raw], out=offsets[1:])
    parsed = np.stack([np.loadtxt(p, delimiter=",", dtype=np.float32) for p in pick])
    np.savez_compressed(OUT / "csv_bytes.npz", text=np.frombuffer(b"".join(raw), dtype=np.uint8), offsets=offsets,
                        parsed=parsed, names=np.array([p.name for p in pick]))


def main():
    OUT.mkdir(parents=True, exist_ok=True)
    csv_bytes_fixture()
    torch.set_num_threads(1)           # deterministic reduction order in the fixtures
    stub_brainflow()
    from Utilities.lstm_eeg_model import EEG_LSTM, SimplePredictor, CLASS_NAMES
    from Utilities.preprocessor import PreProcessor
    import Utilities.tester as ref_tester

    pth = REF / "DeepLearning" / "LSTM_Model" / "lstm_classifier_Water_Food_Bg_Noise.pth"
    sd = torch.load(pth, map_location="cpu", weights_only=True)
    keys = list(sd.keys())
    np.savez(OUT / "checkpoint_3class.npz", __order__=np.array(keys), **{k: v.numpy() for k, v in sd.items()})

    X, names, prefix = load_windows()
    assert X.shape == (324, 625, 8), X.shape
    np.savez_compressed(OUT / "eeg_windows.npz", X=X, names=np.array(names), prefix=np.array(prefix))

    model = EEG_LSTM()
    model.load_state_dict(sd, strict=True)
    model.eval()
    xt = torch.from_numpy(X)
    with torch.inference_mode():
        logits_b1 = torch.cat([model(xt[i:i + 1]) for i in range(len(X))]).numpy()   # SimplePredictor's B=1 path
        logits_batched = model(xt).numpy()
    out = {"logits_raw_b1": logits_b1, "logits_raw_batched": logits_batched}

    # MindsAI-filtered path (the live path: preprocessor.py:21-36 -> model).  Logits for all
    # windows, filtered windows stored for a subset only (size).
    pre = PreProcessor(sr=125, tailoring_lambda=1.25e-29)
    sub = np.arange(0, len(X), 10)
    filt_all = np.stack([pre.transform(X[i]) for i in range(len(X))]).astype(np.float32)
    with torch.inference_mode():
        out["logits_filtered_b1"] = torch.cat(
            [model(torch.from_numpy(filt_all[i:i + 1])) for i in range(len(X))]).numpy()
    out["filtered_subset_idx"] = sub
    out["filtered_subset"] = filt_all[sub]

    # SimplePredictor.predict on the subset (lstm_eeg_model.py:86-101)
    pred = SimplePredictor(str(pth), sr=125, device="cpu")
    pp = [pred.predict(X[i]) for i in sub]
    out["predict_probs_subset"] = np.stack([p for p, _ in pp])
    out["predict_labels_subset"] = np.array([l for _, l in pp])
    out["class_names"] = np.array(CLASS_NAMES)
    np.savez_compressed(OUT / "ref_outputs_3class.npz", **out)

    # Gradients of mean CE in eval mode on the first 16 three-class windows (torch autograd on
    # the reference module: the only backward the reference has -- SURVEY 8(a) a15).
    idx3 = [i for i, p in enumerate(prefix) if p in THREE]
    sel = np.array(idx3[:: max(1, len(idx3) // 16)][:16])
    y = torch.tensor([THREE[prefix[i]] for i in sel])
    model.zero_grad()
    logits = model(xt[sel])
    loss = torch.nn.functional.cross_entropy(logits, y)
    loss.backward()
    g = {k: p.grad.numpy().copy() for k, p in model.named_parameters()}
    np.savez(OUT / "ref_grads_3class_eval_b16.npz", sel=sel, y=y.numpy(), loss=loss.item(),
             logits=logits.detach().numpy(), **g)

    # fp64 "truth" for the same 16 windows from the oracle's explicit restatement.  The reference's own
    # fp32 autograd is up to 1.1e-5 away from it on attn.weight (softmax-over-time cancellation), so a
    # 1e-5 parity test needs to know how much of a difference is the reference's own rounding.
    from oracle.torch_ref import explicit_forward
    sd64 = {k: v.double().requires_grad_(True) for k, v in sd.items()}
    l64 = torch.nn.functional.cross_entropy(explicit_forward(xt[sel].double(), sd64), y)
    l64.backward()
    np.savez(OUT / "fp64_grads_3class_eval_b16.npz", loss=l64.item(), **{k: v.grad.numpy() for k, v in sd64.items()})

    # Seeded default init: manual_seed(7); EEG_LSTM()  (same RNG draw order)
    torch.manual_seed(7)
    m7 = EEG_LSTM()
    np.savez(OUT / "ref_init_seed7.npz", **{k: v.numpy() for k, v in m7.state_dict().items()})

    # 5-class variant (config 4): seeded init, logits on 32 windows, grads on 16
    torch.manual_seed(1234)
    m5 = EEG_LSTM(num_classes=5).eval()
    sel5 = np.arange(0, len(X), 10)[:32]
    y5 = torch.tensor([FIVE[prefix[i]] for i in sel5])
    with torch.inference_mode():
        l5 = m5(xt[sel5]).numpy()
    m5.zero_grad()
    lo = m5(xt[sel5[:16]])
    ls = torch.nn.functional.cross_entropy(lo, y5[:16])
    ls.backward()
    np.savez(OUT / "ref_5class.npz", sel=sel5, y=y5.numpy(), logits=l5, loss=ls.item(),
             **{"sd." + k: v.numpy() for k, v in m5.state_dict().items()},
             **{"grad." + k: p.grad.numpy().copy() for k, p in m5.named_parameters()})

    # Stress-shape numerics at reduced batch/time (config 5 arithmetic: H=192): seeded init, T=200
    torch.manual_seed(4321)
    ms = EEG_LSTM(hidden_size=192).eval()
    gs = torch.Generator().manual_seed(5)
    xs = torch.randn(4, 200, 8, generator=gs) * 2.73
    ys = torch.tensor([0, 1, 2, 1])
    ms.zero_grad()
    los = ms(xs)
    lss = torch.nn.functional.cross_entropy(los, ys)
    lss.backward()
    np.savez_compressed(OUT / "ref_stress_h192.npz", x=xs.numpy(), y=ys.numpy(), logits=los.detach().numpy(),
                        loss=lss.item(),
                        **{"sd." + k: v.numpy() for k, v in ms.state_dict().items()},
                        **{"grad." + k: p.grad.numpy().copy() for k, p in ms.named_parameters()})

    # run_trials (tester.py:30-110) end to end with a fake producer pushing CSV windows
    trial_idx = [i for i, p in enumerate(prefix) if p == "water"][:10]

    class FakeProducer:
        def __init__(self, serial_port, num_channels, window_seconds, out_queue):
            self.q = out_queue
            self.recording_flag = types.SimpleNamespace(value=False)
            self._alive = True

        def _run(self):
            for i in trial_idx:
                self.q.put({"sr": 125, "channels": list(range(1, 9)), "data": X[i], "t_emit": 0.0})

        def start(self):
            threading.Thread(target=self._run, daemon=True).start()

        def is_alive(self):
            return self._alive

        def stop(self):
            self._alive = False

        def join(self, timeout=None):
            pass

    ref_tester.StreamingProcess = FakeProducer
    res = ref_tester.run_trials(trials=10, model_path=str(pth), verbose=False)
    per_trial = np.stack([pred.predict(X[i])[0] for i in trial_idx])
    np.savez(OUT / "ref_run_trials.npz", trial_idx=np.array(trial_idx), trials=res.trials,
             avg_probs=res.avg_probs, avg_chunk=res.avg_chunk, per_trial_probs=per_trial,
             filtered=np.stack([pre.transform(X[i]) for i in trial_idx]).astype(np.float32))

    # normalize_eeg (Frontend/app.py:166-170): exec the one function out of the file
    src = (REF / "Neuro-Alpha-App" / "Frontend" / "app.py").read_text()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "normalize_eeg")
    ns = {"np": np}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "app.py", "exec"), ns)
    zi = np.array([0, 57, 123, 200, 323])
    np.savez(OUT / "ref_zscore.npz", idx=zi, z=np.stack([ns["normalize_eeg"](X[i]) for i in zi]),
             z_avg_chunk=ns["normalize_eeg"](res.avg_chunk))

    for p in sorted(OUT.iterdir()):
        print(f"{p.name:40s} {p.stat().st_size:>10d} B")


if __name__ == "__main__":
    main()
