"""The trainer that replaces the reference's missing lstm_trainer.ipynb: labels from file-name prefixes,
loss goes down, and the saved state_dict loads into the reference architecture with strict=True."""
import numpy as np
import pytest
import torch

from oracle.torch_ref import RefEEGLSTM


def _data(windows, T=None, n=48):
    from neural_speech_decoding_b200.trainer import DEFAULT_CLASSES
    idx = {c: i for i, c in enumerate(DEFAULT_CLASSES)}
    keep = [i for i, p in enumerate(windows["prefix"]) if p in idx][::3][:n]
    X = windows["X"][keep]
    if T:
        X = np.ascontiguousarray(X[:, :T])
    y = np.array([idx[windows["prefix"][i]] for i in keep], dtype=np.int64)
    return X, y


def test_load_windows_from_npz_and_csv(golden_dir, tmp_path, windows):
    from neural_speech_decoding_b200.trainer import load_windows
    X, y = load_windows(str(golden_dir / "eeg_windows.npz"), ["food", "water", "backgroundnoise"])
    assert X.shape == (179, 625, 8) and np.bincount(y).tolist() == [69, 70, 40]
    X5, y5 = load_windows(str(golden_dir / "eeg_windows.npz"), ["yes", "no", "water", "food", "backgroundnoise"])
    assert X5.shape[0] == 324 and np.bincount(y5).tolist() == [74, 71, 70, 69, 40]
    for i in (0, 100):                                   # the collector's CSV format: %.7f, comma, no header
        np.savetxt(tmp_path / windows["names"][i], windows["X"][i], fmt="%.7f", delimiter=",")
    Xc, yc = load_windows(str(tmp_path), ["food", "water", "backgroundnoise", "yes", "no"])
    assert Xc.shape == (2, 625, 8) and np.abs(Xc[0] - windows["X"][0]).max() < 1e-6


def test_training_loop_host_logic(cpu_backend, windows):
    from neural_speech_decoding_b200.trainer import train
    X, y = _data(windows, T=12, n=24)
    model, hist = train(X, y, 3, epochs=3, batch=8, lr=5e-3, val_frac=0.25, dropout=0.0, device=torch.device("cpu"), log=None)
    assert len(hist) == 3 and hist[-1]["train_loss"] < hist[0]["train_loss"]
    assert np.array(hist[-1]["val_confusion"]).sum() == 6
    ref = RefEEGLSTM(dropout=0.0)
    ref.load_state_dict({k: v.detach().cpu() for k, v in model.state_dict().items()}, strict=True)


@pytest.mark.gpu
@pytest.mark.parametrize("bf16", [False, True])
def test_training_on_repo_windows_gpu(windows, tmp_path, bf16):
    from neural_speech_decoding_b200.trainer import train
    X, y = _data(windows, n=48)
    model, hist = train(X, y, 3, epochs=6, batch=16, lr=3e-3, val_frac=0.25, seed=1, bf16=bf16, log=None)
    assert hist[-1]["train_loss"] < hist[0]["train_loss"]
    out = tmp_path / "m.pth"
    torch.save({k: v.detach().cpu() for k, v in model.state_dict().items()}, out)
    ref = RefEEGLSTM().eval()
    ref.load_state_dict(torch.load(out, map_location="cpu"), strict=True)
    with torch.inference_mode():                         # the trained weights decode identically on the CPU port
        want = ref(torch.from_numpy(X[:8])).numpy()
        got = model.eval()(torch.from_numpy(X[:8]).cuda()).float().cpu().numpy()
    assert np.abs(got - want).max() / np.abs(want).max() < (2e-2 if bf16 else 1e-5)
