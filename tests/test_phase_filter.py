"""the preprocessing filter in front of the decoder, on the GPU (opt-in).  Pins: the reference's own
filtered windows and the logits of its filtered path (tests/golden/ref_outputs_3class.npz)."""
import numpy as np
import pytest
import torch


def test_gpu_filter_is_opt_in():
    from neural_speech_decoding_b200.preprocess_gpu import PhaseCouplingFilterGPU
    with pytest.raises(PermissionError):
        PhaseCouplingFilterGPU()


@pytest.mark.gpu
def test_gpu_filter_matches_reference_windows_and_logits(windows, checkpoint, golden_dir):
    from neural_speech_decoding_b200.lstm_eeg_model import EEG_LSTM
    from neural_speech_decoding_b200.preprocess_gpu import PhaseCouplingFilterGPU
    f = np.load(golden_dir / "ref_outputs_3class.npz")
    dev = torch.device("cuda:0")
    pre = PhaseCouplingFilterGPU(sr=125, tailoring_lambda=1.25e-29, device=dev, accept_noncommercial_terms=True)
    X = torch.from_numpy(windows["X"]).to(dev)
    Y = pre.transform_batch(X)
    pre.check()
    got = Y.cpu().numpy()
    want = f["filtered_subset"]
    sub = got[f["filtered_subset_idx"]]
    err = np.abs(sub - want).max(axis=(1, 2)) / np.abs(want).max(axis=(1, 2))
    assert err.max() < 1e-5, err.max()
    # single-window entry point (the reference's PreProcessor contract) == the batched one, bit for bit
    one = pre.transform(windows["X"][7])
    assert one.dtype == np.float32 and one.shape == (625, 8) and np.array_equal(one, got[7])
    with pytest.raises(ValueError):
        pre.transform(windows["X"][:2])
    # the whole live path on the GPU: filter -> decoder, against the reference's filtered-path logits of all 324 windows
    m = EEG_LSTM()
    m.load_state_dict(checkpoint, strict=True)
    m = m.to(dev).eval()
    with torch.inference_mode():
        logits = m(Y).cpu().numpy()
    ref = f["logits_filtered_b1"]
    assert np.abs(logits - ref).max() / np.abs(ref).max() < 5e-5
    margin = np.sort(ref, axis=1)
    safe = (margin[:, -1] - margin[:, -2]) > 1e-3
    assert safe.sum() >= 320 and np.array_equal(logits.argmax(1)[safe], ref.argmax(1)[safe])


@pytest.mark.gpu
def test_gpu_filter_random_windows_vs_oracle_and_predictor(checkpoint, tmp_path):
    from oracle.phase_filter import phase_coupling_filter
    from neural_speech_decoding_b200.lstm_eeg_model import SimplePredictor
    from neural_speech_decoding_b200.preprocess_gpu import PhaseCouplingFilterGPU
    rng = np.random.default_rng(3)
    t = np.arange(625)[:, None] / 125.0
    x = (rng.standard_normal((20, 625, 8)) * 2.0 + 3.0 * np.sin(2 * np.pi * (8 + np.arange(8))[None, None, :] * t[None])).astype(np.float32)
    x[3] = 0.0                                     # an all-zero window: phases are 0, P = 0, the filter is the identity
    dev = torch.device("cuda:0")
    for lam in (1.25e-29, 1e-25, 0.0):
        pre = PhaseCouplingFilterGPU(tailoring_lambda=lam, device=dev, accept_noncommercial_terms=True)
        got = pre.transform_batch(torch.from_numpy(x).to(dev)).cpu().numpy()
        pre.check()
        for i in range(x.shape[0]):
            want = phase_coupling_filter(x[i], lam)
            assert np.abs(got[i] - want).max() <= 1e-5 * max(np.abs(want).max(), 1e-6), (lam, i)
    assert np.array_equal(got[3], x[3])
    # SimplePredictor with the GPU front stage: predict_batch filters on the device (no per-window CPU loop)
    torch.save(checkpoint, tmp_path / "m.pth")
    pre = PhaseCouplingFilterGPU(device=dev, accept_noncommercial_terms=True)
    sp = SimplePredictor(str(tmp_path / "m.pth"), sr=125, preprocessor=pre)
    pb = sp.predict_batch(x)
    p0, label = sp.predict(x[0])
    assert pb.shape == (20, 3) and np.abs(pb[0] - p0).max() < 1e-6 and label in sp.class_names
