"""CPU tests of the host logic above the C ABI (layouts, padding, autograd wiring, state_dict).

They run the real Python product code against tests/fake_lib.FakeLib (a numpy statement of the ABI
on host pointers) -- see the ``cpu_backend`` fixture.  What they prove: the host side composes the
ABI calls correctly and the hand-derived backward formulas (the ones the CUDA kernels implement)
are the true gradients.  They do NOT prove kernel parity; the ``-m gpu`` tests do.
"""
import numpy as np
import pytest
import torch

from oracle.torch_ref import RefEEGLSTM, explicit_forward
from oracle import numpy_oracle as no


def _model(sd=None, **kw):
    from neural_speech_decoding_b200.lstm_eeg_model import EEG_LSTM
    m = EEG_LSTM(**kw)
    if sd is not None:
        m.load_state_dict(sd, strict=True)
    return m


def test_state_dict_layout_and_strict_load(checkpoint):
    m = _model(checkpoint)
    sd = m.state_dict()
    assert list(sd.keys()) == list(checkpoint.keys())
    for k in sd:
        assert sd[k].shape == checkpoint[k].shape and sd[k].dtype == torch.float32
        assert torch.equal(sd[k], checkpoint[k])
    assert sum(p.numel() for p in m.parameters()) == 31764


def test_seeded_init_matches_reference(golden_dir):
    ref = np.load(golden_dir / "ref_init_seed7.npz")
    torch.manual_seed(7)
    m = _model()
    for k, v in m.state_dict().items():
        assert np.array_equal(v.numpy(), ref[k]), k


def test_cpu_tensor_raises():
    m = _model()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(2, 5, 8))
    with pytest.raises(ValueError):
        m(torch.zeros(5, 8))


def test_eval_forward_matches_reference_logits(cpu_backend, checkpoint, windows, golden_dir):
    ref = np.load(golden_dir / "ref_outputs_3class.npz")["logits_raw_b1"]
    m = _model(checkpoint).eval()
    x = torch.from_numpy(windows["X"][:5, :60].copy())          # B=5 (padding to 32), short T
    with torch.no_grad():
        got = m(x).numpy()
    want = no.decoder_forward(x.numpy(), {k: v.numpy() for k, v in checkpoint.items()}, np.float64)
    assert got.shape == (5, 3)
    np.testing.assert_allclose(got, want, rtol=0, atol=2e-6 * np.abs(want).max())
    x = torch.from_numpy(windows["X"][:3].copy())
    with torch.no_grad():
        got = m(x).numpy()
    np.testing.assert_allclose(got, ref[:3], rtol=0, atol=1e-5 * np.abs(ref[:3]).max())


@pytest.mark.parametrize("train", [False, True])
def test_autograd_matches_oracle(cpu_backend, checkpoint, train):
    torch.manual_seed(0)
    B, T, H = 5, 20, 48
    m = _model(checkpoint)
    m.train(train)
    x = (torch.randn(B, T, 8) * 2.73).requires_grad_(True)
    y = torch.tensor([0, 1, 2, 1, 0])
    noise = {}
    if train:
        noise = dict(drop1=(torch.rand(1, B, T, H) >= 0.6).float(),
                     rrelu_slope=torch.empty(B, 32).uniform_(1 / 8, 1 / 3),
                     drop2=(torch.rand(B, 32) >= 0.6).float())
        m.inject_noise(**noise)
    loss = torch.nn.functional.cross_entropy(m(x), y)
    loss.backward()

    sd = {k: v.detach().double().requires_grad_(True) for k, v in checkpoint.items()}
    xr = x.detach().double().requires_grad_(True)
    lo = explicit_forward(xr, sd, 2, 0.6,
                          noise.get("drop1")[0].double() if train else None,
                          noise.get("rrelu_slope").double() if train else None,
                          noise.get("drop2").double() if train else None)
    lr = torch.nn.functional.cross_entropy(lo, y)
    lr.backward()
    assert abs(loss.item() - lr.item()) < 1e-5
    for (k, p) in m.named_parameters():
        g, r = p.grad.double(), sd[k].grad
        # attn.bias: analytically zero (softmax shift invariance) -> judge it on attn.weight's scale
        denom = sd["attn.weight"].grad.abs().max().item() if k == "attn.bias" else r.abs().max().item()
        assert (g - r).abs().max().item() / max(denom, 1e-12) < 2e-5, k
    assert (x.grad.double() - xr.grad).abs().max().item() / xr.grad.abs().max().item() < 2e-5


def test_five_class_and_other_sizes(cpu_backend):
    torch.manual_seed(3)
    m = _model(input_size=4, hidden_size=20, num_layers=3, num_classes=5, dropout=0.25).eval()
    ref = RefEEGLSTM(4, 20, 3, 5, 0.25).eval()
    ref.load_state_dict(m.state_dict(), strict=True)
    x = torch.randn(7, 11, 4)
    with torch.no_grad():
        np.testing.assert_allclose(m(x).numpy(), ref(x).numpy(), atol=2e-6)


def test_window_zscore_stream_and_batch(cpu_backend, windows, golden_dir):
    from neural_speech_decoding_b200 import ops
    z = np.load(golden_dir / "ref_zscore.npz")
    x = torch.from_numpy(windows["X"][z["idx"]].copy())
    got = ops.window_zscore(x, 625, 625, True, False, False).numpy()
    np.testing.assert_allclose(got, z["z"], atol=1e-6)
    stream = torch.arange(100 * 8, dtype=torch.float32).reshape(100, 8)
    w = ops.window_zscore(stream, 30, 10, False, False, False)
    assert w.shape == (8, 30, 8) and torch.equal(w[3], stream[30:60])
    tm = ops.window_zscore(stream, 30, 10, False, True, False)
    assert tm.shape == (30, 32, 8) and torch.equal(tm[:, 3], stream[30:60]) and tm[:, 8:].abs().sum() == 0


def test_trial_mean_rounding(cpu_backend, golden_dir):
    from neural_speech_decoding_b200 import ops
    t = np.load(golden_dir / "ref_run_trials.npz")
    got = ops.trial_mean(torch.from_numpy(t["per_trial_probs"])).numpy()
    assert np.array_equal(got, t["avg_probs"])


def test_bind_to_gpu_numa_is_harmless_without_a_gpu():
    """dp.bind_to_gpu_numa must never raise: no NVML / no GPU / restricted cpuset -> None and the affinity is untouched."""
    import os
    from neural_speech_decoding_b200 import dp
    before = os.sched_getaffinity(0)
    got = dp.bind_to_gpu_numa(0)
    assert got is None or (isinstance(got, list) and set(got) <= before)
    assert os.sched_getaffinity(0) == (set(got) if got else before)


def test_shard_batch_covers_everything_once():
    from neural_speech_decoding_b200.dp import shard_batch
    for n, w in ((4096, 8), (10, 3), (1, 4), (0, 2)):
        idx = [i for r in range(w) for i in range(n)[shard_batch(n, r, w)]]
        assert idx == list(range(n))


def test_wide_tile_layout_transposes_are_inverse_and_contiguous():
    """ops._to_wtl / _from_wtl (host side of the wide training tier): the per-thread layout the recurrence kernels read is a
    contiguous tensor (a reshape of a permuted view may silently stay a view), the element order is the accumulator column
    order, and the two transposes are inverses."""
    from neural_speech_decoding_b200 import ops
    T, Bp = 3, 256
    for H in (96, 144, 192):
        nch = H // 48
        g = torch.arange(T * Bp * 4 * H, dtype=torch.float32).reshape(T * Bp, 4 * H).to(torch.float16)
        w = ops._to_wtl(g, T, Bp, H)
        assert w.is_contiguous() and tuple(w.shape) == (T, Bp // 128, nch * 3, 8, 128, 8)
        for (t, tile, k, grp, p, row, e) in [(0, 0, 0, 0, 0, 0, 0), (1, 1, nch - 1, 2, 5, 77, 3), (2, 0, 1, 1, 7, 127, 7)]:
            e64 = p * 8 + e
            q, gate, ur = e64 // 16, (e64 % 16) // 4, e64 % 4
            unit = 48 * k + 16 * grp + 4 * q + ur
            assert w[t, tile, k * 3 + grp, p, row, e] == g[t * Bp + tile * 128 + row, gate * H + unit]
        back = ops._from_wtl(w, T, Bp, H)
        assert back.is_contiguous() and torch.equal(back, g)
        for dt in (torch.float32, torch.float16):
            c = torch.randn(T * Bp, H).to(dt)
            wc = ops._to_wtl(c, T, Bp, H)
            epp = 16 // c.element_size()
            assert wc.is_contiguous() and tuple(wc.shape) == (T, Bp // 128, nch * 3, 16 // epp, 128, epp)
            assert wc[1, 1, 2, 1, 9, 2] == c[Bp + 128 + 9, 48 * 0 + 16 * 2 + epp * 1 + 2]
            assert torch.equal(ops._from_wtl(wc, T, Bp, H), c)
