"""collector-side filter chain.  PARITY UNPINNED (BrainFlow absent): the oracle is
oracle/filter_chain.py, a scipy restatement of BrainFlow's published algorithm."""
import numpy as np
import pytest
import torch
from scipy import signal

from oracle import filter_chain as fc


def test_butterworth_design_matches_scipy():
    """CPU: the from-first-principles SOS design has the same transfer function as scipy.signal.butter."""
    from neural_speech_decoding_b200 import filters
    w = np.linspace(0.01, np.pi - 0.01, 2000)
    cases = list(filters.COLLECTOR_CHAIN) + [("bandpass", 8.0, 13.0, 3), ("bandstop", 20.0, 30.0, 5), ("bandpass", 1.0, 60.0, 1)]
    for kind, lo, hi, order in cases:
        mine = filters.butter_band_sos(order, lo, hi, kind, 125.0)
        ref = signal.butter(order, [lo, hi], btype=kind, fs=125.0, output="sos")
        assert mine.shape == ref.shape == (order, 6)
        _, h1 = signal.sosfreqz(mine, worN=w)
        _, h2 = signal.sosfreqz(ref, worN=w)
        assert np.abs(h1 - h2).max() < 1e-11, (kind, lo, hi, order)
        assert np.all(np.abs(np.roots(s[3:])) < 1.0 for s in mine)            # stable
    with pytest.raises(ValueError):
        filters.butter_band_sos(2, 30.0, 70.0, "bandpass", 125.0)              # above Nyquist
    with pytest.raises(ValueError):
        filters.butter_band_sos(2, 3.0, 48.0, "lowpass", 125.0)


def test_oracle_cascade_matches_scipy_sosfilt():
    """CPU: the explicit-loop DirectFormII cascade of the oracle == scipy.signal.sosfilt (same recurrence)."""
    rng = np.random.default_rng(0)
    x = rng.standard_normal(300)
    sos = fc.scipy_sos("bandstop", 49.5, 50.5, 4, 125.0)
    got = fc.cascade(x, sos, np.zeros((4, 2)))
    np.testing.assert_allclose(got, signal.sosfilt(sos, x), rtol=0, atol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("carry", [True, False])
def test_filter_chain_matches_oracle(windows, carry):
    from neural_speech_decoding_b200 import filters
    rng = np.random.default_rng(1)
    raw = np.concatenate([windows["X"][:5] * 7.0 + 3.0,                         # recorded windows, rescaled + offset
                          (rng.standard_normal((3, 625, 8)) * 40 + 100 * np.sin(np.arange(625) * 2 * np.pi * 50 / 125)[None, :, None]
                           ).astype(np.float32)])                               # noise + strong 50 Hz mains
    got = filters.filter_windows(torch.from_numpy(raw).cuda(), carry_state=carry).cpu().numpy()
    sos_list = filters.chain_sos()
    for i in range(raw.shape[0]):
        want = fc.collector_chain(raw[i], sos_list=sos_list, carry_state=carry)           # same coefficients
        np.testing.assert_allclose(got[i], want, rtol=0, atol=2e-6), i
        want_scipy = fc.collector_chain(raw[i], carry_state=carry)                        # scipy's own design
        np.testing.assert_allclose(got[i], want_scipy, rtol=0, atol=5e-6), i
    # the 50 Hz line is gone, the mean is gone
    spec = np.abs(np.fft.rfft(got[-1][:, 0]))
    assert spec[250] < 1e-2 * np.abs(np.fft.rfft(raw[-1][:, 0]))[250]
    assert abs(got[-1].mean()) < 0.5


@pytest.mark.gpu
def test_filter_chain_options_and_errors(windows):
    from neural_speech_decoding_b200 import filters
    x = torch.from_numpy(windows["X"][:2]).cuda()
    y0 = filters.filter_windows(x, chain=(), detrend=True, round_decimals=-1).cpu().numpy()
    np.testing.assert_allclose(y0, windows["X"][:2] - windows["X"][:2].mean(axis=1, keepdims=True), atol=1e-6)
    y1 = filters.filter_windows(x, chain=(("bandpass", 8.0, 13.0, 3),), detrend=False, round_decimals=-1, carry_state=False).cpu().numpy()
    sos = signal.butter(3, [8.0, 13.0], btype="bandpass", fs=125.0, output="sos")
    want = signal.sosfilt(sos, signal.sosfilt(sos, windows["X"][0].astype(np.float64), axis=0)[::-1], axis=0)[::-1]
    np.testing.assert_allclose(y1[0], want, atol=2e-5)
    with pytest.raises(ValueError):
        filters.filter_windows(x[0])
    with pytest.raises(RuntimeError):
        filters.filter_windows(x.cpu())
