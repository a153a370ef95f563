"""CPU restatement (TEST INFRASTRUCTURE) of the preprocessing filter in front of the decoder:
Neuro-Alpha-App/Utilities/preprocessor.py:21-36 -> Utilities/MindsAI/mindsai_filter_python/core.py:14-48.

Written from the mathematics (numpy FFT, float64), not from the package's source; pinned against the reference's own
outputs (tests/golden/ref_outputs_3class.npz: ``filtered_subset``, produced by oracle/make_golden.py calling the real
``PreProcessor``) in tests/test_oracle.py.
"""
from __future__ import annotations

import numpy as np


def analytic_phase(x_txc: np.ndarray) -> np.ndarray:
    """Instantaneous phase of every channel: angle of x + i H[x] (core.py:14-16; scipy.signal.hilbert semantics)."""
    x = np.asarray(x_txc, np.float64)
    n = x.shape[0]
    spec = np.fft.fft(x, axis=0)
    h = np.zeros(n)
    h[0] = 1.0
    if n % 2 == 0:
        h[n // 2] = 1.0
        h[1:n // 2] = 2.0
    else:
        h[1:(n + 1) // 2] = 2.0
    return np.angle(np.fft.ifft(spec * h[:, None], axis=0))


def coupling_operator(phases_txc: np.ndarray, eps: float = 1e-12) -> np.ndarray:
    """P[i, j] = sum_t sin^2(phi_i - phi_j), zero diagonal, then D^-1 P D^-1 with D = sqrt(clip(diag P, eps))
    (core.py:18-34)."""
    ph = np.asarray(phases_txc, np.float64)
    d = np.sin(ph[:, :, None] - ph[:, None, :])
    P = np.sum(d * d, axis=0)
    np.fill_diagonal(P, 0.0)
    scale = 1.0 / np.sqrt(np.clip(np.diag(P), eps, None))
    return (scale[:, None] * P) * scale[None, :]


def phase_coupling_filter(x_txc: np.ndarray, lambd: float = 1.25e-29) -> np.ndarray:
    """[T, C] float32 window -> filtered [T, C] float32 (preprocessor.py:21-36 with core.py:36-48)."""
    x = np.asarray(x_txc, np.float32).astype(np.float64)
    P = coupling_operator(analytic_phase(x))
    M = np.linalg.inv(np.eye(P.shape[0]) + lambd * (P.T @ P))
    return (M @ x.T).T.astype(np.float32)
