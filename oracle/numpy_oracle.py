"""Pure-numpy restatement of the reference decoder arithmetic (TEST INFRASTRUCTURE).

Every function cites the reference lines it restates (paths relative to
``/root/reference``).  The arithmetic of ``nn.LSTM`` / ``nn.LayerNorm`` /
``nn.RReLU`` lives in torch (un-vendored, un-pinned: ``requirements.txt`` has no
torch line); the equations below are torch's documented ones and are pinned
against the real reference's outputs in ``tests/test_oracle.py``.

All functions take a ``dtype`` (np.float32 to mirror the reference, np.float64
to measure rounding noise) and are vectorised over the batch only -- the time
loop is explicit, exactly as the recurrence is defined.
"""
from __future__ import annotations

import numpy as np

# nn.RReLU() defaults lower=1/8, upper=1/3; in eval mode the slope is their mean
# (Neuro-Alpha-App/Utilities/lstm_eeg_model.py:27).
RRELU_LOWER = 1.0 / 8.0
RRELU_UPPER = 1.0 / 3.0
RRELU_EVAL_SLOPE = (RRELU_LOWER + RRELU_UPPER) / 2.0
LN_EPS = 1e-5  # nn.LayerNorm default, lstm_eeg_model.py:23


def _sigmoid(v):
    return 1.0 / (1.0 + np.exp(-v))


def lstm_layer_forward(x, w_ih, w_hh, b_ih, b_hh, dtype=np.float32, return_cell=False):
    """One ``nn.LSTM`` layer, batch_first, h0=c0=0, gate order [i,f,g,o].

    Restates the layer that ``self.lstm(x)`` runs at lstm_eeg_model.py:16-22,34.
    x: [B,T,K] -> h: [B,T,H] (and c: [B,T,H]).
    """
    x = np.asarray(x, dtype)
    w_ih, w_hh = np.asarray(w_ih, dtype), np.asarray(w_hh, dtype)
    bias = np.asarray(b_ih, dtype) + np.asarray(b_hh, dtype)
    B, T, _ = x.shape
    H = w_hh.shape[1]
    h = np.zeros((B, H), dtype)
    c = np.zeros((B, H), dtype)
    hs = np.empty((B, T, H), dtype)
    cs = np.empty((B, T, H), dtype)
    for t in range(T):
        g = x[:, t] @ w_ih.T + h @ w_hh.T + bias
        i = _sigmoid(g[:, 0 * H:1 * H])
        f = _sigmoid(g[:, 1 * H:2 * H])
        gg = np.tanh(g[:, 2 * H:3 * H])
        o = _sigmoid(g[:, 3 * H:4 * H])
        c = (f * c + i * gg).astype(dtype)
        h = (o * np.tanh(c)).astype(dtype)
        hs[:, t] = h
        cs[:, t] = c
    return (hs, cs) if return_cell else hs


def head_forward(out, sd, dtype=np.float32, rrelu_slope=None, drop_mask=None, p=0.6):
    """Attention pool -> LayerNorm -> Linear -> RReLU -> Dropout -> Linear.

    Restates lstm_eeg_model.py:35-39 (and the module definitions :23-30).
    ``rrelu_slope`` [B,32] / ``drop_mask`` [B,32] (0/1) inject the train-mode noise;
    ``None`` gives eval-mode behaviour (slope = mean, dropout = identity).
    out: [B,T,H] -> logits [B,K].
    """
    out = np.asarray(out, dtype)
    g = lambda k: np.asarray(sd[k], dtype)
    scores = out @ g("attn.weight")[0] + g("attn.bias")[0]            # :35
    scores = scores - scores.max(axis=1, keepdims=True)
    w = np.exp(scores)
    w = w / w.sum(axis=1, keepdims=True)                              # :36 softmax over time
    z = (out * w[..., None]).sum(axis=1)                              # :37
    mu = z.mean(axis=-1, keepdims=True)
    var = ((z - mu) ** 2).mean(axis=-1, keepdims=True)                # biased variance
    z = (z - mu) / np.sqrt(var + dtype(LN_EPS)) * g("ln.weight") + g("ln.bias")   # :38
    a = z @ g("fc.0.weight").T + g("fc.0.bias")
    slope = dtype(RRELU_EVAL_SLOPE) if rrelu_slope is None else np.asarray(rrelu_slope, dtype)
    a = np.where(a >= 0, a, a * slope)
    if drop_mask is not None:
        a = a * np.asarray(drop_mask, dtype) / dtype(1.0 - p)
    return (a @ g("fc.3.weight").T + g("fc.3.bias")).astype(dtype)   # :39


def decoder_forward(x, sd, dtype=np.float32, num_layers=2, drop1_mask=None,
                    rrelu_slope=None, drop2_mask=None, p=0.6):
    """``EEG_LSTM.forward`` (lstm_eeg_model.py:32-39). x [B,T,C] -> logits [B,K].

    ``drop1_mask`` [B,T,H] is the inter-layer LSTM dropout mask applied to every
    layer's output except the last (lstm_eeg_model.py:21).
    """
    h = np.asarray(x, dtype)
    for l in range(num_layers):
        h = lstm_layer_forward(h, sd[f"lstm.weight_ih_l{l}"], sd[f"lstm.weight_hh_l{l}"],
                               sd[f"lstm.bias_ih_l{l}"], sd[f"lstm.bias_hh_l{l}"], dtype)
        if drop1_mask is not None and l < num_layers - 1:
            h = h * np.asarray(drop1_mask, dtype) / dtype(1.0 - p)
    return head_forward(h, sd, dtype, rrelu_slope, drop2_mask, p)


def softmax(logits, dtype=np.float32):
    """``F.softmax(logits, dim=-1)`` at lstm_eeg_model.py:97."""
    v = np.asarray(logits, dtype)
    v = v - v.max(axis=-1, keepdims=True)
    e = np.exp(v)
    return (e / e.sum(axis=-1, keepdims=True)).astype(dtype)


def trial_mean(probs_rk):
    """``run_trials`` averaging (tester.py:54,89,97): f32 zeros, ``+=`` in arrival
    order, then divide by the trial count.  probs_rk: [R,...,K] -> [...,K]."""
    p = np.asarray(probs_rk, np.float32)
    acc = np.zeros(p.shape[1:], np.float32)
    for r in range(p.shape[0]):
        acc += p[r]
    return acc / np.float32(p.shape[0]) if p.dtype == np.float32 else acc / p.shape[0]


def zscore_window(chunk):
    """``normalize_eeg`` (Frontend/app.py:166-170): per-window, per-channel
    ``(x-mean)/(std+1e-6)`` over the time axis, population std, f32 in/out.
    chunk: [...,T,C]."""
    x = np.asarray(chunk, np.float32)
    mu = x.mean(axis=-2, keepdims=True)
    sigma = x.std(axis=-2, keepdims=True) + 1e-6
    return ((x - mu) / sigma).astype(np.float32)
