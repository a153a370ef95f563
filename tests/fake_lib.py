"""Host-memory stand-in for libneuroalpha_b200.so, for CPU-only tests of the HOST logic.

TEST INFRASTRUCTURE.  It implements the C ABI of include/neuroalpha.h in numpy on raw HOST
pointers, following the same formulas the CUDA kernels implement (forward recurrence, fused
BPTT, head backward, reductions).  `tests/conftest.py::cpu_backend` swaps it in for
``_lib.load()`` and registers CPU kernels for the ``neuroalpha::*`` custom ops, so the Python
layers above the ABI (layouts, padding, autograd wiring, DP trainer) can be exercised where there
is no GPU.  The product never imports this module, and the `-m gpu` tests never use it.
"""
from __future__ import annotations

import ctypes

import numpy as np

FC = 32
RRELU_EVAL = np.float32((1.0 / 8.0 + 1.0 / 3.0) / 2.0)


def _arr(ptr, shape, dtype=np.float32):
    if ptr is None:
        return None
    n = int(np.prod(shape))
    if n == 0:
        return np.empty(shape, dtype)
    ct = {np.float32: ctypes.c_float, np.uint16: ctypes.c_uint16}[dtype]
    buf = (ct * n).from_address(int(ptr))
    return np.ctypeslib.as_array(buf).reshape(shape)


def _sig(v):
    return 1.0 / (1.0 + np.exp(-v))


def _bf16_bits(v32):
    u = np.asarray(v32, np.float32).view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint16)
    return r


class FakeLib:
    def __init__(self):
        self.launches = 0
        self.err = b""

    # -- plumbing ---------------------------------------------------------------------------
    def na_version(self):
        return 100

    def na_last_error(self):
        return self.err

    def na_launch_count(self):
        return self.launches

    def _fail(self, msg, code=-1):
        self.err = msg.encode()
        return code

    # -- K1 ---------------------------------------------------------------------------------
    def na_window_zscore(self, x, y, B, T, C, hop, normalize, out_tmp, Bp, out_dtype, stream):
        if not out_tmp:
            Bp = B
        if Bp == 0:
            return 0
        n_samples = (B - 1) * hop + T if B > 0 else 0
        xs = _arr(x, (n_samples, C)) if B > 0 else np.zeros((0, C), np.float32)
        win = np.stack([xs[b * hop:b * hop + T] for b in range(B)]) if B > 0 else np.zeros((0, T, C), np.float32)
        if normalize and B > 0:
            mu = win.mean(axis=1, keepdims=True)
            sg = win.std(axis=1, keepdims=True) + np.float32(1e-6)
            win = ((win - mu) / sg).astype(np.float32)
        if out_tmp:
            out = np.zeros((T, Bp, C), np.float32)
            out[:, :B] = win.transpose(1, 0, 2)
        else:
            out = win
        if out_dtype == 0:
            _arr(y, out.shape)[...] = out
        elif out_dtype == 2:
            _arr(y, out.shape, np.uint16)[...] = out.astype(np.float16).view(np.uint16)
        else:
            _arr(y, out.shape, np.uint16)[...] = _bf16_bits(out)
        self.launches += 1
        return 0

    # -- K3 ---------------------------------------------------------------------------------
    def na_pack_lstm_layer(self, w_ih, w_hh, b_ih, b_hh, wt, bias, K, H, stream):
        G = 4 * H
        wi, wh = _arr(w_ih, (G, K)), _arr(w_hh, (G, H))
        _arr(wt, (K + H, G))[...] = np.concatenate([wi.T, wh.T], axis=0)
        _arr(bias, (G,))[...] = _arr(b_ih, (G,)) + _arr(b_hh, (G,))
        self.launches += 1
        return 0

    def na_lstm_layer_fwd_f32(self, inp, wt, bias, hout, cout, gates, drop_mask, drop_scale, hout_drop,
                              T, Bp, K, H, stream):
        if Bp % 32 or Bp <= 0:
            return self._fail("na_lstm_layer_fwd_f32: bad shape")
        if (drop_mask is None) != (hout_drop is None):
            return self._fail("na_lstm_layer_fwd_f32: drop_mask and hout_drop must be given together")
        G = 4 * H
        x = _arr(inp, (T, Bp, K)).astype(np.float64)
        W = _arr(wt, (K + H, G)).astype(np.float64)
        b = _arr(bias, (G,)).astype(np.float64)
        ho, co, go = _arr(hout, (T, Bp, H)), _arr(cout, (T, Bp, H)), _arr(gates, (T, Bp, G))
        dm, hd = _arr(drop_mask, (T, Bp, H)), _arr(hout_drop, (T, Bp, H))
        h = np.zeros((Bp, H))
        c = np.zeros((Bp, H))
        for t in range(T):
            g = np.concatenate([x[t], h], axis=1) @ W + b
            i, f, gg, o = _sig(g[:, :H]), _sig(g[:, H:2 * H]), np.tanh(g[:, 2 * H:3 * H]), _sig(g[:, 3 * H:])
            c = f * c + i * gg
            h = o * np.tanh(c)
            ho[t] = h
            if co is not None:
                co[t] = c
            if go is not None:
                go[t] = np.concatenate([i, f, gg, o], axis=1)
            if hd is not None:
                hd[t] = h * dm[t] * drop_scale
        self.launches += 1
        return 0

    def na_lstm_layer_bwd_f32(self, dh_out, gates, cstate, w_ih, w_hh, dgates, din, in_drop_mask, drop_scale,
                              T, Bp, K, H, stream):
        G = 4 * H
        dho = _arr(dh_out, (T, Bp, H)).astype(np.float64)
        ga = _arr(gates, (T, Bp, G)).astype(np.float64)
        cs = _arr(cstate, (T, Bp, H)).astype(np.float64)
        wi, wh = _arr(w_ih, (G, K)).astype(np.float64), _arr(w_hh, (G, H)).astype(np.float64)
        dg_out, di_out = _arr(dgates, (T, Bp, G)), _arr(din, (T, Bp, K))
        mask = _arr(in_drop_mask, (T, Bp, K))
        dhrec = np.zeros((Bp, H))
        dc = np.zeros((Bp, H))
        for t in range(T - 1, -1, -1):
            i, f, g, o = ga[t, :, :H], ga[t, :, H:2 * H], ga[t, :, 2 * H:3 * H], ga[t, :, 3 * H:]
            cp = cs[t - 1] if t > 0 else np.zeros((Bp, H))
            tc = np.tanh(cs[t])
            dh = dho[t] + dhrec
            d_o = dh * tc
            dct = dc + dh * o * (1 - tc * tc)
            d_i, d_g, d_f = dct * g, dct * i, dct * cp
            dc = dct * f
            pre = np.concatenate([d_i * i * (1 - i), d_f * f * (1 - f), d_g * (1 - g * g), d_o * o * (1 - o)], axis=1)
            dg_out[t] = pre
            dhrec = pre @ wh
            if di_out is not None:
                v = pre @ wi
                if mask is not None:
                    v = v * mask[t] * drop_scale
                di_out[t] = v
        self.launches += 1
        return 0

    def na_wgrad_partial_floats(self, K, H):
        return 64

    def na_lstm_layer_wgrad_f32(self, dgates, inp, h, dw_ih, dw_hh, db, partials, T, Bp, K, H, stream):
        G = 4 * H
        dg = _arr(dgates, (T * Bp, G)).astype(np.float64)
        x = _arr(inp, (T * Bp, K)).astype(np.float64)
        hh = _arr(h, (T * Bp, H)).astype(np.float64)
        _arr(dw_ih, (G, K))[...] = dg.T @ x
        _arr(dw_hh, (G, H))[...] = dg[Bp:].T @ hh[:-Bp] if T > 1 else 0.0
        _arr(db, (G,))[...] = dg.sum(axis=0)
        self.launches += 5
        return 0

    # -- K4 ---------------------------------------------------------------------------------
    @staticmethod
    def _head_params(ptrs, H, NC):
        aw, ab, lw, lb, w0, b0, w3, b3 = ptrs
        return dict(aw=_arr(aw, (H,)).astype(np.float64), ab=float(_arr(ab, (1,))[0]),
                    lw=_arr(lw, (H,)).astype(np.float64), lb=_arr(lb, (H,)).astype(np.float64),
                    w0=_arr(w0, (FC, H)).astype(np.float64), b0=_arr(b0, (FC,)).astype(np.float64),
                    w3=_arr(w3, (NC, FC)).astype(np.float64), b3=_arr(b3, (NC,)).astype(np.float64))

    @staticmethod
    def _tail(zp, P, slope, keep):
        mean = zp.mean(axis=1, keepdims=True)
        var = ((zp - mean) ** 2).mean(axis=1, keepdims=True)
        rstd = 1.0 / np.sqrt(var + 1e-5)
        xhat = (zp - mean) * rstd
        zn = xhat * P["lw"] + P["lb"]
        a_pre = zn @ P["w0"].T + P["b0"]
        act_grad = np.where(a_pre >= 0, 1.0, slope) * keep
        a_post = np.where(a_pre >= 0, a_pre, a_pre * slope) * keep
        return xhat, rstd, zn, a_pre, a_post, act_grad

    def na_head_fwd_f32(self, h, aw, ab, lw, lb, w0, b0, w3, b3, rrelu_slope, drop_mask, drop_scale,
                        logits, probs, stats, zpool, T, B, Bp, H, NC, stream):
        if NC > 16:
            return self._fail("na_head_fwd_f32: num_classes outside [1,16]", -3)
        P = self._head_params((aw, ab, lw, lb, w0, b0, w3, b3), H, NC)
        hh = _arr(h, (T, Bp, H)).astype(np.float64)[:, :B]
        s = hh @ P["aw"] + P["ab"]                       # [T,B]
        m = s.max(axis=0)
        e = np.exp(s - m)
        l = e.sum(axis=0)
        zp = np.einsum("tb,tbh->bh", e, hh) / l[:, None]
        slope = _arr(rrelu_slope, (B, FC)).astype(np.float64) if rrelu_slope is not None else float(RRELU_EVAL)
        keep = _arr(drop_mask, (B, FC)).astype(np.float64) * drop_scale if drop_mask is not None else 1.0
        _, _, _, _, a_post, _ = self._tail(zp, P, slope, keep)
        lg = a_post @ P["w3"].T + P["b3"]
        _arr(logits, (B, NC))[...] = lg
        if probs is not None:
            ee = np.exp(lg - lg.max(axis=1, keepdims=True))
            _arr(probs, (B, NC))[...] = ee / ee.sum(axis=1, keepdims=True)
        if stats is not None:
            st = _arr(stats, (B, 2))
            st[:, 0], st[:, 1] = m, l
        if zpool is not None:
            _arr(zpool, (B, H))[...] = zp
        self.launches += 1
        return 0

    def na_head_param_floats(self, H, NC):
        return H + 1 + 2 * H + FC * H + FC + FC * NC + NC

    def na_head_partial_floats(self, B, H, NC):
        return 64

    def na_head_bwd_f32(self, dlogits, h, stats, zpool, aw, ab, lw, lb, w0, b0, w3, b3, rrelu_slope, drop_mask,
                        drop_scale, dh, dparams, partials, T, B, Bp, H, NC, stream):
        P = self._head_params((aw, ab, lw, lb, w0, b0, w3, b3), H, NC)
        hh = _arr(h, (T, Bp, H)).astype(np.float64)[:, :B]
        st = _arr(stats, (B, 2)).astype(np.float64)
        zp = _arr(zpool, (B, H)).astype(np.float64)
        dl = _arr(dlogits, (B, NC)).astype(np.float64)
        slope = _arr(rrelu_slope, (B, FC)).astype(np.float64) if rrelu_slope is not None else float(RRELU_EVAL)
        keep = _arr(drop_mask, (B, FC)).astype(np.float64) * drop_scale if drop_mask is not None else 1.0
        xhat, rstd, zn, a_pre, a_post, act_grad = self._tail(zp, P, slope, keep)
        da_pre = (dl @ P["w3"]) * act_grad
        dzn = da_pre @ P["w0"]
        dxh = dzn * P["lw"]
        m1 = dxh.mean(axis=1, keepdims=True)
        m2 = (dxh * xhat).mean(axis=1, keepdims=True)
        dz = rstd * (dxh - m1 - xhat * m2)
        s = hh @ P["aw"] + P["ab"]
        alpha = np.exp(s - st[:, 0]) / st[:, 1]                      # [T,B]
        hc = hh - zp[None]                                           # centred form, as in na_head.cu
        ds = alpha * np.einsum("tbh,bh->tb", hc, dz)
        dho = _arr(dh, (T, Bp, H))
        dho[...] = 0.0
        dho[:, :B] = alpha[..., None] * dz[None] + ds[..., None] * P["aw"]
        out = _arr(dparams, (self.na_head_param_floats(H, NC),))
        parts = [np.einsum("tb,tbh->h", ds, hc), [ds.sum()], (dzn * xhat).sum(0), dzn.sum(0),
                 (da_pre.T @ zn).ravel(), da_pre.sum(0), (dl.T @ a_post).ravel(), dl.sum(0)]
        out[...] = np.concatenate([np.asarray(p, np.float64).ravel() for p in parts])
        self.launches += 14
        return 0

    # -- K5 ---------------------------------------------------------------------------------
    def na_trial_mean_f32(self, inp, out, R, N, stream):
        if N == 0:
            return 0
        x = _arr(inp, (R, N))
        acc = np.zeros((N,), np.float32)
        for r in range(R):
            acc += x[r]
        _arr(out, (N,))[...] = acc / np.float32(R)
        self.launches += 1
        return 0


def install():
    """Process-wide installation (for spawned worker processes of the gloo DP test)."""
    import torch
    from neural_speech_decoding_b200 import _lib, ops
    fake = FakeLib()
    _lib.load = lambda path=None: fake
    ops._require_cuda = lambda *t: None
    ops._stream = lambda: None
    ops.compute_device = lambda d: torch.device("cpu")
    ops.EXACT_TC = False             # the fake implements the granular fp32 entry points (host-logic path)
    ops.EXACT_TC_TRAIN = False
    for op in ops.all_custom_ops():
        op.register_kernel("cpu")(op._init_fn)
    return fake
