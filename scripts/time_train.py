import sys, numpy as np, torch
sys.path.insert(0, '.')
from neural_speech_decoding_b200.lstm_eeg_model import EEG_LSTM
from neural_speech_decoding_b200 import ops
ck = np.load('tests/golden/checkpoint_3class.npz')
sd = {str(k): torch.from_numpy(ck[str(k)].copy()) for k in ck['__order__']}
dev = torch.device('cuda:0')
m = EEG_LSTM(); m.load_state_dict(sd); m = m.to(dev).train()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T = 625
x = torch.randn(B, T, 8, device=dev) * 2.73
def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
with torch.no_grad():
    xt = ops.window_zscore(x, T, T, False, True, False)
    Bp = xt.shape[1]
    p0 = ops.pack_lstm_layer(*m.lstm.layer(0)); p1 = ops.pack_lstm_layer(*m.lstm.layer(1))
    mask = (torch.rand(T, Bp, 48, device=dev) >= 0.6).float()
    res = {}
    res['pack_x'] = t(lambda: ops.window_zscore(x, T, T, False, True, False))
    res['noise'] = t(lambda: m._draw_noise(B, T, dev))
    res['fwd_l0_save+drop'] = t(lambda: ops.lstm_layer_fwd(xt, p0[0], p0[1], mask, 2.5, True))
    h0, c0, g0, hd0 = ops.lstm_layer_fwd(xt, p0[0], p0[1], mask, 2.5, True)
    res['fwd_l1_save'] = t(lambda: ops.lstm_layer_fwd(hd0, p1[0], p1[1], None, 1.0, True))
    h1, c1, g1, _ = ops.lstm_layer_fwd(hd0, p1[0], p1[1], None, 1.0, True)
    head = [p.detach() for p in m._head_params()]
    res['head_fwd'] = t(lambda: ops.head_fwd(h1, B, head, None, None, 2.5, False, True))
    lg, _, st, zp = ops.head_fwd(h1, B, head, None, None, 2.5, False, True)
    dl = torch.randn_like(lg)
    res['head_bwd'] = t(lambda: ops.head_bwd(dl, h1, st, zp, head, None, None, 2.5))
    dh, dpar = ops.head_bwd(dl, h1, st, zp, head, None, None, 2.5)
    w = [p.detach() for p in m.lstm.layer(1)]
    res['bwd_l1'] = t(lambda: ops.lstm_layer_bwd(dh, g1, c1, w[0], w[1], mask, 2.5, True))
    dg1, din1 = ops.lstm_layer_bwd(dh, g1, c1, w[0], w[1], mask, 2.5, True)
    res['wgrad_l1'] = t(lambda: ops.lstm_layer_wgrad(dg1, hd0, h1))
    w0 = [p.detach() for p in m.lstm.layer(0)]
    res['bwd_l0'] = t(lambda: ops.lstm_layer_bwd(din1, g0, c0, w0[0], w0[1], None, 1.0, False))
    dg0, _ = ops.lstm_layer_bwd(din1, g0, c0, w0[0], w0[1], None, 1.0, False)
    res['wgrad_l0'] = t(lambda: ops.lstm_layer_wgrad(dg0, xt, h0))
tot = sum(res.values())
for k, v in res.items(): print(f"{k:20s} {v:8.3f} ms  {100*v/tot:5.1f}%")
print(f"sum {tot:.2f} ms for B={B} -> {B/tot*1e3:.0f} windows/s")
y = torch.randint(0, 3, (B,), device=dev)
opt = torch.optim.Adam(m.parameters(), lr=1e-3)
def step():
    opt.zero_grad(); torch.nn.functional.cross_entropy(m(x), y).backward(); opt.step()
ms = t(step)
print(f"full train step {ms:.2f} ms -> {B/ms*1e3:.0f} windows/s")
