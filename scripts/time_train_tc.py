import sys, numpy as np, torch
sys.path.insert(0, '.')
from neural_speech_decoding_b200.lstm_eeg_model import EEG_LSTM
from neural_speech_decoding_b200 import ops
ck = np.load('tests/golden/checkpoint_3class.npz')
sd = {str(k): torch.from_numpy(ck[str(k)].copy()) for k in ck['__order__']}
dev = torch.device('cuda:0')
m = EEG_LSTM(); m.load_state_dict(sd); m = m.to(dev).train()
m.compute_dtype = torch.bfloat16
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T = 625
x = torch.randn(B, T, 8, device=dev) * 2.73
y = torch.randint(0, 3, (B,), device=dev)
def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
opt = torch.optim.Adam(m.parameters(), lr=1e-3)
def step():
    opt.zero_grad(); torch.nn.functional.cross_entropy(m(x), y).backward(); opt.step()
ms = t(step)
print(f"TC train step B={B}: {ms:.2f} ms -> {B/ms*1e3:.0f} windows/s")
with torch.no_grad():
    xt = ops.window_zscore(x, T, T, False, True, 2, 128)
    Bp = xt.shape[1]
    packed = ops.decoder_pack_bf16(m.lstm.layer(0) + m.lstm.layer(1))
    mask = (torch.rand(T, Bp, 48, device=dev) >= 0.6).to(torch.uint8)
    res = {}
    res['noise'] = t(lambda: m._draw_noise(B, T, dev, 128, torch.uint8))
    res['pack_x'] = t(lambda: ops.window_zscore(x, T, T, False, True, 2, 128))
    res['fwd_train'] = t(lambda: ops.lstm2_fwd_train_bf16(xt, packed, None, 1234, 26214, 2.5))
    h0, h0d, c0, h1, h1f, c1 = ops.lstm2_fwd_train_bf16(xt, packed, None, 1234, 26214, 2.5)
    head = [p.detach() for p in m._head_params()]
    res['head_fwd'] = t(lambda: ops.head_fwd(h1f, B, head, None, None, 2.5, False, True))
    lg, _, st, zp = ops.head_fwd(h1f, B, head, None, None, 2.5, False, True)
    dl = torch.randn_like(lg) * 1e-4
    res['head_bwd'] = t(lambda: ops.head_bwd(dl * 1000, h1f, st, zp, head, None, None, 2.5))
    dh, _ = ops.head_bwd(dl * 1000, h1f, st, zp, head, None, None, 2.5)
    w1 = [p.detach() for p in m.lstm.layer(1)]; w0 = [p.detach() for p in m.lstm.layer(0)]
    res['bwd_l1'] = t(lambda: ops.lstm_bwd_bf16(1, h0d, h1, c1, dh, packed, w1[0], w1[1], None, 1234, 26214, 2.5))
    din1 = ops.lstm_bwd_bf16(1, h0d, h1, c1, dh, packed, w1[0], w1[1], None, 1234, 26214, 2.5)[0]
    res['bwd_l0'] = t(lambda: ops.lstm_bwd_bf16(0, xt, h0, c0, din1, packed, w0[0], w0[1], None, 0, 65536, 1.0))
tot = sum(res.values())
for k, v in res.items(): print(f"  {k:12s} {v:8.3f} ms  {100*v/tot:5.1f}%")
print(f"  sum {tot:.2f} ms")
