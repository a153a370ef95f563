import sys, numpy as np, torch
sys.path.insert(0, '.')
from neural_speech_decoding_b200.lstm_eeg_model import EEG_LSTM
from neural_speech_decoding_b200 import _lib
ck = np.load('tests/golden/checkpoint_3class.npz')
sd = {str(k): torch.from_numpy(ck[str(k)].copy()) for k in ck['__order__']}
dev = torch.device('cuda:0')
B = int(sys.argv[1]) if len(sys.argv) > 1 else 18944
x = torch.randn(B, 625, 8, device=dev) * 2.73
y = torch.randint(0, 3, (B,), device=dev)
grads = {}
for v2 in (0, 1):
    _lib.call("na_set_tuning", b"tc_train_fwd_v2", v2)
    torch.manual_seed(0)
    m = EEG_LSTM(); m.load_state_dict(sd); m = m.to(dev).train(); m.compute_dtype = torch.bfloat16
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    def step():
        opt.zero_grad(); torch.nn.functional.cross_entropy(m(x), y).backward(); opt.step()
    torch.manual_seed(1); opt.zero_grad(); loss = torch.nn.functional.cross_entropy(m(x), y); loss.backward()
    grads[v2] = {k: p.grad.clone() for k, p in m.named_parameters()}
    lossv = float(loss)
    step(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"train fwd v2={v2}: step {ms:.2f} ms -> {B/ms*1e3:.0f} windows/s, loss {lossv:.5f}")
worst = max(float((grads[0][k] - grads[1][k]).abs().max() / grads[0][k].abs().max()) for k in grads[0])
print("worst relative gradient difference v1 vs v2 (same dropout seed):", worst)
