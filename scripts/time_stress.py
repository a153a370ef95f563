"""configs[4] stress shape: T=2500, H=192 (4x window length, 4x hidden width), generic tier."""
import sys, torch
sys.path.insert(0, '.')
from neural_speech_decoding_b200.lstm_eeg_model import EEG_LSTM
dev = torch.device('cuda:0')
torch.manual_seed(0)
m = EEG_LSTM(hidden_size=192).to(dev)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
x = torch.randn(B, 2500, 8, device=dev) * 2.73
y = torch.randint(0, 3, (B,), device=dev)
def t(fn, reps=2):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
m.eval()
with torch.inference_mode():
    ms = t(lambda: m(x))
print(f"stress fwd  B={B}: {ms:.1f} ms -> {B/ms*1e3:.0f} windows/s, {2244492480*B/ms/1e9:.2f} TFLOP/s")
m.train()
opt = torch.optim.Adam(m.parameters(), lr=1e-3)
def step():
    opt.zero_grad(); torch.nn.functional.cross_entropy(m(x), y).backward(); opt.step()
ms = t(step)
print(f"stress train B={B}: {ms:.1f} ms -> {B/ms*1e3:.0f} windows/s, {6702757440*B/ms/1e9:.2f} TFLOP/s")
