import sys, numpy as np, torch
sys.path.insert(0, '.')
from neural_speech_decoding_b200.lstm_eeg_model import EEG_LSTM
from neural_speech_decoding_b200 import ops
dev = torch.device('cuda:0'); H, T, N = 192, 2500, 18944
torch.manual_seed(0)
m = EEG_LSTM(hidden_size=H).to(dev).eval(); m.compute_dtype = torch.bfloat16
x = torch.randn(N, T, 8, device=dev) * 2.73
def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
outs = {}
with torch.inference_mode():
    for fused in (False, True):
        ops.FUSED_INPUT_WIDE = fused
        ms = t(lambda: m.decode(x))
        outs[fused] = m.decode(x)[0].cpu().numpy()
        print(f"wide H={H} T={T} N={N} fused_input={fused}: {ms:.2f} ms -> {N/ms:.1f} k windows/s")
print("bit-identical:", np.array_equal(outs[False], outs[True]))
