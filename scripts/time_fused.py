import sys, numpy as np, torch
sys.path.insert(0, '.')
from neural_speech_decoding_b200.lstm_eeg_model import EEG_LSTM
from neural_speech_decoding_b200 import ops
ck = np.load('tests/golden/checkpoint_3class.npz')
sd = {str(k): torch.from_numpy(ck[str(k)].copy()) for k in ck['__order__']}
dev = torch.device('cuda:0')
m = EEG_LSTM(); m.load_state_dict(sd); m = m.to(dev).eval(); m.compute_dtype = torch.bfloat16
def t(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for N in (1024, 4096, 18944, 40960):
    x = torch.randn(N, 625, 8, device=dev) * 2.73
    with torch.inference_mode():
        outs = {}
        for fused in (False, True):
            ops.FUSED_INPUT = fused
            ms = t(lambda: m.decode(x))
            outs[fused] = m.decode(x)[0].cpu().numpy()
            print(f"N={N} fused_input={fused}: {ms:.3f} ms -> {N/ms*1e3/1e6:.3f} M windows/s", flush=True)
        print("   bit-identical:", np.array_equal(outs[False], outs[True]))
