import sys, numpy as np, torch
sys.path.insert(0, '.')
from neural_speech_decoding_b200.lstm_eeg_model import EEG_LSTM
from neural_speech_decoding_b200 import ops
ck = np.load('tests/golden/checkpoint_3class.npz')
sd = {str(k): torch.from_numpy(ck[str(k)].copy()) for k in ck['__order__']}
dev = torch.device('cuda:0')
m = EEG_LSTM(); m.load_state_dict(sd); m = m.to(dev).eval()
X = torch.from_numpy(np.load('tests/golden/eeg_windows.npz')['X']).to(dev)
ref = np.load('tests/golden/ref_outputs_3class.npz')['logits_raw_b1']
def rel(a, b): return np.abs(a - b).max() / np.abs(b).max()
with torch.inference_mode():
    for flag in (False, True):
        ops.EXACT_TC = flag
        got = m(X).cpu().numpy()
        print(f"EXACT_TC={flag}: rel err vs reference logits on the 324 windows = {rel(got, ref):.2e}, argmax equal: {np.array_equal(got.argmax(1), ref.argmax(1))}", flush=True)
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for N in (4096, 18944, 40960):
    x = torch.randn(N, 625, 8, device=dev) * 2.73
    outs = {}
    with torch.inference_mode():
        for flag in (False, True):
            ops.EXACT_TC = flag
            ms = t(lambda: m.decode(x), reps=3)
            outs[flag] = m.decode(x)[0].cpu().numpy()
            print(f"N={N} EXACT_TC={flag}: {ms:.3f} ms -> {N/ms*1e3/1e6:.3f} M windows/s", flush=True)
    print(f"   tensor-core exact vs FFMA exact: rel diff {rel(outs[True], outs[False]):.2e}")
