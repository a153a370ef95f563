"""Time the tcgen05 decoder kernel alone at batch sizes that are whole / partial tile rounds (148 SMs x 128 windows)."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from neural_speech_decoding_b200.lstm_eeg_model import EEG_LSTM
from neural_speech_decoding_b200 import ops, _lib
ck = np.load('tests/golden/checkpoint_3class.npz')
sd = {str(k): torch.from_numpy(ck[str(k)].copy()) for k in ck['__order__']}
dev = torch.device('cuda:0')
m = EEG_LSTM(); m.load_state_dict(sd); m = m.to(dev).eval(); m.compute_dtype = torch.bfloat16
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
knobs = [a for a in sys.argv[1:] if '=' in a]
for kv in knobs:
    k, v = kv.split('='); _lib.call("na_set_tuning", k.encode(), int(v))
for N in (4736, 9472, 14208, 18944, 23680, 37888, 40960, 56832):
    x = torch.randn(N, 625, 8, device=dev) * 2.73
    with torch.inference_mode():
        xt = ops.window_zscore(x, 625, 625, False, True, 2, 128)
        packed = m._packed_tc(); head = m._head_params()
        for hs in (2, 3):
            _lib.call("na_set_tuning", b"tc_infer_hs", hs)
            ms = t(lambda: ops.decoder_infer_bf16(xt, packed, head, N, True))
            print(f"N={N} ({N/18944:.2f} rounds) HS={hs}: {ms:.3f} ms -> {N/ms*1e3/1e6:.3f} M windows/s", flush=True)
