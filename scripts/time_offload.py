"""A/B timing of the MUFU -> FMA-pipe offloads: x3_rcp_fma in {0,1,2,3} (exact tier) and tc_infer_tanh_fma in {0,1} (16-bit tier),
40,960 windows (the benchmark step), one full round (148 x 128) and one short round (148 x 32), + error vs the FFMA kernels."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from neural_speech_decoding_b200 import _lib, ops
from neural_speech_decoding_b200.lstm_eeg_model import EEG_LSTM
ck = np.load('tests/golden/checkpoint_3class.npz')
sd = {str(k): torch.from_numpy(ck[str(k)].copy()) for k in ck['__order__']}
dev = torch.device('cuda:0')
m = EEG_LSTM(); m.load_state_dict(sd); m = m.to(dev).eval()
x = torch.randn(40960, 625, 8, device=dev) * 2.73
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
with torch.inference_mode():
    ops.EXACT_TC = False
    ref = m.decode(x[:2048])[0].clone()
    ops.EXACT_TC = True
    for nr in (0, 1, 2, 3):
        _lib.call("na_set_tuning", b"x3_rcp_fma", nr)
        err = float((m.decode(x[:2048])[0] - ref).abs().max() / ref.abs().max())
        print(f"x3 rcp_fma={nr}: 40960 {t(lambda: m.decode(x)):.3f} ms  full round {t(lambda: m.decode(x[:18944])):.3f} ms  short {t(lambda: m.decode(x[:4736])):.3f} ms  err vs FFMA {err:.2e}")
    _lib.call("na_set_tuning", b"x3_rcp_fma", 1)
    m.compute_dtype = torch.bfloat16
    for nt in (0, 1):
        _lib.call("na_set_tuning", b"tc_infer_tanh_fma", nt)
        err = float((m.decode(x[:2048])[0] - ref).abs().max() / ref.abs().max())
        print(f"v2 tanh_fma={nt}: 40960 {t(lambda: m.decode(x)):.3f} ms  full round {t(lambda: m.decode(x[:18944])):.3f} ms  short {t(lambda: m.decode(x[:4736])):.3f} ms  err vs FFMA {err:.2e}")
