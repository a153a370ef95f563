import sys, numpy as np, torch
sys.path.insert(0, '.')
from neural_speech_decoding_b200.lstm_eeg_model import EEG_LSTM
from neural_speech_decoding_b200 import ops
ck = np.load('tests/golden/checkpoint_3class.npz')
sd = {str(k): torch.from_numpy(ck[str(k)].copy()) for k in ck['__order__']}
dev = torch.device('cuda:0')
m = EEG_LSTM(); m.load_state_dict(sd); m = m.to(dev).eval()
m.compute_dtype = torch.bfloat16
N = int(sys.argv[1]) if len(sys.argv) > 1 else 40960
x = torch.randn(N, 625, 8, device=dev) * 2.73
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
with torch.inference_mode():
    ms = t(lambda: m.decode(x))
    print(f"bf16 decode {N} windows: {ms:.3f} ms -> {N/ms*1e3/1e6:.3f} M windows/s")
    xt = ops.window_zscore(x, 625, 625, False, True, 2, 128)
    packed = m._packed_tc(); head = m._head_params()
    ms_k = t(lambda: ops.decoder_infer_bf16(xt, packed, head, N, True))
    ms_p = t(lambda: ops.window_zscore(x, 625, 625, False, True, 2, 128))
    print(f"  tcgen05 kernel alone {ms_k:.3f} ms ({ms_k*1e3/625/((N+128*148-1)//(128*148)):.3f} us per step per tile-round), pack kernel {ms_p:.3f} ms")
