"""Time the wide (H = 192, T = 2500: BASELINE configs[4]) tensor-core forward and compare with the exact tier."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from neural_speech_decoding_b200.lstm_eeg_model import EEG_LSTM
from neural_speech_decoding_b200 import ops
dev = torch.device('cuda:0')
H = int(sys.argv[1]) if len(sys.argv) > 1 else 192
T = int(sys.argv[2]) if len(sys.argv) > 2 else 2500
Ns = [int(a) for a in sys.argv[3:]] or [2048, 18944]
torch.manual_seed(0)
m = EEG_LSTM(hidden_size=H).to(dev).eval(); m.compute_dtype = torch.bfloat16
def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
flops = 2 * T * (4 * H * (8 + H) + 4 * H * 2 * H) + 4 * T * H
from neural_speech_decoding_b200 import _lib
for N in Ns:
  for cs, dbg in ((1, 0),):
    _lib.call('na_set_tuning', b'tc_wide_cluster', cs); _lib.call('na_set_tuning', b'tc_wide_dbg', dbg)
    x = torch.randn(N, T, 8, device=dev) * 2.73
    with torch.inference_mode():
        xt = ops.window_zscore(x, T, T, False, True, 2, 128)
        packed = m._packed_tc_wide(); head = m._head_params()
        ms = t(lambda: ops.decoder_infer_wide_bf16(xt, packed, head[2:], N, H, True))
        rounds = -(-N // (148 * 128))
        print(f"cluster={cs} dbg={dbg} H={H} T={T} N={N}: {ms:.2f} ms -> {N/ms:.1f} k windows/s, {N*flops/ms*1e-9:.1f} TFLOP/s, {ms*1e3/T/rounds:.2f} us per step-round", flush=True)
        if N <= 4096:
            got = ops.decoder_infer_wide_bf16(xt, packed, head[2:], N, H, True)[0].cpu().numpy()
            m.compute_dtype = torch.float32
            n = min(N, 256)
            ms32 = t(lambda: m(x[:n]), reps=1)
            ex = m(x[:n]).cpu().numpy(); m.compute_dtype = torch.bfloat16
            err = np.abs(got[:n] - ex).max(axis=1) / np.abs(ex).max()
            print(f"   vs exact tier ({n} windows, {n/ms32:.2f} k windows/s): mean rel err {err.mean():.2e}, max {err.max():.2e}, argmax agree {(got[:n].argmax(1)==ex.argmax(1)).mean():.3f}", flush=True)
