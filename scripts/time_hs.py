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
outs = {}
for N in (18944, 40960):
    x = torch.randn(N, 625, 8, device=dev) * 2.73
    with torch.inference_mode():
        xt = ops.window_zscore(x, 625, 625, False, True, 2, 128)
        packed = m._packed_tc(); head = m._head_params()
        for hs in (1, 2):
            _lib.call("na_set_tuning", b"tc_infer_hs", hs)
            ms = t(lambda: ops.decoder_infer_bf16(xt, packed, head, N, True))
            outs[(N, hs)] = ops.decoder_infer_bf16(xt, packed, head, N, True)[0].cpu().numpy()
            print(f"N={N} HS={hs}: {ms:.3f} ms -> {N/ms*1e3/1e6:.3f} M windows/s")
    print("  max |HS1-HS2| =", np.abs(outs[(N, 1)] - outs[(N, 2)]).max())
