"""Time the tcgen05 decoder kernel alone at batch sizes that are whole / partial tile rounds (148 SMs x 128 windows).
usage: time_tiles.py [hs=2,3] [rep=0,1] [N ...]"""
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
hss, reps_, Ns = (3,), (0, 1), []
for a in sys.argv[1:]:
    if a.startswith('hs='): hss = tuple(int(v) for v in a[3:].split(','))
    elif a.startswith('rep='): reps_ = tuple(int(v) for v in a[4:].split(','))
    else: Ns.append(int(a))
Ns = Ns or [32, 1024, 4096, 4736, 9472, 18944, 23680, 37888, 40960, 56832]
for N in Ns:
    x = torch.randn(N, 625, 8, device=dev) * 2.73
    with torch.inference_mode():
        xt = ops.window_zscore(x, 625, 625, False, True, 2, 128)
        packed = m._packed_tc(); head = m._head_params()
        ref = None
        for hs in hss:
            for rp in (reps_ if hs == 3 else (0,)):
                _lib.call("na_set_tuning", b"tc_infer_hs", hs); _lib.call("na_set_tuning", b"tc_infer_rep", rp)
                ms = t(lambda: ops.decoder_infer_bf16(xt, packed, head, N, True))
                out = ops.decoder_infer_bf16(xt, packed, head, N, True)[0].cpu().numpy()
                if ref is None: ref = out
                print(f"N={N} ({N/18944:.2f} rounds) hs={hs} rep={rp}: {ms:.3f} ms -> {N/ms*1e3/1e6:.3f} M windows/s   max|d logits| vs first = {np.abs(out-ref).max():.2e}", flush=True)
