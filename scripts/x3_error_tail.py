import sys, numpy as np, torch
sys.path.insert(0, '.')
from neural_speech_decoding_b200 import ops
from neural_speech_decoding_b200.lstm_eeg_model import EEG_LSTM
ck = np.load('tests/golden/checkpoint_3class.npz')
sd = {str(k): torch.from_numpy(ck[str(k)].copy()) for k in ck['__order__']}
dev = torch.device('cuda:0')
m = EEG_LSTM(); m.load_state_dict(sd); m = m.to(dev).eval()
from oracle.torch_ref import RefEEGLSTM
ref = RefEEGLSTM().double().eval(); ref.load_state_dict({k: v.double() for k, v in sd.items()})
gen = torch.Generator(device="cpu").manual_seed(21)
for B, T in ((23675, 30), (4096, 625), (20000, 8)):
    x = torch.randn(B, T, 8, generator=gen) * 2.73
    with torch.inference_mode():
        ops.EXACT_TC = True; a = m(x.to(dev)).cpu().numpy()
        ops.EXACT_TC = False; b = m(x.to(dev)).cpu().numpy()
        n = min(B, 4096)
        t = ref(x[:n].double()).numpy()
    s = np.abs(t).max()
    print(f"B={B} T={T}: x3 vs FFMA max {np.abs(a-b).max()/np.abs(b).max():.2e}; vs fp64 truth on {n}: x3 max {np.abs(a[:n]-t).max()/s:.2e} mean {np.abs(a[:n]-t).mean()/s:.2e}; FFMA max {np.abs(b[:n]-t).max()/s:.2e} mean {np.abs(b[:n]-t).mean()/s:.2e}")
