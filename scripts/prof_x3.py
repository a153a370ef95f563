import sys, numpy as np, torch
sys.path.insert(0, '.')
from neural_speech_decoding_b200.lstm_eeg_model import EEG_LSTM
ck = np.load('tests/golden/checkpoint_3class.npz')
sd = {str(k): torch.from_numpy(ck[str(k)].copy()) for k in ck['__order__']}
dev = torch.device('cuda:0')
m = EEG_LSTM(); m.load_state_dict(sd); m = m.to(dev).eval()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 18944
x = torch.randn(N, 625, 8, device=dev) * 2.73
with torch.inference_mode():
    for _ in range(3): out = m.decode(x)
    torch.cuda.synchronize()
print("ok", float(out[0].abs().max()))
