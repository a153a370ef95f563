import sys, numpy as np, torch
sys.path.insert(0, '.')
from neural_speech_decoding_b200.lstm_eeg_model import EEG_LSTM
ck = np.load('tests/golden/checkpoint_3class.npz')
sd = {str(k): torch.from_numpy(ck[str(k)].copy()) for k in ck['__order__']}
dev = torch.device('cuda:0')
m = EEG_LSTM(); m.load_state_dict(sd); m = m.to(dev).eval()
gen = torch.Generator().manual_seed(0)
for (ntile, T) in [(160, 12), (148, 12), (300, 12), (4, 625), (160, 100)]:
    xs = torch.randn(ntile * 128, T, 8, generator=gen) * 2.73
    with torch.inference_mode():
        m.compute_dtype = torch.bfloat16; a = m(xs.to(dev)).cpu().numpy()
        m.compute_dtype = torch.float32; b = m(xs.to(dev)).cpu().numpy()
    err = np.abs(a - b).max(axis=1).reshape(ntile, 128)
    pt = err.max(axis=1)
    print(f"ntile={ntile} T={T} scale={np.abs(b).max():.2f} max err={err.max():.4f} mean err={err.mean():.5f} "
          f"first-round tiles max={pt[:148].max():.4f} later tiles max={pt[148:].max() if ntile > 148 else 0:.4f} "
          f"argmax-mismatch={(a.argmax(1) != b.argmax(1)).mean():.5f}")
    worst = np.argsort(pt)[-3:]
    print("  worst tiles", worst, pt[worst], " rows of worst:", np.argsort(err[worst[-1]])[-3:])
