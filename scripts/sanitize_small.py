"""Smallest end-to-end exercise of every kernel (for compute-sanitizer): both tiers, inference + training."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from neural_speech_decoding_b200.lstm_eeg_model import EEG_LSTM
from neural_speech_decoding_b200.tester import run_trials_batched
ck = np.load('tests/golden/checkpoint_3class.npz')
sd = {str(k): torch.from_numpy(ck[str(k)].copy()) for k in ck['__order__']}
dev = torch.device('cuda:0')
torch.manual_seed(0)
m = EEG_LSTM(); m.load_state_dict(sd); m = m.to(dev)
x = torch.randn(150, 9, 8, device=dev) * 2.73
y = torch.randint(0, 3, (150,), device=dev)
for dtype in (torch.float32, torch.bfloat16):
    m.compute_dtype = dtype
    m.eval()
    with torch.inference_mode():
        a = m(x)
        avg = run_trials_batched(x[:140].reshape(10, 14, 9, 8).cpu().numpy(), m)
    m.train()
    m.zero_grad()
    torch.nn.functional.cross_entropy(m(x), y).backward()
    torch.cuda.synchronize()
    print(dtype, float(a.abs().sum()), float(avg.sum()), float(sum(p.grad.abs().sum() for p in m.parameters())))
print("done")
