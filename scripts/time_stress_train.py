import sys, torch
sys.path.insert(0, '.')
from neural_speech_decoding_b200.lstm_eeg_model import EEG_LSTM
dev = torch.device('cuda:0')
torch.manual_seed(0)
m = EEG_LSTM(hidden_size=192).to(dev).train()
m.compute_dtype = torch.bfloat16
B, T = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (2048, 2500)
x = torch.randn(B, T, 8, device=dev) * 2.73
y = torch.randint(0, 3, (B,), device=dev)
opt = torch.optim.Adam(m.parameters(), lr=1e-3)
def step():
    opt.zero_grad(); torch.nn.functional.cross_entropy(m(x), y).backward(); opt.step()
step(); torch.cuda.synchronize()
print("peak GB after first step", torch.cuda.max_memory_allocated() / 1e9)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); step(); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"wide TC train step B={B} T={T}: {ms:.1f} ms -> {B/ms*1e3:.0f} windows/s")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
for r in sorted(prof.key_averages(), key=lambda r: -r.device_time_total)[:14]:
    print(f"   {r.device_time_total/1e3:9.3f} ms  x{r.count:3d}  {r.key[:100]}")
