import sys, numpy as np, torch
sys.path.insert(0, '.')
from neural_speech_decoding_b200.lstm_eeg_model import EEG_LSTM
ck = np.load('tests/golden/checkpoint_3class.npz')
sd = {str(k): torch.from_numpy(ck[str(k)].copy()) for k in ck['__order__']}
dev = torch.device('cuda:0')
m = EEG_LSTM(); m.load_state_dict(sd); m = m.to(dev).train()
m.compute_dtype = torch.bfloat16
for B in [int(a) for a in sys.argv[1:]] or [18944]:
    x = torch.randn(B, 625, 8, device=dev) * 2.73
    y = torch.randint(0, 3, (B,), device=dev)
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    def step():
        opt.zero_grad(); torch.nn.functional.cross_entropy(m(x), y).backward(); opt.step()
    step(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"TC train step B={B}: {ms:.2f} ms -> {B/ms*1e3:.0f} windows/s")
    if "--profile" in sys.argv or True:
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            step(); torch.cuda.synchronize()
        rows = sorted(prof.key_averages(), key=lambda r: -r.device_time_total)[:26]
        allk = [r for r in prof.key_averages() if r.device_time_total > 0]
        print(f"   device total {sum(r.device_time_total for r in allk)/1e3:.3f} ms in {sum(r.count for r in allk)} launches")
        for r in rows: print(f"   {r.device_time_total/1e3:8.3f} ms  x{r.count:3d}  {r.key[:90]}")
