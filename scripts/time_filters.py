import sys, torch
sys.path.insert(0, '.')
from neural_speech_decoding_b200 import filters
dev = torch.device('cuda:0')
for B in (4096, 40960):
    x = torch.randn(B, 625, 8, device=dev) * 30 + 5
    filters.filter_windows(x); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): y = filters.filter_windows(x)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    byt = B * 8 * 625 * (4 + 4 + 7 * 16)
    print(f"filter chain B={B}: {ms:.2f} ms -> {B/ms*1e3/1e6:.2f} M windows/s, {byt/ms/1e6:.0f} GB/s algorithmic")
