import sys, torch
sys.path.insert(0, '.')
from neural_speech_decoding_b200 import filters, _lib
dev = torch.device('cuda:0')
x = torch.randn(40960, 625, 8, device=dev) * 30 + 5
for occ in (1, 2, 3):
    _lib.call("na_set_tuning", b"iir_occ3", occ)
    filters.filter_windows(x); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): y = filters.filter_windows(x)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"iir occ={occ}: {ms:.2f} ms per 40,960 windows -> {40960/ms*1e3/1e6:.2f} M windows/s")
