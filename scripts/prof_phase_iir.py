import sys, torch
sys.path.insert(0, '.')
from neural_speech_decoding_b200 import filters
from neural_speech_decoding_b200.preprocess_gpu import PhaseCouplingFilterGPU
dev = torch.device('cuda:0')
x = torch.randn(40960, 625, 8, device=dev) * 30 + 5
pre = PhaseCouplingFilterGPU(device=dev, accept_noncommercial_terms=True)
for _ in range(2):
    y = filters.filter_windows(x)
    z = pre.transform_batch(x)
torch.cuda.synchronize()
print("ok", float(y.abs().max()), float(z.abs().max()))
