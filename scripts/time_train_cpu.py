"""How much HOST time does one 16-bit train step cost (enqueue only), against its GPU time?  python scripts/time_train_cpu.py [B]"""
import sys, time, cProfile, pstats, io
import numpy as np, torch
sys.path.insert(0, '.')
from neural_speech_decoding_b200.lstm_eeg_model import EEG_LSTM
from neural_speech_decoding_b200.dp import DataParallelTrainer
from neural_speech_decoding_b200.optim import FusedAdam
ck = np.load('tests/golden/checkpoint_3class.npz')
sd = {str(k): torch.from_numpy(ck[str(k)].copy()) for k in ck['__order__']}
dev = torch.device('cuda:0')
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
for dtype in (torch.bfloat16, torch.float32):
    m = EEG_LSTM(); m.load_state_dict(sd); m = m.to(dev).train(); m.compute_dtype = dtype
    x = torch.randn(B, 625, 8, device=dev) * 2.73
    y = torch.randint(0, 3, (B,), device=dev)
    tr = DataParallelTrainer(m, FusedAdam(m.parameters(), lr=1e-3), world_size=1)
    for _ in range(3): tr.step([(x, y)], global_batch=B)
    torch.cuda.synchronize()
    n = 20
    t0 = time.perf_counter()
    for _ in range(n): tr.step([(x, y)], global_batch=B)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"{dtype}: B={B}: enqueue {1e3 * (t1 - t0) / n:.3f} ms/step (host), total {1e3 * (t2 - t0) / n:.3f} ms/step")
    # host-only cost with the GPU idle between steps (sync each step): pure python + launch latency
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(10):
        tr.step([(x, y)], global_batch=B)
    pr.disable()
    torch.cuda.synchronize()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(28)
    print("\n".join(l[:150] for l in s.getvalue().splitlines()[:60]))
