import torch, time
dev = torch.device('cuda:0')
n = 40960 * 625 * 8
h = torch.empty(n, dtype=torch.float32, pin_memory=True); h.normal_()
d = torch.empty(n, dtype=torch.float32, device=dev)
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
for chunks in (1, 2, 10, 40):
    c = n // chunks
    def f():
        for i in range(chunks): d[i*c:(i+1)*c].copy_(h[i*c:(i+1)*c], non_blocking=True)
    s = t(f)
    print(f"H2D {n*4/1e6:.0f} MB in {chunks:3d} chunk(s): {s*1e3:.2f} ms = {n*4/s/1e9:.1f} GB/s")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def f2():
    half = n // 2
    with torch.cuda.stream(s1): d[:half].copy_(h[:half], non_blocking=True)
    with torch.cuda.stream(s2): d[half:].copy_(h[half:], non_blocking=True)
s = t(f2); print(f"two streams: {s*1e3:.2f} ms = {n*4/s/1e9:.1f} GB/s")
import os; print("cpus", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
