"""Multi-rank diagnosis of the 16-bit train step: host enqueue time and step time per rank, with and without the all-reduce.
torchrun --nproc-per-node N scripts/diag_train_n.py [per_gpu_batch]"""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, '.')
from neural_speech_decoding_b200.lstm_eeg_model import EEG_LSTM
from neural_speech_decoding_b200.dp import DataParallelTrainer
from neural_speech_decoding_b200.optim import FusedAdam
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ck = np.load('tests/golden/checkpoint_3class.npz')
sd = {str(k): torch.from_numpy(ck[str(k)].copy()) for k in ck['__order__']}
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
x = torch.randn(B, 625, 8, device=dev) * 2.73
y = torch.randint(0, 3, (B,), device=dev)
def run(ws, label, n=20):
    m = EEG_LSTM(); m.load_state_dict(sd); m = m.to(dev).train(); m.compute_dtype = torch.bfloat16
    tr = DataParallelTrainer(m, FusedAdam(m.parameters(), lr=1e-3), world_size=ws)
    for _ in range(3): tr.step([(x, y)], global_batch=B * world)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n): tr.step([(x, y)], global_batch=B * world)
    e1.record(); t1 = time.perf_counter()
    torch.cuda.synchronize()
    res = torch.tensor([1e3 * (t1 - t0) / n, e0.elapsed_time(e1) / n], device=dev)
    allr = [torch.zeros_like(res) for _ in range(world)]
    dist.all_gather(allr, res)
    if rank == 0:
        print(f"{label}: per-rank host enqueue ms/step {[round(float(a[0]), 2) for a in allr]}  device ms/step {[round(float(a[1]), 2) for a in allr]}", flush=True)
run(1, "no all-reduce (8 independent trainers)")
run(world, "DP: flat all-reduce + loss all-reduce")
run(1, "no all-reduce again")
print_cpu = len(os.sched_getaffinity(0))
if rank == 0: print("cpus visible to a rank:", print_cpu, "os.cpu_count", os.cpu_count(), flush=True)
dist.destroy_process_group()
