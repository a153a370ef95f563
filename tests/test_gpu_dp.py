"""Data-parallel step over NCCL on >= 2 real GPUs: DP gradients == single-GPU gradients on the
concatenated batch; ranks stay in lock-step.  Skipped on a 1-GPU box."""
import os
import sys
import tempfile
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, str(ROOT))
    import torch.distributed as dist
    from tests.conftest import load_checkpoint
    from neural_speech_decoding_b200.lstm_eeg_model import EEG_LSTM
    from neural_speech_decoding_b200.dp import DataParallelTrainer, shard_batch
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    g = torch.Generator().manual_seed(5)
    X = torch.randn(64, 50, 8, generator=g) * 2.73
    Y = torch.randint(0, 3, (64,), generator=g)
    m = EEG_LSTM()
    m.load_state_dict(load_checkpoint(), strict=True)
    m = m.to(dev).eval()
    tr = DataParallelTrainer(m, torch.optim.SGD(m.parameters(), lr=0.1), world_size=world)
    sl = shard_batch(64, rank, world)
    xs, ys = X[sl].to(dev), Y[sl].to(dev)
    loss = tr.step([(xs[:10], ys[:10]), (xs[10:], ys[10:])], global_batch=64)
    np.savez(Path(out_dir) / f"rank{rank}.npz", loss=loss.cpu().numpy(), grad=tr.bucket.flat.cpu().numpy(),
             **{k: v.detach().cpu().numpy() for k, v in m.state_dict().items()})
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_nccl_dp_equals_single_gpu(checkpoint):
    from neural_speech_decoding_b200.lstm_eeg_model import EEG_LSTM
    from neural_speech_decoding_b200.dp import DataParallelTrainer
    world = 2
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(world, 29600 + os.getpid() % 2000, d), nprocs=world, join=True)
        r = [np.load(Path(d) / f"rank{i}.npz") for i in range(world)]
    g = torch.Generator().manual_seed(5)
    X = torch.randn(64, 50, 8, generator=g) * 2.73
    Y = torch.randint(0, 3, (64,), generator=g)
    dev = torch.device("cuda:0")
    m = EEG_LSTM()
    m.load_state_dict(checkpoint, strict=True)
    m = m.to(dev).eval()
    tr = DataParallelTrainer(m, torch.optim.SGD(m.parameters(), lr=0.1), world_size=1)
    loss = tr.step([(X.to(dev), Y.to(dev))], global_batch=64)
    ref = tr.bucket.flat.cpu().numpy()
    assert np.array_equal(r[0]["grad"], r[1]["grad"])
    assert np.abs(r[0]["grad"] - ref).max() / np.abs(ref).max() < 1e-6
    assert abs(float(r[0]["loss"]) - loss.item()) < 1e-6
    for k in m.state_dict():
        assert np.array_equal(r[0][k], r[1][k]), k
