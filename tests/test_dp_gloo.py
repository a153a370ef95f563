"""world_size-2 data-parallel step over gloo on CPU (host logic of dp.py): the all-reduced DP
gradients equal the single-process gradients on the concatenated batch, and both ranks end the
optimizer step with identical weights.  Kernels are replaced by tests/fake_lib (no GPU here)."""
import os
import sys
import tempfile
from pathlib import Path

import numpy as np
import torch
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, str(ROOT))
    import torch.distributed as dist
    from tests.fake_lib import install
    from tests.conftest import load_checkpoint
    install()
    from neural_speech_decoding_b200.lstm_eeg_model import EEG_LSTM
    from neural_speech_decoding_b200.dp import DataParallelTrainer, shard_batch
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    g = torch.Generator().manual_seed(5)
    X = torch.randn(12, 15, 8, generator=g) * 2.73
    Y = torch.randint(0, 3, (12,), generator=g)
    m = EEG_LSTM()
    m.load_state_dict(load_checkpoint(), strict=True)
    m.eval()                                          # deterministic (autograd works in eval mode)
    tr = DataParallelTrainer(m, torch.optim.SGD(m.parameters(), lr=0.1), world_size=world)
    sl = shard_batch(12, rank, world)
    xs, ys = X[sl], Y[sl]
    loss = tr.step([(xs[:3], ys[:3]), (xs[3:], ys[3:])], global_batch=12)     # two micro-batches per rank
    np.savez(Path(out_dir) / f"rank{rank}.npz", loss=loss.numpy(), grad=tr.bucket.flat.numpy(),
             **{k: v.detach().numpy() for k, v in m.state_dict().items()})
    dist.destroy_process_group()


def test_dp_two_ranks_equals_single_process(cpu_backend, checkpoint):
    from neural_speech_decoding_b200.lstm_eeg_model import EEG_LSTM
    from neural_speech_decoding_b200.dp import DataParallelTrainer
    with tempfile.TemporaryDirectory() as d:
        port = 29500 + (os.getpid() % 2000)
        mp.spawn(_worker, args=(2, port, d), nprocs=2, join=True)
        r0, r1 = np.load(Path(d) / "rank0.npz"), np.load(Path(d) / "rank1.npz")
    g = torch.Generator().manual_seed(5)
    X = torch.randn(12, 15, 8, generator=g) * 2.73
    Y = torch.randint(0, 3, (12,), generator=g)
    m = EEG_LSTM()
    m.load_state_dict(checkpoint, strict=True)
    m.eval()
    tr = DataParallelTrainer(m, torch.optim.SGD(m.parameters(), lr=0.1), world_size=1)
    loss = tr.step([(X, Y)], global_batch=12)
    assert np.array_equal(r0["grad"], r1["grad"])                      # same reduced bucket on both ranks
    scale = np.abs(tr.bucket.flat.numpy()).max()
    assert np.abs(r0["grad"] - tr.bucket.flat.numpy()).max() / scale < 1e-6
    assert abs(float(r0["loss"]) - loss.item()) < 1e-6 and float(r0["loss"]) == float(r1["loss"])
    for k, v in m.state_dict().items():
        assert np.array_equal(r0[k], r1[k]), k                          # ranks stay in lock-step
        assert np.abs(r0[k] - v.numpy()).max() < 1e-6, k


def test_flat_bucket_views_and_shards(cpu_backend, checkpoint):
    from neural_speech_decoding_b200.lstm_eeg_model import EEG_LSTM
    from neural_speech_decoding_b200.dp import FlatGradBucket, shard_batch
    m = EEG_LSTM()
    b = FlatGradBucket(list(m.parameters()))
    assert b.flat.numel() == 31764
    off = 0
    for p in m.parameters():
        assert p.grad.data_ptr() == b.flat[off:].data_ptr()
        off += p.numel()
    assert [shard_batch(10, r, 4) for r in range(4)] == [slice(0, 3), slice(3, 6), slice(6, 9), slice(9, 10)]
    assert shard_batch(2, 3, 4) == slice(2, 2)
