"""Soak test of the persistent tensor-core training kernels: random (B, T), both tiers, full and half tiles, capped grids
(several tiles per CTA), train mode.  Every run must finish (no mbarrier deadlock), give finite gradients and logits that
are bit-identical to the uncapped run.  python scripts/soak_train.py [iterations] [seed]"""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from neural_speech_decoding_b200 import _lib, ops
from neural_speech_decoding_b200.lstm_eeg_model import EEG_LSTM
ck = np.load('tests/golden/checkpoint_3class.npz')
sd = {str(k): torch.from_numpy(ck[str(k)].copy()) for k in ck['__order__']}
dev = torch.device('cuda:0')
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
models = {}
for name, dt in (("fp16", torch.bfloat16), ("exact", torch.float32)):
    m = EEG_LSTM(); m.load_state_dict(sd); m = m.to(dev); m.compute_dtype = dt
    models[name] = m
def run(m, x, y, cap, train):
    _lib.call("na_set_tuning", b"train_max_ctas", cap)
    m.train(train); m.zero_grad()
    torch.manual_seed(5)
    out = m(x)
    torch.nn.functional.cross_entropy(out, y).backward()
    g = torch.cat([p.grad.flatten() for p in m.parameters()])
    return out.detach().clone(), g.clone()
t0 = time.time()
for it in range(iters):
    B = int(rng.choice([1, 2, 63, 64, 65, 127, 128, 129, 300, 777, 1500, 2500, int(rng.integers(1, 3000))]))
    T = int(rng.choice([1, 2, 3, 4, 5, 7, 16, 33, int(rng.integers(1, 90))]))
    half = bool(rng.integers(0, 2)); train = bool(rng.integers(0, 2)); cap = int(rng.integers(1, 4))
    ops.TC_HALF_TILES = half; ops.X3_HALF_TILES = half
    x = (torch.randn(B, T, 8) * 2.73).to(dev); y = torch.randint(0, 3, (B,)).to(dev)
    for name, m in models.items():
        a, ga = run(m, x, y, cap, train)
        b, gb = run(m, x, y, 0, train)
        assert torch.isfinite(a).all() and torch.isfinite(ga).all(), (name, B, T, half, train, cap)
        assert torch.equal(a, b), (name, B, T, half, train, cap, float((a - b).abs().max()))
        err = float((ga - gb).abs().max() / gb.abs().max().clamp_min(1e-30))
        # a capped grid accumulates up to B*T/128 work items per weight-gradient element serially in ONE fp32 TMEM accumulator
        # (cap = 1, B = 1500, T = 88: 2.7e-4 of the largest gradient); the uncapped path is pinned to fp64 truth by the parity tests
        assert err < 2e-3, (name, B, T, half, train, cap, err)
    if it % 10 == 0: print(f"iter {it}: B={B} T={T} half={half} train={train} cap={cap} ok ({time.time() - t0:.0f} s)", flush=True)
_lib.call("na_set_tuning", b"train_max_ctas", 0)
torch.cuda.synchronize()
print(f"soak ok: {iters} configurations x 2 tiers in {time.time() - t0:.0f} s")
