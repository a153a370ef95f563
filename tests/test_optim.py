"""optim.FusedAdam (one launch for the whole decoder) against torch.optim.Adam."""
import numpy as np
import pytest
import torch


@pytest.mark.gpu
def test_fused_adam_matches_torch_adam(checkpoint):
    from neural_speech_decoding_b200.lstm_eeg_model import EEG_LSTM
    from neural_speech_decoding_b200.optim import FusedAdam
    dev = torch.device("cuda:0")
    a, b = EEG_LSTM(), EEG_LSTM()
    a.load_state_dict(checkpoint, strict=True)
    b.load_state_dict(checkpoint, strict=True)
    a, b = a.to(dev), b.to(dev)
    for wd in (0.0, 0.01):
        oa = FusedAdam(a.parameters(), lr=1e-3, weight_decay=wd).attach(a)
        ob = torch.optim.Adam(b.parameters(), lr=1e-3, weight_decay=wd)
        gen = torch.Generator(device="cpu").manual_seed(3)
        for step in range(6):
            for p, q in zip(a.parameters(), b.parameters()):
                g = (torch.randn(p.shape, generator=gen) * (10.0 ** (step % 3 - 2))).to(dev)
                p.grad, q.grad = g.clone(), g.clone()
            oa.step()
            ob.step()
        for (k, p), q in zip(a.named_parameters(), b.parameters()):
            d = (p - q).abs()
            # Adam's first update is lr * g' / (|g'| + eps): where g' = g + wd * p cancels to ~eps it is ill-conditioned (one ulp of
            # g moves the update by up to lr * ulp / eps ~ 6e-5, and torch's foreach kernels round g' differently: measured
            # 4.6e-6 on ONE element of lstm.weight_hh_l1, scripts/debug_adam.py).  Everything else agrees to 2e-6.
            tol = 2e-6 * max(1.0, q.abs().max().item())
            assert d.max().item() <= 1e-4 and (d > tol).sum().item() <= 2, k
    # grad_scale: a device scalar folded into every gradient (the unscale of a loss-scaled backward)
    oa = FusedAdam(a.parameters(), lr=1e-2)
    ob = torch.optim.Adam(b.parameters(), lr=1e-2)
    b.load_state_dict(a.state_dict())
    for p, q in zip(a.parameters(), b.parameters()):
        g = torch.ones_like(p)
        p.grad, q.grad = g * 4.0, g.clone()
    oa.step(grad_scale=torch.tensor([0.25], device=dev))
    ob.step()
    for p, q in zip(a.parameters(), b.parameters()):
        assert torch.allclose(p, q, rtol=0, atol=2e-7)
    # a training loop through the trainer with the fused optimizer still converges on a tiny problem and keeps eval in sync
    from neural_speech_decoding_b200.dp import DataParallelTrainer
    torch.manual_seed(0)
    m = EEG_LSTM().to(dev)
    x = torch.randn(64, 40, 8, device=dev) * 2.73
    y = (x[:, :, 0].mean(1) > 0).long()
    tr = DataParallelTrainer(m, FusedAdam(m.parameters(), lr=5e-3), world_size=1)
    m.eval()
    with torch.no_grad():
        before = m(x).clone()
    m.train()
    losses = [tr.step([(x, y)], global_batch=64).item() for _ in range(25)]
    assert losses[-1] < losses[0]
    m.eval()
    with torch.no_grad():
        after = m(x)
    assert not torch.equal(before, after)            # the packed-weight cache was dropped: eval sees the updated weights
