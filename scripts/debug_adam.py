import sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from conftest import load_checkpoint
from neural_speech_decoding_b200.lstm_eeg_model import EEG_LSTM
from neural_speech_decoding_b200.optim import FusedAdam
dev = torch.device('cuda:0')
ck = load_checkpoint()
a, b, c = EEG_LSTM(), EEG_LSTM(), EEG_LSTM()
for m in (a, b, c): m.load_state_dict(ck, strict=True)
a, b, c = a.to(dev), b.to(dev), c.to(dev).double()
for wd in (0.0, 0.01):
    oa = FusedAdam(a.parameters(), lr=1e-3, weight_decay=wd).attach(a)
    ob = torch.optim.Adam(b.parameters(), lr=1e-3, weight_decay=wd)
    oc = torch.optim.Adam(c.parameters(), lr=1e-3, weight_decay=wd)
    gen = torch.Generator(device="cpu").manual_seed(3)
    for step in range(6):
        for p, q, r in zip(a.parameters(), b.parameters(), c.parameters()):
            g = (torch.randn(p.shape, generator=gen) * (10.0 ** (step % 3 - 2))).to(dev)
            p.grad, q.grad, r.grad = g.clone(), g.clone(), g.double()
        oa.step(); ob.step(); oc.step()
        with torch.no_grad():
            w1 = max((float((p - q).abs().max()), k) for (k, p), q in zip(a.named_parameters(), b.parameters()))
            w2 = max((float((p.double() - r).abs().max()), k) for (k, p), r in zip(a.named_parameters(), c.parameters()))
            w3 = max((float((q.double() - r).abs().max()), k) for (k, q), r in zip(b.named_parameters(), c.parameters()))
        print(wd, step, "fused-torch", w1, "fused-f64", w2, "torch-f64", w3)
    with torch.no_grad():
        for (k, p), q, r in zip(a.named_parameters(), b.parameters(), c.parameters()):
            d = (p - q).abs(); i = int(d.argmax())
            print("  ", k, float(d.max()), "p32", float(p.flatten()[i]), "q32", float(q.flatten()[i]), "r64", float(r.flatten()[i]), "g", float(q.grad.flatten()[i]),
                  "m", float(ob.state[q]["exp_avg"].flatten()[i]), "v", float(ob.state[q]["exp_avg_sq"].flatten()[i]),
                  "v64", float(oc.state[r]["exp_avg_sq"].flatten()[i]))
