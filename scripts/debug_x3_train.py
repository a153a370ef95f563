"""Piece-by-piece check of the fp32-accurate tensor-core training kernels (na_train_x3.cu) against the FFMA / generic
fp32 kernels on the same weights and windows: saved activations, pooled vector, d(gates), din, weight gradients."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from neural_speech_decoding_b200 import ops
from neural_speech_decoding_b200.lstm_eeg_model import EEG_LSTM

ck = np.load('tests/golden/checkpoint_3class.npz')
sd = {str(k): torch.from_numpy(ck[str(k)].copy()) for k in ck['__order__']}
dev = torch.device('cuda:0')
torch.manual_seed(0)
B, T = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (200, 40)
drop = "--drop" in sys.argv
m = EEG_LSTM(); m.load_state_dict(sd); m = m.to(dev)
x = (torch.randn(B, T, 8) * 2.73).to(dev)
y = torch.randint(0, 3, (B,)).to(dev)
lstm = [m.lstm.layer(l) for l in range(2)]
head = [ops._f32c(t.detach()) for t in m._head_params()]
flat = [t.detach() for l in lstm for t in l]


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def tclx(t, B):      # [T,NT,12,128,8] fp16 -> [T,B,48] fp32
    T_, NT = t.shape[0], t.shape[1]
    v = t[:, :, :6].float() + t[:, :, 6:].float()
    return v.permute(0, 1, 3, 2, 4).reshape(T_, NT * 128, 48)[:, :B]


def tcl32(t, B):
    T_, NT = t.shape[0], t.shape[1]
    return t.permute(0, 1, 3, 2, 4).reshape(T_, NT * 128, 48)[:, :B]


def dgx(t, B):       # [T,NT,48,128,8] -> [T,B,192] in torch gate order (i|f|g|o blocks of 48)
    T_, NT = t.shape[0], t.shape[1]
    v = (t[:, :, :24].float() + t[:, :, 24:].float()).permute(0, 1, 3, 2, 4).reshape(T_, NT * 128, 192)[:, :B]
    n = torch.arange(192, device=t.device)
    j, gate = (n // 16) * 4 + (n % 4), (n % 16) // 4
    out = torch.empty_like(v)
    out[:, :, gate * 48 + j] = v
    return out


with torch.no_grad():
    Bp32, Bp = ops.padded_batch(B, 32), ops.padded_batch(B, 128)
    keep = (torch.rand(T, Bp, 48, device=dev) >= 0.6) if drop else None
    scale1 = 2.5 if drop else 1.0
    # ---- reference: FFMA / generic fp32 kernels ----
    xt = ops.window_zscore(x, T, T, False, True, False)
    wt0, b0 = ops.pack_lstm_layer(*lstm[0]); wt1, b1 = ops.pack_lstm_layer(*lstm[1])
    mask32 = None
    if drop:
        mask32 = torch.zeros(T, Bp32, 48, device=dev); mask32[:, :B] = keep[:, :B].float()
    rh0, rc0, rg0, rh0d = ops.lstm_layer_fwd(xt, wt0, b0, mask32, scale1, True)
    in1 = rh0d if drop else rh0
    rh1, rc1, rg1, _ = ops.lstm_layer_fwd(in1, wt1, b1, None, 1.0, True)
    rlogits, _, rstats, rz = ops.head_fwd(rh1, B, head, None, None, 1.0, False, True)
    dlogits = torch.softmax(rlogits, 1); dlogits[torch.arange(B), y] -= 1; dlogits /= B
    rdh, rdparams = ops.head_bwd(dlogits, rh1, rstats, rz, head, None, None, 1.0)
    rdg1, rdin1 = ops.lstm_layer_bwd(rdh, rg1, rc1, lstm[1][0].detach(), lstm[1][1].detach(), mask32, scale1, True)
    rdw1 = ops.lstm_layer_wgrad(rdg1, in1, rh1)
    rdg0, _ = ops.lstm_layer_bwd(rdin1, rg0, rc0, lstm[0][0].detach(), lstm[0][1].detach(), None, 1.0, False)
    rdw0 = ops.lstm_layer_wgrad(rdg0, xt, rh0)
    # ---- x3 ----
    xs = ops.x3_split_input(x, Bp)
    packed = ops.decoder_pack_x3(flat)
    mask8 = keep.to(torch.uint8).contiguous() if drop else None
    h0, h0d, c0, _, _ = ops.lstm_fwd_train_x3(0, xs, packed, head[0], head[1], mask8, 0, 65536, scale1, B)
    torch.cuda.synchronize(); print("fwd L0 ran")
    print("h0", rel(tclx(h0, B), rh0[:, :B]), "c0", rel(tcl32(c0, B), rc0[:, :B]))
    if drop: print("h0d", rel(tclx(h0d, B), rh0d[:, :B]))
    i1 = h0d if drop else h0
    h1, _, c1, z, stats = ops.lstm_fwd_train_x3(1, i1, packed, head[0], head[1], None, 0, 65536, 1.0, B)
    torch.cuda.synchronize(); print("fwd L1 ran")
    print("h1", rel(tclx(h1, B), rh1[:, :B]), "c1", rel(tcl32(c1, B), rc1[:, :B]), "z", rel(z, rz), "stats", rel(stats, rstats))
    logits, _ = ops.head_tail_fwd(z, head, None, None, 1.0, False)
    print("logits", rel(logits, rlogits))
    dz, dpar = ops.head_tail_bwd(dlogits.contiguous(), z, head, None, None, 1.0)
    dz, s_t, _ = ops._scale_dz(dz)
    s = float(s_t.item())
    din1, dg1, d_attn = ops.lstm_bwd_x3(1, i1, h1, c1, None, packed, mask8, 0, 65536, scale1, [dz, stats, z, head[0], head[1]], B)
    torch.cuda.synchronize(); print("bwd L1 ran")
    print("dg1", rel(dgx(dg1, B) / s, rdg1[:, :B]), "din1", rel(tcl32(din1, B) / s, rdin1[:, :B]),
          "d_attn_w", rel(d_attn[:48] / s, rdparams[:48]), "d_attn_b", float(d_attn[48] / s), float(rdparams[48]))
    dw = ops.lstm_wgrad_x3(1, dg1, i1, h1)
    torch.cuda.synchronize(); print("wgrad L1 ran")
    print("dW_ih1", rel(dw[0] / s, rdw1[0]), "dW_hh1", rel(dw[1] / s, rdw1[1]), "db1", rel(dw[2] / s, rdw1[2]))
    _, dg0, _ = ops.lstm_bwd_x3(0, xs, h0, c0, din1, packed, None, 0, 65536, 1.0, [], B)
    torch.cuda.synchronize(); print("bwd L0 ran")
    print("dg0", rel(dgx(dg0, B) / s, rdg0[:, :B]), "finite", bool(torch.isfinite(dg0.float()).all()), "max|dg0|", float(dg0.float().abs().max()),
          "max|dg1|", float(dg1.float().abs().max()), "s", s)
    dw = ops.lstm_wgrad_x3(0, dg0, xs, h0)
    torch.cuda.synchronize(); print("wgrad L0 ran")
    print("dW_ih0", rel(dw[0] / s, rdw0[0]), "dW_hh0", rel(dw[1] / s, rdw0[1]), "db0", rel(dw[2] / s, rdw0[2]))

# ---- end to end through the module: x3 training tier vs the FFMA tier ----
def grads(flag):
    ops.EXACT_TC_TRAIN = flag
    m.zero_grad()
    m.eval()
    out = m(x)
    torch.nn.functional.cross_entropy(out, y).backward()
    return out.detach().clone(), {k: p.grad.clone() for k, p in m.named_parameters()}
la, ga = grads(True)
lb, gb = grads(False)
print("module logits", rel(la, lb))
for k in ga: print(f"  grad {k:22s} {rel(ga[k], gb[k]):.3e}")
print("done")
