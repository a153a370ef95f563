import sys, torch
sys.path.insert(0, '.')
from neural_speech_decoding_b200.lstm_eeg_model import EEG_LSTM
from neural_speech_decoding_b200 import ops, _lib
dev = torch.device('cuda:0'); H = 192; T = int(sys.argv[1]) if len(sys.argv) > 1 else 500; N = 18944
cs = int(sys.argv[2]) if len(sys.argv) > 2 else 1
torch.manual_seed(0)
m = EEG_LSTM(hidden_size=H).to(dev).eval(); m.compute_dtype = torch.bfloat16
_lib.call('na_set_tuning', b'tc_wide_cluster', cs)
x = torch.randn(N, T, 8, device=dev) * 2.73
with torch.inference_mode():
    xt = ops.window_zscore(x, T, T, False, True, 2, 128)
    packed = m._packed_tc_wide(); head = m._head_params()
    for _ in range(3): out = ops.decoder_infer_wide_bf16(xt, packed, head[2:], N, H, True)
    torch.cuda.synchronize()
print("ok", float(out[0].abs().max()))
