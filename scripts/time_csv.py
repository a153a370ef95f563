import io, sys, numpy as np, torch
sys.path.insert(0, '.')
from neural_speech_decoding_b200 import ingest, _lib, ops
dev = torch.device('cuda:0')
rng = np.random.default_rng(0)
files = []
for i in range(64):
    b = io.BytesIO(); np.savetxt(b, rng.standard_normal((625, 8)) * 2.73, delimiter=",", fmt="%.7f"); files.append(b.getvalue())
files = files * 64
off = np.zeros(len(files) + 1, dtype=np.int64); np.cumsum([len(f) for f in files], out=off[1:])
text = torch.frombuffer(bytearray(b"".join(files)), dtype=torch.uint8).to(dev); offd = torch.from_numpy(off).to(dev)
n = len(files); out = torch.empty((n, 625, 8), device=dev); st = torch.empty((n, 2), dtype=torch.int32, device=dev)
mx = int((off[1:] - off[:-1]).max())
def run(): _lib.call("na_csv_parse_f32", text.data_ptr(), offd.data_ptr(), out.data_ptr(), st.data_ptr(), n, 5000, mx, ops._stream())
run(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
byt = int(off[-1]) + n * 5000 * 4
print(f"csv parse kernel: {ms:.3f} ms for {n} files ({off[-1]/1e6:.1f} MB text): {byt/ms/1e6:.1f} GB/s, {n/ms*1e3:.0f} files/s")
want = np.loadtxt(io.BytesIO(files[5]), delimiter=",", dtype=np.float32)
print("exact:", np.array_equal(out[5].cpu().numpy(), want), st[:2].tolist())
