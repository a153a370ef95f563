"""CSV / stream ingestion.  Oracle = numpy's own reader (np.loadtxt is what the reference data
pipeline uses) on the same bytes; the fixture holds raw bytes of five of the reference's CSV windows."""
import io

import numpy as np
import pytest
import torch


def _savetxt_bytes(a):
    b = io.BytesIO()
    np.savetxt(b, a, delimiter=",", fmt="%.7f")          # the collector's writer (Neural_decoding_data_collector.py:139)
    return b.getvalue()


def test_csv_fixture_is_consistent(golden_dir):
    """CPU: the committed bytes parse (numpy) to the committed values, and equal the windows fixture."""
    f = np.load(golden_dir / "csv_bytes.npz")
    text, off = f["text"].tobytes(), f["offsets"]
    for i in range(len(off) - 1):
        got = np.loadtxt(io.BytesIO(text[off[i]:off[i + 1]]), delimiter=",", dtype=np.float32)
        assert got.shape == (625, 8) and np.array_equal(got, f["parsed"][i])
    w = np.load(golden_dir / "eeg_windows.npz")
    names = [str(n) for n in w["names"]]
    for i, n in enumerate(f["names"]):
        assert np.array_equal(w["X"][names.index(str(n))], f["parsed"][i])


def test_read_files_pinned_and_errors(tmp_path):
    """CPU: host-side packing of files into one buffer + offsets; argument errors."""
    from neural_speech_decoding_b200 import ingest
    a = tmp_path / "food_1.csv"; a.write_bytes(b"1.0,2.0\n")
    b = tmp_path / "water_2.csv"; b.write_bytes(b"-3.5,4.25\n7,8\n")
    buf, off = ingest.read_files_pinned([a, b])
    assert off.tolist() == [0, 8, 22] and bytes(buf.numpy()) == b"1.0,2.0\n-3.5,4.25\n7,8\n"
    with pytest.raises(ValueError):
        ingest.read_files_pinned([])
    with pytest.raises(RuntimeError):
        ingest.load_csv_windows([a], device="cpu")          # no CPU fallback
    with pytest.raises(ValueError):
        ingest.windows_from_stream(torch.zeros(3, 4, 5))


@pytest.mark.gpu
def test_csv_parse_matches_numpy_on_reference_bytes(golden_dir):
    from neural_speech_decoding_b200 import ingest
    f = np.load(golden_dir / "csv_bytes.npz")
    dev = torch.device("cuda:0")
    got = ingest.parse_csv_bytes(torch.from_numpy(f["text"].copy()).to(dev), torch.from_numpy(f["offsets"].copy()).to(dev))
    assert got.shape == (5, 625, 8)
    assert np.array_equal(got.cpu().numpy().view(np.uint32), f["parsed"].view(np.uint32))      # bit-exact


@pytest.mark.gpu
def test_csv_parse_bit_exact_on_written_files(tmp_path, windows):
    """All 324 windows re-written with the collector's writer, plus adversarial values (double-rounding candidates,
    large / tiny magnitudes, negative zero, integers, CR-LF, missing final newline), through the file API."""
    from neural_speech_decoding_b200 import ingest
    rng = np.random.default_rng(5)
    X = windows["X"].astype(np.float64)
    extra = np.stack([
        rng.standard_normal((625, 8)) * 1e4,                                     # 12 significant digits
        rng.standard_normal((625, 8)) * 1e-4,                                    # mostly leading zeros
        np.round(rng.standard_normal((625, 8)) * 50) + 0.5 ** 24 * rng.integers(-3, 4, (625, 8)),   # near fp32 ties
        np.where(rng.random((625, 8)) < 0.5, -0.0, rng.integers(-9, 10, (625, 8)).astype(np.float64)),
    ])
    allw = np.concatenate([X, extra])
    paths, want = [], []
    for i, a in enumerate(allw):
        raw = _savetxt_bytes(a)
        if i % 7 == 3:
            raw = raw.replace(b"\n", b"\r\n")
        if i % 11 == 5:
            raw = raw.rstrip(b"\r\n")
        p = tmp_path / f"food_{i:04d}.csv"
        p.write_bytes(raw)
        paths.append(p)
        want.append(np.loadtxt(io.BytesIO(raw), delimiter=",", dtype=np.float32))
    got = ingest.load_csv_windows(paths).cpu().numpy()
    want = np.stack(want)
    assert got.shape == want.shape
    assert np.array_equal(got, want)                                            # values (-0.0 == 0.0)
    nz = want != 0
    assert np.array_equal(got.view(np.uint32)[nz], want.view(np.uint32)[nz])     # bits


@pytest.mark.gpu
def test_csv_parse_rejects_malformed(tmp_path):
    from neural_speech_decoding_b200 import ingest
    good = _savetxt_bytes(np.zeros((625, 8)))
    for name, raw in (("short", good[: len(good) // 2]), ("nan", good.replace(b"0.0000000", b"nan", 1)),
                      ("exp", good.replace(b"0.0000000", b"1e-3", 1)), ("long", b"0.1234567890123456," + good)):
        p = tmp_path / f"food_{name}.csv"
        p.write_bytes(raw)
        with pytest.raises(ValueError):
            ingest.load_csv_windows([p])


@pytest.mark.gpu
def test_labelled_dir_and_stream_windows(tmp_path, windows):
    from neural_speech_decoding_b200 import ingest
    X = windows["X"]
    for i, pre in enumerate(["food", "water", "yes", "backgroundnoise", "food"]):
        (tmp_path / f"{pre}_{i}.csv").write_bytes(_savetxt_bytes(X[i]))
    Xd, y, names = ingest.load_labelled_dir(tmp_path, ["food", "water", "backgroundnoise"])
    assert Xd.shape == (4, 625, 8) and sorted(names) == names and "yes_2.csv" not in names
    assert y.tolist() == [{"food": 0, "water": 1, "backgroundnoise": 2}[n.split("_")[0]] for n in names]
    # continuous stream -> windows (streaming_process.py: non-overlapping; hop < window: overlapping)
    stream = torch.from_numpy(np.concatenate([X[0], X[1], X[2][:300]])).cuda()
    w = ingest.windows_from_stream(stream)
    assert w.shape == (2, 625, 8) and torch.equal(w[1].cpu(), torch.from_numpy(X[1]))
    w2 = ingest.windows_from_stream(stream, hop=125)
    s = stream.cpu().numpy()
    assert w2.shape[0] == (s.shape[0] - 625) // 125 + 1
    assert np.array_equal(w2[3].cpu().numpy(), s[375:1000])
