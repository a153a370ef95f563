"""In-tree build of libneuroalpha_b200.so (nvcc, sm_100a only).

No JIT cache: the shared object is written next to the sources so that it travels with the
repository snapshot to the GPU box (``*.so`` is git-ignored, not gpurun-ignored).

Every ``csrc/*.cu`` is compiled to its own object (in parallel, re-compiled only when the file, a
header or the flags changed) and the objects are linked into one shared library.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
OBJ_DIR = CSRC / "build"
LIB_PATH = PKG_DIR / "libneuroalpha_b200.so"
STAMP = PKG_DIR / "csrc" / ".build_stamp"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def sources():
    return sorted(CSRC.glob("*.cu"))


def _headers_digest() -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*.cuh")) + [PKG_DIR.parent / "include" / "neuroalpha.h"]):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _digest() -> str:
    h = hashlib.sha256(_headers_digest().encode())
    for p in sources():
        h.update(p.name.encode())
        h.update(p.read_bytes())
    return h.hexdigest()


def find_nvcc() -> str:
    cand = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(cand).exists():
        raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")
    return cand


def _compile_one(nvcc: str, src: Path, hdr: str, force: bool, verbose: bool) -> Path:
    obj = OBJ_DIR / (src.stem + ".o")
    stamp = OBJ_DIR / (src.stem + ".stamp")
    want = hashlib.sha256(hdr.encode() + src.read_bytes()).hexdigest()
    if not force and obj.exists() and stamp.exists() and stamp.read_text().strip() == want:
        return obj
    cmd = [nvcc, *NVCC_FLAGS, *(["-Xptxas", "-v"] if verbose else []), "-c", str(src), "-o", str(obj)]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    if verbose:
        print(proc.stderr)
    stamp.write_text(want)
    return obj


def build_library(force: bool = False, verbose: bool = False) -> Path:
    """Compile every .cu under csrc/ and link one shared object.  Rebuilds only what changed."""
    digest = _digest()
    if not force and LIB_PATH.exists() and STAMP.exists() and STAMP.read_text().strip() == digest:
        return LIB_PATH
    nvcc = find_nvcc()
    OBJ_DIR.mkdir(exist_ok=True)
    hdr = _headers_digest()
    with ThreadPoolExecutor(max_workers=max(1, min(8, os.cpu_count() or 1))) as ex:
        objs = list(ex.map(lambda s: _compile_one(nvcc, s, hdr, force, verbose), sources()))
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", *[str(o) for o in objs], "-o", str(LIB_PATH), "-lcuda"]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("link failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    STAMP.write_text(digest)
    return LIB_PATH


if __name__ == "__main__":
    import sys
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
