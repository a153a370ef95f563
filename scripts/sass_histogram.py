"""Per-kernel SASS opcode histogram of libneuroalpha_b200.so (cuobjdump -sass): the instructions that prove the
tcgen05 / TMEM / TMA / MUFU paths are really in the binary.  Usage: python scripts/sass_histogram.py > profiles/rN_sass_histogram.txt"""
import collections
import re
import subprocess
import sys
from pathlib import Path

LIB = Path(__file__).resolve().parents[1] / "neural-speech-decoding_b200" / "libneuroalpha_b200.so"
WATCH = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTCCP", "UTCATOM", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "LDGSTS",
         "MUFU.EX2", "MUFU.RCP", "MUFU.TANH", "MUFU.RSQ", "MUFU.LG2", "FFMA", "FMUL", "FADD", "DFMA", "HFMA2", "LDS", "STS", "LDG", "STG",
         "SHFL", "BAR", "ELECT", "R2UR"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            kernels[cur]["_total"] += 1
            for w in WATCH:
                if op == w or op.startswith(w + "."):
                    kernels[cur][w] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print(f"# SASS opcode histogram of {LIB.name} (sm_100a), one row per kernel; columns = static instruction counts")
    print("kernel," + ",".join(["total"] + WATCH))
    for (k, c), name in zip(kernels.items(), demangle):
        short = re.sub(r"\(.*", "", name).replace("na::tc::", "").replace("na::", "").replace("void ", "")
        print(short + "," + ",".join(str(c.get(w, 0)) for w in ["_total"] + WATCH))


if __name__ == "__main__":
    sys.exit(main())
