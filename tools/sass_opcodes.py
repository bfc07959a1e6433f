"""Per-kernel counts of the SASS mnemonics that prove the tensor-core / TMA paths (cuobjdump -sass of the in-tree
library): DMMA (FP64 tensor), UTCHMMA / UTCQMMA... (tcgen05.mma), UTMALDG / UTMASTG (TMA), LDTM / STTM (tcgen05.ld/st),
UBLKCP (bulk copies), SYNCS (mbarrier), MULTIMEM / REDG.  Usage: python tools/sass_opcodes.py > profiles/rNN_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "scimloperators.jl_b200", "libscimlop_b200.so")
KEYS = ["DMMA", "UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCOMMA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "SYNCS",
        "MULTIMEM", "HMMA", "FFMA", "DFMA", "LDGSTS", "STAS", "UTCCP"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)(\.[A-Z0-9_.]+)?", line)
        if m:
            op = m.group(1)
            kernels[cur]["_total"] += 1
            for k in KEYS:
                if op.startswith(k):
                    kernels[cur][k + (m.group(2) or "") if k in ("DMMA", "UTMALDG", "UTMASTG") else k] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)} : instructions per kernel (sm_100a)")
    for (name, cnt), dn in zip(kernels.items(), demangle):
        short = re.sub(r"\(.*", "", dn.replace("(anonymous namespace)::", ""))
        hits = ", ".join(f"{k} x{v}" for k, v in sorted(cnt.items()) if k != "_total")
        print(f"{short}\n    total {cnt['_total']}; {hits or '-'}")


if __name__ == "__main__":
    sys.exit(main())
