#!/usr/bin/env python
"""SASS opcode histogram per kernel of libb2rl.so (cuobjdump -sass; runs without a GPU): the mnemonics that prove which
hardware paths a kernel uses — UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld), UTMALDG (TMA tensor loads; .MULTICAST),
UTCBAR (tcgen05.commit), FFMA / FFMA2 (fp32, packed fp32x2), STAS (st.async), SYNCS (mbarrier), LDGSTS (cp.async),
UCGABAR (cluster barrier), ACQBULK / griddepcontrol (PDL).   python tools/sass_histogram.py > profiles/r2_sass_histogram.txt"""
import collections
import re
import subprocess
import sys
from pathlib import Path

LIB = Path(__file__).resolve().parent.parent / "sac_td3_cudagraphs_pytorch_b200" / "libb2rl.so"
WATCH = ["UTCHMMA", "LDTM", "UTMALDG", "UTCBAR", "UTMASTG", "FFMA2", "FFMA", "HMMA", "STAS", "SYNCS", "LDGSTS", "UCGABAR", "ATOM", "RED",
         "LDG", "STG", "LDS", "STS", "SHFL", "MUFU", "BAR"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = kernels.setdefault(re.sub(r"\(.*", "", name), collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", line)
        if m and cur is not None:
            op, mods = m.group(1), m.group(2)
            cur["total"] += 1
            if op in WATCH:
                cur[op] += 1
            if op == "UTMALDG" and ".MULTICAST" in mods:
                cur["UTMALDG.MULTICAST"] += 1
    cols = ["total"] + WATCH[:8] + ["UTMALDG.MULTICAST"] + WATCH[8:]
    print(f"{'kernel':58s} " + " ".join(f"{c[:9]:>9s}" for c in cols))
    for k, c in kernels.items():
        if "b2rl" in k:
            print(f"{k.replace('b2rl::', '')[:58]:58s} " + " ".join(f"{c.get(col, 0):9d}" for col in cols))


if __name__ == "__main__":
    main()
