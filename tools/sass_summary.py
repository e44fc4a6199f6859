#!/usr/bin/env python
"""Opcode histogram per kernel of ppo_cpp_b200/libppo_core.so (cuobjdump -sass) -> profiles/sass_summary.txt.
Shows which kernels carry tcgen05 (UTC*MMA), TMEM loads/stores (LDTM/STTM), bulk copies (UBLKCP), tensor-map TMA
(UTMALDG/UTMASTG) and which run on the CUDA cores only."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "ppo_cpp_b200", "libppo_core.so")
KEY = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "HMMA", "FFMA", "DFMA", "MUFU", "LDG", "STG",
       "LDS", "STS", "LDGSTS", "ATOM", "ATOMG", "RED", "BAR", "MEMBAR", "SHFL", "F2FP", "R2UR"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", SO], check=True, capture_output=True, text=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            kernels[cur][m.group(1).split(".")[0]] += 1
    lines = ["# cuobjdump -sass opcode counts per kernel (sm_100a cubin of libppo_core.so); columns: " + " ".join(KEY), ""]
    for k, c in kernels.items():
        tot = sum(c.values())
        cells = " ".join(f"{n}={c[n]}" for n in KEY if c[n])
        lines.append(f"{k}\n    instructions={tot}  {cells}")
    txt = "\n".join(lines) + "\n"
    dst = os.path.join(ROOT, "profiles", sys.argv[1] if len(sys.argv) > 1 else "sass_summary.txt")
    with open(dst, "w") as f:
        f.write(txt)
    print(txt[:3000])


if __name__ == "__main__":
    main()
