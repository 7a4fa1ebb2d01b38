"""SASS opcode histogram per kernel of libequss_b200.so (cuobjdump -sass): instruction count and the mnemonics that
prove which hardware path a kernel uses -- UTCHMMA/UTCQMMA (tcgen05.mma), LDTM/STTM (tcgen05.ld/st), UTMALDG (TMA tensor
load), UBLKCP (bulk copy), UTCBAR / SYNCS (mbarrier traffic), FMNMX3 / VIMNMX3 (3-input min/max), ATOMS / ATOMG / RED
(atomics), MUFU, HMMA (legacy mma.sync -- must be absent).  Usage: python scripts/sass_histogram.py > profiles/r2_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "expand-and-quantize-for-unsupervised-semantic-segmentation_b200", "libequss_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEY = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "FMNMX3", "FMNMX", "VIMNMX3",
       "HMMA", "IMMA", "ATOMS", "ATOMG", "RED", "ATOM", "MUFU", "FFMA", "LDG", "STG", "LDS", "STS", "SHFL", "MATCH", "DADD", "F2FP", "BAR"]
kern = None
hist = {}
for ln in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        kern = m.group(1)
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
    if m and kern:
        op = m.group(1)
        hist[kern]["_total"] += 1
        hist[kern][op.split(".")[0]] += 1
dem = subprocess.run(["cu++filt"] + list(hist.keys()), capture_output=True, text=True).stdout.splitlines() if hist else []
names = dict(zip(hist.keys(), dem)) if len(dem) == len(hist) else {k: k for k in hist}
print(f"# SASS opcode histogram of {os.path.basename(lib)} (sm_100a), {len(hist)} kernels; columns: total instructions, then non-zero counts of the key mnemonics")
for k in sorted(hist, key=lambda k: names[k]):
    h = hist[k]
    nm = re.sub(r"\(.*", "", names[k])
    parts = [f"{op}={h[op]}" for op in KEY if h.get(op)]
    print(f"{nm[:110]:110s} total={h['_total']:6d}  " + " ".join(parts))
legacy = [names[k] for k in hist if hist[k].get("HMMA") or hist[k].get("IMMA")]
print(f"# kernels with legacy mma.sync (HMMA/IMMA): {len(legacy)}")
tc = sorted({re.sub(r'<.*', '', re.sub(r'\(.*', '', names[k])) for k in hist if hist[k].get('UTCHMMA') or hist[k].get('UTCQMMA')})
print(f"# kernels issuing tcgen05.mma: {', '.join(tc)}")
