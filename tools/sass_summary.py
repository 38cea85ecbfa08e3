"""SASS evidence for the Blackwell paths: per kernel of vamb_b200/_vk.so, the counts of the mnemonics that prove
TMA (UTMALDG), tcgen05.mma (UTCHMMA) / commit (UTCBAR), tensor-memory stores / loads (STTM / LDTM), mbarrier traffic
(SYNCS), cluster barriers (UCGABAR_*), cp.async (LDGSTS), and the fences / atomics of the completion protocol.
    python tools/sass_summary.py > profiles/r02_sass_summary.txt          (CPU box: cuobjdump only)
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "vamb_b200", "_vk.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
KEEP = ("UTMALDG", "UTCHMMA", "UTCBAR", "STTM", "LDTM", "SYNCS", "UCGABAR_ARV", "UCGABAR_WAIT", "LDGSTS", "UTCATOMSWS",
        "MEMBAR", "ERRBAR", "ATOMG", "ATOMS", "RED", "FFMA", "DFMA", "DADD", "MAPA")
pat = re.compile(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z_0-9]+)")
cnt, regs, cur = collections.defaultdict(collections.Counter), {}, None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = pat.match(line)
    if cur and m:
        op = m.group(1)
        cnt[cur]["_all"] += 1
        if op in KEEP:
            cnt[cur][op] += 1
names = subprocess.run(["c++filt"], input="\n".join(cnt), capture_output=True, text=True).stdout.splitlines()
print(f"# cuobjdump -sass {os.path.relpath(so, ROOT)} (sm_100a): instruction counts per kernel")
rows = []
for mangled, name in zip(cnt, names):
    short = name.split("(")[0].replace("void ", "").replace("(anonymous namespace)::", "")
    rows.append((short, cnt[mangled]))
for short, c in sorted(rows):
    marks = "  ".join(f"{k}={c[k]}" for k in KEEP if c[k])
    print(f"{short:45s} total={c['_all']:6d}  {marks}")
