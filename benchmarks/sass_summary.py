"""SASS evidence for profiles/: per-kernel counts of the sm_100a tensor-core / TMA / TMEM mnemonics in the built library.

    python benchmarks/sass_summary.py > profiles/r02_sass_summary.txt

UTCHMMA = tcgen05.mma kind::f16/tf32 (.2CTA = cta_group::2), UTMALDG / UTMASTG = TMA tensor load / store, LDTM / STTM =
tcgen05.ld / st, UTCBAR = tcgen05.commit (.2CTA.MULTICAST = commit to both CTAs of a pair), UCGABAR = cluster barrier.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "peneo_b200", "libpeneo_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
KEYS = ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "UCGABAR", "MUFU.TANH", "HFMA2.BF16", "HMMA", "WGMMA")
pat = re.compile(r"^\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)")
fn, per = None, collections.defaultdict(collections.Counter)
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        fn = m.group(1)
        continue
    m = pat.match(line)
    if m and fn:
        op = m.group(2)
        for k in KEYS:
            if op.startswith(k):
                per[fn][op if k.startswith(("UTCHMMA", "UTMALDG", "UTMASTG", "UTCBAR")) else k] += 1
tot = collections.Counter()
rows = []
for f, c in per.items():
    tot.update(c)
    if any(k.startswith(("UTCHMMA", "UTMALDG", "LDTM")) for k in c):
        rows.append((subprocess.run(["c++filt", f], capture_output=True, text=True).stdout.strip().split("(")[0], c))
print("library:", os.path.relpath(lib, ROOT), os.path.getsize(lib), "bytes;  arch sm_100a;  cuobjdump -sass mnemonic counts")
print("whole library:", ", ".join(f"{k} {v}" for k, v in sorted(tot.items())))
print("(no HMMA / WGMMA: every tensor-core instruction is tcgen05)" if not (tot["HMMA"] or tot["WGMMA"]) else "")
print()
for name, c in sorted(rows):
    print(f"{name}\n    " + ", ".join(f"{k} {v}" for k, v in sorted(c.items())))
