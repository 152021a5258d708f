"""Per-source-line totals of an .ncu-rep captured with --import-source on:  python ncu_lines.py rep [column] [topN]"""
import csv, collections, subprocess, sys
rep = sys.argv[1]
col = sys.argv[2] if len(sys.argv) > 2 else "Instructions Executed"
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; cur = None
agg = collections.defaultdict(lambda: [0.0, ""])
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) > 2 and r[0] == "Line No": hdr = r; continue
    if len(r) > 2 and hdr:
        d = dict(zip(hdr, r))
        try: ln = int(r[0]); v = float(d.get(col) or 0)
        except ValueError: continue
        e = agg[(cur, ln)]; e[0] += v; e[1] = r[1]
tot = sum(e[0] for e in agg.values())
print(col, "total", tot)
for (f, ln), e in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    print(f"{f[:18]:18s} L{ln:4d} {e[0]:14.0f} {100*e[0]/max(tot,1):5.1f}%  {e[1][:120]}")
