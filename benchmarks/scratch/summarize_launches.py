import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
h = rows[hi]; ki = h.index('Kernel Name'); vi = h.index('Metric Value')
d = collections.defaultdict(lambda: [0, 0.0, 0.0])
for r in rows[hi + 1:]:
    if len(r) <= vi: continue
    try: v = float(r[vi].replace(',', ''))
    except Exception: continue
    e = d[r[ki][:78]]; e[0] += 1; e[1] += v; e[2] = max(e[2], v)
tot = sum(v[1] for v in d.values())
for k, v in sorted(d.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print(f"{k:80s} {v[0]:5d} tot {v[1]/1e6:9.3f} ms {100*v[1]/tot:5.1f}%  avg {v[1]/v[0]/1e3:8.1f} us")
print("total", tot / 1e6, "ms")
