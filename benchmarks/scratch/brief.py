"""One-line summary of bench.py JSON lines: python benchmarks/scratch/brief.py file.json [...]"""
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable:", e); continue
    r = d.get("roofline", {})
    print(f, "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "ms", round(d["ms_per_step"], 3), "k2_ms", round(r.get("kernel_ms", 0), 3),
          "burst", round(r.get("frac_burst", 0), 3), "sust", round(r.get("frac_sustained", 0), 3), "k3", round(d.get("roofline_decode", {}).get("kernel_ms", 0), 4),
          "k3frac", round(d.get("roofline_decode", {}).get("frac", 0), 3), "clk", d["clocks"]["sm_mhz"], "fused", d.get("roofline_decode", {}).get("note", "")[:0],
          "train", [(t["shape"][:6], round(t["ms_per_step"], 2), round(t.get("ms_per_step_median", 0), 2)) for t in (d.get("train") or [])])
