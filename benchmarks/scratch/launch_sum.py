import csv,collections,sys
for f in sys.argv[1:]:
    lines=[l for l in open(f) if l.startswith('"')]
    r=csv.DictReader(lines)
    agg=collections.defaultdict(lambda:[0,0.0])
    for row in r:
        if row.get("Metric Name")!="gpu__time_duration.sum": continue
        k=row["Kernel Name"][:70]
        v=float(row["Metric Value"].replace(",","")); u=row["Metric Unit"]
        if u=="ns": v/=1e3
        elif u=="ms": v*=1e3
        agg[k][0]+=1; agg[k][1]+=v
    tot=sum(v[1] for v in agg.values())
    print("==",f,"total us",round(tot))
    for k,v in sorted(agg.items(),key=lambda x:-x[1][1])[:12]:
        print(f"  {v[1]:10.0f} us  n={v[0]:4d} avg {v[1]/v[0]:8.1f} {k}")
