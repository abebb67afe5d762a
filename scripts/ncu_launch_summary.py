import csv, sys, collections, re
rows=[]
with open(sys.argv[1]) as f:
    lines=[l for l in f if not l.startswith("==")]
r=csv.DictReader(lines)
agg=collections.defaultdict(lambda:[0.0,0])
for row in r:
    name=row.get("Kernel Name") or ""
    try: v=float(row["Metric Value"].replace(",",""))
    except: continue
    unit=row.get("Metric Unit","")
    if unit in ("nsecond","ns"): v/=1e3
    elif unit in ("msecond","ms"): v*=1e3
    elif unit in ("second","s"): v*=1e6
    name=re.sub(r"\(.*","",name)
    agg[name][0]+=v; agg[name][1]+=1
tot=sum(v[0] for v in agg.values())
print("total us",round(tot,1))
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][0])[:30]:
    print(f"| `{k}` | {v[0]/1e3:.3f} | {v[1]} | {v[0]/v[1]:.1f} | {100*v[0]/tot:.1f} % |")
