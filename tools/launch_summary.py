import csv, collections, sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10 and r[0].isdigit()]
agg=collections.OrderedDict()
for r in rows:
    name=r[4].split('(')[0].replace('vpl::','').replace('<unnamed>::','').replace('void ','')
    agg.setdefault(name,[]).append(float(r[-1]))
tot=sum(sum(v) for v in agg.values())
print(f"{'kernel':32s} {'n':>4s} {'mean us':>10s} {'total ms':>9s} {'share':>6s}")
for k,v in agg.items():
    print(f"{k:32s} {len(v):4d} {sum(v)/len(v)/1e3:10.1f} {sum(v)/1e6:9.2f} {sum(v)/tot*100:5.1f}%")
