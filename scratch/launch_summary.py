import csv, collections, sys
path=sys.argv[1]; per_step=int(sys.argv[2]) if len(sys.argv)>2 else None
with open(path) as f:
    lines=[l for l in f if not l.startswith('==')]
agg=collections.OrderedDict()
for row in csv.DictReader(lines):
    name=row['Kernel Name'].split('(')[0].replace('void ','').replace('isg::','')
    agg.setdefault(name,[]).append(float(row['Metric Value'].replace(',','')))
tot=sum(sum(v) for v in agg.values())
print(f"{'kernel':44s} {'n':>3s} {'mean us':>9s} {'share':>7s}")
for k,v in agg.items():
    print(f"{k[:44]:44s} {len(v):3d} {sum(v)/len(v)/1000:9.2f} {100*sum(v)/tot:6.1f}%")
n=max(len(v) for v in agg.values())
print('sum of means (one step, serialised, cold):', round(sum(sum(v)/len(v) for v in agg.values())/1000,1),'us')
