import csv,collections,sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=None; d=collections.OrderedDict()
for r in rows:
    if len(r)>5 and r[0]=='ID': hdr=r; continue
    if hdr and len(r)==len(hdr):
        rec=dict(zip(hdr,r))
        if rec.get('Metric Name')=='gpu__time_duration.sum':
            name=rec['Kernel Name'].split('(')[0].replace('void ','')
            v=float(rec['Metric Value']); u=rec['Metric Unit']
            if u=='ns': v/=1000
            if u=='ms': v*=1000
            d.setdefault(name,[]).append(v)
tot=sum(sum(v)/len(v) for v in d.values())
print("%-46s %4s %9s %7s" % ("kernel","n","mean us","share"))
for k,v in d.items():
    m=sum(v)/len(v); print("%-46s %4d %9.2f %6.1f%%" % (k[:46],len(v),m,100*m/tot))
print("sum of means (one step, serialised, cold): %.1f us" % tot)
