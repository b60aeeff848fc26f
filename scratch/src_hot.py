# usage: src_hot.py report.ncu-rep [kernel-substring] [topN]
import csv, subprocess, sys, io
rep=sys.argv[1]; sub=sys.argv[2] if len(sys.argv)>2 else ''; topn=int(sys.argv[3]) if len(sys.argv)>3 else 25
out=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','cuda,sass'],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(out)))
cur_file=None; hdr=None; kern=None; agg={}; order=[]
for r in rows:
    if not r: continue
    if r[0]=='File Path': cur_file=r[1].split('/')[-1]; continue
    if r[0]=='Function Name':
        kern=r[1][:60]; 
        if kern not in agg: agg[kern]={}; order.append(kern)
        continue
    if r[0]=='Line No': hdr=r; continue
    if hdr is None or len(r)<10 or not r[0].isdigit(): continue
    def col(n):
        v=r[hdr.index(n)]; return int(v) if v.isdigit() else 0
    a=agg[kern].setdefault((cur_file,int(r[0])),[r[1],0,0,{}])
    a[1]+=col('Instructions Executed'); a[2]+=col('# Samples')
    for n in hdr:
        if n.startswith('stall_') and 'Not Issued' not in n:
            v=r[hdr.index(n)]
            if v.isdigit() and int(v): a[3][n]=a[3].get(n,0)+int(v)
for k in order:
    if sub not in k: continue
    d=agg[k]; tot=sum(a[1] for a in d.values()) or 1; ts=sum(a[2] for a in d.values()) or 1
    print('=====',k,'inst',tot,'samples',ts)
    for key,a in sorted(d.items(), key=lambda kv:-kv[1][2])[:topn]:
        top=sorted(a[3].items(), key=lambda kv:-kv[1])[:2]
        print(f"{key[0][:10]:10s}:{key[1]:4d} samp={a[2]/ts*100:5.1f}% inst={a[1]/tot*100:5.1f}% {','.join(f'{n[6:]}={v}' for n,v in top):28s} {a[0].strip()[:80]}")
