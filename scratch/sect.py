import csv, subprocess, io, sys
rep=sys.argv[1]; npx=float(sys.argv[2]) if len(sys.argv)>2 else 8*1024*2048
out=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','cuda,sass'],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(out)))
cur_file=None; hdr=None; agg={}; samp={}; ops={}
for r in rows:
    if not r: continue
    if r[0]=='File Path': cur_file=r[1].split('/')[-1]; continue
    if r[0]=='Function Name': continue
    if r[0]=='Line No': hdr=r; continue
    if hdr is None or len(r)<10: continue
    ie=hdr.index('Instructions Executed'); isp=hdr.index('# Samples')
    if r[0].isdigit():
        cur=(cur_file,int(r[0]),r[1].strip()[:90]); agg.setdefault(cur,0); samp.setdefault(cur,0)
        if r[isp].isdigit(): samp[cur]+=int(r[isp])
    elif r[0]=='' and r[ie].isdigit():
        agg[cur]+=int(r[ie]); op=r[3].strip().split()[0] if r[3].strip() else ''
        if op.startswith('@'): op=r[3].strip().split()[1]
        ops[op]=ops.get(op,0)+int(r[ie])
tot=sum(agg.values()); ts=sum(samp.values()) or 1
print('total warp instr', tot, ' thread-slots per pixel', round(tot*32/npx,1))
for k,v in sorted(agg.items(), key=lambda kv:-kv[1])[:int(sys.argv[3]) if len(sys.argv)>3 else 30]:
    print(f'{k[0][:14]:14s}:{k[1]:4d} inst={v/tot*100:5.1f}% ({v*32/npx:5.1f}/px) samp={samp[k]/ts*100:5.1f}%  {k[2]}')
print(sorted(ops.items(), key=lambda kv:-kv[1])[:30])
