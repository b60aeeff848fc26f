import csv, sys, subprocess, io
rep = sys.argv[1]
out = subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","cuda,sass"],capture_output=True,text=True).stdout
rows=[]; cur=None; hdr=None
for r in csv.reader(io.StringIO(out)):
    if not r: continue
    if r[0]=="File Path": cur=r[1]; continue
    if r[0]=="Function Name": continue
    if r[0]=="Line No": hdr=r; continue
    if r[0]=="Kernel Name" or r[0]=="Address": hdr=None; continue
    if hdr and r[0].isdigit():
        ii=hdr.index("Instructions Executed"); si=hdr.index("# Samples")
        if len(r)!=len(hdr) or not r[ii].isdigit(): continue
        rows.append((int(r[ii]), int(r[si]), cur.split("/")[-1], int(r[0]), r[1].strip()[:110]))
tot=sum(x[0] for x in rows); ts=sum(x[1] for x in rows)
print("total inst", tot, "samples", ts)
rows.sort(reverse=True)
for n,s,f,l,src in rows[:int(sys.argv[2]) if len(sys.argv)>2 else 60]:
    print(f"{100*n/tot:5.1f}% {100*s/ts:5.1f}%s {f}:{l}  {src}")
