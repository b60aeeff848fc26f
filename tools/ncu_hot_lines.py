import csv, sys, subprocess, io
rep = sys.argv[1]; kern = sys.argv[2]; topn=int(sys.argv[3]) if len(sys.argv)>3 else 25
out = subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","cuda,sass","-k","regex:"+kern],capture_output=True,text=True).stdout
rows=[]; cur=None; hdr=None
for r in csv.reader(io.StringIO(out)):
    if not r: continue
    if r[0]=="File Path": cur=r[1]; continue
    if r[0]=="Function Name": continue
    if r[0]=="Line No": hdr=r; continue
    if r[0]=="Kernel Name" or r[0]=="Address": hdr=None; continue
    if hdr and r[0].isdigit() and len(r)==len(hdr):
        si=hdr.index("# Samples"); ii=hdr.index("Instructions Executed")
        if r[si].isdigit():
            rows.append((int(r[si]), int(r[ii]), cur.split("/")[-1], int(r[0]), r[1].strip()[:105]))
ts=sum(x[0] for x in rows); ti=sum(x[1] for x in rows)
print("samples", ts, "inst", ti)
rows.sort(reverse=True)
for s,i,f,l,src in rows[:topn]:
    print(f"{100*s/max(ts,1):5.1f}%s {100*i/max(ti,1):5.1f}%i {f}:{l}  {src}")
