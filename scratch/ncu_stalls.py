import csv, sys, subprocess, io
rep = sys.argv[1]; col = sys.argv[2] if len(sys.argv)>2 else "stall_long_sb"
out = subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","cuda,sass"],capture_output=True,text=True).stdout
rows=[]; cur=None; hdr=None
for r in csv.reader(io.StringIO(out)):
    if not r: continue
    if r[0]=="File Path": cur=r[1]; continue
    if r[0]=="Function Name": continue
    if r[0]=="Line No": hdr=r; continue
    if r[0]=="Kernel Name" or r[0]=="Address": hdr=None; continue
    if hdr and r[0].isdigit() and len(r)==len(hdr):
        ci=hdr.index(col); si=hdr.index("# Samples"); ii=hdr.index("Instructions Executed")
        if r[ci].isdigit():
            rows.append((int(r[ci]), int(r[si]), int(r[ii]), cur.split("/")[-1], int(r[0]), r[1].strip()[:100]))
tot=sum(x[0] for x in rows); ts=sum(x[1] for x in rows)
print("total", col, tot, "of samples", ts)
rows.sort(reverse=True)
for n,s,i,f,l,src in rows[:int(sys.argv[3]) if len(sys.argv)>3 else 25]:
    print(f"{100*n/max(tot,1):5.1f}%  smp {s:5d} inst {i:8d} {f}:{l}  {src}")
