"""SASS digest of libisg.so: per kernel, the instruction counts that show what the hardware is asked to do —
TMA (UTMALDG = tensor-map tile loads, UBLKCP = 1-D bulk copies), mbarrier traffic (SYNCS), cluster / DSMEM use,
MUFU, FP64, global atomics; plus proof that no tensor-core instruction exists (nothing on this path is a contraction).
  python tools/sass_digest.py > profiles/r2_sass_digest.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "instance-segmentation_b200", "libisg.so")
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
cur, stats = None, collections.OrderedDict()
PAT = [("UTMALDG", r"\bUTMALDG"), ("UBLKCP", r"\bUBLKCP"), ("SYNCS", r"\bSYNCS"), ("UCGABAR/cluster", r"\bUCGABAR|\bCGAERRBAR"),
       ("LD/ST.shared::cluster", r"\b(LD|ST)\.E?\.?.*\.SHARED::CLUSTER|\bLDS\..*CLUSTER|MAPA"), ("MUFU", r"\bMUFU"), ("F64", r"\bD(ADD|MUL|FMA|SETP)"),
       ("ATOMG/RED", r"\b(ATOMG|RED)\b|\bREDG|\bATOM\b"), ("MATCH/VOTE/SHFL", r"\b(MATCH|VOTE|SHFL)"), ("tensor (HMMA/UTCMMA/…)", r"\b(HMMA|IMMA|DMMA|UTC\w*MMA|QGMMA|HGMMA)")]
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        stats[cur] = collections.Counter()
        continue
    if cur and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
        stats[cur]["instr"] += 1
        for name, pat in PAT:
            if re.search(pat, line):
                stats[cur][name] += 1


def demangle(n):
    try:
        d = subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip()
        d = d.split(">(")[0] + ">" if ">(" in d else d.split("(")[0]
        return d.replace("(int)", "").replace("(bool)", "").replace("void ", "") or n
    except Exception:
        return n


print("# cuobjdump -sass %s   (architectures: %s)" % (os.path.relpath(LIB, ROOT), ", ".join(arch)))
cols = ["instr"] + [p[0] for p in PAT]
print("%-64s " % "kernel" + " ".join("%9s" % c[:9] for c in cols))
tot = collections.Counter()
for k, c in stats.items():
    name = demangle(k).replace("isg::", "")
    print("%-64s " % name[:64] + " ".join("%9d" % c[x] for x in cols))
    tot.update(c)
print("%-64s " % "TOTAL" + " ".join("%9d" % tot[x] for x in cols))
print("# tensor-core instructions in the library: %d (no kernel on this path is a dense contraction)" % tot["tensor (HMMA/UTCMMA/…)"])
