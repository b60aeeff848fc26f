"""Turns the raw output of tools/collect_evidence.sh (gpurun_out/r2ev_*) into the tracked summaries under profiles/:
  r2_config_lines.jsonl   one bench line per configuration / variant (as printed by bench.py, plus a "name" key)
  r2_launch_summary.txt   per-kernel means of the ncu launch lists (cold, serialised launches: shares, not bench values)
  r2_ncu_digests.txt      the `--set full` metrics quoted in DESIGN.md for every captured kernel
  r2_aux_timings.txt      secondary entry points next to the CPU oracle
  dense_traffic.json      dram bytes per launch of the roofline kernels (bench.py's `roofline.traffic`)
"""
import csv, glob, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
PREFIX = sys.argv[1] if len(sys.argv) > 1 else "r2ev"

lines = []
for f in sorted(glob.glob(os.path.join(G, PREFIX + "_line_*.json"))):
    txt = open(f).read().strip().splitlines()
    if not txt:
        continue
    d = json.loads(txt[-1])
    lines.append({"name": os.path.basename(f)[len(PREFIX) + 6:-5], **d})
with open(os.path.join(P, "r2_config_lines.jsonl"), "w") as f:
    for d in lines:
        f.write(json.dumps(d) + "\n")

with open(os.path.join(P, "r2_launch_summary.txt"), "w") as out:
    out.write("# ncu --metrics gpu__time_duration.sum --clock-control none, python bench.py ... (tools/collect_evidence.sh)\n"
              "# per-launch times are cold-cache and serialised: they give each kernel's SHARE of a step, not bench values\n")
    for name, what in (("default", "cityscapes_1024x2048_b8_n100, --ring 1 (one pipeline, steps not overlapped)"),
                       ("coco", "coco_800x1333_c80_n1000 (isg_mask_nms)"), ("kmeans", "crowd_1024x2048_b4_n500_kmeans, --ring 1"),
                       ("split", "cityscapes_1024x2048_b8_n100, --ring 1, ISG_SPLIT_KEEP=1 (isg_topk_keep + labels-only dense kernel)")):
        path = os.path.join(G, "%s_launches_%s.csv" % (PREFIX, name))
        if not os.path.exists(path):
            continue
        res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_launch_summary.py"), path], capture_output=True, text=True).stdout
        out.write("\n## %s\n%s" % (what, res))

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'lts__t_sector_hit_rate.pct',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
traffic = {}
try:
    traffic = json.load(open(os.path.join(P, "dense_traffic.json")))
except Exception:
    pass
with open(os.path.join(P, "r2_ncu_digests.txt"), "w") as out:
    out.write("# ncu --set full --clock-control none --import-source on -k regex:<kernel> -c 1 (tools/collect_evidence.sh); one launch each\n")
    for rep in sorted(glob.glob(os.path.join(G, PREFIX + "_*.ncu-rep"))):
        name = os.path.basename(rep)[len(PREFIX) + 1:-8]
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        if len(rows) < 3:
            continue
        h, u, v = rows[0], rows[1], rows[2]
        kn = v[h.index("Kernel Name")] if "Kernel Name" in h else "?"
        out.write("\n## %s  (%s)\n" % (name, kn[:90]))
        vals = {}
        for i, k in enumerate(h):
            if k in KEYS or ("issue_stalled" in k and "per_issue_active" in k and v[i] not in ("0", "0.00", "")):
                out.write("%-92s %-14s %s\n" % (k, u[i], v[i]))
            vals[k] = v[i]
        try:
            mul = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
            rd = float(vals["dram__bytes_read.sum"].replace(",", "")) * mul[u[h.index("dram__bytes_read.sum")]]
            wr = float(vals["dram__bytes_write.sum"].replace(",", "")) * mul[u[h.index("dram__bytes_write.sum")]]
            traffic["ncu:" + name] = int(rd + wr)
        except Exception:
            pass
if "ncu:dense" in traffic and traffic["ncu:dense"]:
    traffic["cityscapes_1024x2048_b8_n100"] = traffic["ncu:dense"]
if traffic.get("ncu:dense_wide"):
    traffic["cityscapes_1024x2048_b8_n100:wide"] = traffic["ncu:dense_wide"]
if all(traffic.get("ncu:" + k) for k in ("mask_area", "mask_pair", "nms_scan")):
    traffic["coco_800x1333_c80_n1000"] = sum(traffic["ncu:" + k] for k in ("mask_area", "mask_pair", "nms_scan"))
json.dump(traffic, open(os.path.join(P, "dense_traffic.json"), "w"), indent=1)
for src, dst, head in (("_fill_polygons.txt", "r2_fill_polygons.txt", "# python tools/bench_fill.py on a B200\n"),
                       ("_overlap.txt", "r2_overlap_ablation.txt", "# python tools/overlap_experiment.py on a B200 (ring = independent pipelines used round-robin; spare = SMs the dense\n# kernel leaves free, ISG_DENSE_SPARE; 200 steps per point, best of 2)\n")):
    path = os.path.join(G, PREFIX + src)
    if os.path.exists(path):
        open(os.path.join(P, dst), "w").write(head + open(path).read())
aux = os.path.join(G, PREFIX + "_aux_timings.txt")
if os.path.exists(aux):
    open(os.path.join(P, "r2_aux_timings.txt"), "w").write("# python tests/aux_timings.py on a B200 (device time through the drop-in call incl. its read-back; CPU oracle on the box's host)\n" + open(aux).read())
for d in lines:
    r, e, c = d.get("roofline") or {}, d.get("e2e") or {}, d.get("cpu_baseline") or {}
    print("%-14s value %10.0f  ms/step %.4f  roof %.3f (kern %.4f ms)  e2e %8.0f  cpu %6.2f (%s)" % (
        d["name"], d["value"], d["ms_per_step"], r.get("frac") or 0, r.get("kernel_ms") or 0, e.get("value") or 0, c.get("value") or 0, c.get("kind")))
