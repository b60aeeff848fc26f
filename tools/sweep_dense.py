"""GPU sweep of the dense kernel geometries: time + equality against the first configuration (v1 = the plain-LDG kernel)."""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))   # repo root (this file lives in tools/)
sys.path.insert(0, ROOT)
import bench
import isg_b200  # noqa
from isg_b200 import _lib, engine

def main():
    wlname = sys.argv[1] if len(sys.argv) > 1 else "cityscapes_1024x2048_b8_n100"
    cfgs = sys.argv[2].split(",") if len(sys.argv) > 2 else ["v1", "2x8x2", "4x4x3", "4x4x4", "4x2x6", "2x4x4", "2x4x6", "2x8x3"]
    wl = bench.WORKLOADS[wlname]
    dev = torch.device("cuda", 0)
    B, H, W, N = wl["B"], wl["H"], wl["W"], wl["N"]
    host = bench.make_batch(wl, 0)
    d = {k: v.to(dev) for k, v in host.items()}
    A, C = d["classification"].shape[1], d["classification"].shape[2]
    max_keep = max(64, 1 << int(np.ceil(np.log2(N * 1.3))))
    bplan = engine.BoxPlan(B, A, C, H, W, dev, cap=1024, max_keep=max_keep)
    dplan = engine.DecodePlan(B, H, W, bplan.N, wl["kp_th"], dev, "dense", want_score=False, wh_delta=0.1)
    pipe = engine.DecodePipeline(bplan, dplan)
    ref = None
    for cfg in cfgs:
        os.environ.pop("ISG_DENSE_V1", None); os.environ.pop("ISG_DENSE_CFG", None)
        os.environ.pop("ISG_DENSE_TAIL", None); os.environ.pop("ISG_DENSE_SPARE", None)
        if cfg == "v1": os.environ["ISG_DENSE_V1"] = "1"
        else:
            c = cfg
            if "-" in c:
                c, spare = c.split("-"); os.environ["ISG_DENSE_SPARE"] = spare
            os.environ.pop("ISG_DENSE_DEBUG", None)
            if "!" in c:
                c, dbg = c.split("!"); os.environ["ISG_DENSE_DEBUG"] = dbg
            if "@" in c:
                c, tail = c.split("@"); os.environ["ISG_DENSE_TAIL"] = tail
            os.environ["ISG_DENSE_CFG"] = c
        _lib.lib().isg_debug_reload_tuning()
        dplan.label_map.fill_(-7); dplan.keepbits.fill_(0)
        try:
            for _ in range(3):
                pipe.run(d["kp"], d["ae"], d["anchors"], d["regression"], d["classification"], bench.CLS_TH, bench.IOU_TH)
            torch.cuda.synchronize()
            dplan.events = []
            for _ in range(30):
                pipe.run(d["kp"], d["ae"], d["anchors"], d["regression"], d["classification"], bench.CLS_TH, bench.IOU_TH, time_main=True)
            torch.cuda.synchronize()
        except Exception as e:
            print(cfg, "FAILED", repr(e)[:200], flush=True); continue
        ms = float(np.mean([a.elapsed_time(b) for a, b in dplan.events]))
        mn = float(np.min([a.elapsed_time(b) for a, b in dplan.events]))
        out = dict(lab=dplan.label_map.clone(), kb=dplan.keepbits.clone(), st=dplan.stats.clone(), idx=dplan.idx.clone(),
                   cnt=dplan.count.clone(), l=dplan.label.clone(), off=dplan.offsets.clone())
        msg = ""
        if ref is None: ref = out
        else:
            ndiff = int((out["lab"] != ref["lab"]).sum().item())
            kb = bool(torch.equal(out["kb"], ref["kb"])); st = bool(torch.equal(out["st"], ref["st"]))
            cnt = bool(torch.equal(out["cnt"], ref["cnt"]))
            M = out["cnt"].cpu().numpy(); okl = all(torch.equal(out["l"][b, :M[b]], ref["l"][b, :M[b]]) for b in range(B))
            msg = "label_map diff px=%d keepbits_eq=%s stats_eq=%s count_eq=%s keep_labels_eq=%s offsets_eq=%s" % (ndiff, kb, st, cnt, okl, bool(torch.equal(out["off"], ref["off"])))
        gbs = 24.0 * B * H * W / (ms * 1e-3) / 1e9
        print("%-8s mean %.1f us  min %.1f us  %.0f GB/s  frac %.3f  %s" % (cfg, ms * 1e3, mn * 1e3, gbs, gbs / 6463.7, msg), flush=True)

main()
