"""Per-phase clock64 stamps of instance_polygons_kernel on the bench workload.  Needs a debug build of the library:
`ISG_NVCC_EXTRA=-DISG_POLY_DEBUG python instance-segmentation_b200/build.py --force` (adds the stamp stores and the
`isg_debug_poly_stamps` export; rebuild without the flag afterwards).  Not a bench line."""
import os, sys, ctypes
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import isg_b200
from isg_b200 import engine, _lib
wl = bench.WORKLOADS[bench.DEFAULT_WORKLOAD]
dev = torch.device("cuda", 0)
B, H, W, N = wl["B"], wl["H"], wl["W"], wl["N"]
d = {k: v.to(dev) for k, v in bench.make_batch(wl, 0).items()}
A, C = d["classification"].shape[1], d["classification"].shape[2]
bplan = engine.BoxPlan(B, A, C, H, W, dev, cap=1024, max_keep=max(64, 1 << int(np.ceil(np.log2(N * 1.3)))))
dplan = engine.DecodePlan(B, H, W, bplan.N, wl["kp_th"], dev, "dense", want_score=False, wh_delta=bench.WH_DELTA)
pipe = engine.DecodePipeline(bplan, dplan)
for _ in range(5):
    pipe.run(d["kp"], d["ae"], d["anchors"], d["regression"], d["classification"], bench.CLS_TH, bench.IOU_TH, tail="polygons", obj_pixel_th=bench.OBJ_PIXEL_TH)
torch.cuda.synchronize()
lib = _lib.lib()
buf = np.zeros(4096 * 8, np.int64)
lib.isg_debug_poly_stamps(buf.ctypes.data_as(ctypes.c_void_p))
st = buf.reshape(4096, 8)[: B * bplan.N]
act = st[:, 7] > 0
dd = st[act]
ph = np.diff(dd, axis=1)
names = ["cand count+scan", "list+labels+scan", "emit", "internal pt", "angles", "sort", "gather+centre test"]
print("active CTAs", act.sum(), "of", len(st))
for k, nme in enumerate(names):
    print("%-20s mean %7.0f  p90 %7.0f  max %7.0f cycles" % (nme, ph[:, k].mean(), np.percentile(ph[:, k], 90), ph[:, k].max()))
tot = dd[:, 7] - dd[:, 0]
print("total mean %.0f p90 %.0f max %.0f cycles (%.1f us at 1.9 GHz)" % (tot.mean(), np.percentile(tot, 90), tot.max(), tot.max() / 1900))
cnt = dplan.inst_count.cpu().numpy().reshape(-1)[: len(st)]
# CTA id = inst * B + b ; inst_count index = b * N + inst
cid = np.nonzero(act)[0]; inst = cid // B; b = cid % B
K = dplan.inst_count.cpu().numpy()[b, inst]
o = np.argsort(tot)[-6:]
print("slowest K:", K[o].tolist(), "total", tot[o].tolist(), "phases", ph[o].tolist())
print("corr(total, K) = %.2f" % np.corrcoef(K, tot)[0, 1])
