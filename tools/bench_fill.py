"""Timing of the device polygon rasteriser (isg_fill_polygons, §8 f2) on the polygons of the bench workload, next to the
reference's per-detection cv2.fillPoly loop on the host (utils/eval_util.py:116).  Not a bench line."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import isg_b200  # noqa
from isg_b200 import engine
from isg_b200._lib import call
from isg_b200.engine import ptr, stream_ptr
from isg_b200.utils import image
import cv2

wl_name = sys.argv[1] if len(sys.argv) > 1 else bench.DEFAULT_WORKLOAD
wl = bench.WORKLOADS[wl_name]
dev = torch.device("cuda", 0)
B, H, W, N = wl["B"], wl["H"], wl["W"], wl["N"]
d = {k: v.to(dev) for k, v in bench.make_batch(wl, 0).items()}
A, C = d["classification"].shape[1], d["classification"].shape[2]
bplan = engine.BoxPlan(B, A, C, H, W, dev, cap=1024, max_keep=max(64, 1 << int(np.ceil(np.log2(N * 1.3)))))
dplan = engine.DecodePlan(B, H, W, bplan.N, wl["kp_th"], dev, "dense", want_score=False, wh_delta=bench.WH_DELTA)
pipe = engine.DecodePipeline(bplan, dplan)
pipe.run(d["kp"], d["ae"], d["anchors"], d["regression"], d["classification"], bench.CLS_TH, bench.IOU_TH, tail="polygons",
         obj_pixel_th=bench.OBJ_PIXEL_TH)
torch.cuda.synchronize()

filled = image.fill_instances(dplan)
ok = filled.desc[:, 0] == image.FILL_OK
n_poly, n_vert = int(ok.sum()), int(filled.desc[ok, 7].sum())
print("workload %s: %d polygons, %d vertices, %.2f MB of bit-packed boxes (%.1f MB as full frames)" % (
    wl_name, n_poly, n_vert, filled.used * 4 / 1e6, n_poly * H * ((W + 31) // 32) * 4 / 1e6))

# raw kernel timing, buffers preallocated
base = torch.arange(B, device=dev, dtype=torch.int32)[:, None] * dplan.cap
start = (dplan.inst_start + base).reshape(-1).contiguous()
count = torch.where(dplan.inst_flags == 1, dplan.inst_count, torch.zeros_like(dplan.inst_count)).reshape(-1).contiguous()
n = B * bplan.N
desc = torch.empty((n, 8), dtype=torch.int32, device=dev)
total = torch.empty(1, dtype=torch.int64, device=dev)
for full in (0, 1):
    cap = n * H * ((W + 31) // 32) if full else max(filled.used, 1)
    words = torch.empty(cap, dtype=torch.int32, device=dev)
    fn = lambda: call("isg_fill_polygons", ptr(dplan.poly_points), ptr(start), ptr(count), n, H, W, full, ptr(words), cap,
                      ptr(desc), ptr(total), stream_ptr(dev))
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
    print("isg_fill_polygons %-10s %8.1f us per batch (%d polygons)  -> %.1f GB/s of mask words written" % (
        "full-frame" if full else "compact", 1e3 * ms, n_poly, cap * 4 / ms / 1e6))
    del words

# host loop of the reference on the same polygons
masks = filled.masks()
pts = dplan.poly_points.view(-1, 2).cpu().numpy()
st, ct = start.cpu().numpy(), count.cpu().numpy()
polys = [pts[st[i]: st[i] + ct[i]] for i in range(n) if ct[i] > 0]
t0 = time.perf_counter()
ref = [cv2.fillPoly(np.zeros((H, W), np.int32), [p.astype(np.int32)], 1) for p in polys]
t_cpu = time.perf_counter() - t0
got = [m for m, c in zip(masks, ct) if c > 0]
assert len(got) == len(ref) and all(np.array_equal(a, b) for a, b in zip(got, ref)), "device masks differ from cv2.fillPoly"
print("host poly_to_mask loop (cv2.fillPoly, 1 core): %.1f ms for %d polygons; device masks identical" % (1e3 * t_cpu, len(ref)))
