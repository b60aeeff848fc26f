"""device polygon stage vs host stage on the bench workload: identical detections? timing split"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))   # repo root (this file lives in tools/)
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
import isg_b200  # noqa
from isg_b200.utils import decode as dec
from helpers import DecodeCfg, IdentityTransforms, TransInfo
wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cityscapes_1024x2048_b8_n100"]
dev = torch.device("cuda", 0)
B, H, W = wl["B"], wl["H"], wl["W"]
host = bench.make_batch(wl, 0)
d = {k: v.to(dev) for k, v in host.items()}
cfg = DecodeCfg(kp_th=wl["kp_th"], cls_th=bench.CLS_TH, iou_th=bench.IOU_TH, wh_delta=bench.WH_DELTA)
infos = [TransInfo("/nonexistent.png", (H, W))] * B
inputs = torch.empty((B, 3, H, W), device="meta")
outs = ((d["kp"], d["ae"], None), d["regression"], d["classification"], d["anchors"])
res = {}
for name, flag in (("host", False), ("device", True)):
    dec.decode_mode, dec.device_polygon_stage = "dense", flag
    r = dec.decode_output(inputs, outs, infos, IdentityTransforms(), cfg, dev)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        r = dec.decode_output(inputs, outs, infos, IdentityTransforms(), cfg, dev)
    torch.cuda.synchronize()
    print(name, "ms/step", (time.perf_counter() - t0) / 5 * 1e3, "instances", sum(len(x) for x in r), dec.last_timing, flush=True)
    res[name] = r
nd = ne = nt = 0
for b in range(B):
    assert len(res["host"][b]) == len(res["device"][b]), (b, len(res["host"][b]), len(res["device"][b]))
    for (c1, f1, k1, p1), (c2, f2, k2, p2) in zip(res["host"][b], res["device"][b]):
        assert int(c1) == int(c2) and f1 == f2 and np.array_equal(k1, k2)
        nt += 1
        if np.array_equal(p1, p2): ne += 1
        else:
            key = lambda a: a[np.lexsort((a[:, 0], a[:, 1]))]
            assert p1.shape == p2.shape and np.array_equal(key(p1), key(p2))
            nd += 1
print("polygons", nt, "identical", ne, "tie-order differences", nd)
from isg_b200 import engine as _e
for key, plan in _e._plans.items():
    if key[0] == "d" and hasattr(plan, "inst_count"):
        cnt = plan.inst_count.cpu().numpy().ravel()
        fl = plan.inst_flags.cpu().numpy().ravel()
        print("plan", key[1:6], "instances", int((cnt > 0).sum()), "K max", int(cnt.max()), "top10", sorted(cnt.tolist())[-10:], "mean", float(cnt[cnt > 0].mean()), "flags", np.bincount(fl, minlength=3).tolist())
