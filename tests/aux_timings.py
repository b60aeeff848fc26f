"""(Run by hand: `python tests/aux_timings.py` on a B200; not collected by pytest.  Lives under tests/ because it times the
CPU oracle next to the device.)
Timings of the secondary entry points on the BASELINE.json parity configurations (not bench lines): the seeded k-means of
config 4, mask NMS of config 5, py_cpu_nms and select_points — device (CUDA events) next to the CPU oracle (wall clock)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import isg_b200  # noqa
from isg_b200 import synth
from isg_b200.utils import decode as dec, kmeans as km, nms
from oracle import ref_decode as rd, ref_kmeans_nms as rk

dev = torch.device("cuda", 0)


def gpu_ms(fn, n=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def cpu_ms(fn, n=1):
    fn()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    return (time.perf_counter() - t0) * 1e3 / n


rows = []
# select_points at full size
kp = torch.randn(1024, 2048)
kpd = kp.to(dev)
rows.append(("select_points 1024x2048 k=20000", gpu_ms(lambda: dec.select_points(kpd, 20000)), cpu_ms(lambda: rd.select_points(kp, 20000))))
# k-means of config 4: M=20000 points, N=500 clusters, D=2
rs = np.random.RandomState(0)
ctr = rs.uniform(0, 1, size=(500, 2)).astype(np.float32)
X = (ctr[rs.randint(0, 500, size=20000)] + rs.normal(0, 0.01, size=(20000, 2))).astype(np.float32)
allow = np.full(500, 0.05, np.float32)
Xd, cd = torch.from_numpy(X).to(dev), torch.from_numpy(ctr).to(dev)
rows.append(("kmeans M=20000 N=500 D=2", gpu_ms(lambda: km.kmeans(Xd, 500, cd, allow, device=dev), 5),
             cpu_ms(lambda: rk.kmeans(torch.from_numpy(X), 500, torch.from_numpy(ctr), allow))))
# py_cpu_nms n=1000
dets = synth.make_nms_boxes(3, 1000)
dd = torch.from_numpy(dets).to(dev)
rows.append(("py_cpu_nms n=1000", gpu_ms(lambda: nms.py_cpu_nms(dd, 0.5)), cpu_ms(lambda: rk.py_cpu_nms(dets, 0.5), 3)))
# mask NMS of config 5: 1000 masks at 800x1333, 80 classes (bit-packed, as synth.make_masks emits them)
masks, boxes, scores, cls = synth.make_masks(5, 1000, 800, 1333, 80)
md = torch.from_numpy(masks.view(np.int32)).to(dev)
sc, cl = torch.from_numpy(scores).to(dev), torch.from_numpy(cls).to(dev)
rows.append(("mask_nms n=1000 800x1333 C=80", gpu_ms(lambda: nms.mask_nms(md, sc, cl, 0.5), 3),
             cpu_ms(lambda: rk.mask_nms(masks, scores, cls, 0.5))))
# f4: the kp / ae heads of the decoder at inference (B=8, 16 channels, 1024x2048): 64 B/px read + 20 B/px written
from isg_b200.utils.heads import InferenceHeads
kc, ac = torch.nn.Conv2d(16, 1, 1), torch.nn.Conv2d(16, 4, 1)
heads = InferenceHeads(kc, ac)
xf = torch.randn((8, 16, 1024, 2048), device=dev)
g_ms = gpu_ms(lambda: heads(xf), 10)
xc = xf[:1].cpu()
c_ms = cpu_ms(lambda: (torch.nn.functional.conv2d(xc, kc.weight, kc.bias), torch.nn.functional.conv2d(xc, ac.weight, ac.bias),
                       torch.nn.functional.conv2d(xc, torch.zeros(2, 16, 1, 1), None))) * 8
rows.append(("decoder heads kp+ae B=8 16ch 1024x2048 (%.0f GB/s)" % (84 * 8 * 1024 * 2048 / g_ms / 1e6), g_ms, c_ms))
del xf
for name, g, c in rows:
    print("%-58s device %9.3f ms   cpu %10.2f ms" % (name, g, c), flush=True)
