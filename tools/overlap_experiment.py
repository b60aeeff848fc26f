"""GPU experiment: step time of the decode ring (engine.DecodeRing, one isg_decode_step call per step) against the ring
size, the SMs the dense kernel leaves free (ISG_DENSE_SPARE) and ablations that show each stage's marginal cost:
  python tools/overlap_experiment.py [workload] [steps]"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import isg_b200  # noqa
from isg_b200 import _lib, engine


def main():
    wlname = sys.argv[1] if len(sys.argv) > 1 else "cityscapes_1024x2048_b8_n100"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    wl = bench.WORKLOADS[wlname]
    dev = torch.device("cuda", 0)
    B, H, W, N = wl["B"], wl["H"], wl["W"], wl["N"]
    host = bench.make_batch(wl, 0)
    d = {k: v.to(dev) for k, v in host.items()}
    A, C = d["classification"].shape[1], d["classification"].shape[2]
    max_keep = max(64, 1 << int(np.ceil(np.log2(N * 1.3))))
    make = lambda: engine.make_pipeline(B, A, C, H, W, H, W, wl["kp_th"], dev, cand_cap=1024 if N < 300 else 2048, max_keep=max_keep)
    rings = {n: engine.DecodeRing(make, n) for n in (1, 2, 3, 4, 6)}

    def run(ring, n_steps, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(n_steps):
            ring.submit(d["kp"], d["ae"], d["anchors"], d["regression"], d["classification"], bench.CLS_TH, bench.IOU_TH,
                        obj_pixel_th=2, **kw)
        ring.wait()
        e1.record()
        host_ms = (time.perf_counter() - t0) * 1e3 / n_steps
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n_steps, host_ms

    def tune(**env):
        for k in ("ISG_DENSE_SPARE", "ISG_DENSE_DEBUG"):
            os.environ.pop(k, None)
        for k, v in env.items():
            os.environ[k] = str(v)
        _lib.lib().isg_debug_reload_tuning()

    for spare in (0, 8):
        tune(ISG_DENSE_SPARE=spare)
        for n, ring in rings.items():
            run(ring, 16)
            ms, host_ms = min(run(ring, steps) for _ in range(2))
            print("spare %2d ring %d: %.4f ms/step (host enqueue %.4f)  %.1f Gpix/s" % (spare, n, ms, host_ms, B * H * W / ms / 1e6), flush=True)
    ring = rings[4]
    for name, env, kw in (("full step", {}, {}), ("no polygon stage", {}, dict(polygons=False)),
                          ("dense consumers release without computing (loads only)", dict(ISG_DENSE_DEBUG=1), {}),
                          ("loads only + no polygon stage", dict(ISG_DENSE_DEBUG=1), dict(polygons=False)),
                          ("sparse assignment instead of dense", {}, dict(assign="sparse")),
                          ("sparse assignment, no polygon stage", {}, dict(assign="sparse", polygons=False))):
        tune(**env)
        run(ring, 16, **kw)
        ms, _ = min(run(ring, steps, **kw) for _ in range(2))
        print("ablation ring 4: %-58s %.4f ms/step" % (name, ms), flush=True)
    tune()


main()
