"""GPU experiment: how much of a decode step can be hidden by running neighbouring steps concurrently.
  python tools/overlap_experiment.py [workload] [steps]
Variants: n independent pipelines (own plans, own streams), steps issued round-robin; the persistent dense kernel leaves
`spare` SMs to the small kernels of the other pipelines (ISG_DENSE_SPARE)."""
import os, sys
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

    def make_pipe():
        bplan = engine.BoxPlan(B, A, C, H, W, dev, cap=1024, max_keep=max_keep)
        dplan = engine.DecodePlan(B, H, W, bplan.N, wl["kp_th"], dev, "dense", want_score=False, wh_delta=0.1)
        return engine.DecodePipeline(bplan, dplan)

    pipes = [make_pipe() for _ in range(4)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(4)]
    main_s = torch.cuda.current_stream(dev)

    def run(n_pipes, n_steps, pipelined_tail):
        start = torch.cuda.Event(enable_timing=True); stop = torch.cuda.Event(enable_timing=True)
        start.record(main_s)
        for s in streams[:n_pipes]:
            s.wait_event(start)
        for i in range(n_steps):
            p, s = pipes[i % n_pipes], streams[i % n_pipes]
            with torch.cuda.stream(s):
                p.run(d["kp"], d["ae"], d["anchors"], d["regression"], d["classification"], bench.CLS_TH, bench.IOU_TH,
                      tail="polygons", obj_pixel_th=2, pipelined=pipelined_tail)
        for p, s in zip(pipes[:n_pipes], streams[:n_pipes]):
            with torch.cuda.stream(s):
                p.finish()
            ev = torch.cuda.Event(); ev.record(s); main_s.wait_event(ev)
        stop.record(main_s)
        torch.cuda.synchronize()
        return start.elapsed_time(stop) / n_steps

    ref = None
    for spare in (0, 8, 16, 24, 32):
        os.environ["ISG_DENSE_SPARE"] = str(spare)
        _lib.lib().isg_debug_reload_tuning()
        for n_pipes in (1, 2, 3, 4):
            for tail in (False, True):
                run(n_pipes, 12, tail)
                ms = min(run(n_pipes, steps, tail) for _ in range(2))
                # identical results in every pipeline
                out = [(p.dplan.inst_count.clone(), p.dplan.inst_flags.clone(), p.dplan.img_total.clone()) for p in pipes[:n_pipes]]
                if ref is None:
                    ref = out[0]
                same = all(torch.equal(o[0], ref[0]) and torch.equal(o[1], ref[1]) and torch.equal(o[2], ref[2]) for o in out)
                print("spare %2d pipes %d tail-pipelined %-5s  %.4f ms/step  %.1f Gpix/s  results_equal=%s" %
                      (spare, n_pipes, tail, ms, B * H * W / ms / 1e6, same), flush=True)


main()
