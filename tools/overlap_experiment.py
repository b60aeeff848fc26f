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

    import time
    host_ms = [0.0]

    def run(n_pipes, n_steps, pipelined_tail):
        start = torch.cuda.Event(enable_timing=True); stop = torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
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
        host_ms[0] = (time.perf_counter() - t0) * 1e3 / n_steps
        torch.cuda.synchronize()
        return start.elapsed_time(stop) / n_steps

    def run_graph(n_pipes, n_cycles, pipelined_tail, reps):
        """capture n_cycles * n_pipes steps into one CUDA graph, replay it `reps` times"""
        cap_s = torch.cuda.Stream(device=dev)
        g = torch.cuda.CUDAGraph()
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=cap_s):
            cur = torch.cuda.current_stream(dev)
            fork = torch.cuda.Event(); fork.record(cur)
            for s in streams[:n_pipes]:
                s.wait_event(fork)
            for i in range(n_cycles * n_pipes):
                p, s = pipes[i % n_pipes], streams[i % n_pipes]
                with torch.cuda.stream(s):
                    p.run(d["kp"], d["ae"], d["anchors"], d["regression"], d["classification"], bench.CLS_TH, bench.IOU_TH,
                          tail="polygons", obj_pixel_th=2, pipelined=pipelined_tail)
            for p, s in zip(pipes[:n_pipes], streams[:n_pipes]):
                with torch.cuda.stream(s):
                    p.finish()
                ev = torch.cuda.Event(); ev.record(s); cur.wait_event(ev)
        g.replay(); torch.cuda.synchronize()
        start = torch.cuda.Event(enable_timing=True); stop = torch.cuda.Event(enable_timing=True)
        start.record()
        for _ in range(reps):
            g.replay()
        stop.record(); torch.cuda.synchronize()
        return start.elapsed_time(stop) / (reps * n_cycles * n_pipes)

    ref = None
    for spare in (0, 8, 16):
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
                try:
                    gms = run_graph(n_pipes, 4, tail, 25)
                except Exception as e:
                    gms = float("nan"); print("graph capture failed:", repr(e)[:300], flush=True)
                print("spare %2d pipes %d tail-pipelined %-5s  %.4f ms/step (host enqueue %.4f ms/step)  graph replay %.4f ms/step  results_equal=%s" %
                      (spare, n_pipes, tail, ms, host_ms[0], gms, same), flush=True)


main()
