#!/usr/bin/env python
"""Benchmark of the decode hot path (BASELINE.json metric: decoded Mpix/s at 1024x2048).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

One step = one pass of the device decode over one batch per GPU: box-head front-end + class-aware NMS ->
seeds -> top-k threshold -> fused dense embedding/membership/assignment -> compaction -> grouping.
`value` is device-timed (CUDA events, inputs resident in HBM); `e2e` is the same decode through the drop-in
`utils.decode.decode_output` with pinned HOST tensors in and Python polygon lists out (wall clock);
`--impl reference` times the CPU oracle port of the reference decode on the host cores.
Under torchrun every rank decodes its own batch (weak scaling, no collective on the data path); NCCL only
reduces the timings.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # per-GPU share of BASELINE config[2] (batch 64 at 1024x2048 over 8 GPUs, ~100 seeds/image) + the NMS of config[1]
    "cityscapes_1024x2048_b8_n100": dict(B=8, H=1024, W=2048, N=100, C=8, kp_th=20000, n_dup=2),
    # BASELINE config[1]
    "half_512x1024_b8_n50": dict(B=8, H=512, W=1024, N=50, C=8, kp_th=20000, n_dup=2),
    # BASELINE config[3] without k-means (dense crowd)
    "crowd_1024x2048_b4_n500": dict(B=4, H=1024, W=2048, N=500, C=8, kp_th=20000, n_dup=1),
    # CI-sized
    "tiny_256x512_b2_n12": dict(B=2, H=256, W=512, N=12, C=8, kp_th=3000, n_dup=2),
}
DEFAULT_WORKLOAD = "cityscapes_1024x2048_b8_n100"
CLS_TH, IOU_TH, WH_DELTA, OBJ_PIXEL_TH = 0.3, 0.2, 0.1, 2   # obj_pixel_th of the reference configs/decode_cfg.yaml
ALGO_BYTES_PER_PIXEL = 24   # dense kernel: kp + 4 ae planes read (20 B) + int32 label written (4 B); SURVEY.md §8d


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--mode", default="dense", choices=["dense", "sparse"])
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-images", type=int, default=2, help="images of the batch decoded by the CPU baseline")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--pipeline", action="store_true",
                    help="overlap the polygon tail of step s with the box head / NMS / top-k of step s+1 "
                         "(engine.DecodePipeline.run(pipelined=True)); default: every step runs on its own")
    return ap.parse_args()


def make_batch(wl: dict, rank: int, count: int | None = None):
    from isg_b200 import synth
    B = wl["B"] if count is None else count
    anchors = synth.make_anchors(wl["H"], wl["W"])
    scenes = [synth.make_scene(10_000 * (rank + 1) + b, wl["H"], wl["W"], wl["N"], wl["C"], anchors, wl["n_dup"], CLS_TH, IOU_TH)
              for b in range(B)]
    return dict(
        kp=torch.from_numpy(np.stack([s[0].kp for s in scenes])),                 # [B,1,H,W]
        ae=torch.from_numpy(np.stack([s[0].ae for s in scenes])),                 # [B,4,H,W]
        regression=torch.from_numpy(np.stack([s[1] for s in scenes])),            # [B,A,4]
        classification=torch.from_numpy(np.stack([s[2] for s in scenes])),        # [B,A,C]
        anchors=torch.from_numpy(anchors),                                        # [1,A,4]
    )


class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {}
        for n in ("HwSlowdown", "HwThermalSlowdown", "SwThermalSlowdown", "SwPowerCap", "HwPowerBrakeSlowdown"):
            v = getattr(nv, "nvmlClocksEventReason" + n, None) or getattr(nv, "nvmlClocksThrottleReason" + n, None)
            if v is not None:
                names[int(v)] = n
        get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                mask = int(get(self.h))
                for bit, n in names.items():
                    if mask & bit:
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.002)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        if self.nv is None or not self.samples:
            return None
        snake = {"HwSlowdown": "hw_slowdown", "HwThermalSlowdown": "hw_thermal_slowdown", "SwThermalSlowdown": "sw_thermal_slowdown",
                 "SwPowerCap": "sw_power_cap", "HwPowerBrakeSlowdown": "hw_power_brake_slowdown"}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "samples": len(self.samples),
                "reasons": sorted(snake[r] for r in self.reasons)}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(workload: str):
    """dram bytes per launch of the dense kernel from the committed ncu capture, if one exists for this workload"""
    try:
        with open(os.path.join(ROOT, "profiles", "dense_traffic.json")) as f:
            return json.load(f).get(workload)
    except Exception:
        return None


def oracle_decode_images(batch, wl, n_images):
    """CPU reference port: full decode_output (box head + NMS + grouping + polygons) of n_images."""
    from oracle import ref_decode as rd
    outs = ((batch["kp"][:n_images], batch["ae"][:n_images], None), batch["regression"][:n_images],
            batch["classification"][:n_images], batch["anchors"])
    t0 = time.perf_counter()
    res = rd.decode_output(wl["H"], wl["W"], outs, kp_th=wl["kp_th"], cls_th=CLS_TH, iou_th=IOU_TH, wh_delta=WH_DELTA)
    return time.perf_counter() - t0, res


def run_reference(args, wl, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n_img = max(1, min(args.cpu_images, wl["B"]))
    batch = make_batch(wl, 0, n_img)
    for _ in range(max(1, min(args.warmup, 1))):
        oracle_decode_images(batch, wl, 1)
    steps = max(1, min(args.steps, 5))
    t = 0.0
    for _ in range(steps):
        dt, _ = oracle_decode_images(batch, wl, n_img)
        t += dt
    mpix = n_img * wl["H"] * wl["W"] * steps / t / 1e6
    sample = "%d of %d images per step x %d steps (steps capped at 5): oracle decode_output = box head + torchvision NMS + group_kp + polygons" % (n_img, wl["B"], steps)
    line = {"impl": "reference", "metric": "decoded Mpix/s", "value": mpix, "unit": "Mpix/s", "n_gpus": args.gpus, "steps": steps,
            "warmup": 1, "ms_per_step": 1e3 * t / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, **{k: wl[k] for k in ("B", "H", "W", "N", "C", "kp_th")}},
            "cpu_baseline": {"value": mpix, "unit": "Mpix/s", "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
            "e2e": {"value": mpix, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_ours(args, wl, rank, world, local_rank):
    import isg_b200  # noqa: F401
    from isg_b200 import _lib, engine
    from isg_b200.utils import decode as dec
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import DecodeCfg, IdentityTransforms, TransInfo

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    _lib.lib()
    B, H, W, N = wl["B"], wl["H"], wl["W"], wl["N"]
    host = make_batch(wl, rank)
    pinned = {k: v.pin_memory() for k, v in host.items()}
    d = {k: v.to(dev) for k, v in host.items()}
    A, C = d["classification"].shape[1], d["classification"].shape[2]
    max_keep = max(64, 1 << int(np.ceil(np.log2(N * 1.3))))
    bplan = engine.BoxPlan(B, A, C, H, W, dev, cap=1024, max_keep=max_keep)
    dplan = engine.DecodePlan(B, H, W, bplan.N, wl["kp_th"], dev, args.mode, want_score=False, wh_delta=WH_DELTA)
    dplan.events = []
    pipe = engine.DecodePipeline(bplan, dplan)

    pipelined = args.mode == "dense" and args.pipeline

    def step(timed_kernel=False, overlap=None):
        # consecutive steps are pipelined: the polygon tail of step s overlaps the box head / NMS / top-k of step s+1
        # (engine.DecodePipeline.run(pipelined=True)); every step still runs all of its kernels on its own batch
        pipe.run(d["kp"], d["ae"], d["anchors"], d["regression"], d["classification"], CLS_TH, IOU_TH, time_main=timed_kernel,
                 tail="polygons" if args.mode == "dense" else "lists", obj_pixel_th=OBJ_PIXEL_TH,
                 pipelined=pipelined if overlap is None else overlap)

    # one untimed pass with the list tail: keep-pixel counts for the config block
    pipe.run(d["kp"], d["ae"], d["anchors"], d["regression"], d["classification"], CLS_TH, IOU_TH)
    for _ in range(max(args.warmup, 3)):
        step()
    pipe.finish()
    torch.cuda.synchronize(dev)
    n_keep = bplan.n_keep.cpu().numpy()
    n_cand = bplan.cand_count.cpu().numpy()
    assert n_keep.max() <= bplan.N and n_cand.max() <= bplan.cap, (n_keep, n_cand)
    counts = dplan.count.cpu().numpy()

    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize(dev)
    l0 = _lib.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        e0.record()
        for i in range(args.steps):
            # the roofline kernel is bracketed with CUDA events on every 8th step only: an event record between two
            # launches keeps the next kernel from being launched programmatically behind its predecessor
            step(timed_kernel=(i % 8 == 0))
        pipe.finish()                      # the last step's polygon tail is inside the timed region
        e1.record()
        torch.cuda.synchronize(dev)
    launches = _lib.launch_count - l0
    iso_ms = None
    if pipelined:   # latency of an isolated step (no overlap with a neighbour), reported next to the pipelined throughput
        l0e, l1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_iso = max(1, min(50, args.steps))
        l0e.record()
        for _ in range(n_iso):
            step(overlap=False)
        l1e.record()
        torch.cuda.synchronize(dev)
        iso_ms = l0e.elapsed_time(l1e) / n_iso
    ms = e0.elapsed_time(e1)
    kern_ms = float(np.mean([a.elapsed_time(b) for a, b in dplan.events])) if dplan.events else None
    t_max = ms
    if world > 1:
        import torch.distributed as dist
        tt = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_max = float(tt.item())
    value = world * B * H * W * args.steps / (t_max * 1e-3) / 1e6

    # ---- end to end through the drop-in API: pinned host tensors in, polygon lists out -------------
    e2e = None
    if not args.no_e2e:
        dec.decode_mode = args.mode
        cfg = DecodeCfg(kp_th=wl["kp_th"], cls_th=CLS_TH, iou_th=IOU_TH, wh_delta=WH_DELTA)
        infos = [TransInfo("/nonexistent.png", (H, W))] * B
        inputs = torch.empty((B, 3, H, W), device="meta")
        outs = ((pinned["kp"], pinned["ae"], None), pinned["regression"], pinned["classification"], d["anchors"])
        tf = IdentityTransforms()
        res = dec.decode_output(inputs, outs, infos, tf, cfg, dev)      # warm-up (allocates the plans)
        torch.cuda.synchronize(dev)
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        t0 = time.perf_counter()
        host_s = 0.0
        d2h = 0
        for _ in range(args.e2e_steps):
            res = dec.decode_output(inputs, outs, infos, tf, cfg, dev)
            host_s += dec.last_timing.get("host_polygons_s", 0.0)
            d2h = int(dec.last_timing.get("d2h_bytes", 0))
        torch.cuda.synchronize(dev)
        te = time.perf_counter() - t0
        # the H2D copy alone, for the split reported next to the e2e number
        hc0, hc1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        _dst = [torch.empty_like(pinned[k], device=dev) for k in ("kp", "ae", "regression", "classification")]
        torch.cuda.synchronize(dev)
        hc0.record()
        for _d, k in zip(_dst, ("kp", "ae", "regression", "classification")):
            _d.copy_(pinned[k], non_blocking=True)
        hc1.record()
        torch.cuda.synchronize(dev)
        h2d_ms = hc0.elapsed_time(hc1)
        del _dst
        if world > 1:
            tt = torch.tensor([te], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            te = float(tt.item())
        h2d = sum(pinned[k].numel() * 4 for k in ("kp", "ae", "regression", "classification"))
        n_inst = sum(len(r) for r in res)
        e2e = {"value": world * B * H * W * args.e2e_steps / te / 1e6, "unit": "Mpix/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "steps": args.e2e_steps, "ms_per_step": 1e3 * te / args.e2e_steps,
               "instances_per_step": n_inst, "h2d_ms_per_step": h2d_ms, "host_assembly_ms_per_step": 1e3 * host_s / args.e2e_steps,
               "polygons": "device (isg_instance_polygons); the host only slices the read-back buffers into the result lists",
               "h2d": "pinned host tensors, uploaded in chunks of %d images overlapped with the decode of the previous chunk" % dec.host_chunk_images}

    if rank != 0:
        return
    peak, peak_src = measured_peak()
    roofline = None
    if kern_ms:
        achieved = ALGO_BYTES_PER_PIXEL * B * H * W / (kern_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": "dense_v4_kernel (isg_assign_dense; its tile lists are prebuilt on the box branch)" if args.mode == "dense" else "assign_sparse_kernel",
                    "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic(args.workload),
                    "peak_source": peak_src, "kernel_ms": kern_ms, "algorithmic_bytes_per_launch": ALGO_BYTES_PER_PIXEL * B * H * W,
                    "kernel_timing": "CUDA events on the launching stream around the kernel on every 8th step of the timed region (%d launches)" % len(dplan.events)}
    cpu = None
    if not args.no_cpu and world == 1:      # the CPU baseline is taken on rank 0 at N=1 only
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        n_img = max(1, min(args.cpu_images, B))
        oracle_decode_images(host, wl, 1)
        dt, _ = oracle_decode_images(host, wl, n_img)
        cpu = {"value": n_img * H * W / dt / 1e6, "unit": "Mpix/s", "cores": torch.get_num_threads(), "kind": "port",
               "sample": "%d of the %d images of one step, once (after a 1-image warm-up): oracle decode_output incl. polygons" % (n_img, B)}
    line = {"metric": "decoded Mpix/s", "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": t_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "B_per_gpu": B, "H": H, "W": W, "seeds_per_image": [int(v) for v in n_keep],
                       "candidates_per_image": [int(v) for v in n_cand], "keep_pixels_per_image": [int(v) for v in counts],
                       "anchors": A, "classes": C, "kp_th": wl["kp_th"], "mode": args.mode,
                       "l2": "inputs are %.0f MB per step (> 126 MB L2); no flush" % (sum(v.numel() * 4 for v in d.values()) / 1e6),
                       "step": "box head + NMS + seeds + top-k + tile lists + fused assign + per-instance polygons (point sets, internal point, angular sort, centre test)",
                       "pipelining": ("consecutive steps overlap: the polygon tail of step s runs on its own stream next to the box head / NMS / top-k of step s+1"
                                      if pipelined else "none: every step runs on its own (--pipeline overlaps neighbours: 0.178 vs 0.181 ms measured in round 1)"),
                       "isolated_step_ms": iso_ms},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks.summary()}
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    wl = WORKLOADS[args.workload]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, wl, rank, world)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU path); use --impl reference for the CPU oracle")
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        # rank 0 prints ONE JSON line on stdout: NCCL writes its version banner ("NCCL version ...", printed at
        # NCCL_DEBUG=VERSION and above, which this image presets) to file descriptor 1 when the first communicator is
        # created, so stdout points at stderr while the process group comes up and the first collective runs
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            warm = torch.zeros(1, device=torch.device("cuda", local_rank))
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    try:
        run_ours(args, wl, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
