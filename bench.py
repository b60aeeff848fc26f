#!/usr/bin/env python
"""Benchmark of the decode hot path (BASELINE.json metric: decoded Mpix/s at 1024x2048).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME] [--inputs default|wide]
                  [--global-batch G] [--min-seconds S]

One step = one pass of the device decode over one batch per GPU: box-head front-end + class-aware NMS -> seeds ->
top-k threshold -> tile lists -> fused dense embedding/membership/assignment -> per-instance polygons, enqueued by ONE host
call (isg_decode_step).  Consecutive steps go round-robin through a ring of independent pipelines (engine.DecodeRing), so
the small kernels of neighbouring steps fill the machine around each step's dense kernel; every step runs all of its
kernels on its batch.
`value` is device-timed (CUDA events, inputs resident in HBM); `e2e` is the same decode through the drop-in
`utils.decode.decode_output` with pinned HOST tensors in and Python polygon lists out (wall clock);
`--impl reference` times the UNMODIFIED reference's decode_output (baseline/_ref, copied at build time; the oracle port
when that copy is missing) on the host cores.
Under torchrun every rank decodes its own batch (weak scaling) or its shard of --global-batch (strong scaling, BASELINE
config 3); there is no collective on the data path, NCCL only reduces the timings (isg_b200.dist).
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
    # BASELINE config[3]: dense crowd; the _kmeans variant adds the seeded k-means refinement of utils/kmeans.py on the
    # embeddings of the keep pixels (X = e[M,2], centres = seed coordinates, allow = 0.05)
    "crowd_1024x2048_b4_n500": dict(B=4, H=1024, W=2048, N=500, C=8, kp_th=20000, n_dup=1),
    "crowd_1024x2048_b4_n500_kmeans": dict(B=4, H=1024, W=2048, N=500, C=8, kp_th=20000, n_dup=1, kmeans=True),
    # BASELINE config[4]: 1000 candidate masks at 800x1333, 80 classes -> bit-packed mask-IoU NMS (isg_mask_nms)
    "coco_800x1333_c80_n1000": dict(B=1, H=800, W=1333, N=1000, C=80, kind="mask_nms", thr=0.5),
    # CI-sized
    "tiny_256x512_b2_n12": dict(B=2, H=256, W=512, N=12, C=8, kp_th=3000, n_dup=2),
}
DEFAULT_WORKLOAD = "cityscapes_1024x2048_b8_n100"
CLS_TH, IOU_TH, WH_DELTA, OBJ_PIXEL_TH = 0.3, 0.2, 0.1, 2   # the reference's configs/decode_cfg.yaml
ALGO_BYTES_PER_PIXEL = 24   # dense kernel: kp + 4 ae planes read (20 B) + int32 label written (4 B); SURVEY.md §8d
KERNEL_TIMING_LAUNCHES = 24


def l2_note(in_bytes, out_bytes):
    tot = (in_bytes + out_bytes) / 1e6
    return ("a step reads %.0f MB of inputs and writes a %.0f MB label map: %.0f MB %s the 126 MB L2; no flush%s"
            % (in_bytes / 1e6, out_bytes / 1e6, tot, ">" if tot > 126 else "<", "" if tot > 126 else " (CI-sized workload, not a bench line)"))


def workload_config(args, wl):
    """The `config` block of a bench line: static facts of the workload only, so that the repo arm and the reference arm
    (--impl reference) print the SAME dictionary for the same command line.  Measured facts of the batch (seeds, candidates,
    keep pixels per image, ...) go to `workload_stats`."""
    if wl.get("kind") == "mask_nms":
        H, W, n = wl["H"], wl["W"], wl["N"]
        return {"workload": args.workload, "masks": n, "H": H, "W": W, "classes": wl["C"], "iou_thr": wl["thr"],
                "step": "mask-IoU NMS of %d bit-packed masks: per-mask popcount + tight bbox, score sort, same-class pair IoU on the bbox "
                        "intersection, greedy scan" % n,
                "l2": "masks are %.0f MB per step (> 126 MB L2); no flush" % (n * H * ((W + 31) // 32) * 4 / 1e6)}
    H, W = wl["H"], wl["W"]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    B = -(-args.global_batch // world) if args.global_batch > 0 else wl["B"]      # rank 0's shard (dist.shard_range)
    A, C = 9 * sum((H // s) * (W // s) for s in (8, 16, 32, 64, 128)), wl["C"]     # utils/utils.py:419-443
    cfg = {"workload": args.workload, "inputs": args.inputs, "B_per_gpu": B, "H": H, "W": W, "seeds_nominal": wl["N"], "anchors": A,
           "classes": C, "kp_th": wl["kp_th"], "cls_th": CLS_TH, "iou_th": IOU_TH, "wh_delta": WH_DELTA, "obj_pixel_th": OBJ_PIXEL_TH,
           "l2": l2_note(B * (5 * H * W + A * (4 + C)) * 4 + A * 16, B * H * W * 4),
           "step": "decode_output of one batch: box head + NMS + seeds + top-k + 3x3 peaks + embedding / membership / assignment + "
                   "per-instance polygons (point sets, internal point, angular sort, centre test)"
                   + (" + k-means refinement of every image" if wl.get("kmeans") else "")}
    if args.global_batch > 0:
        cfg["global_batch"] = args.global_batch
    return cfg


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--inputs", default="default", choices=["default", "wide"],
                    help="wide: embedding offsets ae[0:2] ~ N(0,1) off the outlines (unbounded logits: the tanh of the dense "
                         "kernel takes its exp/divide branch instead of the |x| < 0.55 polynomial)")
    ap.add_argument("--ring", type=int, default=6, help="independent pipelines used round-robin (1 = every step on its own)")
    ap.add_argument("--global-batch", type=int, default=0,
                    help="strong scaling: decode this many images per step in total, sharded over the ranks (config 3: 64)")
    ap.add_argument("--min-seconds", type=float, default=0.0, help="repeat the timed region until it lasted this long (sustained line)")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-images", type=int, default=1, help="images of the batch decoded by the CPU baseline")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    return ap.parse_args()


def make_batch(wl: dict, rank: int, count: int | None = None, inputs: str = "default"):
    from isg_b200 import synth
    B = wl["B"] if count is None else count
    anchors = synth.make_anchors(wl["H"], wl["W"])
    n_scenes = min(B, 8)           # larger batches repeat the 8 scenes (distinct memory, same content)
    scenes = [synth.make_scene(10_000 * (rank + 1) + b, wl["H"], wl["W"], wl["N"], wl["C"], anchors, wl["n_dup"], CLS_TH, IOU_TH)
              for b in range(n_scenes)]
    if inputs == "wide":
        for b, s in enumerate(scenes):
            rs = np.random.RandomState(77 + b)
            off = s[0].owner < 0
            s[0].ae[0][off] = rs.normal(0.0, 1.0, size=int(off.sum())).astype(np.float32)
            s[0].ae[1][off] = rs.normal(0.0, 1.0, size=int(off.sum())).astype(np.float32)
    pick = [scenes[b % n_scenes] for b in range(B)]
    return dict(
        kp=torch.from_numpy(np.stack([s[0].kp for s in pick])),                 # [B,1,H,W]
        ae=torch.from_numpy(np.stack([s[0].ae for s in pick])),                 # [B,4,H,W]
        regression=torch.from_numpy(np.stack([s[1] for s in pick])),            # [B,A,4]
        classification=torch.from_numpy(np.stack([s[2] for s in pick])),        # [B,A,C]
        anchors=torch.from_numpy(anchors),                                      # [1,A,4]
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


def ncu_traffic(key: str):
    """dram bytes per launch of the roofline kernel from the committed ncu capture, if one exists for this workload"""
    try:
        with open(os.path.join(ROOT, "profiles", "dense_traffic.json")) as f:
            return json.load(f).get(key)
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------------------------
# CPU side: the reference itself (baseline/_ref) or, when that copy is missing, the oracle port
# ------------------------------------------------------------------------------------------------------------------
_reference = None


def load_reference():
    """the unmodified reference's modules (oracle/reference_loader.py), or None"""
    global _reference
    if _reference is None:
        try:
            from oracle import reference_loader as rl
            _reference = rl.Reference() if rl.available() else False
        except Exception as e:            # e.g. a clashing top-level `utils` package
            print("reference unavailable: %r" % (e,), file=sys.stderr)
            _reference = False
    return _reference or None


class _Timer:
    def __init__(self):
        self.t = {}

    def wrap(self, mod, name):
        fn = getattr(mod, name)

        def timed(*a, **k):
            t0 = time.perf_counter()
            try:
                return fn(*a, **k)
            finally:
                self.t[name] = self.t.get(name, 0.0) + time.perf_counter() - t0
        setattr(mod, name, timed)
        return fn


def cpu_decode_images(batch, wl, n_images):
    """decode_output of the first n_images on the host.  Returns (seconds, results, kind, per-function split in ms)."""
    outs = ((batch["kp"][:n_images], batch["ae"][:n_images].clone(), None), batch["regression"][:n_images].clone(),
            batch["classification"][:n_images].clone(), batch["anchors"])
    ref = load_reference()
    if ref is None:
        from oracle import ref_decode as rd
        t0 = time.perf_counter()
        res = rd.decode_output(wl["H"], wl["W"], outs, kp_th=wl["kp_th"], cls_th=CLS_TH, iou_th=IOU_TH, wh_delta=WH_DELTA)
        return time.perf_counter() - t0, res, "port", None
    rdec = ref.decode
    cfg = ref.cfg
    cfg.kp_th, cfg.cls_th, cfg.iou_th, cfg.wh_delta, cfg.obj_pixel_th, cfg.draw_flag = wl["kp_th"], CLS_TH, IOU_TH, WH_DELTA, OBJ_PIXEL_TH, False
    infos = [ref.TransInfo("/nonexistent.png", (wl["H"], wl["W"]))] * n_images
    inputs = torch.empty((n_images, 3, wl["H"], wl["W"]), device="meta")      # only its shape is read (utils/decode.py:378,445)
    tm = _Timer()
    saved = {n: tm.wrap(rdec, n) for n in ("decode_boxes", "select_points", "aug_group", "group_kp")}
    devnull = os.open(os.devnull, os.O_WRONLY)
    err = os.dup(2)
    try:
        os.dup2(devnull, 2)               # cv2.imread of the missing image file warns once per image (:335)
        t0 = time.perf_counter()
        res = rdec.decode_output(inputs, outs, infos, ref.transforms, cfg, torch.device("cpu"))
        dt = time.perf_counter() - t0
    finally:
        os.dup2(err, 2); os.close(err); os.close(devnull)
        for n, fn in saved.items():
            setattr(rdec, n, fn)
    t = tm.t
    split = {"decode_boxes": 1e3 * t.get("decode_boxes", 0.0), "select_points": 1e3 * t.get("select_points", 0.0),
             "aug_group (polygons)": 1e3 * t.get("aug_group", 0.0),
             "group_kp remainder (embedding, masked_select, membership, per-instance filter)":
                 1e3 * (t.get("group_kp", 0.0) - t.get("select_points", 0.0) - t.get("aug_group", 0.0)),
             "total": 1e3 * dt}
    return dt, res, "reference", split


def cpu_side_functions():
    """kmeans at the config-4 shape and py_cpu_nms at n=1000 on the host (reference when available), in ms"""
    from isg_b200 import synth
    rs = np.random.RandomState(0)
    ctr = rs.uniform(0, 1, size=(500, 2)).astype(np.float32)
    X = (ctr[rs.randint(0, 500, size=20000)] + rs.normal(0, 0.01, size=(20000, 2))).astype(np.float32)
    allow = np.full(500, 0.05, np.float32)
    dets = synth.make_nms_boxes(3, 1000)
    ref = load_reference()
    if ref is not None:
        km = lambda: ref.kmeans.kmeans(torch.from_numpy(X), 500, torch.from_numpy(ctr), allow)
        nm = lambda: ref.nms.py_cpu_nms(dets, 0.5)
    else:
        from oracle import ref_kmeans_nms as rk
        km = lambda: rk.kmeans(torch.from_numpy(X), 500, torch.from_numpy(ctr), allow)
        nm = lambda: rk.py_cpu_nms(dets, 0.5)
    out = {}
    for name, fn in (("kmeans M=20000 N=500", km), ("py_cpu_nms n=1000", nm)):
        fn()
        t0 = time.perf_counter()
        fn()
        out[name] = 1e3 * (time.perf_counter() - t0)
    return out


def run_reference(args, wl, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    if wl.get("kind") == "mask_nms":
        return run_reference_mask_nms(args, wl)
    n_img = max(1, min(args.cpu_images, wl["B"]))
    batch = make_batch(wl, 0, n_img, args.inputs)
    for _ in range(max(1, min(args.warmup, 1))):
        cpu_decode_images(batch, wl, 1)
    steps = max(1, min(args.steps, 3))
    t, split, kind = 0.0, None, "port"
    for _ in range(steps):
        dt, _, kind, sp = cpu_decode_images(batch, wl, n_img)
        t += dt
        if sp is not None:
            split = sp if split is None else {k: split[k] + v for k, v in sp.items()}
    if split is not None:
        split = {k: v / (steps * n_img) for k, v in split.items()}
        split.update({k + " (separate call)": v for k, v in cpu_side_functions().items()})
    mpix = n_img * wl["H"] * wl["W"] * steps / t / 1e6
    what = ("the unmodified reference's utils.decode.decode_output (baseline/_ref, torch CPU + numpy + cv2)" if kind == "reference"
            else "oracle port of decode_output (baseline/_ref missing)")
    sample = "%d of %d images per step x %d steps (steps capped at 3, 1 warm-up image): %s" % (n_img, wl["B"], steps, what)
    line = {"impl": "reference", "metric": "decoded Mpix/s", "value": mpix, "unit": "Mpix/s", "n_gpus": args.gpus, "steps": steps,
            "warmup": 1, "ms_per_step": 1e3 * t / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, wl),
            "cpu_baseline": {"value": mpix, "unit": "Mpix/s", "cores": torch.get_num_threads(), "kind": kind, "sample": sample,
                             "split_ms_per_image": split},
            "e2e": {"value": mpix, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
# config 5: bit-packed mask-IoU NMS
# ------------------------------------------------------------------------------------------------------------------
def make_masks(wl):
    from isg_b200 import synth
    return synth.make_masks(5, wl["N"], wl["H"], wl["W"], wl["C"])


def cpu_mask_nms(masks, scores, cls, thr, n):
    """greedy mask NMS of the first n masks on the host: IoU of utils/image.py:188-191 inside the loop of utils/nms.py:23-37"""
    from oracle import ref_kmeans_nms as rk
    t0 = time.perf_counter()
    keep = rk.mask_nms(masks[:n], scores[:n], cls[:n], thr)
    return time.perf_counter() - t0, keep


def run_reference_mask_nms(args, wl):
    masks, _, scores, cls = make_masks(wl)
    n = min(wl["N"], 250)
    dt, keep = cpu_mask_nms(masks, scores, cls, wl["thr"], n)
    mpix = n * wl["H"] * wl["W"] / dt / 1e6
    sample = ("first %d of the %d masks, once: oracle port of the greedy loop of utils/nms.py:23-37 with the mask IoU of "
              "utils/image.py:188-191 on bit-packed masks (the reference has no mask NMS function; pair count grows with n^2)" % (n, wl["N"]))
    line = {"impl": "reference", "metric": "decoded Mpix/s", "value": mpix, "unit": "Mpix/s", "n_gpus": args.gpus, "steps": 1,
            "warmup": 0, "ms_per_step": 1e3 * dt, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
            "data": "synthetic", "config": workload_config(args, wl),
            "cpu_baseline": {"value": mpix, "unit": "Mpix/s", "cores": 1, "kind": "port", "sample": sample},
            "e2e": {"value": mpix, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_mask_nms(args, wl, rank, world, local_rank):
    import isg_b200  # noqa: F401
    from isg_b200 import _lib, dist as idist, engine
    from isg_b200.utils import nms
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    masks, boxes, scores, cls = make_masks(wl)
    n, H, W = wl["N"], wl["H"], wl["W"]
    Ww = (W + 31) // 32
    md = torch.from_numpy(masks.view(np.int32)).to(dev)
    sc = torch.from_numpy(scores).to(dev)
    cl = torch.from_numpy(np.unique(cls, return_inverse=True)[1].astype(np.int32)).to(dev)
    pin = torch.from_numpy(masks.view(np.int32)).pin_memory()
    keep = torch.empty(n, dtype=torch.int32, device=dev)
    n_keep = torch.empty(1, dtype=torch.int32, device=dev)
    ws_bytes = int(_lib.lib().isg_mask_nms_workspace_bytes(n))
    ws, ws_ptr = engine.aligned_workspace(ws_bytes, dev)

    def step():     # the C entry point itself: per-mask popcount + tight bbox, score sort, same-class pair IoUs, greedy scan
        _lib.call("isg_mask_nms", md.data_ptr(), n, H, Ww, 0, sc.data_ptr(), cl.data_ptr(), float(wl["thr"]), keep.data_ptr(),
                  n_keep.data_ptr(), ws_ptr, ws_bytes, engine.stream_ptr(dev))

    for _ in range(max(args.warmup, 3)):
        step()
    idist.barrier()
    torch.cuda.synchronize(dev)
    l0 = _lib.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = args.steps
    with ClockSampler(local_rank) as clocks:
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize(dev)
    launches = _lib.launch_count - l0
    t_max = idist.reduce_max(e0.elapsed_time(e1), dev)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(KERNEL_TIMING_LAUNCHES)]
    for a, b in ev:
        a.record(); step(); b.record()
        torch.cuda.synchronize(dev)
    kern_ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    kept = int(n_keep.item())
    # end to end through the drop-in: pinned host masks in, keep list (numpy) out
    te = None
    if not args.no_e2e:
        nms.mask_nms(pin, scores, cls, wl["thr"])
        torch.cuda.synchronize(dev)
        idist.barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            k2 = nms.mask_nms(pin, scores, cls, wl["thr"])
        torch.cuda.synchronize(dev)
        te = idist.reduce_max(time.perf_counter() - t0, dev)
        assert len(k2) == kept
    if rank != 0:
        return
    peak, peak_src = measured_peak()
    algo = n * H * Ww * 4
    achieved = algo / (kern_ms * 1e-3) / 1e9
    cpu = None
    if not args.no_cpu and world == 1:
        nc = 150
        dt, _ = cpu_mask_nms(masks, scores, cls, wl["thr"], nc)
        cpu = {"value": nc * H * W / dt / 1e6, "unit": "Mpix/s", "cores": 1, "kind": "port",
               "sample": "first %d of the %d masks, once: oracle port of the greedy loop of utils/nms.py:23-37 with the mask IoU of "
                         "utils/image.py:188-191 (the reference has no mask NMS function); pair work grows with n^2" % (nc, n)}
    line = {"metric": "decoded Mpix/s", "value": world * n * H * W * steps / (t_max * 1e-3) / 1e6, "unit": "Mpix/s", "n_gpus": world,
            "steps": steps, "warmup": max(args.warmup, 3), "ms_per_step": t_max / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": workload_config(args, wl), "workload_stats": {"kept": kept},
            "roofline": {"bound": "hbm", "kernel": "isg_mask_nms (mask_area_kernel + nms_sort + mask_pair_kernel + nms_scan); algorithmic bytes = one read of every mask word",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic(args.workload),
                         "peak_source": peak_src, "kernel_ms": kern_ms, "algorithmic_bytes_per_launch": algo,
                         "kernel_timing": "CUDA events on the launching stream around the entry point, %d isolated calls after the timed region" % KERNEL_TIMING_LAUNCHES},
            "cpu_baseline": cpu,
            "e2e": None if te is None else {"value": world * n * H * W * args.e2e_steps / te / 1e6, "unit": "Mpix/s", "h2d_bytes_per_step": algo + 8 * n,
                                            "d2h_bytes_per_step": 4 * kept + 4, "steps": args.e2e_steps, "ms_per_step": 1e3 * te / args.e2e_steps},
            "gpu_launches": launches, "clocks": clocks.summary()}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
# decode workloads
# ------------------------------------------------------------------------------------------------------------------
def run_ours(args, wl, rank, world, local_rank):
    import isg_b200  # noqa: F401
    from isg_b200 import _lib, dist as idist, engine
    from isg_b200.utils import decode as dec, kmeans as km
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import DecodeCfg, IdentityTransforms, TransInfo

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    _lib.lib()
    H, W, N = wl["H"], wl["W"], wl["N"]
    strong = args.global_batch > 0
    if strong:
        lo, hi = idist.shard_range(args.global_batch, rank, world)
        B = hi - lo
    else:
        B = wl["B"]
    host = make_batch(wl, rank, B, args.inputs)
    pinned = {k: v.pin_memory() for k, v in host.items()}
    d = {k: v.to(dev) for k, v in host.items()}
    A, C = d["classification"].shape[1], d["classification"].shape[2]
    max_keep = max(64, 1 << int(np.ceil(np.log2(N * 1.3))))
    make = lambda: engine.make_pipeline(B, A, C, H, W, H, W, wl["kp_th"], dev, cand_cap=1024 if N < 300 else 2048, max_keep=max_keep,
                                        wh_delta=WH_DELTA)
    ring = engine.DecodeRing(make, max(1, args.ring))
    use_kmeans = bool(wl.get("kmeans"))

    def submit(timed_kernel=False):
        return ring.submit(d["kp"], d["ae"], d["anchors"], d["regression"], d["classification"], CLS_TH, IOU_TH,
                           obj_pixel_th=OBJ_PIXEL_TH, assign="dense", time_main=timed_kernel)

    # warm-up; the kept / candidate / keep-pixel counts of the batch for the config block
    for _ in range(max(args.warmup, 3)):
        submit()
    ring.wait()
    torch.cuda.synchronize(dev)
    p0 = ring.pipes[0]
    p0.bplan.arena.fetch_async(); p0.bplan.arena.wait()
    n_keep = p0.bplan.host["n_keep"].numpy().copy()
    n_cand = p0.bplan.host["cand_count"].numpy().copy()
    assert n_keep.max() <= p0.bplan.N and n_cand.max() <= p0.bplan.cap, (n_keep, n_cand)
    counts = np.array([int(v) for v in engine.popcount_rows(p0.dplan.keepbits)])

    km_state = None
    if use_kmeans:
        # X = embeddings of the keep pixels, centres = seed coordinates of the image (utils/kmeans.py:16-93 semantics)
        km_state = dict(emb=torch.empty((B, p0.dplan.cap, 2), dtype=torch.float32, device=dev), iters=[],
                        allow=np.full(int(n_keep.max()), 0.05, np.float32), ms=[])

    def refine(slot):
        """config 4: seeded k-means over the embeddings of every image's keep pixels"""
        pipe = ring.pipes[slot]
        dp, bp = pipe.dplan, pipe.bplan
        ring.wait(slot)
        s = engine.stream_ptr(dev)
        _lib.call("isg_compact_points", engine.ptr(dp.keepbits), B, H, W, dp.cap, engine.ptr(dp.idx), engine.ptr(dp.count), s)
        _lib.call("isg_gather_embeddings", engine.ptr(d["ae"]), d["ae"].stride(0), d["ae"].stride(1), engine.ptr(dp.idx),
                  engine.ptr(dp.count), dp.cap, B, H, W, engine.ptr(dp.ys), engine.ptr(dp.xs), engine.ptr(km_state["emb"]), s)
        for b in range(B):
            n, m = int(n_keep[b]), int(counts[b])
            centres = dp.seeds[b, :n, 4:6].view(torch.float32)
            km.kmeans(km_state["emb"][b, :m], n, centres, km_state["allow"][:n], device=dev)
            km_state["iters"].append(km.kmeans.last_iterations)

    if use_kmeans:
        refine(submit())
    idist.barrier()
    torch.cuda.synchronize(dev)
    l0 = _lib.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    total_ms, total_steps = 0.0, 0
    with ClockSampler(local_rank) as clocks:
        while True:
            e0.record()
            for i in range(args.steps):
                slot = submit()
                if use_kmeans:
                    refine(slot)
            ring.wait()                        # every step's polygon stage is inside the timed region
            e1.record()
            torch.cuda.synchronize(dev)
            total_ms += e0.elapsed_time(e1)
            total_steps += args.steps
            if total_ms >= 1e3 * args.min_seconds:
                break
    launches = _lib.launch_count - l0
    t_max = idist.reduce_max(total_ms, dev)
    n_img_total = args.global_batch if strong else world * B
    value = n_img_total * H * W * total_steps / (t_max * 1e-3) / 1e6

    # ---- the roofline kernel: CUDA events around isg_assign_dense ----------------------------------------------------
    # (a) in flight: every 8th step of a ring pass (other pipelines' kernels share the machine with it)
    for p in ring.pipes:
        p.dplan.events = []
    for i in range(8 * len(ring.pipes)):
        submit(timed_kernel=(i % 8 == 0))
    ring.wait()
    torch.cuda.synchronize(dev)
    in_flight = [a.elapsed_time(b) for p in ring.pipes for a, b in p.dplan.events]
    # (b) on its own: KERNEL_TIMING_LAUNCHES isolated steps (the GPU is idle when each step starts)
    p0.dplan.events = []
    for _ in range(KERNEL_TIMING_LAUNCHES):
        p0.run_native(d["kp"], d["ae"], d["anchors"], d["regression"], d["classification"], CLS_TH, IOU_TH, obj_pixel_th=OBJ_PIXEL_TH,
                      assign="dense", time_main=True)
        torch.cuda.synchronize(dev)
    kern_ms = float(np.mean([a.elapsed_time(b) for a, b in p0.dplan.events]))
    iso0, iso1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_iso = 50
    iso0.record()
    for _ in range(n_iso):
        p0.run_native(d["kp"], d["ae"], d["anchors"], d["regression"], d["classification"], CLS_TH, IOU_TH, obj_pixel_th=OBJ_PIXEL_TH)
    iso1.record()
    torch.cuda.synchronize(dev)
    iso_ms = iso0.elapsed_time(iso1) / n_iso

    # ---- end to end through the drop-in API: pinned host tensors in, polygon lists out -------------
    e2e = None
    if not args.no_e2e:
        dec.decode_mode = "dense"
        cfg = DecodeCfg(kp_th=wl["kp_th"], cls_th=CLS_TH, iou_th=IOU_TH, wh_delta=WH_DELTA)
        infos = [TransInfo("/nonexistent.png", (H, W))] * B
        inputs = torch.empty((B, 3, H, W), device="meta")
        outs = ((pinned["kp"], pinned["ae"], None), pinned["regression"], pinned["classification"], d["anchors"])
        tf = IdentityTransforms()
        res = dec.decode_output(inputs, outs, infos, tf, cfg, dev)      # warm-up (allocates the plans)
        res = dec.decode_output(inputs, outs, infos, tf, cfg, dev)
        torch.cuda.synchronize(dev)
        idist.barrier()
        t0 = time.perf_counter()
        host_s = 0.0
        for _ in range(args.e2e_steps):
            res = dec.decode_output(inputs, outs, infos, tf, cfg, dev)
            host_s += dec.last_timing.get("host_polygons_s", 0.0)
        torch.cuda.synchronize(dev)
        te = idist.reduce_max(time.perf_counter() - t0, dev)
        d2h = int(dec.last_timing.get("d2h_bytes", 0))
        full = sum(pinned[k].numel() * 4 for k in ("kp", "ae", "regression", "classification"))
        uploaded = int(dec.last_timing.get("h2d_bytes", full))
        gathered = 16 * int(counts.sum()) + 16 * int(n_cand.sum()) if uploaded < full else 0
        # the upload alone, for the split reported next to the e2e number
        hc0, hc1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        keys = ("kp", "classification") if uploaded < full else ("kp", "ae", "regression", "classification")
        _dst = [torch.empty_like(pinned[k], device=dev) for k in keys]
        torch.cuda.synchronize(dev)
        hc0.record()
        for _d, k in zip(_dst, keys):
            _d.copy_(pinned[k], non_blocking=True)
        hc1.record()
        torch.cuda.synchronize(dev)
        h2d_ms = hc0.elapsed_time(hc1)
        del _dst
        n_inst = sum(len(r) for r in res)
        e2e = {"value": n_img_total * H * W * args.e2e_steps / te / 1e6, "unit": "Mpix/s", "h2d_bytes_per_step": uploaded + gathered,
               "d2h_bytes_per_step": d2h, "steps": args.e2e_steps, "ms_per_step": 1e3 * te / args.e2e_steps,
               "instances_per_step": n_inst, "h2d_ms_per_step": h2d_ms, "host_assembly_ms_per_step": 1e3 * host_s / args.e2e_steps,
               "uploaded_bytes_per_step": uploaded, "gathered_over_pcie_bytes_per_step": gathered, "model_output_bytes_per_step": full,
               "polygons": "device (isg_instance_polygons); the host only slices the read-back buffers into the result lists",
               "h2d": ("pinned host tensors: kp and classification are uploaded in chunks of %d images; ae is gathered at the keep pixels and "
                       "regression at the candidate anchors by the kernels out of the pinned buffers (zero-copy)" % dec.host_chunk_images)
                      if uploaded < full else "pinned host tensors, uploaded in chunks of %d images" % dec.host_chunk_images}

    if rank != 0:
        return
    peak, peak_src = measured_peak()
    # ISG_SPLIT_KEEP=1 (labels-only dense kernel, keep bits from the top-k candidates): kp is not read by this kernel
    algo = (ALGO_BYTES_PER_PIXEL - 4 if engine.SPLIT_KEEP else ALGO_BYTES_PER_PIXEL) * B * H * W
    achieved = algo / (kern_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": ("dense_v4_kernel, labels-only form (isg_assign_labels, 20 B/pixel; keep bits from isg_topk_keep)" if engine.SPLIT_KEEP else
                           "dense_v4_kernel (isg_assign_dense; its tile lists are prebuilt on the box branch)"),
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic(args.workload + ("" if args.inputs == "default" else ":" + args.inputs)),
                "peak_source": peak_src, "kernel_ms": kern_ms, "algorithmic_bytes_per_launch": algo,
                "kernel_timing": "CUDA events on the launching stream around the kernel, %d isolated launches after the timed region" % KERNEL_TIMING_LAUNCHES,
                "kernel_ms_in_flight": float(np.mean(in_flight)) if in_flight else None,
                "in_flight_note": "same events on every 8th step of a ring pass: kernels of the neighbouring steps share the SMs and the HBM with it"}
    cpu = None
    if not args.no_cpu and world == 1:      # the CPU baseline is taken on rank 0 at N=1 only
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        n_img = max(1, min(args.cpu_images, B))
        cpu_decode_images(host, wl, 1)
        dt, _, kind, split = cpu_decode_images(host, wl, n_img)
        cpu = {"value": n_img * H * W / dt / 1e6, "unit": "Mpix/s", "cores": torch.get_num_threads(), "kind": kind,
               "sample": "%d of the %d images of one step, once (after a 1-image warm-up): %s" % (
                   n_img, B, "the unmodified reference's decode_output (baseline/_ref)" if kind == "reference" else "oracle port of decode_output"),
               "split_ms_per_image": None if split is None else {k: v / n_img for k, v in split.items()}}
        if kind == "reference":               # the port next to it: the same algorithm with the Python loops vectorised
            from oracle import ref_decode as rd
            outs = ((host["kp"][:n_img], host["ae"][:n_img], None), host["regression"][:n_img], host["classification"][:n_img], host["anchors"])
            t0 = time.perf_counter()
            rd.decode_output(H, W, outs, kp_th=wl["kp_th"], cls_th=CLS_TH, iou_th=IOU_TH, wh_delta=WH_DELTA)
            cpu["port_value"] = n_img * H * W / (time.perf_counter() - t0) / 1e6
    config = workload_config(args, wl)
    stats = {"seeds_per_image": [int(v) for v in n_keep[:8]], "candidates_per_image": [int(v) for v in n_cand[:8]],
             "keep_pixels_per_image": [int(v) for v in counts[:8]], "mode": "dense",
             "pipelining": "steps go round-robin through %d independent pipelines (own plans / streams, one isg_decode_step call each); "
                           "neighbouring steps overlap, every step runs all of its kernels" % len(ring.pipes),
             "isolated_step_ms": iso_ms, "timed_region_s": t_max * 1e-3}
    line = {"metric": "decoded Mpix/s", "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": total_steps,
            "warmup": max(args.warmup, 3), "ms_per_step": t_max / total_steps, "higher_is_better": True,
            "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "workload_stats": stats, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks.summary()}
    if use_kmeans:
        line["kmeans"] = {"iterations_per_image": float(np.mean(km_state["iters"])), "calls": len(km_state["iters"]),
                          "points_per_image": [int(v) for v in counts], "clusters_per_image": [int(v) for v in n_keep]}
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
        raise SystemExit("bench.py needs a CUDA device (there is no CPU path); use --impl reference for the CPU reference")
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
        if wl.get("kind") == "mask_nms":
            run_mask_nms(args, wl, rank, world, local_rank)
        else:
            run_ours(args, wl, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
