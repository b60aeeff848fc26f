"""Batched device pipeline behind the drop-in modules.

PyTorch is used for device memory, streams and host<->device copies only; every computation is a
libisg.so kernel enqueued on torch's current stream.  A plan owns all outputs and workspaces for a
fixed (batch, height, width, max seeds) shape so that a decode step allocates nothing and never
synchronises with the host; results are read back once per batch.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import _lib
from ._lib import call

GRID_H, GRID_W = 1024, 2048   # utils/utils.py:453-458 of the reference


def require_cuda(device) -> torch.device:
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("isg_b200 has no CPU path: pass a CUDA device (got %s)" % device)
    if not torch.cuda.is_available():
        raise RuntimeError("isg_b200 needs a CUDA device (B200, sm_100a); none is visible")
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    return device


_device_checked = set()


def check_device(device: torch.device) -> None:
    if device.index in _device_checked:
        return
    if not _lib.lib().isg_device_supported(device.index):
        raise RuntimeError("libisg.so is built for sm_100a (B200) only; device %s is not supported" % device)
    _device_checked.add(device.index)


def stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


def device_address(t: torch.Tensor) -> int:
    """Address a kernel can dereference: the data pointer of a CUDA tensor, or the mapped device address of a PINNED
    host tensor (isg_host_device_pointer).  Pageable host memory raises."""
    if t.is_cuda:
        return t.data_ptr()
    if not t.is_pinned():
        raise RuntimeError("a host tensor handed to a kernel must be pinned (page-locked, mapped)")
    import ctypes
    out = ctypes.c_void_p()
    rc = _lib.lib().isg_host_device_pointer(t.data_ptr(), ctypes.byref(out))
    if rc != 0:
        raise _lib.IsgError(rc, "isg_host_device_pointer")
    return int(out.value)


_tables = {}


def coordinate_tables(H: int, W: int, device: torch.device):
    """ys [H], xs [W] on `device`, the slices of the reference grid used at utils/decode.py:304."""
    if H > GRID_H or W > GRID_W:
        raise RuntimeError("the reference coordinate grid is 1024x2048 (utils/utils.py:453-458); got %dx%d" % (H, W))
    key = (H, W, device.index)
    if key not in _tables:
        ys = torch.linspace(0, 1, GRID_H)[:H].contiguous().to(device)
        xs = torch.linspace(0, 2, GRID_W)[:W].contiguous().to(device)
        _tables[key] = (ys, xs)
    return _tables[key]


def _rows_contiguous(t: torch.Tensor) -> bool:
    return t.stride(-1) == 1 and t.stride(-2) == t.shape[-1]


def as_f32_planes(t: torch.Tensor, device: torch.device) -> torch.Tensor:
    """fp32 tensor on `device` whose last two dims are contiguous rows (copies only when needed)."""
    if t.device != device:
        t = t.to(device, non_blocking=True)
    if t.dtype != torch.float32:
        t = t.float()
    if not _rows_contiguous(t):
        t = t.contiguous()
    return t


def popcount_rows(bits: torch.Tensor):
    """number of set bits per leading index of a bit-packed tensor (host side; diagnostics only)"""
    a = bits.detach().cpu().numpy()
    return np.unpackbits(a.reshape(a.shape[0], -1).view(np.uint8), axis=1).sum(axis=1)


# Box NMS of the decode (torchvision.ops.batched_nms at utils/decode.py:400).  ISG_NMS_TV_BATCHED reproduces the batched_nms
# of current torchvision on CPU tensors (coordinate trick for <= 1000 candidates, per-class NMS above) - what the reference
# computes in this image; ISG_NMS_TV_TRICK is torchvision 0.5.0 (the reference's pin); ISG_NMS_TV_GT never shifts the boxes.
BOX_NMS_CONVENTION = _lib.ISG_NMS_TV_BATCHED
# Dense steps of isg_decode_step: threshold and keep bits from the top-k candidate list in one go (isg_topk_keep) + the labels-only dense
# kernel (isg_assign_labels), so that kp is read from HBM once per step; False (default) = the fused form
# (isg_assign_dense).  Both give bit-identical label maps and keep bits.  Measured: the labels-only kernel
# takes 56.9 us instead of 67.6, but the scattered neighbour loads of the peak test at the ~20 k selected pixels per image
# cost the select kernel the same 10 us - no net gain, so the fused form stays the default.
SPLIT_KEEP = os.environ.get("ISG_SPLIT_KEEP", "0") == "1"


def aligned_workspace(nbytes: int, device: torch.device):
    """(tensor, 256-byte aligned device pointer) of at least `nbytes`"""
    t = torch.empty(int(nbytes) + 256, dtype=torch.uint8, device=device)
    return t, t.data_ptr() + ((-t.data_ptr()) % 256)


class Arena:
    """One device buffer + a pinned host mirror for everything the host reads back after a step (counts, detection
    tables, per-instance polygon tables and points): the read-back is ONE device->host copy and one wait instead of a
    blocking `.cpu()` per buffer.  Plans take their host-visible outputs from an arena when they are given one."""

    def __init__(self, nbytes: int, device):
        self.device = require_cuda(device)
        self.nbytes = (int(nbytes) + 255) // 256 * 256
        self.dev = torch.empty(self.nbytes, dtype=torch.uint8, device=self.device)
        self.host = torch.empty(self.nbytes, dtype=torch.uint8).pin_memory()
        self.used = 0
        self.ready = torch.cuda.Event()

    def alloc(self, shape, dtype):
        """(device tensor, host mirror tensor) of `shape`/`dtype` carved from the arena (256-byte aligned)"""
        n = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        off = (self.used + 255) // 256 * 256
        if off + n > self.nbytes:
            raise RuntimeError("arena too small: %d + %d > %d" % (off, n, self.nbytes))
        self.used = off + n
        return (self.dev[off:off + n].view(dtype).view(shape), self.host[off:off + n].view(dtype).view(shape))

    def fetch_async(self) -> None:
        """enqueue the device->host copy of the used part on the current stream"""
        self.host[:self.used].copy_(self.dev[:self.used], non_blocking=True)
        self.ready.record(torch.cuda.current_stream(self.device))

    def wait(self) -> int:
        self.ready.synchronize()
        return self.used

    @staticmethod
    def bytes_for(B: int, N: int, cap: int) -> int:
        return B * (cap * 8 + N * 64) + 64 * 256


class DecodePlan:
    """select -> assign (dense or sparse) -> compact -> group for a batch of B images of HxW."""

    def __init__(self, B: int, H: int, W: int, max_seeds: int, kp_th: int, device, mode: str = "sparse",
                 want_score: bool = True, wh_delta: float = 0.1, scale: float = 1.0, fused_stats: bool = False,
                 min_cap: int = 0, arena: Arena | None = None):
        if mode not in ("dense", "sparse"):
            raise ValueError("mode must be 'dense' or 'sparse'")
        self.device = require_cuda(device)
        check_device(self.device)
        self.B, self.H, self.W, self.N = int(B), int(H), int(W), max(int(max_seeds), 1)
        self.kp_th = int(kp_th)
        if self.kp_th > H * W:
            raise RuntimeError("selected index k out of range")   # torch.topk, utils/decode.py:81
        # room for the keep pixels of one image: k, unless the caller saw more (ties at the k-th value select every tied
        # pixel, so a plateau can exceed k; the drop-in re-plans with the reported count, see utils/decode.py)
        self.cap = min(max(min(self.kp_th, H * W), int(min_cap), 1), H * W)
        self.mode, self.want_score = mode, bool(want_score)
        # dense mode: per-instance count/bbox either inside the fused kernel (one atomic set per keep pixel in the hot
        # loop) or in the gather pass over the compacted keep pixels (default: same numbers, cheaper)
        self.fused_stats = bool(fused_stats)
        # wh_delta None disables the device ghost filter (non-identity val transforms filter on the host)
        self.ghost_k = -1.0 if wh_delta is None else float(np.float32(0.5 + wh_delta))
        self.scale = float(scale)
        self.Ww = (W + 31) // 32
        self.events = []
        d, i32, f32 = self.device, torch.int32, torch.float32
        B, N, cap = self.B, self.N, self.cap
        self.arena, self.host = arena, {}

        def out(name, shape, dtype):
            """a buffer the host reads back: from the arena (with a pinned mirror in self.host) when there is one"""
            if arena is None:
                return torch.empty(shape, dtype=dtype, device=d)
            t, h = arena.alloc(shape, dtype)
            self.host[name] = h
            return t
        self.ys, self.xs = coordinate_tables(H, W, d)
        self.thr_key = torch.empty(B, dtype=i32, device=d)
        self.keepbits = torch.empty((B, H, self.Ww), dtype=i32, device=d)
        self.idx = torch.empty((B, cap, 2), dtype=i32, device=d)
        self.count = out("count", (B,), i32)
        self.label = torch.empty((B, cap), dtype=i32, device=d)
        self.score = torch.empty((B, cap), dtype=f32, device=d) if want_score else None
        self.flag = torch.empty((B, cap), dtype=torch.uint8, device=d)
        self.stats = torch.empty((B, N, _lib.STAT_WORDS), dtype=i32, device=d)
        self.seeds = torch.empty((B, N, _lib.SEED_WORDS), dtype=i32, device=d)
        self.ghost = torch.empty((B, N, _lib.GHOST_WORDS), dtype=f32, device=d)
        self.offsets = torch.empty((B, N + 1), dtype=i32, device=d)
        self.points = torch.empty((B, cap, 2), dtype=f32, device=d)
        self.ws_bytes = int(_lib.lib().isg_topk_workspace_bytes(B, H, W, self.kp_th))
        self.ws, self.ws_ptr = aligned_workspace(self.ws_bytes, d)
        if mode == "dense":
            # device polygon stage (isg_instance_polygons)
            self.img_total = out("img_total", (B,), i32)
            self.inst_start = out("inst_start", (B, N), i32)
            self.inst_count = out("inst_count", (B, N), i32)
            self.inst_flags = out("inst_flags", (B, N), torch.uint8)
            self.poly_points = out("poly_points", (B, cap, 2), f32)
            self.inst_internal = torch.empty((B, N, 2), dtype=f32, device=d)
            self.poly_ws_bytes = int(_lib.lib().isg_instance_polygons_workspace_bytes(B, cap))
            self.poly_ws, self.poly_ws_ptr = aligned_workspace(self.poly_ws_bytes, d)
            self.label_map = torch.empty((B, H, W), dtype=i32, device=d)
            # tile scheduler words + per-tile seed lists of the dense kernel
            self.dense_ws_bytes = int(_lib.lib().isg_assign_dense_workspace_bytes(B, N, H, W))
            self.dense_ws = torch.empty(self.dense_ws_bytes, dtype=torch.uint8, device=d)
            self.score_map = torch.empty((B, H, W), dtype=f32, device=d) if want_score else None
        else:
            self.label_map = self.score_map = None

    # ------------------------------------------------------------------------------------------
    def _check(self, kp, ae):
        B, H, W = self.B, self.H, self.W
        if kp.dim() == 4:
            kp = kp[:, 0]
        assert kp.shape == (B, H, W), kp.shape
        assert kp.dtype == torch.float32 and _rows_contiguous(kp)
        if ae is not None:
            assert ae.shape == (B, 4, H, W) and ae.dtype == torch.float32 and _rows_contiguous(ae), ae.shape
        return kp

    def run_topk(self, kp: torch.Tensor) -> None:
        """Stage 1 (independent of the boxes): exact k-th largest kp value per image -> self.thr_key."""
        kp = self._check(kp, None)
        B, H, W = self.B, self.H, self.W
        kp_stride = kp.stride(0) if B > 1 else H * W
        call("isg_topk_threshold", ptr(kp), B, H, W, kp_stride, self.kp_th, ptr(self.thr_key), self.ws_ptr,
             self.ws_bytes, stream_ptr(self.device))

    def run_assign(self, kp: torch.Tensor, ae: torch.Tensor, rois: torch.Tensor, n_seeds: torch.Tensor,
                   layout: int = _lib.ISG_BOX_XYXY, time_main: bool = False, tail: str = "lists",
                   obj_pixel_th: int = 0, seeds_ready: bool = False, lists_ready: bool = False) -> None:
        """Stage 2: seeds -> assignment (dense or sparse) -> tail.  Needs self.thr_key.
        tail "lists": compaction + per-pixel labels + per-instance point sets (idx/label/flag/offsets/points);
        tail "polygons" (dense mode, XYXY rois): isg_instance_polygons straight from the label map
        (poly_points/inst_start/inst_count/inst_flags)."""
        if tail not in ("lists", "polygons", "defer"):
            raise ValueError("tail must be 'lists', 'polygons' or 'defer' (the caller enqueues run_polygons itself)")
        if tail in ("polygons", "defer") and (self.mode != "dense" or self.ghost_k < 0):
            raise ValueError("the device polygon stage needs dense mode and the device ghost filter")
        B, H, W, N, cap = self.B, self.H, self.W, self.N, self.cap
        kp = self._check(kp, ae)
        assert rois.shape == (B, N, 4) and rois.is_contiguous() and rois.dtype == torch.float32
        assert n_seeds.shape == (B,) and n_seeds.dtype == torch.int32
        s = stream_ptr(self.device)
        kp_stride = kp.stride(0) if B > 1 else H * W
        ae_img = ae.stride(0) if B > 1 else 4 * H * W
        ae_plane = ae.stride(1)
        if not seeds_ready:   # the pipeline builds seeds / ghost / stats together with the detection tables
            call("isg_build_seeds", ptr(rois), layout, ptr(n_seeds), B, N, ptr(self.ys), ptr(self.xs), H, W, self.ghost_k,
                 self.scale, ptr(self.seeds), ptr(self.ghost), s)
            call("isg_stats_init", ptr(self.stats), B, N, s)
        ev = None
        if time_main:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        if self.mode == "dense":
            if ev:
                ev[0].record()
            call("isg_assign_dense", ptr(kp), kp_stride, ptr(ae), ae_img, ae_plane, ptr(self.thr_key), ptr(self.seeds),
                 ptr(self.ghost), ptr(n_seeds), B, N, H, W, ptr(self.ys), ptr(self.xs), ptr(self.label_map),
                 ptr(self.score_map), ptr(self.keepbits), ptr(self.stats) if self.fused_stats else 0, ptr(self.dense_ws),
                 self.dense_ws_bytes, 1 if lists_ready else 0, s)
            if ev:
                ev[1].record()
            if tail in ("polygons", "defer"):
                if tail == "polygons":
                    self.run_polygons(rois, n_seeds, layout, obj_pixel_th, totals_zeroed=seeds_ready)
                if ev:
                    self.events.append(ev)
                return
            call("isg_compact_points", ptr(self.keepbits), B, H, W, cap, ptr(self.idx), ptr(self.count), s)
            call("isg_gather_labels", ptr(self.label_map), ptr(self.score_map), ptr(self.idx), ptr(self.count), cap,
                 ptr(self.ghost), B, N, H, W, ptr(self.label), ptr(self.score), ptr(self.flag),
                 0 if self.fused_stats else ptr(self.stats), s)
        else:
            call("isg_keep_points", ptr(kp), B, H, W, kp_stride, ptr(self.thr_key), ptr(self.keepbits), 0, s)
            call("isg_compact_points", ptr(self.keepbits), B, H, W, cap, ptr(self.idx), ptr(self.count), s)
            if ev:
                ev[0].record()
            call("isg_assign_sparse", ptr(ae), ae_img, ae_plane, ptr(self.idx), ptr(self.count), cap, ptr(self.seeds),
                 ptr(self.ghost), ptr(n_seeds), B, N, H, W, ptr(self.ys), ptr(self.xs), ptr(self.label),
                 ptr(self.score), ptr(self.flag), ptr(self.stats), s)
            if ev:
                ev[1].record()
        call("isg_group_points", ptr(self.idx), ptr(self.label), ptr(self.flag), ptr(self.count), cap, ptr(n_seeds),
             B, N, ptr(self.offsets), ptr(self.points), s)
        if ev:
            self.events.append(ev)

    def run_polygons(self, rois: torch.Tensor, n_seeds: torch.Tensor, layout: int = _lib.ISG_BOX_XYXY, obj_pixel_th: int = 0,
                     totals_zeroed: bool = False) -> None:
        """The polygon tail of a dense step (isg_instance_polygons) on the current stream."""
        call("isg_instance_polygons", ptr(self.keepbits), ptr(self.label_map), ptr(rois), layout, ptr(self.ghost), ptr(n_seeds),
             self.B, self.N, self.H, self.W, self.cap, int(obj_pixel_th), ptr(self.poly_points), ptr(self.inst_start),
             ptr(self.inst_count), ptr(self.inst_flags), ptr(self.inst_internal), ptr(self.img_total),
             0 if self.fused_stats else ptr(self.stats), self.poly_ws_ptr, self.poly_ws_bytes,
             1 if totals_zeroed else 0, stream_ptr(self.device))

    def run(self, kp: torch.Tensor, ae: torch.Tensor, rois: torch.Tensor, n_seeds: torch.Tensor,
            layout: int = _lib.ISG_BOX_XYXY, time_main: bool = False, tail: str = "lists", obj_pixel_th: int = 0) -> None:
        """kp [B,1,H,W] or [B,H,W]; ae [B,4,H,W]; rois [B,N,4] fp32 ((x1,y1,x2,y2) or (cy,cx,h,w), see `layout`);
        n_seeds [B] int32 — all on the plan's device.  Enqueues the kernels on the current stream and returns
        without synchronising.  time_main=True brackets the assignment kernel with CUDA events on the launching
        stream and appends the pair to self.events (bench.py's roofline measurement)."""
        self.run_topk(kp)
        self.run_assign(kp, ae, rois, n_seeds, layout, time_main, tail, obj_pixel_th)


class BoxPlan:
    """decode_boxes on the device: front-end -> class-aware NMS -> per-image detection tables."""

    def __init__(self, B: int, A: int, C: int, H: int, W: int, device, cap: int = 4096, max_keep: int = 1024,
                 arena: Arena | None = None):
        self.device = require_cuda(device)
        check_device(self.device)
        self.B, self.A, self.C, self.H, self.W = int(B), int(A), int(C), int(H), int(W)
        self.cap = int(min(cap, A))
        self.N = int(min(max_keep, self.cap))
        d, i32, f32 = self.device, torch.int32, torch.float32
        B, cap, N = self.B, self.cap, self.N
        self.arena, self.host = arena, {}

        def out(name, shape, dtype):
            if arena is None:
                return torch.empty(shape, dtype=dtype, device=d)
            t, h = arena.alloc(shape, dtype)
            self.host[name] = h
            return t
        self.cand_boxes = torch.empty((B, cap, 4), dtype=f32, device=d)
        self.cand_scores = torch.empty((B, cap), dtype=f32, device=d)
        self.cand_cls = torch.empty((B, cap), dtype=i32, device=d)
        self.cand_anchor = torch.empty((B, cap), dtype=i32, device=d)
        self.cand_count = out("cand_count", (B,), i32)
        self.keep = torch.empty((B, cap), dtype=i32, device=d)
        self.n_keep = out("n_keep", (B,), i32)
        self.rois = out("rois", (B, N, 4), f32)
        self.scores = out("scores", (B, N), f32)
        self.cls = out("cls", (B, N), i32)
        self.n_seeds = torch.empty(B, dtype=i32, device=d)
        self.ws_bytes = int(_lib.lib().isg_box_nms_workspace_bytes(B, cap))
        self.ws, self.ws_ptr = aligned_workspace(self.ws_bytes, d)

    def run(self, anchors: torch.Tensor, regression: torch.Tensor, classification: torch.Tensor, cls_th: float,
            iou_th: float, gather: bool = True) -> None:
        """gather=False leaves the kept candidates in (keep, n_keep): the pipeline gathers them together with the seed
        records of the decode (isg_gather_build_seeds)"""
        B, A, C = self.B, self.A, self.C
        assert anchors.numel() == A * 4 and regression.shape == (B, A, 4) and classification.shape == (B, A, C)
        assert anchors.is_contiguous() and regression.is_contiguous() and classification.is_contiguous()
        s = stream_ptr(self.device)
        call("isg_decode_boxes", ptr(anchors), ptr(regression), ptr(classification), B, A, C, self.H, self.W,
             float(np.float32(cls_th)), self.cap, ptr(self.cand_boxes), ptr(self.cand_scores), ptr(self.cand_cls),
             ptr(self.cand_anchor), ptr(self.cand_count), s)
        call("isg_box_nms", ptr(self.cand_boxes), ptr(self.cand_scores), ptr(self.cand_cls), ptr(self.cand_anchor),
             ptr(self.cand_count), B, self.cap, float(iou_th), BOX_NMS_CONVENTION, ptr(self.keep), ptr(self.n_keep),
             self.ws_ptr, self.ws_bytes, s)
        if gather:
            call("isg_gather_kept", ptr(self.cand_boxes), ptr(self.cand_scores), ptr(self.cand_cls), ptr(self.keep),
                 ptr(self.n_keep), B, self.cap, self.N, ptr(self.rois), ptr(self.scores), ptr(self.cls), ptr(self.n_seeds), s)


_plans = {}


def get_decode_plan(B, H, W, max_seeds, kp_th, device, mode, want_score=True, wh_delta=0.1, scale=1.0, min_cap=0) -> DecodePlan:
    device = require_cuda(device)
    key = ("d", B, H, W, max_seeds, kp_th, device.index, mode, want_score, wh_delta, float(scale), int(min_cap))
    if key not in _plans:
        if len(_plans) > 8:
            _plans.clear()
        _plans[key] = DecodePlan(B, H, W, max_seeds, kp_th, device, mode, want_score, wh_delta, scale, min_cap=min_cap)
    return _plans[key]


def get_box_plan(B, A, C, H, W, device, cap=4096, max_keep=1024) -> BoxPlan:
    device = require_cuda(device)
    key = ("b", B, A, C, H, W, device.index, cap, max_keep)
    if key not in _plans:
        if len(_plans) > 8:
            _plans.clear()
        _plans[key] = BoxPlan(B, A, C, H, W, device, cap, max_keep)
    return _plans[key]


class DecodePipeline:
    """box head + decode for one batch.  The top-k threshold does not depend on the boxes, so it runs on a side
    stream concurrently with the box head / NMS branch and joins before the assignment kernel."""

    def __init__(self, bplan: BoxPlan, dplan: DecodePlan):
        assert bplan.device == dplan.device and bplan.B == dplan.B and bplan.N == dplan.N
        self.bplan, self.dplan, self.device = bplan, dplan, dplan.device
        self.side = torch.cuda.Stream(device=self.device)
        self.fork = torch.cuda.Event()
        self.join = torch.cuda.Event()
        # pipelined steps: the polygon tail of step s runs on its own stream while step s+1's box head / NMS / top-k run
        self.tail_stream = torch.cuda.Stream(device=self.device)
        self.dense_done = torch.cuda.Event()
        self.tail_done = torch.cuda.Event()
        self.tail_pending = False
        self.tail_deferred = None      # (obj_pixel_th,) of a pipelined step whose polygon tail is not enqueued yet
        self._step = None              # isg_decode_step_t with every plan pointer filled in (run_native)
        self.time_events = None        # (begin, end) CUDA events around the assignment of the last timed native step

    def _native_step(self) -> "_lib.DecodeStep":
        if self._step is not None:
            return self._step
        bp, dp = self.bplan, self.dplan
        if dp.mode != "dense" or dp.ghost_k < 0:
            raise ValueError("the native step needs a dense-mode plan with the device ghost filter")
        st = _lib.DecodeStep()
        st.struct_bytes = int(_lib.lib().isg_decode_step_bytes())
        st.B, st.H, st.W, st.img_h, st.img_w, st.A, st.C = dp.B, dp.H, dp.W, bp.H, bp.W, bp.A, bp.C
        st.Nmax, st.cand_cap, st.cap, st.kp_th = dp.N, bp.cap, dp.cap, dp.kp_th
        st.ghost_k, st.scale = dp.ghost_k, dp.scale
        st.ys, st.xs = ptr(dp.ys), ptr(dp.xs)
        for name in ("cand_boxes", "cand_scores", "cand_cls", "cand_anchor", "cand_count", "keep", "n_keep", "rois", "scores",
                     "cls", "n_seeds"):
            setattr(st, name, ptr(getattr(bp, name)))
        st.nms_ws, st.nms_ws_bytes = bp.ws_ptr, bp.ws_bytes
        for name in ("thr_key", "seeds", "ghost", "stats", "keepbits", "label_map", "idx", "count", "label", "poly_points",
                     "inst_start", "inst_count", "inst_flags", "inst_internal", "img_total"):
            setattr(st, name, ptr(getattr(dp, name)))
        st.topk_ws, st.topk_ws_bytes = dp.ws_ptr, dp.ws_bytes
        st.dense_ws, st.dense_ws_bytes = ptr(dp.dense_ws), dp.dense_ws_bytes
        st.poly_ws, st.poly_ws_bytes = dp.poly_ws_ptr, dp.poly_ws_bytes
        # torch creates the cudaEvent_t lazily: record once so that the handles exist
        cur = torch.cuda.current_stream(self.device)
        self._t0, self._t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for ev in (self.fork, self.join, self._t0, self._t1):
            ev.record(cur)
        st.side, st.fork_event, st.join_event = self.side.cuda_stream, self.fork.cuda_event, self.join.cuda_event
        self._step = st
        return st

    def run_native(self, kp, ae, anchors, regression, classification, cls_th, iou_th, obj_pixel_th: int = 0,
                   assign: str = "dense", polygons: bool = True, time_main: bool = False) -> None:
        """One whole step through isg_decode_step (a single host call) on the current stream + the side stream.
        assign "sparse": only the keep pixels are assigned (isg_assign_sparse + isg_scatter_labels); `ae` and `regression`
        may then be PINNED HOST tensors - they are gathered over PCIe at the keep pixels / candidate anchors instead of
        being uploaded.  Results as for run(tail="polygons")."""
        st = self._native_step()
        dp, bp = self.dplan, self.bplan
        B, H, W = dp.B, dp.H, dp.W
        if kp.dim() == 4:
            kp = kp[:, 0]
        assert kp.shape == (B, H, W) and kp.dtype == torch.float32 and _rows_contiguous(kp) and kp.is_cuda
        assert ae.shape == (B, 4, H, W) and ae.dtype == torch.float32 and _rows_contiguous(ae)
        assert regression.shape == (B, bp.A, 4) and classification.shape == (B, bp.A, bp.C) and anchors.numel() == bp.A * 4
        assert regression.is_contiguous() and classification.is_contiguous() and anchors.is_contiguous() and classification.is_cuda
        sparse = assign == "sparse"
        if not sparse and not (ae.is_cuda and regression.is_cuda):
            raise ValueError("host-resident ae / regression need assign='sparse'")
        st.assign = _lib.ISG_ASSIGN_SPARSE if sparse else _lib.ISG_ASSIGN_DENSE
        st.polygons, st.obj_pixel_th = 1 if polygons else 0, int(obj_pixel_th)
        st.nms_convention = BOX_NMS_CONVENTION
        st.split_keep = 1 if (SPLIT_KEEP and not sparse) else 0
        st.cls_th, st.iou_th = float(np.float32(cls_th)), float(iou_th)
        st.kp, st.kp_img_stride = ptr(kp), kp.stride(0) if B > 1 else H * W
        st.ae, st.ae_img_stride, st.ae_plane_stride = device_address(ae), ae.stride(0) if B > 1 else 4 * H * W, ae.stride(1)
        st.anchors, st.regression, st.classification = ptr(anchors), device_address(regression), ptr(classification)
        st.main = stream_ptr(self.device)
        if time_main:
            st.time_begin, st.time_end = self._t0.cuda_event, self._t1.cuda_event
        else:
            st.time_begin = st.time_end = None
        import ctypes
        rc = _lib.lib().isg_decode_step(ctypes.byref(st))
        if rc != 0:
            raise _lib.IsgError(rc, "isg_decode_step")
        _lib.launch_count += 8 + (1 if polygons else 0) + (2 if sparse else 0)
        if time_main:
            ev = (self._t0, self._t1)
            self._t0, self._t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self._t0.record(torch.cuda.current_stream(self.device)); self._t1.record(torch.cuda.current_stream(self.device))
            dp.events.append(ev)

    def _launch_tail(self) -> None:
        """Enqueue the polygon tail of the last pipelined step on the tail stream (behind its dense kernel)."""
        if self.tail_deferred is None:
            return
        (obj_pixel_th,) = self.tail_deferred
        self.tail_deferred = None
        bp, dp = self.bplan, self.dplan
        self.tail_stream.wait_event(self.dense_done)
        with torch.cuda.stream(self.tail_stream):
            dp.run_polygons(bp.rois, bp.n_seeds, _lib.ISG_BOX_XYXY, obj_pixel_th, totals_zeroed=True)
            self.tail_done.record(self.tail_stream)
        self.tail_pending = True

    def finish(self) -> None:
        """Make the current stream wait for the polygon tail of the last pipelined step (no-op otherwise)."""
        self._launch_tail()
        if self.tail_pending:
            torch.cuda.current_stream(self.device).wait_event(self.tail_done)
            self.tail_pending = False

    def run(self, kp, ae, anchors, regression, classification, cls_th, iou_th, time_main: bool = False,
            tail: str = "lists", obj_pixel_th: int = 0, pipelined: bool = False) -> None:
        """One decode step on the current stream.  pipelined=True (dense mode, tail "polygons"): the polygon tail is
        enqueued on a separate stream and the call returns without making the current stream wait for it, so that the
        box head, NMS and top-k of the NEXT run() overlap it; the kernels of the next step that overwrite what the tail
        reads (seeds, label map) wait for it.  Call finish() before reading the results on the current stream."""
        main = torch.cuda.current_stream(self.device)
        pipelined = pipelined and tail == "polygons" and self.dplan.mode == "dense"
        if not pipelined:
            self.finish()
        # the top-k of this step may start as soon as everything enqueued so far on `main` is done (in a pipelined
        # sequence: the dense kernel of the previous step, which reads the previous threshold)
        self.fork.record(main)
        self.side.wait_event(self.fork)
        with torch.cuda.stream(self.side):
            self.dplan.run_topk(kp)
            self.join.record(self.side)
        bp, dp = self.bplan, self.dplan
        bp.run(anchors, regression, classification, cls_th, iou_th, gather=False)
        # The previous step's polygon tail is enqueued only now, behind this step's box head in the hardware queues: the
        # box head is a short bandwidth-bound grid, NMS occupies one CTA per image, and the tail (whose grid does not fit
        # the GPU in one wave and would hold back every later grid until its last CTA is placed) fills the rest.
        self._launch_tail()
        if self.tail_pending:      # the seeds / label map about to be overwritten are still read by the previous tail
            main.wait_event(self.tail_done)
            self.tail_pending = False
        call("isg_gather_build_seeds", ptr(bp.cand_boxes), ptr(bp.cand_scores), ptr(bp.cand_cls), ptr(bp.keep), ptr(bp.n_keep),
             bp.B, bp.cap, bp.N, ptr(dp.ys), ptr(dp.xs), dp.H, dp.W, dp.ghost_k, dp.scale, ptr(bp.rois), ptr(bp.scores),
             ptr(bp.cls), ptr(bp.n_seeds), ptr(dp.seeds), ptr(dp.ghost), ptr(dp.stats),
             ptr(dp.img_total) if dp.mode == "dense" else 0, stream_ptr(self.device))
        dense = dp.mode == "dense"
        if dense:   # the tile lists of the dense kernel only need the seeds: build them before joining the top-k branch
            call("isg_build_tile_lists", ptr(dp.seeds), ptr(bp.n_seeds), dp.B, dp.N, dp.H, dp.W, ptr(dp.dense_ws),
                 dp.dense_ws_bytes, stream_ptr(self.device))
        main.wait_event(self.join)
        if not pipelined:
            dp.run_assign(kp, ae, bp.rois, bp.n_seeds, _lib.ISG_BOX_XYXY, time_main, tail, obj_pixel_th, seeds_ready=True,
                          lists_ready=dense)
            return
        dp.run_assign(kp, ae, bp.rois, bp.n_seeds, _lib.ISG_BOX_XYXY, time_main, "defer", obj_pixel_th, seeds_ready=True,
                      lists_ready=True)
        self.dense_done.record(main)
        self.tail_deferred = (obj_pixel_th,)


class DecodeRing:
    """n independent pipelines (own plans, arenas and streams) used round-robin.  Steps submitted back to back overlap:
    the box head / NMS / top-k of the next steps and the polygon stage of the previous ones fill the machine around
    each step's dense kernel, which runs one CTA per SM.  Every step still executes all of its kernels
    on its own inputs; results of slot i are valid after wait(i)."""

    def __init__(self, make_pipeline, n: int = 4):
        self.pipes = [make_pipeline() for _ in range(n)]
        self.device = self.pipes[0].device
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(n)]
        self.done = [torch.cuda.Event() for _ in range(n)]
        self.ready = torch.cuda.Event()
        self.next = 0

    def submit(self, *args, fetch: bool = False, **kw) -> int:
        """run_native(*args, **kw) on the next slot's stream, ordered behind the caller's current stream (the inputs).
        fetch=True also enqueues the read-back of the slot's arena.  Returns the slot."""
        i = self.next
        self.next = (i + 1) % len(self.pipes)
        s = self.streams[i]
        self.ready.record(torch.cuda.current_stream(self.device))
        s.wait_event(self.ready)
        with torch.cuda.stream(s):
            self.pipes[i].run_native(*args, **kw)
            if fetch:
                self.pipes[i].bplan.arena.fetch_async()
            self.done[i].record(s)
        return i

    def wait(self, i: int | None = None) -> None:
        """make the caller's current stream wait for slot i (default: every slot)"""
        cur = torch.cuda.current_stream(self.device)
        for j in (range(len(self.pipes)) if i is None else (i,)):
            cur.wait_event(self.done[j])


def make_pipeline(B, A, C, H, W, img_h, img_w, kp_th, device, cand_cap=1024, max_keep=256, min_cap=0, wh_delta=0.1,
                  scale=1.0) -> "DecodePipeline":
    """box plan + dense-mode decode plan + pipeline sharing one read-back arena"""
    device = require_cuda(device)
    N = int(min(max_keep, min(cand_cap, A)))
    cap = min(max(min(int(kp_th), H * W), int(min_cap), 1), H * W)
    arena = Arena(Arena.bytes_for(B, N, cap), device)
    bplan = BoxPlan(B, A, C, img_h, img_w, device, cand_cap, max_keep, arena=arena)
    dplan = DecodePlan(B, H, W, bplan.N, kp_th, device, "dense", want_score=False, wh_delta=wh_delta, scale=scale,
                       min_cap=min_cap, arena=arena)
    return DecodePipeline(bplan, dplan)


_pipes = {}


def get_pipeline(bplan: BoxPlan, dplan: DecodePlan) -> DecodePipeline:
    key = (id(bplan), id(dplan))
    if key not in _pipes:
        if len(_pipes) > 8:
            _pipes.clear()
        _pipes[key] = DecodePipeline(bplan, dplan)
    return _pipes[key]
