"""Drop-in for the reference's utils/decode.py: same function names, argument order and return structures,
device work done by libisg.so (sm_100a CUDA) instead of torch/numpy ops.

Differences that are deliberate:
  * there is no CPU path: tensors are moved to the CUDA `device` argument; a CPU `device` raises;
  * the network outputs are not modified (the reference overwrites hm_ae[0:2] in place at :305 and
    clips the transformed anchors in place at utils/utils.py:357-361; no caller reads them afterwards);
  * the batch is decoded in one set of launches instead of the serial multi_apply map (:458);
  * drawing (decode_cfg.draw_flag) is a host-side debug aid and delegates to the reference's
    utils.visualize when that module is importable, otherwise it is skipped with a warning.
All file:line citations refer to the reference's utils/decode.py unless stated.
"""
from __future__ import annotations

import math
import os
import warnings
from typing import Iterable

import numpy as np
import torch

from .. import _lib, engine
from .._lib import call
from ..engine import ptr, require_cuda, stream_ptr
from . import image, parell_util
from .kmeans import kmeans            # noqa: F401  (imported, never called — like the reference, :16)
from .nms import py_cpu_nms
from .utils import BBoxTransform, ClipBoxes, generate_coordinates   # noqa: F401

base_dir = r""        # :28
target_size = 1       # :29
device = None         # set by test.py:134 / evaluate.py:40
draw_flag = False     # set by evaluate.py:41 (the code reads decode_cfg.draw_flag, like the reference)

# "dense": fused label map for every pixel (isg_assign_dense); "sparse": only the selected pixels, the
# reference's own amount of work (isg_assign_sparse).  Both give identical detections.
decode_mode = os.environ.get("ISG_DECODE_MODE", "dense")
# dense mode, identity val-transform, no drawing: run the per-instance polygon stage on the device
# (isg_instance_polygons); False keeps it on the host (cv2/numpy, the reference's own calls)
device_polygon_stage = os.environ.get("ISG_DEVICE_POLYGONS", "1") != "0"

# wall-clock split of the last decode_output call (seconds): device (H2D + kernels, until the results are on the
# host) and host (polygon stage).  Diagnostic only.
last_timing = {}

_xym = None


def __getattr__(name):
    # `xym` (:31) is a 16 MiB CPU tensor the reference builds at import; built lazily here.
    global _xym
    if name == "xym":
        if _xym is None:
            _xym = generate_coordinates()
        return _xym
    raise AttributeError(name)


def compute_scale(info):            # :34-35
    return target_size


def to_numpy(tensor):               # :38-39
    return tensor.cpu().numpy()


def _device_of(t: torch.Tensor, fallback=None) -> torch.device:
    if t.is_cuda:
        return t.device
    if fallback is not None:
        return require_cuda(fallback)
    if device is not None:
        return require_cuda(device)
    raise RuntimeError("isg_b200 has no CPU path: pass CUDA tensors or set decode.device to a CUDA device")


# ----------------------------------------------------------------------------------------------
# K2
# ----------------------------------------------------------------------------------------------
def nms_hm(heat, kernel=3):
    """:42-48 — uint8 mask of the pixels equal to their kernel x kernel maximum."""
    dev = _device_of(heat)
    h = engine.as_f32_planes(heat, dev)
    if h.dim() < 2:
        raise ValueError("heat must have at least 2 dims")
    H, W = h.shape[-2], h.shape[-1]
    h = h.contiguous()
    planes = h.numel() // (H * W)
    keep = torch.empty(h.shape, dtype=torch.uint8, device=dev)
    call("isg_nms_hm", ptr(h), planes, H, W, int(kernel), ptr(keep), stream_ptr(dev))
    return keep


def select_points(mat, k):
    """:71-85 — uint8 [H,W]: the k largest pixels that are 3x3 maxima of (mat where selected else 0)."""
    dev = _device_of(mat)
    m = engine.as_f32_planes(mat, dev)
    H, W = m.shape
    k = int(k)
    if k > H * W:
        raise RuntimeError("selected index k out of range")     # what torch.topk raises at :81
    lib = _lib.lib()
    ws_bytes = int(lib.isg_select_points_workspace_bytes(1, H, W, k))
    ws, ws_ptr = engine.aligned_workspace(ws_bytes, dev)
    keepbits = torch.empty((H, (W + 31) // 32), dtype=torch.int32, device=dev)
    mask = torch.empty((H, W), dtype=torch.uint8, device=dev)
    call("isg_select_points", ptr(m), 1, H, W, H * W, k, ptr(keepbits), ptr(mask), ws_ptr, ws_bytes, stream_ptr(dev))
    return mask


# ----------------------------------------------------------------------------------------------
# a7 — polygon stage (host glue this round; SURVEY.md §8 f1 moves it to the GPU)
# ----------------------------------------------------------------------------------------------
def find_internal_point(kps, default):
    """:51-68"""
    import cv2
    kps = np.array(kps)
    if cv2.pointPolygonTest(kps, tuple(default), False) > 0:
        return default
    mean = kps.mean(axis=0).reshape(-1)
    if cv2.pointPolygonTest(kps, tuple(mean), False) > 0:
        return mean
    n = kps.shape[0]
    for i in range(n):
        mids = (kps[i][None, :] + kps[1:]) / 2
        for point in mids:
            if cv2.pointPolygonTest(kps, tuple(point), False) > 0:
                return point
    return default


def cartesian2polar(kps, center_loc):
    """:88-113 — [K,2] float32 (theta in [0,2pi), distance); vectorised, same fp32 arithmetic."""
    d = (np.asarray(kps) - center_loc).astype(np.float32)
    dx, dy = d[:, 0], d[:, 1]
    with np.errstate(divide="ignore", invalid="ignore"):
        seta = np.arctan(dy / dx)
        seta = np.where(dx < 0, seta + np.float32(np.pi), seta)
        seta = np.where((dx > 0) & (dy < 0), seta + np.float32(2 * np.pi), seta)
        seta = np.where((dx == 0) & (dy > 0), np.float32(np.pi / 2), seta)
        seta = np.where((dx == 0) & (dy < 0), np.float32(3 * np.pi / 2), seta)
        dist = np.sqrt(dx ** 2 + dy ** 2)
    return np.stack((seta, dist), axis=1).astype(np.float32)


def polar2cartesian(kps, center_loc):
    """:116-128"""
    ang, dist = kps[:, 0], kps[:, 1]
    delta = np.hstack(((dist * np.cos(ang)).reshape(-1, 1), (dist * np.sin(ang)).reshape(-1, 1)))
    return delta + center_loc


def filter_ghost_polygons(polygons, center):
    """:131-141 (unused by the decode; kept for API parity — expects shapely-like polygons)."""
    import cv2
    if not isinstance(polygons, Iterable):
        polygons = [polygons]
    max_area, max_poly = 0, None
    for poly in polygons:
        np_poly = np.array(poly.exterior.coords).astype(np.float32)
        if poly.area > max_area and cv2.pointPolygonTest(np_poly, tuple(center), False) > 0:
            max_area, max_poly = poly.area, np_poly
    return max_poly


def smooth_polygon(polar_pts, sorted_inds, k=360):
    """:144-163 (unused by the decode; kept for API parity)."""
    d_seta = 2 * np.pi / 12
    selected, cur_ind, cur_dist, cur_bin = [], -1, -1, 0
    for ind in sorted_inds:
        index = math.floor(polar_pts[ind][0] / d_seta)
        if index != cur_bin:
            if cur_ind >= 0:
                selected.append(cur_ind)
            cur_ind, cur_dist, cur_bin = -1, -1, index
        elif polar_pts[ind][1] > cur_dist:
            cur_ind, cur_dist = ind, polar_pts[ind][1]
    if cur_ind >= 0:
        selected.append(cur_ind)
    return selected


def _polygon_area_is_zero(sorted_kp) -> bool:
    """`image.poly_to_mask(sorted_kp).sum() == 0` (:187-189) evaluated on the polygon's own bounding box:
    fillPoly on integer vertices is invariant under integer translation, so rasterising the polygon shifted
    to the origin gives the same pixel count without allocating a canvas as large as the image."""
    import cv2
    poly = sorted_kp.astype(np.int32)
    lo = poly.min(0)
    size = (poly.max(0) - lo + 1)[::-1]
    return not cv2.fillPoly(np.zeros(size, dtype=np.uint8), [poly - lo], 1).any()


def aug_group(pts, center_loc):
    """:167-204 — order the points by polar angle about an internal point; None if degenerate."""
    import cv2
    center_loc = center_loc.reshape(-1)
    internal_point = find_internal_point(pts, center_loc)
    polar_pts = cartesian2polar(pts, internal_point)
    sorted_inds = np.argsort(polar_pts[:, 0])
    sorted_kp = pts[sorted_inds]
    if _polygon_area_is_zero(sorted_kp):
        return None
    if cv2.pointPolygonTest(sorted_kp, tuple(center_loc), False) > 0:
        return sorted_kp
    return None


# ----------------------------------------------------------------------------------------------
# drawing helpers (:207-251): debug output, delegated to the reference's utils.visualize if present
# ----------------------------------------------------------------------------------------------
_warned_draw = False


def _visualize():
    global _warned_draw
    try:
        from utils import visualize   # the reference's module, present when used as a drop-in
        return visualize
    except Exception:
        if not _warned_draw:
            warnings.warn("decode_cfg.draw_flag is set but the reference's utils.visualize is not importable; drawing skipped")
            _warned_draw = True
        return None


def draw_kp(img, kps, transforms, kp_threshold, infos, keyword):
    import cv2
    vis = _visualize()
    if vis is None or img is None:
        return img
    for kp in kps:
        img = vis.visualize_kp(img, transforms.detransform_pixel(kp.astype(np.float32), infos))
    cv2.imwrite(r'{}/{}_{}{}.png'.format(base_dir, os.path.basename(infos.img_path), keyword, kp_threshold), img)
    return img


def draw_kp_mask(kp_mask, transforms, kp_threshold, infos, keyword):
    import cv2
    cv2.imwrite(r'{}/mask_{}{}'.format(base_dir, keyword, os.path.basename(infos.img_path)), to_numpy(kp_mask) * 255)
    draw_kp(cv2.imread(infos.img_path), to_numpy(kp_mask.nonzero()), transforms, kp_threshold, infos, keyword)


def draw_box(box_sizes, centers, trans_info, transforms):
    import cv2
    vis = _visualize()
    img = cv2.imread(trans_info.img_path)
    if vis is None or img is None:
        return
    centers = [transforms.detransform_pixel(center, trans_info)[0] for center in centers]
    box_sizes = [box_size[::-1] * compute_scale(trans_info) for box_size in box_sizes]
    img = vis.visualize_box(img, centers, box_sizes, mask=True)
    cv2.imwrite(r'{}/{}_{}.png'.format(base_dir, os.path.basename(trans_info.img_path), "box"), img)


def draw_objs(img, kp_index, kp_mask, transforms, infos):
    """:238-245 — draw, per column of the membership matrix kp_mask [L,C], the key points kp_index [L,2] it selects."""
    l, c = kp_mask.shape
    for i in range(c):
        c_vec = kp_mask[:, i]
        c_kps = to_numpy(kp_index[to_numpy(c_vec.nonzero()), :])
        img = draw_kp(img, c_kps, transforms, i, infos, "objs")
    return img


def draw_candid(kps, lt, rb, img, color):
    import cv2
    if img is None:
        return img
    cv2.rectangle(img, lt, rb, color)
    # :251 reshapes to (-1, 1, 2), which OpenCV >= 4.5 no longer parses as points2f; (-1, 2) is the same point list
    return cv2.drawKeypoints(img, cv2.KeyPoint_convert(np.ascontiguousarray(kps, dtype=np.float32).reshape((-1, 2))), None,
                             color=color)


# ----------------------------------------------------------------------------------------------
# host side of group_kp: per-instance filter + polygons from the grouped device output
# ----------------------------------------------------------------------------------------------
def _identity_transform(transforms) -> bool:
    """True when detransform_pixel is the plain (y,x)->(x,y) flip (utils/tranform.py:157-171 with the default
    val_trans.trans_seq = [])."""
    try:
        return 'resize' not in transforms.configer.get('val_trans', 'trans_seq')
    except Exception:
        return getattr(transforms, "isg_identity", False)


def _polygons_for_image(points, offsets, n, centres_yx, whs, center_cls, center_confs, transforms, info, decode_cfg,
                        identity):
    """:337-369 — `points` [*,2] fp32 (x,y) grouped by instance (ghost-filtered on the device when `identity`)."""
    n_clss, n_confs, n_centers, kps = [], [], [], []
    draw = bool(getattr(decode_cfg, "draw_flag", False))
    img = None
    if draw:
        import cv2
        img = cv2.imread(info.img_path)
        color = [int(e) for e in np.random.randint(0, 257, 3)]
    for i in range(n):
        pts = points[offsets[i]:offsets[i + 1]]
        center_loc = transforms.detransform_pixel(centres_yx[i], info)[0]
        if not identity:
            # non-identity val transform: the device grouped by label only; filter like :339-353 on the host
            h, w = tuple(whs[i] * compute_scale(info))
            true_pixels = transforms.detransform_pixel(pts[:, ::-1], info)
            x, y = center_loc[0], center_loc[1]
            x_mask = (x - (0.5 + decode_cfg.wh_delta) * w < true_pixels[:, 0]) * (true_pixels[:, 0] < x + (0.5 + decode_cfg.wh_delta) * w)
            y_mask = (y - (0.5 + decode_cfg.wh_delta) * h < true_pixels[:, 1]) * (true_pixels[:, 1] < y + (0.5 + decode_cfg.wh_delta) * h)
            pts = true_pixels[x_mask * y_mask]
        if pts.shape[0] < decode_cfg.obj_pixel_th:                 # :355
            continue
        np_poly = aug_group(pts, center_loc)                        # :359
        if np_poly is not None:
            if draw and img is not None:
                h, w = tuple(whs[i] * compute_scale(info))
                x, y = center_loc[0], center_loc[1]
                img = draw_candid(np_poly, (int(x - w / 2), int(y - h / 2)), (int(x + w / 2), int(y + h / 2)), img,
                                  (color[0] * (i + 1) % 256, color[1] * (i + 1) % 256, color[2] * (i + 1) % 256))
            kps.append(np_poly)
            n_centers.append(center_loc)
            n_clss.append(center_cls[i])
            n_confs.append(center_confs[i])
    if draw and img is not None:
        import cv2
        cv2.imwrite(r'{}/{}_{}.png'.format(base_dir, os.path.basename(info.img_path), "candid"), img)
    return n_clss, n_confs, n_centers, kps


def _polar_angles(pts, centres):
    """theta of cartesian2polar for points [K,2] about per-point centres [K,2] (same fp32 arithmetic)."""
    d = (pts - centres).astype(np.float32)
    dx, dy = d[:, 0], d[:, 1]
    with np.errstate(divide="ignore", invalid="ignore"):
        seta = np.arctan(dy / dx)
        seta = np.where(dx < 0, seta + np.float32(np.pi), seta)
        seta = np.where((dx > 0) & (dy < 0), seta + np.float32(2 * np.pi), seta)
        seta = np.where((dx == 0) & (dy > 0), np.float32(np.pi / 2), seta)
        seta = np.where((dx == 0) & (dy < 0), np.float32(3 * np.pi / 2), seta)
    return seta.astype(np.float32)


def _polygons_for_image_fast(points, offsets, n, centres_yx, center_cls, center_confs, obj_pixel_th):
    """The identity-transform, no-drawing case of :337-369 with the per-instance work batched per image:
      * internal points of all instances: one call of libisg's host helper (restates cv2.pointPolygonTest and the
        search of find_internal_point, :51-68);
      * polar angles of all points: one vectorised numpy evaluation (same fp32 arithmetic as cartesian2polar);
      * per instance only np.argsort (kept per instance: its order among equal angles must be the reference's);
      * centre-inside test of all sorted polygons (:201): one helper call.
    The `area == 0` rejection (:187-189) is not evaluated: fillPoly rasterises the outline of the polygon, every
    vertex lies inside the (max+1)-sized canvas, so the area of a polygon with at least one vertex is >= 1
    (tests/test_host_logic.py checks the predicate against the full-canvas rasterisation)."""
    lib = _lib.lib()
    offsets = np.ascontiguousarray(offsets[:n + 1], dtype=np.int32)
    tot = int(offsets[n])
    cnt = np.diff(offsets)
    valid = cnt >= obj_pixel_th                                      # :355
    if tot == 0 or not valid.any():
        return [], [], [], []
    pts = np.ascontiguousarray(points[:tot], dtype=np.float32)
    centers_xy = np.ascontiguousarray(centres_yx[:n, ::-1], dtype=np.float32)   # detransform_pixel flip, (x,y)
    internal = np.empty((n, 2), dtype=np.float32)
    rc = lib.isg_host_internal_points(pts.ctypes.data, offsets.ctypes.data, n, centers_xy.ctypes.data, int(obj_pixel_th),
                                      internal.ctypes.data)
    if rc != 0:
        raise _lib.IsgError(rc, "isg_host_internal_points")
    theta = _polar_angles(pts, np.repeat(internal, cnt, axis=0))
    offs = offsets.tolist()                                          # python ints: no numpy-scalar overhead in the loops
    is_valid = valid.tolist()
    parts = []
    for i in range(n):
        o0, o1 = offs[i], offs[i + 1]
        if is_valid[i]:
            a = theta[o0:o1].argsort()                               # :183
            a += o0
            parts.append(a)
        elif o1 > o0:
            parts.append(np.arange(o0, o1))
    sorted_pts = pts[np.concatenate(parts)]
    inside = np.empty(n, dtype=np.uint8)
    rc = lib.isg_host_centres_inside(sorted_pts.ctypes.data, offsets.ctypes.data, n, centers_xy.ctypes.data, inside.ctypes.data)
    if rc != 0:
        raise _lib.IsgError(rc, "isg_host_centres_inside")
    n_clss, n_confs, n_centers, kps = [], [], [], []
    ok = (valid & (inside != 0)).tolist()
    for i in range(n):
        if ok[i]:                                                    # :201
            kps.append(sorted_pts[offs[i]:offs[i + 1]].copy())
            n_centers.append(centers_xy[i])
            n_clss.append(center_cls[i])
            n_confs.append(center_confs[i])
    return n_clss, n_confs, n_centers, kps


def _overflow(plan, device_polygons) -> int:
    """Keep pixels the plan had no room for: 0, or the per-image capacity a re-run needs.  Ties at the k-th value
    select every tied pixel, so plateaus of the heat map can push an image beyond k = decode_cfg.kp_th; the kernels
    keep counting past the capacity (img_total / count) without storing."""
    seen = int((plan.img_total if device_polygons else plan.count).max().item())
    return seen if seen > plan.cap else 0


def _run_plan(kp, ae, boxes_dev, n_dev, layout, decode_cfg, transforms, dev, max_seeds, min_cap=0):
    """Enqueue select/assign/group for a batch; returns (plan, identity)."""
    B, H, W = kp.shape[0], kp.shape[-2], kp.shape[-1]
    identity = _identity_transform(transforms)
    plan = engine.get_decode_plan(B, H, W, max_seeds, int(decode_cfg.kp_th), dev, decode_mode, want_score=False,
                                  wh_delta=float(decode_cfg.wh_delta) if identity else None,
                                  scale=float(compute_scale(None)), min_cap=min_cap)
    device_polygons = identity and not decode_cfg.draw_flag and decode_mode == "dense" and device_polygon_stage
    plan.run(kp, ae, boxes_dev, n_dev, layout, tail="polygons" if device_polygons else "lists",
             obj_pixel_th=int(decode_cfg.obj_pixel_th))
    return plan, identity, device_polygons


def group_kp(hm_kp, hm_ae, transforms, center_whs, center_indexes, center_cls, center_confs, info, decode_cfg, device):
    """:288-374 — group the boundary key points of ONE image around the detected box centres.
    hm_kp [H,W], hm_ae [4,H,W]; center_indexes / center_whs: per-box (y,x) centres and (h,w) sizes."""
    objs_num = len(center_indexes)
    dev = _device_of(hm_kp, device)
    kp = engine.as_f32_planes(hm_kp, dev)[None]
    ae = engine.as_f32_planes(hm_ae, dev)[None]
    if objs_num == 0:
        return [], [], [], []
    centres = np.vstack(center_indexes).astype(np.float32).reshape(-1, 2)
    whs = np.vstack(center_whs).astype(np.float32).reshape(-1, 2)
    boxes = torch.from_numpy(np.ascontiguousarray(np.concatenate([centres, whs], axis=1))[None]).to(dev)
    n_dev = torch.tensor([objs_num], dtype=torch.int32, device=dev)
    plan, identity, device_polygons = _run_plan(kp, ae, boxes, n_dev, _lib.ISG_BOX_CYCXHW, decode_cfg, transforms, dev, objs_num)
    need = _overflow(plan, device_polygons)
    if need:                      # a plateau at the k-th value selected more pixels than k: decode again with room for them
        plan, identity, device_polygons = _run_plan(kp, ae, boxes, n_dev, _lib.ISG_BOX_CYCXHW, decode_cfg, transforms, dev,
                                                    objs_num, min_cap=need)
    if device_polygons:
        tot = int(plan.img_total[0].item())
        if tot == 0:                                                 # :300
            return [], [], [], []
        st = plan.inst_start[0, :objs_num].cpu().tolist(); ct = plan.inst_count[0, :objs_num].cpu().tolist()
        fl = plan.inst_flags[0, :objs_num].cpu().tolist()
        pts = plan.poly_points[0, :tot].cpu().numpy()
        centers_xy = np.ascontiguousarray(centres[:, ::-1])          # detransform_pixel flip, (x,y) fp32
        n_clss, n_confs, n_centers, kps = [], [], [], []
        for i in range(objs_num):
            poly = None
            if fl[i] == 1:
                poly = pts[st[i]:st[i] + ct[i]].copy()
            elif fl[i] == 2:                                         # more points than the device stage handles
                poly = aug_group(pts[st[i]:st[i] + ct[i]].copy(), centers_xy[i])
            if poly is not None:
                kps.append(poly); n_centers.append(centers_xy[i]); n_clss.append(center_cls[i]); n_confs.append(center_confs[i])
        return n_clss, n_confs, n_centers, kps
    if decode_cfg.draw_flag:
        mask = select_points(kp[0], decode_cfg.kp_th)
        draw_kp_mask(mask, transforms, decode_cfg.kp_th, info, "bound")
    count = int(plan.count[0].item())
    if count == 0:                                                   # :300
        return [], [], [], []
    offsets = plan.offsets[0].cpu().numpy()
    points = plan.points[0, :int(offsets[objs_num])].cpu().numpy()
    if identity and not decode_cfg.draw_flag:
        return _polygons_for_image_fast(points, offsets, objs_num, centres, center_cls, center_confs, decode_cfg.obj_pixel_th)
    return _polygons_for_image(points, offsets, objs_num, centres, whs, center_cls, center_confs, transforms, info,
                               decode_cfg, identity)


# ----------------------------------------------------------------------------------------------
# a2 — box head
# ----------------------------------------------------------------------------------------------
_EMPTY = {'rois': np.array(()), 'class_ids': np.array(()), 'scores': np.array(())}


def _decode_boxes_device(height, width, anchors, regression, classification, threshold, iou_threshold, dev,
                         cap=4096, max_keep=1024):
    """front-end + class-aware NMS on the device; returns the BoxPlan holding the detection tables."""
    reg = engine.as_f32_planes(regression, dev).contiguous()
    cls = engine.as_f32_planes(classification, dev).contiguous()
    anc = engine.as_f32_planes(anchors, dev).contiguous()
    B, A, C = cls.shape
    plan = engine.get_box_plan(B, A, C, height, width, dev, cap, max_keep)
    plan.run(anc, reg, cls, threshold, iou_threshold)
    return plan


def decode_boxes(x, anchors, regression, classification, threshold, iou_threshold):
    """:377-419 — list (per image) of {'rois' f32 [n,4] (x1,y1,x2,y2), 'class_ids' i64 [n], 'scores' f32 [n]}
    sorted by score descending; empty images give three empty arrays."""
    dev = _device_of(classification)
    height, width = x.shape[2], x.shape[3]
    cap, max_keep = 1024, 1024
    while True:
        plan = _decode_boxes_device(height, width, anchors, regression, classification, threshold, iou_threshold, dev,
                                    cap, max_keep)
        n_cand = plan.cand_count.cpu().numpy()
        n_keep = plan.n_keep.cpu().numpy()
        if n_cand.max(initial=0) > plan.cap:          # plan.cap < A here: every anchor can be a candidate at most
            cap = max(cap * 4, int(n_cand.max())); continue
        if n_keep.max(initial=0) > plan.N:
            max_keep = min(max_keep * 4, plan.cap); continue
        break
    rois, scores, cls = plan.rois.cpu().numpy(), plan.scores.cpu().numpy(), plan.cls.cpu().numpy()
    dets = []
    for b in range(plan.B):
        n = int(n_keep[b])
        if n == 0:
            dets.append(dict(_EMPTY))                               # :389-393, :413-417
        else:
            dets.append({'rois': rois[b, :n].copy(), 'class_ids': cls[b, :n].astype(np.int64), 'scores': scores[b, :n].copy()})
    return dets


def decode_single(kp_heat, ae_mat, boxes, info, transforms, decode_cfg, device):
    """:422-441"""
    hm_kp_mat = kp_heat[0]
    center_cls = boxes["class_ids"]
    if center_cls.shape[0] == 0:
        return ([],)
    lt = boxes["rois"][:, :2][:, ::-1]
    rb = boxes["rois"][:, 2:][:, ::-1]
    center_indexes = (lt + rb) / 2
    center_confs = boxes["scores"]
    center_whs = rb - lt
    if decode_cfg.draw_flag:
        draw_box(center_whs, center_indexes, info, transforms)
    center_cls, center_confs, center_indexes, groups = group_kp(hm_kp_mat, ae_mat, transforms, center_whs, center_indexes,
                                                                center_cls, center_confs, info, decode_cfg, device)
    return ([e for e in zip(center_cls, center_confs, center_indexes, groups)],)


def decode_ct_hm(conf_mat, cls_mat, wh, num_classes, cls_th, transforms, info):
    """:254-285 (dead code in the reference; the only caller of py_cpu_nms).  `cls_th` is the COUNT handed to
    select_points (:256).  The peak selection (top-k + 3x3 maximum) and the per-class NMS run on the device; the values
    at the <= k peaks are gathered on the host from the maps as the caller holds them."""
    cat, height, width = wh.size()
    dev = _device_of(conf_mat)
    center_mask = to_numpy(select_points(conf_mat, cls_th)).astype(bool)            # :256
    ys, xs = np.nonzero(center_mask)                                                # row-major, like nonzero() at :258
    center_indexes = np.stack((ys, xs), axis=1)
    center_cls = to_numpy(cls_mat)[ys, xs]                                          # :257
    center_confs = to_numpy(conf_mat)[ys, xs].astype(np.float32)                    # :259
    center_whs = to_numpy(wh)[:, ys, xs].reshape(cat, -1)                           # :260
    keep_cls, keep_idx, keep_confs, keep_whs = [], [], [], []
    for c_i in range(0, num_classes):
        sel = center_cls == c_i
        if sel.sum() == 0:
            continue
        cls, confs, whs, centers = center_cls[sel], center_confs[sel], center_whs[:, sel], center_indexes[sel, :]
        tc = transforms.detransform_pixel(centers, info)[:, ::-1]
        swh = whs * compute_scale(info)
        boxes = np.array([[*(tc[j] - swh[:, j] / 2), *(tc[j] + swh[:, j] / 2), confs[j]] for j in range(tc.shape[0])],
                         dtype=np.float32)
        keep = py_cpu_nms(torch.from_numpy(boxes).to(dev), thresh=0.5)
        keep_cls.extend(cls[keep]); keep_idx.extend(centers[keep]); keep_confs.extend(confs[keep]); keep_whs.extend(whs[:, keep].T)
    return keep_cls, keep_idx, keep_confs, keep_whs


def _host(t):
    """device tensor -> numpy, counting the bytes read back (bench.py reports them)"""
    last_timing["d2h_bytes"] = last_timing.get("d2h_bytes", 0) + t.numel() * t.element_size()
    return t.cpu().numpy()


def _dets_from_polygon_tables(B, n_keep, rois, scores, cls, totals, starts, cnts, flags, pts):
    """decode_output's result lists from the tables of isg_instance_polygons (numpy views, e.g. of an arena's pinned
    mirror).  Per image: one copy of the image's points, polygons are slices of it; no per-instance device traffic."""
    dets = []
    for b in range(B):
        n = int(n_keep[b])
        tot = int(totals[b])
        if n == 0 or tot == 0:                                      # :426-427, :300
            dets.append([]); continue
        fl = flags[b, :n]
        sel = np.flatnonzero(fl == 1)
        r = rois[b, :n]
        centres_xy = (r[:, :2] + r[:, 2:]) / 2                       # :428-432 + detransform_pixel flip -> (x,y)
        pb = pts[b, :tot].copy()
        st = starts[b, :n]
        en = st + cnts[b, :n]
        cls_b = cls[b, :n].astype(np.int64)
        sc_b = scores[b, :n].copy()
        if sel.size == n or not (fl == 2).any():
            stl, enl = st[sel].tolist(), en[sel].tolist()
            polys = [pb[a:e] for a, e in zip(stl, enl)]
            dets.append(list(zip(cls_b[sel], sc_b[sel], centres_xy[sel], polys)))
            continue
        out = []                                                    # rare: instances beyond the device stage's capacity
        for i in range(n):
            if fl[i] == 1:
                out.append((cls_b[i], sc_b[i], centres_xy[i], pb[st[i]:en[i]]))
            elif fl[i] == 2:
                poly = aug_group(pb[st[i]:en[i]].copy(), centres_xy[i])
                if poly is not None:
                    out.append((cls_b[i], sc_b[i], centres_xy[i], poly))
        dets.append(out)
    return dets


def _dets_from_device_polygons(plan, B, n_keep, rois, scores, cls, decode_cfg):
    """Assemble decode_output's result from isg_instance_polygons' buffers (one read-back per buffer)."""
    totals = _host(plan.img_total)
    tot = int(totals.max(initial=0))
    starts = _host(plan.inst_start); cnts = _host(plan.inst_count); flags = _host(plan.inst_flags)
    pts = _host(plan.poly_points[:, :max(tot, 1)])
    return _dets_from_polygon_tables(B, n_keep, rois, scores, cls, totals, starts, cnts, flags, pts)


def _dets_from_arena(pipe, B):
    """the same from the pinned mirror of the pipeline's arena (after arena.wait())"""
    hb, hd = pipe.bplan.host, pipe.dplan.host
    return _dets_from_polygon_tables(B, hb["n_keep"].numpy(), hb["rois"].numpy(), hb["scores"].numpy(), hb["cls"].numpy(),
                                     hd["img_total"].numpy(), hd["inst_start"].numpy(), hd["inst_count"].numpy(),
                                     hd["inst_flags"].numpy(), hd["poly_points"].numpy())


# host-resident model outputs are uploaded and decoded in chunks of this many images, the upload of the later chunks
# overlapping the kernels / read-back / list assembly of the earlier ones (0 disables the chunking)
host_chunk_images = int(os.environ.get("ISG_HOST_CHUNK", "2"))
# pinned host outputs: upload only kp and classification; the kernels gather `ae` at the keep pixels and `regression` at
# the candidate anchors straight from the pinned buffers (the reference's own amount of work, :312-315,:395-398)
host_zero_copy = os.environ.get("ISG_HOST_ZERO_COPY", "1") != "0"
e2e_trace = os.environ.get("ISG_E2E_TRACE", "0") == "1"       # host timeline of the zero-copy path in last_timing["trace_ms"]
_copy_streams = {}
_rings = {}


def _get_ring(n_slots, B, A, C, H, W, height, width, kp_th, dev, cand_cap, max_keep, min_cap, wh_delta, scale):
    key = (n_slots, B, A, C, H, W, height, width, kp_th, dev.index, cand_cap, max_keep, min_cap, wh_delta, scale)
    if key not in _rings:
        if len(_rings) > 6:
            _rings.clear()
        _rings[key] = engine.DecodeRing(lambda: engine.make_pipeline(B, A, C, H, W, height, width, kp_th, dev, cand_cap, max_keep,
                                                                     min_cap, wh_delta, scale), n_slots)
    return _rings[key]


def _plan_overflow(pipe, sparse=False):
    """(cand_cap, max_keep, min_cap) a re-run needs, or None when everything fitted (after arena.wait())"""
    bp, dp = pipe.bplan, pipe.dplan
    n_cand = int(bp.host["cand_count"].max()); n_keep = int(bp.host["n_keep"].max()); tot = int(dp.host["img_total"].max())
    if sparse:                     # the compaction stores at most `cap` keep pixels per image and keeps counting
        tot = max(tot, int(dp.host["count"].max()))
    if n_cand <= bp.cap and n_keep <= bp.N and tot <= dp.cap:
        return None
    cand_cap = min(max(bp.cap * 4, n_cand), bp.A) if n_cand > bp.cap else bp.cap
    return cand_cap, (min(max(bp.N * 4, n_keep), cand_cap) if n_keep > bp.N else bp.N), (tot if tot > dp.cap else 0)


def _fast_path(transforms, decode_cfg) -> bool:
    """identity val-transform without drawing, dense mode: the whole per-instance stage runs on the device"""
    return _identity_transform(transforms) and not decode_cfg.draw_flag and decode_mode == "dense" and device_polygon_stage


def decode_output(inputs, outs, infos, transforms, decode_cfg, device):
    """:444-461 — decode the model output of a batch into per-image lists of
    (class id, confidence, centre (x,y) fp32[2], polygon fp32[K,2] (x,y)).
    Outputs that live in (pinned) host memory are streamed to the device chunk by chunk."""
    kp_out, regression, classification, anchors = outs
    dev = require_cuda(device if device is not None else globals()["device"])
    B = kp_out[0].shape[0]
    cb = host_chunk_images
    on_host = all(t.device.type == "cpu" for t in (kp_out[0], kp_out[1], regression, classification))
    if not (on_host and cb > 0 and B > cb):
        return _decode_output_batch(inputs, outs, infos, transforms, decode_cfg, dev)
    if dev.index not in _copy_streams:
        _copy_streams[dev.index] = torch.cuda.Stream(device=dev)
    cs = _copy_streams[dev.index]
    main = torch.cuda.current_stream(dev)
    anc = engine.as_f32_planes(anchors, dev).contiguous()
    zero_copy = (host_zero_copy and _fast_path(transforms, decode_cfg) and
                 all(t.is_pinned() and t.dtype == torch.float32 and t.is_contiguous() for t in (kp_out[0], kp_out[1], regression, classification)))
    if zero_copy:
        return _decode_output_zero_copy(inputs, kp_out[0], kp_out[1], regression, classification, anc, infos, transforms,
                                        decode_cfg, dev, cb, cs, main)

    def upload(b0):
        b1 = min(b0 + cb, B)
        with torch.cuda.stream(cs):
            t = [x[b0:b1].to(dev, non_blocking=True) for x in (kp_out[0], kp_out[1], regression, classification)]
            ev = torch.cuda.Event()
            ev.record(cs)
        return b0, b1, t, ev

    dets, timing = [], {}
    nxt = upload(0)
    while nxt is not None:
        b0, b1, t, ev = nxt
        nxt = upload(b1) if b1 < B else None                       # enqueue the next upload before decoding this chunk
        main.wait_event(ev)
        sub_inputs = inputs[b0:b1] if inputs.shape[0] == B else inputs
        dets += _decode_output_batch(sub_inputs, ((t[0], t[1], None), t[2], t[3], anc), infos[b0:b1], transforms, decode_cfg, dev)
        for k, v in last_timing.items():
            timing[k] = timing.get(k, 0.0) + v
    last_timing.update(timing)
    return dets


_learned_sizes = {}    # (A, C, H, W, kp_th) -> (cand_cap, max_keep, min_cap) that the last batch of this shape needed


def _decode_output_zero_copy(inputs, kp_h, ae_h, reg_h, cls_h, anc, infos, transforms, decode_cfg, dev, cb, cs, main):
    """decode_output for PINNED host outputs.  Only kp and classification are uploaded (every element of them is
    needed: top-k / 3x3 maxima and the class maximum); `ae` is gathered at the ~k keep pixels and `regression` at the
    candidate anchors by the kernels themselves, out of the pinned buffers.  All uploads are enqueued up front on the copy
    stream; the chunks go through two-slot rings (isg_decode_step, sparse assignment + device polygon stage), so the
    kernels of a chunk overlap the read-back and list assembly of the previous one.  The last chunk is decoded image by
    image: what follows the end of the upload (kernel chain, read-back, assembly of the last piece) is the exposed tail."""
    import time as _time
    B, H, W = kp_h.shape[0], kp_h.shape[-2], kp_h.shape[-1]
    height, width = inputs.shape[2], inputs.shape[3]
    A, C = cls_h.shape[1], cls_h.shape[2]
    size_key = (A, C, H, W, int(decode_cfg.kp_th))
    sizes = _learned_sizes.get(size_key, (1024, 256, 0))

    spans = [(b0, min(cb, B - b0)) for b0 in range(0, B, cb)]
    if spans[-1][1] > 1:                                            # shorter tail: the last chunk image by image
        b0, nb = spans.pop()
        spans += [(b, 1) for b in range(b0, b0 + nb)]
    # one ring slot per chunk (up to 4 per chunk size): every chunk is submitted right away - its kernels wait for its
    # upload on the device - and the host only ever blocks on results
    n_slots = {nb: min(4, sum(1 for _, m in spans if m == nb)) for _, nb in spans}

    def get_ring(nb):
        return _get_ring(n_slots[nb], nb, A, C, H, W, height, width, int(decode_cfg.kp_th), dev, sizes[0], sizes[1], sizes[2],
                         float(decode_cfg.wh_delta), float(compute_scale(None)))
    trace = [] if e2e_trace else None
    _t_start = _time.perf_counter()
    chunks = []
    with torch.cuda.stream(cs):
        for b0, nb in spans:
            kp_d = kp_h[b0:b0 + nb].to(dev, non_blocking=True)
            cls_d = cls_h[b0:b0 + nb].to(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(cs)
            chunks.append((b0, nb, kp_d, cls_d, ev))
    dets, pending, host_s, d2h = [], [], 0.0, 0

    def submit(b0, nb, kp_d, cls_d):
        r = get_ring(nb)
        return r, r.submit(kp_d, ae_h[b0:b0 + nb], anc, reg_h[b0:b0 + nb], cls_d, decode_cfg.cls_th, decode_cfg.iou_th,
                           obj_pixel_th=int(decode_cfg.obj_pixel_th), assign="sparse", fetch=True)

    def finish(r, slot, b0, nb, kp_d, cls_d):
        nonlocal host_s, d2h, sizes
        pipe = r.pipes[slot]
        d2h += pipe.bplan.arena.wait()
        if trace is not None:
            trace.append(("images %d-%d on host" % (b0, b0 + nb - 1), _time.perf_counter() - _t_start))
        over = _plan_overflow(pipe, sparse=True)
        while over is not None:     # more candidates / kept boxes / keep pixels than planned: larger plans from here on
            sizes = (max(sizes[0], over[0]), max(sizes[1], over[1]), max(sizes[2], over[2]))
            _learned_sizes[size_key] = sizes
            r, slot = submit(b0, nb, kp_d, cls_d)
            pipe = r.pipes[slot]
            d2h += pipe.bplan.arena.wait()
            over = _plan_overflow(pipe, sparse=True)
        t0 = _time.perf_counter()
        out = _dets_from_arena(pipe, nb)
        host_s += _time.perf_counter() - t0
        return out

    if trace is not None:
        trace.append(("uploads enqueued", _time.perf_counter() - _t_start))
    for b0, nb, kp_d, cls_d, ev in chunks:
        r = get_ring(nb)
        while sum(1 for p in pending if p[0] is r) >= len(r.pipes):   # a ring slot (round-robin) is read before it is reused
            dets += finish(*pending.pop(0))
        main.wait_event(ev)
        r, slot = submit(b0, nb, kp_d, cls_d)
        pending.append((r, slot, b0, nb, kp_d, cls_d))
        if trace is not None:
            trace.append(("images %d-%d submitted" % (b0, b0 + nb - 1), _time.perf_counter() - _t_start))
    while pending:
        dets += finish(*pending.pop(0))
    if trace is not None:
        trace.append(("all lists assembled", _time.perf_counter() - _t_start))
        last_timing["trace_ms"] = [(n, round(1e3 * t, 3)) for n, t in trace]
    last_timing.update(d2h_bytes=d2h, readback_s=0.0, host_polygons_s=host_s,
                       h2d_bytes=sum(c[2].numel() * 4 + c[3].numel() * 4 for c in chunks))
    return dets


def _decode_output_batch(inputs, outs, infos, transforms, decode_cfg, device):
    """decode_output for one device-sized batch"""
    kp_out, regression, classification, anchors = outs
    dev = require_cuda(device if device is not None else globals()["device"])
    kp = engine.as_f32_planes(kp_out[0], dev)
    ae = engine.as_f32_planes(kp_out[1], dev)
    B, H, W = kp.shape[0], kp.shape[-2], kp.shape[-1]
    height, width = inputs.shape[2], inputs.shape[3]
    cap, max_keep, min_cap = 1024, 256, 0
    reg = engine.as_f32_planes(regression, dev).contiguous()
    cls_t = engine.as_f32_planes(classification, dev).contiguous()
    anc = engine.as_f32_planes(anchors, dev).contiguous()
    A, C = cls_t.shape[1], cls_t.shape[2]
    identity = _identity_transform(transforms)
    import time as _time
    if _fast_path(transforms, decode_cfg):
        # one host call per step (isg_decode_step) and one read-back (the pipeline's arena)
        while True:
            ring = _get_ring(1, B, A, C, H, W, height, width, int(decode_cfg.kp_th), dev, cap, max_keep, min_cap,
                             float(decode_cfg.wh_delta), float(compute_scale(None)))
            slot = ring.submit(kp, ae, anc, reg, cls_t, decode_cfg.cls_th, decode_cfg.iou_th,
                               obj_pixel_th=int(decode_cfg.obj_pixel_th), assign="dense", fetch=True)
            pipe = ring.pipes[slot]
            last_timing["d2h_bytes"] = pipe.bplan.arena.wait()
            over = _plan_overflow(pipe)
            if over is None:
                break
            cap, max_keep, min_cap = over
        _t0 = _time.perf_counter()
        dets = _dets_from_arena(pipe, B)
        last_timing.update(readback_s=0.0, host_polygons_s=_time.perf_counter() - _t0)
        return dets
    while True:
        bplan = engine.get_box_plan(B, A, C, height, width, dev, cap, max_keep)
        plan = engine.get_decode_plan(B, H, W, bplan.N, int(decode_cfg.kp_th), dev, decode_mode, want_score=False,
                                      wh_delta=float(decode_cfg.wh_delta) if identity else None,
                                      scale=float(compute_scale(None)), min_cap=min_cap)
        # the device emits the per-instance point sets, the host runs the polygon stage (non-identity val transform,
        # drawing, sparse mode or ISG_DEVICE_POLYGONS=0)
        engine.get_pipeline(bplan, plan).run(kp, ae, anc, reg, cls_t, decode_cfg.cls_th, decode_cfg.iou_th, tail="lists",
                                             obj_pixel_th=int(decode_cfg.obj_pixel_th))
        # one read-back for the whole batch
        last_timing["d2h_bytes"] = 0
        n_cand = _host(bplan.cand_count)
        n_keep = _host(bplan.n_keep)
        if n_cand.max(initial=0) > bplan.cap:         # bplan.cap < A here
            cap = max(cap * 4, int(n_cand.max())); continue
        if n_keep.max(initial=0) > bplan.N:
            max_keep = min(max(max_keep * 4, int(n_keep.max())), bplan.cap); continue
        need = int(_host(plan.count).max(initial=0))
        if need > plan.cap:       # a plateau at the k-th value selected more pixels than k: decode again with room for them
            min_cap = need; continue
        break
    _t0 = _time.perf_counter()
    rois = _host(bplan.rois); scores = _host(bplan.scores); cls = _host(bplan.cls)
    counts = _host(plan.count)
    offsets = _host(plan.offsets)
    tot = int(offsets[np.arange(B), np.minimum(n_keep, bplan.N)].max(initial=0))
    points = _host(plan.points[:, :max(tot, 1)])
    _t1 = _time.perf_counter()
    dets = []
    for b in range(B):
        n = int(n_keep[b])
        if n == 0 or counts[b] == 0:                                # :426-427, :300
            dets.append([]); continue
        r = rois[b, :n]
        lt, rb = r[:, :2][:, ::-1], r[:, 2:][:, ::-1]
        centres, whs = (lt + rb) / 2, rb - lt                       # :428-432
        if decode_cfg.draw_flag:
            draw_box(whs, centres, infos[b], transforms)
        if identity and not decode_cfg.draw_flag:
            c, f, ctr, g = _polygons_for_image_fast(points[b], offsets[b], n, centres, cls[b, :n].astype(np.int64),
                                                    scores[b, :n], decode_cfg.obj_pixel_th)
        else:
            c, f, ctr, g = _polygons_for_image(points[b], offsets[b], n, centres, whs, cls[b, :n].astype(np.int64),
                                               scores[b, :n], transforms, infos[b], decode_cfg, identity)
        dets.append([e for e in zip(c, f, ctr, g)])
    last_timing.update(readback_s=_t1 - _t0, host_polygons_s=_time.perf_counter() - _t1)
    return dets
