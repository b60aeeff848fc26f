"""ORACLE — test infrastructure only.  CPU restatement of the reference decode path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package; the product (instance-segmentation_b200/) never does.

Parity status: the reference ships no tests, golden vectors or fixtures for this path (SURVEY.md §4,
§8c).  This restatement is pinned instead against outputs of the reference itself, executed in the
build container by oracle/make_golden.py (fixtures in tests/golden/, checked by
tests/test_oracle_golden.py).

Every function cites the reference lines it follows (paths relative to the reference tree).  The
floating-point elementwise ops (tanh, exp, max_pool2d, topk) are torch CPU fp32, the same library
calls the reference makes, so scores agree with the reference bit for bit; control flow is restated
in closed form instead of the reference's Python loops.

Third-party arithmetic the reference delegates to and that is present in this image is called the
same way: torchvision.ops.batched_nms (utils/decode.py:400), cv2.pointPolygonTest / cv2.fillPoly
(utils/decode.py:58,201; utils/image.py:185).  `nms_torchvision_numpy` restates torchvision's
published algorithm and is checked against the library in the tests.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

GRID_H, GRID_W = 1024, 2048


def generate_coordinates():
    """utils/utils.py:453-458 — [2,1024,2048]: ch0 = y in [0,1], ch1 = x in [0,2]."""
    xm = torch.linspace(0, 2, GRID_W).view(1, 1, -1).expand(1, GRID_H, GRID_W)
    ym = torch.linspace(0, 1, GRID_H).view(1, -1, 1).expand(1, GRID_H, GRID_W)
    return torch.cat((ym, xm), 0)


# ----------------------------------------------------------------------------------------------
# a3 — select_points / nms_hm
# ----------------------------------------------------------------------------------------------
def nms_hm(heat: torch.Tensor, kernel: int = 3) -> torch.Tensor:
    """utils/decode.py:42-48."""
    pad = (kernel - 1) // 2
    hmax = F.max_pool2d(heat, (kernel, kernel), stride=1, padding=pad)
    return (hmax == heat).to(torch.uint8)


def select_points(mat: torch.Tensor, k: int) -> torch.Tensor:
    """utils/decode.py:71-85 in closed form (tie-free inputs): the k largest pixels that equal the 3x3
    maximum of (mat where selected else 0)."""
    h, w = mat.shape
    if k > h * w:
        raise RuntimeError("selected index k out of range")          # torch.topk at :81
    if k == 0:
        return torch.zeros((h, w), dtype=torch.uint8)
    thr = torch.topk(mat.reshape(-1), k).values[-1]
    sel = mat >= thr
    v = torch.where(sel, mat, torch.zeros_like(mat))                # mat * mask, :84
    hmax = F.max_pool2d(v[None, None], (3, 3), stride=1, padding=1)[0, 0]
    return ((hmax == v) & sel).to(torch.uint8)                      # :85


# ----------------------------------------------------------------------------------------------
# a13 + a4 + a5 — seeds, embedding, membership, assignment
# ----------------------------------------------------------------------------------------------
def box_geometry(rois: np.ndarray):
    """decode_single, utils/decode.py:428-432: (y,x) centres and sizes from (x1,y1,x2,y2) rois."""
    rois = np.asarray(rois, dtype=np.float32).reshape(-1, 4)
    lt = rois[:, :2][:, ::-1]
    rb = rois[:, 2:][:, ::-1]
    centres = (lt + rb) / 2
    whs = rb - lt
    return centres.astype(np.float32), whs.astype(np.float32)


def membership(e: torch.Tensor, sig: torch.Tensor, pix: torch.Tensor, centres: np.ndarray, whs: np.ndarray,
               ys: torch.Tensor, xs: torch.Tensor, chunk: int = 65536):
    """utils/decode.py:316-328 for M pixels: e [M,2] embeddings, sig [M,2] sigmas, pix [M,2] (y,x) int64.
    Returns (score f32 [M], label i64 [M])."""
    ci = torch.from_numpy(np.ascontiguousarray(centres)).to(torch.int64)          # truncation, :317
    C = torch.stack((ys[ci[:, 0]], xs[ci[:, 1]]), dim=1)[None]                    # [1,N,2]
    c_t = torch.from_numpy(np.ascontiguousarray(centres))
    wh_t = torch.from_numpy(np.ascontiguousarray(whs))
    lt = (c_t - wh_t / 2)[None]                                                   # :321
    rb = (c_t + wh_t / 2)[None]                                                   # :322
    scores, labels = [], []
    for s in range(0, e.shape[0], chunk):
        p = pix[s:s + chunk].float()[:, None]                                     # :323
        mask = (p - lt >= 0).all(dim=2) * (rb - p >= 0).all(dim=2)                # :325
        d = torch.exp(-1 * torch.sum(torch.pow(e[s:s + chunk, None] - C, 2) * sig[s:s + chunk, None], 2))  # :326-327
        sc, lb = (d * mask.float()).max(1)                                        # :328
        scores.append(sc); labels.append(lb)
    if not scores:
        return torch.zeros(0), torch.zeros(0, dtype=torch.int64)
    return torch.cat(scores), torch.cat(labels)


def group_core(hm_kp: torch.Tensor, hm_ae: torch.Tensor, rois: np.ndarray, kp_th: int):
    """The arithmetic core of group_kp (utils/decode.py:299-328) on one image.
    Returns dict(idx [M,2] i64 (y,x) row-major, label [M] i64, score [M] f32, centres, whs)."""
    h, w = hm_kp.shape
    ys = torch.linspace(0, 1, GRID_H)[:h]
    xs = torch.linspace(0, 2, GRID_W)[:w]
    kp_mask = select_points(hm_kp, kp_th)
    idx = kp_mask.nonzero()                                                       # :312
    centres, whs = box_geometry(rois)
    yy, xx = idx[:, 0], idx[:, 1]
    e = torch.stack((torch.tanh(hm_ae[0, yy, xx]) + ys[yy], torch.tanh(hm_ae[1, yy, xx]) + xs[xx]), dim=1)  # :305,313
    sig = torch.exp(torch.stack((hm_ae[2, yy, xx], hm_ae[3, yy, xx]), dim=1))     # :315
    score, label = membership(e, sig, idx, centres, whs, ys, xs)
    return dict(idx=idx, label=label, score=score, centres=centres, whs=whs, mask=kp_mask)


def dense_labels(hm_ae: torch.Tensor, rois: np.ndarray, rows_per_chunk: int = 16):
    """The same membership/assignment for EVERY pixel (the dense kernel's contract): (score [H,W], label [H,W])."""
    _, h, w = hm_ae.shape
    ys = torch.linspace(0, 1, GRID_H)[:h]
    xs = torch.linspace(0, 2, GRID_W)[:w]
    centres, whs = box_geometry(rois)
    score = torch.zeros((h, w)); label = torch.zeros((h, w), dtype=torch.int64)
    for y0 in range(0, h, rows_per_chunk):
        y1 = min(h, y0 + rows_per_chunk)
        yy, xx = torch.meshgrid(torch.arange(y0, y1), torch.arange(w), indexing="ij")
        yy, xx = yy.reshape(-1), xx.reshape(-1)
        e = torch.stack((torch.tanh(hm_ae[0, yy, xx]) + ys[yy], torch.tanh(hm_ae[1, yy, xx]) + xs[xx]), dim=1)
        sig = torch.exp(torch.stack((hm_ae[2, yy, xx], hm_ae[3, yy, xx]), dim=1))
        sc, lb = membership(e, sig, torch.stack((yy, xx), dim=1), centres, whs, ys, xs)
        score[y0:y1] = sc.view(y1 - y0, w); label[y0:y1] = lb.view(y1 - y0, w)
    return score, label


# ----------------------------------------------------------------------------------------------
# a6 — per-instance point sets with the ghost filter
# ----------------------------------------------------------------------------------------------
def detransform_pixel(pixels: np.ndarray, img_size, resize_target=None) -> np.ndarray:
    """CommonTransforms.detransform_pixel (utils/tranform.py:157-171): (y,x) -> (x,y), then — when the validation
    transform resized the image by 1/resize_target — the inverse affine map back to the original img_size (h,w),
    clipped to the frame (utils/image.py:48-82)."""
    import cv2
    rev = pixels.reshape(-1, 2)[:, ::-1]
    if resize_target is None:
        return rev
    height, width = img_size
    out_size = (int(round(width * (1 / resize_target))), int(round(height * (1 / resize_target))))   # :166-167
    in_size = tuple(img_size)[::-1]
    src = np.array([[0, 0], [0, in_size[1] - 1], [in_size[0] - 1, in_size[1] - 1]], dtype=np.float32)    # image.py:56
    dst = np.array([[0, 0], [0, out_size[1] - 1], [out_size[0] - 1, out_size[1] - 1]], dtype=np.float32)
    t = cv2.getAffineTransform(dst, src).astype(np.float32)                                              # inv=True, :63,78
    pts_h = np.hstack((rev, np.ones((rev.shape[0], 1), dtype=np.float32)))
    out = np.dot(t, pts_h.T).T
    out[:, 0] = out[:, 0].clip(min=0, max=in_size[0] - 1)
    out[:, 1] = out[:, 1].clip(min=0, max=in_size[1] - 1)
    return out[:, :2]


def instance_points(idx: torch.Tensor, label: torch.Tensor, centres: np.ndarray, whs: np.ndarray, wh_delta: float,
                    scale=1, img_size=None, resize_target=None):
    """utils/decode.py:337-353.  With resize_target None the val transform is the identity (detransform_pixel =
    (y,x)->(x,y) flip, utils/tranform.py:157-159); otherwise pixels and centres go through the inverse resize and
    `scale` (= decode.target_size, :34-35) multiplies the box sizes.
    Returns list over instances of (points f32 [K,2] (x,y), centre f32 [2] (x,y))."""
    out = []
    idx_np = idx.numpy()
    lab_np = label.numpy()
    for i in range(centres.shape[0]):
        h, w = tuple(whs[i] * scale)                                             # :339
        sel = np.nonzero(lab_np == i)[0]                                         # :342
        true_pixels = detransform_pixel(idx_np[sel].astype(np.float32), img_size, resize_target)   # :343-345
        center_loc = detransform_pixel(centres[i], img_size, resize_target)[0]   # :347-348
        x, y = center_loc[0], center_loc[1]
        xm = (x - (0.5 + wh_delta) * w < true_pixels[:, 0]) * (true_pixels[:, 0] < x + (0.5 + wh_delta) * w)  # :351
        ym = (y - (0.5 + wh_delta) * h < true_pixels[:, 1]) * (true_pixels[:, 1] < y + (0.5 + wh_delta) * h)  # :352
        out.append((true_pixels[xm * ym], center_loc))
    return out


# ----------------------------------------------------------------------------------------------
# a7 — polygon stage
# ----------------------------------------------------------------------------------------------
def cartesian2polar(kps: np.ndarray, center_loc: np.ndarray) -> np.ndarray:
    """utils/decode.py:88-113, vectorised with the same fp32 arithmetic; returns [K,2] (theta, d) fp32."""
    d = (kps - center_loc).astype(np.float32)
    dx, dy = d[:, 0], d[:, 1]
    with np.errstate(divide="ignore", invalid="ignore"):
        seta = np.arctan(dy / dx)                                                # :104
        seta = np.where(dx < 0, seta + np.float32(np.pi), seta)                  # :105-106
        seta = np.where((dx > 0) & (dy < 0), seta + np.float32(2 * np.pi), seta)  # :107-108
        seta = np.where((dx == 0) & (dy > 0), np.float32(np.pi / 2), seta)       # :99-100
        seta = np.where((dx == 0) & (dy < 0), np.float32(3 * np.pi / 2), seta)   # :101-102
        dist = np.sqrt(dx ** 2 + dy ** 2)                                        # :110
    return np.stack((seta, dist), axis=1).astype(np.float32)


def find_internal_point(kps: np.ndarray, default: np.ndarray):
    """utils/decode.py:51-68."""
    import cv2
    kps = np.array(kps)
    if cv2.pointPolygonTest(kps, tuple(default), False) > 0:
        return default
    mean = kps.mean(axis=0).reshape(-1)
    if cv2.pointPolygonTest(kps, tuple(mean), False) > 0:
        return mean
    for i in range(kps.shape[0]):
        for j in range(1, kps.shape[0]):
            point = (kps[i] + kps[j]) / 2
            if cv2.pointPolygonTest(kps, tuple(point), False) > 0:
                return point
    return default


def poly_to_mask(poly: np.ndarray, img_size=None) -> np.ndarray:
    """utils/image.py:180-185."""
    import cv2
    poly = poly.astype(np.int32)
    if img_size is None:
        img_size = (poly.max(0) + 1)[::-1]
    return cv2.fillPoly(np.zeros(img_size, dtype=np.int32), [poly], 1)


def aug_group(pts: np.ndarray, center_loc: np.ndarray):
    """utils/decode.py:167-204."""
    import cv2
    center_loc = center_loc.reshape(-1)
    internal = find_internal_point(pts, center_loc)
    polar = cartesian2polar(pts, internal)
    order = np.argsort(polar[:, 0])                                              # :183
    sorted_kp = pts[order]
    if poly_to_mask(sorted_kp).sum() == 0:                                       # :187-189
        return None
    if cv2.pointPolygonTest(sorted_kp, tuple(center_loc), False) > 0:            # :201
        return sorted_kp
    return None


def group_kp(hm_kp, hm_ae, rois, class_ids, scores, kp_th=20000, wh_delta=0.1, obj_pixel_th=2, scale=1, img_size=None,
             resize_target=None):
    """group_kp (utils/decode.py:288-374) with draw_flag False; identity val transform unless resize_target is given."""
    n = len(rois)
    if n == 0:
        return [], [], [], []
    core = group_core(hm_kp, hm_ae, rois, kp_th)
    if core["idx"].shape[0] == 0:                                                # :300
        return [], [], [], []
    clss, confs, centers, polys = [], [], [], []
    for i, (pts, center_loc) in enumerate(instance_points(core["idx"], core["label"], core["centres"], core["whs"], wh_delta,
                                                           scale, img_size, resize_target)):
        if pts.shape[0] < obj_pixel_th:                                          # :355
            continue
        poly = aug_group(pts, center_loc)
        if poly is not None:
            polys.append(poly); centers.append(center_loc); clss.append(class_ids[i]); confs.append(scores[i])
    return clss, confs, centers, polys


def decode_single(kp_heat, ae_mat, boxes, kp_th=20000, wh_delta=0.1, obj_pixel_th=2, scale=1, img_size=None,
                  resize_target=None):
    """utils/decode.py:422-441."""
    if boxes["class_ids"].shape[0] == 0:
        return ([],)
    c, f, ctr, g = group_kp(kp_heat[0], ae_mat, boxes["rois"], boxes["class_ids"], boxes["scores"], kp_th, wh_delta,
                            obj_pixel_th, scale, img_size, resize_target)
    return ([e for e in zip(c, f, ctr, g)],)


def decode_ct_hm(conf_mat: torch.Tensor, cls_mat: torch.Tensor, wh: torch.Tensor, num_classes: int, k: int):
    """utils/decode.py:254-285 with the identity val transform and target_size 1: the `k` (= cls_th, a COUNT) best
    3x3 peaks of conf_mat -> per class boxes [c - wh/2, c + wh/2, conf] in (x,y) order -> py_cpu_nms(0.5).
    Returns (classes i64 [n], centre indexes i64 [n,2] (y,x), confidences f32 [n], sizes f32 [n,2])."""
    from .ref_kmeans_nms import py_cpu_nms
    cat = wh.shape[0]
    mask = select_points(conf_mat, k).bool()                                     # :256
    center_cls = cls_mat[mask].numpy()                                           # :257
    center_indexes = mask.nonzero().numpy()                                      # :258
    center_confs = conf_mat[mask].numpy().astype(np.float32)                     # :259
    center_whs = wh[:, mask].numpy().reshape(cat, -1)                            # :260
    kc, ki, kf, kw = [], [], [], []
    for c_i in range(num_classes):                                               # :266
        sel = center_cls == c_i
        if sel.sum() == 0:
            continue
        cls, confs, whs, centers = center_cls[sel], center_confs[sel], center_whs[:, sel], center_indexes[sel, :]
        tc = detransform_pixel(centers, None)[:, ::-1]                           # :274 (flip twice: back to (y,x))
        boxes = np.array([[*(tc[j] - whs[:, j] / 2), *(tc[j] + whs[:, j] / 2), confs[j]] for j in range(tc.shape[0])],
                         dtype=np.float32)                                       # :276
        keep = py_cpu_nms(boxes, 0.5)                                            # :277
        kc.extend(cls[keep]); ki.extend(centers[keep]); kf.extend(confs[keep]); kw.extend(whs[:, keep].T)
    return (np.asarray(kc, dtype=np.int64), np.asarray(ki, dtype=np.int64).reshape(-1, 2), np.asarray(kf, dtype=np.float32),
            np.asarray(kw, dtype=np.float32).reshape(-1, 2))


# ----------------------------------------------------------------------------------------------
# a2 — box head
# ----------------------------------------------------------------------------------------------
def bbox_transform(anchors: torch.Tensor, regression: torch.Tensor) -> torch.Tensor:
    """utils/utils.py:318-346."""
    yca = (anchors[..., 0] + anchors[..., 2]) / 2
    xca = (anchors[..., 1] + anchors[..., 3]) / 2
    ha = anchors[..., 2] - anchors[..., 0]
    wa = anchors[..., 3] - anchors[..., 1]
    w = regression[..., 3].exp() * wa
    h = regression[..., 2].exp() * ha
    yc = regression[..., 0] * ha + yca
    xc = regression[..., 1] * wa + xca
    return torch.stack([xc - w / 2., yc - h / 2., xc + w / 2., yc + h / 2.], dim=2)


def clip_boxes(boxes: torch.Tensor, height: int, width: int) -> torch.Tensor:
    """utils/utils.py:349-363 (out of place)."""
    b = boxes.clone()
    b[:, :, 0] = torch.clamp(b[:, :, 0], min=0)
    b[:, :, 1] = torch.clamp(b[:, :, 1], min=0)
    b[:, :, 2] = torch.clamp(b[:, :, 2], max=width - 1)
    b[:, :, 3] = torch.clamp(b[:, :, 3], max=height - 1)
    return b


def nms_torchvision_numpy(boxes: np.ndarray, scores: np.ndarray, thr: float) -> np.ndarray:
    """torchvision/csrc/ops/cpu/nms_kernel.cpp (0.26): areas without +1, visit by score descending,
    suppress j iff inter/(area_i+area_j-inter) > thr, fp32 arithmetic."""
    b = boxes.astype(np.float32)
    order = np.argsort(-scores.astype(np.float32), kind="stable")
    areas = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    dead = np.zeros(len(b), dtype=bool)
    keep = []
    for pos, i in enumerate(order):
        if dead[i]:
            continue
        keep.append(i)
        rest = order[pos + 1:]
        xx1 = np.maximum(b[i, 0], b[rest, 0]); yy1 = np.maximum(b[i, 1], b[rest, 1])
        xx2 = np.minimum(b[i, 2], b[rest, 2]); yy2 = np.minimum(b[i, 3], b[rest, 3])
        inter = np.maximum(np.float32(0), xx2 - xx1) * np.maximum(np.float32(0), yy2 - yy1)
        with np.errstate(divide="ignore", invalid="ignore"):
            ovr = inter / (areas[i] + areas[rest] - inter)
        dead[rest[ovr.astype(np.float64) > thr]] = True
    return np.asarray(keep, dtype=np.int64)


def batched_nms_numpy(boxes: np.ndarray, scores: np.ndarray, classes: np.ndarray, thr: float) -> np.ndarray:
    """Class-aware NMS = torchvision.ops.batched_nms semantics evaluated per class on the ORIGINAL
    coordinates (the library's coordinate-offset trick changes low-order bits only; inputs keep IoUs
    1e-4 away from the threshold).  Result sorted by score descending like the library's."""
    keep = []
    for c in np.unique(classes):
        ids = np.nonzero(classes == c)[0]
        keep.extend(ids[nms_torchvision_numpy(boxes[ids], scores[ids], thr)])
    keep = np.asarray(keep, dtype=np.int64)
    return keep[np.argsort(-scores[keep], kind="stable")]


def batched_nms_trick_numpy(boxes: np.ndarray, scores: np.ndarray, classes: np.ndarray, thr: float) -> np.ndarray:
    """torchvision.ops.boxes._batched_nms_coordinate_trick restated (the only form of batched_nms in torchvision 0.5.0,
    the reference's pin, and the CPU path of current torchvision for up to 1000 boxes): every box is shifted by
    float32(class) * (boxes.max() + 1) in fp32 and ONE class-agnostic NMS runs on the shifted coordinates."""
    if len(boxes) == 0:
        return np.zeros(0, dtype=np.int64)
    b = boxes.astype(np.float32)
    off = classes.astype(np.float32) * (b.max() + np.float32(1))
    return nms_torchvision_numpy(b + off[:, None], scores, thr)


def decode_boxes(height, width, anchors, regression, classification, threshold, iou_threshold, use_torchvision=True):
    """utils/decode.py:377-419; x is only used for its H, W."""
    from torchvision.ops.boxes import batched_nms
    boxes = clip_boxes(bbox_transform(anchors, regression), height, width)
    scores = torch.max(classification, dim=2, keepdim=True)[0]
    over = (scores > threshold)[:, :, 0]
    dets = []
    empty = {"rois": np.array(()), "class_ids": np.array(()), "scores": np.array(())}
    for i in range(classification.shape[0]):
        if over[i].sum() == 0:
            dets.append(dict(empty)); continue
        cls_per = classification[i, over[i, :], ...].permute(1, 0)
        box_per = boxes[i, over[i, :], ...]
        sc_per = scores[i, over[i, :], ...]
        sc_, cl_ = cls_per.max(dim=0)
        if use_torchvision:
            keep = batched_nms(box_per, sc_per[:, 0], cl_, iou_threshold=iou_threshold)
        else:
            keep = torch.from_numpy(batched_nms_numpy(box_per.numpy(), sc_per[:, 0].numpy(), cl_.numpy(), iou_threshold))
        if keep.shape[0] != 0:
            dets.append({"rois": box_per[keep, :].numpy(), "class_ids": cl_[keep].numpy(), "scores": sc_[keep].numpy()})
        else:
            dets.append(dict(empty))
    return dets


def decode_output(height, width, outs, kp_th=20000, cls_th=0.3, iou_th=0.2, wh_delta=0.1, obj_pixel_th=2):
    """utils/decode.py:444-461 (serial over the batch like utils/parell_util.py:5-8)."""
    kp_out, regression, classification, anchors = outs
    det_boxes = decode_boxes(height, width, anchors, regression, classification, cls_th, iou_th)
    res = []
    for b in range(kp_out[0].shape[0]):
        res.append(decode_single(kp_out[0][b], kp_out[1][b].clone(), det_boxes[b], kp_th, wh_delta, obj_pixel_th)[0])
    return res
