"""Deterministic synthetic inputs for the decode hot path.

Shapes follow what EfficientSeg emits (models/efficient.py:615-626 of the reference):
``kp [1,H,W]`` boundary-keypoint logits, ``ae [4,H,W]`` (2 embedding offsets + 2 log-sigmas), detected
boxes ``rois [N,4]`` (x1,y1,x2,y2), and for the box head ``anchors [1,A,4]`` (y1,x1,y2,x2),
``regression [A,4]`` (dy,dx,dh,dw), ``classification [A,C]`` (already sigmoid).

Everything is tie-free by construction so that label maps and NMS keep lists are well defined:
  * every kp value of an image is distinct (top-k and the 3x3 peak test never see a tie);
  * box centres sit at integer+0.5 with even sizes, so the truncated centre index and the inclusive
    in-box test are stable against 1-ulp differences in exp();
  * outline pixels whose best/second-best membership margin (fp64) is below 1e-3 are removed;
  * candidate boxes whose IoU is within 1e-4 of the NMS threshold are removed.
Only numpy RandomState is used (frozen bit streams).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

GRID_H, GRID_W = 1024, 2048  # utils/utils.py:453-458


def coordinate_tables(H: int, W: int):
    """ys[H], xs[W] exactly as the reference slices them (utils/decode.py:304)."""
    if H > GRID_H or W > GRID_W:
        raise ValueError("the reference coordinate grid is 1024x2048 (utils/utils.py:453-458)")
    ys = torch.linspace(0, 1, GRID_H)[:H].contiguous()
    xs = torch.linspace(0, 2, GRID_W)[:W].contiguous()
    return ys, xs


@dataclass
class Image:
    kp: np.ndarray          # [1,H,W] float32
    ae: np.ndarray          # [4,H,W] float32
    rois: np.ndarray        # [N,4] float32 (x1,y1,x2,y2), sorted by score desc
    class_ids: np.ndarray   # [N] int64
    scores: np.ndarray      # [N] float32
    owner: np.ndarray       # [H,W] int32: instance whose outline covers the pixel, -1 elsewhere


def _distinct_float32(v: np.ndarray) -> np.ndarray:
    """Return a float32 array with the same ordering as `v` (ties broken by position) and no two EQUAL values
    (-0.0 never appears, so key-distinct implies value-distinct)."""
    flat = v.astype(np.float32).ravel() + np.float32(0.0)                 # -0.0 -> +0.0
    u = flat.view(np.uint32)
    key = np.where(u & 0x80000000, ~u, u | 0x80000000).astype(np.int64)  # monotone in the float value
    k0 = 0x7FFFFFFF                                                      # key of -0.0: removed from the key space
    key = key - (key > k0)
    order = np.argsort(key, kind="stable")
    ks = key[order]
    i = np.arange(ks.size, dtype=np.int64)
    ks2 = np.maximum.accumulate(ks - i) + i                              # strictly increasing
    out_key = np.empty_like(ks2)
    out_key[order] = ks2
    out_key = out_key + (out_key >= k0)
    ku = out_key.astype(np.uint64).astype(np.uint32)
    back = np.where(ku & 0x80000000, ku & 0x7FFFFFFF, ~ku).astype(np.uint32)
    return back.view(np.float32).reshape(v.shape)


def make_boxes(rs: np.random.RandomState, H: int, W: int, N: int, C: int = 8, min_sep: float = 6.0):
    """N boxes with centres at integer+0.5, even sizes, inside the image, centres >= min_sep px apart."""
    hmax = max(10, min(198, (H // 3) // 2 * 2))
    wmax = max(10, min(298, (W // 3) // 2 * 2))
    hmin, wmin = min(30, hmax), min(30, wmax)
    centres, rois = [], []
    tries = 0
    while len(rois) < N:
        tries += 1
        if tries > 200 * N + 1000:
            raise RuntimeError("could not place %d separated boxes in %dx%d" % (N, H, W))
        h = int(rs.randint(hmin // 2, hmax // 2 + 1)) * 2
        w = int(rs.randint(wmin // 2, wmax // 2 + 1)) * 2
        cy = int(rs.randint(h // 2 + 1, H - h // 2 - 1)) + 0.5
        cx = int(rs.randint(w // 2 + 1, W - w // 2 - 1)) + 0.5
        if centres:
            c = np.asarray(centres)
            if np.min(np.hypot(c[:, 0] - cy, c[:, 1] - cx)) < min_sep:
                continue
        centres.append((cy, cx))
        rois.append((cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2))
    rois = np.asarray(rois, dtype=np.float32).reshape(-1, 4)
    scores = np.sort(_distinct_float32(rs.uniform(0.3, 1.0, size=N).astype(np.float32)))[::-1].copy()
    class_ids = rs.randint(0, C, size=N).astype(np.int64)
    return rois, class_ids, scores


def make_image(seed: int, H: int, W: int, N: int, C: int = 8, sigma: float = 200.0) -> Image:
    rs = np.random.RandomState(seed)
    ys, xs = coordinate_tables(H, W)
    ys64, xs64 = ys.double().numpy(), xs.double().numpy()
    rois, class_ids, scores = make_boxes(rs, H, W, N, C)

    kp = rs.normal(-6.0, 0.5, size=(H, W)).astype(np.float32)
    ae = np.empty((4, H, W), dtype=np.float32)
    ae[0:2] = rs.normal(0.0, 0.01, size=(2, H, W)).astype(np.float32)
    ae[2:4] = (np.log(sigma) + rs.normal(0.0, 0.1, size=(2, H, W))).astype(np.float32)
    owner = np.full((H, W), -1, dtype=np.int32)

    # seed coordinates exactly as group_kp builds them (utils/decode.py:316-317)
    cy = (rois[:, 1] + rois[:, 3]) / 2
    cx = (rois[:, 0] + rois[:, 2]) / 2
    hh = rois[:, 3] - rois[:, 1]
    ww = rois[:, 2] - rois[:, 0]
    Cy = ys64[cy.astype(np.int64)]
    Cx = xs64[cx.astype(np.int64)]

    for j in range(N):
        n_t = int(4 * (hh[j] + ww[j]))
        t = np.linspace(0.0, 2 * np.pi, n_t, endpoint=False)
        py = np.rint(cy[j] + (hh[j] / 2 - 1.5) * np.sin(t)).astype(np.int64)
        px = np.rint(cx[j] + (ww[j] / 2 - 1.5) * np.cos(t)).astype(np.int64)
        owner[py, px] = j

    oy, ox = np.nonzero(owner >= 0)
    oj = owner[oy, ox]
    kp[oy, ox] = rs.uniform(4.0, 5.0, size=oy.size).astype(np.float32)
    ae[0, oy, ox] = (np.arctanh(Cy[oj] - ys64[oy]) + rs.normal(0, 1e-4, size=oy.size)).astype(np.float32)
    ae[1, oy, ox] = (np.arctanh(Cx[oj] - xs64[ox]) + rs.normal(0, 1e-4, size=oy.size)).astype(np.float32)

    # fp64 margin certification on every outline pixel (a superset of the pixels the decode keeps)
    if oy.size:
        ey = np.tanh(ae[0, oy, ox].astype(np.float64)) + ys64[oy]
        ex = np.tanh(ae[1, oy, ox].astype(np.float64)) + xs64[ox]
        sy = np.exp(ae[2, oy, ox].astype(np.float64))
        sx = np.exp(ae[3, oy, ox].astype(np.float64))
        lty, ltx = cy - hh / 2, cx - ww / 2
        rby, rbx = cy + hh / 2, cx + ww / 2
        bad = np.zeros(oy.size, dtype=bool)
        step = max(1, 4_000_000 // max(N, 1))
        for s in range(0, oy.size, step):
            e = slice(s, s + step)
            inb = ((oy[e, None] >= lty[None]) & (ox[e, None] >= ltx[None]) &
                   (oy[e, None] <= rby[None]) & (ox[e, None] <= rbx[None]))
            P = np.exp(-((ey[e, None] - Cy[None]) ** 2 * sy[e, None] + (ex[e, None] - Cx[None]) ** 2 * sx[e, None])) * inb
            top = np.argmax(P, axis=1)
            best = P[np.arange(P.shape[0]), top]
            P[np.arange(P.shape[0]), top] = -1.0
            second = P.max(axis=1) if N > 1 else np.zeros_like(best)
            bad[e] = (top != oj[e]) | (best - np.maximum(second, 0.0) < 1e-3)
        if bad.any():
            by, bx = oy[bad], ox[bad]
            owner[by, bx] = -1
            kp[by, bx] = rs.normal(-6.0, 0.5, size=by.size).astype(np.float32)
            ae[0, by, bx] = rs.normal(0.0, 0.01, size=by.size).astype(np.float32)
            ae[1, by, bx] = rs.normal(0.0, 0.01, size=by.size).astype(np.float32)

    kp = _distinct_float32(kp)
    return Image(kp=kp[None], ae=ae, rois=rois, class_ids=class_ids, scores=scores, owner=owner)


# --------------------------------------------------------------------------------------------
# box head
# --------------------------------------------------------------------------------------------
def make_anchors(H: int, W: int, strides=(8, 16, 32, 64, 128), anchor_scale: float = 4.0) -> np.ndarray:
    """EfficientDet-style anchors [1,A,4] (y1,x1,y2,x2); A = 9 * sum((H/s)*(W/s))."""
    scales = (2 ** 0, 2 ** (1 / 3), 2 ** (2 / 3))
    ratios = ((1.0, 1.0), (1.4, 0.7), (0.7, 1.4))
    levels = []
    for s in strides:
        if H % s or W % s:
            raise ValueError("H and W must be divisible by every stride")
        per = []
        for sc in scales:
            for rx, ry in ratios:
                base = anchor_scale * s * sc
                ax2, ay2 = base * rx / 2.0, base * ry / 2.0
                x = np.arange(s / 2, W, s)
                y = np.arange(s / 2, H, s)
                xv, yv = np.meshgrid(x, y)
                xv, yv = xv.reshape(-1), yv.reshape(-1)
                per.append(np.vstack((yv - ay2, xv - ax2, yv + ay2, xv + ax2)).T[:, None, :])
        levels.append(np.concatenate(per, axis=1).reshape(-1, 4))
    return np.ascontiguousarray(np.vstack(levels), dtype=np.float32)[None]


def _iou_matrix(b: np.ndarray) -> np.ndarray:
    b = b.astype(np.float64).reshape(-1, 4)
    area = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    x1 = np.maximum(b[:, None, 0], b[None, :, 0]); y1 = np.maximum(b[:, None, 1], b[None, :, 1])
    x2 = np.minimum(b[:, None, 2], b[None, :, 2]); y2 = np.minimum(b[:, None, 3], b[None, :, 3])
    inter = np.clip(x2 - x1, 0, None) * np.clip(y2 - y1, 0, None)
    with np.errstate(divide="ignore", invalid="ignore"):
        return inter / (area[:, None] + area[None, :] - inter)


def make_box_head(seed: int, H: int, W: int, C: int, K: int, anchors: np.ndarray | None = None,
                  cls_th: float = 0.3, iou_th: float = 0.2):
    """regression [A,4], classification [A,C] with K candidate anchors above cls_th (distinct scores)."""
    rs = np.random.RandomState(seed)
    if anchors is None:
        anchors = make_anchors(H, W)
    A = anchors.shape[1]
    regression = rs.normal(0.0, 0.1, size=(A, 4)).astype(np.float32)
    classification = rs.uniform(0.0, 0.05, size=(A, C)).astype(np.float32)
    chosen = rs.choice(A, size=min(K, A), replace=False)
    sc = _distinct_float32(rs.uniform(0.31, 1.0, size=chosen.size).astype(np.float32))
    cl = rs.randint(0, C, size=chosen.size)
    # drop candidates that sit within 1e-4 of the IoU threshold against another candidate
    a = anchors[0, chosen].astype(np.float64); r = regression[chosen].astype(np.float64)
    yc = r[:, 0] * (a[:, 2] - a[:, 0]) + (a[:, 0] + a[:, 2]) / 2
    xc = r[:, 1] * (a[:, 3] - a[:, 1]) + (a[:, 1] + a[:, 3]) / 2
    h = np.exp(r[:, 2]) * (a[:, 2] - a[:, 0]); w = np.exp(r[:, 3]) * (a[:, 3] - a[:, 1])
    bx = np.stack([np.clip(xc - w / 2, 0, None), np.clip(yc - h / 2, 0, None),
                   np.clip(xc + w / 2, None, W - 1), np.clip(yc + h / 2, None, H - 1)], axis=1)
    iou = _iou_matrix(bx)
    np.fill_diagonal(iou, 0.0)
    near = np.abs(iou - iou_th) < 1e-4
    keep = np.ones(chosen.size, dtype=bool)
    for i in np.nonzero(near.any(axis=1))[0]:
        if keep[i] and (near[i] & keep).any():
            keep[i] = False
    chosen, sc, cl = chosen[keep], sc[keep], cl[keep]
    classification[chosen, cl] = sc
    return anchors, regression, classification


def make_nms_boxes(seed: int, n: int, extent: float = 1000.0, thr: float = 0.5, plus1: bool = True):
    """dets [n,5] float32 (x1,y1,x2,y2,score) with distinct scores and no IoU within 1e-4 of thr."""
    rs = np.random.RandomState(seed)
    ctr = rs.uniform(0, extent, size=(n, 2))
    wh = rs.uniform(20, 200, size=(n, 2))
    b = np.concatenate([ctr - wh / 2, ctr + wh / 2], axis=1).astype(np.float32)
    s = _distinct_float32(rs.uniform(0.05, 1.0, size=n).astype(np.float32))
    bb = b.astype(np.float64).copy()
    if plus1:
        bb[:, 2:] += 1.0
    iou = _iou_matrix(bb)
    np.fill_diagonal(iou, 0.0)
    near = (np.abs(iou - thr) < 1e-4).any(axis=1)
    b, s = b[~near], s[~near]
    return np.concatenate([b, s[:, None]], axis=1).astype(np.float32)


def make_near_threshold_boxes(seed: int, n_pairs: int, thr: float, ncls: int = 8, extent: float = 2000.0):
    """Adversarial NMS input: pairs of same-class boxes whose IoU sits within a few fp32 ulps of `thr` (bisection on a
    horizontal shift in fp64), so that the low-order bits of the coordinates the IoU is evaluated on decide the
    outcome - torchvision's coordinate trick and its per-class NMS resolve a few percent of the pairs differently.
    Returns boxes [2n,4] (x1,y1,x2,y2) float32, scores [2n] float32 (distinct), classes [2n] int64."""
    rs = np.random.RandomState(seed)
    out_b, out_c = [], []
    for _ in range(n_pairs):
        w, h = rs.uniform(30, 200, 2)
        x, y = rs.uniform(0, extent - 2 * w), rs.uniform(0, extent - h)
        lo, hi = 0.0, w
        for _ in range(60):                 # IoU of two equal boxes shifted by d in x: (w - d) / (w + d)
            mid = 0.5 * (lo + hi)
            if (w - mid) / (w + mid) > thr:
                lo = mid
            else:
                hi = mid
        d = lo + rs.uniform(-3e-5, 3e-5) * w
        c = rs.randint(0, ncls)
        out_b += [[x, y, x + w, y + h], [x + d, y, x + d + w, y + h]]
        out_c += [c, c]
    s = _distinct_float32(rs.uniform(0.1, 1.0, size=2 * n_pairs).astype(np.float32))
    return np.asarray(out_b, np.float32), s, np.asarray(out_c, np.int64)


def make_masks(seed: int, n: int, H: int, W: int, C: int = 80):
    """n ellipse masks bit-packed to uint32 [n,H,ceil(W/32)], boxes int32 [n,4] (x0,y0,x1,y1), scores, classes."""
    rs = np.random.RandomState(seed)
    Ww = (W + 31) // 32
    masks = np.zeros((n, H, Ww), dtype=np.uint32)
    boxes = np.zeros((n, 4), dtype=np.int32)
    yy = np.arange(H)[:, None]
    xx = np.arange(W)[None, :]
    n_groups = max(1, n // 4)
    gcy = rs.uniform(0.1 * H, 0.9 * H, size=n_groups); gcx = rs.uniform(0.1 * W, 0.9 * W, size=n_groups)
    gcls = rs.randint(0, C, size=n_groups)
    cls = np.zeros(n, dtype=np.int64)
    for i in range(n):
        g = rs.randint(0, n_groups)
        cls[i] = gcls[g] if rs.rand() < 0.8 else rs.randint(0, C)
        ry, rx = rs.uniform(0.03 * H, 0.12 * H), rs.uniform(0.03 * W, 0.12 * W)
        cy, cx = gcy[g] + rs.normal(0, 0.3 * ry), gcx[g] + rs.normal(0, 0.3 * rx)
        y0, y1 = max(0, int(cy - ry)), min(H - 1, int(cy + ry))
        x0, x1 = max(0, int(cx - rx)), min(W - 1, int(cx + rx))
        sub = (((yy[y0:y1 + 1] - cy) / ry) ** 2 + ((xx[:, x0:x1 + 1] - cx) / rx) ** 2) <= 1.0
        full = np.zeros((y1 - y0 + 1, Ww * 32), dtype=bool)
        full[:, x0:x1 + 1] = sub
        packed = np.packbits(full.reshape(full.shape[0], Ww, 32), axis=2, bitorder="little").view(np.uint32)[..., 0]
        masks[i, y0:y1 + 1] = packed
        boxes[i] = (x0, y0, x1, y1)
    scores = _distinct_float32(rs.uniform(0.05, 1.0, size=n).astype(np.float32))
    return masks, boxes, scores, cls


# --------------------------------------------------------------------------------------------
# full scene: image + a box head whose decoded, NMS-surviving boxes are the image's instances
# --------------------------------------------------------------------------------------------
def make_scene(seed: int, H: int, W: int, N: int, C: int = 8, anchors: np.ndarray | None = None, n_dup: int = 2,
               cls_th: float = 0.3, iou_th: float = 0.2):
    """Returns (Image, regression [A,4], classification [A,C], anchors [1,A,4]).

    Every instance box of the image is produced by one anchor (regression = inverse BBoxTransform, score and
    class of the instance); `n_dup` lower-scored, shifted duplicates per instance give the NMS real work.
    Duplicates keep integer+0.5 centres and even sizes, and no candidate pair has an IoU within 1e-3 of
    `iou_th`, so the kept set is stable against ulp-level differences in exp().
    """
    if anchors is None:
        anchors = make_anchors(H, W)
    img = make_image(seed, H, W, N, C)
    rs = np.random.RandomState(seed + 7919)
    A = anchors.shape[1]
    an = anchors[0].astype(np.float64)
    a_cy, a_cx = (an[:, 0] + an[:, 2]) / 2, (an[:, 1] + an[:, 3]) / 2
    a_h, a_w = an[:, 2] - an[:, 0], an[:, 3] - an[:, 1]
    regression = rs.normal(0.0, 0.1, size=(A, 4)).astype(np.float32)
    classification = rs.uniform(0.0, 0.05, size=(A, C)).astype(np.float32)

    cand = []   # (box xyxy float64, score, cls)
    for j in range(len(img.rois)):
        x1, y1, x2, y2 = img.rois[j].astype(np.float64)
        cand.append(((x1, y1, x2, y2), float(img.scores[j]), int(img.class_ids[j])))
        for d in range(n_dup):
            dx, dy = int(rs.randint(-4, 5)), int(rs.randint(-4, 5))
            gw, gh = 2 * int(rs.randint(-2, 3)), 2 * int(rs.randint(-2, 3))
            bx = (x1 + dx - gw / 2, y1 + dy - gh / 2, x2 + dx + gw / 2, y2 + dy + gh / 2)
            if bx[0] < 0 or bx[1] < 0 or bx[2] > W - 1 or bx[3] > H - 1 or bx[2] - bx[0] < 4 or bx[3] - bx[1] < 4:
                continue
            cand.append((bx, float(img.scores[j]) * float(rs.uniform(0.45, 0.95)), int(img.class_ids[j])))
    boxes = np.array([c[0] for c in cand], dtype=np.float64).reshape(-1, 4)
    scores = _distinct_float32(np.array([c[1] for c in cand], dtype=np.float32))
    classes = np.array([c[2] for c in cand], dtype=np.int64)
    ok = scores > cls_th + 1e-3
    iou = _iou_matrix(boxes)
    np.fill_diagonal(iou, 0.0)
    near = np.abs(iou - iou_th) < 1e-3
    for i in np.nonzero(near.any(axis=1))[0]:
        if i >= len(img.rois) and ok[i] and (near[i] & ok).any():     # only ever drop duplicates
            ok[i] = False
    used = np.zeros(A, dtype=bool)
    for i in np.nonzero(ok)[0]:
        x1, y1, x2, y2 = boxes[i]
        cy, cx, h, w = (y1 + y2) / 2, (x1 + x2) / 2, y2 - y1, x2 - x1
        cost = np.abs(np.log(a_h / h)) + np.abs(np.log(a_w / w)) + (np.abs(a_cy - cy) / a_h + np.abs(a_cx - cx) / a_w)
        cost[used] = np.inf
        a = int(np.argmin(cost))
        used[a] = True
        regression[a] = np.array([(cy - a_cy[a]) / a_h[a], (cx - a_cx[a]) / a_w[a], np.log(h / a_h[a]), np.log(w / a_w[a])],
                                 dtype=np.float32)
        classification[a, classes[i]] = scores[i]
    return img, regression, classification, anchors


# --------------------------------------------------------------------------------------------
# seeded k-means (BASELINE config 4): embeddings around seed coordinates, label-stable by construction
# --------------------------------------------------------------------------------------------
def _kmeans_fp64(X, init, allow, tol):
    """utils/kmeans.py:55-93 in fp64, recording every iteration's two smallest distances per point.
    Returns (labels, centres, iterations, worst label margin per point, worst relative gap of shift^2 to tol)."""
    X = X.astype(np.float64); c = init.astype(np.float64).copy(); allow = allow.astype(np.float64)
    N = c.shape[0]
    margin = np.full(X.shape[0], np.inf)
    stop_gap = np.inf
    it = 0
    while True:
        d = np.sqrt(((X[:, None, :] - c[None]) ** 2).sum(-1))                  # [M,N]
        part = np.partition(d, 1, axis=1) if N > 1 else np.concatenate([d, np.full_like(d, np.inf)], axis=1)
        best = d.argmin(1)
        d1, d2 = part[:, 0], part[:, 1]
        margin = np.minimum(margin, np.minimum(d2 - d1, np.abs(d1 - allow[best])))
        lab = np.where(d1 < allow[best], best, N)
        new = c.copy()
        shift = 0.0
        for k in np.unique(lab[lab < N]):
            new[k] = X[lab == k].mean(0)
            shift += np.sqrt(((new[k] - c[k]) ** 2).sum())
        c = new
        it += 1
        stop_gap = min(stop_gap, abs(shift * shift - tol) / tol)
        if shift * shift < tol or it > 500:
            return lab, c, it, margin, stop_gap


def make_kmeans_case(seed: int, M: int, N: int, noise: float = 0.004, allow: float = 0.05, tol: float = 1e-4,
                     eps: float = 1e-5):
    """X [M',2] fp32 (M' <= M), initial centres [N,2] fp32, allow [N] fp32 for utils/kmeans.py::kmeans, such that along
    the whole Lloyd trajectory (followed in fp64) every point's nearest centre beats the runner-up by more than `eps`,
    its distance stays more than `eps` away from the allowed distance, and center_shift^2 stays 1 % away from `tol`:
    labels and the iteration count are then independent of fp32 rounding and of the order of the mean's additions
    (centre perturbations ~1e-7).  Points that violate a margin are removed and the trajectory is re-certified."""
    rs = np.random.RandomState(seed)
    cen = np.stack([rs.uniform(0.05, 0.95, N), rs.uniform(0.05, 1.95, N)], axis=1).astype(np.float32)
    X = (cen[rs.randint(0, N, size=M)] + rs.normal(0, noise, size=(M, 2))).astype(np.float32)
    far = rs.rand(M) < 0.01                                                     # outliers beyond every allowed distance
    X[far] += rs.choice([-1.0, 1.0], size=(int(far.sum()), 2)).astype(np.float32) * np.float32(3.0)
    init = (cen + rs.normal(0, noise / 2, size=cen.shape)).astype(np.float32)
    allow_v = np.full(N, allow, dtype=np.float32)
    for _ in range(20):
        lab, c, it, margin, gap = _kmeans_fp64(X, init, allow_v, tol)
        bad = margin <= eps
        if not bad.any():
            if gap < 1e-2:
                raise RuntimeError("kmeans case %d: center_shift^2 within 1%% of tol; pick another seed" % seed)
            return X, init, allow_v, lab.astype(np.int64), c.astype(np.float32), it
        X = X[~bad]
    raise RuntimeError("kmeans case %d could not be certified" % seed)
