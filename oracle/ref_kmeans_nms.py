"""ORACLE — test infrastructure only.  CPU restatement of utils/kmeans.py, utils/nms.py and the mask-IoU
formula of utils/image.py.  See oracle/ref_decode.py for the usage rules and the parity status
(pinned against reference-run fixtures by tests/test_oracle_golden.py)."""
from __future__ import annotations

import numpy as np
import torch


# ----------------------------------------------------------------------------------------------
# a8 — k-means
# ----------------------------------------------------------------------------------------------
def pairwise_distance(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """utils/kmeans.py:96-109."""
    return ((a.unsqueeze(1) - b.unsqueeze(0)) ** 2.0).sum(dim=-1).sqrt()


def pairwise_cosine(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """utils/kmeans.py:112-130."""
    A = a.unsqueeze(1); B = b.unsqueeze(0)
    An = A / A.norm(dim=-1, keepdim=True)
    Bn = B / B.norm(dim=-1, keepdim=True)
    return 1 - (An * Bn).sum(dim=-1).squeeze()


def kmeans(X, num_clusters, cluster_centers, allow_distances, distance="euclidean", tol=1e-4, max_iter=100000):
    """utils/kmeans.py:16-93.  The per-cluster Python loop (:66-75) is restated with a scatter-add in
    fp64 (the reference's fp32 `mean` agrees to ~1e-7 relative).  Returns (labels i64 [M], centres f32 [N,D], iters)."""
    if distance == "euclidean":
        dist_fn = pairwise_distance
    elif distance == "cosine":
        dist_fn = pairwise_cosine
    else:
        raise NotImplementedError
    X = X.float()
    allow = torch.from_numpy(np.asarray(allow_distances))
    state = cluster_centers.clone().float()
    it = 0
    while True:
        dis = dist_fn(X, state)
        if dis.dim() == 1:
            dis = dis.view(X.shape[0], state.shape[0])
        min_d, choice = torch.min(dis, dim=1)                                    # :57
        ok = (min_d < allow[choice]).long()                                      # :60
        choice = choice * ok + (1 - ok) * num_clusters                           # :61
        sums = torch.zeros((num_clusters + 1, X.shape[1]), dtype=torch.float64)
        sums.index_add_(0, choice, X.double())
        cnt = torch.bincount(choice, minlength=num_clusters + 1)[:num_clusters]
        new_state = state.clone()
        nz = cnt > 0
        new_state[nz] = (sums[:num_clusters][nz] / cnt[nz].double()[:, None]).float()   # :70-71
        shift = torch.zeros(1, dtype=torch.float32)
        per = (new_state - state).pow(2).sum(dim=1).sqrt()                       # :72
        for k in torch.nonzero(nz).flatten().tolist():                           # cluster order, fp32 adds
            shift += per[k]
        state = new_state
        it += 1
        if shift ** 2 < tol:                                                     # :90
            break
        if it >= max_iter:
            raise RuntimeError("oracle kmeans: no convergence")
    return choice, state, it


# ----------------------------------------------------------------------------------------------
# a9 / a10 — greedy box NMS
# ----------------------------------------------------------------------------------------------
def py_cpu_nms(dets: np.ndarray, thresh: float):
    """utils/nms.py:11-39 restated with a `dead` array instead of the shrinking `order` array; every
    elementwise operation keeps the reference's fp32 form (+1 areas, survivor iff ovr <= thresh)."""
    dets = np.asarray(dets)
    x1, y1, x2, y2, scores = dets[:, 0], dets[:, 1], dets[:, 2], dets[:, 3], dets[:, 4]
    areas = (x2 - x1 + 1) * (y2 - y1 + 1)                                        # :19
    order = scores.argsort()[::-1]                                               # :20
    dead = np.zeros(len(dets), dtype=bool)
    keep = []
    for pos, i in enumerate(order):
        if dead[i]:
            continue
        keep.append(i)
        rest = order[pos + 1:]
        rest = rest[~dead[rest]]
        w = np.maximum(0.0, np.minimum(x2[i], x2[rest]) - np.maximum(x1[i], x1[rest]) + 1)   # :26-31
        h = np.maximum(0.0, np.minimum(y2[i], y2[rest]) - np.maximum(y1[i], y1[rest]) + 1)
        inter = w * h
        with np.errstate(divide="ignore", invalid="ignore"):
            ovr = inter / (areas[i] + areas[rest] - inter)                       # :33
        dead[rest[~(ovr <= thresh)]] = True                                      # :35-36
    return keep


def boxes_nms(dets: dict, thresh: float):
    """utils/nms.py:42-65 — the INTENDED behaviour (the reference raises TypeError at :51 for any non-empty
    input): per-class py_cpu_nms, survivors merged in descending confidence."""
    cls_ids = np.asarray(dets["class_ids"])
    if len(np.unique(cls_ids)) <= 0:
        return [], [], []
    rois = np.asarray(dets["rois"], dtype=np.float32).reshape(-1, 4)
    scores = np.asarray(dets["scores"], dtype=np.float32)
    picked = []
    for c in np.unique(cls_ids):
        ids = np.nonzero(cls_ids == c)[0]
        d = np.concatenate([rois[ids], scores[ids, None]], axis=1).astype(np.float32)
        picked.extend(ids[k] for k in py_cpu_nms(d, thresh))
    picked = sorted(picked, key=lambda i: -float(scores[i]))
    return [cls_ids[i] for i in picked], [rois[i] for i in picked], [scores[i] for i in picked]


# ----------------------------------------------------------------------------------------------
# a12 — mask IoU and the mask NMS composition
# ----------------------------------------------------------------------------------------------
def compute_iou_for_mask(mask1: np.ndarray, mask2: np.ndarray) -> float:
    """utils/image.py:188-191."""
    return float((mask1 & mask2).sum() + 1) / float((mask1 | mask2).sum() + 1)


def is_cover(mask1: np.ndarray, mask2: np.ndarray) -> bool:
    """utils/image.py:205-207."""
    inter = (mask1 * mask2).sum()
    return bool(mask1.sum() == inter or mask2.sum() == inter)


_POP8 = np.array([bin(i).count("1") for i in range(256)], dtype=np.int64)


def popcount(a: np.ndarray) -> int:
    return int(_POP8[a.view(np.uint8)].sum())


def mask_nms(masks: np.ndarray, scores: np.ndarray, classes, thresh: float):
    """Greedy loop of utils/nms.py:20-37 with IoU = (|A&B|+1)/(|A|B|+1) (utils/image.py:188-191) on bit-packed
    masks [n,H,Ww] uint32; class aware (like utils/decode.py:400) when `classes` is given."""
    n = masks.shape[0]
    area = np.array([popcount(masks[i]) for i in range(n)], dtype=np.int64)
    order = np.asarray(scores).argsort()[::-1]
    dead = np.zeros(n, dtype=bool)
    keep = []
    for pos, i in enumerate(order):
        if dead[i]:
            continue
        keep.append(int(i))
        for j in order[pos + 1:]:
            if dead[j] or (classes is not None and classes[i] != classes[j]):
                continue
            inter = popcount(masks[i] & masks[j])
            iou = float(inter + 1) / float(area[i] + area[j] - inter + 1)
            if not (iou <= thresh):
                dead[j] = True
    return keep
