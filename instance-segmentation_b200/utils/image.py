"""The four symbols of the reference's utils/image.py on the decode path:
poly_to_mask (:180-185, host rasterisation via cv2 like the reference), compute_iou_for_mask (:188-191),
compute_iou_for_poly (:194-202), is_cover (:205-207).  Mask statistics are popcounts on bit-packed masks
computed by libisg.so."""
from __future__ import annotations

import numpy as np
import torch

from .. import engine
from .._lib import call
from ..engine import ptr, require_cuda, stream_ptr

device = None   # CUDA device used when the inputs are host arrays (defaults to the current device)


def poly_to_mask(poly, img_size=None):
    import cv2
    poly = poly.astype(np.int32)
    if img_size is None:
        img_size = (poly.max(0) + 1)[::-1]
    mask = np.zeros(img_size, dtype=np.int32)
    return cv2.fillPoly(mask, [poly], 1)


def _dev():
    return require_cuda(device if device is not None else "cuda")


def pack_masks(dense, dev=None) -> torch.Tensor:
    """dense 0/1 masks [n,H,W] (numpy or torch, any integer/bool dtype) -> int32 [n,H,ceil(W/32)] on the device."""
    dev = dev or _dev()
    t = torch.as_tensor(np.ascontiguousarray(dense) if isinstance(dense, np.ndarray) else dense)
    if t.dim() == 2:
        t = t[None]
    t = (t != 0).to(torch.uint8).to(dev).contiguous()
    n, H, W = t.shape
    bits = torch.empty((n, H, (W + 31) // 32), dtype=torch.int32, device=dev)
    call("isg_pack_masks", ptr(t), n, H, W, ptr(bits), stream_ptr(dev))
    return bits


def mask_pair_counts(bits: torch.Tensor, pairs) -> np.ndarray:
    """(|A&B|, |A|B|) for index pairs into bit-packed masks; int64 [n_pairs,2] on the host."""
    dev = bits.device
    n, H, Ww = bits.shape
    p = torch.as_tensor(np.asarray(pairs, dtype=np.int32).reshape(-1, 2)).to(dev).contiguous()
    out = torch.empty((p.shape[0], 2), dtype=torch.int64, device=dev)
    call("isg_mask_pair_counts", ptr(bits), n, H, Ww, ptr(p), p.shape[0], ptr(out), stream_ptr(dev))
    return out.cpu().numpy()


def compute_iou_for_mask(mask1, mask2):
    """(|m1 & m2| + 1) / (|m1 | m2| + 1) for 0/1 masks of equal shape."""
    bits = pack_masks(np.stack((np.asarray(mask1), np.asarray(mask2))))
    inter, union = mask_pair_counts(bits, [(0, 1)])[0]
    return float(inter + 1) / float(union + 1)


def compute_iou_for_poly(poly1, poly2, img_size=None):
    if img_size is None:
        img_size = (np.max(np.vstack((poly1.max(0), poly2.max(0))), axis=0).astype(np.int32) + 1)[::-1]
    return compute_iou_for_mask(poly_to_mask(poly1, img_size), poly_to_mask(poly2, img_size))


def is_cover(mask1, mask2):
    """True iff one 0/1 mask contains the other."""
    bits = pack_masks(np.stack((np.asarray(mask1), np.asarray(mask2))))
    c = mask_pair_counts(bits, [(0, 1), (0, 0), (1, 1)])
    inter, a, b = int(c[0, 0]), int(c[1, 0]), int(c[2, 0])
    return a == inter or b == inter
