"""The four symbols of the reference's utils/image.py on the decode path:
poly_to_mask (:180-185, host rasterisation via cv2 like the reference), compute_iou_for_mask (:188-191),
compute_iou_for_poly (:194-202), is_cover (:205-207).  Mask statistics are popcounts on bit-packed masks
computed by libisg.so.  `fill_polygons` / `polys_to_masks` are the batched device form of poly_to_mask
(`isg_fill_polygons`): all polygons of a call are rasterised by one kernel launch."""
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


FILL_OK, FILL_EMPTY, FILL_OUTSIDE, FILL_OVERFLOW = 0, 1, 2, 3
_FILL_ERRORS = {FILL_OUTSIDE: "a vertex lies outside the %d x %d frame (the device rasteriser does not clip edges)",
                FILL_OVERFLOW: "output capacity exceeded (frame %d x %d)"}


class FilledPolygons:
    """Bit-packed masks of n polygons on the device, as written by `isg_fill_polygons`.

    words: int32 [cap] device buffer; desc: int32 [n,8] (host copy, see include/isg.h); in full-frame mode
    `bits` is the [n,H,ceil(W/32)] view that `isg_mask_nms` / `mask_pair_counts` read."""

    def __init__(self, words, desc, size, full_frame, used=None):
        self.words, self.desc, self.size, self.full_frame = words, desc, (int(size[0]), int(size[1])), full_frame
        self.used = int(words.numel() if used is None else min(used, words.numel()))     # words actually written

    def __len__(self):
        return self.desc.shape[0]

    @property
    def bits(self) -> torch.Tensor:
        if not self.full_frame:
            raise ValueError("bits: only for full_frame=True")
        H, W = self.size
        return self.words[: len(self) * H * ((W + 31) // 32)].view(len(self), H, (W + 31) // 32)

    def _host_words(self):
        if getattr(self, "_host", None) is None:
            self._host = self.words[: self.used].cpu().numpy().view(np.uint32)     # one D2H copy of the packed words
        return self._host

    def mask(self, i, dtype=np.int32):
        """[H,W] 0/1 array of polygon i, equal to poly_to_mask(poly_i, size)."""
        H, W = self.size
        st, x0, y0, rows, wpr, lo, hi, _k = (int(v) for v in self.desc[i])
        m = np.zeros((H, W), dtype=dtype)
        if st == FILL_OK:
            off = (lo & 0xFFFFFFFF) | (hi << 32)
            blk = self._host_words()[off: off + rows * wpr].reshape(rows, wpr)
            px = np.unpackbits(blk.view(np.uint8), axis=1, bitorder="little")           # [rows, 32*wpr]
            x1 = min(W, x0 + 32 * wpr)
            m[y0: y0 + rows, x0: x1] = px[:, : x1 - x0]
        return m

    def masks(self, dtype=np.int32):
        """list of [H,W] 0/1 arrays equal to poly_to_mask(poly, size) (one D2H copy of the packed words)."""
        return [self.mask(i, dtype) for i in range(len(self))]


def fill_polygons(polys, img_size, full_frame=False, dev=None) -> FilledPolygons:
    """Rasterise polygons ([K_i,2] (x,y) arrays, or a device float32 [*,2] tensor with `(start, count)` given as
    polys=(points, start, count)) into an H x W frame exactly like poly_to_mask, on the device.
    Raises ValueError if a polygon has a vertex outside the frame."""
    dev = dev or _dev()
    H, W = int(img_size[0]), int(img_size[1])
    if isinstance(polys, tuple) and torch.is_tensor(polys[0]):
        pts, start, count = polys
        pts = pts.to(dev, torch.float32).contiguous()
        start = torch.as_tensor(start).to(dev, torch.int32).contiguous()
        count = torch.as_tensor(count).to(dev, torch.int32).contiguous()
        n = int(count.numel())
        bound = None
    else:
        arrs = [np.asarray(p, dtype=np.float32).reshape(-1, 2) for p in polys]
        n = len(arrs)
        cnt = np.array([a.shape[0] for a in arrs], dtype=np.int32)
        st = np.zeros(n, dtype=np.int32)
        if n:
            st[1:] = np.cumsum(cnt)[:-1]
        flat = np.concatenate(arrs) if n and cnt.sum() else np.zeros((1, 2), np.float32)
        pts = torch.from_numpy(np.ascontiguousarray(flat)).to(dev)
        start, count = torch.from_numpy(st).to(dev), torch.from_numpy(cnt).to(dev)
        # exact output size of the compact layout: rows x words of every bounding box
        bound = 0
        for a in arrs:
            if a.shape[0]:
                ai = a.astype(np.int32)
                bound += int(ai[:, 1].max() - ai[:, 1].min() + 1) * int((ai[:, 0].max() >> 5) - (ai[:, 0].min() >> 5) + 1)
    if n == 0:
        return FilledPolygons(torch.zeros(1, dtype=torch.int32, device=dev), np.zeros((0, 8), np.int32), (H, W), full_frame)
    Ww = (W + 31) // 32
    if full_frame:
        cap = n * H * Ww
    elif bound is not None:
        cap = max(bound, 1)
    else:
        cap = min(n * H * Ww, 1 << 23)                    # device-resident polygons: 32 MB first, the exact size on overflow
    desc = torch.empty((n, 8), dtype=torch.int32, device=dev)
    total = torch.empty(1, dtype=torch.int64, device=dev)
    while True:
        words = torch.empty(cap, dtype=torch.int32, device=dev)
        call("isg_fill_polygons", ptr(pts), ptr(start), ptr(count), n, H, W, 1 if full_frame else 0, ptr(words), cap, ptr(desc),
             ptr(total), stream_ptr(dev))
        d = desc.cpu().numpy()
        used = n * H * Ww if full_frame else int(total.item())
        if full_frame or bound is not None or used <= cap:
            break
        cap = used                                        # the kernel reports the words it needed: run again with room for all
    for code, msg in _FILL_ERRORS.items():
        if (d[:, 0] == code).any():
            raise ValueError("fill_polygons: polygon %d: " % int(np.nonzero(d[:, 0] == code)[0][0]) + msg % (H, W))
    return FilledPolygons(words, d, (H, W), full_frame, used)


def fill_instances(plan, full_frame=False) -> FilledPolygons:
    """Masks of the polygons the device polygon stage left in a DecodePlan (`run_assign(..., tail="polygons")`),
    without a host round trip: entry b * N + i is instance i of image b (status FILL_EMPTY where no polygon was
    accepted).  Full-frame output is the [B*N, H, ceil(W/32)] input of `nms.mask_nms`."""
    B, N, cap = plan.B, plan.N, plan.cap
    base = torch.arange(B, device=plan.device, dtype=torch.int32)[:, None] * cap
    start = (plan.inst_start + base).reshape(-1)
    count = torch.where(plan.inst_flags == 1, plan.inst_count, torch.zeros_like(plan.inst_count)).reshape(-1)
    return fill_polygons((plan.poly_points.view(-1, 2), start, count), (plan.H, plan.W), full_frame, plan.device)


def polys_to_masks(polys, img_size=None):
    """[poly_to_mask(p, img_size) for p in polys] with one kernel launch and one D2H copy of bit-packed boxes.
    img_size None: every mask is cropped to its own polygon's extent + 1 like poly_to_mask does."""
    arrs = [np.asarray(p, dtype=np.float32).reshape(-1, 2) for p in polys]
    if img_size is not None:
        return fill_polygons(arrs, img_size).masks()
    ext = [(a.astype(np.int32).max(0) + 1)[::-1] for a in arrs]                # per polygon (rows, cols)
    H, W = max(int(e[0]) for e in ext), max(int(e[1]) for e in ext)
    return [m[: e[0], : e[1]].copy() for m, e in zip(fill_polygons(arrs, (H, W)).masks(), ext)]
