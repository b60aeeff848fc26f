"""The symbols of the reference's utils/utils.py on the decode path:
BBoxTransform (:318-346), ClipBoxes (:349-363), Anchors (:366-450), generate_coordinates (:453-458).  The arithmetic
runs in libisg.so; the classes keep the reference's nn.Module call surface."""
from __future__ import annotations

import ctypes
import itertools

import numpy as np
import torch
import torch.nn as nn

from .. import _lib, engine
from .._lib import call
from ..engine import ptr, require_cuda, stream_ptr


def _cuda_dev(t: torch.Tensor) -> torch.device:
    if not t.is_cuda:
        raise RuntimeError("isg_b200 has no CPU path: BBoxTransform/ClipBoxes need CUDA tensors")
    return require_cuda(t.device)


class BBoxTransform(nn.Module):
    def forward(self, anchors, regression):
        """anchors [1|B,A,4] (y1,x1,y2,x2), regression [B,A,4] (dy,dx,dh,dw) -> [B,A,4] (xmin,ymin,xmax,ymax)."""
        dev = _cuda_dev(regression)
        reg = engine.as_f32_planes(regression, dev).contiguous()
        B, A = reg.shape[0], reg.shape[1]
        anc = engine.as_f32_planes(anchors, dev)
        if anc.dim() == 3 and anc.shape[0] == B and B > 1:
            out = torch.empty((B, A, 4), dtype=torch.float32, device=dev)
            for b in range(B):   # per-image anchors: one launch each
                call("isg_bbox_transform", ptr(anc[b].contiguous()), ptr(reg[b]), 1, A, 0, 1, 1, ptr(out[b]), stream_ptr(dev))
            return out
        anc = anc.reshape(-1, 4).contiguous()
        out = torch.empty((B, A, 4), dtype=torch.float32, device=dev)
        call("isg_bbox_transform", ptr(anc), ptr(reg), B, A, 0, 1, 1, ptr(out), stream_ptr(dev))
        return out


class ClipBoxes(nn.Module):
    def __init__(self):
        super(ClipBoxes, self).__init__()

    def forward(self, boxes, img):
        """Clamp boxes [B,A,4] into the image, IN PLACE like the reference (:357-361), and return them."""
        dev = _cuda_dev(boxes)
        _, _, height, width = img.shape
        if boxes.dtype != torch.float32 or not boxes.is_contiguous():
            raise RuntimeError("ClipBoxes expects a contiguous float32 tensor (it clips in place)")
        call("isg_clip_boxes", ptr(boxes), boxes.numel() // 4, int(height), int(width), stream_ptr(dev))
        return boxes


class Anchors(nn.Module):
    """:366-450 — multi-level anchor table [1,A,4] (y1,x1,y2,x2) for `image` [B,C,H,W], generated on the image's
    CUDA device by isg_generate_anchors, bit-identical to the reference's numpy construction.  Same constructor
    arguments, per-(shape, device) cache and ValueError as the reference."""

    def __init__(self, anchor_scale=4., pyramid_levels=None, **kwargs):
        super().__init__()
        self.anchor_scale = anchor_scale
        self.pyramid_levels = [3, 4, 5, 6, 7] if pyramid_levels is None else pyramid_levels
        self.strides = kwargs.get('strides', [2 ** x for x in self.pyramid_levels])
        self.scales = np.array(kwargs.get('scales', [2 ** 0, 2 ** (1.0 / 3.0), 2 ** (2.0 / 3.0)]))
        self.ratios = kwargs.get('ratios', [(1.0, 1.0), (1.4, 0.7), (0.7, 1.4)])
        self.last_anchors = {}
        self.last_shape = None

    def forward(self, image, dtype=torch.float32):
        image_shape = image.shape[2:]
        if image_shape == self.last_shape and image.device in self.last_anchors:      # :401-402
            return self.last_anchors[image.device]
        if self.last_shape is None or self.last_shape != image_shape:
            self.last_shape = image_shape
        dev = _cuda_dev(image)
        half = dtype == torch.float16                                                   # :407-410
        H, W = int(image_shape[0]), int(image_shape[1])
        pairs = list(itertools.product(self.scales, self.ratios))                       # :415
        half_sizes = []
        for stride in self.strides:
            if W % stride != 0:                                                         # :416-417
                raise ValueError('input size must be divided by the stride.')
            for scale, ratio in pairs:
                base_anchor_size = self.anchor_scale * stride * scale                   # :418-420, python/numpy float64
                half_sizes += [float(base_anchor_size * ratio[0] / 2.0), float(base_anchor_size * ratio[1] / 2.0)]
        n_levels, per_cell = len(self.strides), len(pairs)
        strides = (ctypes.c_int * n_levels)(*[int(v) for v in self.strides])
        sizes = (ctypes.c_double * len(half_sizes))(*half_sizes)
        A = int(_lib.lib().isg_anchor_count(H, W, strides, n_levels, per_cell))
        if A <= 0:
            raise ValueError("unsupported anchor configuration (levels <= 8, anchors per cell <= 16, positive sizes)")
        out = torch.empty((1, A, 4), dtype=torch.float16 if half else torch.float32, device=dev)
        call("isg_generate_anchors", H, W, strides, n_levels, sizes, per_cell, 1 if half else 0, ptr(out), stream_ptr(dev))
        self.last_anchors[image.device] = out                                           # :447
        return out


def generate_coordinates():
    """:453-458 — [2,1024,2048] CPU tensor: channel 0 = y in [0,1], channel 1 = x in [0,2]."""
    xm = torch.linspace(0, 2, 2048).view(1, 1, -1).expand(1, 1024, 2048)
    ym = torch.linspace(0, 1, 1024).view(1, -1, 1).expand(1, 1024, 2048)
    return torch.cat((ym, xm), 0)
