"""The three symbols of the reference's utils/utils.py that the decode path uses:
BBoxTransform (:318-346), ClipBoxes (:349-363), generate_coordinates (:453-458).  The arithmetic runs in
libisg.so; the classes keep the reference's nn.Module call surface."""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import engine
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


def generate_coordinates():
    """:453-458 — [2,1024,2048] CPU tensor: channel 0 = y in [0,1], channel 1 = x in [0,2]."""
    xm = torch.linspace(0, 2, 2048).view(1, 1, -1).expand(1, 1024, 2048)
    ym = torch.linspace(0, 1, 1024).view(1, -1, 1).expand(1, 1024, 2048)
    return torch.cat((ym, xm), 0)
