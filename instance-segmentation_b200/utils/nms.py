"""Drop-in for the reference's utils/nms.py (py_cpu_nms :11-39, boxes_nms :42-65) plus the mask-IoU NMS that
BASELINE.json config 5 names (IoU of utils/image.py:188-191 inside the greedy loop of utils/nms.py:23-37).
Computed by libisg.so: bitmask suppression matrix + chunked greedy scan."""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib
from .._lib import call
from ..engine import check_device, ptr, require_cuda, stream_ptr

device = None   # CUDA device for host inputs (defaults to the current device)


def _dev(t=None) -> torch.device:
    if isinstance(t, torch.Tensor) and t.is_cuda:
        dev = require_cuda(t.device)
    else:
        dev = require_cuda(device if device is not None else "cuda")
    check_device(dev)
    return dev


def _run_box_nms(boxes, scores, cls, thr, convention, dev):
    n = boxes.shape[0]       # more than ISG_NMS_MAX_BOXES: the library's tiled large-set path (no suppression matrix)
    lib = _lib.lib()
    count = torch.tensor([n], dtype=torch.int32, device=dev)
    keep = torch.empty(n, dtype=torch.int32, device=dev)
    n_keep = torch.empty(1, dtype=torch.int32, device=dev)
    ws_bytes = int(lib.isg_box_nms_workspace_bytes(1, n))
    ws = torch.empty(ws_bytes + 256, dtype=torch.uint8, device=dev)
    off = (-ws.data_ptr()) % 256
    call("isg_box_nms", ptr(boxes), ptr(scores), ptr(cls), 0, ptr(count), 1, n, float(thr), convention, ptr(keep),
         ptr(n_keep), ws.data_ptr() + off, ws_bytes, stream_ptr(dev))
    k = int(n_keep.item())
    return keep[:k].cpu().numpy().astype(np.int64)


def py_cpu_nms(dets, thresh):
    """dets [n,5] (x1,y1,x2,y2,score) -> list of kept indices in pick order.  Fast R-CNN "+1" convention,
    survivor iff IoU <= thresh; arithmetic in fp32 (the dtype of the reference's call site, utils/decode.py:277)."""
    d = torch.as_tensor(dets)
    dev = _dev(d)
    if d.shape[0] == 0:
        return []
    d = d.float().to(dev)
    boxes = d[:, :4].contiguous()
    scores = d[:, 4].contiguous()
    keep = _run_box_nms(boxes, scores, None, thresh, _lib.ISG_NMS_PLUS1_LE, dev)
    return [k for k in keep]


def boxes_nms(dets, thresh):
    """Per-class py_cpu_nms over a decode_boxes-style dict, survivors merged in descending confidence:
    (cls_ids, boxes, confs) lists.  NOTE: this is the behaviour utils/nms.py:42-65 INTENDS; the reference itself
    returns ([],[],[]) for empty input and raises TypeError at :51 for anything else (and has no caller)."""
    cls_ids = np.asarray(dets["class_ids"])
    if len(np.unique(cls_ids)) <= 0:
        return [], [], []
    dev = _dev()
    rois = np.asarray(dets["rois"], dtype=np.float32).reshape(-1, 4)
    confs = np.asarray(dets["scores"], dtype=np.float32).reshape(-1)
    uniq, inv = np.unique(cls_ids, return_inverse=True)
    keep = _run_box_nms(torch.from_numpy(rois).to(dev).contiguous(), torch.from_numpy(confs).to(dev).contiguous(),
                        torch.from_numpy(inv.astype(np.int32)).to(dev).contiguous(), thresh, _lib.ISG_NMS_PLUS1_LE, dev)
    return [cls_ids[i] for i in keep], [rois[i] for i in keep], [confs[i] for i in keep]


def mask_nms(masks, scores, class_ids=None, thresh=0.5, bboxes=None):
    """Greedy mask NMS.  masks: bit-packed int32/uint32 [n,H,ceil(W/32)] (torch CUDA tensor or numpy; see
    utils.image.pack_masks), scores [n], class_ids [n] or None (class agnostic), bboxes int32 [n,4]
    (x0,y0,x1,y1 inclusive) or None.  Returns kept indices (int64) in pick order."""
    if isinstance(masks, np.ndarray):
        masks = torch.from_numpy(masks.view(np.int32) if masks.dtype == np.uint32 else masks)
    dev = _dev(masks)
    m = masks.to(dev).contiguous()
    if m.dtype != torch.int32:
        raise TypeError("masks must be bit-packed 32-bit words")
    n, H, Ww = m.shape
    if n == 0:
        return np.zeros(0, dtype=np.int64)
    if n > _lib.ISG_NMS_MAX_BOXES:
        raise RuntimeError("mask NMS supports at most %d masks per call" % _lib.ISG_NMS_MAX_BOXES)
    sc = torch.as_tensor(np.asarray(scores, dtype=np.float32) if not isinstance(scores, torch.Tensor) else scores).float().to(dev).contiguous()
    cl = None
    if class_ids is not None:
        c = class_ids.cpu().numpy() if isinstance(class_ids, torch.Tensor) else np.asarray(class_ids)
        cl = torch.from_numpy(np.unique(c, return_inverse=True)[1].astype(np.int32)).to(dev).contiguous()
    bb = None
    if bboxes is not None:
        bb = torch.as_tensor(bboxes).to(torch.int32).to(dev).contiguous()
    lib = _lib.lib()
    keep = torch.empty(n, dtype=torch.int32, device=dev)
    n_keep = torch.empty(1, dtype=torch.int32, device=dev)
    ws_bytes = int(lib.isg_mask_nms_workspace_bytes(n))
    ws = torch.empty(ws_bytes + 256, dtype=torch.uint8, device=dev)
    off = (-ws.data_ptr()) % 256
    call("isg_mask_nms", ptr(m), n, H, Ww, ptr(bb), ptr(sc), ptr(cl), float(thresh), ptr(keep), ptr(n_keep),
         ws.data_ptr() + off, ws_bytes, stream_ptr(dev))
    return keep[:int(n_keep.item())].cpu().numpy().astype(np.int64)
