"""f4 - the producer side of the decode: the 1x1 heads of the reference's EfficientDecoder
(models/efficient.py:508-510 build one nn.Conv2d(16, c, kernel_size=1) per header {"kp": 1, "ae": 4, "tan": 2}; :536-541
apply them) evaluated for inference by libisg.so: one pass over the decoder's last feature map, only the five channels the
decode reads (`tan` is dropped: decode_output ignores it, utils/decode.py:447), written in the planar layout the decode
kernels consume.  No CPU path."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from .._lib import call
from ..engine import check_device, ptr, require_cuda, stream_ptr


class InferenceHeads(nn.Module):
    """Built from the reference decoder's `kp` and `ae` head convolutions (any object with .weight [c,Cin,1,1] and .bias [c]).
    forward(x [B,Cin,H,W] CUDA fp32) -> (kp [B,1,H,W], ae [B,4,H,W], None) - the `kp_out` triple decode_output unpacks."""

    def __init__(self, kp_conv, ae_conv):
        super().__init__()
        wk, wa = kp_conv.weight.detach().float().cpu(), ae_conv.weight.detach().float().cpu()
        if wk.dim() != 4 or wk.shape[0] != 1 or wk.shape[2:] != (1, 1) or wa.shape[0] != 4 or wa.shape[1:] != wk.shape[1:]:
            raise ValueError("expected the 1x1 kp (1 channel) and ae (4 channels) heads of EfficientDecoder")
        zeros = lambda n: torch.zeros(n)
        bk = kp_conv.bias.detach().float().cpu() if kp_conv.bias is not None else zeros(1)
        ba = ae_conv.bias.detach().float().cpu() if ae_conv.bias is not None else zeros(4)
        self.cin = int(wk.shape[1])
        self._w_kp = np.ascontiguousarray(wk.reshape(1, self.cin).numpy())
        self._w_ae = np.ascontiguousarray(wa.reshape(4, self.cin).numpy())
        self._b_kp = np.ascontiguousarray(bk.numpy())
        self._b_ae = np.ascontiguousarray(ba.numpy())

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("isg_b200 has no CPU path: InferenceHeads needs a CUDA tensor")
        dev = require_cuda(x.device)
        check_device(dev)
        x = x.float().contiguous()
        B, Cin, H, W = x.shape
        if Cin != self.cin:
            raise ValueError("feature map has %d channels, the heads expect %d" % (Cin, self.cin))
        kp = torch.empty((B, 1, H, W), dtype=torch.float32, device=dev)
        ae = torch.empty((B, 4, H, W), dtype=torch.float32, device=dev)
        call("isg_decode_heads", ptr(x), B, Cin, H, W, self._w_kp.ctypes.data, self._b_kp.ctypes.data, self._w_ae.ctypes.data,
             self._b_ae.ctypes.data, ptr(kp), ptr(ae), stream_ptr(dev))
        return kp, ae, None
