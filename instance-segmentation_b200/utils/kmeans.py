"""Drop-in for the reference's utils/kmeans.py (kmeans :16-93, pairwise_distance :96-109, pairwise_cosine
:112-130), computed by libisg.so.  Same signatures; `device` must be a CUDA device (there is no CPU path —
the reference's default torch.device('cpu') is replaced by the current CUDA device)."""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib
from .._lib import IsgError, call
from ..engine import check_device, ptr, require_cuda, stream_ptr

_METRIC = {"euclidean": _lib.ISG_KMEANS_EUCLIDEAN, "cosine": _lib.ISG_KMEANS_COSINE}
max_iterations = 100000   # the reference loops without bound (:55); a bound turns a livelock into an error


def _resolve(device) -> torch.device:
    if device is None or torch.device(device).type == "cpu":
        device = "cuda"
    dev = require_cuda(device)
    check_device(dev)
    return dev


def kmeans(X, num_clusters, cluster_centers, allow_distances, distance='euclidean', tol=1e-4, device=None):
    """Seeded Lloyd iterations; label `num_clusters` marks points farther than allow_distances[nearest].
    Returns (labels int64 [M] w.r.t. the pre-update centres of the last iteration, centres fp32 [N,D])."""
    if distance not in _METRIC:
        raise NotImplementedError
    dev = _resolve(device)
    Xd = torch.as_tensor(X).float().to(dev).contiguous()                               # :42-45
    allow = torch.from_numpy(np.asarray(allow_distances)).float().to(dev).contiguous()  # :48
    centers = torch.as_tensor(cluster_centers).float().to(dev).contiguous().clone()    # :51
    M, D = Xd.shape
    N = int(num_clusters)
    if centers.shape != (N, D) or allow.shape != (N,):
        raise ValueError("cluster_centers must be [num_clusters, D] and allow_distances [num_clusters]")
    labels = torch.empty(M, dtype=torch.int32, device=dev)
    lib = _lib.lib()
    ws_bytes = int(lib.isg_kmeans_workspace_bytes(M, N, D))
    ws = torch.empty(ws_bytes + 256, dtype=torch.uint8, device=dev)
    off = (-ws.data_ptr()) % 256
    import ctypes
    iters = ctypes.c_int(0)
    rc = lib.isg_kmeans(ptr(Xd), M, D, ptr(centers), ptr(allow), N, float(np.float32(tol)), _METRIC[distance],
                        int(max_iterations), ptr(labels), ctypes.addressof(iters), ws.data_ptr() + off, ws_bytes,
                        stream_ptr(dev))
    _lib.launch_count += 1          # the whole Lloyd loop is one cooperative launch
    if rc != 0:
        raise IsgError(rc, "isg_kmeans")
    kmeans.last_iterations = iters.value
    return labels.long(), centers


kmeans.last_iterations = 0


def _pairwise(data1, data2, device, metric):
    dev = _resolve(device)
    a = torch.as_tensor(data1).float().to(dev).contiguous()
    b = torch.as_tensor(data2).float().to(dev).contiguous()
    out = torch.empty((a.shape[0], b.shape[0]), dtype=torch.float32, device=dev)
    call("isg_pairwise", ptr(a), a.shape[0], ptr(b), b.shape[0], a.shape[1], metric, ptr(out), stream_ptr(dev))
    return out


def pairwise_distance(data1, data2, device=None):
    """[M,N] euclidean distances."""
    return _pairwise(data1, data2, device, _lib.ISG_KMEANS_EUCLIDEAN)


def pairwise_cosine(data1, data2, device=None):
    """[M,N] cosine distances, squeezed like the reference (:129)."""
    return _pairwise(data1, data2, device, _lib.ISG_KMEANS_COSINE).squeeze()
