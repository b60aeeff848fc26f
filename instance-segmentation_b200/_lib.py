"""ctypes binding of libisg.so (include/isg.h).  No fallback: a missing library is an error."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libisg.so")

P = C.c_void_p
I = C.c_int
I64 = C.c_int64
F = C.c_float
D = C.c_double
SZ = C.c_size_t



class DecodeStep(C.Structure):
    """isg_decode_step_t (include/isg.h), field for field"""
    _fields_ = [
        ("struct_bytes", I), ("assign", I), ("polygons", I),
        ("B", I), ("H", I), ("W", I), ("img_h", I), ("img_w", I), ("A", I), ("C", I), ("Nmax", I), ("cand_cap", I), ("cap", I),
        ("kp_th", I), ("obj_pixel_th", I), ("nms_convention", I),
        ("cls_th", F), ("ghost_k", F), ("scale", F), ("iou_th", D),
        ("kp", P), ("kp_img_stride", I64), ("ae", P), ("ae_img_stride", I64), ("ae_plane_stride", I64),
        ("anchors", P), ("regression", P), ("classification", P), ("ys", P), ("xs", P),
        ("cand_boxes", P), ("cand_scores", P), ("cand_cls", P), ("cand_anchor", P), ("cand_count", P),
        ("keep", P), ("n_keep", P), ("nms_ws", P), ("nms_ws_bytes", SZ),
        ("rois", P), ("scores", P), ("cls", P), ("n_seeds", P),
        ("thr_key", P), ("topk_ws", P), ("topk_ws_bytes", SZ),
        ("seeds", P), ("ghost", P), ("stats", P),
        ("keepbits", P), ("label_map", P), ("dense_ws", P), ("dense_ws_bytes", SZ),
        ("idx", P), ("count", P), ("label", P),
        ("poly_points", P), ("inst_start", P), ("inst_count", P), ("inst_flags", P), ("inst_internal", P),
        ("img_total", P), ("poly_ws", P), ("poly_ws_bytes", SZ),
        ("main", P), ("side", P), ("fork_event", P), ("join_event", P), ("time_begin", P), ("time_end", P),
        ("split_keep", I),
    ]


ISG_ASSIGN_DENSE, ISG_ASSIGN_SPARSE = 0, 1

# name -> (restype, argtypes); mirrors include/isg.h one to one
PROTOTYPES = {
    "isg_abi_version": (I, []),
    "isg_strerror": (C.c_char_p, [I]),
    "isg_device_supported": (I, [I]),
    "isg_debug_reload_tuning": (None, []),
    "isg_topk_workspace_bytes": (SZ, [I, I, I, I]),
    "isg_topk_threshold": (I, [P, I, I, I, I64, I, P, P, SZ, P]),
    "isg_keep_points": (I, [P, I, I, I, I64, P, P, P, P]),
    "isg_select_points_workspace_bytes": (SZ, [I, I, I, I]),
    "isg_select_points": (I, [P, I, I, I, I64, I, P, P, P, SZ, P]),
    "isg_nms_hm": (I, [P, I, I, I, I, P, P]),
    "isg_compact_points": (I, [P, I, I, I, I, P, P, P]),
    "isg_build_seeds": (I, [P, I, P, I, I, P, P, I, I, F, F, P, P, P]),
    "isg_stats_init": (I, [P, I, I, P]),
    "isg_gather_build_seeds": (I, [P, P, P, P, P, I, I, I, P, P, I, I, F, F, P, P, P, P, P, P, P, P, P]),
    "isg_assign_sparse": (I, [P, I64, I64, P, P, I, P, P, P, I, I, I, I, P, P, P, P, P, P, P]),
    "isg_scatter_labels": (I, [P, P, I, P, I, I, I, P, P]),
    "isg_gather_embeddings": (I, [P, I64, I64, P, P, I, I, I, I, P, P, P, P]),
    "isg_host_device_pointer": (I, [P, P]),
    "isg_assign_dense_workspace_bytes": (SZ, [I, I, I, I]),
    "isg_build_tile_lists": (I, [P, P, I, I, I, I, P, SZ, P]),
    "isg_assign_dense": (I, [P, I64, P, I64, I64, P, P, P, P, I, I, I, I, P, P, P, P, P, P, P, SZ, I, P]),
    "isg_topk_keep": (I, [P, I, I, I, I64, I, P, P, I, P, SZ, P]),
    "isg_assign_labels": (I, [P, I64, I64, P, P, P, I, I, I, I, P, P, P, P, P, SZ, I, P]),
    "isg_gather_labels": (I, [P, P, P, P, I, P, I, I, I, I, P, P, P, P, P]),
    "isg_group_points": (I, [P, P, P, P, I, P, I, I, P, P, P]),
    "isg_decode_boxes": (I, [P, P, P, I, I, I, I, I, F, I, P, P, P, P, P, P]),
    "isg_bbox_transform": (I, [P, P, I, I, I, I, I, P, P]),
    "isg_clip_boxes": (I, [P, I64, I, I, P]),
    "isg_anchor_count": (I64, [I, I, P, I, I]),
    "isg_generate_anchors": (I, [I, I, P, I, P, I, I, P, P]),
    "isg_pack_masks": (I, [P, I, I, I, P, P]),
    "isg_gather_kept": (I, [P, P, P, P, P, I, I, I, P, P, P, P, P]),
    "isg_box_nms_workspace_bytes": (SZ, [I, I]),
    "isg_box_nms": (I, [P, P, P, P, P, I, I, D, I, P, P, P, SZ, P]),
    "isg_mask_nms_workspace_bytes": (SZ, [I]),
    "isg_mask_nms": (I, [P, I, I, I, P, P, P, D, P, P, P, SZ, P]),
    "isg_mask_pair_counts": (I, [P, I, I, I, P, I, P, P]),
    "isg_kmeans_workspace_bytes": (SZ, [I, I, I]),
    "isg_kmeans": (I, [P, I, I, P, P, I, F, I, I, P, P, P, SZ, P]),
    "isg_pairwise": (I, [P, I, P, I, I, I, P, P]),
    "isg_instance_polygons_workspace_bytes": (SZ, [I, I]),
    "isg_instance_polygons": (I, [P, P, P, I, P, P, I, I, I, I, I, I, P, P, P, P, P, P, P, P, SZ, I, P]),
    "isg_fill_polygons": (I, [P, P, P, I, I, I, I, P, SZ, P, P, P]),
    "isg_decode_heads": (I, [P, I, I, I, I, P, P, P, P, P, P, P]),
    "isg_decode_step": (I, [P]),
    "isg_decode_step_bytes": (SZ, []),
    "isg_host_point_in_polygon": (I, [P, I, F, F]),
    "isg_host_internal_points": (I, [P, P, I, P, I, P]),
    "isg_host_centres_inside": (I, [P, P, I, P, P]),
}

ISG_NMS_PLUS1_LE = 0
ISG_NMS_TV_GT = 1
ISG_NMS_TV_TRICK = 2
ISG_NMS_TV_BATCHED = 3
ISG_NMS_MAX_BOXES = 16384
ISG_KMEANS_EUCLIDEAN = 0
ISG_KMEANS_COSINE = 1
ISG_ENOTCONVERGED = -4
SEED_WORDS, GHOST_WORDS, STAT_WORDS = 8, 4, 5
ISG_BOX_XYXY, ISG_BOX_CYCXHW = 0, 1

_lib = None


class IsgError(RuntimeError):
    def __init__(self, code: int, where: str):
        self.code = code
        msg = lib().isg_strerror(code)
        super().__init__("%s failed: %s (code %d)" % (where, msg.decode() if msg else "?", code))


def lib() -> C.CDLL:
    """Load libisg.so (once).  Raises if it has not been built — there is no CPU path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "libisg.so is missing (%s). Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `python instance-segmentation_b200/build.py`; this package has no CPU fallback." % LIB_PATH)
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(handle, name)   # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(code: int, where: str) -> None:
    if code != 0:
        raise IsgError(code, where)


# number of kernels/launches enqueued through this binding (the bench's `gpu_launches` claim)
launch_count = 0

# kernels each entry point enqueues (excluding memsets); an int, or a function of the call's arguments where
# the entry point picks a path by size.  Keep in sync with csrc/.
_LAUNCHES = {
    "isg_topk_threshold": 3, "isg_keep_points": 1, "isg_select_points": 4, "isg_nms_hm": 1,
    "isg_compact_points": 1, "isg_build_seeds": 1, "isg_stats_init": 1, "isg_assign_sparse": 1,
    "isg_gather_build_seeds": 1, "isg_scatter_labels": 1, "isg_gather_embeddings": 1, "isg_build_tile_lists": 1, "isg_assign_dense": lambda a: 1 if a[21] else 2, "isg_gather_labels": 1, "isg_instance_polygons": lambda a: 1 if a[21] else 2, "isg_decode_boxes": 1,
    "isg_gather_kept": 1, "isg_mask_nms": 5, "isg_mask_pair_counts": 1, "isg_pairwise": 1,
    "isg_bbox_transform": 1, "isg_clip_boxes": 1, "isg_generate_anchors": 1, "isg_decode_heads": 1, "isg_pack_masks": 1, "isg_fill_polygons": 1,
    # (idx,label,flag,count,cap,n_seeds,B,Nmax,...): one multisplit kernel unless the seed table is huge
    "isg_group_points": lambda a: 1 if (16 * a[7] + a[7] + 1 + a[4]) * 4 <= 200 * 1024 else 3,
    # (boxes,scores,cls,tiebreak,count,B,cap,...): fused single-CTA kernel for cap <= 1024
    "isg_box_nms": lambda a: 1 if a[6] <= 1024 else 3,
}


def call(name: str, *args) -> None:
    """Call an int-returning entry point and raise IsgError on a non-zero code."""
    global launch_count
    rc = getattr(lib(), name)(*args)
    if rc != 0:
        raise IsgError(rc, name)
    n = _LAUNCHES.get(name, 0)
    launch_count += n(args) if callable(n) else n
