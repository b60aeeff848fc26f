"""CPU: the C-ABI library loads and exports every symbol include/isg.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "isg.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(isg_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    g.build()
    import isg_b200  # noqa: F401
    from isg_b200 import _lib
    return _lib


def test_header_declares_functions():
    syms = header_symbols()
    assert len(syms) >= 25 and "isg_assign_dense" in syms and "isg_box_nms" in syms and "isg_kmeans" in syms


def test_library_exports_every_header_symbol(built):
    handle = ctypes.CDLL(built.LIB_PATH)
    for name in header_symbols():
        assert hasattr(handle, name), name


def test_binding_table_matches_header(built):
    assert sorted(built.PROTOTYPES) == header_symbols()
    lib = built.lib()
    assert lib.isg_abi_version() == 6
    assert lib.isg_strerror(0) == b"ok" and lib.isg_strerror(-1) == b"invalid argument"


def test_host_side_queries_need_no_gpu(built):
    lib = built.lib()
    assert lib.isg_decode_step_bytes() == ctypes.sizeof(built.DecodeStep)      # the ctypes mirror of isg_decode_step_t
    step = built.DecodeStep()
    assert lib.isg_decode_step(ctypes.byref(step)) == -1                         # struct_bytes not set
    assert lib.isg_topk_workspace_bytes(8, 1024, 2048, 20000) >= 8 * (3 * 2048 * 4 + 8 * 20000 * 4)
    assert lib.isg_topk_workspace_bytes(1, 0, 4, 1) == 0
    assert lib.isg_select_points_workspace_bytes(1, 64, 64, 10) >= lib.isg_topk_workspace_bytes(1, 64, 64, 10) + 4
    assert lib.isg_box_nms_workspace_bytes(2, 1000) > 2 * 1000 * 16 * 8
    assert lib.isg_kmeans_workspace_bytes(100, 10, 2) > 0
    assert lib.isg_mask_nms_workspace_bytes(100) > 0
    strides = (ctypes.c_int * 5)(8, 16, 32, 64, 128)
    assert lib.isg_anchor_count(1024, 2048, strides, 5, 9) == 392832      # A at 1024x2048
    assert lib.isg_anchor_count(512, 1024, strides, 5, 9) == 98208
    assert lib.isg_anchor_count(100, 64, strides, 2, 9) == 9 * (12 * 8 + 6 * 4)   # ragged height: ceil((H - s/2) / s) rows
    assert lib.isg_anchor_count(64, 64, strides, 9, 9) == -1


def test_bad_arguments_are_rejected_without_touching_the_device(built):
    lib = built.lib()
    assert lib.isg_topk_threshold(None, 1, 4, 4, 16, 1, None, None, 0, None) == -1
    assert lib.isg_nms_hm(None, 1, 4, 4, 3, None, None) == -1
    assert lib.isg_box_nms(None, None, None, None, None, 1, 10, 0.5, 0, None, None, None, 0, None) == -1
    assert lib.isg_kmeans(None, 1, 2, None, None, 1, 1e-4, 0, 10, None, None, None, 0, None) == -1


def test_missing_library_fails_loudly(built, monkeypatch):
    monkeypatch.setattr(built, "_lib", None)
    monkeypatch.setattr(built, "LIB_PATH", "/nonexistent/libisg.so")
    with pytest.raises(ImportError):
        built.lib()
