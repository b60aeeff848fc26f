"""CPU: host-side logic of the drop-in modules (polygon glue, generator guarantees, API surface)."""
import inspect

import numpy as np
import pytest
import torch

from helpers import DecodeCfg, IdentityTransforms, TransInfo


@pytest.fixture(scope="module")
def dec():
    import isg_b200  # noqa: F401
    from isg_b200.utils import decode
    return decode


def test_call_surface_matches_reference(dec):
    """names and argument order of SURVEY.md §8b"""
    from isg_b200.utils import kmeans, nms
    sig = lambda f: list(inspect.signature(f).parameters)
    assert sig(dec.decode_output) == ["inputs", "outs", "infos", "transforms", "decode_cfg", "device"]
    assert sig(dec.decode_single) == ["kp_heat", "ae_mat", "boxes", "info", "transforms", "decode_cfg", "device"]
    assert sig(dec.decode_boxes) == ["x", "anchors", "regression", "classification", "threshold", "iou_threshold"]
    assert sig(dec.group_kp) == ["hm_kp", "hm_ae", "transforms", "center_whs", "center_indexes", "center_cls", "center_confs",
                                 "info", "decode_cfg", "device"]
    assert sig(dec.select_points) == ["mat", "k"] and sig(dec.nms_hm) == ["heat", "kernel"]
    assert sig(dec.decode_ct_hm) == ["conf_mat", "cls_mat", "wh", "num_classes", "cls_th", "transforms", "info"]
    assert sig(kmeans.kmeans) == ["X", "num_clusters", "cluster_centers", "allow_distances", "distance", "tol", "device"]
    assert sig(kmeans.pairwise_distance) == ["data1", "data2", "device"] and sig(kmeans.pairwise_cosine) == ["data1", "data2", "device"]
    assert sig(nms.py_cpu_nms) == ["dets", "thresh"] and sig(nms.boxes_nms) == ["dets", "thresh"]
    assert dec.base_dir == "" and dec.target_size == 1 and dec.xym.shape == (2, 1024, 2048)


def test_cartesian2polar_matches_scalar_loop(dec):
    """the vectorised form against the reference's per-point control flow (utils/decode.py:96-112)"""
    rs = np.random.RandomState(0)
    pts = rs.randint(0, 50, size=(400, 2)).astype(np.float32)
    for c in (np.array([25.0, 25.0], np.float32), np.array([24.5, 10.5], np.float32), pts[7].copy()):
        want = []
        with np.errstate(divide="ignore", invalid="ignore"):
            for p in pts:
                d_x, d_y = tuple(p - c)
                if d_x == 0 and d_y > 0:
                    seta = np.pi / 2
                elif d_x == 0 and d_y < 0:
                    seta = 3 * np.pi / 2
                else:
                    seta = np.arctan(d_y / d_x)
                    if d_x < 0:
                        seta = seta + np.pi
                    elif d_x > 0 and d_y < 0:
                        seta = seta + 2 * np.pi
                want.append(np.array([[seta, np.sqrt(d_x ** 2 + d_y ** 2)]], dtype=np.float32))
        want = np.vstack(want)
        got = dec.cartesian2polar(pts, c)
        assert np.array_equal(got, want, equal_nan=True)


@pytest.mark.parametrize("name", ["s0", "s1", "s2", "s3"])
def test_aug_group_reproduces_reference_polygons(dec, golden, name):
    """host polygon glue fed with the point sets the reference handed to its own aug_group"""
    g = golden("decode_single_" + name)
    polys = [g["det_poly_%d" % i] for i in range(int(g["n_dets"]))]
    out = []
    for i in range(int(g["n_groups"])):
        p = dec.aug_group(g["grp_pts_%d" % i], g["grp_ctr_%d" % i])
        if p is not None:
            out.append(p)
    assert len(out) == len(polys)
    for a, b in zip(out, polys):
        assert np.array_equal(a, b)


def test_polygon_area_predicate_equals_full_canvas(dec):
    from isg_b200.utils import image
    rs = np.random.RandomState(3)
    for _ in range(300):
        k = rs.randint(1, 30)
        p = (rs.randint(0, 40, size=(k, 2)) + rs.randint(0, 300, size=2)).astype(np.float32)
        assert dec._polygon_area_is_zero(p) == (image.poly_to_mask(p).sum() == 0)


def test_degenerate_polygon_is_dropped(dec):
    # centre outside the point set -> pointPolygonTest fails -> None (utils/decode.py:201-204)
    pts = np.array([[10, 10], [20, 10], [20, 20], [10, 20]], np.float32)
    assert dec.aug_group(pts, np.array([50.0, 50.0], np.float32)) is None
    assert dec.aug_group(pts, np.array([15.0, 15.0], np.float32)) is not None


def test_no_cpu_path(dec):
    with pytest.raises(RuntimeError):
        dec.select_points(torch.zeros(8, 8), 3)
    with pytest.raises(RuntimeError):
        dec.decode_output(torch.zeros(1, 3, 8, 8), ((torch.zeros(1, 1, 8, 8), torch.zeros(1, 4, 8, 8), None), torch.zeros(1, 4, 4),
                                                    torch.zeros(1, 4, 2), torch.zeros(1, 4, 4)), [TransInfo("x", (8, 8))],
                          IdentityTransforms(), DecodeCfg(), torch.device("cpu"))


def test_generator_guarantees():
    import isg_b200  # noqa: F401
    from isg_b200 import synth
    a = synth.make_image(5, 96, 160, 5)
    b = synth.make_image(5, 96, 160, 5)
    assert np.array_equal(a.kp, b.kp) and np.array_equal(a.ae, b.ae) and np.array_equal(a.rois, b.rois)   # deterministic
    assert np.unique(a.kp).size == a.kp.size                                                              # tie-free
    c = (a.rois[:, :2] + a.rois[:, 2:]) / 2
    assert np.all(c - np.floor(c) == 0.5) and np.all((a.rois[:, 2:] - a.rois[:, :2]) % 2 == 0)
    assert np.unique(a.scores).size == a.scores.size and np.all(np.diff(a.scores) < 0)
    x = synth._distinct_float32(np.array([1.0, 1.0, -0.0, 0.0, 1.0, np.float32(1.0000001)], np.float32))
    assert np.unique(x).size == 6 and np.all(np.argsort(x, kind="stable") == np.argsort(np.array([1, 1, -0.0, 0.0, 1, 1.0000001], np.float32), kind="stable"))
    anc = synth.make_anchors(128, 256)
    assert anc.shape == (1, 9 * (16 * 32 + 8 * 16 + 4 * 8 + 2 * 4 + 1 * 2), 4) and anc.flags["C_CONTIGUOUS"]


def test_scene_box_head_decodes_to_the_image_boxes():
    """oracle decode_boxes on a synthetic scene returns (a subset of) the image's instance boxes, exactly"""
    import isg_b200  # noqa: F401
    from isg_b200 import synth
    from oracle import ref_decode as rd
    img, reg, cls, anc = synth.make_scene(11, 128, 256, 6)
    det = rd.decode_boxes(128, 256, torch.from_numpy(anc), torch.from_numpy(reg)[None], torch.from_numpy(cls)[None], 0.3, 0.2)[0]
    assert 1 <= len(det["class_ids"]) <= 6
    ctr = (det["rois"][:, :2] + det["rois"][:, 2:]) / 2
    assert np.all(ctr - np.floor(ctr) == 0.5)
    empty = synth.make_scene(12, 128, 256, 0)
    assert (empty[2].max(axis=1) > 0.3).sum() == 0


def test_fast_polygon_path_equals_per_instance_aug_group(dec):
    """the batched host polygon path against the per-instance reference-shaped path, on a full-size image"""
    import isg_b200  # noqa: F401
    from isg_b200 import synth
    from oracle import ref_decode as rd
    img = synth.make_image(77, 512, 1024, 40)
    core = rd.group_core(torch.from_numpy(img.kp[0]), torch.from_numpy(img.ae), img.rois, 20000)
    groups = rd.instance_points(core["idx"], core["label"], core["centres"], core["whs"], 0.1)
    offsets = np.concatenate([[0], np.cumsum([len(p) for p, _ in groups])]).astype(np.int32)
    points = np.concatenate([p for p, _ in groups]).astype(np.float32)
    n = len(groups)
    cls, conf = np.arange(n, dtype=np.int64), np.linspace(0.9, 0.4, n).astype(np.float32)
    c1, f1, ctr1, p1 = dec._polygons_for_image_fast(points, offsets, n, core["centres"], cls, conf, 2)
    c2, f2, ctr2, p2 = dec._polygons_for_image(points, offsets, n, core["centres"], core["whs"], cls, conf, IdentityTransforms(),
                                               TransInfo("x", (512, 1024)), DecodeCfg(), True)
    assert len(p1) == len(p2) > 10
    assert list(c1) == list(c2) and list(f1) == list(f2)
    for a, b in zip(ctr1, ctr2):
        assert np.array_equal(a, b)
    for a, b in zip(p1, p2):
        assert np.array_equal(a, b)


def test_host_point_in_polygon_equals_cv2():
    """libisg's host restatement of cv2.pointPolygonTest (measureDist=False, fp32 contour) against cv2 itself,
    including points on edges / vertices and degenerate contours"""
    import ctypes
    import cv2
    import isg_b200  # noqa: F401
    from isg_b200 import _lib
    lib = _lib.lib()
    rs = np.random.RandomState(0)
    n_checked = 0
    for trial in range(400):
        K = int(rs.randint(1, 40))
        pts = rs.randint(0, 24, size=(K, 2)).astype(np.float32)
        if trial % 3 == 0:
            pts += rs.choice([0.0, 0.5], size=(K, 2)).astype(np.float32)
        tests = [pts[rs.randint(0, K)], pts.mean(axis=0), (pts[0] + pts[-1]) / 2]
        tests += [rs.uniform(-2, 26, size=2).astype(np.float32) for _ in range(6)]
        tests += [rs.randint(0, 24, size=2).astype(np.float32) for _ in range(6)]
        for t in tests:
            want = cv2.pointPolygonTest(pts, (np.float32(t[0]), np.float32(t[1])), False)
            got = lib.isg_host_point_in_polygon(pts.ctypes.data, K, float(t[0]), float(t[1]))
            assert got == int(want), (pts, t, got, want)
            n_checked += 1
    assert n_checked > 5000


def test_host_internal_points_equal_reference_search(dec):
    import isg_b200  # noqa: F401
    from isg_b200 import _lib
    lib = _lib.lib()
    rs = np.random.RandomState(1)
    segs, ctrs = [], []
    for _ in range(120):
        K = int(rs.randint(2, 30))
        base = rs.randint(0, 400, size=2)
        # ring-like and random clouds, sorted row-major like the device emits them
        p = (base + rs.randint(0, 30, size=(K, 2))).astype(np.float32)
        p = p[np.lexsort((p[:, 0], p[:, 1]))]
        segs.append(p)
        ctrs.append((p.mean(axis=0) + rs.uniform(-8, 8, size=2)).astype(np.float32))
    offsets = np.concatenate([[0], np.cumsum([len(p) for p in segs])]).astype(np.int32)
    points = np.ascontiguousarray(np.concatenate(segs), dtype=np.float32)
    centers = np.ascontiguousarray(np.stack(ctrs), dtype=np.float32)
    internal = np.empty_like(centers)
    assert lib.isg_host_internal_points(points.ctypes.data, offsets.ctypes.data, len(segs), centers.ctypes.data, 2, internal.ctypes.data) == 0
    n_fallback = 0
    for i, (p, c) in enumerate(zip(segs, ctrs)):
        want = dec.find_internal_point(p, c)
        assert np.array_equal(internal[i], np.asarray(want, dtype=np.float32)), i
        n_fallback += not np.array_equal(want, c)
    assert n_fallback > 10


def test_assembly_from_device_polygon_buffers(dec):
    """decode_output's host side for the device polygon stage: slicing of the read-back buffers, box-order of the
    instances, empty images, and the flag-2 path (an instance the device left unsorted is finished by aug_group)"""
    from types import SimpleNamespace
    B, N, cap = 3, 4, 64
    ring = lambda cx, cy, r, n: np.stack([cx + r * np.cos(np.linspace(0, 2 * np.pi, n, endpoint=False)),
                                          cy + r * np.sin(np.linspace(0, 2 * np.pi, n, endpoint=False))], 1).round().astype(np.float32)
    p0, p1, p2 = ring(20, 20, 8, 12), ring(50, 30, 6, 10), ring(30, 30, 10, 16)
    pts = np.zeros((B, cap, 2), np.float32)
    # image 0: instance 1 allocated before instance 0 (blocks are handed out in completion order), instance 2 invalid
    pts[0, 0:10], pts[0, 10:22] = p1, p0
    rs = np.random.RandomState(0)
    raw2 = p2[rs.permutation(len(p2))]                                 # image 2: one instance left unsorted (flag 2)
    pts[2, 5:5 + len(raw2)] = raw2
    plan = SimpleNamespace(
        img_total=torch.tensor([22, 0, 21], dtype=torch.int32),
        inst_start=torch.tensor([[10, 0, 22, 0], [0, 0, 0, 0], [5, 0, 0, 0]], dtype=torch.int32),
        inst_count=torch.tensor([[12, 10, 1, 0], [0, 0, 0, 0], [16, 0, 0, 0]], dtype=torch.int32),
        inst_flags=torch.tensor([[1, 1, 0, 0], [0, 0, 0, 0], [2, 0, 0, 0]], dtype=torch.uint8),
        poly_points=torch.from_numpy(pts))
    rois = np.zeros((B, N, 4), np.float32)
    rois[0, 0], rois[0, 1], rois[0, 2] = (10, 10, 30, 30), (42, 22, 58, 38), (0, 0, 4, 4)
    rois[2, 0] = (18, 18, 42, 42)
    scores = np.linspace(0.9, 0.1, B * N, dtype=np.float32).reshape(B, N)
    cls = np.arange(B * N, dtype=np.int32).reshape(B, N) % 5
    dets = dec._dets_from_device_polygons(plan, B, np.array([3, 0, 1]), rois, scores, cls, DecodeCfg())
    assert [len(d) for d in dets] == [2, 0, 1]
    (c0, f0, k0, g0), (c1, f1, k1, g1) = dets[0]
    assert np.array_equal(g0, p0) and np.array_equal(g1, p1) and (int(c0), int(c1)) == (0, 1)
    assert f0 == scores[0, 0] and np.array_equal(k0, np.array([20, 20], np.float32)) and np.array_equal(k1, np.array([50, 30], np.float32))
    # flag 2: same polygon as aug_group on the raw set (sorted by angle about the centre)
    want = dec.aug_group(raw2.copy(), np.array([30, 30], np.float32))
    assert want is not None and np.array_equal(dets[2][0][3], want)


def test_dets_json_round_trip(tmp_path):
    """`{epoch}_dets.json` / `_infos.json` (reference utils/eval_util.py:23-32,65-70): numpy scalars and arrays
    serialise to plain JSON and load back as lists."""
    import importlib
    eval_util = importlib.import_module("isg_b200.utils.eval_util")
    dets = [[(np.int64(3), np.float32(0.75), np.array([4.0, 5.0], np.float32), np.array([[1, 2], [3, 4], [5, 1]], np.float32))], []]
    infos = [("/a/b_leftImg8bit.png", (8, 16)), ("/a/c.png", (8, 16))]
    eval_util.save_dets(dets, infos, str(tmp_path), 12)
    d2, i2 = eval_util.load_dets(str(tmp_path), 12)
    assert d2 == [[[3, 0.75, [4.0, 5.0], [[1.0, 2.0], [3.0, 4.0], [5.0, 1.0]]]], []]
    assert i2 == [["/a/b_leftImg8bit.png", [8, 16]], ["/a/c.png", [8, 16]]]


def test_peak_test_threshold_free_form_equals_select_points():
    """DESIGN.md 4.1 (split form, next step): for a SELECTED pixel every larger neighbour is selected too, so
    keep(p) = selected(p) and p >= every in-image neighbour (raw values) and (p >= 0 or every in-image neighbour is
    selected) - what a filter pass can evaluate per candidate before the exact threshold is known (raw local maximum +
    smallest neighbour).  Checked against the oracle's select_points (utils/decode.py:71-85) on maps with positive,
    negative and mixed selected values."""
    import torch
    import torch.nn.functional as F
    from oracle import ref_decode as rd
    g = torch.Generator().manual_seed(5)
    for H, W, k, shift in ((40, 56, 300, 0.0), (33, 47, 900, -3.0), (24, 24, 500, 0.4), (16, 64, 1024, -1.0), (31, 29, 1, 2.0)):
        mat = torch.randn((H, W), generator=g) + shift
        want = rd.select_points(mat, k).bool()
        thr = torch.topk(mat.reshape(-1), k).values[-1]
        sel = mat >= thr
        ninf, pinf = float("-inf"), float("inf")
        nb_max = F.max_pool2d(F.pad(mat[None, None], (1, 1, 1, 1), value=ninf), 3, stride=1)[0, 0]       # includes the centre
        nb_min = -F.max_pool2d(F.pad(-mat[None, None], (1, 1, 1, 1), value=ninf), 3, stride=1)[0, 0]     # smallest value of the window
        local_max = mat >= nb_max
        all_selected = nb_min >= thr                                   # the centre is selected, so this is about the neighbours
        got = sel & local_max & ((mat >= 0) | all_selected)
        assert torch.equal(got, want), (H, W, k, shift)
