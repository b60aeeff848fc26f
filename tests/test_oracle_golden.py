"""Pins the oracle (oracle/) against outputs of the reference itself (tests/golden/, written by
oracle/make_golden.py from the unmodified reference).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import ref_decode as rd
from oracle import ref_kmeans_nms as rk


def test_select_points_matches_reference(golden):
    g = golden("select_points")
    for name in "abcd":
        out = rd.select_points(torch.from_numpy(g["in_" + name]), int(g["k_" + name])).numpy()
        assert np.array_equal(out, g["out_" + name]), name


def test_select_points_k_too_large_raises():
    with pytest.raises(RuntimeError):
        rd.select_points(torch.zeros(4, 4), 17)


def test_nms_hm_matches_reference(golden):
    g = golden("select_points")
    heat = torch.from_numpy(g["heat"])
    assert np.array_equal(rd.nms_hm(heat, 3).numpy(), g["heat_keep3"])
    assert np.array_equal(rd.nms_hm(heat, 5).numpy(), g["heat_keep5"])


@pytest.mark.parametrize("name", ["s0", "s1", "s2", "s3"])
def test_decode_single_matches_reference(golden, name):
    g = golden("decode_single_" + name)
    kp, ae = torch.from_numpy(g["kp"]), torch.from_numpy(g["ae"])
    kp_th = int(g["kp_th"])
    assert np.array_equal(rd.select_points(kp[0], kp_th).numpy(), g["mask"])
    core = rd.group_core(kp[0], ae, g["rois"], kp_th)
    groups = [(p, c) for p, c in rd.instance_points(core["idx"], core["label"], core["centres"], core["whs"], 0.1) if p.shape[0] >= 2]
    assert len(groups) == int(g["n_groups"])
    for i, (pts, ctr) in enumerate(groups):
        assert np.array_equal(pts, g["grp_pts_%d" % i])
        assert np.array_equal(ctr, g["grp_ctr_%d" % i])
    boxes = {"rois": g["rois"], "class_ids": g["class_ids"], "scores": g["scores"]}
    (dets,) = rd.decode_single(kp, ae, boxes, kp_th)
    assert len(dets) == int(g["n_dets"])
    for i, (cls, conf, ctr, poly) in enumerate(dets):
        assert int(cls) == int(g["det_cls_%d" % i])
        assert np.float32(conf) == g["det_conf_%d" % i]
        assert np.array_equal(ctr, g["det_ctr_%d" % i])
        assert np.array_equal(poly, g["det_poly_%d" % i])


def test_decode_single_with_rejections_matches_reference(golden):
    """s4: one polygon rejected by the centre-inside test (aug_group -> None), one instance below obj_pixel_th, and
    label-0 background pixels removed by the ghost filter — all decided by the reference itself"""
    g = golden("decode_single_s4")
    kp, ae = torch.from_numpy(g["kp"]), torch.from_numpy(g["ae"])
    kp_th, n = int(g["kp_th"]), len(g["rois"])
    assert np.array_equal(rd.select_points(kp[0], kp_th).numpy(), g["mask"])
    core = rd.group_core(kp[0], ae, g["rois"], kp_th)
    inst = rd.instance_points(core["idx"], core["label"], core["centres"], core["whs"], 0.1)
    assert int((core["label"] == 0).sum()) > inst[0][0].shape[0] > 0            # the ghost filter removed label-0 pixels
    groups = [(p, c) for p, c in inst if p.shape[0] >= 2]
    assert len(groups) == int(g["n_groups"]) == n - 1                             # one instance below obj_pixel_th
    rejected = []
    for i, (pts, ctr) in enumerate(groups):
        assert np.array_equal(pts, g["grp_pts_%d" % i])
        assert np.array_equal(ctr, g["grp_ctr_%d" % i])
        rejected.append(rd.aug_group(pts, ctr) is None)
    assert rejected == g["grp_rejected"].tolist() and sum(rejected) == 1
    boxes = {"rois": g["rois"], "class_ids": g["class_ids"], "scores": g["scores"]}
    (dets,) = rd.decode_single(kp, ae, boxes, kp_th)
    assert len(dets) == int(g["n_dets"]) == n - 2
    for i, (cls, conf, ctr, poly) in enumerate(dets):
        assert int(cls) == int(g["det_cls_%d" % i]) and np.float32(conf) == g["det_conf_%d" % i]
        assert np.array_equal(ctr, g["det_ctr_%d" % i]) and np.array_equal(poly, g["det_poly_%d" % i])


def test_decode_single_resize_transform_matches_reference(golden):
    """val_trans = resize(target_size 2) + decode.target_size = 2 (utils/tranform.py:157-171, test.py:58)"""
    g = golden("decode_single_resize")
    kp, ae = torch.from_numpy(g["kp"]), torch.from_numpy(g["ae"])
    boxes = {"rois": g["rois"], "class_ids": g["class_ids"], "scores": g["scores"]}
    ts, size = int(g["target_size"]), tuple(int(v) for v in g["img_size"])
    (dets,) = rd.decode_single(kp, ae, boxes, int(g["kp_th"]), scale=ts, img_size=size, resize_target=ts)
    assert len(dets) == int(g["n_dets"]) > 0
    for i, (cls, conf, ctr, poly) in enumerate(dets):
        assert int(cls) == int(g["det_cls_%d" % i]) and np.float32(conf) == g["det_conf_%d" % i]
        assert np.array_equal(ctr, g["det_ctr_%d" % i]) and np.array_equal(poly, g["det_poly_%d" % i])
    assert max(float(g["det_poly_%d" % i][:, 0].max()) for i in range(len(dets))) > kp.shape[-1]   # original-image pixels


def test_decode_ct_hm_matches_reference(golden):
    g = golden("decode_ct_hm")
    cls, idx, conf, wh = rd.decode_ct_hm(torch.from_numpy(g["conf"]), torch.from_numpy(g["cls"]), torch.from_numpy(g["wh"]),
                                         int(g["num_classes"]), int(g["k"]))
    assert len(cls) > 0 and len(cls) < int(g["k"])
    assert np.array_equal(cls, g["keep_cls"]) and np.array_equal(idx, g["keep_idx"])
    assert np.array_equal(conf, g["keep_conf"]) and np.array_equal(wh, g["keep_wh"])


def test_dense_labels_agree_with_sparse(golden):
    g = golden("decode_single_s1")
    kp, ae = torch.from_numpy(g["kp"]), torch.from_numpy(g["ae"])
    core = rd.group_core(kp[0], ae, g["rois"], int(g["kp_th"]))
    score, label = rd.dense_labels(ae, g["rois"])
    yy, xx = core["idx"][:, 0], core["idx"][:, 1]
    assert torch.equal(label[yy, xx], core["label"])
    assert torch.equal(score[yy, xx], core["score"])


@pytest.mark.parametrize("use_tv", [True, False])
def test_decode_boxes_matches_reference(golden, use_tv):
    g = golden("decode_boxes")
    dets = rd.decode_boxes(int(g["H"]), int(g["W"]), torch.from_numpy(g["anchors"]), torch.from_numpy(g["regression"]),
                           torch.from_numpy(g["classification"]), 0.3, 0.2, use_torchvision=use_tv)
    for b, det in enumerate(dets):
        assert np.array_equal(np.asarray(det["rois"]), g["rois_%d" % b])
        assert np.array_equal(np.asarray(det["class_ids"]), g["cls_%d" % b])
        assert np.array_equal(np.asarray(det["scores"]), g["scores_%d" % b])
    assert len(dets[2]["class_ids"]) == 0


def test_kmeans_matches_reference(golden):
    g = golden("kmeans")
    lab, ctr, _ = rk.kmeans(torch.from_numpy(g["X"]), 10, torch.from_numpy(g["init"]), g["allow"])
    assert np.array_equal(lab.numpy(), g["labels"])
    np.testing.assert_allclose(ctr.numpy(), g["centers"], rtol=1e-5, atol=1e-7)
    lab, ctr, _ = rk.kmeans(torch.from_numpy(g["X"] + 1.0), 10, torch.from_numpy(g["init"] + 1.0),
                            np.full(10, 0.002, dtype=np.float32), distance="cosine")
    assert np.array_equal(lab.numpy(), g["labels_cos"])
    np.testing.assert_allclose(ctr.numpy(), g["centers_cos"], rtol=1e-5, atol=1e-7)
    assert np.array_equal(rk.pairwise_distance(torch.from_numpy(g["X"][:40]), torch.from_numpy(g["init"])).numpy(), g["pd"])
    assert np.array_equal(rk.pairwise_cosine(torch.from_numpy(g["X"][:40] + 1.0), torch.from_numpy(g["init"] + 1.0)).numpy(), g["pc"])


def test_kmeans_crowd_matches_reference(golden):
    """BASELINE config 4 size (M ~ 20000, N = 500): labels of the reference itself, bit for bit"""
    g = golden("kmeans_crowd")
    lab, ctr, _ = rk.kmeans(torch.from_numpy(g["X"]), 500, torch.from_numpy(g["init"]), g["allow"])
    assert np.array_equal(lab.numpy(), g["labels"].astype(np.int64))
    np.testing.assert_allclose(ctr.numpy(), g["centers"], rtol=1e-6, atol=1e-7)


def test_py_cpu_nms_matches_reference(golden):
    g = golden("nms")
    for name in "abc":
        keep = rk.py_cpu_nms(g["dets_" + name], float(g["thr_" + name]))
        assert np.array_equal(np.asarray(keep, dtype=np.int64), g["keep_" + name]), name
    # the reference's boxes_nms: ([],[],[]) on empty input, TypeError otherwise (utils/nms.py:45-51)
    assert int(g["boxes_nms_empty"]) == 0 and int(g["boxes_nms_raises"]) == 1
    assert rk.boxes_nms({"class_ids": np.array(()), "rois": np.array(()), "scores": np.array(())}, 0.5) == ([], [], [])


def test_mask_iou_matches_reference(golden):
    g = golden("mask_iou")
    masks, H, W = g["masks"], int(g["H"]), int(g["W"])
    dense = np.unpackbits(masks.view(np.uint8).reshape(len(masks), H, -1), axis=2, bitorder="little")[:, :, :W].astype(np.int32)
    for i in range(len(masks)):
        for j in range(len(masks)):
            assert rk.compute_iou_for_mask(dense[i], dense[j]) == g["iou"][i, j]
            assert rk.is_cover(dense[i], dense[j]) == bool(g["cover"][i, j])
            inter = rk.popcount(masks[i] & masks[j]); uni = rk.popcount(masks[i] | masks[j])
            assert float(inter + 1) / float(uni + 1) == g["iou"][i, j]


def test_anchor_restatement_matches_reference(golden):
    """synth.make_anchors (the generator every box-head test uses) against the reference's Anchors.forward"""
    import hashlib
    from isg_b200 import synth
    g = golden("anchors")
    assert np.array_equal(synth.make_anchors(128, 256), g["a_128x256"])
    for h, w in ((1024, 2048), (512, 1024)):
        a = synth.make_anchors(h, w)
        assert a.shape[1] == int(g["count_%dx%d" % (h, w)])
        assert np.array_equal(np.frombuffer(hashlib.sha256(a.tobytes()).digest(), dtype=np.uint8), g["sha_%dx%d" % (h, w)])
        assert np.array_equal(a[0, g["rows_%dx%d" % (h, w)]], g["vals_%dx%d" % (h, w)])


def test_batched_nms_restatements_vs_torchvision_near_threshold():
    """pairs whose IoU sits within a few fp32 ulps of the threshold: the coordinate trick and the per-class NMS of
    torchvision disagree on some of them; each restatement must follow its own library function exactly"""
    import isg_b200  # noqa: F401
    from isg_b200 import synth
    from oracle import ref_decode as rd
    tvb = pytest.importorskip("torchvision.ops.boxes")
    differ = 0
    for seed in range(3):
        b, s, c = synth.make_near_threshold_boxes(seed, 300, 0.5)
        tb, ts, tc = torch.from_numpy(b), torch.from_numpy(s), torch.from_numpy(c)
        trick = tvb._batched_nms_coordinate_trick(tb, ts, tc, 0.5).numpy()
        vanilla = tvb._batched_nms_vanilla(tb, ts, tc, 0.5).numpy()
        assert np.array_equal(rd.batched_nms_trick_numpy(b, s, c, 0.5), trick)
        assert np.array_equal(rd.batched_nms_numpy(b, s, c, 0.5), vanilla)
        assert np.array_equal(tvb.batched_nms(tb, ts, tc, 0.5).numpy(), trick)      # 600 boxes: numel <= 4000
        differ += len(set(trick.tolist()) ^ set(vanilla.tolist()))
    assert differ > 0      # the inputs do separate the two conventions
