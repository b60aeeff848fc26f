"""Generate tests/golden/*.npz by running the UNMODIFIED reference (read-only at /root/reference) on small
synthetic inputs.  Runs only in the build container (the GPU box has no /root/reference); the fixtures it
writes are committed.  Usage:  python oracle/make_golden.py [--ref /root/reference]

The reference needs three compatibility shims on this image; none touches the
decode arithmetic:  webcolors / skimage stubs (imported, never called) and a uint8->bool cast for
Tensor.masked_select (utils/decode.py:313 passes a uint8 mask, legal in the torch 1.4 the reference pins).
"""
from __future__ import annotations

import argparse
import importlib.util
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")


def load_synth():
    spec = importlib.util.spec_from_file_location("isg_synth", os.path.join(ROOT, "instance-segmentation_b200", "synth.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["isg_synth"] = mod
    spec.loader.exec_module(mod)
    return mod


def install_shims():
    wc = types.ModuleType("webcolors")

    class _RGB:
        red = green = blue = 0
    wc.name_to_rgb = lambda name: _RGB()
    sys.modules["webcolors"] = wc
    sk, skm = types.ModuleType("skimage"), types.ModuleType("skimage.measure")
    skm.find_contours = lambda *a, **k: []
    sk.measure = skm
    sys.modules["skimage"], sys.modules["skimage.measure"] = sk, skm
    _ms = torch.Tensor.masked_select
    torch.Tensor.masked_select = lambda self, m: _ms(self, m.bool() if m.dtype == torch.uint8 else m)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    args = ap.parse_args()
    synth = load_synth()
    install_shims()
    sys.path.insert(0, args.ref)
    os.chdir(args.ref)
    torch.set_num_threads(1)
    from configs import Config, Configer
    from utils import decode as rdecode
    from utils import image as rimage
    from utils.kmeans import kmeans as rkmeans, pairwise_cosine, pairwise_distance
    from utils.nms import py_cpu_nms, boxes_nms
    from utils.tranform import CommonTransforms, TransInfo

    os.makedirs(OUT, exist_ok=True)
    cfg = Config(os.path.join(args.ref, "configs", "decode_cfg.yaml"))
    cfg.draw_flag = False
    tf = CommonTransforms(Configer(configs=os.path.join(args.ref, "configs", "trans_cfg.json")), "val")
    dev = torch.device("cpu")

    # ---- select_points / nms_hm on unstructured maps (incl. the negative-selected quirk) ----------
    rs = np.random.RandomState(7)
    sel_cases = {}
    for name, (h, w, k, loc) in {"a": (33, 47, 200, 0.0), "b": (64, 128, 1500, -3.0), "c": (40, 40, 1600, 1.0),
                                 "d": (17, 130, 1, 0.0)}.items():
        m = synth._distinct_float32(rs.normal(loc, 1.0, size=(h, w)).astype(np.float32))
        out = rdecode.select_points(torch.from_numpy(m), k).numpy()
        sel_cases["in_" + name] = m
        sel_cases["k_" + name] = np.int64(k)
        sel_cases["out_" + name] = out
    heat = rs.randint(0, 6, size=(2, 3, 19, 23)).astype(np.float32)   # ties on purpose: nms_hm is plain equality
    sel_cases["heat"] = heat
    sel_cases["heat_keep3"] = rdecode.nms_hm(torch.from_numpy(heat), 3).numpy()
    sel_cases["heat_keep5"] = rdecode.nms_hm(torch.from_numpy(heat), 5).numpy()
    np.savez_compressed(os.path.join(OUT, "select_points.npz"), **sel_cases)

    # ---- decode_single: polygons + the per-instance point sets handed to aug_group ---------------
    for name, (seed, h, w, n, kp_th) in {"s0": (11, 64, 128, 3, 300), "s1": (12, 128, 256, 6, 1500),
                                         "s2": (13, 192, 320, 12, 20000), "s3": (14, 96, 160, 5, 100)}.items():
        img = synth.make_image(seed, h, w, n)
        cfg.kp_th = kp_th
        captured = []
        orig = rdecode.aug_group

        def spy(pts, center_loc, _o=orig, _c=captured):
            _c.append((np.array(pts, dtype=np.float32), np.array(center_loc, dtype=np.float32).reshape(-1)))
            return _o(pts, center_loc)
        rdecode.aug_group = spy
        try:
            boxes = {"rois": img.rois, "class_ids": img.class_ids, "scores": img.scores}
            mask = rdecode.select_points(torch.from_numpy(img.kp[0]), kp_th).numpy()
            (dets,) = rdecode.decode_single(torch.from_numpy(img.kp), torch.from_numpy(img.ae.copy()), boxes,
                                            TransInfo("/nonexistent.png", (h, w)), tf, cfg, dev)
        finally:
            rdecode.aug_group = orig
        d = dict(kp=img.kp, ae=img.ae, rois=img.rois, class_ids=img.class_ids, scores=img.scores,
                 kp_th=np.int64(kp_th), mask=mask, n_groups=np.int64(len(captured)), n_dets=np.int64(len(dets)))
        for i, (pts, c) in enumerate(captured):
            d["grp_pts_%d" % i] = pts
            d["grp_ctr_%d" % i] = c
        for i, (cls, conf, ctr, poly) in enumerate(dets):
            d["det_cls_%d" % i] = np.int64(cls)
            d["det_conf_%d" % i] = np.float32(conf)
            d["det_ctr_%d" % i] = np.asarray(ctr, dtype=np.float32)
            d["det_poly_%d" % i] = np.asarray(poly, dtype=np.float32)
        np.savez_compressed(os.path.join(OUT, "decode_single_%s.npz" % name), **d)
        print(name, "groups", len(captured), "dets", len(dets), "kept px", int(mask.sum()))

    # ---- decode_single with rejections: an instance whose polygon does not contain its centre (aug_group -> None,
    #      :201-204), an instance below obj_pixel_th (:355), and background pixels outside every box that land on
    #      label 0 (:328) and are removed by the ghost filter (:351-352) ------------------------------------------------
    seed, h, w, n, kp_th = 15, 128, 256, 8, 6000
    img = synth.make_image(seed, h, w, n)
    rs4 = np.random.RandomState(1500)
    kp4 = img.kp[0].copy()
    cx = (img.rois[:, 0] + img.rois[:, 2]) / 2
    ww = img.rois[:, 2] - img.rois[:, 0]
    yy, xx = np.nonzero(img.owner == 2)
    cut = xx >= cx[2] - ww[2] / 4                                   # instance 2 keeps only the left quarter of its outline
    kp4[yy[cut], xx[cut]] = rs4.normal(-6.0, 0.5, size=int(cut.sum())).astype(np.float32)
    yy, xx = np.nonzero(img.owner == 5)
    kp4[yy[1:], xx[1:]] = rs4.normal(-6.0, 0.5, size=yy.size - 1).astype(np.float32)   # instance 5 keeps one pixel
    # stray peaks outside every box: all-zero membership row -> label 0 (:328); those outside instance 0's ghost band
    # are removed by the filter (:351-352), the ones inside the band (0.5w..0.6w from its centre) stay
    yy, xx = np.mgrid[0:h, 0:w]
    covered = np.zeros((h, w), dtype=bool)
    for x1, y1, x2, y2 in img.rois:
        covered |= (xx >= x1 - 1) & (xx <= x2 + 1) & (yy >= y1 - 1) & (yy <= y2 + 1)
    cy0, hh0 = (img.rois[0, 1] + img.rois[0, 3]) / 2, img.rois[0, 3] - img.rois[0, 1]
    band = ~covered & (np.abs(xx - cx[0]) < 0.58 * ww[0]) & (np.abs(yy - cy0) < 0.58 * hh0)
    far = ~covered & ((np.abs(xx - cx[0]) > 0.7 * ww[0]) | (np.abs(yy - cy0) > 0.7 * hh0))
    for region, cnt in ((band, 3), (far, 12)):
        ry, rx = np.nonzero(region)
        pick = rs4.choice(ry.size, size=min(cnt, ry.size), replace=False)
        kp4[ry[pick], rx[pick]] = rs4.uniform(4.0, 5.0, size=pick.size).astype(np.float32)
    n_band, n_far = int(min(3, band.sum())), int(min(12, far.sum()))
    kp4 = synth._distinct_float32(kp4)[None]
    cfg.kp_th = kp_th
    captured = []
    orig = rdecode.aug_group

    def spy4(pts, center_loc, _o=orig, _c=captured):
        out = _o(pts, center_loc)
        _c.append((np.array(pts, dtype=np.float32), np.array(center_loc, dtype=np.float32).reshape(-1), out is None))
        return out
    rdecode.aug_group = spy4
    try:
        boxes = {"rois": img.rois, "class_ids": img.class_ids, "scores": img.scores}
        mask = rdecode.select_points(torch.from_numpy(kp4[0]), kp_th).numpy()
        (dets,) = rdecode.decode_single(torch.from_numpy(kp4), torch.from_numpy(img.ae.copy()), boxes,
                                        TransInfo("/nonexistent.png", (h, w)), tf, cfg, dev)
    finally:
        rdecode.aug_group = orig
    d = dict(kp=kp4, ae=img.ae, rois=img.rois, class_ids=img.class_ids, scores=img.scores, kp_th=np.int64(kp_th), mask=mask,
             n_groups=np.int64(len(captured)), n_dets=np.int64(len(dets)),
             grp_rejected=np.array([c[2] for c in captured], dtype=bool))
    for i, (pts, c, _) in enumerate(captured):
        d["grp_pts_%d" % i] = pts
        d["grp_ctr_%d" % i] = c
    for i, (cls, conf, ctr, poly) in enumerate(dets):
        d["det_cls_%d" % i] = np.int64(cls)
        d["det_conf_%d" % i] = np.float32(conf)
        d["det_ctr_%d" % i] = np.asarray(ctr, dtype=np.float32)
        d["det_poly_%d" % i] = np.asarray(poly, dtype=np.float32)
    assert len(dets) < len(captured) < n and n_band > 0 and n_far > 0, (len(dets), len(captured), n, n_band, n_far)
    np.savez_compressed(os.path.join(OUT, "decode_single_s4.npz"), **d)
    print("s4 boxes", n, "groups", len(captured), "dets", len(dets), "kept px", int(mask.sum()))

    # ---- decode_single under a resize validation transform (utils/tranform.py:157-171) with decode.target_size = 2
    #      (test.py:58): the network sees the half-size image, polygons come out in original-image pixels --------------
    import json
    import tempfile
    trans = json.load(open(os.path.join(args.ref, "configs", "trans_cfg.json")))
    trans["val_trans"] = {"trans_seq": ["resize"], "resize": {"target_size": 2}}
    with tempfile.NamedTemporaryFile("w", suffix=".json", delete=False) as f:
        json.dump(trans, f)
    tf2 = CommonTransforms(Configer(configs=f.name), "val")
    os.unlink(f.name)
    seed, h, w, n, kp_th = 16, 96, 192, 5, 2500
    img = synth.make_image(seed, h, w, n)
    cfg.kp_th = kp_th
    saved_ts = rdecode.target_size
    rdecode.target_size = 2
    captured = []

    def spy5(pts, center_loc, _o=orig, _c=captured):
        _c.append((np.array(pts, dtype=np.float32), np.array(center_loc, dtype=np.float32).reshape(-1)))
        return _o(pts, center_loc)
    rdecode.aug_group = spy5
    try:
        boxes = {"rois": img.rois, "class_ids": img.class_ids, "scores": img.scores}
        (dets,) = rdecode.decode_single(torch.from_numpy(img.kp), torch.from_numpy(img.ae.copy()), boxes,
                                        TransInfo("/nonexistent.png", (2 * h, 2 * w)), tf2, cfg, dev)
    finally:
        rdecode.aug_group = orig
        rdecode.target_size = saved_ts
    d = dict(kp=img.kp, ae=img.ae, rois=img.rois, class_ids=img.class_ids, scores=img.scores, kp_th=np.int64(kp_th),
             img_size=np.array([2 * h, 2 * w], dtype=np.int64), target_size=np.int64(2),
             n_groups=np.int64(len(captured)), n_dets=np.int64(len(dets)))
    for i, (pts, c) in enumerate(captured):
        d["grp_pts_%d" % i] = pts
        d["grp_ctr_%d" % i] = c
    for i, (cls, conf, ctr, poly) in enumerate(dets):
        d["det_cls_%d" % i] = np.int64(cls)
        d["det_conf_%d" % i] = np.float32(conf)
        d["det_ctr_%d" % i] = np.asarray(ctr, dtype=np.float32)
        d["det_poly_%d" % i] = np.asarray(poly, dtype=np.float32)
    np.savez_compressed(os.path.join(OUT, "decode_single_resize.npz"), **d)
    print("resize groups", len(captured), "dets", len(dets))

    # ---- decode_ct_hm (:254-285): peaks of a centre heat map -> per-class boxes -> py_cpu_nms(0.5) --------------------
    rs6 = np.random.RandomState(61)
    h, w, ncls, k = 48, 80, 3, 60
    conf = synth._distinct_float32(rs6.uniform(0.0, 1.0, size=(h, w)).astype(np.float32))
    clsm = rs6.randint(0, ncls, size=(h, w)).astype(np.int64)
    whm = rs6.uniform(6.0, 30.0, size=(2, h, w)).astype(np.float32)
    out = rdecode.decode_ct_hm(torch.from_numpy(conf), torch.from_numpy(clsm), torch.from_numpy(whm), ncls, k, tf,
                               TransInfo("/nonexistent.png", (h, w)))
    np.savez_compressed(os.path.join(OUT, "decode_ct_hm.npz"), conf=conf, cls=clsm, wh=whm, num_classes=np.int64(ncls),
                        k=np.int64(k), keep_cls=np.asarray(out[0], dtype=np.int64), keep_idx=np.asarray(out[1], dtype=np.int64).reshape(-1, 2),
                        keep_conf=np.asarray(out[2], dtype=np.float32), keep_wh=np.asarray(out[3], dtype=np.float32).reshape(-1, 2))
    print("decode_ct_hm kept", len(out[0]))

    # ---- decode_boxes ------------------------------------------------------------------------------
    H, W, C = 128, 256, 8
    anchors = synth.make_anchors(H, W)
    regs, clss = [], []
    for b, K in enumerate((60, 25)):
        _, r, c = synth.make_box_head(100 + b, H, W, C, K, anchors)
        regs.append(r); clss.append(c)
    regs.append(regs[0].copy()); clss.append(np.full_like(clss[0], 0.01))      # image with no candidate
    regression = torch.from_numpy(np.stack(regs)); classification = torch.from_numpy(np.stack(clss))
    x = torch.zeros((3, 3, H, W))
    dets = rdecode.decode_boxes(x, torch.from_numpy(anchors), regression.clone(), classification.clone(), 0.3, 0.2)
    d = dict(anchors=anchors, regression=regression.numpy(), classification=classification.numpy(), H=np.int64(H), W=np.int64(W))
    for b, det in enumerate(dets):
        d["rois_%d" % b] = np.asarray(det["rois"]); d["cls_%d" % b] = np.asarray(det["class_ids"]); d["scores_%d" % b] = np.asarray(det["scores"])
        print("decode_boxes img", b, "kept", len(det["class_ids"]))
    np.savez_compressed(os.path.join(OUT, "decode_boxes.npz"), **d)

    # ---- Anchors.forward (utils/utils.py:366-450): full table at a small shape, digests + samples at the bench shapes
    import hashlib
    from utils.utils import Anchors as RAnchors
    d = {}
    small = RAnchors()(torch.zeros(1, 3, 128, 256)).numpy()
    d["a_128x256"] = small
    d["a_half_128x256"] = RAnchors()(torch.zeros(1, 3, 128, 256), dtype=torch.float16).numpy()
    d["a_custom_96x160"] = RAnchors(anchor_scale=3., pyramid_levels=[2, 3, 4], scales=[1.0, 1.5], ratios=[(1.0, 1.0), (2.0, 0.5)])(
        torch.zeros(2, 3, 96, 160)).numpy()                      # H % 16 == 0 is not required by the reference, only W (:416)
    d["a_ragged_100x64"] = RAnchors(pyramid_levels=[3, 4])(torch.zeros(1, 3, 100, 64)).numpy()   # H not divisible by the stride
    for (h, w) in ((1024, 2048), (512, 1024)):
        a = RAnchors()(torch.zeros(1, 3, h, w)).numpy()
        d["sha_%dx%d" % (h, w)] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)
        d["count_%dx%d" % (h, w)] = np.int64(a.shape[1])
        pick = np.random.RandomState(h).randint(0, a.shape[1], size=64)
        d["rows_%dx%d" % (h, w)] = pick
        d["vals_%dx%d" % (h, w)] = a[0, pick]
    np.savez_compressed(os.path.join(OUT, "anchors.npz"), **d)
    print("anchors", small.shape, int(d["count_1024x2048"]))

    # ---- f4: the 1x1 heads of the reference's EfficientDecoder (models/efficient.py:508-510,536-541) -------------------------
    from models.efficient import EfficientDecoder
    torch.manual_seed(4)
    dec_mod = EfficientDecoder([320, 112, 40, 24, 16], {"kp": 1, "ae": 4, "tan": 2}).eval()
    xf = torch.randn(2, 16, 24, 40)
    with torch.no_grad():
        heads = [dec_mod.__getattr__(h)(xf) for h in dec_mod.headers.keys()]      # the loop at :537-539
    np.savez_compressed(os.path.join(OUT, "heads.npz"), x=xf.numpy(), kp=heads[0].numpy(), ae=heads[1].numpy(),
                        w_kp=dec_mod.kp.weight.detach().numpy(), b_kp=dec_mod.kp.bias.detach().numpy(),
                        w_ae=dec_mod.ae.weight.detach().numpy(), b_ae=dec_mod.ae.bias.detach().numpy())
    print("heads", [tuple(h.shape) for h in heads])

    # ---- kmeans / pairwise ---------------------------------------------------------------------------
    rs = np.random.RandomState(21)
    cen = rs.uniform(0, 1, size=(10, 2)).astype(np.float32)
    X = (cen[rs.randint(0, 10, size=600)] + rs.normal(0, 0.02, size=(600, 2))).astype(np.float32)
    X[::50] += 0.5                                                           # outliers
    init = (cen + rs.normal(0, 0.01, size=cen.shape)).astype(np.float32)
    allow = np.full(10, 0.08, dtype=np.float32)
    lab, ctr = rkmeans(torch.from_numpy(X), 10, torch.from_numpy(init), allow)
    labc, ctrc = rkmeans(torch.from_numpy(X + 1.0), 10, torch.from_numpy(init + 1.0), np.full(10, 0.002, dtype=np.float32), distance="cosine")
    np.savez_compressed(os.path.join(OUT, "kmeans.npz"), X=X, init=init, allow=allow, labels=lab.numpy(), centers=ctr.numpy(),
                        labels_cos=labc.numpy(), centers_cos=ctrc.numpy(),
                        pd=pairwise_distance(torch.from_numpy(X[:40]), torch.from_numpy(init)).numpy(),
                        pc=pairwise_cosine(torch.from_numpy(X[:40] + 1.0), torch.from_numpy(init + 1.0)).numpy())

    # BASELINE config 4 size: M ~ 20000 embeddings, N = 500 seeds, allow 0.05; margins certified by the generator
    Xc, initc, allowc, _, _, _ = synth.make_kmeans_case(4, 20000, 500)
    labc4, ctrc4 = rkmeans(torch.from_numpy(Xc), 500, torch.from_numpy(initc), allowc)
    np.savez_compressed(os.path.join(OUT, "kmeans_crowd.npz"), X=Xc, init=initc, allow=allowc,
                        labels=labc4.numpy().astype(np.int16), centers=ctrc4.numpy())
    print("kmeans crowd", Xc.shape, "outliers", int((labc4 == 500).sum()))

    # ---- py_cpu_nms / boxes_nms -----------------------------------------------------------------------
    d = {}
    for name, (seed, n, thr) in {"a": (31, 300, 0.5), "b": (32, 64, 0.3), "c": (33, 1, 0.5)}.items():
        dets_np = synth.make_nms_boxes(seed, n, extent=600.0, thr=thr, plus1=True)
        d["dets_" + name] = dets_np
        d["thr_" + name] = np.float64(thr)
        d["keep_" + name] = np.asarray(py_cpu_nms(dets_np, thr), dtype=np.int64)
    d["boxes_nms_empty"] = np.int64(len(boxes_nms({"class_ids": np.array(()), "rois": np.array(()), "scores": np.array(())}, 0.5)[0]))
    try:
        boxes_nms({"class_ids": np.array([1, 1]), "rois": np.zeros((2, 4), np.float32), "scores": np.array([0.5, 0.4], np.float32)}, 0.5)
        d["boxes_nms_raises"] = np.int64(0)
    except TypeError:
        d["boxes_nms_raises"] = np.int64(1)
    np.savez_compressed(os.path.join(OUT, "nms.npz"), **d)

    # ---- mask IoU ---------------------------------------------------------------------------------------
    masks, _, _, _ = synth.make_masks(41, 6, 48, 70, C=2)
    dense = np.unpackbits(masks.view(np.uint8).reshape(6, 48, -1), axis=2, bitorder="little")[:, :, :70].astype(np.int32)
    iou = np.array([[rimage.compute_iou_for_mask(dense[i], dense[j]) for j in range(6)] for i in range(6)], dtype=np.float64)
    cov = np.array([[rimage.is_cover(dense[i], dense[j]) for j in range(6)] for i in range(6)], dtype=bool)
    np.savez_compressed(os.path.join(OUT, "mask_iou.npz"), masks=masks, iou=iou, cover=cov, H=np.int64(48), W=np.int64(70))
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
