"""GPU parity of the device polygon rasteriser (isg_fill_polygons) against the library call the
reference makes (cv2.fillPoly through poly_to_mask, utils/image.py:180-185) — bit-exact."""
import numpy as np
import pytest
import torch

from helpers import DecodeCfg, IdentityTransforms, TransInfo, unpack_bits
from test_fill_oracle import polygon_cases

cv2 = pytest.importorskip("cv2")
pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def image():
    from isg_b200 import _lib
    from isg_b200.utils import image
    _lib.lib()
    return image


def cv_mask(poly, size):
    return cv2.fillPoly(np.zeros(size, np.int32), [np.asarray(poly).astype(np.int32)], 1)


@pytest.mark.parametrize("seed", [0, 3])
def test_fill_matches_opencv_fuzz(image, seed):
    # all polygons of one frame size in one launch
    rng = np.random.RandomState(seed)
    for size in [(37, 91), (160, 240), (64, 1000)]:
        polys = []
        for (h, w), pts in polygon_cases(seed + size[0], 120, max_hw=size):
            sx, sy = rng.randint(0, size[1] - w + 1), rng.randint(0, size[0] - h + 1)
            polys.append(pts + np.array([sx, sy], np.float32))
        got = image.fill_polygons(polys, size).masks()
        for k, p in enumerate(polys):
            assert np.array_equal(got[k], cv_mask(p, size)), (size, k, p.tolist())


def test_fill_full_frame_layout_feeds_mask_statistics(image):
    size = (96, 200)                                   # W not a multiple of 32: padding bits must stay 0
    polys = [pts for _, pts in polygon_cases(11, 40, max_hw=size)]
    filled = image.fill_polygons(polys, size, full_frame=True)
    ref = np.stack([cv_mask(p, size) for p in polys])
    want_bits = image.pack_masks(ref)
    assert torch.equal(filled.bits, want_bits)
    assert all(np.array_equal(m, r) for m, r in zip(filled.masks(), ref))
    pairs = [(0, 1), (2, 3), (5, 5)]
    assert np.array_equal(image.mask_pair_counts(filled.bits, pairs), image.mask_pair_counts(want_bits, pairs))


def test_fill_tall_and_wide_polygons_take_several_passes(image):
    # bounding boxes larger than the shared-memory bit planes: several row chunks per polygon
    size = (1024, 2048)
    rng = np.random.RandomState(2)
    polys = []
    for _ in range(6):
        k = rng.randint(3, 40)
        polys.append(np.stack([rng.randint(0, 2048, k), rng.randint(0, 1024, k)], 1).astype(np.float32))
    k = 1500                                            # a large jagged outline, angle-sorted
    ang = np.sort(rng.uniform(0, 2 * np.pi, k))
    polys.append(np.stack([1024 + 900 * np.cos(ang) + rng.randint(-3, 4, k), 512 + 480 * np.sin(ang) + rng.randint(-3, 4, k)],
                          1).astype(np.float32))
    for ff in (False, True):
        got = image.fill_polygons(polys, size, full_frame=ff).masks()
        for g, p in zip(got, polys):
            assert np.array_equal(g, cv_mask(p, size))


def test_fill_degenerate_polygons(image):
    size = (20, 40)
    polys = [np.array([[5, 7]], np.float32), np.array([[3, 3], [30, 15]], np.float32),
             np.array([[2, 9], [35, 9], [17, 9]], np.float32), np.array([[8, 2], [8, 18], [8, 11]], np.float32),
             np.array([[4, 4], [4, 4], [4, 4], [9, 9]], np.float32), np.array([[0, 0], [39, 0], [39, 19], [0, 19]], np.float32),
             np.array([[1.9, 2.7], [30.2, 3.99], [12.5, 17.1]], np.float32)]       # truncation like astype(int32)
    got = image.polys_to_masks(polys, size)
    for g, p in zip(got, polys):
        assert g.dtype == np.int32 and np.array_equal(g, cv_mask(p, size)), p.tolist()
    # img_size=None: cropped to the polygon's own extent like poly_to_mask
    for g, p in zip(image.polys_to_masks(polys), polys):
        assert np.array_equal(g, image.poly_to_mask(p))
    assert image.polys_to_masks([], size) == []


def test_fill_vertex_outside_frame_raises(image):
    with pytest.raises(ValueError):
        image.fill_polygons([np.array([[0, 0], [9, 3], [3, 12]], np.float32)], (8, 8))
    with pytest.raises(ValueError):
        image.fill_polygons([np.array([[-1, 0], [5, 3], [3, 6]], np.float32)], (8, 8))


def _decode_scene(B=2, H=256, W=512):
    from isg_b200 import synth
    from isg_b200.utils import decode as dec
    anchors = synth.make_anchors(H, W)
    scenes = [synth.make_scene(410 + b, H, W, [14, 9, 5][b % 3], 8, anchors) for b in range(B)]
    kp = torch.from_numpy(np.stack([s[0].kp for s in scenes])); ae = torch.from_numpy(np.stack([s[0].ae for s in scenes]))
    reg = torch.from_numpy(np.stack([s[1] for s in scenes])); cls = torch.from_numpy(np.stack([s[2] for s in scenes]))
    infos = [TransInfo("/data/img_%d_leftImg8bit.png" % b, (H, W)) for b in range(B)]
    dets = dec.decode_output(torch.zeros((B, 3, H, W)), ((kp.to(DEV), ae.to(DEV), None), reg.to(DEV), cls.to(DEV),
                             torch.from_numpy(anchors).to(DEV)), infos, IdentityTransforms(), DecodeCfg(kp_th=3000), torch.device(DEV))
    return dets, infos, (H, W)


def test_fill_decode_polygons_equal_poly_to_mask(image):
    """The writer's use: masks of the polygons decode_output returns."""
    dets, _, size = _decode_scene()
    polys = [d[3] for img in dets for d in img]
    assert len(polys) > 10
    for g, p in zip(image.polys_to_masks(polys, size), polys):
        assert np.array_equal(g, image.poly_to_mask(np.array(p), size))


def test_results_writer_matches_reference_files(image, tmp_path):
    """write_results (utils/eval_util.py:100-125 layout): same pred.txt lines and the same PNG bytes as the
    reference's per-detection poly_to_mask + cv2.imwrite loop."""
    import os
    from isg_b200.utils import eval_util
    dets, infos, size = _decode_scene()
    names, ids = ["c%d" % j for j in range(8)], [24 + j for j in range(8)]
    eval_util.write_results(dets, [(i.img_path, i.img_size) for i in infos], str(tmp_path), names, ids)
    ref_dir = tmp_path / "ref"
    ref_dir.mkdir()
    for i, det in enumerate(dets):
        base = os.path.splitext(os.path.basename(infos[i].img_path))[0]
        lines = open(os.path.join(str(tmp_path), base + "pred.txt")).read().splitlines()
        want = []
        for j in range(8):
            for k, (c, conf, _, poly) in enumerate(det):
                if c != j:
                    continue
                png = os.path.join("results", base + "_" + names[j] + "_{}.png".format(k))
                want.append("{} {} {}".format(png, ids[j], float(conf)))
                ref_png = str(ref_dir / "m.png")
                cv2.imwrite(ref_png, image.poly_to_mask(np.array(poly), img_size=size) * 255)      # reference :116-124
                assert open(os.path.join(str(tmp_path), png), "rb").read() == open(ref_png, "rb").read()
        assert lines == want and len(want) == len(det)
    # the json pair survives a round trip and feeds the writer unchanged
    eval_util.save_dets(dets, [(i.img_path, i.img_size) for i in infos], str(tmp_path), 7)
    d2, i2 = eval_util.load_dets(str(tmp_path), 7)
    out2 = tmp_path / "again"
    out2.mkdir()
    eval_util.write_results(d2, i2, str(out2), names, ids)
    for f in os.listdir(str(tmp_path / "results")):
        assert open(str(tmp_path / "results" / f), "rb").read() == open(str(out2 / "results" / f), "rb").read()


def test_fill_instances_from_device_polygons(image):
    """Masks straight from the device polygon buffers of a DecodePlan (no host round trip of the polygons)."""
    from isg_b200 import engine, synth
    from isg_b200 import _lib
    B, H, W, C = 2, 256, 512, 8
    dev = torch.device(DEV)
    anchors = synth.make_anchors(H, W)
    scenes = [synth.make_scene(520 + b, H, W, [11, 7][b], C, anchors) for b in range(B)]
    kp = torch.from_numpy(np.stack([s[0].kp for s in scenes])).to(dev); ae = torch.from_numpy(np.stack([s[0].ae for s in scenes])).to(dev)
    reg = torch.from_numpy(np.stack([s[1] for s in scenes])).to(dev); cls = torch.from_numpy(np.stack([s[2] for s in scenes])).to(dev)
    bplan = engine.BoxPlan(B, anchors.reshape(-1, 4).shape[0], C, H, W, dev, cap=1024, max_keep=64)
    dplan = engine.DecodePlan(B, H, W, bplan.N, 3000, dev, "dense", want_score=False, wh_delta=0.1)
    engine.DecodePipeline(bplan, dplan).run(kp, ae, torch.from_numpy(anchors).to(dev), reg, cls, 0.3, 0.2, tail="polygons", obj_pixel_th=2)
    for full in (False, True):
        filled = image.fill_instances(dplan, full_frame=full)
        masks = filled.masks()
        pts = dplan.poly_points.cpu().numpy(); st = dplan.inst_start.cpu().numpy(); ct = dplan.inst_count.cpu().numpy()
        fl = dplan.inst_flags.cpu().numpy()
        assert (fl == 1).sum() >= 10
        for b in range(B):
            for i in range(bplan.N):
                m = masks[b * bplan.N + i]
                if fl[b, i] == 1:
                    assert np.array_equal(m, cv_mask(pts[b, st[b, i]: st[b, i] + ct[b, i]], (H, W)))
                else:
                    assert filled.desc[b * bplan.N + i, 0] == image.FILL_EMPTY and not m.any()


def test_fill_device_polygons_retry_on_overflow(image, monkeypatch):
    """Device-resident polygons: the first output buffer is a guess; an overflow is answered with the exact size."""
    rng = np.random.RandomState(4)
    size = (1024, 2048)
    polys = [np.stack([rng.randint(0, 2048, 5), rng.randint(0, 1024, 5)], 1).astype(np.float32) for _ in range(400)]
    flat = torch.from_numpy(np.concatenate(polys)).to(DEV)
    start = torch.arange(0, 5 * len(polys), 5, dtype=torch.int32)
    count = torch.full((len(polys),), 5, dtype=torch.int32)
    filled = image.fill_polygons((flat, start, count), size)          # 400 large boxes: more than the 32 MB first guess
    assert filled.used > (1 << 23) and (filled.desc[:, 0] == image.FILL_OK).all()
    for i in (0, 7, 399):
        assert np.array_equal(filled.mask(i), cv_mask(polys[i], size))
