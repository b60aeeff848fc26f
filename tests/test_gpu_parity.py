"""GPU parity: libisg.so (through the drop-in modules, i.e. through the C ABI) against the oracle and the
reference-run golden fixtures.  Integer / index results must be bit-exact; floating-point results within the
tolerance north_star states (1e-5 relative, 1e-7 absolute for the far tail)."""
import numpy as np
import pytest
import torch

from helpers import DecodeCfg, IdentityTransforms, TransInfo, unpack_bits

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-5, 1e-7


@pytest.fixture(scope="module")
def mods():
    import isg_b200
    from isg_b200 import _lib, engine, synth
    from isg_b200.utils import decode, image, kmeans, nms, utils
    _lib.lib()
    return dict(lib=_lib, engine=engine, synth=synth, decode=decode, image=image, kmeans=kmeans, nms=nms, utils=utils)


@pytest.fixture(scope="module")
def oracle():
    from oracle import ref_decode, ref_kmeans_nms
    return ref_decode, ref_kmeans_nms


DEV = "cuda:0"


# ---------------------------------------------------------------------------------------------- K2
def test_select_points_golden(mods, golden):
    g = golden("select_points")
    for name in "abcd":
        out = mods["decode"].select_points(torch.from_numpy(g["in_" + name]).to(DEV), int(g["k_" + name]))
        assert out.dtype == torch.uint8 and out.is_cuda
        assert np.array_equal(out.cpu().numpy(), g["out_" + name]), name


@pytest.mark.parametrize("shape,k", [((64, 128), 0), ((64, 128), 64 * 128), ((5, 7), 3), ((1, 1), 1), ((3, 300), 17),
                                     ((130, 36), 999), ((257, 515), 20000)])
def test_select_points_shapes_vs_oracle(mods, oracle, shape, k):
    rd, _ = oracle
    rs = np.random.RandomState(shape[0] * 1000 + shape[1])
    m = mods["synth"]._distinct_float32(rs.normal(0.0, 2.0, size=shape).astype(np.float32))
    want = rd.select_points(torch.from_numpy(m), k).numpy()
    got = mods["decode"].select_points(torch.from_numpy(m).to(DEV), k).cpu().numpy()
    assert np.array_equal(got, want)


def test_select_points_k_too_large_raises(mods):
    with pytest.raises(RuntimeError):
        mods["decode"].select_points(torch.zeros(4, 4, device=DEV), 17)


def test_select_points_ties_select_all_tied(mods):
    # documented tie rule: every pixel tied with the k-th value is selected
    m = torch.zeros(8, 8, device=DEV)
    m[2, 2] = 5.0
    m[6, 6] = 5.0
    out = mods["decode"].select_points(m, 1).cpu().numpy()
    assert out[2, 2] == 1 and out[6, 6] == 1 and out.sum() == 2


def test_topk_count_property_full_size(mods):
    """size-independent property at BASELINE size: exactly k pixels are >= the selected threshold."""
    lib, eng = mods["lib"], mods["engine"]
    H, W, k = 1024, 2048, 20000
    g = torch.Generator(device="cpu").manual_seed(5)
    kp = torch.randn((2, H, W), generator=g).to(DEV)
    kp = kp + torch.arange(H * W, device=DEV).view(1, H, W) * 1e-9   # still may tie; count property allows >= k
    ws_bytes = int(lib.lib().isg_topk_workspace_bytes(2, H, W, k))
    ws, ws_ptr = eng.aligned_workspace(ws_bytes, torch.device(DEV))
    thr = torch.empty(2, dtype=torch.int32, device=DEV)
    lib.call("isg_topk_threshold", kp.data_ptr(), 2, H, W, H * W, k, thr.data_ptr(), ws_ptr, ws_bytes,
             eng.stream_ptr(torch.device(DEV)))
    for b in range(2):
        kth = torch.topk(kp[b].reshape(-1), k).values[-1]
        u = np.array([thr[b].item()], dtype=np.int32).view(np.uint32)[0]
        f = np.array([u & 0x7FFFFFFF if u & 0x80000000 else ~u], dtype=np.uint32).view(np.float32)[0]
        assert f == kth.item()


@pytest.mark.parametrize("shape,k,kind", [((1024, 2048), 1, "normal"), ((1024, 2048), 700000, "normal"), ((512, 1024), 20000, "flat"),
                                          ((512, 1024), 5000, "sorted"), ((300, 500), 149999, "normal"), ((256, 256), 65536, "flat")])
@pytest.mark.parametrize("path", ["sample", "radix", "sample+cluster-select"])
def test_topk_threshold_paths(mods, shape, k, kind, path, monkeypatch):
    """every host-selected path (sample/filter/select, legacy multi-CTA, single-CTA) and the in-kernel
    fallback (flat / sorted images defeat the sample bound) return the exact k-th largest value"""
    lib, eng = mods["lib"], mods["engine"]
    # sample/filter/select or two-level radix; libisg reads its tuning variables once, the debug hook re-reads them
    monkeypatch.setenv("ISG_TOPK_PATH", path.split("+")[0])
    if path.endswith("cluster-select"):       # the 8-CTA cluster form of the select step (default: one CTA per image)
        monkeypatch.setenv("ISG_TOPK_SELECT", "cluster")
    lib.lib().isg_debug_reload_tuning()
    H, W = shape
    g = torch.Generator(device="cpu").manual_seed(H + k)
    if kind == "normal":
        kp = torch.randn((1, H, W), generator=g)
    elif kind == "flat":      # two values only: the k-th largest is heavily tied
        kp = (torch.rand((1, H, W), generator=g) < 0.01).float() * 3.0 - 1.0
    else:                     # monotone ramp: the strided sample sees a biased subset
        kp = torch.arange(H * W, dtype=torch.float32).view(1, H, W) * 1e-3
    kp = kp.to(DEV)
    ws_bytes = int(lib.lib().isg_topk_workspace_bytes(1, H, W, k))
    ws, ws_ptr = eng.aligned_workspace(ws_bytes, torch.device(DEV))
    thr = torch.empty(1, dtype=torch.int32, device=DEV)
    lib.call("isg_topk_threshold", kp.data_ptr(), 1, H, W, H * W, k, thr.data_ptr(), ws_ptr, ws_bytes, eng.stream_ptr(torch.device(DEV)))
    kth = torch.topk(kp[0].reshape(-1), k).values[-1].item()
    u = np.array([thr[0].item()], dtype=np.int32).view(np.uint32)[0]
    f = np.array([u & 0x7FFFFFFF if u & 0x80000000 else ~u], dtype=np.uint32).view(np.float32)[0]
    monkeypatch.delenv("ISG_TOPK_PATH")
    monkeypatch.delenv("ISG_TOPK_SELECT", raising=False)
    lib.lib().isg_debug_reload_tuning()
    assert f == np.float32(kth)


def test_nms_hm_golden(mods, golden):
    g = golden("select_points")
    heat = torch.from_numpy(g["heat"]).to(DEV)
    assert np.array_equal(mods["decode"].nms_hm(heat, 3).cpu().numpy(), g["heat_keep3"])
    assert np.array_equal(mods["decode"].nms_hm(heat, 5).cpu().numpy(), g["heat_keep5"])


# ---------------------------------------------------------------------------------------------- K1/K3
def _run_plan(mods, img, kp_th, mode, want_score=True, fused_stats=False):
    eng = mods["engine"]
    H, W = img.kp.shape[-2:]
    N = len(img.rois)
    plan = eng.DecodePlan(1, H, W, N, kp_th, DEV, mode, want_score=want_score, fused_stats=fused_stats)
    kp = torch.from_numpy(img.kp)[None].to(DEV)
    ae = torch.from_numpy(img.ae)[None].to(DEV)
    rois = torch.from_numpy(img.rois)[None].to(DEV).contiguous()
    n = torch.tensor([N], dtype=torch.int32, device=DEV)
    plan.run(kp, ae, rois, n)
    torch.cuda.synchronize()
    return plan


@pytest.mark.parametrize("mode", ["sparse", "dense", "dense-fused-stats"])
@pytest.mark.parametrize("shape,N,kp_th", [((256, 512), 20, 20000), ((96, 160), 5, 100), ((130, 257), 6, 3000),
                                          ((33, 64), 2, 500), ((70, 260), 4, 2000), ((48, 36), 2, 300)])
def test_group_core_vs_oracle(mods, oracle, mode, shape, N, kp_th):
    rd, _ = oracle
    img = mods["synth"].make_image(77 + N, shape[0], shape[1], N)
    core = rd.group_core(torch.from_numpy(img.kp[0]), torch.from_numpy(img.ae), img.rois, kp_th)
    # dense mode: per-instance statistics from the gather pass (default) or accumulated inside the fused kernel
    plan = _run_plan(mods, img, kp_th, mode.split("-")[0], fused_stats=mode.endswith("fused-stats"))
    M = int(plan.count[0].item())
    assert M == core["idx"].shape[0]
    assert np.array_equal(unpack_bits(plan.keepbits[0].cpu().numpy(), shape[1]), core["mask"].numpy())
    assert np.array_equal(plan.idx[0, :M].cpu().numpy(), core["idx"].numpy().astype(np.int32))
    assert np.array_equal(plan.label[0, :M].cpu().numpy(), core["label"].numpy().astype(np.int32))     # bit-exact
    np.testing.assert_allclose(plan.score[0, :M].cpu().numpy(), core["score"].numpy(), rtol=RTOL, atol=ATOL)
    # per-instance point sets (ghost filter + grouping) — exact
    want = rd.instance_points(core["idx"], core["label"], core["centres"], core["whs"], 0.1)
    off = plan.offsets[0].cpu().numpy()
    pts = plan.points[0].cpu().numpy()
    stats = plan.stats[0].cpu().numpy()
    for i, (p, _) in enumerate(want):
        got = pts[off[i]:off[i + 1]]
        assert np.array_equal(got, p), i
        assert stats[i, 0] == p.shape[0]
        if p.shape[0]:
            assert (stats[i, 1], stats[i, 2], stats[i, 3], stats[i, 4]) == (p[:, 1].min(), p[:, 0].min(), p[:, 1].max(), p[:, 0].max())


@pytest.mark.parametrize("shape,N", [((128, 256), 6), ((130, 257), 5), ((256, 512), 40)])
def test_dense_label_map_vs_oracle(mods, oracle, shape, N):
    rd, _ = oracle
    img = mods["synth"].make_image(5 + N, shape[0], shape[1], N)
    score, label = rd.dense_labels(torch.from_numpy(img.ae), img.rois)
    plan = _run_plan(mods, img, 2000, "dense")
    got_l = plan.label_map[0].cpu().numpy()
    got_s = plan.score_map[0].cpu().numpy()
    np.testing.assert_allclose(got_s, score.numpy(), rtol=RTOL, atol=ATOL)
    # labels are bit-exact wherever the oracle's best/second-best margin is not a float-rounding tie
    P = _all_memberships(rd, img)
    top2 = np.sort(P, axis=2)[:, :, -2:] if P.shape[2] > 1 else np.concatenate([np.zeros_like(P), P], axis=2)
    safe = (top2[:, :, 1] - top2[:, :, 0]) > 1e-5 * np.maximum(top2[:, :, 1], 1e-30)
    safe |= top2[:, :, 1] == 0
    assert safe.mean() > 0.99
    assert np.array_equal(got_l[safe], label.numpy()[safe].astype(np.int32))
    # the dense map restricted to the keep pixels equals the sparse labels
    M = int(plan.count[0].item())
    idx = plan.idx[0, :M].cpu().numpy()
    assert np.array_equal(got_l[idx[:, 0], idx[:, 1]], plan.label[0, :M].cpu().numpy())


def _all_memberships(rd, img):
    """fp32 P[H,W,N] per the oracle's formulas (for margin masks only)."""
    ae = torch.from_numpy(img.ae)
    _, h, w = ae.shape
    ys = torch.linspace(0, 1, 1024)[:h]; xs = torch.linspace(0, 2, 2048)[:w]
    centres, whs = rd.box_geometry(img.rois)
    ci = torch.from_numpy(centres).long()
    C = torch.stack((ys[ci[:, 0]], xs[ci[:, 1]]), dim=1)
    e0 = torch.tanh(ae[0]) + ys[:, None]; e1 = torch.tanh(ae[1]) + xs[None, :]
    s0, s1 = torch.exp(ae[2]), torch.exp(ae[3])
    c_t, wh_t = torch.from_numpy(centres), torch.from_numpy(whs)
    lt, rb = c_t - wh_t / 2, c_t + wh_t / 2
    yy = torch.arange(h).float()[:, None, None]; xx = torch.arange(w).float()[None, :, None]
    inb = (yy - lt[:, 0] >= 0) & (xx - lt[:, 1] >= 0) & (rb[:, 0] - yy >= 0) & (rb[:, 1] - xx >= 0)
    P = torch.exp(-((e0[..., None] - C[:, 0]) ** 2 * s0[..., None] + (e1[..., None] - C[:, 1]) ** 2 * s1[..., None])) * inb
    return P.numpy()


@pytest.mark.parametrize("name", ["s0", "s1", "s2", "s3", "s4"])
@pytest.mark.parametrize("mode", ["sparse", "dense", "dense-host-polygons"])
def test_decode_single_golden(mods, golden, name, mode):
    """the drop-in decode_single against polygons produced by the reference itself (s4: one polygon rejected by the
    centre-inside test, one instance below obj_pixel_th, label-0 strays removed by the ghost filter)"""
    g = golden("decode_single_" + name)
    dec = mods["decode"]
    saved = dec.decode_mode, dec.device_polygon_stage
    dec.decode_mode, dec.device_polygon_stage = mode.split("-")[0], not mode.endswith("host-polygons")
    try:
        h, w = g["kp"].shape[-2:]
        boxes = {"rois": g["rois"], "class_ids": g["class_ids"], "scores": g["scores"]}
        (dets,) = dec.decode_single(torch.from_numpy(g["kp"]).to(DEV), torch.from_numpy(g["ae"]).to(DEV), boxes,
                                    TransInfo("/nonexistent.png", (h, w)), IdentityTransforms(),
                                    DecodeCfg(kp_th=int(g["kp_th"])), torch.device(DEV))
    finally:
        dec.decode_mode, dec.device_polygon_stage = saved
    assert len(dets) == int(g["n_dets"])
    for i, (cls, conf, ctr, poly) in enumerate(dets):
        assert int(cls) == int(g["det_cls_%d" % i])
        assert np.float32(conf) == g["det_conf_%d" % i]
        assert np.array_equal(ctr, g["det_ctr_%d" % i])
        if mode == "dense":      # device polygon stage: equal-angle points may come in a different order
            assert_polygon_equivalent(dec, poly, g["det_poly_%d" % i], ctr)
        else:
            assert np.array_equal(poly, g["det_poly_%d" % i])


@pytest.mark.parametrize("mode", ["sparse", "dense"])
@pytest.mark.parametrize("draw", [False, True])
def test_decode_single_resize_transform_golden(mods, golden, mode, draw, tmp_path):
    """non-identity validation transform (resize, utils/tranform.py:157-171) with decode.target_size = 2 (test.py:58):
    the ghost filter and the polygon stage run on the host in original-image pixels; draw=True is the stock
    configs/decode_cfg.yaml setting (drawing itself is skipped when the reference's utils.visualize is absent)"""
    import cv2
    import warnings
    from helpers import ResizeTransforms
    g = golden("decode_single_resize")
    dec = mods["decode"]
    ts, size = int(g["target_size"]), tuple(int(v) for v in g["img_size"])
    img_path = str(tmp_path / "frame.png")
    cv2.imwrite(img_path, np.zeros(size + (3,), np.uint8))
    saved = dec.decode_mode, dec.target_size, dec.base_dir
    dec.decode_mode, dec.target_size, dec.base_dir = mode, ts, str(tmp_path)
    try:
        boxes = {"rois": g["rois"], "class_ids": g["class_ids"], "scores": g["scores"]}
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            (dets,) = dec.decode_single(torch.from_numpy(g["kp"]).to(DEV), torch.from_numpy(g["ae"]).to(DEV), boxes,
                                        TransInfo(img_path, size), ResizeTransforms(ts),
                                        DecodeCfg(kp_th=int(g["kp_th"]), draw_flag=draw), torch.device(DEV))
    finally:
        dec.decode_mode, dec.target_size, dec.base_dir = saved
    assert len(dets) == int(g["n_dets"]) > 0
    for i, (cls, conf, ctr, poly) in enumerate(dets):
        assert int(cls) == int(g["det_cls_%d" % i]) and np.float32(conf) == g["det_conf_%d" % i]
        assert np.array_equal(ctr, g["det_ctr_%d" % i])
        assert np.array_equal(poly, g["det_poly_%d" % i])
    assert (tmp_path / "frame.png_candid.png").exists() == draw          # the debug image of :370-372


def test_decode_ct_hm_golden(mods, golden):
    """decode_ct_hm (utils/decode.py:254-285) against the reference's own output"""
    g = golden("decode_ct_hm")
    h, w = g["conf"].shape
    mods["decode"].device = torch.device(DEV)          # what test.py:134 / evaluate.py:40 set
    for where in ("cpu", DEV):           # the reference's callers hand CPU tensors; device tensors work too
        out = mods["decode"].decode_ct_hm(torch.from_numpy(g["conf"]).to(where), torch.from_numpy(g["cls"]).to(where),
                                          torch.from_numpy(g["wh"]).to(where), int(g["num_classes"]), int(g["k"]),
                                          IdentityTransforms(), TransInfo("/nonexistent.png", (h, w)))
        assert np.array_equal(np.asarray(out[0], dtype=np.int64), g["keep_cls"])
        assert np.array_equal(np.asarray(out[1], dtype=np.int64).reshape(-1, 2), g["keep_idx"])
        assert np.array_equal(np.asarray(out[2], dtype=np.float32), g["keep_conf"])
        assert np.array_equal(np.asarray(out[3], dtype=np.float32).reshape(-1, 2), g["keep_wh"])


def test_decode_single_no_boxes(mods):
    dec = mods["decode"]
    boxes = {"rois": np.array(()), "class_ids": np.array(()), "scores": np.array(())}
    out = dec.decode_single(torch.zeros(1, 8, 8, device=DEV), torch.zeros(4, 8, 8, device=DEV), boxes,
                            TransInfo("x", (8, 8)), IdentityTransforms(), DecodeCfg(), torch.device(DEV))
    assert out == ([],)


def test_cpu_device_is_refused(mods):
    with pytest.raises(RuntimeError):
        mods["decode"].group_kp(torch.zeros(8, 8), torch.zeros(4, 8, 8), IdentityTransforms(), [np.zeros(2, np.float32)],
                                [np.zeros(2, np.float32)], [0], [0.5], TransInfo("x", (8, 8)), DecodeCfg(), torch.device("cpu"))


# ---------------------------------------------------------------------------------------------- a2 / K5
def test_decode_boxes_golden(mods, golden):
    g = golden("decode_boxes")
    H, W = int(g["H"]), int(g["W"])
    x = torch.zeros((3, 3, H, W))
    dets = mods["decode"].decode_boxes(x, torch.from_numpy(g["anchors"]).to(DEV), torch.from_numpy(g["regression"]).to(DEV),
                                       torch.from_numpy(g["classification"]).to(DEV), 0.3, 0.2)
    for b, det in enumerate(dets):
        assert np.array_equal(det["class_ids"], g["cls_%d" % b])
        assert np.array_equal(det["scores"], g["scores_%d" % b])
        np.testing.assert_allclose(det["rois"], g["rois_%d" % b].reshape(det["rois"].shape), rtol=1e-6, atol=1e-4)
    assert dets[2]["rois"].shape == (0,)


def test_anchors_golden(mods, golden):
    """device anchor generation (isg_generate_anchors behind the Anchors module) is bit-identical to the reference's
    Anchors.forward (utils/utils.py:366-450), incl. the fp16 branch, custom levels/scales/ratios and a ragged height"""
    import hashlib
    g = golden("anchors")
    Anchors = mods["utils"].Anchors
    img = lambda b, h, w: torch.empty((b, 3, h, w), device=DEV)
    a = Anchors()(img(1, 128, 256))
    assert a.dtype == torch.float32 and a.device.type == "cuda" and np.array_equal(a.cpu().numpy(), g["a_128x256"])
    assert np.array_equal(Anchors()(img(1, 128, 256), dtype=torch.float16).cpu().numpy(), g["a_half_128x256"])
    custom = Anchors(anchor_scale=3., pyramid_levels=[2, 3, 4], scales=[1.0, 1.5], ratios=[(1.0, 1.0), (2.0, 0.5)])
    assert np.array_equal(custom(img(2, 96, 160)).cpu().numpy(), g["a_custom_96x160"])
    assert np.array_equal(Anchors(pyramid_levels=[3, 4])(img(1, 100, 64)).cpu().numpy(), g["a_ragged_100x64"])
    for h, w in ((1024, 2048), (512, 1024)):
        mod = Anchors()
        t = mod(img(1, h, w))
        assert mod(img(1, h, w)) is t                                  # cached per (shape, device), :401-402
        arr = t.cpu().numpy()
        assert arr.shape == (1, int(g["count_%dx%d" % (h, w)]), 4)
        assert np.array_equal(np.frombuffer(hashlib.sha256(arr.tobytes()).digest(), dtype=np.uint8), g["sha_%dx%d" % (h, w)])
        assert np.array_equal(arr[0, g["rows_%dx%d" % (h, w)]], g["vals_%dx%d" % (h, w)])
    with pytest.raises(ValueError):
        Anchors()(img(1, 128, 200))                                    # W not divisible by the stride, :416-417
    with pytest.raises(RuntimeError):
        Anchors()(torch.empty((1, 3, 128, 256)))                       # no CPU path


def test_inference_heads_golden(mods, golden):
    """f4: the kp / ae heads of the reference's EfficientDecoder (1x1 convolutions, models/efficient.py:508-510,536-541) in
    one pass, `tan` dropped; against the reference module's own outputs (fp32, summation order differs: rtol 1e-5)"""
    from isg_b200.utils.heads import InferenceHeads
    g = golden("heads")
    kp_conv = torch.nn.Conv2d(16, 1, 1); ae_conv = torch.nn.Conv2d(16, 4, 1)
    with torch.no_grad():
        kp_conv.weight.copy_(torch.from_numpy(g["w_kp"])); kp_conv.bias.copy_(torch.from_numpy(g["b_kp"]))
        ae_conv.weight.copy_(torch.from_numpy(g["w_ae"])); ae_conv.bias.copy_(torch.from_numpy(g["b_ae"]))
    heads = InferenceHeads(kp_conv, ae_conv)
    kp, ae, tan = heads(torch.from_numpy(g["x"]).to(DEV))
    assert tan is None and kp.shape == (2, 1, 24, 40) and ae.shape == (2, 4, 24, 40)
    np.testing.assert_allclose(kp.cpu().numpy(), g["kp"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(ae.cpu().numpy(), g["ae"], rtol=1e-5, atol=1e-6)
    # the outputs feed the decode directly: full-size planes, same layout as the model's
    x = torch.randn((1, 16, 256, 512), device=DEV)
    kp, ae, _ = heads(x)
    want = torch.nn.functional.conv2d(x.cpu(), torch.from_numpy(g["w_ae"]), torch.from_numpy(g["b_ae"]))
    np.testing.assert_allclose(ae.cpu().numpy(), want.numpy(), rtol=1e-5, atol=1e-5)
    with pytest.raises(RuntimeError):
        heads(torch.zeros(1, 16, 8, 8))


def test_bbox_transform_and_clip_vs_oracle(mods, oracle, golden):
    rd, _ = oracle
    g = golden("decode_boxes")
    anchors, reg = torch.from_numpy(g["anchors"]), torch.from_numpy(g["regression"])
    want = rd.bbox_transform(anchors, reg)
    got = mods["utils"].BBoxTransform()(anchors.to(DEV), reg.to(DEV))
    np.testing.assert_allclose(got.cpu().numpy(), want.numpy(), rtol=1e-6, atol=1e-4)
    img = torch.zeros((3, 3, int(g["H"]), int(g["W"])))
    clipped = mods["utils"].ClipBoxes()(got, img)
    assert clipped.data_ptr() == got.data_ptr()
    np.testing.assert_allclose(clipped.cpu().numpy(), rd.clip_boxes(want, int(g["H"]), int(g["W"])).numpy(), rtol=1e-6, atol=1e-4)


def test_py_cpu_nms_golden(mods, golden):
    g = golden("nms")
    for name in "abc":
        keep = mods["nms"].py_cpu_nms(g["dets_" + name], float(g["thr_" + name]))
        assert isinstance(keep, list)
        assert np.array_equal(np.asarray(keep, dtype=np.int64), g["keep_" + name]), name
    assert mods["nms"].py_cpu_nms(np.zeros((0, 5), np.float32), 0.5) == []


@pytest.mark.parametrize("n,thr", [(1000, 0.5), (2500, 0.3), (65, 0.7)])
def test_py_cpu_nms_vs_oracle(mods, oracle, n, thr):
    _, rk = oracle
    dets = mods["synth"].make_nms_boxes(n, n, extent=1500.0, thr=thr, plus1=True)
    assert np.array_equal(np.asarray(mods["nms"].py_cpu_nms(dets, thr)), np.asarray(rk.py_cpu_nms(dets, thr)))


@pytest.mark.parametrize("n", [300, 1200])
def test_nms_suppression_chain_falls_back_to_sequential_scan(mods, oracle, n):
    """every box suppresses only its successor: the parallel suppression scan (one round per level of the chain) gives up
    after its round limit and the sequential scan finishes - small fused kernel (n <= 1024) and staged scan"""
    _, rk = oracle
    x = 30.0 * np.arange(n, dtype=np.float32)
    dets = np.stack([x, np.zeros_like(x), x + 100.0, np.full_like(x, 50.0), np.linspace(0.99, 0.01, n).astype(np.float32)], axis=1)
    want = rk.py_cpu_nms(dets, 0.5)
    assert list(want) == list(range(0, n, 2))
    assert np.array_equal(np.asarray(mods["nms"].py_cpu_nms(dets, 0.5)), np.asarray(want))


@pytest.mark.parametrize("rounds", ["0", "2"])
@pytest.mark.parametrize("n", [700, 1250])
def test_nms_parallel_and_sequential_scans_agree(mods, oracle, n, rounds, monkeypatch):
    """ISG_NMS_ROUNDS=0 disables the parallel scan, 2 makes it give up on most inputs: same keep list either way"""
    _, rk = oracle
    lib = mods["lib"]
    dets = mods["synth"].make_nms_boxes(n + 1, n, extent=900.0, thr=0.4, plus1=True)
    want = np.asarray(rk.py_cpu_nms(dets, 0.4))
    assert np.array_equal(np.asarray(mods["nms"].py_cpu_nms(dets, 0.4)), want)
    monkeypatch.setenv("ISG_NMS_ROUNDS", rounds)
    lib.lib().isg_debug_reload_tuning()
    try:
        got = np.asarray(mods["nms"].py_cpu_nms(dets, 0.4))
    finally:
        monkeypatch.delenv("ISG_NMS_ROUNDS")
        lib.lib().isg_debug_reload_tuning()
    assert np.array_equal(got, want)


def test_boxes_nms_intended_semantics(mods, oracle):
    _, rk = oracle
    dets = mods["synth"].make_nms_boxes(9, 400, extent=500.0, thr=0.4, plus1=True)
    rs = np.random.RandomState(1)
    d = {"rois": dets[:, :4], "scores": dets[:, 4], "class_ids": rs.randint(0, 5, size=len(dets))}
    c1, b1, s1 = mods["nms"].boxes_nms(d, 0.4)
    c2, b2, s2 = rk.boxes_nms(d, 0.4)
    assert np.array_equal(np.asarray(c1), np.asarray(c2)) and np.array_equal(np.asarray(b1), np.asarray(b2))
    assert np.array_equal(np.asarray(s1), np.asarray(s2))
    assert mods["nms"].boxes_nms({"class_ids": np.array(()), "rois": np.array(()), "scores": np.array(())}, 0.5) == ([], [], [])


def test_tv_nms_vs_torchvision(mods):
    from torchvision.ops.boxes import batched_nms
    lib, eng = mods["lib"], mods["engine"]
    n, thr = 3000, 0.2
    dets = mods["synth"].make_nms_boxes(4, n, extent=2000.0, thr=thr, plus1=False)
    n = len(dets)
    rs = np.random.RandomState(2)
    cls = rs.randint(0, 8, size=n).astype(np.int32)
    want = batched_nms(torch.from_numpy(dets[:, :4].copy()), torch.from_numpy(dets[:, 4].copy()), torch.from_numpy(cls).long(), thr).numpy()
    d = torch.device(DEV)
    boxes = torch.from_numpy(dets[:, :4].copy()).to(d).contiguous(); scores = torch.from_numpy(dets[:, 4].copy()).to(d)
    clsd = torch.from_numpy(cls).to(d); count = torch.tensor([n], dtype=torch.int32, device=d)
    keep = torch.empty(n, dtype=torch.int32, device=d); nk = torch.empty(1, dtype=torch.int32, device=d)
    wsb = int(lib.lib().isg_box_nms_workspace_bytes(1, n)); ws = torch.empty(wsb + 256, dtype=torch.uint8, device=d)
    off = (-ws.data_ptr()) % 256
    lib.call("isg_box_nms", boxes.data_ptr(), scores.data_ptr(), clsd.data_ptr(), 0, count.data_ptr(), 1, n, thr,
             lib.ISG_NMS_TV_GT, keep.data_ptr(), nk.data_ptr(), ws.data_ptr() + off, wsb, eng.stream_ptr(d))
    got = keep[:int(nk.item())].cpu().numpy()
    assert np.array_equal(got, want)


def _box_nms(mods, boxes, scores, cls, thr, convention):
    lib, eng = mods["lib"], mods["engine"]
    d = torch.device(DEV)
    n = len(boxes)
    bt = torch.from_numpy(boxes.copy()).to(d).contiguous(); st = torch.from_numpy(scores.copy()).to(d)
    ct = torch.from_numpy(cls.astype(np.int32)).to(d); count = torch.tensor([n], dtype=torch.int32, device=d)
    keep = torch.empty(n, dtype=torch.int32, device=d); nk = torch.empty(1, dtype=torch.int32, device=d)
    wsb = int(lib.lib().isg_box_nms_workspace_bytes(1, n))
    ws, ws_ptr = eng.aligned_workspace(max(wsb, 256), d)
    lib.call("isg_box_nms", bt.data_ptr(), st.data_ptr(), ct.data_ptr(), 0, count.data_ptr(), 1, n, thr, convention,
             keep.data_ptr(), nk.data_ptr(), ws_ptr, wsb, eng.stream_ptr(d))
    return keep[:int(nk.item())].cpu().numpy()


@pytest.mark.parametrize("n_pairs", [300, 600])        # 600 / 1200 boxes: fused small kernel / staged path
def test_box_nms_conventions_follow_torchvision_on_near_threshold_pairs(mods, n_pairs):
    """IoUs within a few fp32 ulps of the threshold: the coordinate trick (shifted fp32 coordinates) and the per-class
    NMS resolve some pairs differently; each convention of isg_box_nms must reproduce its torchvision function bit for
    bit, and ISG_NMS_TV_BATCHED the dispatch of batched_nms itself (trick up to 1000 boxes, per-class above)"""
    tvb = pytest.importorskip("torchvision.ops.boxes")
    lib = mods["lib"]
    differ = 0
    for seed in (11, 12):
        b, s, c = mods["synth"].make_near_threshold_boxes(seed, n_pairs, 0.5)
        tb, ts, tc = torch.from_numpy(b), torch.from_numpy(s), torch.from_numpy(c)
        trick = tvb._batched_nms_coordinate_trick(tb, ts, tc, 0.5).numpy()
        vanilla = tvb._batched_nms_vanilla(tb, ts, tc, 0.5).numpy()
        assert np.array_equal(_box_nms(mods, b, s, c, 0.5, lib.ISG_NMS_TV_TRICK), trick)
        assert np.array_equal(_box_nms(mods, b, s, c, 0.5, lib.ISG_NMS_TV_GT), vanilla)
        assert np.array_equal(_box_nms(mods, b, s, c, 0.5, lib.ISG_NMS_TV_BATCHED), tvb.batched_nms(tb, ts, tc, 0.5).numpy())
        differ += len(set(trick.tolist()) ^ set(vanilla.tolist()))
    assert differ > 0


def _random_boxes(seed, n, extent, ncls):
    rs = np.random.RandomState(seed)
    ctr = rs.uniform(0, extent, size=(n, 2)); wh = rs.uniform(20, 200, size=(n, 2))
    b = np.concatenate([ctr - wh / 2, ctr + wh / 2], axis=1).astype(np.float32)
    s = rs.permutation(n).astype(np.float32) / np.float32(n) + np.float32(0.001)          # distinct scores
    return b, s, rs.randint(0, ncls, size=n).astype(np.int64)


@pytest.mark.parametrize("rounds", [None, "0"])
def test_box_nms_beyond_the_matrix_limit_follows_torchvision(mods, rounds, monkeypatch):
    """more candidates than ISG_NMS_MAX_BOXES (an untrained head fires on every anchor): the tiled large-set path (radix
    sort + tile-by-tile resolution, no suppression matrix) gives torchvision's keep list; rounds=0 forces the sequential
    scan inside every tile"""
    tvb = pytest.importorskip("torchvision.ops.boxes")
    lib = mods["lib"]
    n = 40000
    assert n > lib.ISG_NMS_MAX_BOXES
    b, s, c = _random_boxes(5, n, 3000.0, 8)
    tb, ts, tc = torch.from_numpy(b), torch.from_numpy(s), torch.from_numpy(c)
    if rounds is not None:
        monkeypatch.setenv("ISG_NMS_ROUNDS", rounds)
        lib.lib().isg_debug_reload_tuning()
    try:
        got_gt = _box_nms(mods, b, s, c, 0.5, lib.ISG_NMS_TV_GT)
        got_batched = _box_nms(mods, b, s, c, 0.5, lib.ISG_NMS_TV_BATCHED)
        got_trick = _box_nms(mods, b, s, c, 0.5, lib.ISG_NMS_TV_TRICK)
    finally:
        if rounds is not None:
            monkeypatch.delenv("ISG_NMS_ROUNDS")
            lib.lib().isg_debug_reload_tuning()
    vanilla = tvb._batched_nms_vanilla(tb, ts, tc, 0.5).numpy()
    assert np.array_equal(got_gt, vanilla)
    assert np.array_equal(got_batched, tvb.batched_nms(tb, ts, tc, 0.5).numpy())
    assert np.array_equal(got_trick, tvb._batched_nms_coordinate_trick(tb, ts, tc, 0.5).numpy())


def test_py_cpu_nms_beyond_the_matrix_limit(mods, oracle):
    _, rk = oracle
    b, s, _ = _random_boxes(6, 20000, 2500.0, 1)
    dets = np.concatenate([b, s[:, None]], axis=1).astype(np.float32)
    assert np.array_equal(np.asarray(mods["nms"].py_cpu_nms(dets, 0.5)), np.asarray(rk.py_cpu_nms(dets, 0.5)))


def test_decode_boxes_every_anchor_fires(mods, oracle):
    """all of ~24.5 k anchors above cls_th (what an untrained classification head produces), two images: the reference
    runs batched_nms on all of them (utils/decode.py:395-400); the drop-in re-plans past ISG_NMS_MAX_BOXES instead of raising"""
    rd, _ = oracle
    H, W, B = 256, 512, 2
    anchors = mods["utils"].Anchors()(torch.zeros((1, 3, H, W), device=DEV)).cpu()
    A, C = anchors.shape[1], 4
    assert A > mods["lib"].ISG_NMS_MAX_BOXES
    g = torch.Generator().manual_seed(3)
    regression = torch.randn((B, A, 4), generator=g) * 0.2
    regression[..., 2:] = 0.0            # exp(0) is exact on both sides: the decoded boxes are bit-identical
    classification = torch.full((B, A, C), 0.26)                 # distinct maxima: the order of equal scores is unspecified
    for b in range(B):
        top = torch.randperm(A, generator=g).float() / A * 0.5 + 0.3
        classification[b, torch.arange(A), torch.randint(0, C, (A,), generator=g)] = top
    want = rd.decode_boxes(H, W, anchors, regression, classification, 0.25, 0.4)
    got = mods["decode"].decode_boxes(torch.zeros((B, 3, H, W)), anchors.to(DEV), regression.to(DEV), classification.to(DEV), 0.25, 0.4)
    for b in range(B):
        assert len(want[b]["scores"]) > 100
        assert np.array_equal(got[b]["class_ids"], want[b]["class_ids"])
        assert np.array_equal(got[b]["scores"], want[b]["scores"])
        np.testing.assert_allclose(got[b]["rois"], want[b]["rois"], rtol=1e-6, atol=1e-4)


def test_decode_boxes_near_threshold_pairs_follow_the_reference(mods, oracle):
    """decode_boxes on anchors that form near-threshold pairs (regression 0, so the decoded boxes are the anchors): the
    kept set must be the one torchvision.ops.batched_nms gives the reference (utils/decode.py:400), bit for bit"""
    rd, _ = oracle
    b, s, c = mods["synth"].make_near_threshold_boxes(21, 350, 0.5, extent=1000.0)
    A, C, H, W = len(b), 8, 1024, 2048
    anchors = torch.from_numpy(np.ascontiguousarray(b[:, [1, 0, 3, 2]]))[None]          # y1,x1,y2,x2
    regression = torch.zeros((1, A, 4))
    classification = torch.zeros((1, A, C))
    classification[0, torch.arange(A), torch.from_numpy(c)] = torch.from_numpy(s)
    want = rd.decode_boxes(H, W, anchors, regression, classification, 0.05, 0.5)[0]
    got = mods["decode"].decode_boxes(torch.zeros((1, 3, H, W)), anchors.to(DEV), regression.to(DEV), classification.to(DEV), 0.05, 0.5)[0]
    assert np.array_equal(got["class_ids"], want["class_ids"])
    assert np.array_equal(got["scores"], want["scores"])
    assert np.array_equal(got["rois"], want["rois"])


# ---------------------------------------------------------------------------------------------- full decode
@pytest.mark.parametrize("mode", ["sparse", "dense", "dense-host-polygons"])
def test_decode_output_vs_oracle(mods, oracle, mode):
    rd, _ = oracle
    synth, dec = mods["synth"], mods["decode"]
    H, W, C, B = 256, 512, 8, 3
    anchors = synth.make_anchors(H, W)
    scenes = [synth.make_scene(300 + b, H, W, [12, 0, 25][b], C, anchors) for b in range(B)]
    kp = torch.from_numpy(np.stack([s[0].kp for s in scenes])); ae = torch.from_numpy(np.stack([s[0].ae for s in scenes]))
    reg = torch.from_numpy(np.stack([s[1] for s in scenes])); cls = torch.from_numpy(np.stack([s[2] for s in scenes]))
    anc = torch.from_numpy(anchors)
    want = rd.decode_output(H, W, ((kp, ae, None), reg, cls, anc), kp_th=3000)
    saved = dec.decode_mode, dec.device_polygon_stage
    dec.decode_mode, dec.device_polygon_stage = mode.split("-")[0], not mode.endswith("host-polygons")
    try:
        inputs = torch.zeros((B, 3, H, W))
        infos = [TransInfo("/nonexistent.png", (H, W))] * B
        got = dec.decode_output(inputs, ((kp.to(DEV), ae.to(DEV), None), reg.to(DEV), cls.to(DEV), anc.to(DEV)), infos,
                                IdentityTransforms(), DecodeCfg(kp_th=3000), torch.device(DEV))
    finally:
        dec.decode_mode, dec.device_polygon_stage = saved
    assert len(got) == B and len(got[1]) == 0
    for b in range(B):
        assert len(got[b]) == len(want[b]) and (b == 1 or len(got[b]) > 0)
        for (c1, f1, ctr1, p1), (c2, f2, ctr2, p2) in zip(got[b], want[b]):
            assert int(c1) == int(c2) and np.float32(f1) == np.float32(f2)
            assert np.array_equal(ctr1, ctr2)
            assert_polygon_equivalent(dec, p1, p2, ctr2)


def test_decode_output_replans_past_the_nms_matrix_limit(mods):
    """every anchor above cls_th: decode_output grows its candidate / seed plans past ISG_NMS_MAX_BOXES (tiled large-set NMS)
    instead of raising, and every detection it returns is one of decode_boxes' boxes"""
    synth, dec = mods["synth"], mods["decode"]
    H, W, C = 256, 512, 4
    anchors = synth.make_anchors(H, W)
    img = synth.make_scene(77, H, W, 10, C, anchors)[0]
    A = anchors.shape[1] if anchors.ndim == 3 else anchors.shape[0]
    assert A > mods["lib"].ISG_NMS_MAX_BOXES
    g = torch.Generator().manual_seed(4)
    reg = torch.randn((1, A, 4), generator=g) * 0.2
    reg[..., 2:] = 0.0
    cls = torch.full((1, A, C), 0.26)
    cls[0, torch.arange(A), torch.randint(0, C, (A,), generator=g)] = torch.randperm(A, generator=g).float() / A * 0.5 + 0.3
    anc = torch.from_numpy(anchors).reshape(1, A, 4)
    kp, ae = torch.from_numpy(img.kp)[None], torch.from_numpy(img.ae)[None]
    boxes = dec.decode_boxes(torch.zeros((1, 3, H, W)), anc.to(DEV), reg.to(DEV), cls.to(DEV), 0.25, 0.1)[0]
    cfg = DecodeCfg(kp_th=3000)
    cfg.cls_th, cfg.iou_th = 0.25, 0.1
    got = dec.decode_output(torch.zeros((1, 3, H, W)), ((kp.to(DEV), ae.to(DEV), None), reg.to(DEV), cls.to(DEV), anc.to(DEV)),
                            [TransInfo("/nonexistent.png", (H, W))], IdentityTransforms(), cfg, torch.device(DEV))
    assert len(got) == 1 and len(boxes["scores"]) > 64
    known = {(int(c), np.float32(f).item()) for c, f in zip(boxes["class_ids"], boxes["scores"])}
    for c, f, ctr, poly in got[0]:
        assert (int(c), np.float32(f).item()) in known and len(poly) >= 1


@pytest.mark.parametrize("mode", ["sparse", "dense", "dense-host-polygons"])
def test_decode_output_plateau_overflow_is_not_silent(mods, mode):
    """A plateau at the k-th value: every tied pixel is selected, so an image keeps more pixels than k.  The drop-in must
    notice (the kernels keep counting past the plan's capacity) and decode again with room - never return fewer
    detections silently.  Checked against the same decode with k = H*W, where nothing can overflow."""
    synth, dec = mods["synth"], mods["decode"]
    H, W, B = 128, 256, 2
    anchors = synth.make_anchors(H, W)
    scenes = [synth.make_scene(700 + b, H, W, 3, 8, anchors) for b in range(B)]
    kp = torch.zeros((B, 1, H, W)); kp[1] = 0.5                      # flat heat maps: all H*W pixels tie and are 3x3 maxima
    ae = torch.from_numpy(np.stack([s[0].ae for s in scenes]))
    reg = torch.from_numpy(np.stack([s[1] for s in scenes])); cls = torch.from_numpy(np.stack([s[2] for s in scenes]))
    outs = ((kp.to(DEV), ae.to(DEV), None), reg.to(DEV), cls.to(DEV), torch.from_numpy(anchors).to(DEV))
    infos = [TransInfo("/nonexistent.png", (H, W))] * B
    saved = dec.decode_mode, dec.device_polygon_stage
    dec.decode_mode, dec.device_polygon_stage = mode.split("-")[0], not mode.endswith("host-polygons")
    try:
        got = dec.decode_output(torch.zeros((B, 3, H, W)), outs, infos, IdentityTransforms(), DecodeCfg(kp_th=100), torch.device(DEV))
        want = dec.decode_output(torch.zeros((B, 3, H, W)), outs, infos, IdentityTransforms(), DecodeCfg(kp_th=H * W), torch.device(DEV))
        boxes = dec.decode_boxes(torch.zeros((B, 3, H, W)), outs[3], outs[1], outs[2], 0.3, 0.2)
        (single,) = dec.decode_single(outs[0][0][0], outs[0][1][0], boxes[0], infos[0], IdentityTransforms(), DecodeCfg(kp_th=100),
                                      torch.device(DEV))
    finally:
        dec.decode_mode, dec.device_polygon_stage = saved
    assert sum(len(g) for g in got) > 0
    for res, ref in ((got[0], want[0]), (got[1], want[1]), (single, want[0])):
        assert len(res) == len(ref)
        for (c1, f1, ctr1, p1), (c2, f2, ctr2, p2) in zip(res, ref):
            assert int(c1) == int(c2) and np.float32(f1) == np.float32(f2) and np.array_equal(ctr1, ctr2)
            assert p1.shape[0] > 100 and np.array_equal(p1, p2)      # more points than k in ONE instance


def assert_polygon_equivalent(dec, got, want, centre_xy):
    """Bit-exact, except that points with EQUAL polar angle about the internal point may come in any order
    (np.argsort's order among equal keys is unspecified; the device sort keeps them row-major)."""
    if np.array_equal(got, want):
        return
    assert got.shape == want.shape and got.dtype == want.dtype
    key = lambda a: a[np.lexsort((a[:, 0], a[:, 1]))]
    assert np.array_equal(key(got), key(want)), "different point sets"
    # the internal point is found on the UNSORTED point set (row-major pixel order, utils/decode.py:342-359): its
    # crossing test depends on the vertex order, so it must not be recomputed from the sorted polygon
    internal = np.asarray(dec.find_internal_point(key(want), np.asarray(centre_xy, dtype=np.float32)), dtype=np.float32)
    th_g = dec._polar_angles(got, np.repeat(internal[None], len(got), 0))
    th_w = dec._polar_angles(want, np.repeat(internal[None], len(want), 0))
    assert np.all(np.diff(th_w[~np.isnan(th_w)]) >= 0), "the expected polygon is not sorted about this internal point"
    assert np.array_equal(th_g, th_w, equal_nan=True), "polygons differ beyond the order of equal-angle points"


def test_decode_output_from_host_tensors(mods):
    """the e2e entry: pinned host tensors in, python lists out"""
    synth, dec = mods["synth"], mods["decode"]
    H, W = 128, 256
    anchors = synth.make_anchors(H, W)
    img, reg, cls, _ = synth.make_scene(9, H, W, 6, 8, anchors)
    outs = ((torch.from_numpy(img.kp)[None].pin_memory(), torch.from_numpy(img.ae)[None].pin_memory(), None),
            torch.from_numpy(reg)[None].pin_memory(), torch.from_numpy(cls)[None].pin_memory(), torch.from_numpy(anchors))
    got = dec.decode_output(torch.zeros((1, 3, H, W)), outs, [TransInfo("x", (H, W))], IdentityTransforms(),
                            DecodeCfg(kp_th=2000), torch.device(DEV))
    assert len(got) == 1 and len(got[0]) >= 1
    assert got[0][0][3].dtype == np.float32 and got[0][0][3].shape[1] == 2


# ---------------------------------------------------------------------------------------------- K4
def test_kmeans_golden(mods, golden):
    g = golden("kmeans")
    km = mods["kmeans"]
    lab, ctr = km.kmeans(torch.from_numpy(g["X"]), 10, torch.from_numpy(g["init"]), g["allow"], device=torch.device(DEV))
    assert lab.dtype == torch.int64
    assert np.array_equal(lab.cpu().numpy(), g["labels"])
    np.testing.assert_allclose(ctr.cpu().numpy(), g["centers"], rtol=RTOL, atol=ATOL)
    lab, ctr = km.kmeans(torch.from_numpy(g["X"] + 1.0), 10, torch.from_numpy(g["init"] + 1.0), np.full(10, 0.002, dtype=np.float32),
                         distance="cosine", device=torch.device(DEV))
    assert np.array_equal(lab.cpu().numpy(), g["labels_cos"])
    np.testing.assert_allclose(ctr.cpu().numpy(), g["centers_cos"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(km.pairwise_distance(torch.from_numpy(g["X"][:40]), torch.from_numpy(g["init"]), torch.device(DEV)).cpu().numpy(),
                               g["pd"], rtol=1e-6, atol=0)
    np.testing.assert_allclose(km.pairwise_cosine(torch.from_numpy(g["X"][:40] + 1.0), torch.from_numpy(g["init"] + 1.0), torch.device(DEV)).cpu().numpy(),
                               g["pc"], rtol=0, atol=2e-7)
    with pytest.raises(NotImplementedError):
        km.kmeans(torch.zeros(4, 2), 2, torch.zeros(2, 2), np.ones(2, np.float32), distance="manhattan")


def test_kmeans_crowd_golden(mods, golden):
    """BASELINE config 4 shape (M ~ 20000 embeddings, N = 500 seeds, allow 0.05): labels of the REFERENCE, bit for bit.
    The generator certified every point's label margin along the whole trajectory (synth.make_kmeans_case)."""
    g = golden("kmeans_crowd")
    km = mods["kmeans"]
    lab, ctr = km.kmeans(torch.from_numpy(g["X"]), 500, torch.from_numpy(g["init"]), g["allow"], device=torch.device(DEV))
    assert np.array_equal(lab.cpu().numpy(), g["labels"].astype(np.int64))
    np.testing.assert_allclose(ctr.cpu().numpy(), g["centers"], rtol=1e-6, atol=1e-7)
    again, ctr2 = km.kmeans(torch.from_numpy(g["X"]), 500, torch.from_numpy(g["init"]), g["allow"], device=torch.device(DEV))
    assert torch.equal(again, lab) and torch.equal(ctr2, ctr)                 # fixed order of additions: reproducible


@pytest.mark.parametrize("M,N,seed", [(6000, 200, 7), (3000, 33, 8), (100, 500, 9)])
def test_kmeans_vs_oracle_exact(mods, oracle, M, N, seed):
    """labels and iteration count bit-exact against the oracle on margin-certified cases (incl. more clusters than points)"""
    _, rk = oracle
    X, init, allow, _, _, _ = mods["synth"].make_kmeans_case(seed, M, N)
    lab_w, ctr_w, it_w = rk.kmeans(torch.from_numpy(X), N, torch.from_numpy(init), allow)
    lab, ctr = mods["kmeans"].kmeans(torch.from_numpy(X), N, torch.from_numpy(init), allow, device=torch.device(DEV))
    assert np.array_equal(lab.cpu().numpy(), lab_w.numpy())
    assert mods["kmeans"].kmeans.last_iterations == it_w
    np.testing.assert_allclose(ctr.cpu().numpy(), ctr_w.numpy(), rtol=1e-6, atol=1e-7)


def test_kmeans_max_iterations_is_an_error(mods):
    km = mods["kmeans"]
    saved = km.max_iterations
    km.max_iterations = 1
    try:
        with pytest.raises(mods["lib"].IsgError):
            km.kmeans(torch.tensor([[0.0, 0.0], [1.0, 1.0], [4.0, 4.0]]), 1, torch.tensor([[3.0, 3.0]]), np.array([100.0], np.float32),
                      device=torch.device(DEV))
    finally:
        km.max_iterations = saved


# ---------------------------------------------------------------------------------------------- K6
def test_mask_iou_golden(mods, golden):
    g = golden("mask_iou")
    H, W = int(g["H"]), int(g["W"])
    dense = unpack_bits(g["masks"], W).astype(np.int32)
    im = mods["image"]
    bits = im.pack_masks(dense)
    assert np.array_equal(bits.cpu().numpy().view(np.uint32), g["masks"])
    for i in range(len(dense)):
        for j in range(len(dense)):
            assert im.compute_iou_for_mask(dense[i], dense[j]) == g["iou"][i, j]
            assert im.is_cover(dense[i], dense[j]) == bool(g["cover"][i, j])


@pytest.mark.parametrize("n,H,W,C,thr,with_boxes", [(60, 96, 130, 3, 0.5, True), (200, 200, 333, 10, 0.3, False), (33, 64, 64, 1, 0.7, True)])
def test_mask_nms_vs_oracle(mods, oracle, n, H, W, C, thr, with_boxes):
    _, rk = oracle
    masks, boxes, scores, cls = mods["synth"].make_masks(n + H, n, H, W, C)
    want = rk.mask_nms(masks, scores, cls, thr)
    got = mods["nms"].mask_nms(masks, scores, cls, thr, bboxes=boxes if with_boxes else None)
    assert np.array_equal(got, np.asarray(want, dtype=np.int64))
    got_agn = mods["nms"].mask_nms(masks, scores, None, thr)
    assert np.array_equal(got_agn, np.asarray(rk.mask_nms(masks, scores, None, thr), dtype=np.int64))


# ---------------------------------------------------------------------------------------------- full size
def test_full_size_batch_properties(mods, oracle):
    """BASELINE full size (1024x2048, ~100 seeds): dense and sparse agree with each other on every keep pixel,
    labels of keep pixels match the oracle on one image, and the per-instance counts are consistent."""
    rd, _ = oracle
    synth, eng = mods["synth"], mods["engine"]
    H, W, N, B = 1024, 2048, 100, 2
    imgs = [synth.make_image(1000 + b, H, W, N) for b in range(B)]
    kp = torch.from_numpy(np.stack([i.kp for i in imgs])).to(DEV)
    ae = torch.from_numpy(np.stack([i.ae for i in imgs])).to(DEV)
    rois = torch.from_numpy(np.stack([i.rois for i in imgs])).to(DEV).contiguous()
    n = torch.full((B,), N, dtype=torch.int32, device=DEV)
    out = {}
    for mode in ("sparse", "dense"):
        plan = eng.DecodePlan(B, H, W, N, 20000, DEV, mode, want_score=True)
        plan.run(kp, ae, rois, n)
        torch.cuda.synchronize()
        out[mode] = {k: getattr(plan, k).cpu().numpy() for k in ("count", "idx", "label", "score", "flag", "offsets", "points", "stats")}
    for k in ("count", "offsets", "stats"):
        assert np.array_equal(out["sparse"][k], out["dense"][k]), k
    for b in range(B):
        M = int(out["sparse"]["count"][b]); T = int(out["sparse"]["offsets"][b, N])
        for k in ("idx", "label", "flag"):
            assert np.array_equal(out["sparse"][k][b, :M], out["dense"][k][b, :M]), k
        assert np.array_equal(out["sparse"]["points"][b, :T], out["dense"]["points"][b, :T])
    core = rd.group_core(torch.from_numpy(imgs[0].kp[0]), torch.from_numpy(imgs[0].ae), imgs[0].rois, 20000)
    M = int(out["sparse"]["count"][0])
    assert M == core["idx"].shape[0]
    assert np.array_equal(out["sparse"]["idx"][0, :M], core["idx"].numpy().astype(np.int32))
    assert np.array_equal(out["sparse"]["label"][0, :M], core["label"].numpy().astype(np.int32))
    np.testing.assert_allclose(out["sparse"]["score"][0, :M], core["score"].numpy(), rtol=RTOL, atol=ATOL)
    assert out["sparse"]["stats"][0, :, 0].sum() == out["sparse"]["flag"][0, :M].sum() == out["sparse"]["offsets"][0, N]
    # device polygon tail at full size: every instance's polygon is a permutation of its point set from the list tail,
    # its vertices come in non-decreasing polar angle about the internal point, and the statistics agree
    dec = mods["decode"]
    plan = eng.DecodePlan(B, H, W, N, 20000, DEV, "dense", want_score=False)
    plan.run(kp, ae, rois, n, tail="polygons", obj_pixel_th=2)
    torch.cuda.synchronize()
    st, ct, fl = plan.inst_start.cpu().numpy(), plan.inst_count.cpu().numpy(), plan.inst_flags.cpu().numpy()
    internal, pts = plan.inst_internal.cpu().numpy(), plan.poly_points.cpu().numpy()
    assert np.array_equal(plan.stats.cpu().numpy(), out["dense"]["stats"])
    key = lambda a: a[np.lexsort((a[:, 0], a[:, 1]))]
    n_poly = 0
    for b in range(B):
        off = out["dense"]["offsets"][b]
        assert int(plan.img_total[b].item()) == off[N]
        for i in range(N):
            want = out["dense"]["points"][b, off[i]:off[i + 1]]
            got = pts[b, st[b, i]:st[b, i] + ct[b, i]]
            assert ct[b, i] == want.shape[0] and np.array_equal(key(got), key(want)), (b, i)
            if fl[b, i] == 1:
                th = dec._polar_angles(got, np.repeat(internal[b, i][None], len(got), 0))
                assert np.all(np.diff(th[~np.isnan(th)]) >= -1e-6), (b, i)      # numpy vs device arctan: <= 1-2 ulp
                n_poly += 1
            else:
                assert fl[b, i] == 0
    assert n_poly > N


# ---------------------------------------------------------------------------------------------- transcendentals
def test_fast_tanh_exp_accuracy(mods):
    """tanh_fast / exp_fast (csrc/common.cuh) observed through the sparse kernel: a 1-seed image whose score is
    exp(-(tanh(a0)+y - cy)^2 * exp(a2)); compared with fp64 over a sweep of arguments."""
    eng = mods["engine"]
    H, W = 64, 256
    rs = np.random.RandomState(0)
    a0 = rs.uniform(-3.0, 3.0, size=(H, W)).astype(np.float32)
    a0[:8] = rs.uniform(-0.6, 0.6, size=(8, W)).astype(np.float32)
    a0[8:12] = rs.uniform(-1e-3, 1e-3, size=(4, W)).astype(np.float32)
    a2 = rs.uniform(-3.0, 6.0, size=(H, W)).astype(np.float32)
    ae = np.zeros((4, H, W), np.float32); ae[0] = a0; ae[1] = 0.0; ae[2] = a2; ae[3] = -30.0   # x term ~ 0
    kp = mods["synth"]._distinct_float32(rs.normal(0, 1, size=(1, H, W)).astype(np.float32))
    rois = np.array([[-0.5, -0.5, W - 0.5, H - 0.5]], np.float32)          # one box covering the image
    plan = eng.DecodePlan(1, H, W, 1, H * W, DEV, "dense", want_score=True)
    plan.run(torch.from_numpy(kp)[None].to(DEV), torch.from_numpy(ae)[None].to(DEV), torch.from_numpy(rois)[None].to(DEV),
             torch.tensor([1], dtype=torch.int32, device=DEV))
    torch.cuda.synchronize()
    got = plan.score_map[0].cpu().numpy().astype(np.float64)
    ys = torch.linspace(0, 1, 1024)[:H].double().numpy(); xs = torch.linspace(0, 2, 2048)[:W].double().numpy()
    cy, cx = ys[int((H - 1) / 2)], xs[int((W - 1) / 2)]
    ey = (np.tanh(a0.astype(np.float64)).astype(np.float32) + ys.astype(np.float32)[:, None]).astype(np.float64)
    q = (ey - cy) ** 2 * np.exp(a2.astype(np.float64)) + (xs[None, :] - cx) ** 2 * np.exp(-30.0)
    want = np.exp(-q)
    ok = want > 1e-30
    rel = np.abs(got[ok] - want[ok]) / want[ok]
    # error budget: fp32 rounding of (e-c), its square, the product and the sum (~4 * 6e-8 * q) plus 2-ulp exp
    bound = 4e-6 + 1e-6 * q[ok]
    assert np.all(rel < bound), float((rel / bound).max())


# ---------------------------------------------------------------------------------------------- round-1 additions
def test_dense_tile_list_overflow(mods, oracle):
    """crowded image: tiles overlapped by more seed boxes than the records staged next to a tile (16), so the dense
    kernel walks the overflow index list; labels must still be the oracle's"""
    rd, _ = oracle
    H, W, N = 128, 256, 60
    img = mods["synth"].make_image(4242, H, W, N)
    x1, y1, x2, y2 = img.rois.T
    worst = 0
    for ty in range(0, H, 16):
        for tx in range(0, W, 128):
            hit = (np.ceil(y1) <= ty + 15) & (np.floor(y2) >= ty) & (np.ceil(x1) <= tx + 127) & (np.floor(x2) >= tx)
            worst = max(worst, int(hit.sum()))
    assert worst > 16, worst
    core = rd.group_core(torch.from_numpy(img.kp[0]), torch.from_numpy(img.ae), img.rois, 3000)
    plan = _run_plan(mods, img, 3000, "dense")
    M = int(plan.count[0].item())
    assert M == core["idx"].shape[0]
    assert np.array_equal(plan.label[0, :M].cpu().numpy(), core["label"].numpy().astype(np.int32))
    np.testing.assert_allclose(plan.score[0, :M].cpu().numpy(), core["score"].numpy(), rtol=RTOL, atol=ATOL)
    score, label = rd.dense_labels(torch.from_numpy(img.ae), img.rois)
    P = _all_memberships(rd, img)
    top2 = np.sort(P, axis=2)[:, :, -2:]
    safe = ((top2[:, :, 1] - top2[:, :, 0]) > 1e-5 * np.maximum(top2[:, :, 1], 1e-30)) | (top2[:, :, 1] == 0)
    assert safe.mean() > 0.98
    assert np.array_equal(plan.label_map[0].cpu().numpy()[safe], label.numpy()[safe].astype(np.int32))
    # the sparse kernel (independent code path: exp + arg-max) agrees on every keep pixel
    sp = _run_plan(mods, img, 3000, "sparse")
    assert np.array_equal(sp.label[0, :M].cpu().numpy(), plan.label[0, :M].cpu().numpy())


def test_decode_output_host_chunks_equal_device_batch(mods):
    """host tensors are uploaded and decoded in chunks (uneven last chunk); same detections as one device batch"""
    synth, dec = mods["synth"], mods["decode"]
    H, W, C, B = 128, 256, 8, 5
    anchors = synth.make_anchors(H, W)
    scenes = [synth.make_scene(700 + b, H, W, [6, 3, 0, 9, 5][b], C, anchors) for b in range(B)]
    kp = torch.from_numpy(np.stack([s[0].kp for s in scenes])); ae = torch.from_numpy(np.stack([s[0].ae for s in scenes]))
    reg = torch.from_numpy(np.stack([s[1] for s in scenes])); cls = torch.from_numpy(np.stack([s[2] for s in scenes]))
    anc = torch.from_numpy(anchors)
    infos = [TransInfo("/nonexistent.png", (H, W))] * B
    cfg, tf, inputs = DecodeCfg(kp_th=2000), IdentityTransforms(), torch.zeros((B, 3, H, W))
    saved = dec.decode_mode, dec.host_chunk_images
    dec.decode_mode, dec.host_chunk_images = "dense", 2
    try:
        got = dec.decode_output(inputs, ((kp.pin_memory(), ae.pin_memory(), None), reg.pin_memory(), cls.pin_memory(), anc),
                                infos, tf, cfg, torch.device(DEV))
        want = dec.decode_output(inputs, ((kp.to(DEV), ae.to(DEV), None), reg.to(DEV), cls.to(DEV), anc.to(DEV)), infos, tf,
                                 cfg, torch.device(DEV))
    finally:
        dec.decode_mode, dec.host_chunk_images = saved
    assert len(got) == len(want) == B and len(got[2]) == 0 and sum(len(g) for g in got) > 5
    for g, w in zip(got, want):
        assert len(g) == len(w)
        for (c1, f1, k1, p1), (c2, f2, k2, p2) in zip(g, w):
            assert int(c1) == int(c2) and f1 == f2 and np.array_equal(k1, k2) and np.array_equal(p1, p2)


@pytest.mark.parametrize("B", [4, 5])
def test_decode_output_zero_copy_equals_device_batch(mods, B):
    """pinned host outputs: only kp / classification are uploaded, ae and regression are gathered by the kernels out of
    the pinned buffers (isg_decode_step, sparse assignment); the detections are those of the dense device-resident decode"""
    synth, dec = mods["synth"], mods["decode"]
    H, W, C = 256, 512, 8
    anchors = synth.make_anchors(H, W)
    scenes = [synth.make_scene(760 + b, H, W, [14, 0, 9, 22, 5][b], C, anchors) for b in range(B)]
    kp = torch.from_numpy(np.stack([s[0].kp for s in scenes])); ae = torch.from_numpy(np.stack([s[0].ae for s in scenes]))
    reg = torch.from_numpy(np.stack([s[1] for s in scenes])); cls = torch.from_numpy(np.stack([s[2] for s in scenes]))
    anc = torch.from_numpy(anchors)
    infos = [TransInfo("/nonexistent.png", (H, W))] * B
    cfg, tf, inputs = DecodeCfg(kp_th=3000), IdentityTransforms(), torch.zeros((B, 3, H, W))
    pinned = ((kp.pin_memory(), ae.pin_memory(), None), reg.pin_memory(), cls.pin_memory(), anc)
    ae_before = pinned[0][1].clone()
    saved = dec.decode_mode, dec.host_chunk_images, dec.host_zero_copy
    dec.decode_mode, dec.host_chunk_images = "dense", 2
    try:
        dec.host_zero_copy = True
        got = dec.decode_output(inputs, pinned, infos, tf, cfg, torch.device(DEV))
        h2d = dec.last_timing["h2d_bytes"]
        again = dec.decode_output(inputs, pinned, infos, tf, cfg, torch.device(DEV))
        dec.host_zero_copy = False
        uploaded = dec.decode_output(inputs, pinned, infos, tf, cfg, torch.device(DEV))
        want = dec.decode_output(inputs, ((kp.to(DEV), ae.to(DEV), None), reg.to(DEV), cls.to(DEV), anc.to(DEV)), infos, tf,
                                 cfg, torch.device(DEV))
    finally:
        dec.decode_mode, dec.host_chunk_images, dec.host_zero_copy = saved
    assert h2d == B * (kp[0].numel() + cls[0].numel()) * 4                   # kp + classification only
    assert torch.equal(pinned[0][1], ae_before)                              # inputs are not modified
    assert len(got) == len(want) == B and len(got[1]) == 0 and sum(len(g) for g in got) > 20
    for res in (got, again, uploaded):
        for g, w in zip(res, want):
            assert len(g) == len(w)
            for (c1, f1, k1, p1), (c2, f2, k2, p2) in zip(g, w):
                assert int(c1) == int(c2) and f1 == f2 and np.array_equal(k1, k2) and np.array_equal(p1, p2)


def test_decode_ring_overlapped_steps_equal_isolated_steps(mods):
    """engine.DecodeRing: steps submitted back to back run concurrently on independent pipelines (isg_decode_step, one host
    call per step); every slot must hold exactly the result of its own batch, for the dense and the sparse assignment."""
    synth, engine = mods["synth"], mods["engine"]
    B, H, W, C = 2, 256, 512, 8
    dev = torch.device(DEV)
    anchors = synth.make_anchors(H, W)
    anc = torch.from_numpy(anchors).to(dev)
    A = anchors.reshape(-1, 4).shape[0]

    def batch(seed, counts):
        sc = [synth.make_scene(seed + b, H, W, counts[b], C, anchors) for b in range(B)]
        return [torch.from_numpy(np.stack(x)).to(dev) for x in ([s[0].kp for s in sc], [s[0].ae for s in sc], [s[1] for s in sc], [s[2] for s in sc])]

    batches = [batch(1200, [12, 3]), batch(1210, [0, 25]), batch(1220, [7, 7])]
    make = lambda: engine.make_pipeline(B, A, C, H, W, H, W, 3000, dev, cand_cap=1024, max_keep=64)

    def tables(pipe):
        pipe.bplan.arena.wait()
        h = {**pipe.bplan.host, **pipe.dplan.host}
        n = h["n_keep"].numpy().copy()
        out = {}
        for b in range(B):
            for i in range(int(n[b])):
                s0, c0 = int(h["inst_start"][b, i]), int(h["inst_count"][b, i])
                out[(b, i)] = (int(h["inst_flags"][b, i]), h["poly_points"][b, s0:s0 + c0].numpy().copy())
        return n, h["rois"].numpy().copy(), out

    # reference results: the Python multi-call pipeline, one batch at a time
    ref_pipe = engine.DecodePipeline(engine.BoxPlan(B, A, C, H, W, dev, cap=1024, max_keep=64),
                                     engine.DecodePlan(B, H, W, 64, 3000, dev, "dense", want_score=False, wh_delta=0.1))
    want = []
    for x in batches:
        ref_pipe.run(x[0], x[1], anc, x[2], x[3], 0.3, 0.2, tail="polygons", obj_pixel_th=2)
        torch.cuda.synchronize(dev)
        bp, dp = ref_pipe.bplan, ref_pipe.dplan
        n = bp.n_keep.cpu().numpy()
        st, ct, fl, pts = dp.inst_start.cpu().numpy(), dp.inst_count.cpu().numpy(), dp.inst_flags.cpu().numpy(), dp.poly_points.cpu().numpy()
        want.append((n, bp.rois.cpu().numpy(), {(b, i): (int(fl[b, i]), pts[b, st[b, i]:st[b, i] + ct[b, i]].copy())
                                                for b in range(B) for i in range(int(n[b]))}))
    for assign in ("dense", "sparse"):
        ring = engine.DecodeRing(make, 3)
        for seq in ([0, 1, 2], [2, 2, 0, 1, 1, 0, 2, 1, 0]):
            slots = []
            for k in seq:
                x = batches[k]
                slots.append((ring.submit(x[0], x[1], anc, x[2], x[3], 0.3, 0.2, obj_pixel_th=2, assign=assign, fetch=True), k))
                if len(slots) == 3:               # a slot is read before it is reused
                    slot, kk = slots.pop(0)
                    n, rois, got = tables(ring.pipes[slot])
                    wn, wrois, wgot = want[kk]
                    assert np.array_equal(n, wn) and got.keys() == wgot.keys()
                    assert all(np.array_equal(rois[b, :n[b]], wrois[b, :n[b]]) for b in range(B))
                    for key in got:
                        assert got[key][0] == wgot[key][0] and np.array_equal(got[key][1], wgot[key][1]), (assign, key)
            for slot, kk in slots:
                n, rois, got = tables(ring.pipes[slot])
                wn, wrois, wgot = want[kk]
                assert np.array_equal(n, wn) and got.keys() == wgot.keys()
                for key in got:
                    assert got[key][0] == wgot[key][0] and np.array_equal(got[key][1], wgot[key][1]), (assign, key)


def test_polygon_stage_large_instance(mods):
    """an instance with more boundary points than fit in shared memory (2048) is finished by the global-memory variant
    of the device stage: same polygon as the all-host path (up to the order of equal-angle points)"""
    dec = mods["decode"]
    H, W = 256, 512
    rs = np.random.RandomState(11)
    kp = torch.from_numpy(rs.permutation(H * W).astype(np.float32).reshape(H, W) / 1000.0)     # distinct values
    ae = torch.zeros((4, H, W)); ae[2:] = np.log(200.0)
    centre, wh = [np.array([128.5, 256.5], np.float32)], [np.array([250.0, 500.0], np.float32)]
    cfg = DecodeCfg(kp_th=30000)
    saved = dec.decode_mode, dec.device_polygon_stage
    out = {}
    try:
        for name, flag in (("device", True), ("host", False)):
            dec.decode_mode, dec.device_polygon_stage = "dense", flag
            out[name] = dec.group_kp(kp.to(DEV), ae.to(DEV), IdentityTransforms(), wh, centre, [3], [0.9],
                                     TransInfo("x", (H, W)), cfg, torch.device(DEV))
    finally:
        dec.decode_mode, dec.device_polygon_stage = saved
    assert len(out["host"][3]) == len(out["device"][3]) == 1
    for p_host, p_dev, ctr in zip(out["host"][3], out["device"][3], out["host"][2]):
        assert p_host.shape[0] > 2048
        assert_polygon_equivalent(dec, p_dev, p_host, ctr)


def test_pipelined_steps_equal_isolated_steps(mods):
    """engine.DecodePipeline.run(pipelined=True): the polygon tail of a step overlaps the head of the next one; the
    results of every step must be those of the same step run on its own."""
    synth, engine = mods["synth"], mods["engine"]
    B, H, W, C = 3, 256, 512, 8
    dev = torch.device(DEV)
    anchors = synth.make_anchors(H, W)
    anc = torch.from_numpy(anchors).to(dev)

    def batch(seed, counts):
        sc = [synth.make_scene(seed + b, H, W, counts[b], C, anchors) for b in range(B)]
        return [torch.from_numpy(np.stack(x)).to(dev) for x in ([s[0].kp for s in sc], [s[0].ae for s in sc], [s[1] for s in sc], [s[2] for s in sc])]

    batches = [batch(900, [12, 0, 20]), batch(940, [5, 17, 9])]
    bplan = engine.BoxPlan(B, anchors.reshape(-1, 4).shape[0], C, H, W, dev, cap=1024, max_keep=64)
    dplan = engine.DecodePlan(B, H, W, bplan.N, 3000, dev, "dense", want_score=False, wh_delta=0.1)
    pipe = engine.DecodePipeline(bplan, dplan)

    def snapshot():
        torch.cuda.synchronize(dev)
        out = {}
        cnt = dplan.inst_count.cpu().numpy(); st = dplan.inst_start.cpu().numpy(); pts = dplan.poly_points.cpu().numpy()
        n = bplan.n_seeds.cpu().numpy()
        for b in range(B):
            for i in range(int(n[b])):
                out[(b, i)] = (int(dplan.inst_flags[b, i]), pts[b, st[b, i]: st[b, i] + cnt[b, i]].copy())
        return n.copy(), out

    def run(x, pipelined):
        pipe.run(x[0], x[1], anc, x[2], x[3], 0.3, 0.2, tail="polygons", obj_pixel_th=2, pipelined=pipelined)

    want = []
    for x in batches:
        run(x, False)
        want.append(snapshot())
    for seq in ([0, 1], [0, 1, 0], [1, 1, 0, 1, 0, 1, 0, 0, 1]):
        for k in seq:
            run(batches[k], True)
        pipe.finish()
        n, got = snapshot()
        wn, wgot = want[seq[-1]]
        assert np.array_equal(n, wn) and got.keys() == wgot.keys()
        for key in got:
            assert got[key][0] == wgot[key][0] and np.array_equal(got[key][1], wgot[key][1]), key
    # a non-pipelined step behind a pipelined one waits for the pending tail by itself
    run(batches[0], True)
    run(batches[1], False)
    n, got = snapshot()
    assert np.array_equal(n, want[1][0]) and all(np.array_equal(got[k][1], want[1][1][k][1]) for k in got)


# ---------------------------------------------------------------------------------------------- split dense step
def _split_vs_fused(mods, kp, ae, rois, n_seeds, kp_th, N):
    """label map + keep bits of isg_assign_dense (fused) and of isg_topk_keep + isg_assign_labels (split), through the C ABI"""
    lib, engine = mods["lib"], mods["engine"]
    call, ptr, sp = lib.call, engine.ptr, engine.stream_ptr
    B, _, H, W = ae.shape
    dev = torch.device(DEV)
    plan = engine.DecodePlan(B, H, W, N, kp_th, dev, "dense", want_score=False)
    kp_d, ae_d = kp.to(dev).contiguous(), ae.to(dev).contiguous()
    rois_d, ns_d = rois.to(dev).contiguous(), n_seeds.to(dev)
    plan.run(kp_d, ae_d, rois_d, ns_d)                                   # fused: thr_key, seeds, label_map, keepbits
    torch.cuda.synchronize()
    fused = plan.label_map.clone(), plan.keepbits.clone()
    thr_f = plan.thr_key.clone()
    kb = torch.full_like(plan.keepbits, -1)                              # garbage: the call must zero the plane itself
    lab = torch.full_like(plan.label_map, -7)
    s = sp(dev)
    k2 = kp_d[:, 0] if kp_d.dim() == 4 else kp_d
    plan.thr_key.fill_(0)
    call("isg_topk_keep", ptr(k2), B, H, W, k2.stride(0) if B > 1 else H * W, kp_th, ptr(plan.thr_key), ptr(kb), 0, plan.ws_ptr,
         plan.ws_bytes, s)
    torch.cuda.synchronize()
    assert torch.equal(plan.thr_key, thr_f)
    rc = lib.lib().isg_assign_labels(ptr(ae_d), ae_d.stride(0) if B > 1 else 4 * H * W, ae_d.stride(1), ptr(plan.seeds),
                                     ptr(plan.ghost), ptr(ns_d), B, N, H, W, ptr(plan.ys), ptr(plan.xs), ptr(lab), None,
                                     ptr(plan.dense_ws), plan.dense_ws_bytes, 0, s)
    torch.cuda.synchronize()
    return fused, (lab, kb), rc


@pytest.mark.parametrize("shape,N,kp_th,kind", [
    ((512, 1024), 40, 20000, "scene"),       # sample -> filter -> select: complete candidate list
    ((384, 640), 20, 3000, "scene"),         # ragged tiles (H % 16, W % 128 != 0)
    ((128, 256), 6, 500, "scene"),           # small image: no candidate list (whole-image select) -> full pass inside the kernel
    ((256, 512), 6, 100000, "scene"),        # large k: multi-CTA radix select, no candidate list -> streaming keep kernel
    ((256, 512), 6, 100, "flat"),            # plateau: every pixel ties at the threshold, the candidate list overflows
    ((256, 512), 6, 5000, "negative"),       # selected negative values next to unselected (0) neighbours are dropped
])
def test_split_dense_step_equals_fused(mods, shape, N, kp_th, kind):
    synth = mods["synth"]
    H, W = shape
    B = 2
    imgs = [synth.make_image(900 + b, H, W, N) for b in range(B)]
    kp = torch.from_numpy(np.stack([im.kp for im in imgs]))
    if kind == "flat":
        kp = torch.zeros_like(kp); kp[1] = 0.5
    elif kind == "negative":
        kp = -kp.abs() - 0.25
    ae = torch.from_numpy(np.stack([im.ae for im in imgs]))
    rois = torch.from_numpy(np.stack([im.rois for im in imgs]))
    n_seeds = torch.tensor([N] * B, dtype=torch.int32)
    (lab_f, kb_f), (lab_s, kb_s), rc = _split_vs_fused(mods, kp, ae, rois, n_seeds, kp_th, N)
    assert rc == 0
    assert torch.equal(kb_s, kb_f), "keep bits differ"
    assert torch.equal(lab_s, lab_f), "label maps differ"
    if kind == "scene":
        assert int(unpack_bits(kb_f.cpu().numpy(), W).sum()) > 0


def test_assign_labels_refuses_widths_the_tensor_map_cannot_take(mods):
    """W % 4 != 0: isg_assign_labels answers ISG_EUNSUPPORTED (the step then runs the fused form, which serves any width)"""
    synth = mods["synth"]
    H, W, N = 96, 258, 4
    im = synth.make_image(950, H, W, N)
    (_, _), (_, _), rc = _split_vs_fused(mods, torch.from_numpy(im.kp)[None], torch.from_numpy(im.ae)[None],
                                         torch.from_numpy(im.rois)[None], torch.tensor([N], dtype=torch.int32), 300, N)
    assert rc == -3          # ISG_EUNSUPPORTED


def test_decode_ring_split_keep_equals_fused_steps(mods):
    """whole steps through isg_decode_step: split (keep bits from the top-k candidates + labels-only dense kernel) against
    the fused dense kernel - identical polygon tables"""
    synth, engine = mods["synth"], mods["engine"]
    H, W, B, N = 256, 512, 2, 10
    dev = torch.device(DEV)
    anchors = synth.make_anchors(H, W)
    scenes = [synth.make_scene(960 + b, H, W, N, 8, anchors) for b in range(B)]
    kp = torch.from_numpy(np.stack([s[0].kp for s in scenes])).to(dev)
    ae = torch.from_numpy(np.stack([s[0].ae for s in scenes])).to(dev)
    reg = torch.from_numpy(np.stack([s[1] for s in scenes])).to(dev); cls = torch.from_numpy(np.stack([s[2] for s in scenes])).to(dev)
    anc = torch.from_numpy(anchors).to(dev)
    res = {}
    saved = engine.SPLIT_KEEP
    try:
        for split in (False, True):
            engine.SPLIT_KEEP = split
            pipe = engine.make_pipeline(B, anchors.shape[1], 8, H, W, H, W, 3000, dev, cand_cap=512, max_keep=64)
            pipe.run_native(kp, ae, anc, reg, cls, 0.3, 0.2, obj_pixel_th=2)
            torch.cuda.synchronize()
            dp = pipe.dplan
            res[split] = [t.clone() for t in (dp.label_map, dp.keepbits, dp.img_total, dp.inst_count, dp.inst_flags)]
            st, ct, pts = dp.inst_start.cpu().numpy(), dp.inst_count.cpu().numpy(), dp.poly_points.cpu().numpy()
            # an instance's slot in poly_points depends on the order the CTAs finish in: compare instance by instance
            res[split].append([[pts[b, st[b, i]:st[b, i] + ct[b, i]].copy() for i in range(dp.N)] for b in range(B)])
    finally:
        engine.SPLIT_KEEP = saved
    assert int(res[True][2].sum().item()) > 0
    for a, b in zip(res[False][:5], res[True][:5]):
        assert torch.equal(a, b)
    for pa, pb in zip(res[False][5], res[True][5]):
        assert len(pa) == len(pb) and all(np.array_equal(x, y) for x, y in zip(pa, pb))
