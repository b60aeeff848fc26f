"""The polygon-fill restatement (oracle/ref_fill.py) pinned against the library the reference calls
(`cv2.fillPoly`, utils/image.py:185) — CPU only."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
from oracle import ref_fill


def polygon_cases(seed, n, max_hw=(160, 240)):
    rng = np.random.RandomState(seed)
    for t in range(n):
        H, W = rng.randint(4, max_hw[0]), rng.randint(4, max_hw[1])
        k = rng.randint(1, 48)
        kind = t % 4
        if kind == 0:                                   # arbitrary (self-intersecting) vertex lists
            pts = np.stack([rng.randint(0, W, k), rng.randint(0, H, k)], 1)
        elif kind == 1:                                 # star-shaped: what the angular sort produces
            ang = np.sort(rng.uniform(0, 2 * np.pi, k))
            r = rng.uniform(0.3, 1.0, k)
            pts = np.stack([W / 2 + r * np.cos(ang) * W / 2.2, H / 2 + r * np.sin(ang) * H / 2.2], 1)
        elif kind == 2:                                 # tight cluster: repeated vertices, horizontal edges, ties
            pts = np.stack([rng.randint(W // 2 - 5, W // 2 + 6, k), rng.randint(H // 2 - 3, H // 2 + 4, k)], 1)
        else:                                           # jagged outline of an ellipse, angle-sorted
            k = rng.randint(20, 300)
            ang = np.sort(rng.uniform(0, 2 * np.pi, k))
            pts = np.stack([W / 2 + (W / 2 - 2) * np.cos(ang) + rng.randint(-1, 2, k),
                            H / 2 + (H / 2 - 2) * np.sin(ang) + rng.randint(-1, 2, k)], 1)
        pts = np.clip(pts, 0, [W - 1, H - 1]).astype(np.float32)
        yield (H, W), pts


@pytest.mark.parametrize("seed", [0, 1])
def test_fill_restatement_matches_opencv(seed):
    for size, pts in polygon_cases(seed, 400):
        ref = cv2.fillPoly(np.zeros(size, np.int32), [pts.astype(np.int32)], 1)
        assert np.array_equal(ref_fill.fill_poly(pts, size), ref), (size, pts.tolist())
        assert np.array_equal(ref_fill.fill_poly_parity(pts, size), ref), (size, pts.tolist())


def test_fill_restatement_long_edges():
    rng = np.random.RandomState(5)
    for _ in range(12):
        size = (1024, 2048)
        k = rng.randint(3, 12)
        pts = np.stack([rng.randint(0, 2048, k), rng.randint(0, 1024, k)], 1).astype(np.float32)
        ref = cv2.fillPoly(np.zeros(size, np.int32), [pts.astype(np.int32)], 1)
        assert np.array_equal(ref_fill.fill_poly(pts, size), ref)


def test_fill_restatement_rejects_outside_vertices():
    with pytest.raises(ValueError):
        ref_fill.fill_poly(np.array([[0, 0], [9, 3], [3, 12]], np.float32), (8, 8))
