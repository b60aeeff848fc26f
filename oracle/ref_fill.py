"""ORACLE — test infrastructure only.  CPU restatement of the polygon rasteriser behind
`poly_to_mask` (reference utils/image.py:180-185, called by the results writer utils/eval_util.py:116).

The reference delegates the arithmetic to a third-party library: `cv2.fillPoly(mask, [poly], 1)`
(OpenCV; the reference pins opencv-python without a version, this image has OpenCV 4.13).  OpenCV's
sources are not under /root/reference, so this file restates the published algorithm of
`fillPoly` for integer vertices, 8-connected lines, shift 0:

  * every polygon edge is drawn with the 8-connected Bresenham iterator, always walked from its left
    end point to its right one (`Line` -> `LineIterator(..., leftToRight=true)`);
  * the interior comes from the scan-line edge table (`CollectPolyEdges` / `FillEdgeCollection`): an
    edge (x0,y0)-(x1,y1), y0 < y1, is active on rows y0 <= y < y1, its abscissa is the 16.16 fixed-point
    x0 + (y - y0) * dx with dx = ((x1 - x0) << 16) / (y1 - y0) (C integer division, truncating); on every
    row the sorted abscissae are paired and the integer pixels ceil(xa) .. floor(xb) of each pair are set.

Parity status: pinned against the library itself — tests/test_fill_oracle.py fuzzes this restatement
against `cv2.fillPoly` of the installed OpenCV (random, star-shaped, degenerate and decode-produced
polygons).  Only vertices inside the frame are covered (the decode path never produces others); vertices
outside the frame go through OpenCV's line clipping, which is not restated here.

Only tests/ (and smoke / the bench's CPU legs) may import this module.
"""
from __future__ import annotations

import numpy as np

XY_SHIFT = 16
XY_ONE = 1 << XY_SHIFT


def _line8(mask: np.ndarray, x0: int, y0: int, x1: int, y1: int) -> None:
    """8-connected Bresenham line, walked left to right (major axis steps every pixel; the minor axis
    steps when the running error is negative)."""
    if x1 < x0:
        x0, y0, x1, y1 = x1, y1, x0, y0
    dx, dy = x1 - x0, y1 - y0
    sy = -1 if dy < 0 else 1
    dy = abs(dy)
    x, y = x0, y0
    if dy > dx:                          # y is the major axis
        err, plus, minus = dy - 2 * dx, 2 * dy, -2 * dx
        for _ in range(dy + 1):
            mask[y, x] = 1
            neg = err < 0
            err += minus + (plus if neg else 0)
            y += sy
            if neg:
                x += 1
    else:
        err, plus, minus = dx - 2 * dy, 2 * dx, -2 * dy
        for _ in range(dx + 1):
            mask[y, x] = 1
            neg = err < 0
            err += minus + (plus if neg else 0)
            x += 1
            if neg:
                y += sy


def _cdiv(a: int, b: int) -> int:
    """C integer division (truncation toward zero)."""
    q = abs(a) // abs(b)
    return q if (a < 0) == (b < 0) else -q


def fill_poly(poly: np.ndarray, img_size) -> np.ndarray:
    """`cv2.fillPoly(np.zeros(img_size, int32), [poly.astype(int32)], 1)` for vertices inside the frame."""
    H, W = int(img_size[0]), int(img_size[1])
    pts = [(int(p[0]), int(p[1])) for p in np.asarray(poly).astype(np.int32).reshape(-1, 2)]
    mask = np.zeros((H, W), np.int32)
    for x, y in pts:
        if not (0 <= x < W and 0 <= y < H):
            raise ValueError("vertex outside the frame")
    edges = []
    for i in range(len(pts)):
        (x0, y0), (x1, y1) = pts[i - 1], pts[i]
        _line8(mask, x0, y0, x1, y1)
        if y0 == y1:
            continue
        dxf = _cdiv((x1 - x0) << XY_SHIFT, y1 - y0)
        edges.append((y0, y1, x0 << XY_SHIFT, dxf) if y0 < y1 else (y1, y0, x1 << XY_SHIFT, dxf))
    if len(edges) < 2:
        return mask
    for y in range(min(e[0] for e in edges), max(e[1] for e in edges)):
        xs = sorted(e[2] + (y - e[0]) * e[3] for e in edges if e[0] <= y < e[1])
        for k in range(0, len(xs) - 1, 2):
            a = (xs[k] + XY_ONE - 1) >> XY_SHIFT
            b = xs[k + 1] >> XY_SHIFT
            if a <= b:
                mask[y, a:b + 1] = 1
    return mask


def fill_poly_parity(poly: np.ndarray, img_size) -> np.ndarray:
    """The same rasterisation in the sort-free form the device kernel uses: on a row, pixel p is inside a
    pair iff an odd number of active edges lie strictly left of p, or an edge passes exactly through p.
    Checked equal to `fill_poly` in the tests."""
    H, W = int(img_size[0]), int(img_size[1])
    pts = [(int(p[0]), int(p[1])) for p in np.asarray(poly).astype(np.int32).reshape(-1, 2)]
    mask = np.zeros((H, W), np.int32)
    toggles = np.zeros((H, W + 1), np.int32)
    exact = np.zeros((H, W), np.int32)
    for i in range(len(pts)):
        (x0, y0), (x1, y1) = pts[i - 1], pts[i]
        _line8(mask, x0, y0, x1, y1)
        if y0 == y1:
            continue
        dxf = _cdiv((x1 - x0) << XY_SHIFT, y1 - y0)
        ya, yb, xa = (y0, y1, x0 << XY_SHIFT) if y0 < y1 else (y1, y0, x1 << XY_SHIFT)
        for y in range(ya, yb):
            x = xa + (y - ya) * dxf
            toggles[y, (x >> XY_SHIFT) + 1] ^= 1
            if (x & (XY_ONE - 1)) == 0:
                exact[y, x >> XY_SHIFT] = 1
    inside = np.bitwise_xor.accumulate(toggles, axis=1)[:, :W]
    return mask | inside | exact
