// Host-side helpers of the polygon stage (a7, utils/decode.py:51-68,167-204 of the reference).  Plain C++, no CUDA:
// the polygon stage is host glue this round (SURVEY.md §8 f1 moves it to the GPU).  These functions only remove the
// Python-loop overhead of many tiny cv2 calls; their arithmetic restates OpenCV's so results are identical.
//
// point_in_polygon restates cv::pointPolygonTest(contour CV_32F, pt, measureDist=false)
// (OpenCV 4.x modules/imgproc/src/geometry.cpp, the non-integer branch): a crossing count over the closed polyline
// with the on-edge / on-vertex cases returning 0.  tests/test_host_logic.py checks it against cv2 itself.
#include <cstdint>
#include "../../include/isg.h"

namespace {

inline int point_in_polygon(const float* pts, int K, float px, float py) {
  if (K <= 0) return -1;
  int counter = 0;
  float vx = pts[2 * (K - 1)], vy = pts[2 * (K - 1) + 1];
  for (int i = 0; i < K; ++i) {
    const float v0x = vx, v0y = vy;
    vx = pts[2 * i]; vy = pts[2 * i + 1];
    if ((v0y <= py && vy <= py) || (v0y > py && vy > py) || (v0x < px && vx < px)) {
      if (py == vy && (px == vx || (py == v0y && ((v0x <= px && px <= vx) || (vx <= px && px <= v0x))))) return 0;
      continue;
    }
    double dist = (double)(py - v0y) * (vx - v0x) - (double)(px - v0x) * (vy - v0y);
    if (dist == 0) return 0;
    if (vy < v0y) dist = -dist;
    counter += dist > 0;
  }
  return counter % 2 == 0 ? -1 : 1;
}

}  // namespace

extern "C" int isg_host_point_in_polygon(const float* pts, int K, float px, float py) {
  if (!pts || K < 0) return ISG_EINVAL - 100;   // outside {-1,0,1}
  return point_in_polygon(pts, K, px, py);
}

// find_internal_point (utils/decode.py:51-68) for n instances.  points: [*,2] fp32 (x,y) grouped by instance,
// offsets: [n+1].  centers: [n,2] (x,y).  Instances with fewer than min_pts points are skipped (internal = centre).
// internal[i] = centre if it is strictly inside the (unordered) point polyline, else the fp32 mean of the points if
// that is inside, else the first pair midpoint (kps[i]+kps[j])/2, i in [0,K), j in [1,K), that is inside, else centre.
extern "C" int isg_host_internal_points(const float* points, const int32_t* offsets, int n, const float* centers,
                                        int min_pts, float* internal) {
  if (!points || !offsets || !centers || !internal || n < 0) return ISG_EINVAL;
  for (int q = 0; q < n; ++q) {
    const float* p = points + 2 * (size_t)offsets[q];
    const int K = offsets[q + 1] - offsets[q];
    const float cx = centers[2 * q], cy = centers[2 * q + 1];
    internal[2 * q] = cx; internal[2 * q + 1] = cy;
    if (K < min_pts || K <= 0) continue;
    if (point_in_polygon(p, K, cx, cy) > 0) continue;
    // numpy mean(axis=0) of a C-contiguous [K,2] fp32 array: sequential fp32 sums over the rows, then / K in fp32
    float sx = 0.0f, sy = 0.0f;
    for (int i = 0; i < K; ++i) { sx += p[2 * i]; sy += p[2 * i + 1]; }
    const float mx = sx / (float)K, my = sy / (float)K;
    if (point_in_polygon(p, K, mx, my) > 0) { internal[2 * q] = mx; internal[2 * q + 1] = my; continue; }
    bool found = false;
    for (int i = 0; i < K && !found; ++i)
      for (int j = 1; j < K; ++j) {
        const float qx = (p[2 * i] + p[2 * j]) / 2.0f, qy = (p[2 * i + 1] + p[2 * j + 1]) / 2.0f;
        if (point_in_polygon(p, K, qx, qy) > 0) { internal[2 * q] = qx; internal[2 * q + 1] = qy; found = true; break; }
      }
  }
  return ISG_OK;
}

// inside[i] = pointPolygonTest(polygon i, centre i) > 0 for n polygons stored back to back (offsets [n+1])
extern "C" int isg_host_centres_inside(const float* points, const int32_t* offsets, int n, const float* centers,
                                       uint8_t* inside) {
  if (!points || !offsets || !centers || !inside || n < 0) return ISG_EINVAL;
  for (int q = 0; q < n; ++q) {
    const int K = offsets[q + 1] - offsets[q];
    inside[q] = (K > 0 && point_in_polygon(points + 2 * (size_t)offsets[q], K, centers[2 * q], centers[2 * q + 1]) > 0) ? 1 : 0;
  }
  return ISG_OK;
}
