// f1 — per-instance point sets and polygon extraction on the device.
// Reference: the per-instance loop of group_kp (utils/decode.py:337-356) and aug_group (:167-204) with
// find_internal_point (:51-68) and cartesian2polar (:88-113), for the identity val-transform (the default).
//
// One CTA per (image, instance):
//   1. collect, in row-major order, the keep pixels (keepbits) whose label_map entry is the instance and that lie
//      strictly inside its ghost bounds (:351-352); the scan is restricted to the integer pixel range of the bounds;
//   2. fewer than obj_pixel_th points -> no polygon (:355);
//   3. internal point: the box centre if cv2.pointPolygonTest(points, centre) > 0, else the fp32 mean of the
//      points, else the first pair midpoint (kps[i]+kps[j])/2 (i outer, j from 1) that is inside, else the centre;
//   4. polar angle of every point about the internal point in fp32 (same branches as cartesian2polar), ascending
//      sort; EQUAL angles keep their row-major order (np.argsort's order among equal keys is unspecified);
//   5. polygon is valid iff the box centre is strictly inside the sorted polygon (:201).
// The `area == 0` rejection (:187-189) cannot fire for a polygon with at least one vertex (fillPoly rasterises the
// outline) and is not evaluated.  pointPolygonTest is restated from OpenCV's float branch (crossing count with the
// on-edge cases returning 0, the cross product in double) - the same restatement as csrc/host_polygon.cpp, which
// tests/test_host_logic.py checks against cv2 itself.
#include <algorithm>
#include "common.cuh"

namespace isg {

constexpr int kPolyThreads = 256;
constexpr int kPolyWarps = kPolyThreads / 32;
constexpr int kPolyMaxPoints = 2048;     // points of one instance held in shared memory (more: the global-memory variant)
constexpr int kPolyRankSort = 512;       // up to this many points: rank sort instead of the bitonic network
constexpr int kPolyMaxCand = 4096;       // keep pixels inside one instance's ghost range listed in shared memory

struct PolySmem {
  float2 pts[kPolyMaxPoints];
  unsigned long long keys[kPolyMaxPoints];
  int warp_tot[kPolyWarps];
  int red_i[kPolyWarps];
  int red_j[kPolyWarps];
  float misc[8];
};

// crossing-count contribution of edge (v0 -> v) for the query point; returns 0 / 1 to add to the counter and sets
// on_edge when the point lies on the edge (pointPolygonTest returns 0 then)
__device__ __forceinline__ int pip_edge(float v0x, float v0y, float vx, float vy, float px, float py, bool& on_edge) {
  if ((v0y <= py && vy <= py) || (v0y > py && vy > py) || (v0x < px && vx < px)) {
    if (py == vy && (px == vx || (py == v0y && ((v0x <= px && px <= vx) || (vx <= px && px <= v0x))))) on_edge = true;
    return 0;
  }
  double dist = (double)(py - v0y) * (double)(vx - v0x) - (double)(px - v0x) * (double)(vy - v0y);
  if (dist == 0) { on_edge = true; return 0; }
  if (vy < v0y) dist = -dist;
  return dist > 0 ? 1 : 0;
}

// pointPolygonTest(pts[0..K), (px,py), False) by one warp: +1 inside, 0 on the polyline, -1 outside
__device__ __forceinline__ int pip_warp(const float2* pts, int K, float px, float py, int lane) {
  int cnt = 0;
  bool on_edge = false;
  for (int i = lane; i < K; i += 32) {
    const float2 v0 = pts[i == 0 ? K - 1 : i - 1], v = pts[i];
    cnt += pip_edge(v0.x, v0.y, v.x, v.y, px, py, on_edge);
  }
  cnt = warp_sum(cnt);
  const bool any_on = __any_sync(0xffffffffu, on_edge);
  if (any_on) return 0;
  return (cnt & 1) ? 1 : -1;
}

// the same by the whole CTA (result broadcast to every thread); uses s.red_i / s.red_j
__device__ __forceinline__ int pip_block(PolySmem& s, const float2* pts, int K, float px, float py) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int cnt = 0;
  bool on_edge = false;
  for (int i = tid; i < K; i += kPolyThreads) {
    const float2 v0 = pts[i == 0 ? K - 1 : i - 1], v = pts[i];
    cnt += pip_edge(v0.x, v0.y, v.x, v.y, px, py, on_edge);
  }
  cnt = warp_sum(cnt);
  const bool any_on = __any_sync(0xffffffffu, on_edge);
  __syncthreads();
  if (lane == 0) { s.red_i[warp] = cnt; s.red_j[warp] = any_on ? 1 : 0; }
  __syncthreads();
  int tot = 0, on = 0;
#pragma unroll
  for (int w = 0; w < kPolyWarps; ++w) { tot += s.red_i[w]; on |= s.red_j[w]; }
  if (on) return 0;
  return (tot & 1) ? 1 : -1;
}

__device__ __forceinline__ float polar_theta(float dx, float dy) {
  const float PI_F = 3.14159274101257324f;          // float32(np.pi)
  if (dx == 0.0f && dy > 0.0f) return 1.57079637050628662f;    // float32(np.pi / 2)
  if (dx == 0.0f && dy < 0.0f) return 4.71238899230957031f;    // float32(3 * np.pi / 2)
  float seta = (float)atan((double)__fdiv_rn(dy, dx));         // fp32 arctan of the fp32 ratio (NaN for 0/0)
  if (dx < 0.0f) seta = __fadd_rn(seta, PI_F);
  else if (dx > 0.0f && dy < 0.0f) seta = __fadd_rn(seta, __fmul_rn(2.0f, PI_F));
  return seta;
}

__device__ __forceinline__ float2 box_centre(const float4 r4, int layout) {
  if (layout == ISG_BOX_XYXY)     // x1,y1,x2,y2: centre = (lt+rb)/2 (:430)
    return make_float2(__fmul_rn(__fadd_rn(r4.x, r4.z), 0.5f), __fmul_rn(__fadd_rn(r4.y, r4.w), 0.5f));
  return make_float2(r4.y, r4.x);   // cy,cx,h,w: group_kp's center_indexes
}

// Steps 3-5 for one instance: internal point, polar angles, stable sort, centre test.  pts[0..K) holds the row-major
// point set and receives the sorted polygon (also written to out); keys needs room for the next power of two >= K.
// LARGE = false: pts / keys are the CTA's shared-memory arrays (K <= kPolyMaxPoints).  LARGE = true: global memory,
// tmp[0..K) is scratch for the permutation (the sequential-mean fall-back stages chunks through s.pts).
// Returns 1 when the polygon is valid.  Must be called by the whole CTA.
#ifdef ISG_POLY_DEBUG
__device__ long long g_poly_dbg[4096 * 8];
#define PSTAMP(k) do { const int cid__ = blockIdx.y * gridDim.x + blockIdx.x; if (threadIdx.x == 0 && cid__ < 4096) g_poly_dbg[cid__ * 8 + (k)] = clock64(); } while (0)
#else
#define PSTAMP(k) do {} while (0)
#endif
template <bool LARGE>
__device__ int finish_polygon(PolySmem& s, float2* pts, unsigned long long* keys, float2* tmp, float2* out, int K, int W, int H,
                              float cx, float cy, float2* internal_out) {
  // ---- internal point (:51-68) ----
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float ix = cx, iy = cy;
  if (pip_block(s, pts, K, cx, cy) <= 0) {
    // numpy mean(axis=0) of a C-contiguous [K,2] fp32 array: sequential fp32 sums over the rows, then / K.  The
    // coordinates are integers, so while K * max(W,H) < 2^24 every partial sum is an exactly representable integer
    // and the sum does not depend on the order: reduce in parallel; otherwise add sequentially like numpy.
    if ((long long)K * max(W, H) < (1ll << 24)) {
      int sxi = 0, syi = 0;
      for (int i = tid; i < K; i += kPolyThreads) { sxi += (int)pts[i].x; syi += (int)pts[i].y; }
      sxi = warp_sum(sxi); syi = warp_sum(syi);
      __syncthreads();
      if (lane == 0) { s.red_i[warp] = sxi; s.red_j[warp] = syi; }
      __syncthreads();
      if (tid == 0) {
        int tx = 0, ty = 0;
        for (int w = 0; w < kPolyWarps; ++w) { tx += s.red_i[w]; ty += s.red_j[w]; }
        s.misc[0] = __fdiv_rn((float)tx, (float)K); s.misc[1] = __fdiv_rn((float)ty, (float)K);
      }
    } else {
      // sequential like numpy; the points are staged through shared memory chunk by chunk so that the one adding
      // thread only pays the dependent adds
      float sx = 0.0f, sy = 0.0f;
      for (int c0 = 0; c0 < K; c0 += kPolyMaxPoints) {
        const int m = min(kPolyMaxPoints, K - c0);
        __syncthreads();
        for (int i = tid; i < m; i += kPolyThreads) s.pts[i] = pts[c0 + i];
        __syncthreads();
        if (tid == 0)
          for (int i = 0; i < m; ++i) { sx = __fadd_rn(sx, s.pts[i].x); sy = __fadd_rn(sy, s.pts[i].y); }
      }
      if (tid == 0) { s.misc[0] = __fdiv_rn(sx, (float)K); s.misc[1] = __fdiv_rn(sy, (float)K); }
    }
    __syncthreads();
    const float mx = s.misc[0], my = s.misc[1];
    if (pip_block(s, pts, K, mx, my) > 0) { ix = mx; iy = my; }
    else {
      // pair midpoints in the reference's order (i outer over [0,K), j inner over [1,K)); one candidate per warp
      const long long ncand = (long long)K * (K - 1);
      long long found = -1;
      for (long long c0 = 0; c0 < ncand && found < 0; c0 += kPolyWarps) {
        const long long c = c0 + warp;
        int hit = 0;
        if (c < ncand) {
          const int i = (int)(c / (K - 1)), j = 1 + (int)(c - (long long)i * (K - 1));
          const float qx = __fdiv_rn(__fadd_rn(pts[i].x, pts[j].x), 2.0f), qy = __fdiv_rn(__fadd_rn(pts[i].y, pts[j].y), 2.0f);
          hit = pip_warp(pts, K, qx, qy, lane) > 0;
        }
        __syncthreads();
        if (lane == 0) s.red_i[warp] = hit;
        __syncthreads();
        for (int w = 0; w < kPolyWarps; ++w) if (s.red_i[w]) { found = c0 + w; break; }
      }
      if (found >= 0) {
        const int i = (int)(found / (K - 1)), j = 1 + (int)(found - (long long)i * (K - 1));
        ix = __fdiv_rn(__fadd_rn(pts[i].x, pts[j].x), 2.0f); iy = __fdiv_rn(__fadd_rn(pts[i].y, pts[j].y), 2.0f);
      }
    }
  }
  if (tid == 0 && internal_out) *internal_out = make_float2(ix, iy);
  PSTAMP(4);

  // ---- polar angles + stable ascending sort (bitonic on (angle bits, index)) ----
  int Kp = 1;
  while (Kp < K) Kp <<= 1;
  for (int i = tid; i < Kp; i += kPolyThreads) {
    unsigned long long key = ~0ull;
    if (i < K) {
      const float2 p = pts[i];
      float th = polar_theta(__fsub_rn(p.x, ix), __fsub_rn(p.y, iy));
      th = th + 0.0f;                                                 // -0 -> +0
      uint32_t tb = __float_as_uint(th);
      if (th != th) tb = 0x7fffffffu;                                 // NaN sorts last (np.argsort)
      else if (tb & 0x80000000u) tb = 0;                              // defensive: angles are never negative
      key = ((unsigned long long)tb << 32) | (unsigned)i;
    }
    keys[i] = key;
  }
  __syncthreads();
  PSTAMP(5);
  if (!LARGE && K <= kPolyRankSort) {
    // small sets: rank sort - every key counts the keys below it (broadcast reads, no barriers); keys are distinct
    // because they carry the index, so the ranks are a permutation
    unsigned long long mykey[kPolyRankSort / kPolyThreads];
    int myrank[kPolyRankSort / kPolyThreads];
#pragma unroll
    for (int u = 0; u < kPolyRankSort / kPolyThreads; ++u) {
      const int i = tid + u * kPolyThreads;
      mykey[u] = (i < K) ? keys[i] : ~0ull;
      myrank[u] = 0;
    }
    if (K <= kPolyThreads) {                    // one key per thread; warps without keys skip the loop
      if ((tid & ~31) < K)
        for (int j = 0; j < K; ++j) myrank[0] += keys[j] < mykey[0];
    } else {
      for (int j = 0; j < K; ++j) {
        const unsigned long long kj = keys[j];
#pragma unroll
        for (int u = 0; u < kPolyRankSort / kPolyThreads; ++u) myrank[u] += kj < mykey[u];
      }
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < kPolyRankSort / kPolyThreads; ++u)
      if (tid + u * kPolyThreads < K) keys[myrank[u]] = mykey[u];
    __syncthreads();
  } else {
    for (int k2 = 2; k2 <= Kp; k2 <<= 1) {
      for (int j2 = k2 >> 1; j2 > 0; j2 >>= 1) {
        for (int i = tid; i < Kp; i += kPolyThreads) {
          const int l = i ^ j2;
          if (l > i) {
            const unsigned long long a = keys[i], c = keys[l];
            const bool up = (i & k2) == 0;
            if ((a > c) == up) { keys[i] = c; keys[l] = a; }
          }
        }
        __syncthreads();
      }
    }
  }
  PSTAMP(6);
  // sorted polygon -> global (and, for the shared-memory variant, a sorted copy in place for the final test)
  if (!LARGE) {
    float2 mine_pt[kPolyMaxPoints / kPolyThreads];
#pragma unroll
    for (int u = 0; u < kPolyMaxPoints / kPolyThreads; ++u) {
      const int i = tid + u * kPolyThreads;
      if (i < K) mine_pt[u] = pts[(int)(keys[i] & 0xffffffffu)];
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < kPolyMaxPoints / kPolyThreads; ++u) {
      const int i = tid + u * kPolyThreads;
      if (i < K) { pts[i] = mine_pt[u]; out[i] = mine_pt[u]; }
    }
  } else {
    for (int i = tid; i < K; i += kPolyThreads) tmp[i] = pts[(int)(keys[i] & 0xffffffffu)];
    __syncthreads();
    for (int i = tid; i < K; i += kPolyThreads) { const float2 v = tmp[i]; pts[i] = v; out[i] = v; }
  }
  __syncthreads();
  // ---- centre strictly inside the sorted polygon (:201) ----
  const int ok__ = pip_block(s, pts, K, cx, cy) > 0 ? 1 : 0;
  PSTAMP(7);
  return ok__;
}


__global__ void __launch_bounds__(kPolyThreads)
instance_polygons_kernel(const uint32_t* __restrict__ keepbits, const int32_t* __restrict__ label_map,
                         const float4* __restrict__ rois, int layout, const float4* __restrict__ ghost,
                         const int32_t* __restrict__ n_seeds, int Nmax, int H, int W, int Wwords, int cap,
                         int obj_pixel_th, float2* __restrict__ poly_points, int32_t* __restrict__ inst_start,
                         int32_t* __restrict__ inst_count, uint8_t* __restrict__ inst_flags,
                         float2* __restrict__ inst_internal, int32_t* __restrict__ img_total,
                         int32_t* __restrict__ stats, unsigned long long* __restrict__ keys_ws,
                         float2* __restrict__ tmp_ws) {
  extern __shared__ __align__(16) unsigned char poly_smem_raw[];
  PolySmem& s = *reinterpret_cast<PolySmem*>(poly_smem_raw);
  // launched with programmatic stream serialisation behind the dense kernel: nothing of its output is read before the
  // whole preceding grid has completed (a no-op for an ordinary launch)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  // image index fastest: the CTAs of the real instances (inst < n_seeds) are contiguous in launch order and spread
  // evenly over the SMs; the empty tail of the instance table comes last
  const int b = blockIdx.x, inst = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  PSTAMP(0);
  const size_t io = (size_t)b * Nmax + inst;
  const int n = min(n_seeds[b], Nmax);
  if (inst >= n) {
    if (tid == 0) { inst_start[io] = 0; inst_count[io] = 0; inst_flags[io] = 0; }
    return;
  }
  // ---- scan range: integer pixels strictly inside the ghost bounds ----
  const float4 g = ghost[io];                                   // x_lo, x_hi, y_lo, y_hi (strict)
  const float fx_lo = fmaxf(floorf(g.x) + 1.0f, 0.0f), fx_hi = fminf(ceilf(g.y) - 1.0f, (float)(W - 1));
  const float fy_lo = fmaxf(floorf(g.z) + 1.0f, 0.0f), fy_hi = fminf(ceilf(g.w) - 1.0f, (float)(H - 1));
  const bool empty_range = !(fx_lo <= fx_hi) || !(fy_lo <= fy_hi);          // also true for NaN bounds
  const int x_lo = empty_range ? 0 : (int)fx_lo, x_hi = empty_range ? -1 : (int)fx_hi;
  const int y_lo = empty_range ? 0 : (int)fy_lo, y_hi = empty_range ? -1 : (int)fy_hi;
  const int w_lo = x_lo >> 5, w_hi = x_hi >> 5;
  const int Wb = empty_range ? 0 : (w_hi - w_lo + 1), R = empty_range ? 0 : (y_hi - y_lo + 1);
  const int NW = Wb * R;
  const uint32_t* kb = keepbits + (size_t)b * H * Wwords;
  const int32_t* lm = label_map + (size_t)b * H * W;

  auto box_word = [&](int w, int& y, int& xb) -> uint32_t {     // keep bits of word w of the box (row-major), clipped
    const int r = w / Wb, c = w - r * Wb;
    const int wx = w_lo + c;
    y = y_lo + r; xb = wx << 5;
    uint32_t m = __ldg(kb + (size_t)y * Wwords + wx);
    if (wx == w_lo) m &= 0xffffffffu << (x_lo & 31);
    if (wx == w_hi) m &= 0xffffffffu >> (31 - (x_hi & 31));
    return m;
  };
  // exclusive prefix of `v` over the CTA in thread order; `total` = sum (every thread calls it)
  auto block_scan = [&](int v, int& total) -> int {
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
    __syncthreads();
    if (lane == 31) s.warp_tot[warp] = incl;
    __syncthreads();
    int base = incl - v;
    total = 0;
#pragma unroll
    for (int w = 0; w < kPolyWarps; ++w) { const int t = s.warp_tot[w]; if (w < warp) base += t; total += t; }
    return base;
  };

  // ---- stage the keep words of the range in shared memory: strided, independent loads (all in flight at once);
  // the blocked passes below would otherwise pay one global round trip per word, one after the other ----
  uint32_t* wcache = reinterpret_cast<uint32_t*>(s.pts);          // [2 * kPolyMaxPoints] words; free until the emit step
  const bool cached = NW <= 2 * kPolyMaxPoints;
  if (cached) {
    for (int w = tid; w < NW; w += 4 * kPolyThreads) {
      uint32_t m[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { int y, xb; m[u] = (w + u * kPolyThreads < NW) ? box_word(w + u * kPolyThreads, y, xb) : 0u; }
#pragma unroll
      for (int u = 0; u < 4; ++u) if (w + u * kPolyThreads < NW) wcache[w + u * kPolyThreads] = m[u];
    }
    __syncthreads();
  }
  auto range_word = [&](int w, int& y, int& xb) -> uint32_t {     // box_word through the staged copy
    if (!cached) return box_word(w, y, xb);
    const int r = w / Wb, c = w - r * Wb;
    y = y_lo + r; xb = (w_lo + c) << 5;
    return wcache[w];
  };
  // ---- candidates: keep pixels inside the range, row-major (blocked word assignment keeps the order) ----
  const int q = (NW + kPolyThreads - 1) / kPolyThreads;
  const int w0 = min(tid * q, NW), w1 = min(w0 + q, NW);
  int ncand_mine = 0;
  if (cached) { for (int w = w0; w < w1; ++w) ncand_mine += __popc(wcache[w]); }
  else { for (int w = w0; w < w1; ++w) { int y, xb; ncand_mine += __popc(box_word(w, y, xb)); } }
  int C = 0;
  const int cbase = block_scan(ncand_mine, C);
  PSTAMP(1);
  uint32_t* cand = reinterpret_cast<uint32_t*>(s.keys);          // [kPolyMaxCand] packed (y << 16 | x), shares the sort buffer
  const bool listed = C <= kPolyMaxCand && H <= 65535 && W <= 65535;
  int K = 0, base = 0;
  if (listed) {
    {
      int pos = cbase;
      for (int w = w0; w < w1; ++w) {
        int y, xb;
        uint32_t m = range_word(w, y, xb);
        while (m) { const int bit = __ffs(m) - 1; m &= m - 1; cand[pos++] = ((uint32_t)y << 16) | (uint32_t)(xb + bit); }
      }
    }
    __syncthreads();
    // label look-ups: one candidate per thread, all in flight at once
    for (int c = tid; c < C; c += kPolyThreads) {
      const uint32_t v = cand[c];
      if (__ldg(lm + (size_t)(v >> 16) * W + (v & 0xffffu)) != inst) cand[c] = 0xffffffffu;
    }
    __syncthreads();
    const int qc = (C + kPolyThreads - 1) / kPolyThreads;
    const int c0 = min(tid * qc, C), c1 = min(c0 + qc, C);
    int mine = 0;
    for (int c = c0; c < c1; ++c) mine += cand[c] != 0xffffffffu;
    base = block_scan(mine, K);
  } else {
    // very many keep pixels in the range: test the labels word by word
    int mine = 0;
    for (int w = w0; w < w1; ++w) {
      int y, xb;
      uint32_t m = box_word(w, y, xb);
      while (m) { const int bit = __ffs(m) - 1; m &= m - 1; mine += __ldg(lm + (size_t)y * W + xb + bit) == inst; }
    }
    base = block_scan(mine, K);
  }

  PSTAMP(2);
  __shared__ int s_start;
  if (tid == 0) {
    s_start = (K > 0) ? atomicAdd(img_total + b, K) : 0;
    inst_count[io] = K;
  }
  __syncthreads();
  const int start = s_start;
  if (tid == 0) inst_start[io] = start;
  const bool fits = K <= kPolyMaxPoints && start + K <= cap;
  float2* out = poly_points + (size_t)b * cap + start;

  // ---- emit the points (x,y) fp32, row-major; bbox for the statistics ----
  int bx0 = 0x7fffffff, by0 = 0x7fffffff, bx1 = -1, by1 = -1;
  auto emit = [&](int pos, int x, int y) {
    bx0 = min(bx0, x); bx1 = max(bx1, x); by0 = min(by0, y); by1 = max(by1, y);
    const float2 p = make_float2((float)x, (float)y);
    if (fits) s.pts[pos] = p;
    else if (start + pos < cap) out[pos] = p;                     // too many points for the device stage: raw set
  };
  if (listed) {
    const int qc = (C + kPolyThreads - 1) / kPolyThreads;
    const int c0 = min(tid * qc, C), c1 = min(c0 + qc, C);
    int pos = base;
    for (int c = c0; c < c1; ++c) {
      const uint32_t v = cand[c];
      if (v != 0xffffffffu) emit(pos++, (int)(v & 0xffffu), (int)(v >> 16));
    }
  } else {
    int pos = base;
    for (int w = w0; w < w1; ++w) {
      int y, xb;
      uint32_t m = box_word(w, y, xb);
      while (m) {
        const int bit = __ffs(m) - 1; m &= m - 1;
        if (__ldg(lm + (size_t)y * W + xb + bit) == inst) emit(pos++, xb + bit, y);
      }
    }
  }
  if (stats) {
    int32_t* st = stats + io * ISG_STAT_WORDS;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      bx0 = min(bx0, __shfl_xor_sync(0xffffffffu, bx0, o)); by0 = min(by0, __shfl_xor_sync(0xffffffffu, by0, o));
      bx1 = max(bx1, __shfl_xor_sync(0xffffffffu, bx1, o)); by1 = max(by1, __shfl_xor_sync(0xffffffffu, by1, o));
    }
    if (lane == 0 && bx1 >= 0) { atomicMin(st + 1, by0); atomicMin(st + 2, bx0); atomicMax(st + 3, by1); atomicMax(st + 4, bx1); }
    if (tid == 0) st[0] = K;
  }
  __syncthreads();
  PSTAMP(3);
  if (!fits) {
    // more points than the shared-memory arrays hold: the raw row-major set is in `out`; finish it in global memory
    // (sort keys / permutation scratch in the caller's workspace), or flag it for the caller when there is none
    if (K < obj_pixel_th || start + K > cap) { if (tid == 0) inst_flags[io] = 0; return; }
    if (!keys_ws) { if (tid == 0) inst_flags[io] = 2; return; }
    const float2 centre = box_centre(rois[io], layout);
    __threadfence_block();
    const int ok = finish_polygon<true>(s, out, keys_ws + ((size_t)b * cap + start) * 2, tmp_ws + (size_t)b * cap + start, out, K,
                                        W, H, centre.x, centre.y, inst_internal ? inst_internal + io : nullptr);
    if (tid == 0) inst_flags[io] = ok ? 1 : 0;
    return;
  }
  if (K < obj_pixel_th || K == 0) {                                                 // :355
    for (int i = tid; i < K; i += kPolyThreads) out[i] = s.pts[i];
    if (tid == 0) inst_flags[io] = 0;
    return;
  }

  const float2 centre = box_centre(rois[io], layout);
  const int ok = finish_polygon<false>(s, s.pts, s.keys, nullptr, out, K, W, H, centre.x, centre.y,
                                       inst_internal ? inst_internal + io : nullptr);
  if (tid == 0) inst_flags[io] = ok ? 1 : 0;
}

__global__ void zero_i32_kernel(int32_t* p, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = 0;
}

}  // namespace isg

using namespace isg;

extern "C" size_t isg_instance_polygons_workspace_bytes(int B, int cap) {
  if (B <= 0 || cap <= 0) return 0;
  // large-instance stage: sort keys [B][2*cap] u64 + permutation scratch [B][cap] float2
  return (size_t)B * cap * (2 * sizeof(unsigned long long) + sizeof(float2)) + 256;
}

extern "C" int isg_instance_polygons(const uint32_t* keepbits, const int32_t* label_map, const float* rois,
                                     int layout, const float* ghost, const int32_t* n_seeds, int B, int Nmax, int H, int W,
                                     int cap, int obj_pixel_th, float* poly_points, int32_t* inst_start,
                                     int32_t* inst_count, uint8_t* inst_flags, float* inst_internal,
                                     int32_t* img_total, int32_t* stats, void* workspace, size_t workspace_bytes,
                                     int totals_zeroed, isg_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!keepbits || !label_map || !rois || !ghost || !n_seeds || !poly_points || !inst_start || !inst_count || !inst_flags ||
      !img_total)
    return ISG_EINVAL;
  if (B <= 0 || Nmax <= 0 || H <= 0 || W <= 0 || cap <= 0 || B > 65535) return ISG_EINVAL;
  if (layout != ISG_BOX_XYXY && layout != ISG_BOX_CYCXHW) return ISG_EINVAL;
  if (((uintptr_t)rois & 15) || ((uintptr_t)ghost & 15) || ((uintptr_t)poly_points & 7)) return ISG_EINVAL;
  if (workspace && (workspace_bytes < isg_instance_polygons_workspace_bytes(B, cap) || ((uintptr_t)workspace & 255)))
    return ISG_EWORKSPACE;
  if (!totals_zeroed) {
    zero_i32_kernel<<<cdiv(B, 256), 256, 0, stream>>>(img_total, B);
    ISG_LAUNCH_CHECK();
  }
  const size_t smem = sizeof(PolySmem);
  ISG_CUDA(cudaFuncSetAttribute(instance_polygons_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(B, Nmax);
  unsigned long long* keys_ws = static_cast<unsigned long long*>(workspace);     // [B][2*cap]
  float2* tmp_ws = workspace ? reinterpret_cast<float2*>(keys_ws + (size_t)B * cap * 2) : nullptr;   // [B][cap]
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = dim3(kPolyThreads); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = totals_zeroed ? 1 : 0;      // directly behind the dense kernel: overlap the launch
  const float4* rois4 = reinterpret_cast<const float4*>(rois);
  const float4* ghost4 = reinterpret_cast<const float4*>(ghost);
  const int Wwords = cdiv(W, 32);
  ISG_CUDA(cudaLaunchKernelEx(&cfg, instance_polygons_kernel, keepbits, label_map, rois4, layout, ghost4, n_seeds, Nmax, H, W, Wwords,
                              cap, obj_pixel_th, reinterpret_cast<float2*>(poly_points), inst_start, inst_count, inst_flags,
                              reinterpret_cast<float2*>(inst_internal), img_total, stats, keys_ws, tmp_ws));
  ISG_LAUNCH_CHECK();
  return ISG_OK;
}

#ifdef ISG_POLY_DEBUG
extern "C" int isg_debug_poly_stamps(long long* host_out) {
  return (int)cudaMemcpyFromSymbol(host_out, isg::g_poly_dbg, sizeof(long long) * 4096 * 8);
}
#endif
