// 3x3 peak test on the thresholded boundary-keypoint map, shared by the stand-alone keep kernel
// and the fused dense kernel.  Reference: select_points / nms_hm, utils/decode.py:42-48,71-85.
//
//   v(p)    = selected(p) ? kp[p] : 0            (mat * mask, :84)
//   keep(p) = selected(p) && v(p) == max3x3(v)   (max_pool2d with -inf padding, :45-47,85)
// The 3x3 maximum is evaluated separably: h(p) = max(v(x-1), v(x), v(x+1)) per row, then the maximum of
// h over the three rows; the centre is part of its own window, so keep <=> v(p) >= that maximum.
#pragma once
#include "common.cuh"

namespace isg {

// monotone signed key: a < b as floats (total order, -0 < +0)  <=>  skey(a) < skey(b) as int32
__device__ __forceinline__ int skey(float x) {
  int s = __float_as_int(x);
  return s ^ ((s >> 31) & 0x7fffffff);
}
__device__ __forceinline__ int skey_from_ukey(uint32_t k) { return (int)(k ^ 0x80000000u); }
__device__ __forceinline__ float float_from_ukey(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// Selection threshold.  Selection is defined on the total order of the keys (that is what the radix select
// counted); for every threshold except +-0 that is exactly the float comparison x >= thr, which is cheaper.
struct Thr {
  float f;
  int s;
  bool use_int;   // threshold is +-0: -0 and +0 compare equal as floats but are distinct keys
};
__device__ __forceinline__ Thr make_thr(uint32_t ukey) {
  Thr t;
  t.f = float_from_ukey(ukey);
  t.s = skey_from_ukey(ukey);
  t.use_int = (t.f == 0.0f) || (ukey == 0xffffffffu) || (t.f != t.f);   // +-0, "select nothing", NaN keys
  return t;
}
template <bool USE_INT>
__device__ __forceinline__ bool selected(float x, const Thr& t) { return USE_INT ? (skey(x) >= t.s) : (x >= t.f); }

// One image row as seen by a lane owning pixels x0..x0+3.
struct RowH {
  float h[4];      // horizontal 3-max of v centred on the lane's 4 pixels
  float v[4];      // v of the lane's 4 pixels (-inf outside the image)
  uint32_t sel;    // bit i: pixel x0+i is inside the image and selected
};

__device__ __forceinline__ RowH rowh_outside() {
  RowH o;
  const float ninf = __int_as_float(0xff800000);
#pragma unroll
  for (int i = 0; i < 4; ++i) { o.h[i] = ninf; o.v[i] = ninf; }
  o.sel = 0;
  return o;
}

// Build a RowH from the lane's 4 raw pixels.  `nin` = number of the lane's pixels inside the image (0..4),
// `left_raw`/`right_raw` are only read by lane 0 / lane 31 (halo pixels the neighbouring lanes do not own) and
// `has_left`/`has_right` say whether those halo pixels exist.  The interior halo comes from the neighbouring
// lanes by shuffle (the lanes of a warp own consecutive 4-pixel groups of ONE row).
template <bool USE_INT>
__device__ __forceinline__ RowH make_rowh(const float (&raw)[4], int nin, float left_raw, bool has_left, float right_raw,
                                          bool has_right, const Thr& thr, int lane) {
  RowH o;
  const float ninf = __int_as_float(0xff800000);
  o.sel = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const bool s = selected<USE_INT>(raw[i], thr);
    o.v[i] = s ? raw[i] : 0.0f;
    o.sel |= (s ? 1u : 0u) << i;
  }
  if (nin < 4) {   // only the lanes straddling / beyond the right image border
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (i >= nin) { o.v[i] = ninf; o.sel &= ~(1u << i); }
  }
  float left = __shfl_up_sync(0xffffffffu, o.v[3], 1);
  float right = __shfl_down_sync(0xffffffffu, o.v[0], 1);
  if (lane == 0) left = has_left ? (selected<USE_INT>(left_raw, thr) ? left_raw : 0.0f) : ninf;
  if (lane == 31) right = has_right ? (selected<USE_INT>(right_raw, thr) ? right_raw : 0.0f) : ninf;
  o.h[0] = fmaxf(fmaxf(left, o.v[0]), o.v[1]);
  o.h[1] = fmaxf(fmaxf(o.v[0], o.v[1]), o.v[2]);
  o.h[2] = fmaxf(fmaxf(o.v[1], o.v[2]), o.v[3]);
  o.h[3] = fmaxf(fmaxf(o.v[2], o.v[3]), right);
  return o;
}

template <bool VEC>
__device__ __forceinline__ void load_raw4(const float* __restrict__ row, int x0, int W, float (&v)[4]) {
  if (VEC && x0 + 3 < W) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(row + x0));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = (x0 + i < W) ? __ldg(row + x0 + i) : 0.0f;
  }
}

// row `y` of a global-memory image
template <bool VEC, bool USE_INT>
__device__ __forceinline__ RowH load_rowh(const float* __restrict__ img, int y, int x0, int H, int W, const Thr& thr,
                                          int lane) {
  if (y < 0 || y >= H) return rowh_outside();   // warp-uniform
  const float* row = img + (int64_t)y * W;
  float raw[4];
  load_raw4<VEC>(row, x0, W, raw);
  const bool has_left = x0 - 1 >= 0 && x0 - 1 < W, has_right = x0 + 4 < W;
  float lr = 0.0f, rr = 0.0f;
  if (lane == 0 && has_left) lr = __ldg(row + x0 - 1);
  if (lane == 31 && has_right) rr = __ldg(row + x0 + 4);
  return make_rowh<USE_INT>(raw, min(max(W - x0, 0), 4), lr, has_left, rr, has_right, thr, lane);
}

// keep nibble of the lane's 4 pixels of the middle row
__device__ __forceinline__ uint32_t keep_nibble(const RowH& up, const RowH& mid, const RowH& dn) {
  uint32_t nib = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float m = fmaxf(fmaxf(up.h[i], mid.h[i]), dn.h[i]);
    if (mid.v[i] >= m) nib |= 1u << i;
  }
  return nib & mid.sel;
}

// Combine the nibbles of 8 consecutive lanes into one 32-bit word (valid in lanes with lane%8==0).
__device__ __forceinline__ uint32_t nibbles_to_word(uint32_t nib, int lane) {
  uint32_t w = nib << (4 * (lane & 7));
  w |= __shfl_xor_sync(0xffffffffu, w, 1);
  w |= __shfl_xor_sync(0xffffffffu, w, 2);
  w |= __shfl_xor_sync(0xffffffffu, w, 4);
  return w;
}

}  // namespace isg
