// 3x3 peak test on the thresholded boundary-keypoint map, shared by the stand-alone keep kernel
// and the fused dense kernel.  Reference: select_points / nms_hm, utils/decode.py:42-48,71-85.
#pragma once
#include "common.cuh"

namespace isg {

// monotone signed key: a < b as floats (total order, -0 < +0)  <=>  skey(a) < skey(b) as int32
__device__ __forceinline__ int skey(float x) {
  int s = __float_as_int(x);
  return s ^ ((s >> 31) & 0x7fffffff);
}
__device__ __forceinline__ int skey_from_ukey(uint32_t k) { return (int)(k ^ 0x80000000u); }

// v(p) = selected ? value : 0   (mat * mask, utils/decode.py:84)
__device__ __forceinline__ float selv(float x, int thr_skey) { return skey(x) >= thr_skey ? x : 0.0f; }

// One row of the thresholded map as seen by a thread owning pixels x0..x0+3:
// r[0] = v(x0-1), r[1..4] = v(x0..x0+3), r[5] = v(x0+4); out-of-image -> -inf (max_pool2d padding).
struct Row6 { float r[6]; };

// raw4: the thread's 4 raw pixel values (only meaningful where in-image).
// The lanes of a warp own consecutive 4-pixel groups of ONE row, so the halo comes from the
// neighbouring lanes by shuffle; only lane 0 / lane 31 touch memory for it.
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

template <bool VEC>
__device__ __forceinline__ Row6 load_vrow(const float* __restrict__ img, int y, int x0, int H, int W,
                                          int thr_skey, int lane, float (&raw)[4]) {
  Row6 o;
  const float ninf = __int_as_float(0xff800000);
  if (y < 0 || y >= H) {  // warp-uniform
#pragma unroll
    for (int i = 0; i < 6; ++i) o.r[i] = ninf;
#pragma unroll
    for (int i = 0; i < 4; ++i) raw[i] = 0.0f;
    return o;
  }
  const float* row = img + (int64_t)y * W;
  load_raw4<VEC>(row, x0, W, raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) o.r[i + 1] = (x0 + i < W) ? selv(raw[i], thr_skey) : ninf;
  float left = __shfl_up_sync(0xffffffffu, o.r[4], 1);
  float right = __shfl_down_sync(0xffffffffu, o.r[1], 1);
  if (lane == 0) left = (x0 - 1 >= 0 && x0 - 1 < W) ? selv(__ldg(row + x0 - 1), thr_skey) : ninf;
  if (lane == 31) right = (x0 + 4 < W) ? selv(__ldg(row + x0 + 4), thr_skey) : ninf;
  o.r[0] = left;
  o.r[5] = right;
  return o;
}

// keep nibble of the thread's 4 pixels: bit i = selected(x0+i) && v == max3x3(v)
__device__ __forceinline__ uint32_t keep_nibble(const Row6& up, const Row6& mid, const Row6& down,
                                                const float (&raw)[4], int x0, int W, int thr_skey) {
  uint32_t nib = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float m = fmaxf(fmaxf(up.r[i], up.r[i + 1]), up.r[i + 2]);
    m = fmaxf(m, fmaxf(mid.r[i], mid.r[i + 2]));
    m = fmaxf(m, fmaxf(fmaxf(down.r[i], down.r[i + 1]), down.r[i + 2]));
    const bool sel = (x0 + i < W) && (skey(raw[i]) >= thr_skey);
    if (sel && mid.r[i + 1] >= m) nib |= 1u << i;
  }
  return nib;
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
