// K2 — boundary-keypoint selection: exact k-th largest (3-pass radix select), 3x3 peak test,
// row-major compaction.  Reference: select_points / nms_hm (utils/decode.py:42-48,71-85) and
// kp_mask.nonzero() (utils/decode.py:312).
#include "keep.cuh"

namespace isg {

constexpr int kHistBins = 2048;          // 11-bit digits (11 + 11 + 10 = 32)
constexpr int kHistThreads = 256;
constexpr int kHistPxPerBlock = 16384;   // 16 float4 per thread

__device__ __forceinline__ uint32_t digit_of(uint32_t key, int pass) {
  return pass == 0 ? (key >> 21) : pass == 1 ? ((key >> 10) & 0x7ffu) : (key & 0x3ffu);
}

// Block-wide (256 threads): find the digit d with  #(digit > d) < krem <= #(digit >= d)
// in a 2048-bin histogram; returns d and the rank remaining inside that bin.
__device__ void resolve_digit(const uint32_t* __restrict__ hist, uint32_t krem, uint32_t* out_digit,
                              uint32_t* out_krem) {
  __shared__ uint32_t warp_tot[kHistThreads / 32];
  __shared__ uint32_t res[2];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  uint32_t h[8];
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) { h[i] = hist[t * 8 + i]; s += h[i]; }
  // inclusive suffix sum over threads (thread t gets sum over t' >= t)
  uint32_t suf = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t v = __shfl_down_sync(0xffffffffu, suf, o);
    if (lane + o < 32) suf += v;
  }
  if (lane == 0) warp_tot[warp] = suf;
  if (t == 0) { res[0] = 0; res[1] = 0; }
  __syncthreads();
  uint32_t above = 0;
  for (int w = warp + 1; w < kHistThreads / 32; ++w) above += warp_tot[w];
  uint32_t cum = suf - s + above;  // count of keys in bins owned by higher threads
#pragma unroll
  for (int i = 7; i >= 0; --i) {
    if (cum < krem && krem <= cum + h[i]) { res[0] = (uint32_t)(t * 8 + i); res[1] = krem - cum; }
    cum += h[i];
  }
  __syncthreads();
  *out_digit = res[0];
  *out_krem = res[1];
  __syncthreads();
}

// warp-aggregated shared-memory histogram increment (background pixels fall in a handful of bins;
// a plain atomicAdd would serialise 32 ways on one bank)
__device__ __forceinline__ void hist_add(uint32_t* sh, uint32_t bin, bool valid, int lane) {
#pragma unroll
  for (int round = 0; round < 2; ++round) {
    const unsigned active = __ballot_sync(0xffffffffu, valid);
    if (active == 0) return;
    const int leader = __ffs(active) - 1;
    const uint32_t lb = __shfl_sync(0xffffffffu, bin, leader);
    const unsigned same = __ballot_sync(0xffffffffu, valid && bin == lb);
    if (lane == leader) atomicAdd(&sh[lb], (uint32_t)__popc(same));
    valid = valid && (bin != lb);
  }
  if (valid) atomicAdd(&sh[bin], 1u);
}

// hist layout: [B][3][2048] uint32, zeroed before pass 0.
template <int PASS>
__global__ void __launch_bounds__(kHistThreads)
hist_pass_kernel(const float* __restrict__ kp, int64_t img_stride, int npx, uint32_t k,
                 uint32_t* __restrict__ hist_all, bool vec) {
  __shared__ uint32_t sh[kHistBins];
  const int b = blockIdx.y;
  uint32_t* hist = hist_all + (size_t)b * 3 * kHistBins;
  const int t = threadIdx.x, lane = t & 31;
  for (int i = t; i < kHistBins; i += kHistThreads) sh[i] = 0;

  uint32_t prefix = 0, pmask = 0;
  if (PASS >= 1) {
    uint32_t d0, k1;
    resolve_digit(hist, k, &d0, &k1);
    prefix = d0 << 21; pmask = 0xffe00000u;
    if (PASS >= 2) {
      uint32_t d1, k2;
      resolve_digit(hist + kHistBins, k1, &d1, &k2);
      prefix |= d1 << 10; pmask = 0xfffffc00u;
    }
  }
  __syncthreads();

  const float* img = kp + (int64_t)b * img_stride;
  const int base = blockIdx.x * kHistPxPerBlock;
  const int end = min(base + kHistPxPerBlock, npx);
  if (vec) {
    // warp-uniform trip count (the ballots in hist_add need the whole warp); npx % 4 == 0 here
    for (int p0 = base; p0 < end; p0 += kHistThreads * 4) {
      const int p = p0 + t * 4;
      const bool in = p < end;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (in) v = __ldg(reinterpret_cast<const float4*>(img + p));
      const float a[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t key = float_key(a[i]);
        hist_add(sh, digit_of(key, PASS), in && (key & pmask) == prefix, lane);
      }
    }
  } else {
    // keep the warp converged for the ballots: iterate to a warp-uniform bound
    for (int p0 = base; p0 < end; p0 += kHistThreads) {
      const int p = p0 + t;
      const bool in = p < end;
      const uint32_t key = in ? float_key(__ldg(img + p)) : 0u;
      hist_add(sh, digit_of(key, PASS), in && (key & pmask) == prefix, lane);
    }
  }
  __syncthreads();
  for (int i = t; i < kHistBins; i += kHistThreads) {
    const uint32_t c = sh[i];
    if (c) atomicAdd(&hist[PASS * kHistBins + i], c);
  }
}

__global__ void __launch_bounds__(kHistThreads)
finish_threshold_kernel(const uint32_t* __restrict__ hist_all, uint32_t k, uint32_t* __restrict__ thr_key) {
  const int b = blockIdx.x;
  const uint32_t* hist = hist_all + (size_t)b * 3 * kHistBins;
  uint32_t d0, d1, d2, k1, k2, k3;
  resolve_digit(hist, k, &d0, &k1);
  resolve_digit(hist + kHistBins, k1, &d1, &k2);
  resolve_digit(hist + 2 * kHistBins, k2, &d2, &k3);
  if (threadIdx.x == 0) thr_key[b] = (d0 << 21) | (d1 << 10) | d2;
}

__global__ void fill_u32_kernel(uint32_t* p, int n, uint32_t v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// ---- stand-alone keep kernel ---------------------------------------------------------------
constexpr int kKeepRowsPerWarp = 4;
constexpr int kKeepWarps = 8;

template <bool VEC>
__global__ void __launch_bounds__(32 * kKeepWarps)
keep_kernel(const float* __restrict__ kp, int64_t img_stride, int H, int W, int Wwords,
            const uint32_t* __restrict__ thr_key, uint32_t* __restrict__ keepbits,
            uint8_t* __restrict__ mask_u8) {
  const int b = blockIdx.z;
  const int lane = threadIdx.x, warp = threadIdx.y;
  const int x0 = (blockIdx.x * 32 + lane) * 4;
  const int ybeg = (blockIdx.y * kKeepWarps + warp) * kKeepRowsPerWarp;
  if (ybeg >= H) return;  // warp-uniform
  const float* img = kp + (int64_t)b * img_stride;
  const int thr = skey_from_ukey(thr_key[b]);
  float raw_up[4], raw_mid[4], raw_dn[4];
  Row6 up = load_vrow<VEC>(img, ybeg - 1, x0, H, W, thr, lane, raw_up);
  Row6 mid = load_vrow<VEC>(img, ybeg, x0, H, W, thr, lane, raw_mid);
#pragma unroll
  for (int r = 0; r < kKeepRowsPerWarp; ++r) {
    const int y = ybeg + r;
    if (y >= H) break;
    Row6 dn = load_vrow<VEC>(img, y + 1, x0, H, W, thr, lane, raw_dn);
    const uint32_t nib = keep_nibble(up, mid, dn, raw_mid, x0, W, thr);
    const uint32_t word = nibbles_to_word(nib, lane);
    if ((lane & 7) == 0 && x0 < W) keepbits[((size_t)b * H + y) * Wwords + (x0 >> 5)] = word;
    if (mask_u8) {
      uint8_t* mrow = mask_u8 + ((size_t)b * H + y) * W;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (x0 + i < W) mrow[x0 + i] = (nib >> i) & 1u;
    }
    up = mid; mid = dn;
#pragma unroll
    for (int i = 0; i < 4; ++i) raw_mid[i] = raw_dn[i];
  }
}

// ---- generic k x k heat-map NMS (nms_hm) ----------------------------------------------------
__global__ void nms_hm_kernel(const float* __restrict__ heat, int H, int W, int pad, uint8_t* __restrict__ keep) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= W || y >= H) return;
  const float* img = heat + (size_t)blockIdx.z * H * W;
  const float c = img[(size_t)y * W + x];
  float m = c;
  for (int dy = -pad; dy <= pad; ++dy) {
    const int yy = y + dy;
    if (yy < 0 || yy >= H) continue;
    for (int dx = -pad; dx <= pad; ++dx) {
      const int xx = x + dx;
      if (xx < 0 || xx >= W) continue;
      m = fmaxf(m, __ldg(img + (size_t)yy * W + xx));
    }
  }
  keep[(size_t)blockIdx.z * H * W + (size_t)y * W + x] = (m == c) ? 1 : 0;
}

// ---- ordered compaction: one CTA per image ---------------------------------------------------
constexpr int kCompactThreads = 1024;

__global__ void __launch_bounds__(kCompactThreads)
compact_kernel(const uint32_t* __restrict__ keepbits, int H, int Wwords, int W, int cap,
               int32_t* __restrict__ idx, int32_t* __restrict__ count) {
  __shared__ int warp_tot[kCompactThreads / 32];
  const int b = blockIdx.x;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int nwords = H * Wwords;
  const uint32_t* bits = keepbits + (size_t)b * nwords;
  const int per = (nwords + kCompactThreads - 1) / kCompactThreads;
  const int w0 = min(t * per, nwords), w1 = min(w0 + per, nwords);
  int c = 0;
  for (int w = w0; w < w1; ++w) c += __popc(bits[w]);
  // block exclusive scan
  int inc = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int v = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += v;
  }
  if (lane == 31) warp_tot[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int v = warp_tot[lane];
    int s = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int u = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += u;
    }
    warp_tot[lane] = s - v;  // exclusive
  }
  __syncthreads();
  int off = warp_tot[warp] + inc - c;
  if (t == kCompactThreads - 1) count[b] = off + c;
  int32_t* out = idx + (size_t)b * cap * 2;
  for (int w = w0; w < w1; ++w) {
    uint32_t m = bits[w];
    const int y = w / Wwords, xb = (w - y * Wwords) * 32;
    while (m) {
      const int i = __ffs(m) - 1;
      m &= m - 1;
      if (off < cap) { out[2 * off] = y; out[2 * off + 1] = xb + i; }
      ++off;
    }
  }
}

}  // namespace isg

using namespace isg;

extern "C" size_t isg_topk_workspace_bytes(int B) {
  return B > 0 ? (size_t)B * 3 * kHistBins * sizeof(uint32_t) : 0;
}

extern "C" int isg_topk_threshold(const float* kp, int B, int H, int W, int64_t img_stride, int k,
                                  uint32_t* thr_key, void* ws, size_t ws_bytes, isg_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!kp || !thr_key || B <= 0 || H <= 0 || W <= 0 || k < 0) return ISG_EINVAL;
  const int64_t npx64 = (int64_t)H * W;
  if (npx64 > (int64_t)1 << 30 || img_stride < npx64) return ISG_EINVAL;
  if ((int64_t)k > npx64) return ISG_EINVAL;  // torch.topk raises (utils/decode.py:81)
  const int npx = (int)npx64;
  if (k == 0) {
    fill_u32_kernel<<<cdiv(B, 128), 128, 0, stream>>>(thr_key, B, 0xffffffffu);
    ISG_LAUNCH_CHECK();
    return ISG_OK;
  }
  if (!ws || ws_bytes < isg_topk_workspace_bytes(B) || ((uintptr_t)ws & 15)) return ISG_EWORKSPACE;
  uint32_t* hist = (uint32_t*)ws;
  ISG_CUDA(cudaMemsetAsync(hist, 0, isg_topk_workspace_bytes(B), stream));
  const bool vec = (npx % 4 == 0) && (img_stride % 4 == 0) && (((uintptr_t)kp & 15) == 0);
  dim3 grid(cdiv(npx, kHistPxPerBlock), B);
  hist_pass_kernel<0><<<grid, kHistThreads, 0, stream>>>(kp, img_stride, npx, (uint32_t)k, hist, vec);
  hist_pass_kernel<1><<<grid, kHistThreads, 0, stream>>>(kp, img_stride, npx, (uint32_t)k, hist, vec);
  hist_pass_kernel<2><<<grid, kHistThreads, 0, stream>>>(kp, img_stride, npx, (uint32_t)k, hist, vec);
  finish_threshold_kernel<<<B, kHistThreads, 0, stream>>>(hist, (uint32_t)k, thr_key);
  ISG_LAUNCH_CHECK();
  return ISG_OK;
}

extern "C" int isg_keep_points(const float* kp, int B, int H, int W, int64_t img_stride,
                               const uint32_t* thr_key, uint32_t* keepbits, uint8_t* mask_u8,
                               isg_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!kp || !thr_key || !keepbits || B <= 0 || H <= 0 || W <= 0) return ISG_EINVAL;
  if (img_stride < (int64_t)H * W || B > 65535) return ISG_EINVAL;
  const int Wwords = cdiv(W, 32);
  const bool vec = (W % 4 == 0) && (img_stride % 4 == 0) && (((uintptr_t)kp & 15) == 0);
  dim3 block(32, kKeepWarps);
  dim3 grid(cdiv(W, 128), cdiv(H, kKeepWarps * kKeepRowsPerWarp), B);
  if (vec) keep_kernel<true><<<grid, block, 0, stream>>>(kp, img_stride, H, W, Wwords, thr_key, keepbits, mask_u8);
  else keep_kernel<false><<<grid, block, 0, stream>>>(kp, img_stride, H, W, Wwords, thr_key, keepbits, mask_u8);
  ISG_LAUNCH_CHECK();
  return ISG_OK;
}

extern "C" size_t isg_select_points_workspace_bytes(int B) {
  return B > 0 ? isg_topk_workspace_bytes(B) + (((size_t)B * 4 + 15) & ~(size_t)15) : 0;
}

extern "C" int isg_select_points(const float* kp, int B, int H, int W, int64_t img_stride, int k,
                                 uint32_t* keepbits, uint8_t* mask_u8, void* ws, size_t ws_bytes,
                                 isg_stream_t stream) {
  if (B <= 0) return ISG_EINVAL;
  if (!ws || ws_bytes < isg_select_points_workspace_bytes(B) || ((uintptr_t)ws & 15)) return ISG_EWORKSPACE;
  uint32_t* thr = (uint32_t*)((char*)ws + isg_topk_workspace_bytes(B));
  int rc = isg_topk_threshold(kp, B, H, W, img_stride, k, thr, ws, isg_topk_workspace_bytes(B), stream);
  if (rc) return rc;
  return isg_keep_points(kp, B, H, W, img_stride, thr, keepbits, mask_u8, stream);
}

extern "C" int isg_nms_hm(const float* heat, int planes, int H, int W, int kernel, uint8_t* keep,
                          isg_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!heat || !keep || planes <= 0 || H <= 0 || W <= 0) return ISG_EINVAL;
  if (kernel < 1 || kernel > 15 || (kernel & 1) == 0) return ISG_EUNSUPPORTED;
  if (planes > 65535) return ISG_EUNSUPPORTED;
  dim3 block(32, 8), grid(cdiv(W, 32), cdiv(H, 8), planes);
  nms_hm_kernel<<<grid, block, 0, stream>>>(heat, H, W, (kernel - 1) / 2, keep);
  ISG_LAUNCH_CHECK();
  return ISG_OK;
}

extern "C" int isg_compact_points(const uint32_t* keepbits, int B, int H, int W, int cap, int32_t* idx,
                                  int32_t* count, isg_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!keepbits || !idx || !count || B <= 0 || H <= 0 || W <= 0 || cap < 0) return ISG_EINVAL;
  compact_kernel<<<B, kCompactThreads, 0, stream>>>(keepbits, H, cdiv(W, 32), W, cap, idx, count);
  ISG_LAUNCH_CHECK();
  return ISG_OK;
}
