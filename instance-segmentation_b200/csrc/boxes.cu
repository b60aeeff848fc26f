// Box head front-end, K5 greedy box NMS (bitmask matrix + chunked suppression scan) and K6
// bit-packed mask-IoU NMS.
// Reference: decode_boxes (utils/decode.py:377-419), BBoxTransform / ClipBoxes
// (utils/utils.py:318-363), py_cpu_nms (utils/nms.py:11-39), torchvision batched_nms semantics at
// utils/decode.py:400, mask IoU formula (utils/image.py:188-191).
#include <algorithm>
#include <cuda_fp16.h>
#include "common.cuh"

namespace isg {

// ---------------------------------------------------------------------------------------------
// front-end: score = max_c cls, class = argmax_c (first index), threshold, box transform, clip
// ---------------------------------------------------------------------------------------------
constexpr int kFrontThreads = 256;

constexpr int kFrontPer = 2;          // anchors per thread (their score rows are loaded before either is examined)

__global__ void __launch_bounds__(kFrontThreads)
decode_boxes_kernel(const float* __restrict__ anchors, const float* __restrict__ regression,
                    const float* __restrict__ classification, int A, int C, float xmax_clip, float ymax_clip,
                    float thr, int cap, float4* __restrict__ cand_boxes, float* __restrict__ cand_scores,
                    int32_t* __restrict__ cand_cls, int32_t* __restrict__ cand_anchor,
                    int32_t* __restrict__ cand_count, bool vec) {
  pdl_trigger();      // the NMS kernel behind this one may be scheduled early; it waits for this grid (common.cuh)
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31;
  float best[kFrontPer];
  int cls[kFrontPer];
  if (vec && C == 8) {
    // the common head (8 classes): both 128-bit halves of every anchor's score row in flight before any compare
    float4 v[kFrontPer][2];
#pragma unroll
    for (int u = 0; u < kFrontPer; ++u) {
      const int a = (blockIdx.x * kFrontPer + u) * kFrontThreads + threadIdx.x;
      const float* p = classification + ((size_t)b * A + min(a, A - 1)) * 8;
      v[u][0] = ldg_stream4(p); v[u][1] = ldg_stream4(p + 4);
    }
#pragma unroll
    for (int u = 0; u < kFrontPer; ++u) {
      best[u] = -INFINITY; cls[u] = 0;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float4 q = v[u][h];
        if (q.x > best[u]) { best[u] = q.x; cls[u] = 4 * h; }
        if (q.y > best[u]) { best[u] = q.y; cls[u] = 4 * h + 1; }
        if (q.z > best[u]) { best[u] = q.z; cls[u] = 4 * h + 2; }
        if (q.w > best[u]) { best[u] = q.w; cls[u] = 4 * h + 3; }
      }
    }
  } else {
#pragma unroll
    for (int u = 0; u < kFrontPer; ++u) {
      const int a = (blockIdx.x * kFrontPer + u) * kFrontThreads + threadIdx.x;
      best[u] = -INFINITY; cls[u] = 0;
      if (a < A) {
        const float* p = classification + ((size_t)b * A + a) * C;
        if (vec) {
          for (int c = 0; c < C; c += 4) {
            const float4 q = ldg_stream4(p + c);
            if (q.x > best[u]) { best[u] = q.x; cls[u] = c; }
            if (q.y > best[u]) { best[u] = q.y; cls[u] = c + 1; }
            if (q.z > best[u]) { best[u] = q.z; cls[u] = c + 2; }
            if (q.w > best[u]) { best[u] = q.w; cls[u] = c + 3; }
          }
        } else {
          for (int c = 0; c < C; ++c) {
            const float q = ldg_stream1(p + c);
            if (q > best[u]) { best[u] = q; cls[u] = c; }
          }
        }
      }
    }
  }
#pragma unroll
  for (int u = 0; u < kFrontPer; ++u) {
    const int a = (blockIdx.x * kFrontPer + u) * kFrontThreads + threadIdx.x;
    const bool hit = (a < A) && (best[u] > thr);                  // scores > threshold (utils/decode.py:384)
    const unsigned bal = __ballot_sync(0xffffffffu, hit);
    if (bal == 0) continue;
    int base = 0;
    if (lane == __ffs(bal) - 1) base = atomicAdd(cand_count + b, __popc(bal));
    base = __shfl_sync(0xffffffffu, base, __ffs(bal) - 1);
    if (!hit) continue;
    const int pos = base + __popc(bal & ((1u << lane) - 1u));
    if (pos >= cap) continue;
    // BBoxTransform (utils/utils.py:331-346); anchors (y1,x1,y2,x2), regression (dy,dx,dh,dw)
    const float4 an = __ldg(reinterpret_cast<const float4*>(anchors) + a);
    const float4 rg = __ldg(reinterpret_cast<const float4*>(regression) + (size_t)b * A + a);
    const float yca = __fmul_rn(__fadd_rn(an.x, an.z), 0.5f), xca = __fmul_rn(__fadd_rn(an.y, an.w), 0.5f);
    const float ha = __fsub_rn(an.z, an.x), wa = __fsub_rn(an.w, an.y);
    const float w = __fmul_rn(expf(rg.w), wa), h = __fmul_rn(expf(rg.z), ha);
    const float yc = __fadd_rn(__fmul_rn(rg.x, ha), yca), xc = __fadd_rn(__fmul_rn(rg.y, wa), xca);
    float ymin = __fsub_rn(yc, __fmul_rn(h, 0.5f)), xmin = __fsub_rn(xc, __fmul_rn(w, 0.5f));
    float ymax = __fadd_rn(yc, __fmul_rn(h, 0.5f)), xmax = __fadd_rn(xc, __fmul_rn(w, 0.5f));
    // ClipBoxes (utils/utils.py:357-361)
    xmin = fmaxf(xmin, 0.0f); ymin = fmaxf(ymin, 0.0f);
    xmax = fminf(xmax, xmax_clip); ymax = fminf(ymax, ymax_clip);
    const size_t o = (size_t)b * cap + pos;
    cand_boxes[o] = make_float4(xmin, ymin, xmax, ymax);
    cand_scores[o] = best[u];
    cand_cls[o] = cls[u];
    cand_anchor[o] = a;
  }
}

// ---------------------------------------------------------------------------------------------
// NMS step 1: order candidates by (score desc, tiebreak) — single-CTA bitonic sort in shared memory
// ---------------------------------------------------------------------------------------------
constexpr int kSortThreads = 1024;

struct NmsWs {           // per-image views into the workspace
  int32_t* order;        // [cap]   candidate index of rank r
  float4* sbox;          // [cap]   boxes in rank order
  int32_t* scls;         // [cap]
  unsigned long long* mask;  // [cap][nw]
};

__host__ __device__ inline size_t nms_ws_per_image(int cap) {
  const size_t nw = (size_t)(cap + 63) / 64;
  size_t s = 0;
  s += ((size_t)cap * 4 + 15) & ~(size_t)15;   // order
  s += (size_t)cap * 16;                       // sbox
  s += ((size_t)cap * 4 + 15) & ~(size_t)15;   // scls
  s += (size_t)cap * nw * 8;                   // mask
  return (s + 255) & ~(size_t)255;
}
__host__ __device__ inline NmsWs nms_ws_view(void* ws, int b, int cap) {
  char* p = (char*)ws + (size_t)b * nms_ws_per_image(cap);
  NmsWs v;
  v.order = (int32_t*)p; p += ((size_t)cap * 4 + 15) & ~(size_t)15;
  v.sbox = (float4*)p;   p += (size_t)cap * 16;
  v.scls = (int32_t*)p;  p += ((size_t)cap * 4 + 15) & ~(size_t)15;
  v.mask = (unsigned long long*)p;
  return v;
}

// torchvision's coordinate trick (_batched_nms_coordinate_trick): offset = float(class) * (max coordinate + 1), added
// to all four coordinates in fp32.  Used for ISG_NMS_TV_TRICK always and for ISG_NMS_TV_BATCHED while the candidate
// set has at most 1000 boxes (boxes.numel() <= 4000, batched_nms's CPU dispatch rule).
__device__ __forceinline__ bool nms_uses_trick(int convention, int n) {
  return convention == ISG_NMS_TV_TRICK || (convention == ISG_NMS_TV_BATCHED && n <= 1000);
}
__device__ __forceinline__ float4 nms_shift_box(float4 b, int cls, float max_plus_1) {
  const float off = __fmul_rn((float)cls, max_plus_1);
  return make_float4(__fadd_rn(b.x, off), __fadd_rn(b.y, off), __fadd_rn(b.z, off), __fadd_rn(b.w, off));
}
// largest coordinate of the image's n candidate boxes (boxes.max()), + 1; every thread of the CTA gets the result
__device__ float nms_block_max_plus_1(const float4* __restrict__ boxes_b, int n, float* red /* [32] shared */) {
  float m = -INFINITY;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float4 q = boxes_b[i];
    m = fmaxf(m, fmaxf(fmaxf(q.x, q.y), fmaxf(q.z, q.w)));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  float r = -INFINITY;
  for (int w = 0; w < (int)((blockDim.x + 31) >> 5); ++w) r = fmaxf(r, red[w]);
  return __fadd_rn(r, 1.0f);
}

__global__ void __launch_bounds__(kSortThreads)
nms_sort_kernel(const float4* __restrict__ boxes, const float* __restrict__ scores, const int32_t* __restrict__ cls,
                const int32_t* __restrict__ tiebreak, const int32_t* __restrict__ count, int cap, int P,
                int convention, void* ws) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned long long* key = reinterpret_cast<unsigned long long*>(smem_raw);   // [P]
  uint32_t* val = reinterpret_cast<uint32_t*>(key + P);                        // [P]
  pdl_trigger();      // isg_mask_nms runs its mask-area pass (independent of the order) alongside this single-CTA sort
  const int b = blockIdx.x, t = threadIdx.x;
  const int n = min(max(count[b], 0), cap);
  int Pn = 1;
  while (Pn < n) Pn <<= 1;   // <= P
  for (int i = t; i < Pn; i += kSortThreads) {
    unsigned long long k = 0ull;
    if (i < n) {
      const size_t o = (size_t)b * cap + i;
      const uint32_t tb = tiebreak ? (uint32_t)tiebreak[o] : (uint32_t)i;
      // descending sort on the 64-bit key: PLUS1_LE visits the larger tiebreak first (argsort()[::-1]
      // of utils/nms.py:20 on small inputs), TV_GT the smaller (stable descending sort)
      const uint32_t low = (convention == ISG_NMS_PLUS1_LE) ? tb : (0xffffffffu - tb);
      k = ((unsigned long long)float_key(scores[o]) << 32) | low;
    }
    key[i] = k; val[i] = (uint32_t)i;
  }
  __syncthreads();
  for (int size = 2; size <= Pn; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = t; i < (Pn >> 1); i += kSortThreads) {
        const int lo = 2 * i - (i & (stride - 1));
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);   // first half of each bitonic block descending
        const unsigned long long a = key[lo], c = key[hi];
        if ((a < c) == desc) {
          key[lo] = c; key[hi] = a;
          const uint32_t va = val[lo]; val[lo] = val[hi]; val[hi] = va;
        }
      }
      __syncthreads();
    }
  }
  NmsWs v = nms_ws_view(ws, b, cap);
  __shared__ float red[32];
  const bool trick = cls && nms_uses_trick(convention, n);                 // block-uniform
  const float mp1 = trick ? nms_block_max_plus_1(boxes + (size_t)b * cap, n, red) : 0.0f;
  for (int r = t; r < n; r += kSortThreads) {
    const int i = (int)val[r];
    const size_t o = (size_t)b * cap + i;
    const int c = cls ? cls[o] : 0;
    v.order[r] = i;
    v.sbox[r] = trick ? nms_shift_box(boxes[o], c, mp1) : boxes[o];        // sbox only feeds the IoU tests
    v.scls[r] = c;
  }
}

// ---------------------------------------------------------------------------------------------
// NMS step 2: suppression bitmask, 64 x 64 boxes per CTA, upper triangle only
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool suppresses_tv(const float4 a, const float4 c, double thr) {
  // torchvision nms_kernel_impl: no +1, suppress iff ovr > thr (ovr fp32 promoted to double)
  const float areaa = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
  const float areac = __fmul_rn(__fsub_rn(c.z, c.x), __fsub_rn(c.w, c.y));
  const float xx1 = fmaxf(a.x, c.x), yy1 = fmaxf(a.y, c.y), xx2 = fminf(a.z, c.z), yy2 = fminf(a.w, c.w);
  const float w = fmaxf(0.0f, __fsub_rn(xx2, xx1)), h = fmaxf(0.0f, __fsub_rn(yy2, yy1));
  const float inter = __fmul_rn(w, h);
  const float uni = __fsub_rn(__fadd_rn(areaa, areac), inter);
  // disjoint boxes (the common case): 0 / uni is exactly +0 for uni > 0; skipping the division avoids the slow
  // special-operand path of the IEEE divide, which a zero numerator always takes
  const float ovr = (inter == 0.0f && uni > 0.0f) ? 0.0f : __fdiv_rn(inter, uni);
  return (double)ovr > thr;
}
__device__ __forceinline__ bool suppresses_plus1(const float4 a, const float4 c, float thr) {
  // utils/nms.py:19,26-36: +1 convention, survivor iff ovr <= thr (NaN is suppressed)
  const float areaa = __fmul_rn(__fadd_rn(__fsub_rn(a.z, a.x), 1.0f), __fadd_rn(__fsub_rn(a.w, a.y), 1.0f));
  const float areac = __fmul_rn(__fadd_rn(__fsub_rn(c.z, c.x), 1.0f), __fadd_rn(__fsub_rn(c.w, c.y), 1.0f));
  const float xx1 = fmaxf(a.x, c.x), yy1 = fmaxf(a.y, c.y), xx2 = fminf(a.z, c.z), yy2 = fminf(a.w, c.w);
  const float w = fmaxf(0.0f, __fadd_rn(__fsub_rn(xx2, xx1), 1.0f));
  const float h = fmaxf(0.0f, __fadd_rn(__fsub_rn(yy2, yy1), 1.0f));
  const float inter = __fmul_rn(w, h);
  const float uni = __fsub_rn(__fadd_rn(areaa, areac), inter);
  const float ovr = (inter == 0.0f && uni > 0.0f) ? 0.0f : __fdiv_rn(inter, uni);   // see suppresses_tv
  return !(ovr <= thr);
}

constexpr int kMaskCtasPerImage = 128;

__global__ void __launch_bounds__(64)
nms_mask_kernel(const int32_t* __restrict__ count, int cap, double thr, int convention, void* ws) {
  const int b = blockIdx.y;
  const int n = min(max(count[b], 0), cap);
  const int nb = (n + 63) / 64;
  const int npairs = nb * (nb + 1) / 2;
  const NmsWs v = nms_ws_view(ws, b, cap);
  const int nw = (cap + 63) / 64;
  __shared__ float4 cbox[64];
  __shared__ int ccls[64];
  const int t = threadIdx.x;
  const float thr_f = (float)thr;
  for (int p = blockIdx.x; p < npairs; p += gridDim.x) {
    // linear index over the upper triangle -> (row block, column block)
    int rb = 0, rem = p;
    while (rem >= nb - rb) { rem -= nb - rb; ++rb; }
    const int cb = rb + rem;
    __syncthreads();
    const int j = cb * 64 + t;
    if (j < n) { cbox[t] = v.sbox[j]; ccls[t] = v.scls[j]; }
    __syncthreads();
    const int i = rb * 64 + t;
    if (i >= n) continue;
    const float4 a = v.sbox[i];
    const int ac = v.scls[i];
    const int ncol = min(64, n - cb * 64);
    unsigned long long bits = 0ull;
    for (int c = (rb == cb ? t + 1 : 0); c < ncol; ++c) {
      if (ccls[c] != ac) continue;   // class aware (batched_nms); class-agnostic callers pass cls = NULL -> all 0
      const bool s = (convention != ISG_NMS_PLUS1_LE) ? suppresses_tv(a, cbox[c], thr) : suppresses_plus1(a, cbox[c], thr_f);
      if (s) bits |= 1ull << c;
    }
    v.mask[(size_t)i * nw + cb] = bits;
  }
}

// ---------------------------------------------------------------------------------------------
// Parallel suppression scan (shared by the fused small kernel and the staged scan).
// The greedy keep set is the unique solution of  keep[i] = !OR_{j<i} (keep[j] & S[j][i])  (S = the strictly upper
// triangular suppression matrix in rank order).  One ROUND computes  K <- ~OR_{j in K} row[j]  from K = all ones: a
// fully parallel OR-reduction over the rows that are currently kept.  After r rounds the first r ranks are final
// (induction over the rank), so a round that changes nothing has reached the greedy result; the number of rounds is
// the depth of the longest suppression chain + 2 - a handful for detections - instead of one dependent step per
// candidate.  A pathological chain (every box suppressing only its successor) would need ~n/2 rounds: after
// `max_rounds` the function gives up and the caller runs the sequential scan.
// rows: [n][rstride] 64-bit words in shared memory, word w of row j valid for w >= j/64.  K: [nw] keep words (result),
// R32: [2*nw] scratch.  Lane layout: a lane owns ONE word column (conflict-free 8-byte shared loads) and a share of
// the rows, ORs its kept rows in a register and meets the other lanes once per round.  Called by the whole CTA.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long nms_valid_word(int w, int n) {
  const int left = n - w * 64;
  return left >= 64 ? ~0ull : (left <= 0 ? 0ull : ((1ull << left) - 1ull));
}
__device__ bool nms_parallel_scan(const unsigned long long* rows, int rstride, int n, int nw, unsigned long long* K,
                                  uint32_t* R32, int max_rounds, const unsigned long long* alive = nullptr) {
  const int t = threadIdx.x, T = blockDim.x, lane = t & 31, warp = t >> 5, nwarps = T >> 5;
  int G = 1;
  while (G < nw && G < 32) G <<= 1;             // lanes per row (word columns handled at once)
  const int rpw = 32 / G;                        // rows per warp and sweep
  const int wl = lane & (G - 1), rsub = lane / G;
  // `alive` (nullable, [nw] words): candidates already suppressed from outside never count (tiled large-set path)
  for (int w = t; w < nw; w += T) K[w] = nms_valid_word(w, n) & (alive ? alive[w] : ~0ull);
  for (int w = t; w < 2 * nw; w += T) R32[w] = 0u;
  __syncthreads();
  for (int round = 0; round < max_rounds; ++round) {
    for (int wb = 0; wb < nw; wb += 32) {        // word columns wb .. wb+31 (one pass unless nw > 32)
      const int w = wb + wl;
      unsigned long long acc = 0ull;
      for (int j = warp * rpw + rsub; j < n; j += nwarps * rpw) {
        const int wj = j >> 6;
        if (w >= wj && w < nw && ((K[wj] >> (j & 63)) & 1ull)) acc |= rows[(size_t)j * rstride + w];
      }
      for (int o = G; o < 32; o <<= 1) acc |= __shfl_xor_sync(0xffffffffu, acc, o);
      if (rsub == 0 && w < nw) {
        const uint32_t lo = (uint32_t)acc, hi = (uint32_t)(acc >> 32);
        if (lo) atomicOr(&R32[2 * w], lo);
        if (hi) atomicOr(&R32[2 * w + 1], hi);
      }
    }
    __syncthreads();
    int changed = 0;
    for (int w = t; w < nw; w += T) {
      const unsigned long long nk = ~((unsigned long long)R32[2 * w] | ((unsigned long long)R32[2 * w + 1] << 32)) & nms_valid_word(w, n) &
                                    (alive ? alive[w] : ~0ull);
      changed |= nk != K[w];
      K[w] = nk; R32[2 * w] = 0u; R32[2 * w + 1] = 0u;
    }
    if (!__syncthreads_or(changed)) return true;
  }
  return false;
}
// keep list of a converged scan: the kept ranks in rank order.  pre: [nw + 1] scratch.  Called by the whole CTA.
template <typename IdOf>
__device__ void nms_emit_kept(const unsigned long long* K, int n, int nw, int* pre, int32_t* out, int32_t* n_keep_b, IdOf id_of) {
  const int t = threadIdx.x, T = blockDim.x;
  for (int w = t; w <= nw; w += T) {
    int s = 0;
    for (int u = 0; u < w; ++u) s += __popcll(K[u]);
    pre[w] = s;
  }
  __syncthreads();
  for (int j = t; j < n; j += T) {
    const unsigned long long kw = K[j >> 6];
    if ((kw >> (j & 63)) & 1ull) out[pre[j >> 6] + __popcll(kw & ((1ull << (j & 63)) - 1ull))] = id_of(j);
  }
  if (t == 0) *n_keep_b = pre[nw];
}

// ---------------------------------------------------------------------------------------------
// NMS step 3: suppression scan.  One CTA per image.
//   * whole matrix staged (n * ceil(n/64) words fit in the dynamic shared memory - up to ~1400 candidates whatever the
//     row capacity, the rows are staged compactly): parallel rounds (nms_parallel_scan), no dependent chain at all;
//   * otherwise, or when the rounds do not converge: the sequential scan - 64-box chunks are resolved in order, the
//     within-chunk dependency (64 steps on one 64-bit word) runs in one thread on registers (fully unrolled: test bit
//     q, OR row q's diagonal word), the propagation of the kept rows into the later words is one independent load per
//     row and thread.  The rows of the current chunk (words c .. nchunk-1) are staged in shared memory when they fit,
//     else read from global memory (very large candidate sets).
// ---------------------------------------------------------------------------------------------
constexpr int kScanThreads = 1024;
constexpr int kScanMaxWords = ISG_NMS_MAX_BOXES / 64;   // 256 suppression words a row at most

__global__ void __launch_bounds__(kScanThreads)
nms_scan_kernel(const int32_t* __restrict__ count, int cap, void* ws, int32_t* __restrict__ keep,
                int32_t* __restrict__ n_keep, int smem_words, int max_rounds) {
  extern __shared__ __align__(16) unsigned long long scan_rows[];
  const int b = blockIdx.x, t = threadIdx.x;
  const int n = min(max(count[b], 0), cap);
  const NmsWs v = nms_ws_view(ws, b, cap);
  const int nw = (cap + 63) / 64;
  const int nchunk = (n + 63) / 64;
  __shared__ unsigned long long remv[kScanMaxWords];
  __shared__ unsigned long long s_kept;
  if (t < kScanMaxWords) remv[t] = 0ull;
  int nk = 0;
  int32_t* out = keep + (size_t)b * cap;
  // staging mode (block-uniform): 2 = whole matrix, compact rows [n][nchunk]; 1 = the current chunk; 0 = none
  const int staged = (long long)n * nchunk <= smem_words ? 2 : (64 * nchunk <= smem_words ? 1 : 0);
  if (staged == 2) {
    // words below a row's diagonal block are never read (and never written by nms_mask_kernel): skip them
    const int total = n * nchunk;
    for (int e0 = t; e0 < total; e0 += 4 * kScanThreads) {
      unsigned long long q[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int e = e0 + u * kScanThreads;
        const int j = e / nchunk, w = e - j * nchunk;
        q[u] = (e < total && w >= (j >> 6)) ? v.mask[(size_t)j * nw + w] : 0ull;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) if (e0 + u * kScanThreads < total) scan_rows[e0 + u * kScanThreads] = q[u];
    }
  }
  __syncthreads();
  if (staged == 2 && max_rounds > 0) {
    __shared__ unsigned long long Kw[kScanMaxWords];
    __shared__ uint32_t R32[2 * kScanMaxWords];
    __shared__ int pre[kScanMaxWords + 1];
    if (nms_parallel_scan(scan_rows, nchunk, n, nchunk, Kw, R32, max_rounds)) {   // block-uniform
      const int32_t* order = v.order;
      nms_emit_kept(Kw, n, nchunk, pre, out, n_keep + b, [&](int j) { return order[j]; });
      return;
    }
    __syncthreads();
  }
  for (int c = 0; c < nchunk; ++c) {
    const int base = c * 64;
    const int m = min(64, n - base);
    const int nwc = nchunk - c;                 // staged 1: words c .. nchunk-1 of each row
    const unsigned long long* rows;
    int rstride, woff;
    if (staged == 2) { rows = scan_rows + (size_t)base * nchunk; rstride = nchunk; woff = 0; }
    else if (staged == 1) {
      for (int e = t; e < m * nwc; e += kScanThreads) {
        const int q = e / nwc, w = e - q * nwc;
        scan_rows[q * nwc + w] = v.mask[(size_t)(base + q) * nw + c + w];
      }
      __syncthreads();
      rows = scan_rows; rstride = nwc; woff = -c;
    } else { rows = v.mask + (size_t)base * nw; rstride = nw; woff = 0; }
    if (t == 0) {
      unsigned long long dq[64];
#pragma unroll
      for (int q = 0; q < 64; ++q) dq[q] = q < m ? rows[(size_t)q * rstride + c + woff] : 0ull;
      unsigned long long word = remv[c], kept = 0ull;
#pragma unroll
      for (int q = 0; q < 64; ++q) {
        const bool take = q < m && !((word >> q) & 1ull);
        kept |= take ? (1ull << q) : 0ull;
        word |= take ? dq[q] : 0ull;
      }
      s_kept = kept;
    }
    __syncthreads();
    const unsigned long long kept = s_kept;
    if (t < m && ((kept >> t) & 1ull))
      out[nk + __popcll(kept & ((1ull << t) - 1ull))] = v.order[base + t];
    nk += __popcll(kept);
    // propagate the kept rows of this chunk to the later words (independent loads, no dependent bit scan)
    const int w = t;
    if (w > c && w < nchunk) {
      unsigned long long acc = 0ull;
#pragma unroll 8
      for (int q = 0; q < 64; ++q)
        if ((kept >> q) & 1ull) acc |= rows[(size_t)q * rstride + w + woff];
      remv[w] |= acc;
    }
    __syncthreads();
  }
  if (t == 0) n_keep[b] = nk;
}

// launch helper: the kernel picks its staging mode from the candidate count; the dynamic shared memory is sized for the
// whole matrix when the row capacity allows it at all, else for one 64-row chunk
static inline cudaError_t launch_nms_scan(const int32_t* count, int B, int cap, void* ws, int32_t* keep, int32_t* n_keep,
                                          cudaStream_t stream) {
  const size_t nw = (size_t)(cap + 63) / 64;
  const size_t all = (size_t)cap * nw * 8, budget = 200 * 1024;
  const size_t smem = std::min(all, budget);
  const cudaError_t e = cudaFuncSetAttribute(nms_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  nms_scan_kernel<<<B, kScanThreads, smem, stream>>>(count, cap, ws, keep, n_keep, (int)(smem / 8), tuning().nms_rounds);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// NMS, small problems (cap <= 1024): sort + bitmask + greedy scan fused in ONE CTA per image, everything in
// shared memory.  The three-kernel path above remains for larger candidate sets.
// ---------------------------------------------------------------------------------------------
constexpr int kSmallMax = 1024;
constexpr int kRankSortMax = 512;   // up to this many candidates: rank sort instead of the bitonic network

__global__ void __launch_bounds__(kSortThreads)
nms_small_kernel(const float4* __restrict__ boxes, const float* __restrict__ scores, const int32_t* __restrict__ cls,
                 const int32_t* __restrict__ tiebreak, const int32_t* __restrict__ count, int cap, int P,
                 double thr, int convention, int32_t* __restrict__ keep, int32_t* __restrict__ n_keep, int max_rounds) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int nwP = P / 64 > 0 ? P / 64 : 1;
  float4* sbox = reinterpret_cast<float4*>(smem_raw);                               // [P]
  unsigned long long* key = reinterpret_cast<unsigned long long*>(sbox + P);        // [P]
  unsigned long long* mask = key + P;                                               // [P][nwP]
  uint32_t* val = reinterpret_cast<uint32_t*>(mask + (size_t)P * nwP);              // [P]
  int* scls = reinterpret_cast<int*>(val + P);                                      // [P]
  __shared__ unsigned long long remv[kSmallMax / 64];
  pdl_trigger();
  pdl_wait();         // candidates / counts come from the kernel launched just before
  const int b = blockIdx.x, t = threadIdx.x;
  const int n = min(max(count[b], 0), cap);
  int Pn = 1;
  while (Pn < n) Pn <<= 1;
  for (int i = t; i < Pn; i += kSortThreads) {
    unsigned long long k = 0ull;
    if (i < n) {
      const size_t o = (size_t)b * cap + i;
      const uint32_t tb = tiebreak ? (uint32_t)tiebreak[o] : (uint32_t)i;
      const uint32_t low = (convention == ISG_NMS_PLUS1_LE) ? tb : (0xffffffffu - tb);
      k = ((unsigned long long)float_key(scores[o]) << 32) | low;
    }
    key[i] = k; val[i] = (uint32_t)i;
  }
  __syncthreads();
  if (n <= kRankSortMax) {
    // few candidates: rank sort - every key counts the larger keys (broadcast reads, one barrier); the keys are distinct
    // (they carry the tie-break index), so the ranks are a permutation
    int rank = 0;
    if (t < n) {
      const unsigned long long mine = key[t];
      for (int j = 0; j < n; ++j) rank += key[j] > mine;
    }
    __syncthreads();
    if (t < n) val[rank] = (uint32_t)t;
    __syncthreads();
  } else {
    for (int size = 2; size <= Pn; size <<= 1) {
      for (int stride = size >> 1; stride > 0; stride >>= 1) {
        for (int i = t; i < (Pn >> 1); i += kSortThreads) {
          const int lo = 2 * i - (i & (stride - 1));
          const int hi = lo + stride;
          const bool desc = ((lo & size) == 0);
          const unsigned long long a = key[lo], c = key[hi];
          if ((a < c) == desc) {
            key[lo] = c; key[hi] = a;
            const uint32_t va = val[lo]; val[lo] = val[hi]; val[hi] = va;
          }
        }
        __syncthreads();
      }
    }
  }
  {
    __shared__ float red[32];
    const bool trick = cls && nms_uses_trick(convention, n);               // block-uniform
    const float mp1 = trick ? nms_block_max_plus_1(boxes + (size_t)b * cap, n, red) : 0.0f;
    for (int r = t; r < n; r += kSortThreads) {
      const size_t o = (size_t)b * cap + val[r];
      const int c = cls ? cls[o] : 0;
      sbox[r] = trick ? nms_shift_box(boxes[o], c, mp1) : boxes[o];        // sbox only feeds the IoU tests
      scls[r] = c;
    }
  }
  if (t < kSmallMax / 64) remv[t] = 0ull;
  __syncthreads();
  // suppression bits.  Only boxes of the same class can suppress each other, so the ranks are first bucketed by
  // (class & 63), each bucket keeping rank order; row i then visits only the later members of its bucket instead of
  // every later box.  The `key` array is free after the sort and holds the bucket lists.
  const int nw = (n + 63) / 64;
  const float thr_f = (float)thr;
  __shared__ int bcnt[64], bstart[65];
  uint16_t* memb = reinterpret_cast<uint16_t*>(key);            // [n] ranks grouped by bucket, ascending inside
  uint16_t* posof = memb + P;                                   // [n] position of rank r inside memb
  const int lane = t & 31, warp = t >> 5;
  if (t < 64) bcnt[t] = 0;
  for (int i = t; i < n * nw; i += kSortThreads) mask[(size_t)(i / nw) * nwP + (i % nw)] = 0ull;
  __syncthreads();
  for (int h = warp; h < 64; h += kSortThreads / 32) {          // bucket sizes (one warp per bucket, ballots)
    int c = 0;
    for (int r0 = 0; r0 < n; r0 += 32) {
      const int r = r0 + lane;
      c += __popc(__ballot_sync(0xffffffffu, r < n && (scls[r] & 63) == h));
    }
    if (lane == 0) bcnt[h] = c;
  }
  __syncthreads();
  if (t == 0) {
    int run = 0;
    for (int h = 0; h < 64; ++h) { bstart[h] = run; run += bcnt[h]; }
    bstart[64] = run;
  }
  __syncthreads();
  for (int h = warp; h < 64; h += kSortThreads / 32) {          // ordered fill
    int pos = bstart[h];
    for (int r0 = 0; r0 < n; r0 += 32) {
      const int r = r0 + lane;
      const bool in = r < n && (scls[r] & 63) == h;
      const unsigned bal = __ballot_sync(0xffffffffu, in);
      if (in) { const int q = pos + __popc(bal & ((1u << lane) - 1u)); memb[q] = (uint16_t)r; posof[r] = (uint16_t)q; }
      pos += __popc(bal);
    }
  }
  __syncthreads();
  if (n <= 256) {
    // up to 256 candidates (4 suppression words a row): 4 threads per row collect their share of the row in registers,
    // combine by shuffle and store the row once - no shared-memory atomics (64-bit ones are compare-and-swap loops)
    const int sub = t & 3, i = t >> 2;                          // kSortThreads / 4 = 256 rows in one sweep
    unsigned long long b0 = 0ull, b1 = 0ull, b2 = 0ull, b3 = 0ull;
    if (i < n) {
      const float4 a = sbox[i];
      const int ac = scls[i];
      const int qend = bstart[(ac & 63) + 1];
      for (int q = posof[i] + 1 + sub; q < qend; q += 4) {
        const int c = memb[q];                                  // rank > i, ascending
        if (scls[c] != ac) continue;                            // another class hashed into the same bucket
        const bool sup = (convention != ISG_NMS_PLUS1_LE) ? suppresses_tv(a, sbox[c], thr) : suppresses_plus1(a, sbox[c], thr_f);
        if (!sup) continue;
        const unsigned long long bit = 1ull << (c & 63);
        const int w = c >> 6;
        if (w == 0) b0 |= bit; else if (w == 1) b1 |= bit; else if (w == 2) b2 |= bit; else b3 |= bit;
      }
    }
#pragma unroll
    for (int o = 1; o < 4; o <<= 1) {
      b0 |= __shfl_xor_sync(0xffffffffu, b0, o); b1 |= __shfl_xor_sync(0xffffffffu, b1, o);
      b2 |= __shfl_xor_sync(0xffffffffu, b2, o); b3 |= __shfl_xor_sync(0xffffffffu, b3, o);
    }
    if (i < n && sub < nw) mask[(size_t)i * nwP + sub] = sub == 0 ? b0 : (sub == 1 ? b1 : (sub == 2 ? b2 : b3));
  } else {
    // R threads per row share its successors (the rows are short, the block would otherwise be mostly idle)
    const int R = n <= 512 ? 2 : 1;
    const int sub = t % R;
    for (int i = t / R; i < n; i += kSortThreads / R) {
      const float4 a = sbox[i];
      const int ac = scls[i];
      const int qend = bstart[(ac & 63) + 1];
      int cur_w = -1;
      unsigned long long bits = 0ull;
      for (int q = posof[i] + 1 + sub; q < qend; q += R) {
        const int c = memb[q];                                  // rank > i, ascending
        if (scls[c] != ac) continue;                            // another class hashed into the same bucket
        const bool sup = (convention != ISG_NMS_PLUS1_LE) ? suppresses_tv(a, sbox[c], thr) : suppresses_plus1(a, sbox[c], thr_f);
        if (!sup) continue;
        const int w = c >> 6;
        if (w != cur_w) { if (cur_w >= 0) atomicOr(&mask[(size_t)i * nwP + cur_w], bits); cur_w = w; bits = 0ull; }
        bits |= 1ull << (c & 63);
      }
      if (cur_w >= 0) atomicOr(&mask[(size_t)i * nwP + cur_w], bits);
    }
  }
  __syncthreads();
  // parallel suppression scan (nms_parallel_scan); the sequential scan below only runs when it did not converge
  if (max_rounds > 0) {
    __shared__ unsigned long long Kw[kSmallMax / 64];
    __shared__ uint32_t R32[2 * kSmallMax / 64];
    __shared__ int pre[kSmallMax / 64 + 1];
    if (nms_parallel_scan(mask, nwP, n, nw, Kw, R32, max_rounds)) {               // block-uniform
      nms_emit_kept(Kw, n, nw, pre, keep + (size_t)b * cap, n_keep + b, [&](int j) { return (int32_t)val[j]; });
      return;
    }
    __syncthreads();
  }
  // greedy scan by warp 0, 64-box chunks, no block barriers.  The 64 dependent steps of a chunk run in ONE lane on
  // registers (fully unrolled: test bit q of the running word, OR row q's diagonal word): ~12 cycles a step whether the
  // box is kept or not, instead of a shuffle round trip (~180 cycles) per kept box.  The kept rows are then OR-ed into
  // the later words by all lanes.
  int nk = 0;
  if (warp != 0) return;
  int32_t* out = keep + (size_t)b * cap;
  for (int c = 0; c < nw; ++c) {
    const int base = c * 64;
    const int m = min(64, n - base);
    unsigned long long kept = 0ull;
    if (lane == 0) {
      unsigned long long dq[64];
#pragma unroll
      for (int q = 0; q < 64; ++q) dq[q] = q < m ? mask[(size_t)(base + q) * nwP + c] : 0ull;
      unsigned long long word = remv[c];
#pragma unroll
      for (int q = 0; q < 64; ++q) {
        const bool take = q < m && !((word >> q) & 1ull);
        kept |= take ? (1ull << q) : 0ull;
        word |= take ? dq[q] : 0ull;
      }
    }
    kept = __shfl_sync(0xffffffffu, kept, 0);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int bit = lane + 32 * h;
      if (bit < m && ((kept >> bit) & 1ull)) out[nk + __popcll(kept & ((1ull << bit) - 1ull))] = (int32_t)val[base + bit];
    }
    nk += __popcll(kept);
    for (int w = c + 1; w < nw; ++w) {
      unsigned long long v = 0ull;
      if ((kept >> lane) & 1ull) v |= mask[(size_t)(base + lane) * nwP + w];
      if ((kept >> (lane + 32)) & 1ull) v |= mask[(size_t)(base + 32 + lane) * nwP + w];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v |= __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) remv[w] |= v;
    }
    __syncwarp();
  }
  if (t == 0) n_keep[b] = nk;
}

// ---------------------------------------------------------------------------------------------
// NMS, large candidate sets (cap > ISG_NMS_MAX_BOXES, e.g. an untrained head that fires on every anchor): no
// suppression matrix (it would be cap^2 / 8 bytes).  (1) one CTA per image sorts the candidates by (score desc,
// tie-break) with a stable LSD radix sort in global memory; (2) the ranks are resolved tile by tile (1024 ranks):
// nms_tile_resolve_kernel decides the tile's still-alive members among themselves (1024 x 1024 bit matrix in shared
// memory + the parallel suppression scan) and appends the kept ones to the keep list, nms_tile_suppress_kernel then
// clears the alive bit of every later candidate one of them suppresses.  Work is O(n * kept) like the reference's
// torchvision call; the path exists for completeness, not for speed.
// ---------------------------------------------------------------------------------------------
constexpr int kLargeTile = 1024;
constexpr int kLargeTileWords = kLargeTile / 64;

struct NmsLargeWs {
  unsigned long long* key[2];   // [cap] x 2 (ping-pong)
  uint32_t* val[2];             // [cap] x 2
  int32_t* order;               // [cap] candidate index of rank r
  float4* sbox;                 // [cap] boxes in rank order (shifted for the coordinate trick)
  int32_t* scls;                // [cap]
  uint32_t* alive;              // [ceil(cap/32)]
  float4* tbox;                 // [kLargeTile] kept boxes of the current tile
  int32_t* tcls;                // [kLargeTile]
  int32_t* tcount;              // [4]
};
__host__ __device__ inline size_t nms_large_ws_per_image(int cap) {
  const size_t c = ((size_t)cap + 63) & ~(size_t)63;
  return c * (8 * 2 + 4 * 2 + 4 + 16 + 4) + ((c / 8 + 128 + 15) & ~(size_t)15) + (size_t)kLargeTile * 20 + 256;
}
__host__ __device__ inline NmsLargeWs nms_large_ws_view(void* ws, int b, int cap) {
  const size_t c = ((size_t)cap + 63) & ~(size_t)63;
  char* p = (char*)ws + (size_t)b * ((nms_large_ws_per_image(cap) + 255) & ~(size_t)255);
  NmsLargeWs v;
  v.key[0] = (unsigned long long*)p; p += c * 8;
  v.key[1] = (unsigned long long*)p; p += c * 8;
  v.sbox = (float4*)p; p += c * 16;
  v.val[0] = (uint32_t*)p; p += c * 4;
  v.val[1] = (uint32_t*)p; p += c * 4;
  v.order = (int32_t*)p; p += c * 4;
  v.scls = (int32_t*)p; p += c * 4;
  v.alive = (uint32_t*)p; p += (c / 8 + 128 + 15) & ~(size_t)15;          // one tile of slack: the last tile reads 32 whole words
  v.tbox = (float4*)p; p += (size_t)kLargeTile * 16;
  v.tcls = (int32_t*)p; p += (size_t)kLargeTile * 4;
  v.tcount = (int32_t*)p;
  return v;
}

// stable LSD radix sort (8 passes of 8 bits over the inverted 64-bit key = descending order), one CTA per image
__global__ void __launch_bounds__(1024)
nms_large_sort_kernel(const float4* __restrict__ boxes, const float* __restrict__ scores, const int32_t* __restrict__ cls,
                      const int32_t* __restrict__ tiebreak, const int32_t* __restrict__ count, int cap, int convention,
                      void* ws, int32_t* __restrict__ n_keep) {
  __shared__ uint32_t hist[256], base[256];
  __shared__ uint32_t wcnt[32][257];
  __shared__ float red[32];
  const int b = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int n = min(max(count[b], 0), cap);
  NmsLargeWs v = nms_large_ws_view(ws, b, cap);
  for (int i = t; i < n; i += 1024) {
    const size_t o = (size_t)b * cap + i;
    const uint32_t tb = tiebreak ? (uint32_t)tiebreak[o] : (uint32_t)i;
    const uint32_t low = (convention == ISG_NMS_PLUS1_LE) ? tb : (0xffffffffu - tb);
    v.key[0][i] = ~(((unsigned long long)float_key(scores[o]) << 32) | low);     // ascending on the inverted key
    v.val[0][i] = (uint32_t)i;
  }
  for (int w = t; w < (cap + 63) / 64 * 2 + 32; w += 1024) {
    const int left = n - w * 32;
    v.alive[w] = left >= 32 ? 0xffffffffu : (left <= 0 ? 0u : ((1u << left) - 1u));
  }
  if (t == 0) { n_keep[b] = 0; v.tcount[0] = 0; }
  __syncthreads();
  int cur = 0;
  for (int pass = 0; pass < 8; ++pass) {
    const int shift = 8 * pass;
    const unsigned long long* kin = v.key[cur]; const uint32_t* vin = v.val[cur];
    unsigned long long* kout = v.key[cur ^ 1]; uint32_t* vout = v.val[cur ^ 1];
    if (t < 256) hist[t] = 0u;
    __syncthreads();
    for (int i = t; i < n; i += 1024) atomicAdd(&hist[(uint32_t)(kin[i] >> shift) & 255u], 1u);
    __syncthreads();
    if (t == 0) { uint32_t run = 0; for (int d = 0; d < 256; ++d) { base[d] = run; run += hist[d]; } }
    __syncthreads();
    for (int t0 = 0; t0 < n; t0 += 1024) {
      for (int e = t; e < 32 * 257; e += 1024) (&wcnt[0][0])[e] = 0u;
      __syncthreads();
      const int i = t0 + t;
      const bool valid = i < n;
      unsigned long long k = 0ull; uint32_t vv = 0u; uint32_t d = 256u;
      if (valid) { k = kin[i]; vv = vin[i]; d = (uint32_t)(k >> shift) & 255u; }
      const unsigned same = __match_any_sync(0xffffffffu, d);
      const int rank_in_warp = __popc(same & ((1u << lane) - 1u));
      if (valid && rank_in_warp == 0) wcnt[warp][d] = (uint32_t)__popc(same);
      __syncthreads();
      if (t < 256) {                                   // per digit: offsets of the warps in warp order, advance the base
        uint32_t run = base[t];
        for (int w = 0; w < 32; ++w) { const uint32_t c = wcnt[w][t]; wcnt[w][t] = run; run += c; }
        base[t] = run;
      }
      __syncthreads();
      if (valid) { const uint32_t pos = wcnt[warp][d] + (uint32_t)rank_in_warp; kout[pos] = k; vout[pos] = vv; }
      __syncthreads();
    }
    cur ^= 1;
  }
  const uint32_t* val = v.val[cur];
  const bool trick = cls && nms_uses_trick(convention, n);                 // block-uniform
  const float mp1 = trick ? nms_block_max_plus_1(boxes + (size_t)b * cap, n, red) : 0.0f;
  for (int r = t; r < n; r += 1024) {
    const int i = (int)val[r];
    const size_t o = (size_t)b * cap + i;
    const int c = cls ? cls[o] : 0;
    v.order[r] = i;
    v.sbox[r] = trick ? nms_shift_box(boxes[o], c, mp1) : boxes[o];
    v.scls[r] = c;
  }
}

// tile `tile` (ranks tile*1024 ..): the alive members decide among themselves; kept ones go to the keep list (rank
// order) and, with their boxes, to the tile table nms_tile_suppress_kernel reads
__global__ void __launch_bounds__(1024)
nms_tile_resolve_kernel(const int32_t* __restrict__ count, int cap, int tile, double thr, int convention, void* ws,
                        int32_t* __restrict__ keep, int32_t* __restrict__ n_keep, int max_rounds) {
  extern __shared__ __align__(16) unsigned char tile_smem[];
  float4* sb = reinterpret_cast<float4*>(tile_smem);                                   // [1024]
  unsigned long long* rows = reinterpret_cast<unsigned long long*>(sb + kLargeTile);     // [1024][16]
  int* sc = reinterpret_cast<int*>(rows + (size_t)kLargeTile * kLargeTileWords);         // [1024]
  __shared__ unsigned long long alive[kLargeTileWords], Kw[kLargeTileWords];
  __shared__ uint32_t R32[2 * kLargeTileWords];
  __shared__ int pre[kLargeTileWords + 1];
  const int b = blockIdx.x, t = threadIdx.x;
  const int n = min(max(count[b], 0), cap);
  const int base = tile * kLargeTile;
  NmsLargeWs v = nms_large_ws_view(ws, b, cap);
  if (base >= n) { if (t == 0) v.tcount[0] = 0; return; }
  const int m = min(kLargeTile, n - base);
  if (t < kLargeTileWords) {
    const uint32_t lo = v.alive[(base >> 5) + 2 * t], hi = v.alive[(base >> 5) + 2 * t + 1];
    alive[t] = (unsigned long long)lo | ((unsigned long long)hi << 32);
  }
  if (t < m) { sb[t] = v.sbox[base + t]; sc[t] = v.scls[base + t]; }
  __syncthreads();
  {
    const float thr_f = (float)thr;
    const bool me = t < m && ((alive[t >> 6] >> (t & 63)) & 1ull);
    const float4 a = me ? sb[t] : make_float4(0.f, 0.f, 0.f, 0.f);
    const int ac = me ? sc[t] : 0;
    unsigned long long bits = 0ull;
    int cw = t >> 6;
    for (int w = 0; w < (t >> 6); ++w) rows[(size_t)t * kLargeTileWords + w] = 0ull;
    for (int c = (t & ~63); c < ((m + 63) & ~63); ++c) {                                  // words from the diagonal one on
      if ((c >> 6) != cw) { rows[(size_t)t * kLargeTileWords + cw] = bits; bits = 0ull; cw = c >> 6; }
      if (!me || c <= t || c >= m || sc[c] != ac || !((alive[c >> 6] >> (c & 63)) & 1ull)) continue;
      const bool sup = (convention != ISG_NMS_PLUS1_LE) ? suppresses_tv(a, sb[c], thr) : suppresses_plus1(a, sb[c], thr_f);
      if (sup) bits |= 1ull << (c & 63);
    }
    rows[(size_t)t * kLargeTileWords + cw] = bits;
    for (int w = cw + 1; w < kLargeTileWords; ++w) rows[(size_t)t * kLargeTileWords + w] = 0ull;
  }
  __syncthreads();
  const int nw = (m + 63) >> 6;
  if (!nms_parallel_scan(rows, kLargeTileWords, m, nw, Kw, R32, max_rounds, alive)) {    // block-uniform
    // did not converge (a suppression chain): sequential scan by warp 0
    __syncthreads();
    if (t < 32) {
      unsigned long long rem = 0ull, kept = 0ull;      // lane w < 16 owns removed / kept word w
      for (int i = 0; i < m; ++i) {
        const int wi = i >> 6;
        const unsigned long long rw = __shfl_sync(0xffffffffu, rem, wi), aw = alive[wi];
        const bool take = ((aw >> (i & 63)) & 1ull) && !((rw >> (i & 63)) & 1ull);     // warp-uniform
        if (take) {
          if (t == wi) kept |= 1ull << (i & 63);
          if (t < kLargeTileWords) rem |= rows[(size_t)i * kLargeTileWords + t];
        }
      }
      if (t < kLargeTileWords) Kw[t] = kept;
    }
    __syncthreads();
  }
  const int nk0 = n_keep[b];
  const int32_t* order = v.order;
  __shared__ int tile_kept;
  nms_emit_kept(Kw, m, nw, pre, keep + (size_t)b * cap + nk0, &tile_kept, [&](int j) { return order[base + j]; });
  __syncthreads();
  if (t < m && ((Kw[t >> 6] >> (t & 63)) & 1ull)) {
    const int pos = pre[t >> 6] + __popcll(Kw[t >> 6] & ((1ull << (t & 63)) - 1ull));
    v.tbox[pos] = sb[t]; v.tcls[pos] = sc[t];
  }
  if (t == 0) { v.tcount[0] = tile_kept; n_keep[b] = nk0 + tile_kept; }
}

// every still-alive candidate behind the tile against the tile's kept boxes; a warp owns one word of the alive mask
__global__ void __launch_bounds__(256)
nms_tile_suppress_kernel(const int32_t* __restrict__ count, int cap, int tile, double thr, int convention, void* ws) {
  __shared__ float4 kb[kLargeTile];
  __shared__ int kc[kLargeTile];
  const int b = blockIdx.y, t = threadIdx.x, lane = t & 31;
  const int n = min(max(count[b], 0), cap);
  NmsLargeWs v = nms_large_ws_view(ws, b, cap);
  const int first = (tile + 1) * kLargeTile;
  if (first >= n) return;
  const int nkept = v.tcount[0];
  if (nkept == 0) return;
  for (int i = t; i < nkept; i += 256) { kb[i] = v.tbox[i]; kc[i] = v.tcls[i]; }
  __syncthreads();
  const float thr_f = (float)thr;
  for (long long j0 = (long long)first + ((long long)blockIdx.x * 256 + (t - lane)); j0 < n; j0 += (long long)gridDim.x * 256) {
    const int j = (int)j0 + lane;
    const uint32_t word = v.alive[j0 >> 5];
    bool live = j < n && ((word >> lane) & 1u);
    if (live) {
      const float4 c = v.sbox[j];
      const int cc = v.scls[j];
      for (int i = 0; i < nkept; ++i) {
        if (kc[i] != cc) continue;
        const bool sup = (convention != ISG_NMS_PLUS1_LE) ? suppresses_tv(kb[i], c, thr) : suppresses_plus1(kb[i], c, thr_f);
        if (sup) { live = false; break; }
      }
    }
    const uint32_t still = __ballot_sync(0xffffffffu, live);
    if (lane == 0 && still != word) v.alive[j0 >> 5] = still;
  }
}

static int run_nms_large(const float* boxes, const float* scores, const int32_t* cls, const int32_t* tiebreak,
                         const int32_t* count, int B, int cap, double thr, int convention, int32_t* keep,
                         int32_t* n_keep, void* ws, cudaStream_t stream) {
  nms_large_sort_kernel<<<B, 1024, 0, stream>>>(reinterpret_cast<const float4*>(boxes), scores, cls, tiebreak, count, cap,
                                                convention, ws, n_keep);
  ISG_LAUNCH_CHECK();
  const size_t smem = (size_t)kLargeTile * (16 + 8 * kLargeTileWords + 4);
  ISG_CUDA(cudaFuncSetAttribute(nms_tile_resolve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int tiles = cdiv(cap, kLargeTile);
  const int rounds = tuning().nms_rounds;
  for (int tile = 0; tile < tiles; ++tile) {
    nms_tile_resolve_kernel<<<B, 1024, smem, stream>>>(count, cap, tile, thr, convention, ws, keep, n_keep, rounds);
    const int later = cap - (tile + 1) * kLargeTile;
    if (later > 0) {
      dim3 grid((unsigned)std::min(cdiv(later, 256), 148 * 8), (unsigned)B);
      nms_tile_suppress_kernel<<<grid, 256, 0, stream>>>(count, cap, tile, thr, convention, ws);
    }
  }
  ISG_LAUNCH_CHECK();
  return ISG_OK;
}

// ---------------------------------------------------------------------------------------------
// K6 mask NMS
// ---------------------------------------------------------------------------------------------
struct MaskWs {
  int32_t* order;            // [n]
  int32_t* area;             // [n]
  int4* bbox;                // [n] x0,y0,x1,y1 inclusive (tight), empty mask: x0 > x1
  int32_t* scls;             // [n] (rank order)
  int32_t* cnt;              // [1] = n (device copy for the shared scan kernel)
  unsigned long long* mask;  // [n][nw]
};
// The sort/scan kernels are shared with box NMS, so the layout must start like NmsWs(cap = n).
__host__ __device__ inline size_t mask_extra_bytes(int n) {
  return (((size_t)n * 4 + 15) & ~(size_t)15) + (size_t)n * 16 + 256;
}

// per-mask popcount + tight bounding box: the pass that reads every mask word once (the HBM-bound stage of isg_mask_nms)
__global__ void __launch_bounds__(256)
mask_area_kernel(const uint32_t* __restrict__ masks, int H, int Wwords, const int4* __restrict__ given,
                 int32_t* __restrict__ area, int4* __restrict__ bbox) {
  const int i = blockIdx.x, t = threadIdx.x;
  const uint32_t* m = masks + (size_t)i * H * Wwords;
  int a = 0, bx0 = 0x7fffffff, by0 = 0x7fffffff, bx1 = -1, by1 = -1;
  auto take = [&](uint32_t v, int e) {     // word e of the flat mask; most words of a mask are empty
    if (v) {
      a += __popc(v);
      const int y = e / Wwords, xb = (e - y * Wwords) * 32;
      by0 = min(by0, y); by1 = max(by1, y);
      bx0 = min(bx0, xb + __ffs(v) - 1); bx1 = max(bx1, xb + 31 - __clz(v));
    }
  };
  if (given) {
    const int4 g = given[i];
    const int y0 = max(g.y, 0), y1 = min(g.w, H - 1), w0 = max(g.x >> 5, 0), w1 = min(g.z >> 5, Wwords - 1);
    const int nwords = max(w1 - w0 + 1, 0), nrows = max(y1 - y0 + 1, 0);
    for (int e = t; e < nwords * nrows; e += 256) {
      const int r = e / nwords, w = e - r * nwords;
      take(__ldg(m + (size_t)(y0 + r) * Wwords + w0 + w), (y0 + r) * Wwords + w0 + w);
    }
  } else {
    const int total = H * Wwords;
    const int head = min(total, (int)(((16 - ((uintptr_t)m & 15)) & 15) >> 2));     // words before the first 16-byte boundary
    if (t < head) take(__ldg(m + t), t);
    const uint4* m4 = reinterpret_cast<const uint4*>(m + head);
    const int n4 = (total - head) >> 2;
    constexpr int kU = 4;                                                            // independent 128-bit loads in flight
    for (int e0 = t; e0 < n4; e0 += 256 * kU) {
      uint4 q[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int e = e0 + u * 256;
        q[u] = e < n4 ? __ldg(m4 + e) : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        if (q[u].x | q[u].y | q[u].z | q[u].w) {
          const int e = head + (e0 + u * 256) * 4;
          take(q[u].x, e); take(q[u].y, e + 1); take(q[u].z, e + 2); take(q[u].w, e + 3);
        }
      }
    }
    for (int e = head + n4 * 4 + t; e < total; e += 256) take(__ldg(m + e), e);
  }
  __shared__ int sa[8], s0[8], s1[8], s2[8], s3[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    bx0 = min(bx0, __shfl_xor_sync(0xffffffffu, bx0, o)); by0 = min(by0, __shfl_xor_sync(0xffffffffu, by0, o));
    bx1 = max(bx1, __shfl_xor_sync(0xffffffffu, bx1, o)); by1 = max(by1, __shfl_xor_sync(0xffffffffu, by1, o));
  }
  if ((t & 31) == 0) { sa[t >> 5] = a; s0[t >> 5] = bx0; s1[t >> 5] = by0; s2[t >> 5] = bx1; s3[t >> 5] = by1; }
  __syncthreads();
  if (t == 0) {
    for (int w = 1; w < 8; ++w) {
      a += sa[w]; bx0 = min(bx0, s0[w]); by0 = min(by0, s1[w]); bx1 = max(bx1, s2[w]); by1 = max(by1, s3[w]);
    }
    area[i] = a;
    bbox[i] = make_int4(bx0, by0, bx1, by1);
  }
  // launched with programmatic stream serialisation behind the (independent) rank sort: this grid only completes once
  // the sort has, so that everything later in the stream sees both results (a no-op for an ordinary launch)
  pdl_wait();
}

// One warp per (row ri, block of 32 later ranks rj): the lanes test class equality and bounding-box overlap of their
// own rj in parallel (coalesced table reads), then the warp walks the few surviving pairs together, every lane taking a
// share of the words of the bbox intersection.
__global__ void __launch_bounds__(256)
mask_pair_kernel(const uint32_t* __restrict__ masks, int n, int H, int Wwords, const int32_t* __restrict__ order,
                 const int32_t* __restrict__ scls, const int32_t* __restrict__ area, const int4* __restrict__ bbox,
                 double thr, unsigned long long* __restrict__ mask, int nw) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int nb = (n + 31) >> 5;
  const long long items = (long long)n * nb;
  for (long long p = warp0; p < items; p += nwarps) {
    const int ri = (int)(p / nb), jb = (int)(p - (long long)ri * nb);
    if (jb * 32 + 31 <= ri) continue;                        // the whole block lies on or below the diagonal
    const int i = order[ri];
    const int ci = scls[ri];
    const int4 bi = bbox[i];
    const int rj = jb * 32 + lane;
    int j = 0;
    int4 bj = make_int4(0, 0, -1, -1);
    bool cand = rj < n && rj > ri && scls[rj] == ci;
    if (cand) { j = order[rj]; bj = bbox[j]; }
    const int x0 = max(bi.x, bj.x), y0 = max(bi.y, bj.y), x1 = min(bi.z, bj.z), y1 = min(bi.w, bj.w);
    const bool overlap = cand && x0 <= x1 && y0 <= y1;
    // pairs without a common pixel: IoU = 1 / (union + 1) (utils/image.py:188-191), decided by the lane itself
    if (cand && !overlap) {
      const long long uni = (long long)area[i] + (long long)area[j];
      if (!(1.0 / (double)(uni + 1) <= thr)) atomicOr(&mask[(size_t)ri * nw + (rj >> 6)], 1ull << (rj & 63));
    }
    unsigned todo = __ballot_sync(0xffffffffu, overlap);
    const uint32_t* mi = masks + (size_t)i * H * Wwords;
    while (todo) {
      const int src = __ffs(todo) - 1;
      todo &= todo - 1;
      const int jj = __shfl_sync(0xffffffffu, j, src);
      const int qx0 = __shfl_sync(0xffffffffu, x0, src), qy0 = __shfl_sync(0xffffffffu, y0, src);
      const int qx1 = __shfl_sync(0xffffffffu, x1, src), qy1 = __shfl_sync(0xffffffffu, y1, src);
      const int w0 = qx0 >> 5, wn = (qx1 >> 5) - w0 + 1, nr = qy1 - qy0 + 1;
      const uint32_t* mj = masks + (size_t)jj * H * Wwords;
      int c = 0;
      for (int e = lane; e < wn * nr; e += 32) {
        const int r = e / wn, w = e - r * wn;
        const size_t o = (size_t)(qy0 + r) * Wwords + w0 + w;
        c += __popc(__ldg(mi + o) & __ldg(mj + o));
      }
      const long long inter = warp_sum(c);
      if (lane == src) {
        const long long uni = (long long)area[i] + (long long)area[j] - inter;
        const double iou = (double)(inter + 1) / (double)(uni + 1);   // utils/image.py:188-191
        if (!(iou <= thr)) atomicOr(&mask[(size_t)ri * nw + (rj >> 6)], 1ull << (rj & 63));
      }
    }
  }
}

__global__ void mask_pair_counts_kernel(const uint32_t* __restrict__ masks, int H, int Wwords,
                                        const int32_t* __restrict__ pairs, long long* __restrict__ out) {
  const int p = blockIdx.x, t = threadIdx.x;
  const uint32_t* ma = masks + (size_t)pairs[2 * p] * H * Wwords;
  const uint32_t* mb = masks + (size_t)pairs[2 * p + 1] * H * Wwords;
  long long in = 0, un = 0;
  for (int e = t; e < H * Wwords; e += blockDim.x) {
    const uint32_t a = __ldg(ma + e), c = __ldg(mb + e);
    in += __popc(a & c); un += __popc(a | c);
  }
  __shared__ long long si[8], su[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { in += __shfl_xor_sync(0xffffffffu, in, o); un += __shfl_xor_sync(0xffffffffu, un, o); }
  if ((t & 31) == 0) { si[t >> 5] = in; su[t >> 5] = un; }
  __syncthreads();
  if (t == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { in += si[w]; un += su[w]; }
    out[2 * p] = in; out[2 * p + 1] = un;
  }
}

__global__ void set_int_kernel(int32_t* p, int v) { *p = v; }

// BBoxTransform for every anchor (utils/utils.py:318-346), optional ClipBoxes (:349-363)
__global__ void bbox_transform_kernel(const float4* __restrict__ anchors, const float4* __restrict__ regression, int A,
                                      int clip, float xmax_clip, float ymax_clip, float4* __restrict__ out) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (a >= A) return;
  const float4 an = __ldg(anchors + a);
  const float4 rg = regression[(size_t)b * A + a];
  const float yca = __fmul_rn(__fadd_rn(an.x, an.z), 0.5f), xca = __fmul_rn(__fadd_rn(an.y, an.w), 0.5f);
  const float ha = __fsub_rn(an.z, an.x), wa = __fsub_rn(an.w, an.y);
  const float w = __fmul_rn(expf(rg.w), wa), h = __fmul_rn(expf(rg.z), ha);
  const float yc = __fadd_rn(__fmul_rn(rg.x, ha), yca), xc = __fadd_rn(__fmul_rn(rg.y, wa), xca);
  float ymin = __fsub_rn(yc, __fmul_rn(h, 0.5f)), xmin = __fsub_rn(xc, __fmul_rn(w, 0.5f));
  float ymax = __fadd_rn(yc, __fmul_rn(h, 0.5f)), xmax = __fadd_rn(xc, __fmul_rn(w, 0.5f));
  if (clip) {
    xmin = fmaxf(xmin, 0.0f); ymin = fmaxf(ymin, 0.0f);
    xmax = fminf(xmax, xmax_clip); ymax = fminf(ymax, ymax_clip);
  }
  out[(size_t)b * A + a] = make_float4(xmin, ymin, xmax, ymax);
}

__global__ void clip_boxes_kernel(float4* __restrict__ boxes, long long n, float xmax_clip, float ymax_clip) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 v = boxes[i];
  v.x = fmaxf(v.x, 0.0f); v.y = fmaxf(v.y, 0.0f); v.z = fminf(v.z, xmax_clip); v.w = fminf(v.w, ymax_clip);
  boxes[i] = v;
}

// dense masks (one byte per pixel, non-zero = set) -> bit-packed words; one warp per output word group
__global__ void pack_masks_kernel(const uint8_t* __restrict__ dense, long long rows, int W, int Wwords,
                                  uint32_t* __restrict__ bits) {
  const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;   // global warp = (row, word)
  const int lane = threadIdx.x & 31;
  if (gw >= rows * Wwords) return;
  const long long r = gw / Wwords;
  const int w = (int)(gw - r * Wwords);
  const int x = w * 32 + lane;
  const bool on = (x < W) && dense[r * W + x] != 0;
  const unsigned bal = __ballot_sync(0xffffffffu, on);
  if (lane == 0) bits[gw] = bal;
}

// kept candidates -> dense per-image detection tables in pick (= score descending) order
__global__ void gather_kept_kernel(const float4* __restrict__ cand_boxes, const float* __restrict__ cand_scores,
                                   const int32_t* __restrict__ cand_cls, const int32_t* __restrict__ keep,
                                   const int32_t* __restrict__ n_keep, int cap, int Nmax, float4* __restrict__ rois,
                                   float* __restrict__ scores, int32_t* __restrict__ cls, int32_t* __restrict__ n_out) {
  const int b = blockIdx.y;
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = min(n_keep[b], Nmax);
  if (r == 0) n_out[b] = n;
  if (r >= Nmax) return;
  float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);
  float sc = 0.f;
  int c = 0;
  if (r < n) {
    const size_t o = (size_t)b * cap + keep[(size_t)b * cap + r];
    bx = cand_boxes[o]; sc = cand_scores[o]; c = cand_cls[o];
  }
  const size_t d = (size_t)b * Nmax + r;
  rois[d] = bx; scores[d] = sc; cls[d] = c;
}

}  // namespace isg

using namespace isg;

static bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }
static int next_pow2(int n) { int p = 1; while (p < n) p <<= 1; return p; }

extern "C" int isg_decode_boxes(const float* anchors, const float* regression, const float* classification, int B,
                                int A, int C, int H, int W, float thr, int cap, float* cand_boxes,
                                float* cand_scores, int32_t* cand_cls, int32_t* cand_anchor, int32_t* cand_count,
                                isg_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!anchors || !regression || !classification || !cand_boxes || !cand_scores || !cand_cls || !cand_anchor || !cand_count)
    return ISG_EINVAL;
  if (B <= 0 || A <= 0 || C <= 0 || H <= 0 || W <= 0 || cap <= 0 || B > 65535) return ISG_EINVAL;
  if (!aligned16(anchors) || !aligned16(regression) || !aligned16(cand_boxes)) return ISG_EINVAL;
  const bool vec = (C % 4 == 0) && aligned16(classification);
  ISG_CUDA(cudaMemsetAsync(cand_count, 0, (size_t)B * sizeof(int32_t), stream));
  dim3 grid(cdiv(A, kFrontThreads * kFrontPer), B);
  decode_boxes_kernel<<<grid, kFrontThreads, 0, stream>>>(anchors, regression, classification, A, C, (float)(W - 1),
                                                          (float)(H - 1), thr, cap, reinterpret_cast<float4*>(cand_boxes),
                                                          cand_scores, cand_cls, cand_anchor, cand_count, vec);
  ISG_LAUNCH_CHECK();
  return ISG_OK;
}

extern "C" int isg_gather_kept(const float* cand_boxes, const float* cand_scores, const int32_t* cand_cls,
                               const int32_t* keep, const int32_t* n_keep, int B, int cap, int Nmax, float* rois,
                               float* scores, int32_t* cls, int32_t* n_out, isg_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!cand_boxes || !cand_scores || !cand_cls || !keep || !n_keep || !rois || !scores || !cls || !n_out) return ISG_EINVAL;
  if (B <= 0 || cap <= 0 || Nmax <= 0 || B > 65535) return ISG_EINVAL;
  if (!aligned16(cand_boxes) || !aligned16(rois)) return ISG_EINVAL;
  dim3 grid(cdiv(Nmax, 128), B);
  gather_kept_kernel<<<grid, 128, 0, stream>>>(reinterpret_cast<const float4*>(cand_boxes), cand_scores, cand_cls, keep,
                                               n_keep, cap, Nmax, reinterpret_cast<float4*>(rois), scores, cls, n_out);
  ISG_LAUNCH_CHECK();
  return ISG_OK;
}

extern "C" int isg_bbox_transform(const float* anchors, const float* regression, int B, int A, int clip, int H, int W,
                                  float* boxes, isg_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!anchors || !regression || !boxes || B <= 0 || A <= 0 || B > 65535) return ISG_EINVAL;
  if (!aligned16(anchors) || !aligned16(regression) || !aligned16(boxes)) return ISG_EINVAL;
  dim3 grid(cdiv(A, 256), B);
  bbox_transform_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const float4*>(anchors),
                                                  reinterpret_cast<const float4*>(regression), A, clip, (float)(W - 1),
                                                  (float)(H - 1), reinterpret_cast<float4*>(boxes));
  ISG_LAUNCH_CHECK();
  return ISG_OK;
}

extern "C" int isg_clip_boxes(float* boxes, int64_t n, int H, int W, isg_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!boxes || n <= 0 || !aligned16(boxes)) return ISG_EINVAL;
  clip_boxes_kernel<<<(unsigned)cdiv64(n, 256), 256, 0, stream>>>(reinterpret_cast<float4*>(boxes), n, (float)(W - 1),
                                                                 (float)(H - 1));
  ISG_LAUNCH_CHECK();
  return ISG_OK;
}

// ---- anchor generation (Anchors.forward, utils/utils.py:366-450) ------------------------------------------------
// The reference builds the table in numpy float64 and casts it once (`astype`, :441): every coordinate here is the same
// fp64 expression (cell centre stride/2 + i*stride, exact; minus/plus the fp64 half size computed by the host exactly as
// :423-425 do) followed by ONE round-to-nearest conversion, so the table is bit-identical.
namespace isg {
constexpr int kAnchorLevels = 8, kAnchorsPerCell = 16;
struct AnchorParams {
  int n_levels, per_cell, H, W;
  int stride[kAnchorLevels], nx[kAnchorLevels], ny[kAnchorLevels];
  long long first[kAnchorLevels + 1];                       // first anchor index of each level
  double half_x[kAnchorLevels][kAnchorsPerCell], half_y[kAnchorLevels][kAnchorsPerCell];
};

template <typename T> struct AnchorOut;
template <> struct AnchorOut<float> {
  static __device__ __forceinline__ void store(float* out, long long a, double y1, double x1, double y2, double x2) {
    reinterpret_cast<float4*>(out)[a] = make_float4(__double2float_rn(y1), __double2float_rn(x1), __double2float_rn(y2),
                                                    __double2float_rn(x2));
  }
};
template <> struct AnchorOut<__half> {
  static __device__ __forceinline__ void store(__half* out, long long a, double y1, double x1, double y2, double x2) {
    out[4 * a + 0] = __double2half(y1); out[4 * a + 1] = __double2half(x1);
    out[4 * a + 2] = __double2half(y2); out[4 * a + 3] = __double2half(x2);
  }
};

template <typename T>
__global__ void __launch_bounds__(256) anchors_kernel(const __grid_constant__ AnchorParams p, T* __restrict__ out) {
  const long long a = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= p.first[p.n_levels]) return;
  int l = 0;
  while (l + 1 < p.n_levels && a >= p.first[l + 1]) ++l;
  const long long r = a - p.first[l];
  const int k = (int)(r % p.per_cell);                      // (scale, ratio) pair in itertools.product order (:420)
  const long long cell = r / p.per_cell;                    // meshgrid(x, y).reshape(-1): row-major over (y, x) (:429-431)
  const int ix = (int)(cell % p.nx[l]), iy = (int)(cell / p.nx[l]);
  const double s = (double)p.stride[l];
  const double xc = s / 2 + ix * s, yc = s / 2 + iy * s;    // np.arange(stride / 2, size, stride) (:427-428)
  AnchorOut<T>::store(out, a, yc - p.half_y[l][k], xc - p.half_x[l][k], yc + p.half_y[l][k], xc + p.half_x[l][k]);   // :434-435
}
}  // namespace isg

extern "C" int64_t isg_anchor_count(int H, int W, const int* strides, int n_levels, int per_cell) {
  if (H <= 0 || W <= 0 || !strides || n_levels <= 0 || n_levels > isg::kAnchorLevels || per_cell <= 0 || per_cell > isg::kAnchorsPerCell)
    return -1;
  int64_t n = 0;
  for (int l = 0; l < n_levels; ++l) {
    const int s = strides[l];
    if (s <= 0) return -1;
    // len(np.arange(s/2, size, s)) = ceil((size - s/2) / s)
    const int64_t nx = (2LL * W - s + 2LL * s - 1) / (2LL * s), ny = (2LL * H - s + 2LL * s - 1) / (2LL * s);
    n += (nx > 0 ? nx : 0) * (ny > 0 ? ny : 0) * per_cell;
  }
  return n;
}

extern "C" int isg_generate_anchors(int H, int W, const int* strides, int n_levels, const double* half_sizes, int per_cell,
                                    int half_precision, void* out, isg_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!out || !half_sizes || isg_anchor_count(H, W, strides, n_levels, per_cell) <= 0) return ISG_EINVAL;
  if (!half_precision && !aligned16(out)) return ISG_EINVAL;
  isg::AnchorParams p = {};
  p.n_levels = n_levels; p.per_cell = per_cell; p.H = H; p.W = W;
  p.first[0] = 0;
  for (int l = 0; l < n_levels; ++l) {
    const int s = strides[l];
    p.stride[l] = s;
    p.nx[l] = (int)std::max<int64_t>(0, (2LL * W - s + 2LL * s - 1) / (2LL * s));
    p.ny[l] = (int)std::max<int64_t>(0, (2LL * H - s + 2LL * s - 1) / (2LL * s));
    p.first[l + 1] = p.first[l] + (long long)p.nx[l] * p.ny[l] * per_cell;
    for (int k = 0; k < per_cell; ++k) {
      p.half_x[l][k] = half_sizes[((size_t)l * per_cell + k) * 2 + 0];
      p.half_y[l][k] = half_sizes[((size_t)l * per_cell + k) * 2 + 1];
    }
  }
  const unsigned blocks = (unsigned)cdiv64(p.first[n_levels], 256);
  if (half_precision) isg::anchors_kernel<__half><<<blocks, 256, 0, stream>>>(p, reinterpret_cast<__half*>(out));
  else isg::anchors_kernel<float><<<blocks, 256, 0, stream>>>(p, reinterpret_cast<float*>(out));
  ISG_LAUNCH_CHECK();
  return ISG_OK;
}

// ---- f4: the inference-time heads of EfficientDecoder (models/efficient.py:508-510,536-541) ---------------------------
// The reference applies three 1x1 convolutions (kp: 1, ae: 4, tan: 2 channels) to the decoder's last feature map and the
// decode then ignores `tan` (utils/decode.py:447).  This kernel computes only the five channels the decode reads, in
// ONE pass over the feature map, straight into the planar fp32 layout isg_assign_dense / isg_topk_threshold consume:
// out[c] = bias[c] + sum_k w[c][k] * x[k], accumulated in channel order with FMAs.  HBM-bound: 4*Cin B/px read, 20 B/px written.
namespace isg {
constexpr int kHeadMaxCin = 64;
struct HeadParams { float w[5][kHeadMaxCin]; float b[5]; };

__global__ void __launch_bounds__(256)
decode_heads_kernel(const float* __restrict__ x, int Cin, long long plane, const __grid_constant__ HeadParams p,
                    float* __restrict__ kp, float* __restrict__ ae) {
  const long long q = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;      // first of 4 consecutive pixels
  const int b = blockIdx.y;
  if (q >= plane) return;
  float4 acc[5];
#pragma unroll
  for (int c = 0; c < 5; ++c) acc[c] = make_float4(p.b[c], p.b[c], p.b[c], p.b[c]);
  const float* xb = x + (long long)b * Cin * plane + q;
  for (int k = 0; k < Cin; ++k) {
    const float4 v = ldg_stream4(xb + (long long)k * plane);
#pragma unroll
    for (int c = 0; c < 5; ++c) {
      const float w = p.w[c][k];
      acc[c].x = fmaf(w, v.x, acc[c].x); acc[c].y = fmaf(w, v.y, acc[c].y);
      acc[c].z = fmaf(w, v.z, acc[c].z); acc[c].w = fmaf(w, v.w, acc[c].w);
    }
  }
  *reinterpret_cast<float4*>(kp + (long long)b * plane + q) = acc[0];
#pragma unroll
  for (int c = 0; c < 4; ++c) *reinterpret_cast<float4*>(ae + ((long long)b * 4 + c) * plane + q) = acc[1 + c];
}
}  // namespace isg

extern "C" int isg_decode_heads(const float* x, int B, int Cin, int H, int W, const float* w_kp, const float* b_kp,
                                const float* w_ae, const float* b_ae, float* kp, float* ae, isg_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!x || !w_kp || !b_kp || !w_ae || !b_ae || !kp || !ae || B <= 0 || Cin <= 0 || H <= 0 || W <= 0 || B > 65535) return ISG_EINVAL;
  if (Cin > isg::kHeadMaxCin) return ISG_EUNSUPPORTED;
  const long long plane = (long long)H * W;
  if (plane % 4 != 0 || !aligned16(x) || !aligned16(kp) || !aligned16(ae)) return ISG_EUNSUPPORTED;
  isg::HeadParams p = {};
  for (int k = 0; k < Cin; ++k) {
    p.w[0][k] = w_kp[k];
    for (int c = 0; c < 4; ++c) p.w[1 + c][k] = w_ae[c * Cin + k];
  }
  p.b[0] = b_kp[0];
  for (int c = 0; c < 4; ++c) p.b[1 + c] = b_ae[c];
  dim3 grid((unsigned)cdiv64(plane / 4, 256), B);
  isg::decode_heads_kernel<<<grid, 256, 0, stream>>>(x, Cin, plane, p, kp, ae);
  ISG_LAUNCH_CHECK();
  return ISG_OK;
}

extern "C" int isg_pack_masks(const uint8_t* dense, int n, int H, int W, uint32_t* bits, isg_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!dense || !bits || n <= 0 || H <= 0 || W <= 0) return ISG_EINVAL;
  const int Wwords = cdiv(W, 32);
  const long long rows = (long long)n * H;
  const long long threads = rows * Wwords * 32;
  pack_masks_kernel<<<(unsigned)cdiv64(threads, 256), 256, 0, stream>>>(dense, rows, W, Wwords, bits);
  ISG_LAUNCH_CHECK();
  return ISG_OK;
}

extern "C" size_t isg_box_nms_workspace_bytes(int B, int cap) {
  if (B <= 0 || cap <= 0) return 0;
  if (cap > ISG_NMS_MAX_BOXES) return (size_t)B * ((nms_large_ws_per_image(cap) + 255) & ~(size_t)255);
  return (size_t)B * nms_ws_per_image(cap);
}

static int run_nms_stages(const float* boxes, const float* scores, const int32_t* cls, const int32_t* tiebreak,
                          const int32_t* count, int B, int cap, double thr, int convention, int32_t* keep,
                          int32_t* n_keep, void* ws, cudaStream_t stream, bool box_mask) {
  const int P = next_pow2(cap);
  const size_t smem = (size_t)P * 12;
  ISG_CUDA(cudaFuncSetAttribute(nms_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  nms_sort_kernel<<<B, kSortThreads, smem, stream>>>(reinterpret_cast<const float4*>(boxes), scores, cls, tiebreak,
                                                     count, cap, P, convention, ws);
  if (box_mask) {
    dim3 grid(kMaskCtasPerImage, B);
    nms_mask_kernel<<<grid, 64, 0, stream>>>(count, cap, thr, convention, ws);
    ISG_CUDA(launch_nms_scan(count, B, cap, ws, keep, n_keep, stream));
  }
  ISG_LAUNCH_CHECK();
  return ISG_OK;
}

extern "C" int isg_box_nms(const float* boxes, const float* scores, const int32_t* cls, const int32_t* tiebreak,
                           const int32_t* count, int B, int cap, double thr, int convention, int32_t* keep,
                           int32_t* n_keep, void* ws, size_t ws_bytes, isg_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!boxes || !scores || !count || !keep || !n_keep || B <= 0 || cap <= 0 || B > 65535) return ISG_EINVAL;
  if (convention != ISG_NMS_PLUS1_LE && convention != ISG_NMS_TV_GT && convention != ISG_NMS_TV_TRICK &&
      convention != ISG_NMS_TV_BATCHED)
    return ISG_EINVAL;
  if (cap > (1 << 24)) return ISG_EUNSUPPORTED;
  if (!aligned16(boxes)) return ISG_EINVAL;
  if (cap > ISG_NMS_MAX_BOXES) {   // no suppression matrix: sort + tile-by-tile resolution
    if (!ws || ws_bytes < isg_box_nms_workspace_bytes(B, cap) || ((uintptr_t)ws & 255)) return ISG_EWORKSPACE;
    return run_nms_large(boxes, scores, cls, tiebreak, count, B, cap, thr, convention, keep, n_keep, ws, stream);
  }
  if (cap <= kSmallMax) {   // fused single-CTA path, no workspace needed
    const int P = next_pow2(cap);
    const int nwP = P / 64 > 0 ? P / 64 : 1;
    const size_t smem = (size_t)P * (16 + 8 + 4 + 4) + (size_t)P * nwP * 8;
    ISG_CUDA(cudaFuncSetAttribute(nms_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ISG_CUDA(launch_pdl(nms_small_kernel, dim3(B), dim3(kSortThreads), smem, stream, reinterpret_cast<const float4*>(boxes), scores,
                        cls, tiebreak, count, cap, P, thr, convention, keep, n_keep, tuning().nms_rounds));
    ISG_LAUNCH_CHECK();
    return ISG_OK;
  }
  if (!ws || ws_bytes < isg_box_nms_workspace_bytes(B, cap) || ((uintptr_t)ws & 255)) return ISG_EWORKSPACE;
  return run_nms_stages(boxes, scores, cls, tiebreak, count, B, cap, thr, convention, keep, n_keep, ws, stream, true);
}

extern "C" size_t isg_mask_nms_workspace_bytes(int n) {
  if (n <= 0) return 0;
  return nms_ws_per_image(n) + mask_extra_bytes(n);
}

extern "C" int isg_mask_nms(const uint32_t* masks, int n, int H, int Wwords, const int32_t* bboxes,
                            const float* scores, const int32_t* cls, double thr, int32_t* keep, int32_t* n_keep,
                            void* ws, size_t ws_bytes, isg_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!masks || !scores || !keep || !n_keep || n <= 0 || H <= 0 || Wwords <= 0) return ISG_EINVAL;
  if (n > ISG_NMS_MAX_BOXES) return ISG_EUNSUPPORTED;
  if (bboxes && !aligned16(bboxes)) return ISG_EINVAL;
  if (!ws || ws_bytes < isg_mask_nms_workspace_bytes(n) || ((uintptr_t)ws & 255)) return ISG_EWORKSPACE;
  // layout: [NmsWs(cap=n)] [area n*4 (16-aligned)] [bbox n*16] [cnt]
  char* extra = (char*)ws + nms_ws_per_image(n);
  int32_t* area = (int32_t*)extra;
  int4* bbox = (int4*)(extra + (((size_t)n * 4 + 15) & ~(size_t)15));
  int32_t* cnt = (int32_t*)((char*)bbox + (size_t)n * 16);
  NmsWs v = nms_ws_view(ws, 0, n);
  const int nw = cdiv(n, 64);
  set_int_kernel<<<1, 1, 0, stream>>>(cnt, n);
  // rank order by (score desc, larger index first on ties) — the greedy loop of utils/nms.py:20-37.
  // sbox is not needed: boxes pointer is only dereferenced for sbox, so hand the sort a dummy view.
  int rc = run_nms_stages(reinterpret_cast<const float*>(bbox), scores, cls, nullptr, cnt, 1, n, thr, ISG_NMS_PLUS1_LE,
                          keep, n_keep, ws, stream, false);
  if (rc) return rc;
  // the HBM-bound area / bounding-box pass does not depend on the order: it starts while the single-CTA sort runs
  ISG_CUDA(launch_pdl(mask_area_kernel, dim3(n), dim3(256), 0, stream, masks, H, Wwords, reinterpret_cast<const int4*>(bboxes), area, bbox));
  ISG_CUDA(cudaMemsetAsync(v.mask, 0, (size_t)n * nw * 8, stream));
  const long long items = (long long)n * cdiv(n, 32);            // (row, block of 32 later ranks) per warp
  const int blocks = (int)std::min<long long>((items * 32 + 255) / 256, 148LL * 16);
  mask_pair_kernel<<<blocks, 256, 0, stream>>>(masks, n, H, Wwords, v.order, v.scls, area, bbox, thr, v.mask, nw);
  ISG_CUDA(launch_nms_scan(cnt, 1, n, ws, keep, n_keep, stream));
  ISG_LAUNCH_CHECK();
  return ISG_OK;
}

extern "C" int isg_mask_pair_counts(const uint32_t* masks, int n, int H, int Wwords, const int32_t* pairs,
                                    int n_pairs, int64_t* inter_union, isg_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!masks || !pairs || !inter_union || n <= 0 || H <= 0 || Wwords <= 0 || n_pairs <= 0) return ISG_EINVAL;
  mask_pair_counts_kernel<<<n_pairs, 256, 0, stream>>>(masks, H, Wwords, pairs, reinterpret_cast<long long*>(inter_union));
  ISG_LAUNCH_CHECK();
  return ISG_OK;
}
