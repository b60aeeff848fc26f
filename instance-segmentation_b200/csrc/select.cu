// K2 — boundary-keypoint selection: exact k-th largest (3-pass radix select), 3x3 peak test,
// row-major compaction.  Reference: select_points / nms_hm (utils/decode.py:42-48,71-85) and
// kp_mask.nonzero() (utils/decode.py:312).
#include <cooperative_groups.h>
#include <cstdlib>
#include <algorithm>
#include "keep.cuh"

namespace cg = cooperative_groups;

namespace isg {

constexpr int kHistBins = 2048;          // 11-bit digits (11 + 11 + 10 = 32)
constexpr int kHistThreads = 256;
constexpr int kHistPxPerBlock = 16384;   // 16 float4 per thread

__device__ __forceinline__ uint32_t digit_of(uint32_t key, int pass) {
  return pass == 0 ? (key >> 21) : pass == 1 ? ((key >> 10) & 0x7ffu) : (key & 0x3ffu);
}

// Block-wide: find the digit d with  #(digit > d) < krem <= #(digit >= d)  in a 2048-bin histogram;
// returns d and the rank remaining inside that bin.  The first 256 threads own 8 bins each; any further
// threads of the block only take part in the barriers.
__device__ void resolve_digit(const uint32_t* hist, uint32_t krem, uint32_t* out_digit, uint32_t* out_krem) {
  __shared__ uint32_t warp_tot[kHistThreads / 32];
  __shared__ uint32_t res[2];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const bool own = t < kHistThreads;
  uint32_t h[8];
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) { h[i] = own ? hist[t * 8 + i] : 0u; s += h[i]; }
  // inclusive suffix sum over threads (thread t gets sum over t' >= t)
  uint32_t suf = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t v = __shfl_down_sync(0xffffffffu, suf, o);
    if (lane + o < 32) suf += v;
  }
  if (own && lane == 0) warp_tot[warp] = suf;
  if (t == 0) { res[0] = 0; res[1] = 0; }
  __syncthreads();
  if (own) {
    uint32_t above = 0;
    for (int w = warp + 1; w < kHistThreads / 32; ++w) above += warp_tot[w];
    uint32_t cum = suf - s + above;  // count of keys in bins owned by higher threads
#pragma unroll
    for (int i = 7; i >= 0; --i) {
      if (cum < krem && krem <= cum + h[i]) { res[0] = (uint32_t)(t * 8 + i); res[1] = krem - cum; }
      cum += h[i];
    }
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
                 uint32_t* __restrict__ hist_all, int64_t hist_stride, bool vec) {
  __shared__ uint32_t sh[kHistBins];
  const int b = blockIdx.y;
  uint32_t* hist = hist_all + (size_t)b * hist_stride;
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
finish_threshold_kernel(const uint32_t* __restrict__ hist_all, int64_t hist_stride, uint32_t k,
                        uint32_t* __restrict__ thr_key) {
  const int b = blockIdx.x;
  const uint32_t* hist = hist_all + (size_t)b * hist_stride;
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

// ---- fast path: sample -> filter -> select ----------------------------------------------------
// Only ~k of the H*W pixels can matter.  (1) one CTA per image radix-selects, on a 1/64 sample, a
// conservative lower bound L (the sample rank that corresponds to ~2k pixels); (2) one full pass copies the
// keys >= L into a per-image candidate list (block-aggregated appends); (3) one CTA per image selects the
// exact k-th largest key among the candidates.  If the candidate list does not provably contain the answer
// (fewer than k candidates, or overflow) step (3) falls back to scanning the whole image, so the result
// is always exact.
constexpr int kSelThreads = 1024;
constexpr int kSampleMax = 32768;        // sample keys kept in shared memory (128 KB)
#ifndef ISG_FILTER_THREADS
#define ISG_FILTER_THREADS 256
#endif
#ifndef ISG_FILTER_UNROLL
#define ISG_FILTER_UNROLL 4
#endif
constexpr int kFilterThreads = ISG_FILTER_THREADS;
constexpr int kFilterUnroll = ISG_FILTER_UNROLL;                         // independent 128-bit loads in flight per thread
constexpr int kFilterPxPerBlock = kFilterThreads * 4 * kFilterUnroll;    // the block's candidates always fit in shared memory

struct TopkWs {            // per-image views
  uint32_t* hist;          // [3][2048]   (legacy multi-CTA path)
  uint32_t* lower;         // [8]: 0 lower bound, 2 rank inside the bin / 3 bin-mode flag (two-level radix select),
                           //      4 "the candidate list holds every selected pixel" (written by topk_select_kernel)
  uint32_t* ncand;         // [1]
  uint32_t* cand;          // [cap_c] candidate keys
  uint32_t* cand_pos;      // [cap_c] their pixel indices (y * W + x): isg_keep_from_candidates
};
__host__ __device__ inline size_t topk_cand_cap(int npx, int k) {
  size_t c = (size_t)8 * (size_t)k + 8192;
  return c < (size_t)npx ? c : (size_t)npx;
}
__host__ __device__ inline size_t topk_ws_per_image(int npx, int k) {
  size_t s = 3 * kHistBins * sizeof(uint32_t) + 64 + 2 * topk_cand_cap(npx, k) * sizeof(uint32_t);
  return (s + 255) & ~(size_t)255;
}
__host__ __device__ inline TopkWs topk_ws_view(void* ws, int b, int npx, int k) {
  char* p = (char*)ws + (size_t)b * topk_ws_per_image(npx, k);
  TopkWs v;
  v.hist = (uint32_t*)p; p += 3 * kHistBins * sizeof(uint32_t);
  v.lower = (uint32_t*)p; v.ncand = (uint32_t*)(p + 32); p += 64;
  v.cand = (uint32_t*)p;
  v.cand_pos = v.cand + topk_cand_cap(npx, k);
  return v;
}

// ---- alternative fast path (ISG_TOPK_PATH=radix): two-level radix select with a 15-bit first digit ---------------
// (1) topk_hist15_kernel: ONE pass over the batch builds, per image, the histogram of the top 15 key bits (sign,
//     exponent, 6 mantissa bits) - per-CTA in shared memory (32768 x u32 = 128 KB), merged into global memory with one
//     atomic per non-empty bin.  (2) topk_pick15_kernel finds the bin that holds the k-th largest key and the rank
//     that is left inside it.  (3) topk_filter_kernel (bin mode) copies the keys of that one bin - a few thousand - to
//     the candidate list.  (4) topk_select_kernel picks the exact key among them (bitonic sort in one CTA for up to
//     4096 candidates, cluster radix select above; whole-image select if the bin overflowed the list).  Exact for any
//     input; no sampling, no probabilistic bound.
constexpr int kBins15 = 32768;
constexpr int kH15Threads = 1024;
constexpr int kH15Chunk = kH15Threads * 16;     // pixels per CTA iteration: 4 x float4 per thread
__host__ __device__ inline size_t topk_hist15_bytes(int B) { return (size_t)B * kBins15 * sizeof(uint32_t); }

#ifndef ISG_SEL_CLUSTER
#define ISG_SEL_CLUSTER 8
#endif
constexpr int kSelCluster = ISG_SEL_CLUSTER;   // CTAs per image in the sample / select kernels (one thread-block cluster)

// histogram increment aggregated with match.any: one shared-memory atomic per distinct bin per warp
__device__ __forceinline__ void hist_add_match(uint32_t* sh, uint32_t bin, bool valid, int lane) {
  const unsigned peers = __match_any_sync(0xffffffffu, valid ? bin : 0xffffffffu);
  if (valid && lane == __ffs(peers) - 1) atomicAdd(&sh[bin], (uint32_t)__popc(peers));
}

// resolve_digit over the SUM of the cluster's per-CTA histograms (read through distributed shared memory)
__device__ void resolve_digit_cluster(cg::cluster_group& cluster, uint32_t* sh_hist, uint32_t krem, uint32_t* out_digit,
                                      uint32_t* out_krem) {
  __shared__ uint32_t warp_tot[kHistThreads / 32];
  __shared__ uint32_t res[2];
  __shared__ __align__(16) uint32_t tot[kHistBins];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const bool own = t < kHistThreads;
  // cluster-wide totals: all 1024 threads pull two bins each from every peer (independent remote loads in flight),
  // instead of 256 threads pulling eight bins each - the remote-read latency was the larger part of a pass
  static_assert(kSelThreads * 2 == kHistBins, "two bins per thread");
  {
    uint2 acc = make_uint2(0u, 0u);
    uint2 v[kSelCluster];
#pragma unroll
    for (int r = 0; r < kSelCluster; ++r)
      v[r] = (r < (int)cluster.num_blocks()) ? reinterpret_cast<const uint2*>(cluster.map_shared_rank(sh_hist, r))[t] : make_uint2(0u, 0u);
#pragma unroll
    for (int r = 0; r < kSelCluster; ++r) { acc.x += v[r].x; acc.y += v[r].y; }
    reinterpret_cast<uint2*>(tot)[t] = acc;
  }
  __syncthreads();
  uint32_t h[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (own) {
    const uint4 a = reinterpret_cast<const uint4*>(tot)[t * 2], c = reinterpret_cast<const uint4*>(tot)[t * 2 + 1];
    h[0] = a.x; h[1] = a.y; h[2] = a.z; h[3] = a.w; h[4] = c.x; h[5] = c.y; h[6] = c.z; h[7] = c.w;
  }
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += h[i];
  uint32_t suf = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t v = __shfl_down_sync(0xffffffffu, suf, o);
    if (lane + o < 32) suf += v;
  }
  if (own && lane == 0) warp_tot[warp] = suf;
  if (t == 0) { res[0] = 0; res[1] = 0; }
  __syncthreads();
  if (own) {
    uint32_t above = 0;
    for (int w = warp + 1; w < kHistThreads / 32; ++w) above += warp_tot[w];
    uint32_t cum = suf - s + above;
#pragma unroll
    for (int i = 7; i >= 0; --i) {
      if (cum < krem && krem <= cum + h[i]) { res[0] = (uint32_t)(t * 8 + i); res[1] = krem - cum; }
      cum += h[i];
    }
  }
  __syncthreads();
  *out_digit = res[0];
  *out_krem = res[1];
  __syncthreads();
}

// Cluster-wide exact select of the `rank`-th largest of n keys produced by key_at(i).  Every CTA of the
// cluster (1024 threads each) histograms a contiguous slice; the per-CTA histograms are summed through
// DSMEM, so all CTAs resolve the same digit.  sh_hist: this CTA's 2048-bin histogram (shared memory).
template <typename KeyAt>
__device__ uint32_t cluster_radix_select(cg::cluster_group& cluster, uint32_t* sh_hist2, int n, uint32_t rank, KeyAt key_at) {
  // sh_hist2: TWO 2048-bin histograms; pass p uses buffer p & 1, so a CTA may clear the buffer of the next pass while
  // its peers still read the current one - one cluster barrier per pass instead of two, plus one before returning
  // (a CTA must not exit while a peer can still read its shared memory).
  const int t = threadIdx.x, lane = t & 31;
  const int nb = (int)cluster.num_blocks(), r = (int)cluster.block_rank();
  const int chunk = (n + nb - 1) / nb;
  const int lo = min(r * chunk, n), hi = min(lo + chunk, n);
  uint32_t prefix = 0, pmask = 0, krem = rank;
#pragma unroll 1
  for (int pass = 0; pass < 3; ++pass) {
    uint32_t* sh_hist = sh_hist2 + (pass & 1) * kHistBins;
    for (int i = t; i < kHistBins; i += kSelThreads) sh_hist[i] = 0;
    __syncthreads();
    for (int i0 = lo; i0 < hi; i0 += kSelThreads * 4) {   // warp-uniform trip count, 4 independent loads in flight
      uint32_t key[4];
      bool in[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { const int i = i0 + u * kSelThreads + t; in[u] = i < hi; key[u] = in[u] ? key_at(i) : 0u; }
      // pass 0: the keys share a handful of leading digits -> aggregate per warp (match.any); later passes: the digits
      // of the surviving keys are spread out, a plain atomic per key is cheaper than the match
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const bool ok = in[u] && (key[u] & pmask) == prefix;
        if (pass == 0) hist_add_match(sh_hist, digit_of(key[u], pass), ok, lane);
        else if (ok) atomicAdd(&sh_hist[digit_of(key[u], pass)], 1u);
      }
    }
    cluster.sync();   // every CTA's histogram of this pass is complete and visible
    uint32_t d, k2;
    resolve_digit_cluster(cluster, sh_hist, krem, &d, &k2);
    krem = k2;
    if (pass == 0) { prefix = d << 21; pmask = 0xffe00000u; }
    else if (pass == 1) { prefix |= d << 10; pmask = 0xfffffc00u; }
    else prefix |= d;
  }
  cluster.sync();     // nobody still reads this CTA's histograms
  return prefix;
}

__global__ void __cluster_dims__(kSelCluster, 1, 1) __launch_bounds__(kSelThreads)
topk_sample_kernel(const float* __restrict__ kp, int64_t img_stride, int npx, int k, int stride, void* ws) {
  extern __shared__ uint32_t skeys[];   // this CTA's slice of the sample, [ceil(S / cluster)]
  __shared__ uint32_t sh_hist[2 * kHistBins];
  cg::cluster_group cluster = cg::this_cluster();
  const int b = blockIdx.y, t = threadIdx.x;
  const int r = (int)cluster.block_rank(), nb = (int)cluster.num_blocks();
  const float* img = kp + (int64_t)b * img_stride;
  const int S = npx / stride;           // sample i reads pixel i*stride + (i*37 % stride)
  const int chunk = (S + nb - 1) / nb;
  const int lo = min(r * chunk, S), hi = min(lo + chunk, S);
  for (int i = lo + t; i < hi; i += kSelThreads) {
    const int p = i * stride + (int)(((unsigned)i * 37u) % (unsigned)stride);
    skeys[i - lo] = float_key(__ldg(img + p));
  }
  __syncthreads();
  // sample rank that corresponds to ~1.2k pixels, plus 6 sigma and a constant: the number of pixels above the
  // sample value of that rank then exceeds k with a margin of ~9 sigma (and step 3 falls back if it ever did not)
  const double expect = 1.2 * (double)k * (double)S / (double)npx;
  long long rk = (long long)(expect + 6.0 * sqrt(expect) + 8.0);
  TopkWs v = topk_ws_view(ws, b, npx, k);
  uint32_t lower = 0u;                  // 0: every pixel is a candidate (step 3 then falls back if needed)
  if (rk < (long long)S)                // cluster-uniform
    lower = cluster_radix_select(cluster, sh_hist, S, (uint32_t)rk, [&](int i) { return skeys[i - lo]; });
  if (r == 0 && t == 0) { *v.lower = lower; v.lower[3] = 0u; *v.ncand = 0u; }
}

// Single-CTA form of the sample step: at most kSample1Max samples per image, radix-selected with block barriers
// only (no cluster synchronisation, no distributed shared memory) - the sample is small, so latency is what counts.
constexpr int kSample1Max = 8192;

__global__ void __launch_bounds__(kSelThreads)
topk_sample1_kernel(const float* __restrict__ kp, int64_t img_stride, int npx, int k, int stride, void* ws) {
  __shared__ uint32_t skeys[kSample1Max];
  __shared__ uint32_t sh_hist[kHistBins];
  pdl_trigger();                       // the filter kernel may be scheduled; it waits for this grid before reading `lower`
  const int b = blockIdx.x, t = threadIdx.x, lane = t & 31;
  const float* img = kp + (int64_t)b * img_stride;
  const int S = min(npx / stride, kSample1Max);   // sample i reads pixel i*stride + (i*37 % stride)
  {
    // the sample loads are scattered (one DRAM sector each): issue all of a thread's loads before using any
    constexpr int kPer = kSample1Max / kSelThreads;
    float x[kPer];
#pragma unroll
    for (int u = 0; u < kPer; ++u) {
      const int i = t + u * kSelThreads;
      const int p = i * stride + (int)(((unsigned)i * 37u) % (unsigned)stride);
      x[u] = (i < S) ? __ldg(img + p) : 0.0f;
    }
#pragma unroll
    for (int u = 0; u < kPer; ++u) { const int i = t + u * kSelThreads; if (i < S) skeys[i] = float_key(x[u]); }
  }
  __syncthreads();
  // same bound as topk_sample_kernel: sample rank of ~1.2k pixels plus 6 sigma and a constant
  const double expect = 1.2 * (double)k * (double)S / (double)npx;
  const long long rk = (long long)(expect + 6.0 * sqrt(expect) + 8.0);
  TopkWs v = topk_ws_view(ws, b, npx, k);
  uint32_t lower = 0u;                  // 0: every pixel is a candidate (the select step then falls back if needed)
  if (rk < (long long)S) {              // block-uniform
    // Two radix passes (22 of the 32 key bits): the bound only has to be conservative, and the start of the 10-bit bin that
    // holds the sample's rk-th key is - it admits at most the few extra candidates of that bin.
    uint32_t prefix = 0, pmask = 0, krem = (uint32_t)rk;
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
      for (int i = t; i < kHistBins; i += kSelThreads) sh_hist[i] = 0;
      __syncthreads();
      for (int i0 = 0; i0 < S; i0 += kSelThreads) {          // warp-uniform trip count
        const int i = i0 + t;
        const uint32_t key = i < S ? skeys[i] : 0u;
        const bool valid = i < S && (key & pmask) == prefix;
        if (!__any_sync(0xffffffffu, valid)) continue;       // later passes: most warps hold no key of the prefix
        if (pass == 0) hist_add_match(sh_hist, digit_of(key, pass), valid, lane);   // concentrated digits: aggregate
        else if (valid) atomicAdd(&sh_hist[digit_of(key, pass)], 1u);               // spread digits: plain atomics
      }
      __syncthreads();
      uint32_t d, k2;
      resolve_digit(sh_hist, krem, &d, &k2);
      krem = k2;
      if (pass == 0) { prefix = d << 21; pmask = 0xffe00000u; }
      else prefix |= d << 10;
    }
    lower = prefix;
  }
  if (t == 0) { *v.lower = lower; v.lower[3] = 0u; *v.ncand = 0u; }
}

__global__ void __launch_bounds__(kH15Threads, 1)
topk_hist15_kernel(const float* __restrict__ kp, int64_t img_stride, int npx, int B, uint32_t* __restrict__ hist15, bool vec) {
  extern __shared__ uint32_t sh15[];             // [kBins15]
  const int t = threadIdx.x;
  const int cpi = (npx + kH15Chunk - 1) / kH15Chunk;                 // chunks per image
  const long long total = (long long)B * cpi;
  const long long c_begin = total * blockIdx.x / gridDim.x, c_end = total * (blockIdx.x + 1) / gridDim.x;
  int cur_b = -1;
  auto flush = [&](int b) {
    __syncthreads();
    uint32_t* g = hist15 + (size_t)b * kBins15;
    for (int i = t; i < kBins15; i += kH15Threads) { const uint32_t c = sh15[i]; if (c) atomicAdd(&g[i], c); }
    __syncthreads();
  };
  auto add = [&](float x, bool in) {
    const uint32_t d = float_key(x) >> 17;
    // a warp whose pixels all fall into one bin (flat regions) adds once instead of serialising 32 ways
    const uint32_t d0 = __shfl_sync(0xffffffffu, d, 0);
    const unsigned inm = __ballot_sync(0xffffffffu, in);
    if (__all_sync(0xffffffffu, !in || d == d0)) { if ((t & 31) == 0 && inm) atomicAdd(&sh15[d0], (uint32_t)__popc(inm)); }
    else if (in) atomicAdd(&sh15[d], 1u);
  };
  for (long long c = c_begin; c < c_end; ++c) {
    const int b = (int)(c / cpi);
    if (b != cur_b) {
      if (cur_b >= 0) flush(cur_b);
      for (int i = t; i < kBins15; i += kH15Threads) sh15[i] = 0;
      __syncthreads();
      cur_b = b;
    }
    const float* img = kp + (int64_t)b * img_stride;
    const int base = (int)(c - (long long)b * cpi) * kH15Chunk;
    if (vec) {
      float4 q[4];
      bool in[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int p = base + u * kH15Threads * 4 + t * 4;
        in[u] = p < npx;                                              // npx % 4 == 0 on this path
        q[u] = in[u] ? ldg_stream4(img + p) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) { add(q[u].x, in[u]); add(q[u].y, in[u]); add(q[u].z, in[u]); add(q[u].w, in[u]); }
    } else {
#pragma unroll 4
      for (int u = 0; u < 16; ++u) {
        const int p = base + u * kH15Threads + t;
        const bool in = p < npx;
        add(in ? __ldg(img + p) : 0.0f, in);
      }
    }
  }
  if (cur_b >= 0) flush(cur_b);
}

// one CTA per image: the 15-bit bin that holds the k-th largest key, and the rank left inside the bin
__global__ void __launch_bounds__(1024)
topk_pick15_kernel(const uint32_t* __restrict__ hist15, int npx, int k, void* ws) {
  __shared__ uint32_t warp_tot[32];
  __shared__ uint32_t res[2];
  const int b = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const uint32_t* hist = hist15 + (size_t)b * kBins15;
  uint32_t h[32];
  uint32_t sum = 0;
  const uint4* src = reinterpret_cast<const uint4*>(hist + t * 32);
#pragma unroll
  for (int i = 0; i < 8; ++i) { const uint4 v = src[i]; h[4 * i] = v.x; h[4 * i + 1] = v.y; h[4 * i + 2] = v.z; h[4 * i + 3] = v.w; }
#pragma unroll
  for (int i = 0; i < 32; ++i) sum += h[i];
  uint32_t suf = sum;                                       // inclusive suffix sum over threads (larger bins = larger keys)
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_down_sync(0xffffffffu, suf, o); if (lane + o < 32) suf += v; }
  if (lane == 0) warp_tot[warp] = suf;
  if (t == 0) { res[0] = 0; res[1] = 0; }
  __syncthreads();
  uint32_t above = 0;
  for (int w = warp + 1; w < 32; ++w) above += warp_tot[w];
  uint32_t cum = suf - sum + above;                         // keys in bins owned by higher threads
  const uint32_t kk = (uint32_t)k;
#pragma unroll
  for (int i = 31; i >= 0; --i) {
    if (cum < kk && kk <= cum + h[i]) { res[0] = (uint32_t)(t * 32 + i); res[1] = kk - cum; }
    cum += h[i];
  }
  __syncthreads();
  if (t == 0) {
    TopkWs v = topk_ws_view(ws, b, npx, k);
    v.lower[0] = res[0] << 17;      // smallest key of the bin
    v.lower[2] = res[1];            // rank of the answer among the keys of the bin
    v.lower[3] = 1u;                // bin mode (topk_filter_kernel / topk_select_kernel)
    *v.ncand = 0u;
  }
}

// POS: also record the candidates' pixel indices (isg_topk_keep evaluates the peak test at them)
template <bool POS>
__global__ void __launch_bounds__(kFilterThreads)
topk_filter_kernel(const float* __restrict__ kp, int64_t img_stride, int npx, int k, void* ws, bool vec) {
  __shared__ uint32_t buf[kFilterPxPerBlock];
  __shared__ uint16_t bufp[POS ? kFilterPxPerBlock : 1];     // pixel index of buf[i] relative to the block's first pixel
  static_assert(kFilterPxPerBlock <= 65536, "block-relative pixel offsets are 16-bit");
  __shared__ uint32_t s_count, s_base;
  pdl_trigger();
  pdl_wait();                          // `lower` / `ncand` come from the kernel launched just before
  const int b = blockIdx.y, t = threadIdx.x, lane = t & 31;
  const TopkWs v = topk_ws_view(ws, b, npx, k);
  const uint32_t lower = *v.lower;
  const bool bin_mode = v.lower[3] != 0u;      // candidates = the keys of ONE 15-bit bin (two-level radix select)
  const uint32_t upper = bin_mode ? (lower | 0x1ffffu) : 0xffffffffu;
  const float* img = kp + (int64_t)b * img_stride;
  const float lower_f = float_from_ukey(lower);
  if (t == 0) s_count = 0;
  __syncthreads();
  const int base = blockIdx.x * kFilterPxPerBlock;
  const int end = min(base + kFilterPxPerBlock, npx);
  // candidate keys of one 4-pixel group -> this block's shared-memory list (warp-aggregated append)
  auto append4 = [&](uint32_t (&key)[4], int c, int pfirst, int pstep) {
    if (__any_sync(0xffffffffu, c > 0)) {
      int inc = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += u;
      }
      uint32_t wbase = 0;
      if (lane == 31) wbase = atomicAdd(&s_count, (uint32_t)inc);
      wbase = __shfl_sync(0xffffffffu, wbase, 31);
      uint32_t o2 = wbase + inc - c;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (key[i] != 0xffffffffu) { buf[o2] = key[i]; if (POS) bufp[o2] = (uint16_t)(pfirst + i * pstep - base); ++o2; }
    }
  };
  // NOTE: 0xffffffff marks "not a candidate"; a real key of 0xffffffff (a NaN payload) is dropped, NaNs are
  // outside the contract.
  if (vec) {
    constexpr int kUnroll = kFilterUnroll;                      // the block's pixels in one go
    for (int p0 = base; p0 < end; p0 += kFilterThreads * 4 * kUnroll) {   // warp-uniform trip count
      float4 q[kUnroll];
      bool in[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int p = p0 + u * kFilterThreads * 4 + t * 4;
        in[u] = p < end;
        q[u] = in[u] ? ldg_stream4(img + p) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      // Candidates are a few percent of the pixels: reject in the float domain first (x < lower_f as floats implies
      // key(x) < lower; -0 / NaN / a NaN bound fail the strict compare and take the exact key test).
#ifndef ISG_FILTER_INLINE_APPEND
      // Main loop without any divergence: one test per 128-bit group (its maximum below the bound rejects all four
      // pixels: 3 FMNMX + 1 compare; fmaxf drops NaN operands, and NaNs are outside the contract) sets a flag bit.  Nearly
      // every warp holds a group with a candidate, but only a lane or two of it: the flagged groups are handled AFTER the
      // loop, where the warp runs the append code once per flagged group of its busiest lane (1-2 times) instead of once
      // per pixel position of every group (the append is ~20 instructions: exact key test, one shared-memory atomic for
      // the group's candidates, up to four stores).
      uint32_t hit = 0u;
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const float m4 = fmaxf(fmaxf(q[u].x, q[u].y), fmaxf(q[u].z, q[u].w));
        hit |= (in[u] && !(m4 < lower_f)) ? (1u << u) : 0u;
      }
      while (hit) {
        const int u = __ffs(hit) - 1;
        hit &= hit - 1u;
        float4 g = q[0];                                    // register select (a dynamic index would go to local memory)
#pragma unroll
        for (int w = 1; w < kUnroll; ++w) if (u == w) g = q[w];
        const float x4[4] = {g.x, g.y, g.z, g.w};
        uint32_t kk[4];
        int c = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          kk[i] = float_key(x4[i]);
          const bool ok = !(x4[i] < lower_f) && kk[i] >= lower && kk[i] <= upper && kk[i] != 0xffffffffu;
          if (!ok) kk[i] = 0xffffffffu;
          c += ok ? 1 : 0;
        }
        if (c) {
          uint32_t o2 = atomicAdd(&s_count, (uint32_t)c);
          const uint32_t pg = (uint32_t)(p0 - base + u * kFilterThreads * 4 + t * 4);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (kk[i] != 0xffffffffu) { buf[o2] = kk[i]; if (POS) bufp[o2] = (uint16_t)(pg + i); ++o2; }
        }
      }
#else
      // variant kept for A/B measurements: the survivors append themselves inside the loop, one by one
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        // one test per 128-bit group first: its maximum below the bound rejects all four pixels (3 FMNMX + 1 compare
        // instead of four compare-and-branch pairs; fmaxf drops NaN operands, and NaNs are outside the contract)
        const float m4 = fmaxf(fmaxf(q[u].x, q[u].y), fmaxf(q[u].z, q[u].w));
        if (!in[u] || m4 < lower_f) continue;
        const float x4[4] = {q[u].x, q[u].y, q[u].z, q[u].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (!(x4[i] < lower_f)) {
            const uint32_t kk = float_key(x4[i]);
            if (kk >= lower && kk <= upper && kk != 0xffffffffu) {
              const uint32_t o2 = atomicAdd(&s_count, 1u);
              buf[o2] = kk; if (POS) bufp[o2] = (uint16_t)(p0 - base + u * kFilterThreads * 4 + t * 4 + i);
            }
          }
        }
      }
#endif
    }
  } else {
    for (int p0 = base; p0 < end; p0 += kFilterThreads * 4) {   // warp-uniform trip count
      uint32_t key[4];
      int c = 0;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int p = p0 + i * kFilterThreads + t;
        key[i] = 0xffffffffu;
        if (p < end) { const uint32_t kk = float_key(__ldg(img + p)); if (kk >= lower && kk <= upper) { key[i] = kk; ++c; } }
      }
      append4(key, c, p0 + t, kFilterThreads);
    }
  }
  __syncthreads();
  const uint32_t n = s_count;
  if (n == 0) return;
  if (t == 0) s_base = atomicAdd(v.ncand, n);
  __syncthreads();
  const size_t capc = topk_cand_cap(npx, k);
  const uint32_t g = s_base;
  for (uint32_t i = t; i < n; i += kFilterThreads)
    if ((size_t)g + i < capc) { v.cand[g + i] = buf[i]; if (POS) v.cand_pos[g + i] = (uint32_t)base + bufp[i]; }
}

// ---- keep bits from the top-k candidate list -------------------------------------------------
// The filter pass of the top-k threshold already visited every pixel; its candidate list (keys >= a lower bound, with
// their pixel indices) holds every selected pixel whenever the select step could use it.  The 3x3 peak test (keep.cuh)
// then only has to be evaluated at the ~k selected candidates instead of streaming the whole map a second time: eight
// scattered neighbour loads per selected pixel, one atomicOr per keep pixel into the zeroed bit plane.  The select kernel
// does this itself as soon as it knows the threshold (isg_topk_keep); when the list is not complete (overflow, fall-back
// select, bin mode) it walks the whole image instead.
__device__ __forceinline__ void keep_test_and_set(const float* __restrict__ img, int p, float c, int H, int W, int Wwords,
                                                  uint32_t thr, uint32_t* __restrict__ kb) {
  const int y = p / W, x = p - y * W;
  float m = c;                                   // v(c) = c: the pixel is selected
  float nb[8];
  bool in[8];
  int q = 0;
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) {
      if (dy == 0 && dx == 0) continue;
      const int yy = y + dy, xx = x + dx;
      in[q] = yy >= 0 && yy < H && xx >= 0 && xx < W;
      nb[q] = in[q] ? __ldg(img + (int64_t)yy * W + xx) : 0.0f;
      ++q;
    }
#pragma unroll
  for (int i = 0; i < 8; ++i)                    // unselected neighbours count as 0, pixels outside the image not at all
    if (in[i]) m = fmaxf(m, float_key(nb[i]) >= thr ? nb[i] : 0.0f);
  if (c >= m) atomicOr(kb + (size_t)y * Wwords + (x >> 5), 1u << (x & 31));
}

constexpr int kSmallSel = 4096;   // candidates sorted by one CTA in shared memory

__global__ void __cluster_dims__(kSelCluster, 1, 1) __launch_bounds__(kSelThreads)
topk_select_kernel(const float* __restrict__ kp, int64_t img_stride, int npx, int k, void* ws,
                   uint32_t* __restrict__ thr_key, int H, int W, uint32_t* __restrict__ keepbits) {
  __shared__ uint32_t sh_hist[2 * kHistBins];
  cg::cluster_group cluster = cg::this_cluster();
  pdl_wait();                          // the candidate list comes from the filter kernel
  const int b = blockIdx.y;
  const TopkWs v = topk_ws_view(ws, b, npx, k);
  const uint32_t nc = *v.ncand;
  const bool bin_mode = v.lower[3] != 0u;
  const uint32_t rank = bin_mode ? v.lower[2] : (uint32_t)k;      // rank of the answer among the candidates
  const bool use_cand = nc >= rank && rank >= 1u && (size_t)nc <= topk_cand_cap(npx, k);
  // the candidates are all keys >= lower and the answer is one of them: every selected pixel is in the list
  const bool complete = use_cand && !bin_mode;
  if (cluster.block_rank() == 0 && threadIdx.x == 0) v.lower[4] = complete ? 1u : 0u;
  // keep bits (isg_topk_keep; keepbits zeroed by the caller): the 3x3 peak test at the selected pixels, by `nthr` threads
  // of which this one is number `me`
  auto set_keep_bits = [&](uint32_t thr, int me, int nthr) {
    const float* img = kp + (int64_t)b * img_stride;
    const int Wwords = (W + 31) / 32;
    uint32_t* kb = keepbits + (size_t)b * H * Wwords;
    if (complete) {
      // two candidates per thread and round: their list entries, then all their neighbour loads, are in flight together
      // (the loop is a chain of dependent DRAM / L2 round trips otherwise)
      constexpr int kBatch = 2;
      for (int i0 = me; i0 < (int)nc; i0 += nthr * kBatch) {
        uint32_t ck[kBatch];
        int p[kBatch];
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          const int i = i0 + u * nthr;
          ck[u] = i < (int)nc ? v.cand[i] : 0u;
          p[u] = i < (int)nc ? (int)v.cand_pos[i] : 0;
        }
        float nb[kBatch][8];
        uint32_t inm[kBatch];
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          inm[u] = 0u;
          const bool act = i0 + u * nthr < (int)nc && ck[u] >= thr;
          const int y = p[u] / W, x = p[u] - y * W;
          int q = 0;
#pragma unroll
          for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) {
              if (dy == 0 && dx == 0) continue;
              const int yy = y + dy, xx = x + dx;
              const bool in = act && yy >= 0 && yy < H && xx >= 0 && xx < W;
              nb[u][q] = in ? __ldg(img + (int64_t)yy * W + xx) : 0.0f;
              inm[u] |= in ? (1u << q) : 0u;
              ++q;
            }
          if (act) inm[u] |= 0x100u;
        }
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          if (!(inm[u] & 0x100u)) continue;
          const float c = float_from_ukey(ck[u]);
          float m = c;
#pragma unroll
          for (int q = 0; q < 8; ++q)
            if (inm[u] & (1u << q)) m = fmaxf(m, float_key(nb[u][q]) >= thr ? nb[u][q] : 0.0f);
          const int y = p[u] / W, x = p[u] - y * W;
          if (c >= m) atomicOr(kb + (size_t)y * Wwords + (x >> 5), 1u << (x & 31));
        }
      }
    } else {
      for (int p = me; p < npx; p += nthr) {
        const float c = __ldg(img + p);
        if (float_key(c) >= thr) keep_test_and_set(img, p, c, H, W, Wwords, thr, kb);
      }
    }
  };
  uint32_t key;
  if (use_cand && nc <= (uint32_t)kSmallSel) {   // cluster-uniform: one CTA sorts the few candidates in shared memory
    if (cluster.block_rank() != 0) return;
    __shared__ uint32_t sk[kSmallSel];
    const int t = threadIdx.x;
    int P = 1;
    while (P < (int)nc) P <<= 1;
    for (int i = t; i < P; i += kSelThreads) sk[i] = i < (int)nc ? v.cand[i] : 0u;      // 0 sorts last (descending)
    __syncthreads();
    for (int size = 2; size <= P; size <<= 1) {
      for (int stride = size >> 1; stride > 0; stride >>= 1) {
        for (int i = t; i < (P >> 1); i += kSelThreads) {
          const int lo = 2 * i - (i & (stride - 1)), hi = lo + stride;
          const bool desc = ((lo & size) == 0);
          const uint32_t a = sk[lo], c = sk[hi];
          if ((a < c) == desc) { sk[lo] = c; sk[hi] = a; }
        }
        __syncthreads();
      }
    }
    const uint32_t thr = sk[rank - 1];
    if (t == 0) thr_key[b] = thr;
    if (keepbits) set_keep_bits(thr, t, kSelThreads);
    return;
  }
  if (use_cand) {   // cluster-uniform
    const uint32_t* cand = v.cand;
    key = cluster_radix_select(cluster, sh_hist, (int)nc, rank, [&](int i) { return cand[i]; });
  } else {
    const float* img = kp + (int64_t)b * img_stride;
    key = cluster_radix_select(cluster, sh_hist, npx, (uint32_t)k, [&](int i) { return float_key(__ldg(img + i)); });
  }
  if (cluster.block_rank() == 0 && threadIdx.x == 0) thr_key[b] = key;
  if (keepbits) set_keep_bits(key, (int)cluster.block_rank() * kSelThreads + (int)threadIdx.x, kSelCluster * kSelThreads);
}

// Single-CTA form of the select step (default).  The cluster form above spreads ~25 k candidates per image over 8 CTAs of
// 1024 threads: 64 CTAs, each of which needs a whole SM's register file - in a ring of overlapped steps they cannot share
// an SM with the persistent dense kernel of a neighbouring step, which then starts 64 of its 148 CTAs late.  One CTA per
// image does the three radix passes with block barriers only (no DSMEM histogram sums): about the same latency, an eighth
// of the SMs.  Candidates are re-read from global memory in every pass (100 KB per image, L2-resident).  The whole-image
// fall-back (incomplete candidate list) runs the same code over all pixels: slow, and only for degenerate inputs.
template <typename KeyAt>
__device__ uint32_t block_radix_select(uint32_t* sh_hist, int n, uint32_t rank, KeyAt key_at) {
  const int t = threadIdx.x, lane = t & 31;
  uint32_t prefix = 0, pmask = 0, krem = rank;
#pragma unroll 1
  for (int pass = 0; pass < 3; ++pass) {
    for (int i = t; i < kHistBins; i += kSelThreads) sh_hist[i] = 0;
    __syncthreads();
    for (int i0 = 0; i0 < n; i0 += kSelThreads * 8) {     // warp-uniform trip count, 8 independent loads in flight
      uint32_t key[8];
      bool in[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) { const int i = i0 + u * kSelThreads + t; in[u] = i < n; key[u] = in[u] ? key_at(i) : 0u; }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const bool ok = in[u] && (key[u] & pmask) == prefix;
        if (pass == 0) hist_add_match(sh_hist, digit_of(key[u], pass), ok, lane);
        else if (ok) atomicAdd(&sh_hist[digit_of(key[u], pass)], 1u);
      }
    }
    __syncthreads();
    uint32_t d, k2;
    resolve_digit(sh_hist, krem, &d, &k2);
    krem = k2;
    if (pass == 0) { prefix = d << 21; pmask = 0xffe00000u; }
    else if (pass == 1) { prefix |= d << 10; pmask = 0xfffffc00u; }
    else prefix |= d;
  }
  return prefix;
}

__global__ void __launch_bounds__(kSelThreads)
topk_select1_kernel(const float* __restrict__ kp, int64_t img_stride, int npx, int k, void* ws,
                    uint32_t* __restrict__ thr_key, int H, int W, uint32_t* __restrict__ keepbits) {
  __shared__ uint32_t sh_hist[kHistBins];
  pdl_wait();                          // the candidate list comes from the filter kernel
  const int b = blockIdx.x, t = threadIdx.x;
  const TopkWs v = topk_ws_view(ws, b, npx, k);
  const uint32_t nc = *v.ncand;
  const bool bin_mode = v.lower[3] != 0u;
  const uint32_t rank = bin_mode ? v.lower[2] : (uint32_t)k;
  const bool use_cand = nc >= rank && rank >= 1u && (size_t)nc <= topk_cand_cap(npx, k);
  const bool complete = use_cand && !bin_mode;
  if (t == 0) v.lower[4] = complete ? 1u : 0u;
  const float* img = kp + (int64_t)b * img_stride;
  uint32_t key;
  if (use_cand) {                      // block-uniform
    const uint32_t* cand = v.cand;
    key = block_radix_select(sh_hist, (int)nc, rank, [&](int i) { return cand[i]; });
  } else {
    key = block_radix_select(sh_hist, npx, (uint32_t)k, [&](int i) { return float_key(__ldg(img + i)); });
  }
  if (t == 0) thr_key[b] = key;
  if (!keepbits) return;
  // keep bits (isg_topk_keep): the 3x3 peak test at the selected pixels
  const int Wwords = (W + 31) / 32;
  uint32_t* kb = keepbits + (size_t)b * H * Wwords;
  if (complete) {
    for (int i = t; i < (int)nc; i += kSelThreads) {
      const uint32_t ck = v.cand[i];
      const int p = (int)v.cand_pos[i];
      if (ck >= key) keep_test_and_set(img, p, float_from_ukey(ck), H, W, Wwords, key, kb);
    }
  } else {
    for (int p = t; p < npx; p += kSelThreads) {
      const float c = __ldg(img + p);
      if (float_key(c) >= key) keep_test_and_set(img, p, c, H, W, Wwords, key, kb);
    }
  }
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
  const Thr thr = make_thr(thr_key[b]);
  RowH up, mid;
  if (thr.use_int) { up = load_rowh<VEC, true>(img, ybeg - 1, x0, H, W, thr, lane); mid = load_rowh<VEC, true>(img, ybeg, x0, H, W, thr, lane); }
  else { up = load_rowh<VEC, false>(img, ybeg - 1, x0, H, W, thr, lane); mid = load_rowh<VEC, false>(img, ybeg, x0, H, W, thr, lane); }
#pragma unroll
  for (int r = 0; r < kKeepRowsPerWarp; ++r) {
    const int y = ybeg + r;
    if (y >= H) break;
    const RowH dn = thr.use_int ? load_rowh<VEC, true>(img, y + 1, x0, H, W, thr, lane)
                                : load_rowh<VEC, false>(img, y + 1, x0, H, W, thr, lane);
    const uint32_t nib = keep_nibble(up, mid, dn);
    const uint32_t word = nibbles_to_word(nib, lane);
    if ((lane & 7) == 0 && x0 < W) keepbits[((size_t)b * H + y) * Wwords + (x0 >> 5)] = word;
    if (mask_u8) {
      uint8_t* mrow = mask_u8 + ((size_t)b * H + y) * W;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (x0 + i < W) mrow[x0 + i] = (nib >> i) & 1u;
    }
    up = mid; mid = dn;
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

// ---- ordered compaction: one 8-CTA thread-block cluster per image ----------------------------
// CTA r of the cluster owns a contiguous band of rows; inside it warp w owns rows w, w+32, ...  A row's words are
// read with one coalesced request per 32 words, eight rows in flight per warp.  Phase 1 counts the set bits per
// row, a block scan gives row offsets inside the band, the band totals are exchanged through distributed shared
// memory (exclusive prefix over the lower-ranked CTAs), and phase 2 re-reads the rows (L2 hits) and emits the
// (y,x) pairs in row-major order.
constexpr int kCompactThreads = 1024;
constexpr int kCompactBatch = 8;     // rows in flight per warp
constexpr int kCompactCluster = 8;   // CTAs per image

// WL = words per lane (Wwords <= 32*WL); WL == 0 selects the generic any-width path
template <int WL>
__global__ void __cluster_dims__(kCompactCluster, 1, 1) __launch_bounds__(kCompactThreads, 1)
compact_kernel(const uint32_t* __restrict__ keepbits, int H, int Wwords, int W, int cap,
               int32_t* __restrict__ idx, int32_t* __restrict__ count) {
  extern __shared__ int row_off[];   // [band rows + 1]
  __shared__ int warp_tot[32];
  __shared__ int s_total;
  cg::cluster_group cluster = cg::this_cluster();
  const int b = blockIdx.y;
  const int rank = (int)cluster.block_rank();
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const uint32_t* bits = keepbits + (size_t)b * H * Wwords;
  constexpr int WLc = WL > 0 ? WL : 1;
  const int Hc = (H + kCompactCluster - 1) / kCompactCluster;
  const int yb0 = min(rank * Hc, H), yb1 = min(yb0 + Hc, H);   // this CTA's band [yb0, yb1)
  const int nrows = yb1 - yb0;

  // phase 1: per-row popcounts
  if (WL > 0) {
    for (int yr = warp; yr < nrows; yr += 32 * kCompactBatch) {
      uint32_t v[kCompactBatch][WLc];
#pragma unroll
      for (int r = 0; r < kCompactBatch; ++r)
#pragma unroll
        for (int j = 0; j < WLc; ++j) {
          const int y = yb0 + yr + 32 * r, w = lane + 32 * j;
          v[r][j] = (yr + 32 * r < nrows && w < Wwords) ? __ldg(bits + (size_t)y * Wwords + w) : 0u;
        }
#pragma unroll
      for (int r = 0; r < kCompactBatch; ++r) {
        int c = 0;
#pragma unroll
        for (int j = 0; j < WLc; ++j) c += __popc(v[r][j]);
        c = warp_sum(c);
        if (lane == 0 && yr + 32 * r < nrows) row_off[yr + 32 * r] = c;
      }
    }
  } else {
    for (int yr = warp; yr < nrows; yr += 32) {
      int c = 0;
      for (int w = lane; w < Wwords; w += 32) c += __popc(__ldg(bits + (size_t)(yb0 + yr) * Wwords + w));
      c = warp_sum(c);
      if (lane == 0) row_off[yr] = c;
    }
  }
  __syncthreads();

  // block exclusive scan over the band's row counts (thread t owns a contiguous run of rows)
  const int per = (nrows + kCompactThreads - 1) / kCompactThreads;
  const int y0 = min(t * per, nrows), y1 = min(y0 + per, nrows);
  int c = 0;
  for (int y = y0; y < y1; ++y) c += row_off[y];
  int inc = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += v;
  }
  if (lane == 31) warp_tot[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    const int v = warp_tot[lane];
    int s2 = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, s2, o);
      if (lane >= o) s2 += u;
    }
    warp_tot[lane] = s2 - v;
    if (lane == 31) s_total = s2;
  }
  __syncthreads();
  int run = warp_tot[warp] + inc - c;
  for (int y = y0; y < y1; ++y) { const int v = row_off[y]; row_off[y] = run; run += v; }

  // exchange the band totals: offset of this band = sum of the totals of the lower-ranked CTAs
  cluster.sync();
  int band_base = 0, grand = 0;
  for (int r = 0; r < kCompactCluster; ++r) {
    const int v = *cluster.map_shared_rank(&s_total, r);
    if (r < rank) band_base += v;
    grand += v;
  }
  if (rank == 0 && t == 0) count[b] = grand;
  cluster.sync();   // peers have read s_total (nobody exits early), and row_off is complete for this CTA

  // phase 2: emit.  `off` = first output slot of the 32-word group held by the warp.
  int32_t* out = idx + (size_t)b * cap * 2;
  auto emit_group = [&](int y, int w, uint32_t m, int& off) {
    const int n = __popc(m);
    int pre = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, pre, o);
      if (lane >= o) pre += v;
    }
    int o2 = off + pre - n;
    off += __shfl_sync(0xffffffffu, pre, 31);
    while (m) {
      const int i = __ffs(m) - 1;
      m &= m - 1;
      if (o2 < cap) { out[2 * o2] = y; out[2 * o2 + 1] = w * 32 + i; }
      ++o2;
    }
  };
  if (WL > 0) {
    for (int yr = warp; yr < nrows; yr += 32 * kCompactBatch) {
      uint32_t v[kCompactBatch][WLc];
#pragma unroll
      for (int r = 0; r < kCompactBatch; ++r)
#pragma unroll
        for (int j = 0; j < WLc; ++j) {
          const int y = yb0 + yr + 32 * r, w = lane + 32 * j;
          v[r][j] = (yr + 32 * r < nrows && w < Wwords) ? __ldg(bits + (size_t)y * Wwords + w) : 0u;
        }
#pragma unroll
      for (int r = 0; r < kCompactBatch; ++r) {
        if (yr + 32 * r >= nrows) break;   // warp-uniform
        int off = band_base + row_off[yr + 32 * r];
#pragma unroll
        for (int j = 0; j < WLc; ++j) emit_group(yb0 + yr + 32 * r, lane + 32 * j, v[r][j], off);
      }
    }
  } else {
    for (int yr = warp; yr < nrows; yr += 32) {
      int off = band_base + row_off[yr];
      for (int w0 = 0; w0 < Wwords; w0 += 32) {
        const int w = w0 + lane;
        emit_group(yb0 + yr, w, (w < Wwords) ? __ldg(bits + (size_t)(yb0 + yr) * Wwords + w) : 0u, off);
      }
    }
  }
}

}  // namespace isg

using namespace isg;

extern "C" size_t isg_topk_workspace_bytes(int B, int H, int W, int k) {
  if (B <= 0 || H <= 0 || W <= 0 || k < 0 || (int64_t)H * W > ((int64_t)1 << 30)) return 0;
  return (size_t)B * topk_ws_per_image(H * W, k) + topk_hist15_bytes(B);
}

// keepbits != nullptr: isg_topk_keep - the keep bits are produced together with the threshold
static int topk_threshold_impl(const float* kp, int B, int H, int W, int64_t img_stride, int k, uint32_t* thr_key, void* ws,
                               size_t ws_bytes, uint32_t* keepbits, bool* keep_done, isg_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (keep_done) *keep_done = false;
  if (!kp || !thr_key || B <= 0 || H <= 0 || W <= 0 || k < 0 || B > 65535) return ISG_EINVAL;
  const int64_t npx64 = (int64_t)H * W;
  if (npx64 > (int64_t)1 << 30 || img_stride < npx64) return ISG_EINVAL;
  if ((int64_t)k > npx64) return ISG_EINVAL;  // torch.topk raises (utils/decode.py:81)
  const int npx = (int)npx64;
  if (k == 0) {
    fill_u32_kernel<<<cdiv(B, 128), 128, 0, stream>>>(thr_key, B, 0xffffffffu);
    ISG_LAUNCH_CHECK();
    return ISG_OK;
  }
  if (!ws || ws_bytes < isg_topk_workspace_bytes(B, H, W, k) || ((uintptr_t)ws & 255)) return ISG_EWORKSPACE;
  const bool vec = (npx % 4 == 0) && (img_stride % 4 == 0) && (((uintptr_t)kp & 15) == 0);
  if ((int64_t)4 * k > npx64 && npx >= 65536) {
    // large k: most pixels are candidates; the multi-CTA three-pass radix select is the efficient form
    const size_t per = topk_ws_per_image(npx, k);
    for (int b = 0; b < B; ++b)
      ISG_CUDA(cudaMemsetAsync((char*)ws + (size_t)b * per, 0, 3 * kHistBins * sizeof(uint32_t), stream));
    dim3 grid(cdiv(npx, kHistPxPerBlock), B);
    const int64_t hist_stride = (int64_t)(per / sizeof(uint32_t));
    hist_pass_kernel<0><<<grid, kHistThreads, 0, stream>>>(kp, img_stride, npx, (uint32_t)k, (uint32_t*)ws, hist_stride, vec);
    hist_pass_kernel<1><<<grid, kHistThreads, 0, stream>>>(kp, img_stride, npx, (uint32_t)k, (uint32_t*)ws, hist_stride, vec);
    hist_pass_kernel<2><<<grid, kHistThreads, 0, stream>>>(kp, img_stride, npx, (uint32_t)k, (uint32_t*)ws, hist_stride, vec);
    finish_threshold_kernel<<<B, kHistThreads, 0, stream>>>((uint32_t*)ws, hist_stride, (uint32_t)k, thr_key);
    ISG_LAUNCH_CHECK();
    return ISG_OK;
  }
  if (npx >= 65536) {
    // default: sample -> filter -> select.  ISG_TOPK_PATH=radix selects the sampling-free two-level radix select
    // (exact by construction, measured 82 us vs 71 us on the bench workload in round 1: every stage is latency-bound)
    if (!tuning().topk_radix) {
      int stride = 64;
      if (tuning().topk_cluster_sample) {   // ISG_TOPK_SAMPLE=cluster: the 8-CTA cluster form with a 4x larger sample
        while (npx / stride > kSampleMax) stride *= 2;
        const size_t smem = (size_t)cdiv(npx / stride, kSelCluster) * sizeof(uint32_t);
        ISG_CUDA(cudaFuncSetAttribute(topk_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        topk_sample_kernel<<<dim3(kSelCluster, B), kSelThreads, smem, stream>>>(kp, img_stride, npx, k, stride, ws);
      } else {
        while (npx / stride > kSample1Max) stride *= 2;
        topk_sample1_kernel<<<B, kSelThreads, 0, stream>>>(kp, img_stride, npx, k, stride, ws);
      }
    } else {
      uint32_t* hist15 = reinterpret_cast<uint32_t*>((char*)ws + (size_t)B * topk_ws_per_image(npx, k));
      ISG_CUDA(cudaMemsetAsync(hist15, 0, topk_hist15_bytes(B), stream));
      int dev = 0, sms = kSMs;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      const long long chunks = (long long)B * cdiv(npx, kH15Chunk);
      const int grid15 = (int)std::min<long long>(chunks, sms);
      const size_t smem15 = (size_t)kBins15 * sizeof(uint32_t);
      ISG_CUDA(cudaFuncSetAttribute(topk_hist15_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem15));
      topk_hist15_kernel<<<grid15, kH15Threads, smem15, stream>>>(kp, img_stride, npx, B, hist15, vec);
      topk_pick15_kernel<<<B, 1024, 0, stream>>>(hist15, npx, k, ws);
    }
    dim3 grid(cdiv(npx, kFilterPxPerBlock), B);
    if (keepbits) ISG_CUDA(launch_pdl(topk_filter_kernel<true>, grid, dim3(kFilterThreads), 0, stream, kp, img_stride, npx, k, ws, vec));
    else ISG_CUDA(launch_pdl(topk_filter_kernel<false>, grid, dim3(kFilterThreads), 0, stream, kp, img_stride, npx, k, ws, vec));
  } else {
    // small image: one CTA per image selects over all pixels directly (ncand = 0 -> full-image mode)
    const size_t per = topk_ws_per_image(npx, k);
    for (int b = 0; b < B; ++b)
      ISG_CUDA(cudaMemsetAsync((char*)ws + (size_t)b * per + 3 * kHistBins * sizeof(uint32_t), 0, 64, stream));
  }
  if (tuning().topk_cluster_select)
    ISG_CUDA(launch_pdl(topk_select_kernel, dim3(kSelCluster, B), dim3(kSelThreads), 0, stream, kp, img_stride, npx, k, ws, thr_key,
                        H, W, keepbits));
  else
    ISG_CUDA(launch_pdl(topk_select1_kernel, dim3(B), dim3(kSelThreads), 0, stream, kp, img_stride, npx, k, ws, thr_key, H, W,
                        keepbits));
  ISG_LAUNCH_CHECK();
  if (keep_done) *keep_done = keepbits != nullptr;
  return ISG_OK;
}

extern "C" int isg_topk_threshold(const float* kp, int B, int H, int W, int64_t img_stride, int k,
                                  uint32_t* thr_key, void* ws, size_t ws_bytes, isg_stream_t stream_) {
  return topk_threshold_impl(kp, B, H, W, img_stride, k, thr_key, ws, ws_bytes, nullptr, nullptr, stream_);
}

extern "C" int isg_keep_points(const float* kp, int B, int H, int W, int64_t img_stride, const uint32_t* thr_key,
                               uint32_t* keepbits, uint8_t* mask_u8, isg_stream_t stream_);

extern "C" int isg_topk_keep(const float* kp, int B, int H, int W, int64_t img_stride, int k, uint32_t* thr_key,
                             uint32_t* keepbits, int keepbits_zeroed, void* ws, size_t ws_bytes, isg_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!keepbits || B <= 0 || H <= 0 || W <= 0) return ISG_EINVAL;
  if (!keepbits_zeroed) ISG_CUDA(cudaMemsetAsync(keepbits, 0, (size_t)B * H * cdiv(W, 32) * sizeof(uint32_t), stream));
  bool done = false;
  const int rc = topk_threshold_impl(kp, B, H, W, img_stride, k, thr_key, ws, ws_bytes, keepbits, &done, stream_);
  if (rc != ISG_OK || done) return rc;
  // the threshold paths without a select kernel (k = 0, large k): the streaming keep kernel
  return isg_keep_points(kp, B, H, W, img_stride, thr_key, keepbits, nullptr, stream_);
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

extern "C" size_t isg_select_points_workspace_bytes(int B, int H, int W, int k) {
  const size_t t = isg_topk_workspace_bytes(B, H, W, k);
  return t ? t + (((size_t)B * 4 + 255) & ~(size_t)255) : 0;
}

extern "C" int isg_select_points(const float* kp, int B, int H, int W, int64_t img_stride, int k,
                                 uint32_t* keepbits, uint8_t* mask_u8, void* ws, size_t ws_bytes,
                                 isg_stream_t stream) {
  if (B <= 0 || H <= 0 || W <= 0 || k < 0) return ISG_EINVAL;
  if ((int64_t)k > (int64_t)H * W) return ISG_EINVAL;
  const size_t tb = isg_topk_workspace_bytes(B, H, W, k);
  if (!ws || tb == 0 || ws_bytes < isg_select_points_workspace_bytes(B, H, W, k) || ((uintptr_t)ws & 255)) return ISG_EWORKSPACE;
  uint32_t* thr = (uint32_t*)((char*)ws + tb);
  int rc = isg_topk_threshold(kp, B, H, W, img_stride, k, thr, ws, tb, stream);
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
  if (!keepbits || !idx || !count || B <= 0 || H <= 0 || W <= 0 || cap < 0 || B > 65535) return ISG_EINVAL;
  const int Wwords = cdiv(W, 32);
  const size_t smem = (size_t)(cdiv(H, kCompactCluster) + 1) * sizeof(int);
  if (smem > 160 * 1024) return ISG_EUNSUPPORTED;
#define ISG_COMPACT_LAUNCH(WL_)                                                                                 \
  do {                                                                                                          \
    ISG_CUDA(cudaFuncSetAttribute(compact_kernel<WL_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    compact_kernel<WL_><<<dim3(kCompactCluster, B), kCompactThreads, smem, stream>>>(keepbits, H, Wwords, W, cap, idx, count); \
  } while (0)
  if (Wwords <= 32) ISG_COMPACT_LAUNCH(1);
  else if (Wwords <= 64) ISG_COMPACT_LAUNCH(2);
  else ISG_COMPACT_LAUNCH(0);
#undef ISG_COMPACT_LAUNCH
  ISG_LAUNCH_CHECK();
  return ISG_OK;
}
