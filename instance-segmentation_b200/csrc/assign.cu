// K1 + K3 — embedding, Gaussian membership, assignment, per-instance statistics and grouping.
// Reference: group_kp (utils/decode.py:303-356) and decode_single (utils/decode.py:428-432).
//
// Parity-critical arithmetic uses __f*_rn intrinsics so that nvcc never contracts it into FMAs:
// torch evaluates  (e - c)^2 * sigma  and the 2-term sum with separate roundings (:326-327).
#include <algorithm>
#include <cstdlib>
#include "keep.cuh"

namespace isg {

struct __align__(16) SeedRec {   // ISG_SEED_WORDS = 8
  int y0, y1, x0, x1;            // inclusive in-box pixel bounds (empty: y0 > y1)
  float cy, cx;                  // grid coordinate of the truncated centre (utils/decode.py:317)
  int id;                        // seed index (used by the per-tile compacted copy)
  int pad;
};
static_assert(sizeof(SeedRec) == ISG_SEED_WORDS * 4, "seed record layout");

__device__ __forceinline__ int clamp_to_int(float v) {
  v = fminf(fmaxf(v, -1073741824.0f), 1073741824.0f);
  return (int)v;
}

// ---------------------------------------------------------------------------------------------
// seeds: one thread per (image, box)
// ---------------------------------------------------------------------------------------------
// seed record + ghost bounds of one box (`valid` false: an empty record no tile overlaps)
__device__ __forceinline__ void make_seed(const float4 r, int layout, bool valid, int j, const float* __restrict__ ys,
                                          const float* __restrict__ xs, int H, int W, float ghost_k, float scale,
                                          SeedRec& s, float4& g) {
  // empty box: no tile overlaps it, so the per-pixel range tests never see it
  s.y0 = 0x7fffffff; s.y1 = -0x7fffffff - 1; s.x0 = 0x7fffffff; s.x1 = -0x7fffffff - 1;
  s.cy = 0.f; s.cx = 0.f; s.id = j; s.pad = 0;
  g = make_float4(0.f, 0.f, 0.f, 0.f);
  if (!valid) return;
  float cy, cx, hy, wx;
  if (layout == ISG_BOX_XYXY) {   // x1,y1,x2,y2
    // decode_single: lt = (y1,x1), rb = (y2,x2); centre = (lt+rb)/2; wh = rb-lt   (:428-432)
    cy = __fmul_rn(__fadd_rn(r.y, r.w), 0.5f); cx = __fmul_rn(__fadd_rn(r.x, r.z), 0.5f);
    hy = __fsub_rn(r.w, r.y); wx = __fsub_rn(r.z, r.x);
  } else {                        // cy,cx,h,w : group_kp's center_indexes / center_whs (:288)
    cy = r.x; cx = r.y; hy = r.z; wx = r.w;
  }
  // group_kp: lt = c - wh/2, rb = c + wh/2 ; inbox = (p - lt >= 0) & (rb - p >= 0)   (:321-325)
  const float lty = __fsub_rn(cy, __fmul_rn(hy, 0.5f)), ltx = __fsub_rn(cx, __fmul_rn(wx, 0.5f));
  const float rby = __fadd_rn(cy, __fmul_rn(hy, 0.5f)), rbx = __fadd_rn(cx, __fmul_rn(wx, 0.5f));
  // p is an integer-valued float, so  p - lt >= 0  <=>  p >= ceil(lt)  and  rb - p >= 0  <=>  p <= floor(rb)
  const bool finite = (lty == lty) && (ltx == ltx) && (rby == rby) && (rbx == rbx);
  if (finite) {
    const int y0 = clamp_to_int(ceilf(lty)), y1 = clamp_to_int(floorf(rby));
    const int x0 = clamp_to_int(ceilf(ltx)), x1 = clamp_to_int(floorf(rbx));
    if (y0 <= y1 && x0 <= x1) { s.y0 = y0; s.y1 = y1; s.x0 = x0; s.x1 = x1; }
  }
  // seed coordinate = grid value at the truncated centre (:316-317); clamped into the image
  int iy = clamp_to_int(cy), ix = clamp_to_int(cx);
  iy = min(max(iy, 0), H - 1); ix = min(max(ix, 0), W - 1);
  s.cy = ys[iy]; s.cx = xs[ix];
  // ghost filter bounds (:339-352): x -/+ (0.5+wh_delta)*w, y -/+ (0.5+wh_delta)*h, fp32
  const float w = __fmul_rn(wx, scale), h = __fmul_rn(hy, scale);
  if (ghost_k < 0.0f) {   // ghost filter disabled: every pixel passes
    g = make_float4(-INFINITY, INFINITY, -INFINITY, INFINITY);
  } else {
    g.x = __fsub_rn(cx, __fmul_rn(ghost_k, w)); g.y = __fadd_rn(cx, __fmul_rn(ghost_k, w));
    g.z = __fsub_rn(cy, __fmul_rn(ghost_k, h)); g.w = __fadd_rn(cy, __fmul_rn(ghost_k, h));
  }
}

__global__ void build_seeds_kernel(const float* __restrict__ rois, int layout, const int32_t* __restrict__ n_seeds,
                                   int Nmax, const float* __restrict__ ys, const float* __restrict__ xs,
                                   int H, int W, float ghost_k, float scale, SeedRec* __restrict__ seeds,
                                   float4* __restrict__ ghost) {
  const int b = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= Nmax) return;
  const bool valid = j < n_seeds[b];
  float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
  if (valid) r = reinterpret_cast<const float4*>(rois)[(size_t)b * Nmax + j];
  SeedRec s;
  float4 g;
  make_seed(r, layout, valid, j, ys, xs, H, W, ghost_k, scale, s, g);
  seeds[(size_t)b * Nmax + j] = s;
  ghost[(size_t)b * Nmax + j] = g;
}

// gather_kept + build_seeds + stats_init of the batched pipeline in one launch: thread (b, r) copies the r-th kept
// box of image b into the detection tables, derives its seed record / ghost bounds and resets its statistics
__global__ void gather_build_seeds_kernel(const float4* __restrict__ cand_boxes, const float* __restrict__ cand_scores,
                                          const int32_t* __restrict__ cand_cls, const int32_t* __restrict__ keep,
                                          const int32_t* __restrict__ n_keep, int cap, int Nmax,
                                          const float* __restrict__ ys, const float* __restrict__ xs, int H, int W,
                                          float ghost_k, float scale, float4* __restrict__ rois,
                                          float* __restrict__ scores, int32_t* __restrict__ cls,
                                          int32_t* __restrict__ n_out, SeedRec* __restrict__ seeds,
                                          float4* __restrict__ ghost, int32_t* __restrict__ stats,
                                          int32_t* __restrict__ img_total) {
  pdl_trigger();
  pdl_wait();         // the keep list comes from the NMS kernel launched just before
  const int b = blockIdx.y;
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = min(n_keep[b], Nmax);
  if (r == 0) { n_out[b] = n; if (img_total) img_total[b] = 0; }
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
  SeedRec s;
  float4 g;
  make_seed(bx, ISG_BOX_XYXY, r < n, r, ys, xs, H, W, ghost_k, scale, s, g);
  seeds[d] = s;
  ghost[d] = g;
  if (stats) {
    int32_t* st = stats + d * ISG_STAT_WORDS;
    st[0] = 0; st[1] = 0x7fffffff; st[2] = 0x7fffffff; st[3] = -1; st[4] = -1;
  }
}

__global__ void stats_init_kernel(int32_t* stats, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int32_t* s = stats + (size_t)i * ISG_STAT_WORDS;
  s[0] = 0; s[1] = 0x7fffffff; s[2] = 0x7fffffff; s[3] = -1; s[4] = -1;
}

__device__ __forceinline__ bool ghost_pass(const float4 g, int y, int x) {
  const float fx = (float)x, fy = (float)y;
  return (g.x < fx) && (fx < g.y) && (g.z < fy) && (fy < g.w);   // strict (:351-352)
}
__device__ __forceinline__ void stats_add(int32_t* stats, int y, int x) {
  atomicAdd(stats + 0, 1);
  atomicMin(stats + 1, y); atomicMin(stats + 2, x);
  atomicMax(stats + 3, y); atomicMax(stats + 4, x);
}

// membership of one pixel against one seed centre; returns P = exp(-q)
__device__ __forceinline__ float membership(float ey, float ex, float sy, float sx, float cy, float cx) {
  const float dy = __fsub_rn(ey, cy), dx = __fsub_rn(ex, cx);
  const float qy = __fmul_rn(__fmul_rn(dy, dy), sy);
  const float qx = __fmul_rn(__fmul_rn(dx, dx), sx);
  return exp_fast(-__fadd_rn(qy, qx));
}

// keep nibbles of RW consecutive rows starting at ybeg (rows are re-used through a rolling window)
template <int RW, bool VEC, bool USE_INT>
__device__ __forceinline__ void keep_rows_global(const float* __restrict__ img, int ybeg, int x0, int H, int W,
                                                 const Thr& thr, int lane, uint32_t (&nib)[RW]) {
  RowH up = load_rowh<VEC, USE_INT>(img, ybeg - 1, x0, H, W, thr, lane);
  RowH mid = load_rowh<VEC, USE_INT>(img, ybeg, x0, H, W, thr, lane);
#pragma unroll
  for (int r = 0; r < RW; ++r) {
    const RowH dn = load_rowh<VEC, USE_INT>(img, ybeg + r + 1, x0, H, W, thr, lane);
    nib[r] = (ybeg + r < H) ? keep_nibble(up, mid, dn) : 0u;
    up = mid; mid = dn;
  }
}

// One culled seed against the lane's RW x 4 pixels.  bx = {y0,y1,x0,x1}; the caller has established that the
// seed overlaps the lane's columns and the warp's rows.  Branch-free inside: pixels outside the box compute a
// probability that the range predicate discards.
template <int RW>
__device__ __forceinline__ void seed_update(const int4 bx, float cy, float cx, int id, int ybeg, int x0,
                                            const float (&a)[RW][4][4], float (&best)[RW][4], int (&lab)[RW][4]) {
  const unsigned hy = (unsigned)(bx.y - bx.x), hx = (unsigned)(bx.w - bx.z);
#pragma unroll
  for (int r = 0; r < RW; ++r) {
    const bool rowin = (unsigned)(ybeg + r - bx.x) <= hy;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const bool in = rowin && ((unsigned)(x0 + i - bx.z) <= hx);
      const float P = membership(a[r][0][i], a[r][1][i], a[r][2][i], a[r][3][i], cy, cx);
      if (in && P > best[r][i]) { best[r][i] = P; lab[r][i] = id; }   // strict: first index wins ties (:328)
    }
  }
}

// ---------------------------------------------------------------------------------------------
// sparse: one thread per compacted keep pixel, all seeds of the image staged in shared memory
// ---------------------------------------------------------------------------------------------
constexpr int kSparseThreads = 256;

__global__ void __launch_bounds__(kSparseThreads)
assign_sparse_kernel(const float* __restrict__ ae, int64_t img_stride, int64_t plane_stride,
                     const int32_t* __restrict__ idx, const int32_t* __restrict__ count, int cap,
                     const SeedRec* __restrict__ seeds, const float4* __restrict__ ghost,
                     const int32_t* __restrict__ n_seeds, int Nmax, int W,
                     const float* __restrict__ ys, const float* __restrict__ xs,
                     int32_t* __restrict__ label, float* __restrict__ score, uint8_t* __restrict__ flag,
                     int32_t* __restrict__ stats) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SeedRec* s_seed = reinterpret_cast<SeedRec*>(smem_raw);
  __shared__ __align__(8) uint64_t bar;
  const int b = blockIdx.y;
  const int M = min(count[b], cap);
  if ((int)(blockIdx.x * kSparseThreads) >= M) return;
  const int n = min(n_seeds[b], Nmax);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  __syncthreads();
  if (n > 0) {
    if (threadIdx.x == 0) {
      mbar_expect_tx(&bar, (uint32_t)(n * sizeof(SeedRec)));
      bulk_g2s(s_seed, seeds + (size_t)b * Nmax, (uint32_t)(n * sizeof(SeedRec)), &bar);
    }
  }
  const int m = blockIdx.x * kSparseThreads + threadIdx.x;
  int y = 0, x = 0;
  float ey = 0.f, ex = 0.f, sy = 0.f, sx = 0.f;
  if (m < M) {
    y = idx[((size_t)b * cap + m) * 2];
    x = idx[((size_t)b * cap + m) * 2 + 1];
    const float* p = ae + (int64_t)b * img_stride + (int64_t)y * W + x;
    const float a0 = __ldg(p), a1 = __ldg(p + plane_stride), a2 = __ldg(p + 2 * plane_stride),
                a3 = __ldg(p + 3 * plane_stride);
    ey = __fadd_rn(tanh_fast(a0), ys[y]); ex = __fadd_rn(tanh_fast(a1), xs[x]);   // :305
    sy = exp_fast(a2); sx = exp_fast(a3);                                         // :315
  }
  if (n > 0) mbar_wait(&bar, 0);
  if (m >= M) return;
  float best = 0.0f;
  int lab = 0;
  for (int j = 0; j < n; ++j) {
    const SeedRec s = s_seed[j];
    if (y >= s.y0 && y <= s.y1 && x >= s.x0 && x <= s.x1) {
      const float P = membership(ey, ex, sy, sx, s.cy, s.cx);
      if (P > best) { best = P; lab = j; }   // strict: first index wins ties (:328)
    }
  }
  const size_t o = (size_t)b * cap + m;
  label[o] = lab;
  if (score) score[o] = best;
  bool f = false;
  if (n > 0) {
    f = ghost_pass(ghost[(size_t)b * Nmax + lab], y, x);
    if (f && stats) stats_add(stats + ((size_t)b * Nmax + lab) * ISG_STAT_WORDS, y, x);
  }
  if (flag) flag[o] = f ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------
// dense fused kernel: tile = 128 px x (8 warps * RW rows); lane owns 4 consecutive pixels of a row
// ---------------------------------------------------------------------------------------------
constexpr int kDenseWarps = 8;

template <int RW, bool VEC, bool SCORE>
__global__ void __launch_bounds__(32 * kDenseWarps)
assign_dense_kernel(const float* __restrict__ kp, int64_t kp_img_stride,
                    const float* __restrict__ ae, int64_t ae_img_stride, int64_t ae_plane_stride,
                    const uint32_t* __restrict__ thr_key,
                    const SeedRec* __restrict__ seeds, const float4* __restrict__ ghost,
                    const int32_t* __restrict__ n_seeds, int Nmax, int H, int W, int Wwords,
                    const float* __restrict__ ys, const float* __restrict__ xs,
                    int32_t* __restrict__ label_map, float* __restrict__ score_map,
                    uint32_t* __restrict__ keepbits, int32_t* __restrict__ stats) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SeedRec* s_all = reinterpret_cast<SeedRec*>(smem_raw);   // [Nmax] whole table (bulk copy target)
  SeedRec* s_hit = s_all + Nmax;                           // [Nmax] seeds intersecting this tile
  __shared__ __align__(8) uint64_t bar;
  __shared__ int s_wcnt[kDenseWarps];

  const int b = blockIdx.z;
  const int lane = threadIdx.x, warp = threadIdx.y, tid = warp * 32 + lane;
  const int n = min(n_seeds[b], Nmax);
  const int tile_x0 = blockIdx.x * 128, tile_y0 = blockIdx.y * (kDenseWarps * RW);
  const int tile_x1 = min(tile_x0 + 127, W - 1), tile_y1 = min(tile_y0 + kDenseWarps * RW - 1, H - 1);

  // (1) seed table -> shared memory through the TMA unit (1-D bulk copy), asynchronously
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  __syncthreads();
  if (tid == 0 && n > 0) {
    mbar_expect_tx(&bar, (uint32_t)(n * sizeof(SeedRec)));
    bulk_g2s(s_all, seeds + (size_t)b * Nmax, (uint32_t)(n * sizeof(SeedRec)), &bar);
  }

  // (2) issue every pixel load of this thread (ae: 4 planes x RW rows; streamed, read once)
  const int x0 = tile_x0 + lane * 4;
  const int ybeg = tile_y0 + warp * RW;
  const float* aeb = ae + (int64_t)b * ae_img_stride;
  float a[RW][4][4];   // [row][plane][px]
#pragma unroll
  for (int r = 0; r < RW; ++r) {
    const int y = ybeg + r;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float* p = aeb + c * ae_plane_stride + (int64_t)y * W + x0;
      if (VEC && y < H && x0 + 3 < W) {
        const float4 v = ldg_stream4(p);
        a[r][c][0] = v.x; a[r][c][1] = v.y; a[r][c][2] = v.z; a[r][c][3] = v.w;
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) a[r][c][i] = (y < H && x0 + i < W) ? ldg_stream1(p + i) : 0.0f;
      }
    }
  }

  // (3) keep bits from kp (threshold + separable 3x3 peak); hides the ae latency
  const Thr thr = make_thr(thr_key[b]);
  const float* kpb = kp + (int64_t)b * kp_img_stride;
  uint32_t nib[RW];
  if (thr.use_int) keep_rows_global<RW, VEC, true>(kpb, ybeg, x0, H, W, thr, lane, nib);
  else keep_rows_global<RW, VEC, false>(kpb, ybeg, x0, H, W, thr, lane, nib);
#pragma unroll
  for (int r = 0; r < RW; ++r) {
    const uint32_t word = nibbles_to_word(nib[r], lane);
    if ((lane & 7) == 0 && x0 < W && ybeg + r < H)
      keepbits[((size_t)b * H + ybeg + r) * Wwords + (x0 >> 5)] = word;
  }

  // (4) cull the seed table against this tile, preserving seed order (first-index rule)
  int T = 0;
  if (n > 0) {
    mbar_wait(&bar, 0);
    for (int j0 = 0; j0 < n; j0 += 32 * kDenseWarps) {
      const int j = j0 + tid;
      bool hit = false;
      SeedRec s;
      if (j < n) {
        s = s_all[j];
        hit = s.y0 <= tile_y1 && s.y1 >= tile_y0 && s.x0 <= tile_x1 && s.x1 >= tile_x0;
      }
      const unsigned bal = __ballot_sync(0xffffffffu, hit);
      if (lane == 0) s_wcnt[warp] = __popc(bal);
      __syncthreads();
      int pre = T, tot = 0;
#pragma unroll
      for (int w = 0; w < kDenseWarps; ++w) {
        const int c = s_wcnt[w];
        if (w < warp) pre += c;
        tot += c;
      }
      if (hit) { s.id = j; s_hit[pre + __popc(bal & ((1u << lane) - 1u))] = s; }
      T += tot;
      __syncthreads();
    }
  }

  // (5) embedding (per pixel) : e = tanh(ae01) + grid, s = exp(ae23)      (:305,315)
  float xs4[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) xs4[i] = (x0 + i < W) ? __ldg(xs + x0 + i) : 0.0f;
  {
    float amax = 0.0f;
#pragma unroll
    for (int r = 0; r < RW; ++r)
#pragma unroll
      for (int i = 0; i < 4; ++i) amax = fmaxf(amax, fmaxf(fabsf(a[r][0][i]), fabsf(a[r][1][i])));
    const bool small = __all_sync(0xffffffffu, amax < 0.55f);   // warp-uniform: polynomial branch only
#pragma unroll
    for (int r = 0; r < RW; ++r) {
      const float yv = (ybeg + r < H) ? __ldg(ys + ybeg + r) : 0.0f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float t0 = small ? tanh_poly(a[r][0][i]) : tanh_fast(a[r][0][i]);
        const float t1 = small ? tanh_poly(a[r][1][i]) : tanh_fast(a[r][1][i]);
        a[r][0][i] = __fadd_rn(t0, yv);
        a[r][1][i] = __fadd_rn(t1, xs4[i]);
        a[r][2][i] = exp_fast(a[r][2][i]);
        a[r][3][i] = exp_fast(a[r][3][i]);
      }
    }
  }

  // (6) membership against the culled seeds, ascending seed index
  float best[RW][4];
  int lab[RW][4];
#pragma unroll
  for (int r = 0; r < RW; ++r)
#pragma unroll
    for (int i = 0; i < 4; ++i) { best[r][i] = 0.0f; lab[r][i] = 0; }

  for (int t = 0; t < T; ++t) {
    const int4 bx = *reinterpret_cast<const int4*>(&s_hit[t]);          // y0,y1,x0,x1 (broadcast)
    if (bx.x > ybeg + RW - 1 || bx.y < ybeg) continue;                   // warp-uniform row cull
    if (bx.z > x0 + 3 || bx.w < x0) continue;                            // lane column cull
    const float4 cc = *reinterpret_cast<const float4*>(&s_hit[t].cy);   // cy,cx,id,pad
    seed_update<RW>(bx, cc.x, cc.y, __float_as_int(cc.z), ybeg, x0, a, best, lab);
  }

  // (7) stores + statistics of the keep pixels
#pragma unroll
  for (int r = 0; r < RW; ++r) {
    const int y = ybeg + r;
    if (y >= H) break;
    int32_t* lrow = label_map + ((size_t)b * H + y) * W + x0;
    if (VEC && x0 + 3 < W) {
      stg_stream4(lrow, lab[r][0], lab[r][1], lab[r][2], lab[r][3]);
      if (SCORE) stg_stream4f(score_map + ((size_t)b * H + y) * W + x0, best[r][0], best[r][1], best[r][2], best[r][3]);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (x0 + i < W) {
          lrow[i] = lab[r][i];
          if (SCORE) score_map[((size_t)b * H + y) * W + x0 + i] = best[r][i];
        }
    }
    if (nib[r] && stats && n > 0) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if ((nib[r] >> i) & 1u) {
          const int l = lab[r][i];
          if (ghost_pass(ghost[(size_t)b * Nmax + l], y, x0 + i))
            stats_add(stats + ((size_t)b * Nmax + l) * ISG_STAT_WORDS, y, x0 + i);
        }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// dense mode: read labels / scores of the compacted keep pixels back from the maps
// ---------------------------------------------------------------------------------------------
__global__ void gather_labels_kernel(const int32_t* __restrict__ label_map, const float* __restrict__ score_map,
                                     const int32_t* __restrict__ idx, const int32_t* __restrict__ count, int cap,
                                     const float4* __restrict__ ghost, int Nmax, int H, int W,
                                     int32_t* __restrict__ label, float* __restrict__ score,
                                     uint8_t* __restrict__ flag, int32_t* __restrict__ stats) {
  const int b = blockIdx.y;
  const int M = min(count[b], cap);
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const size_t o = (size_t)b * cap + m;
  const int y = idx[o * 2], x = idx[o * 2 + 1];
  const size_t p = ((size_t)b * H + y) * W + x;
  const int l = label_map[p];
  label[o] = l;
  if (score && score_map) score[o] = score_map[p];
  const bool pass = Nmax > 0 && l >= 0 && l < Nmax && ghost_pass(ghost[(size_t)b * Nmax + l], y, x);
  if (flag) flag[o] = pass ? 1 : 0;
  if (stats && pass) stats_add(stats + ((size_t)b * Nmax + l) * ISG_STAT_WORDS, y, x);
}

// ---------------------------------------------------------------------------------------------
// grouping: one warp per instance scans the image's label list and compacts its flagged pixels in
// row-major order (deterministic, no atomics).
// ---------------------------------------------------------------------------------------------
constexpr int kGroupWarps = 8;

__global__ void __launch_bounds__(32 * kGroupWarps)
group_count_kernel(const int32_t* __restrict__ label, const uint8_t* __restrict__ flag,
                   const int32_t* __restrict__ count, int cap, const int32_t* __restrict__ n_seeds, int Nmax,
                   int32_t* __restrict__ offsets) {
  // offsets[b][i+1] = number of flagged pixels of instance i (scanned in place by group_scan_kernel)
  const int b = blockIdx.y;
  const int i = blockIdx.x * kGroupWarps + threadIdx.y;
  if (i >= Nmax) return;
  const int lane = threadIdx.x;
  const int M = min(count[b], cap);
  int c = 0;
  if (i < n_seeds[b]) {
    const int32_t* lb = label + (size_t)b * cap;
    const uint8_t* fb = flag + (size_t)b * cap;
    for (int m = lane; m < M; m += 32) c += (lb[m] == i && fb[m]) ? 1 : 0;
    c = warp_sum(c);
  }
  if (lane == 0) offsets[(size_t)b * (Nmax + 1) + i + 1] = c;
}

__global__ void group_scan_kernel(int32_t* __restrict__ offsets, int Nmax) {
  // one warp per image: in-place exclusive scan of offsets[b][1..Nmax]
  const int b = blockIdx.x, lane = threadIdx.x;
  int32_t* o = offsets + (size_t)b * (Nmax + 1);
  int carry = 0;
  if (lane == 0) o[0] = 0;
  for (int base = 0; base < Nmax; base += 32) {
    const int i = base + lane;
    const int v = i < Nmax ? o[i + 1] : 0;
    int s = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int u = __shfl_up_sync(0xffffffffu, s, d);
      if (lane >= d) s += u;
    }
    if (i < Nmax) o[i + 1] = carry + s;
    carry += __shfl_sync(0xffffffffu, s, 31);
  }
}

__global__ void __launch_bounds__(32 * kGroupWarps)
group_scatter_kernel(const int32_t* __restrict__ idx, const int32_t* __restrict__ label,
                     const uint8_t* __restrict__ flag, const int32_t* __restrict__ count, int cap,
                     const int32_t* __restrict__ n_seeds, int Nmax, const int32_t* __restrict__ offsets,
                     float* __restrict__ points) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * kGroupWarps + threadIdx.y;
  if (i >= Nmax || i >= n_seeds[b]) return;
  const int lane = threadIdx.x;
  const int M = min(count[b], cap);
  const int32_t* lb = label + (size_t)b * cap;
  const uint8_t* fb = flag + (size_t)b * cap;
  const int32_t* ib = idx + (size_t)b * cap * 2;
  float2* out = reinterpret_cast<float2*>(points) + (size_t)b * cap;
  int off = offsets[(size_t)b * (Nmax + 1) + i];
  for (int m0 = 0; m0 < M; m0 += 32) {
    const int m = m0 + lane;
    const bool hit = m < M && lb[m] == i && fb[m];
    const unsigned bal = __ballot_sync(0xffffffffu, hit);
    if (hit) {
      const int y = ib[2 * m], x = ib[2 * m + 1];
      out[off + __popc(bal & ((1u << lane) - 1u))] = make_float2((float)x, (float)y);   // (x,y) flip
    }
    off += __popc(bal);
  }
}

// ---------------------------------------------------------------------------------------------
// grouping v2: stable multisplit by label, one CTA per image.  Warp w owns a contiguous segment of the
// row-major keep list; per-warp label histograms (phase 1) are prefix-summed across warps and labels
// (phase 2) and each warp then scatters its segment in order (phase 3).  __match_any_sync gives every lane
// its rank among the lanes of the same label, so the output order inside an instance stays row-major.
// ---------------------------------------------------------------------------------------------
constexpr int kSplitWarps = 16;

__global__ void __launch_bounds__(32 * kSplitWarps)
group_split_kernel(const int32_t* __restrict__ idx, const int32_t* __restrict__ label,
                   const uint8_t* __restrict__ flag, const int32_t* __restrict__ count, int cap,
                   const int32_t* __restrict__ n_seeds, int Nmax, int32_t* __restrict__ offsets,
                   float* __restrict__ points) {
  extern __shared__ int sm_split[];
  int* hist = sm_split;                                           // [kSplitWarps][Nmax]
  int* tot = sm_split + kSplitWarps * Nmax;                       // [Nmax + 1]
  int* slab = tot + Nmax + 1;                                     // [M]: label (or -1), later the output slot
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int M = min(count[b], cap), n = min(n_seeds[b], Nmax);
  const int32_t* lb = label + (size_t)b * cap;
  const uint8_t* fb = flag + (size_t)b * cap;
  const int32_t* ib = idx + (size_t)b * cap * 2;
  float2* out = reinterpret_cast<float2*>(points) + (size_t)b * cap;
  int32_t* off_g = offsets + (size_t)b * (Nmax + 1);
  for (int i = tid; i < kSplitWarps * Nmax; i += 32 * kSplitWarps) hist[i] = 0;
  // stage the (flag-filtered) labels in shared memory: independent, coalesced loads
  for (int m = tid; m < M; m += 32 * kSplitWarps) {
    int l = -1;
    if (fb[m]) { l = lb[m]; if (l < 0 || l >= n) l = -1; }
    slab[m] = l;
  }
  __syncthreads();
  const int seg = ((M + kSplitWarps * 32 - 1) / (kSplitWarps * 32)) * 32;
  const int m0 = warp * seg, m1 = min(m0 + seg, M);
  int* myhist = hist + warp * Nmax;
  // phase 1: per-warp label histogram of the warp's contiguous segment
  for (int mb = m0; mb < m1; mb += 32) {
    const int m = mb + lane;
    const int l = m < m1 ? slab[m] : -1;
    const unsigned peers = __match_any_sync(0xffffffffu, l);
    if (l >= 0 && lane == __ffs(peers) - 1) myhist[l] += __popc(peers);
    __syncwarp();
  }
  __syncthreads();
  // phase 2a: exclusive prefix across warps, per label
  for (int l = tid; l < n; l += 32 * kSplitWarps) {
    int run = 0;
#pragma unroll
    for (int w = 0; w < kSplitWarps; ++w) { const int c = hist[w * Nmax + l]; hist[w * Nmax + l] = run; run += c; }
    tot[l] = run;
  }
  __syncthreads();
  // phase 2b: exclusive scan over labels (warp 0)
  if (warp == 0) {
    int carry = 0;
    for (int base = 0; base < n; base += 32) {
      const int i = base + lane;
      const int v = i < n ? tot[i] : 0;
      int sc = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, sc, o);
        if (lane >= o) sc += u;
      }
      if (i < n) { tot[i] = carry + sc - v; off_g[i] = carry + sc - v; }
      carry += __shfl_sync(0xffffffffu, sc, 31);
    }
    for (int i = n + lane; i <= Nmax; i += 32) off_g[i] = carry;
  }
  __syncthreads();
  // phase 3: output slot of every element, in order (shared memory only)
  for (int mb = m0; mb < m1; mb += 32) {
    const int m = mb + lane;
    const int l = m < m1 ? slab[m] : -1;
    const unsigned peers = __match_any_sync(0xffffffffu, l);
    int pos = -1;
    if (l >= 0) pos = tot[l] + myhist[l] + __popc(peers & ((1u << lane) - 1u));
    __syncwarp();
    if (l >= 0 && lane == __ffs(peers) - 1) myhist[l] += __popc(peers);
    if (m < m1) slab[m] = pos;
    __syncwarp();
  }
  __syncthreads();
  // phase 4: scatter (independent loads / stores)
  for (int m = tid; m < M; m += 32 * kSplitWarps) {
    const int pos = slab[m];
    if (pos >= 0) {
      const int2 yx = *reinterpret_cast<const int2*>(ib + 2 * m);
      out[pos] = make_float2((float)yx.y, (float)yx.x);   // (x,y) flip
    }
  }
}

}  // namespace isg

#include "dense_v4.cuh"

using namespace isg;

static bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

extern "C" int isg_build_seeds(const float* rois, int layout, const int32_t* n_seeds, int B, int Nmax, const float* ys,
                               const float* xs, int H, int W, float ghost_k, float scale, uint32_t* seeds,
                               float* ghost, isg_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!rois || !n_seeds || !ys || !xs || !seeds || !ghost || B <= 0 || Nmax <= 0 || H <= 0 || W <= 0) return ISG_EINVAL;
  if (!aligned16(rois) || !aligned16(seeds) || !aligned16(ghost) || B > 65535) return ISG_EINVAL;
  if (layout != ISG_BOX_XYXY && layout != ISG_BOX_CYCXHW) return ISG_EINVAL;
  dim3 grid(cdiv(Nmax, 128), B);
  build_seeds_kernel<<<grid, 128, 0, stream>>>(rois, layout, n_seeds, Nmax, ys, xs, H, W, ghost_k, scale,
                                               reinterpret_cast<SeedRec*>(seeds), reinterpret_cast<float4*>(ghost));
  ISG_LAUNCH_CHECK();
  return ISG_OK;
}

extern "C" int isg_stats_init(int32_t* stats, int B, int Nmax, isg_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!stats || B <= 0 || Nmax <= 0) return ISG_EINVAL;
  const int n = B * Nmax;
  stats_init_kernel<<<cdiv(n, 256), 256, 0, stream>>>(stats, n);
  ISG_LAUNCH_CHECK();
  return ISG_OK;
}

extern "C" int isg_gather_build_seeds(const float* cand_boxes, const float* cand_scores, const int32_t* cand_cls,
                                      const int32_t* keep, const int32_t* n_keep, int B, int cap, int Nmax, const float* ys,
                                      const float* xs, int H, int W, float ghost_k, float scale, float* rois, float* scores,
                                      int32_t* cls, int32_t* n_out, uint32_t* seeds, float* ghost, int32_t* stats,
                                      int32_t* img_total, isg_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!cand_boxes || !cand_scores || !cand_cls || !keep || !n_keep || !ys || !xs || !rois || !scores || !cls || !n_out ||
      !seeds || !ghost)
    return ISG_EINVAL;
  if (B <= 0 || cap <= 0 || Nmax <= 0 || H <= 0 || W <= 0 || B > 65535) return ISG_EINVAL;
  if (!aligned16(cand_boxes) || !aligned16(rois) || !aligned16(seeds) || !aligned16(ghost)) return ISG_EINVAL;
  dim3 grid(cdiv(Nmax, 128), B);
  ISG_CUDA(launch_pdl(gather_build_seeds_kernel, grid, dim3(128), 0, stream, reinterpret_cast<const float4*>(cand_boxes), cand_scores,
                      cand_cls, keep, n_keep, cap, Nmax, ys, xs, H, W, ghost_k, scale, reinterpret_cast<float4*>(rois), scores, cls,
                      n_out, reinterpret_cast<SeedRec*>(seeds), reinterpret_cast<float4*>(ghost), stats, img_total));
  ISG_LAUNCH_CHECK();
  return ISG_OK;
}

extern "C" int isg_assign_sparse(const float* ae, int64_t img_stride, int64_t plane_stride, const int32_t* idx,
                                 const int32_t* count, int cap, const uint32_t* seeds, const float* ghost,
                                 const int32_t* n_seeds, int B, int Nmax, int H, int W, const float* ys,
                                 const float* xs, int32_t* label, float* score, uint8_t* flag, int32_t* stats,
                                 isg_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!ae || !idx || !count || !seeds || !ghost || !n_seeds || !ys || !xs || !label) return ISG_EINVAL;
  if (B <= 0 || Nmax <= 0 || H <= 0 || W <= 0 || cap <= 0 || B > 65535) return ISG_EINVAL;
  if (!aligned16(seeds) || !aligned16(ghost)) return ISG_EINVAL;
  const size_t smem = (size_t)Nmax * sizeof(SeedRec);
  if (smem > 200 * 1024) return ISG_EUNSUPPORTED;
  ISG_CUDA(cudaFuncSetAttribute(assign_sparse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(cdiv(cap, kSparseThreads), B);
  assign_sparse_kernel<<<grid, kSparseThreads, smem, stream>>>(
      ae, img_stride, plane_stride, idx, count, cap, reinterpret_cast<const SeedRec*>(seeds),
      reinterpret_cast<const float4*>(ghost), n_seeds, Nmax, W, ys, xs, label, score, flag, stats);
  ISG_LAUNCH_CHECK();
  return ISG_OK;
}

template <int RW>
static int launch_dense(const float* kp, int64_t kp_img_stride, const float* ae, int64_t ae_img_stride,
                        int64_t ae_plane_stride, const uint32_t* thr_key, const uint32_t* seeds, const float* ghost,
                        const int32_t* n_seeds, int B, int Nmax, int H, int W, const float* ys, const float* xs,
                        int32_t* label_map, float* score_map, uint32_t* keepbits, int32_t* stats, bool vec,
                        cudaStream_t stream) {
  const size_t smem = (size_t)Nmax * sizeof(SeedRec) * 2;
  dim3 block(32, kDenseWarps), grid(cdiv(W, 128), cdiv(H, kDenseWarps * RW), B);
  const int Wwords = cdiv(W, 32);
#define ISG_DENSE_LAUNCH(VEC_, SCORE_)                                                                          \
  do {                                                                                                          \
    ISG_CUDA(cudaFuncSetAttribute(assign_dense_kernel<RW, VEC_, SCORE_>,                                        \
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                     \
    assign_dense_kernel<RW, VEC_, SCORE_><<<grid, block, smem, stream>>>(                                       \
        kp, kp_img_stride, ae, ae_img_stride, ae_plane_stride, thr_key, reinterpret_cast<const SeedRec*>(seeds), \
        reinterpret_cast<const float4*>(ghost), n_seeds, Nmax, H, W, Wwords, ys, xs, label_map, score_map,      \
        keepbits, stats);                                                                                       \
  } while (0)
  if (vec) { if (score_map) ISG_DENSE_LAUNCH(true, true); else ISG_DENSE_LAUNCH(true, false); }
  else     { if (score_map) ISG_DENSE_LAUNCH(false, true); else ISG_DENSE_LAUNCH(false, false); }
#undef ISG_DENSE_LAUNCH
  ISG_LAUNCH_CHECK();
  return ISG_OK;
}

extern "C" size_t isg_assign_dense_workspace_bytes(int B, int Nmax, int H, int W) {
  if (B <= 0 || Nmax <= 0 || H <= 0 || W <= 0) return 0;
  return dense_workspace_bytes(B, Nmax, H, W);
}

extern "C" int isg_build_tile_lists(const uint32_t* seeds, const int32_t* n_seeds, int B, int Nmax, int H, int W,
                                    void* workspace, size_t workspace_bytes, isg_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!seeds || !n_seeds || B <= 0 || Nmax <= 0 || H <= 0 || W <= 0 || B > 65535) return ISG_EINVAL;
  if (!aligned16(seeds)) return ISG_EINVAL;
  if (!workspace || workspace_bytes < dense_workspace_bytes(B, Nmax, H, W) || !aligned16(workspace)) return ISG_EINVAL;
  if (W % 4 != 0) return ISG_OK;      // the tensor-map kernel is not used for this width; nothing to prepare
  const int rc = launch_dense_v4(nullptr, 0, nullptr, 0, 0, nullptr, seeds, nullptr, n_seeds, B, Nmax, H, W, nullptr, nullptr,
                                 nullptr, nullptr, nullptr, nullptr, workspace, workspace_bytes, 1, stream);
  return rc == ISG_EUNSUPPORTED ? ISG_OK : rc;   // unsupported geometry: isg_assign_dense falls back and needs no lists
}

extern "C" int isg_assign_labels(const float* ae, int64_t ae_img_stride, int64_t ae_plane_stride, const uint32_t* seeds,
                                 const float* ghost, const int32_t* n_seeds, int B, int Nmax, int H, int W, const float* ys,
                                 const float* xs, int32_t* label_map, float* score_map, void* workspace,
                                 size_t workspace_bytes, int lists_prebuilt, isg_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!ae || !seeds || !ghost || !n_seeds || !ys || !xs || !label_map) return ISG_EINVAL;
  if (B <= 0 || Nmax <= 0 || H <= 0 || W <= 0 || B > 65535) return ISG_EINVAL;
  if (ae_plane_stride < (int64_t)H * W) return ISG_EINVAL;
  if (!aligned16(seeds) || !aligned16(ghost)) return ISG_EINVAL;
  if (!workspace || workspace_bytes < dense_workspace_bytes(B, Nmax, H, W) || !aligned16(workspace)) return ISG_EINVAL;
  if ((size_t)Nmax * sizeof(SeedRec) * 2 > 200 * 1024) return ISG_EUNSUPPORTED;
  const bool vec = (W % 4 == 0) && (ae_img_stride % 4 == 0) && (ae_plane_stride % 4 == 0) && aligned16(ae) &&
                   aligned16(label_map) && (!score_map || aligned16(score_map));
  if (!vec || tuning().dense_v1) return ISG_EUNSUPPORTED;
  return launch_dense_v4(nullptr, 0, ae, ae_img_stride, ae_plane_stride, nullptr, seeds, ghost, n_seeds, B, Nmax, H, W, ys, xs,
                         label_map, score_map, nullptr, nullptr, workspace, workspace_bytes, lists_prebuilt ? 2 : 0, stream,
                         /*with_kp=*/false);
}

extern "C" int isg_assign_dense(const float* kp, int64_t kp_img_stride, const float* ae, int64_t ae_img_stride,
                                int64_t ae_plane_stride, const uint32_t* thr_key, const uint32_t* seeds,
                                const float* ghost, const int32_t* n_seeds, int B, int Nmax, int H, int W,
                                const float* ys, const float* xs, int32_t* label_map, float* score_map,
                                uint32_t* keepbits, int32_t* stats, void* workspace, size_t workspace_bytes,
                                int lists_prebuilt, isg_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!kp || !ae || !thr_key || !seeds || !ghost || !n_seeds || !ys || !xs || !label_map || !keepbits) return ISG_EINVAL;
  if (B <= 0 || Nmax <= 0 || H <= 0 || W <= 0 || B > 65535) return ISG_EINVAL;
  if (kp_img_stride < (int64_t)H * W || ae_plane_stride < (int64_t)H * W) return ISG_EINVAL;
  if (!aligned16(seeds) || !aligned16(ghost)) return ISG_EINVAL;
  if (!workspace || workspace_bytes < dense_workspace_bytes(B, Nmax, H, W) || !aligned16(workspace)) return ISG_EINVAL;
  if ((size_t)Nmax * sizeof(SeedRec) * 2 > 200 * 1024) return ISG_EUNSUPPORTED;
  const bool vec = (W % 4 == 0) && (kp_img_stride % 4 == 0) && (ae_img_stride % 4 == 0) && (ae_plane_stride % 4 == 0) &&
                   aligned16(kp) && aligned16(ae) && aligned16(label_map) && (!score_map || aligned16(score_map));
  // v4 (persistent, TMA-fed, dynamic scheduler) whenever the layout allows tensor maps; the plain-LDG kernel serves the
  // layouts TMA cannot describe (W % 4 != 0) and, with ISG_DENSE_V1=1, A/B measurements
  const Tuning& tn = tuning();
  if (vec && !tn.dense_v1) {
    const int rc = launch_dense_v4(kp, kp_img_stride, ae, ae_img_stride, ae_plane_stride, thr_key, seeds, ghost, n_seeds, B,
                                   Nmax, H, W, ys, xs, label_map, score_map, keepbits, stats, workspace, workspace_bytes,
                                   lists_prebuilt ? 2 : 0, stream);
    if (rc != ISG_EUNSUPPORTED) return rc;
  }
  const int rw = tn.dense_v1_rw;
  if (rw == 2)
    return launch_dense<2>(kp, kp_img_stride, ae, ae_img_stride, ae_plane_stride, thr_key, seeds, ghost, n_seeds, B,
                           Nmax, H, W, ys, xs, label_map, score_map, keepbits, stats, vec, stream);
  if (rw == 1)
    return launch_dense<1>(kp, kp_img_stride, ae, ae_img_stride, ae_plane_stride, thr_key, seeds, ghost, n_seeds, B,
                           Nmax, H, W, ys, xs, label_map, score_map, keepbits, stats, vec, stream);
  return launch_dense<4>(kp, kp_img_stride, ae, ae_img_stride, ae_plane_stride, thr_key, seeds, ghost, n_seeds, B,
                         Nmax, H, W, ys, xs, label_map, score_map, keepbits, stats, vec, stream);
}

namespace isg {
__global__ void __launch_bounds__(256)
scatter_labels_kernel(const int32_t* __restrict__ idx, const int32_t* __restrict__ count, int cap,
                      const int32_t* __restrict__ label, int H, int W, int32_t* __restrict__ label_map) {
  const int b = blockIdx.y;
  const int m = blockIdx.x * 256 + threadIdx.x;
  if (m >= min(count[b], cap)) return;
  const size_t o = (size_t)b * cap + m;
  const int y = idx[o * 2], x = idx[o * 2 + 1];
  label_map[((size_t)b * H + y) * W + x] = label[o];
}
}  // namespace isg

namespace isg {
__global__ void __launch_bounds__(256)
gather_embeddings_kernel(const float* __restrict__ ae, int64_t img_stride, int64_t plane_stride, const int32_t* __restrict__ idx,
                         const int32_t* __restrict__ count, int cap, int W, const float* __restrict__ ys,
                         const float* __restrict__ xs, float2* __restrict__ emb) {
  const int b = blockIdx.y;
  const int m = blockIdx.x * 256 + threadIdx.x;
  if (m >= min(count[b], cap)) return;
  const size_t o = (size_t)b * cap + m;
  const int y = idx[o * 2], x = idx[o * 2 + 1];
  const float* p = ae + (int64_t)b * img_stride + (int64_t)y * W + x;
  // e = tanh(ae[0:2]) + xym (utils/decode.py:305), the same arithmetic as the assignment kernels
  emb[o] = make_float2(__fadd_rn(tanh_fast(__ldg(p)), ys[y]), __fadd_rn(tanh_fast(__ldg(p + plane_stride)), xs[x]));
}
}  // namespace isg

extern "C" int isg_gather_embeddings(const float* ae, int64_t img_stride, int64_t plane_stride, const int32_t* idx,
                                     const int32_t* count, int cap, int B, int H, int W, const float* ys, const float* xs,
                                     float* emb, isg_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!ae || !idx || !count || !ys || !xs || !emb || B <= 0 || cap <= 0 || H <= 0 || W <= 0 || B > 65535) return ISG_EINVAL;
  if (((uintptr_t)emb & 7) != 0) return ISG_EINVAL;
  isg::gather_embeddings_kernel<<<dim3(cdiv(cap, 256), B), 256, 0, stream>>>(ae, img_stride, plane_stride, idx, count, cap, W, ys,
                                                                              xs, reinterpret_cast<float2*>(emb));
  ISG_LAUNCH_CHECK();
  return ISG_OK;
}

extern "C" int isg_scatter_labels(const int32_t* idx, const int32_t* count, int cap, const int32_t* label, int B, int H,
                                  int W, int32_t* label_map, isg_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!idx || !count || !label || !label_map || B <= 0 || cap <= 0 || H <= 0 || W <= 0 || B > 65535) return ISG_EINVAL;
  isg::scatter_labels_kernel<<<dim3(cdiv(cap, 256), B), 256, 0, stream>>>(idx, count, cap, label, H, W, label_map);
  ISG_LAUNCH_CHECK();
  return ISG_OK;
}

extern "C" int isg_gather_labels(const int32_t* label_map, const float* score_map, const int32_t* idx,
                                 const int32_t* count, int cap, const float* ghost, int B, int Nmax, int H, int W,
                                 int32_t* label, float* score, uint8_t* flag, int32_t* stats, isg_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!label_map || !idx || !count || !ghost || !label || B <= 0 || cap <= 0 || H <= 0 || W <= 0 || Nmax <= 0) return ISG_EINVAL;
  dim3 grid(cdiv(cap, 256), B);
  gather_labels_kernel<<<grid, 256, 0, stream>>>(label_map, score_map, idx, count, cap,
                                                 reinterpret_cast<const float4*>(ghost), Nmax, H, W, label, score, flag, stats);
  ISG_LAUNCH_CHECK();
  return ISG_OK;
}

extern "C" int isg_group_points(const int32_t* idx, const int32_t* label, const uint8_t* flag, const int32_t* count,
                                int cap, const int32_t* n_seeds, int B, int Nmax, int32_t* offsets, float* points,
                                isg_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!idx || !label || !flag || !count || !n_seeds || !offsets || !points) return ISG_EINVAL;
  if (B <= 0 || Nmax <= 0 || cap <= 0 || B > 65535) return ISG_EINVAL;
  const size_t smem = ((size_t)kSplitWarps * Nmax + Nmax + 1 + cap) * sizeof(int);
  if (smem <= 200 * 1024) {
    ISG_CUDA(cudaFuncSetAttribute(group_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    group_split_kernel<<<B, 32 * kSplitWarps, smem, stream>>>(idx, label, flag, count, cap, n_seeds, Nmax, offsets, points);
  } else {   // very large seed tables: one warp per instance scans the list
    dim3 block(32, kGroupWarps), grid(cdiv(Nmax, kGroupWarps), B);
    group_count_kernel<<<grid, block, 0, stream>>>(label, flag, count, cap, n_seeds, Nmax, offsets);
    group_scan_kernel<<<B, 32, 0, stream>>>(offsets, Nmax);
    group_scatter_kernel<<<grid, block, 0, stream>>>(idx, label, flag, count, cap, n_seeds, Nmax, offsets, points);
  }
  ISG_LAUNCH_CHECK();
  return ISG_OK;
}
