// Fused dense kernel, v3: persistent CTAs fed by the TMA unit, row-walking consumers.
//
// What one launch does for every pixel of the batch (reference: select_points / nms_hm, utils/decode.py:42-48,71-85,
// and the core of group_kp, :303-328):
//   keep bit  = selected (kp >= k-th largest) and 3x3 maximum of the thresholded map,
//   embedding = tanh(ae[0:2]) + grid, sigma = exp(ae[2:4]),
//   label     = first seed (in index order) with the largest membership exp(-q) among the seeds whose box
//               contains the pixel, q = (e_y-c_y)^2*s_y + (e_x-c_x)^2*s_x; 0 when no membership is > 0,
//   per-instance count / bbox of the keep pixels that pass the ghost filter.
// Every input pixel is read from HBM once (kp halo rows/columns are re-read from L2), the label is written once.
//
// Structure
//   * one CTA per SM; the last warp is the producer, the others are consumers in G groups of WG warps;
//   * a tile is 128 x (RW*WG) pixels: one 3-D tensor-map box of kp with a 1-pixel halo (out-of-image elements
//     are filled with NaN, which fmax ignores - that is the reference's -inf padding) and one 4-D box with the
//     four ae planes; the producer keeps `nstages` tiles in flight (cp.async.bulk.tensor + mbarrier);
//   * group g consumes the CTA's tiles k = g (mod G); inside a tile a warp owns RW consecutive rows and walks
//     down them with a rolling 3-row window of the separable 3x3 maximum, 4 consecutive pixels per lane;
//   * the membership loop tracks the SMALLEST exponent q instead of the largest exp(-q): exp is monotone, so
//     the winner is the same seed whenever two memberships differ as fp32 numbers, and it saves the
//     transcendental per (pixel, seed) pair.  q >= ln(2^150) is where exp(-q) rounds to 0 in fp32 (label 0);
//   * the seed tables of the images a CTA touches are staged by the producer with 1-D bulk copies into two
//     buffers (image parity), guarded by their own full/empty mbarriers.
#pragma once
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include "keep.cuh"

namespace isg {

constexpr int kV3TileW = 128;
constexpr int kV3KpW = kV3TileW + 8;          // 4 columns of padding on each side keep the box 16-byte granular
constexpr int kV3MaxStages = 12;
constexpr float kQZero = 103.97207708f;       // ln(2^150): exp(-q) == 0 in fp32 (round to nearest, subnormals kept)

template <int RW, int WG>
struct V3Geom {
  static constexpr int TH = RW * WG;                                  // tile rows
  static constexpr int kKpRows = TH + 2;
  static constexpr int kKpBytes = kV3KpW * kKpRows * 4;
  static constexpr int kKpStage = (kKpBytes + 127) / 128 * 128;
  static constexpr int kAePlane = kV3TileW * TH;                      // floats
  static constexpr int kAeBytes = kAePlane * 16;
  static constexpr int kStage = kKpStage + kAeBytes;
  static constexpr uint32_t kTx = kKpBytes + kAeBytes;
};

struct RowK {
  float h[4];      // horizontal 3-max of v centred on the lane's 4 pixels
  float v[4];      // thresholded value (0 when not selected, NaN outside the image)
  uint32_t sel;    // bit i: pixel i is selected
};

// One staged kp row -> RowK.  `row` points at the first float of the box row; the lane's pixels start at
// row[4 + 4*lane]; row[3] / row[132] are the halo pixels of lane 0 / lane 31 (`halo_off` selects one per lane).
template <bool USE_INT>
__device__ __forceinline__ RowK prep_row(const float* __restrict__ row, int lane, int halo_off, const Thr& thr) {
  RowK o;
  const float4 t = *reinterpret_cast<const float4*>(row + 4 + lane * 4);
  const float hv = row[halo_off];
  const float raw[4] = {t.x, t.y, t.z, t.w};
  o.sel = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const bool below = USE_INT ? (skey(raw[i]) < thr.s) : (raw[i] < thr.f);   // NaN (outside) is never below
    o.v[i] = below ? 0.0f : raw[i];
    o.sel |= (below ? 0u : 1u) << i;
  }
  const bool hbelow = USE_INT ? (skey(hv) < thr.s) : (hv < thr.f);
  const float hvv = hbelow ? 0.0f : hv;
  float left = __shfl_up_sync(0xffffffffu, o.v[3], 1);
  float right = __shfl_down_sync(0xffffffffu, o.v[0], 1);
  if (lane == 0) left = hvv;
  if (lane == 31) right = hvv;
  o.h[0] = fmaxf(fmaxf(left, o.v[0]), o.v[1]);
  o.h[1] = fmaxf(fmaxf(o.v[0], o.v[1]), o.v[2]);
  o.h[2] = fmaxf(fmaxf(o.v[1], o.v[2]), o.v[3]);
  o.h[3] = fmaxf(fmaxf(o.v[2], o.v[3]), right);
  return o;
}

// A consumer warp is done with the images of ordinals [o0, o1) of its CTA (ordinal = image - first image of the
// CTA's tile range; buffer = ordinal & 1).  An image the warp never visited is first waited for, so that an arrival
// can never land in an earlier phase of the same buffer's `empty` barrier.
__device__ __forceinline__ void leave_images(uint64_t* seed_full, uint64_t* seed_empty, int o0, int o1) {
  for (int o = o0; o < o1; ++o) {
    mbar_wait(&seed_full[o & 1], (uint32_t)((o >> 1) & 1));
    mbar_arrive(&seed_empty[o & 1]);
  }
}

template <int RW, int WG, int G, bool SCORE>
__global__ void __launch_bounds__(32 * (WG * G + 1), 1)
dense_v3_kernel(const __grid_constant__ CUtensorMap tm_kp, const __grid_constant__ CUtensorMap tm_ae,
                const uint32_t* __restrict__ thr_key, const SeedRec* __restrict__ seeds,
                const float4* __restrict__ ghost, const int32_t* __restrict__ n_seeds, int Nmax, int B, int H, int W,
                int Wwords, int tilesX, int tilesY, int nstages, const float* __restrict__ ys,
                const float* __restrict__ xs, int32_t* __restrict__ label_map, float* __restrict__ score_map,
                uint32_t* __restrict__ keepbits, int32_t* __restrict__ stats) {
  using Geo = V3Geom<RW, WG>;
  constexpr int kConsumers = WG * G;
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* stage_base = smem;
  unsigned char* p_after = smem + (size_t)nstages * Geo::kStage;
  SeedRec* s_seed = reinterpret_cast<SeedRec*>(p_after);                 // [2][Nmax]
  p_after += (size_t)2 * Nmax * sizeof(SeedRec);
  uint16_t* s_list = reinterpret_cast<uint16_t*>(p_after);               // [consumers][2][Nmax]
  p_after += (((size_t)kConsumers * 2 * Nmax * 2 + 15) & ~(size_t)15);
  uint64_t* full = reinterpret_cast<uint64_t*>(p_after);                 // [kV3MaxStages]
  uint64_t* empty = full + kV3MaxStages;                                 // [kV3MaxStages]
  uint64_t* seed_full = empty + kV3MaxStages;                            // [2]
  uint64_t* seed_empty = seed_full + 2;                                  // [2]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long T = (long long)B * tilesY * tilesX;
  const int t_begin = (int)(T * blockIdx.x / gridDim.x), t_end = (int)(T * (blockIdx.x + 1) / gridDim.x);
  const int tiles_per_img = tilesX * tilesY;
  const int b_first = t_begin / tiles_per_img;
  const int b_last = (t_end - 1) / tiles_per_img;      // only used when t_end > t_begin

  if (threadIdx.x == 0) {
    for (int s = 0; s < nstages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], WG); }
    for (int i = 0; i < 2; ++i) { mbar_init(&seed_full[i], 1); mbar_init(&seed_empty[i], kConsumers); }
    mbar_fence_init();
  }
  __syncthreads();
  if (t_begin >= t_end) return;

  if (warp == kConsumers) {
    // ===== producer: one elected lane feeds the tile ring and the seed-table buffers =====
    if (lane == 0) {
      int b = b_first;
      int rem = t_begin - b * tiles_per_img;
      int ty = rem / tilesX, tx = rem - ty * tilesX;
      int staged_b = -1;
      for (int t = t_begin, k = 0; t < t_end; ++t, ++k) {
        if (b != staged_b) {
          const int ord = b - b_first, buf = ord & 1, use = ord >> 1;
          if (use >= 1) mbar_wait_backoff(&seed_empty[buf], (uint32_t)((use - 1) & 1), 100);
          const int n = min(n_seeds[b], Nmax);
          mbar_expect_tx(&seed_full[buf], (uint32_t)(n * sizeof(SeedRec)));
          if (n > 0) bulk_g2s(s_seed + (size_t)buf * Nmax, seeds + (size_t)b * Nmax, (uint32_t)(n * sizeof(SeedRec)), &seed_full[buf]);
          staged_b = b;
        }
        const int s = k % nstages;
        if (k >= nstages) mbar_wait_backoff(&empty[s], (uint32_t)((k / nstages - 1) & 1), 100);
        unsigned char* st = stage_base + (size_t)s * Geo::kStage;
        mbar_expect_tx(&full[s], Geo::kTx);
        tma_load_3d(st, &tm_kp, tx * kV3TileW - 4, ty * Geo::TH - 1, b, &full[s]);
        tma_load_4d(st + Geo::kKpStage, &tm_ae, tx * kV3TileW, ty * Geo::TH, 0, b, &full[s]);
        if (++tx == tilesX) { tx = 0; if (++ty == tilesY) { ty = 0; ++b; } }
      }
    }
    return;
  }

  // ===== consumers =====
  const int grp = warp / WG, wl = warp % WG;
  uint16_t* my_band = s_list + (size_t)warp * 2 * Nmax;   // seeds overlapping this warp's rows of the current band
  uint16_t* my_hit = my_band + Nmax;                      // ... and the columns of the current tile
  const int halo_off = (lane == 31) ? (4 + kV3TileW) : 3;
  const SeedRec* s_all = s_seed;
  int cur_b = b_first - 1, cur_ty = -1, n = 0, Tc = 0;
  Thr thr = make_thr(0xffffffffu);
  int b, tx, ty;
  {
    const int t0 = t_begin + grp;
    b = t0 / tiles_per_img;
    const int rem = t0 - b * tiles_per_img;
    ty = rem / tilesX; tx = rem - ty * tilesX;
  }
  int ybeg = 0;
  for (int t = t_begin + grp, k = grp; t < t_end; t += G, k += G) {
    const int s = k % nstages;
    if (b != cur_b) {
      // leave the images up to b-1 (frees their seed buffers), then wait for the table of image b
      __syncwarp();
      if (lane == 0) leave_images(seed_full, seed_empty, max(cur_b, b_first) - b_first, b - b_first);
      const int ord = b - b_first;
      mbar_wait(&seed_full[ord & 1], (uint32_t)((ord >> 1) & 1));
      s_all = s_seed + (size_t)(ord & 1) * Nmax;
      n = min(__ldg(n_seeds + b), Nmax);
      thr = make_thr(__ldg(thr_key + b));
      cur_b = b; cur_ty = -1;
    }
    if (ty != cur_ty) {
      // new band: ordered list of the seeds whose boxes overlap this warp's rows
      ybeg = ty * Geo::TH + wl * RW;
      const int sy0 = ybeg, sy1 = min(ybeg + RW - 1, H - 1);
      Tc = 0;
      for (int j0 = 0; j0 < n; j0 += 32) {
        const int j = j0 + lane;
        bool hit = false;
        if (j < n) { const int4 bx = *reinterpret_cast<const int4*>(&s_all[j]); hit = bx.x <= sy1 && bx.y >= sy0 && bx.z <= bx.w; }
        const unsigned bal = __ballot_sync(0xffffffffu, hit);
        if (hit) my_band[Tc + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)j;
        Tc += __popc(bal);
      }
      __syncwarp();
      cur_ty = ty;
    }
    const int x0 = tx * kV3TileW + lane * 4;
    const bool colvalid = x0 < W;                       // W % 4 == 0 on this path: a lane is all in or all out

    // (1) ordered culling of the band list against this tile's columns
    int Tw = 0;
    {
      const int sx0 = tx * kV3TileW, sx1 = min(sx0 + kV3TileW - 1, W - 1);
      for (int j0 = 0; j0 < Tc; j0 += 32) {
        const int q = j0 + lane;
        bool hit = false;
        int j = 0;
        if (q < Tc) { j = my_band[q]; const int2 bxx = *reinterpret_cast<const int2*>(&s_all[j].x0); hit = bxx.x <= sx1 && bxx.y >= sx0; }
        const unsigned bal = __ballot_sync(0xffffffffu, hit);
        if (hit) my_hit[Tw + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)j;
        Tw += __popc(bal);
      }
      __syncwarp();
    }
    float xs4[4] = {0.f, 0.f, 0.f, 0.f};
    if (colvalid) { const float4 v = __ldg(reinterpret_cast<const float4*>(xs + x0)); xs4[0] = v.x; xs4[1] = v.y; xs4[2] = v.z; xs4[3] = v.w; }

    // (2) wait for the tile
    mbar_wait(&full[s], (uint32_t)((k / nstages) & 1));
    const float* kbox = reinterpret_cast<const float*>(stage_base + (size_t)s * Geo::kStage);
    const float* abox = reinterpret_cast<const float*>(stage_base + (size_t)s * Geo::kStage + Geo::kKpStage);

    // (3) walk down the warp's rows; box row = tile row + 1
    const float* krow = kbox + (wl * RW) * kV3KpW;
    RowK up, mid;
    if (thr.use_int) { up = prep_row<true>(krow, lane, halo_off, thr); mid = prep_row<true>(krow + kV3KpW, lane, halo_off, thr); }
    else { up = prep_row<false>(krow, lane, halo_off, thr); mid = prep_row<false>(krow + kV3KpW, lane, halo_off, thr); }
    const float* arow = abox + (wl * RW) * kV3TileW + lane * 4;
    int32_t* lrow = label_map + ((size_t)b * H + ybeg) * W + x0;
    float* srow = SCORE ? score_map + ((size_t)b * H + ybeg) * W + x0 : nullptr;
    uint32_t* kbrow = keepbits + ((size_t)b * H + ybeg) * Wwords + (x0 >> 5);
#pragma unroll 1
    for (int r = 0; r < RW; ++r) {
      const int y = ybeg + r;
      if (y >= H) break;                                  // warp-uniform (ragged bottom)
      // --- keep bits ---
      RowK dn;
      if (thr.use_int) dn = prep_row<true>(krow + (r + 2) * kV3KpW, lane, halo_off, thr);
      else dn = prep_row<false>(krow + (r + 2) * kV3KpW, lane, halo_off, thr);
      uint32_t nib = 0;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float m = fmaxf(fmaxf(up.h[i], mid.h[i]), dn.h[i]);
        if (mid.v[i] >= m) nib |= 1u << i;
      }
      nib &= mid.sel;
      if (!colvalid) nib = 0;
      {
        const uint32_t word = nibbles_to_word(nib, lane);
        if ((lane & 7) == 0 && colvalid) kbrow[(size_t)r * Wwords] = word;
      }
      up = mid; mid = dn;

      // --- embedding of the lane's 4 pixels ---
      float ey[4], ex[4], sy[4], sx[4];
      {
        const float4 a0 = *reinterpret_cast<const float4*>(arow + r * kV3TileW);
        const float4 a1 = *reinterpret_cast<const float4*>(arow + Geo::kAePlane + r * kV3TileW);
        const float4 a2 = *reinterpret_cast<const float4*>(arow + 2 * Geo::kAePlane + r * kV3TileW);
        const float4 a3 = *reinterpret_cast<const float4*>(arow + 3 * Geo::kAePlane + r * kV3TileW);
        const float t0[4] = {a0.x, a0.y, a0.z, a0.w}, t1[4] = {a1.x, a1.y, a1.z, a1.w};
        const float t2[4] = {a2.x, a2.y, a2.z, a2.w}, t3[4] = {a3.x, a3.y, a3.z, a3.w};
        float amax = 0.0f;
#pragma unroll
        for (int i = 0; i < 4; ++i) amax = fmaxf(amax, fmaxf(fabsf(t0[i]), fabsf(t1[i])));
        const bool small = __all_sync(0xffffffffu, amax < 0.55f);   // warp-uniform: polynomial branch only
        const float yv = __ldg(ys + y);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          ey[i] = __fadd_rn(small ? tanh_poly(t0[i]) : tanh_fast(t0[i]), yv);
          ex[i] = __fadd_rn(small ? tanh_poly(t1[i]) : tanh_fast(t1[i]), xs4[i]);
          sy[i] = exp_fast(t2[i]);
          sx[i] = exp_fast(t3[i]);
        }
      }

      // --- membership: smallest exponent among the seeds whose box contains the pixel, ascending seed index ---
      float bq[4] = {kQZero, kQZero, kQZero, kQZero};
      int lab[4] = {0, 0, 0, 0};
      for (int q = 0; q < Tw; ++q) {
        const int id = my_hit[q];
        const int4 bx = *reinterpret_cast<const int4*>(&s_all[id]);          // y0,y1,x0,x1 (broadcast)
        if (y < bx.x || y > bx.y) continue;                                  // warp-uniform row test
        const float2 cc = *reinterpret_cast<const float2*>(&s_all[id].cy);
        const unsigned hx = (unsigned)(bx.w - bx.z);
        const int xr = x0 - bx.z;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float dy = __fsub_rn(ey[i], cc.x), dx = __fsub_rn(ex[i], cc.y);
          const float qy = __fmul_rn(__fmul_rn(dy, dy), sy[i]);
          const float qx = __fmul_rn(__fmul_rn(dx, dx), sx[i]);
          const float qq = __fadd_rn(qy, qx);
          if (((unsigned)(xr + i) <= hx) && qq < bq[i]) { bq[i] = qq; lab[i] = id; }   // strict: first index wins ties (:328)
        }
      }

      // --- stores + statistics of the keep pixels ---
      if (colvalid) {
        stg_stream4(lrow + (size_t)r * W, lab[0], lab[1], lab[2], lab[3]);
        if (SCORE) {
          float p[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) p[i] = (bq[i] < kQZero) ? exp_fast(-bq[i]) : 0.0f;
          stg_stream4f(srow + (size_t)r * W, p[0], p[1], p[2], p[3]);
        }
      }
      if (nib && stats && n > 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if ((nib >> i) & 1u) {
            const int l = lab[i];
            if (ghost_pass(ghost[(size_t)b * Nmax + l], y, x0 + i))
              stats_add(stats + ((size_t)b * Nmax + l) * ISG_STAT_WORDS, y, x0 + i);
          }
      }
    }
    // every read of the stage (and of my_hit) is done: hand the stage back to the producer
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
    tx += G;
    while (tx >= tilesX) { tx -= tilesX; if (++ty == tilesY) { ty = 0; ++b; } }
  }
  // leave every remaining image of this CTA's range (the producer may be waiting to reuse a seed buffer)
  __syncwarp();
  if (lane == 0) leave_images(seed_full, seed_empty, max(cur_b, b_first) - b_first, b_last - b_first + 1);
}

// ---- host side -------------------------------------------------------------------------------
template <int RW, int WG, int G>
inline size_t dense_v3_smem_bytes(int Nmax, int nstages) {
  return (size_t)nstages * V3Geom<RW, WG>::kStage + (size_t)2 * Nmax * sizeof(SeedRec) +
         (((size_t)WG * G * 2 * Nmax * 2 + 15) & ~(size_t)15) + (2 * kV3MaxStages + 4) * sizeof(uint64_t);
}

template <int RW, int WG, int G>
inline int launch_dense_v3_cfg(const float* kp, int64_t kp_img_stride, const float* ae, int64_t ae_img_stride,
                               int64_t ae_plane_stride, const uint32_t* thr_key, const uint32_t* seeds, const float* ghost,
                               const int32_t* n_seeds, int B, int Nmax, int H, int W, const float* ys, const float* xs,
                               int32_t* label_map, float* score_map, uint32_t* keepbits, int32_t* stats, int max_stages,
                               cudaStream_t stream) {
  using Geo = V3Geom<RW, WG>;
  if (Nmax > 65535) return ISG_EUNSUPPORTED;
  int nstages = std::min(max_stages, kV3MaxStages);
  while (nstages > 1 && dense_v3_smem_bytes<RW, WG, G>(Nmax, nstages) > 227 * 1024) --nstages;
  if (nstages < G + 1) return ISG_EUNSUPPORTED;
  static isg_encode_tiled_fn encode = get_encode_tiled();
  if (!encode) return ISG_EUNSUPPORTED;
  CUtensorMap tm_kp, tm_ae;
  {
    const cuuint64_t dim[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    const cuuint64_t str[2] = {(cuuint64_t)W * 4, (cuuint64_t)kp_img_stride * 4};
    const cuuint32_t box[3] = {kV3KpW, (cuuint32_t)Geo::kKpRows, 1};
    const cuuint32_t es[3] = {1, 1, 1};
    if (encode(&tm_kp, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(kp), dim, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA) != CUDA_SUCCESS)
      return ISG_EUNSUPPORTED;
  }
  {
    const cuuint64_t dim[4] = {(cuuint64_t)W, (cuuint64_t)H, 4, (cuuint64_t)B};
    const cuuint64_t str[3] = {(cuuint64_t)W * 4, (cuuint64_t)ae_plane_stride * 4, (cuuint64_t)ae_img_stride * 4};
    const cuuint32_t box[4] = {kV3TileW, (cuuint32_t)Geo::TH, 4, 1};
    const cuuint32_t es[4] = {1, 1, 1, 1};
    if (encode(&tm_ae, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(ae), dim, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return ISG_EUNSUPPORTED;
  }
  const int tilesX = cdiv(W, kV3TileW), tilesY = cdiv(H, Geo::TH);
  const long long T = (long long)B * tilesX * tilesY;
  int dev = 0, sms = kSMs;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = (int)std::min<long long>(T, sms);
  const size_t smem = dense_v3_smem_bytes<RW, WG, G>(Nmax, nstages);
  const int Wwords = cdiv(W, 32);
  const int threads = 32 * (WG * G + 1);
  if (score_map) {
    ISG_CUDA(cudaFuncSetAttribute(dense_v3_kernel<RW, WG, G, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dense_v3_kernel<RW, WG, G, true><<<grid, threads, smem, stream>>>(
        tm_kp, tm_ae, thr_key, reinterpret_cast<const SeedRec*>(seeds), reinterpret_cast<const float4*>(ghost), n_seeds, Nmax,
        B, H, W, Wwords, tilesX, tilesY, nstages, ys, xs, label_map, score_map, keepbits, stats);
  } else {
    ISG_CUDA(cudaFuncSetAttribute(dense_v3_kernel<RW, WG, G, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dense_v3_kernel<RW, WG, G, false><<<grid, threads, smem, stream>>>(
        tm_kp, tm_ae, thr_key, reinterpret_cast<const SeedRec*>(seeds), reinterpret_cast<const float4*>(ghost), n_seeds, Nmax,
        B, H, W, Wwords, tilesX, tilesY, nstages, ys, xs, label_map, score_map, keepbits, stats);
  }
  ISG_LAUNCH_CHECK();
  return ISG_OK;
}

// Tuning knob for experiments: ISG_DENSE_CFG = "<RW>x<WG>x<G>[:stages]" selects one of the compiled geometries.
inline int launch_dense_v3(const float* kp, int64_t kp_img_stride, const float* ae, int64_t ae_img_stride,
                           int64_t ae_plane_stride, const uint32_t* thr_key, const uint32_t* seeds, const float* ghost,
                           const int32_t* n_seeds, int B, int Nmax, int H, int W, const float* ys, const float* xs,
                           int32_t* label_map, float* score_map, uint32_t* keepbits, int32_t* stats, cudaStream_t stream) {
  int rw = 2, wg = 8, g = 2, st = kV3MaxStages;
  if (const char* e = getenv("ISG_DENSE_CFG")) {
    int a = 0, b_ = 0, c = 0, d = 0;
    const int got = sscanf(e, "%dx%dx%d:%d", &a, &b_, &c, &d);
    if (got >= 3) { rw = a; wg = b_; g = c; }
    if (got >= 4 && d > 0) st = d;
  }
#define ISG_V3_CASE(RW_, WG_, G_)                                                                                      \
  if (rw == RW_ && wg == WG_ && g == G_)                                                                               \
    return launch_dense_v3_cfg<RW_, WG_, G_>(kp, kp_img_stride, ae, ae_img_stride, ae_plane_stride, thr_key, seeds, ghost, \
                                             n_seeds, B, Nmax, H, W, ys, xs, label_map, score_map, keepbits, stats, st, stream);
  ISG_V3_CASE(2, 8, 2)
  ISG_V3_CASE(4, 4, 3)
  ISG_V3_CASE(4, 2, 6)
  ISG_V3_CASE(4, 2, 8)
  ISG_V3_CASE(2, 4, 4)
  ISG_V3_CASE(2, 4, 6)
  ISG_V3_CASE(8, 2, 3)
  ISG_V3_CASE(8, 1, 8)
  ISG_V3_CASE(4, 1, 12)
#undef ISG_V3_CASE
  return ISG_EUNSUPPORTED;
}

}  // namespace isg
