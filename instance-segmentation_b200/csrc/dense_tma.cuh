// Fused dense kernel, v2: persistent CTAs fed by the TMA unit.
//
//   * one CTA per SM; warp 8 is the producer, warps 0-7 are consumers;
//   * a tile is 128 x 16 pixels: one 3-D tensor-map box for kp (136 x 18 incl. the 1-pixel halo, 4 columns of
//     padding on the left keep the box 16-byte granular) and one 4-D box for the four ae planes;
//   * the producer keeps up to kStages tiles in flight (cp.async.bulk.tensor + mbarrier expect_tx), so the HBM
//     latency is hidden behind the consumers' arithmetic instead of being exposed once per CTA as in v1;
//   * consumers read the staged tile from shared memory (conflict-free 128-bit reads), cull the image's seed
//     table per warp (no CTA-wide barrier on the tile path), and write labels / keep bits straight to HBM;
//   * the seed table of the current image is staged once per image with a 1-D bulk copy.
// Out-of-image halo elements are zero-filled by the TMA unit; the reference pads with -inf, which is restored
// from the coordinates.
#pragma once
#include <cuda.h>
#ifndef ISG_TMA_GROUPS
#define ISG_TMA_GROUPS 2
#endif
#ifndef ISG_TMA_CONSUMER_BACKOFF_NS
#define ISG_TMA_CONSUMER_BACKOFF_NS 0
#endif
#ifndef ISG_TMA_PRODUCER_BACKOFF_NS
#define ISG_TMA_PRODUCER_BACKOFF_NS 200
#endif
#include "keep.cuh"

namespace isg {

constexpr int kTmaTileW = 128, kTmaTileH = 16;
constexpr int kTmaConsumers = 8;                      // consumer warps per group (one group works on one tile)
constexpr int kTmaGroups = ISG_TMA_GROUPS;            // consumer groups; group g takes the CTA's tiles k = g (mod groups)
constexpr int kTmaRW = kTmaTileH / kTmaConsumers;     // rows per warp and tile
constexpr int kKpBoxW = kTmaTileW + 8, kKpBoxH = kTmaTileH + 2;
constexpr int kKpBoxBytes = kKpBoxW * kKpBoxH * 4;                    // 9792
constexpr int kKpStageBytes = (kKpBoxBytes + 127) / 128 * 128;        // 9856
constexpr int kAePlaneFloats = kTmaTileW * kTmaTileH;                 // 2048
constexpr int kAeBoxBytes = kAePlaneFloats * 4 * 4;                   // 32768
constexpr int kStageBytes = kKpStageBytes + kAeBoxBytes;              // 42624 (multiple of 128)
constexpr uint32_t kStageTx = kKpBoxBytes + kAeBoxBytes;
constexpr int kTmaMaxStages = 4;

// mbarrier wait that backs off between polls, so a waiting warp does not eat the issue slots of working warps
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t phase, unsigned ns) {
  uint32_t done;
  while (true) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(done) : "r"(smem_u32(bar)), "r"(phase) : "memory");
    if (done) break;
    __nanosleep(ns);
  }
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      :: "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
      :: "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void group_bar_sync(int group) {   // named barrier 1 + group, 256 threads
  asm volatile("bar.sync %0, %1;" :: "r"(1 + group), "n"(32 * kTmaConsumers) : "memory");
}

// thresholded row from the staged kp box; `yimg` is the image row, `srow` the row inside the box
struct LaneCols { int nin; bool has_left, has_right; };
template <bool USE_INT>
__device__ __forceinline__ RowH smem_rowh(const float* __restrict__ box, int srow, int yimg, int H, const LaneCols& lc,
                                          const Thr& thr, int lane) {
  if (yimg < 0 || yimg >= H) return rowh_outside();   // warp-uniform
  const float* row = box + srow * kKpBoxW;
  const float4 t = *reinterpret_cast<const float4*>(row + 4 + lane * 4);
  const float raw[4] = {t.x, t.y, t.z, t.w};
  float lr = 0.0f, rr = 0.0f;
  if (lane == 0) lr = row[3];
  if (lane == 31) rr = row[4 + kTmaTileW];
  return make_rowh<USE_INT>(raw, lc.nin, lr, lc.has_left, rr, lc.has_right, thr, lane);
}

template <bool USE_INT>
__device__ __forceinline__ void keep_rows_smem(const float* __restrict__ kbox, int r0, int ybeg, int H, const LaneCols& lc,
                                               const Thr& thr, int lane, uint32_t (&nib)[kTmaRW]) {
  // r0 = tile row of this warp's first row; box row = tile row + 1
  RowH up = smem_rowh<USE_INT>(kbox, r0, ybeg - 1, H, lc, thr, lane);
  RowH mid = smem_rowh<USE_INT>(kbox, r0 + 1, ybeg, H, lc, thr, lane);
#pragma unroll
  for (int r = 0; r < kTmaRW; ++r) {
    const RowH dn = smem_rowh<USE_INT>(kbox, r0 + r + 2, ybeg + r + 1, H, lc, thr, lane);
    nib[r] = (ybeg + r < H) ? keep_nibble(up, mid, dn) : 0u;
    up = mid; mid = dn;
  }
}

template <bool SCORE>
__global__ void __launch_bounds__(32 * (kTmaConsumers * kTmaGroups + 1), 1)
assign_dense_tma_kernel(const __grid_constant__ CUtensorMap tm_kp, const __grid_constant__ CUtensorMap tm_ae,
                        const uint32_t* __restrict__ thr_key, const SeedRec* __restrict__ seeds,
                        const float4* __restrict__ ghost, const int32_t* __restrict__ n_seeds, int Nmax, int B, int H,
                        int W, int Wwords, int tilesX, int tilesY, int nstages, const float* __restrict__ ys,
                        const float* __restrict__ xs, int32_t* __restrict__ label_map, float* __restrict__ score_map,
                        uint32_t* __restrict__ keepbits, int32_t* __restrict__ stats) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* stage_base = smem;
  unsigned char* p_after = smem + (size_t)nstages * kStageBytes;
  SeedRec* s_all_base = reinterpret_cast<SeedRec*>(p_after);                                      // [groups][Nmax]
  p_after += (size_t)kTmaGroups * Nmax * sizeof(SeedRec);
  uint16_t* s_hit = reinterpret_cast<uint16_t*>(p_after);                                         // [groups*8][2][Nmax]
  p_after += (((size_t)kTmaGroups * kTmaConsumers * Nmax * 4 + 15) & ~(size_t)15);
  uint64_t* bars = reinterpret_cast<uint64_t*>(p_after);
  uint64_t* full = bars;                    // [nstages]
  uint64_t* empty = bars + kTmaMaxStages;   // [nstages]
  uint64_t* seedbars = bars + 2 * kTmaMaxStages;   // [groups]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long T = (long long)B * tilesY * tilesX;
  const int t_begin = (int)(T * blockIdx.x / gridDim.x), t_end = (int)(T * (blockIdx.x + 1) / gridDim.x);

  if (threadIdx.x == 0) {
    for (int s = 0; s < nstages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kTmaConsumers); }
    for (int g = 0; g < kTmaGroups; ++g) mbar_init(&seedbars[g], 1);
    mbar_fence_init();
  }
  __syncthreads();

  // Tile order inside an image is row-major (tx fastest; neighbouring tiles share DRAM pages).  Consecutive
  // tiles of a consumer warp share their image rows, so the seeds overlapping those rows are culled once per
  // tile row ("band"), and only the column test is left per tile.
  const int tiles_per_img = tilesX * tilesY;
  if (warp == kTmaConsumers * kTmaGroups) {
    // ===== producer: one elected lane feeds the ring =====
    if (lane == 0) {
      int b = t_begin / tiles_per_img;
      int rem = t_begin - b * tiles_per_img;
      int ty = rem / tilesX, tx = rem - ty * tilesX;
      for (int t = t_begin, k = 0; t < t_end; ++t, ++k) {
        const int s = k % nstages;
        if (k >= nstages) {
          if (ISG_TMA_PRODUCER_BACKOFF_NS > 0) mbar_wait_backoff(&empty[s], (uint32_t)((k / nstages - 1) & 1), ISG_TMA_PRODUCER_BACKOFF_NS);
          else mbar_wait(&empty[s], (uint32_t)((k / nstages - 1) & 1));
        }
        unsigned char* st = stage_base + (size_t)s * kStageBytes;
        mbar_expect_tx(&full[s], kStageTx);
        tma_load_3d(st, &tm_kp, tx * kTmaTileW - 4, ty * kTmaTileH - 1, b, &full[s]);
        tma_load_4d(st + kKpStageBytes, &tm_ae, tx * kTmaTileW, ty * kTmaTileH, 0, b, &full[s]);
        if (++tx == tilesX) { tx = 0; if (++ty == tilesY) { ty = 0; ++b; } }
      }
    }
    return;
  }

  // ===== consumers: group `grp` takes tiles k = grp (mod groups); its 8 warps split the tile's 16 rows =====
  const int grp = warp / kTmaConsumers, wl = warp % kTmaConsumers;
  SeedRec* s_all = s_all_base + (size_t)grp * Nmax;
  uint64_t* seedbar = &seedbars[grp];
  int cur_b = -1, cur_ty = -1, n = 0, Tc = 0;
  uint32_t seed_phase = 0;
  Thr thr = make_thr(0xffffffffu);
  uint16_t* my_band = s_hit + (size_t)warp * 2 * Nmax;  // seeds overlapping this warp's rows of the current band
  uint16_t* my_hit = my_band + Nmax;                    // ... and the columns of the current tile
  int b, tx, ty;
  {
    const int t0 = t_begin + grp;
    b = t0 / tiles_per_img;
    const int rem = t0 - b * tiles_per_img;
    ty = rem / tilesX; tx = rem - ty * tilesX;
  }
  int ybeg = 0;
  for (int t = t_begin + grp, k = grp; t < t_end; t += kTmaGroups, k += kTmaGroups) {
    const int s = k % nstages;
    if (b != cur_b) {
      // new image: (re)stage its seed table; every warp of the group reaches this at the same tile index
      group_bar_sync(grp);
      n = min(n_seeds[b], Nmax);
      if (wl == 0 && lane == 0 && n > 0) {
        mbar_expect_tx(seedbar, (uint32_t)(n * sizeof(SeedRec)));
        bulk_g2s(s_all, seeds + (size_t)b * Nmax, (uint32_t)(n * sizeof(SeedRec)), seedbar);
      }
      if (n > 0) { mbar_wait(seedbar, seed_phase); seed_phase ^= 1u; }
      thr = make_thr(thr_key[b]);
      cur_b = b; cur_ty = -1;
    }
    if (ty != cur_ty) {
      // new band: ordered list of the seeds whose boxes overlap this warp's rows
      ybeg = ty * kTmaTileH + wl * kTmaRW;
      const int sy0 = ybeg, sy1 = min(ybeg + kTmaRW - 1, H - 1);
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
    const int x0 = tx * kTmaTileW + lane * 4;
    LaneCols lc;
    lc.nin = min(max(W - x0, 0), 4); lc.has_left = x0 - 1 >= 0; lc.has_right = x0 + 4 < W;

    // (1) per-warp ordered culling of the band list against this tile's columns
    int Tw = 0;
    {
      const int sx0 = tx * kTmaTileW, sx1 = min(sx0 + kTmaTileW - 1, W - 1);
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

    // (2) wait for the tile
    if (ISG_TMA_CONSUMER_BACKOFF_NS > 0) mbar_wait_backoff(&full[s], (uint32_t)((k / nstages) & 1), ISG_TMA_CONSUMER_BACKOFF_NS);
    else mbar_wait(&full[s], (uint32_t)((k / nstages) & 1));
    const float* kbox = reinterpret_cast<const float*>(stage_base + (size_t)s * kStageBytes);
    const float* abox = reinterpret_cast<const float*>(stage_base + (size_t)s * kStageBytes + kKpStageBytes);

    // (3) keep bits
    uint32_t nib[kTmaRW];
    if (thr.use_int) keep_rows_smem<true>(kbox, wl * kTmaRW, ybeg, H, lc, thr, lane, nib);
    else keep_rows_smem<false>(kbox, wl * kTmaRW, ybeg, H, lc, thr, lane, nib);
#pragma unroll
    for (int r = 0; r < kTmaRW; ++r) {
      const uint32_t word = nibbles_to_word(nib[r], lane);
      if ((lane & 7) == 0 && x0 < W && ybeg + r < H)
        keepbits[((size_t)b * H + ybeg + r) * Wwords + (x0 >> 5)] = word;
    }

    // (4) embedding
    float a[kTmaRW][4][4];
#pragma unroll
    for (int r = 0; r < kTmaRW; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float4 v = *reinterpret_cast<const float4*>(abox + c * kAePlaneFloats + (wl * kTmaRW + r) * kTmaTileW + lane * 4);
        a[r][c][0] = v.x; a[r][c][1] = v.y; a[r][c][2] = v.z; a[r][c][3] = v.w;
      }
    // every read of the stage is done: hand it back to the producer
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);

    float xs4[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) xs4[i] = (x0 + i < W) ? __ldg(xs + x0 + i) : 0.0f;
    {
      float amax = 0.0f;
#pragma unroll
      for (int r = 0; r < kTmaRW; ++r)
#pragma unroll
        for (int i = 0; i < 4; ++i) amax = fmaxf(amax, fmaxf(fabsf(a[r][0][i]), fabsf(a[r][1][i])));
      const bool small = __all_sync(0xffffffffu, amax < 0.55f);
#pragma unroll
      for (int r = 0; r < kTmaRW; ++r) {
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

    // (5) membership against the culled seeds, ascending seed index
    float best[kTmaRW][4];
    int lab[kTmaRW][4];
#pragma unroll
    for (int r = 0; r < kTmaRW; ++r)
#pragma unroll
      for (int i = 0; i < 4; ++i) { best[r][i] = 0.0f; lab[r][i] = 0; }
    for (int q = 0; q < Tw; ++q) {
      const int id = my_hit[q];
      const int4 bx = *reinterpret_cast<const int4*>(&s_all[id]);          // y0,y1,x0,x1 (broadcast)
      if (bx.z > x0 + 3 || bx.w < x0) continue;                            // lane column cull
      const float2 cc = *reinterpret_cast<const float2*>(&s_all[id].cy);
      seed_update<kTmaRW>(bx, cc.x, cc.y, id, ybeg, x0, a, best, lab);
    }

    // (6) stores + statistics of the keep pixels
#pragma unroll
    for (int r = 0; r < kTmaRW; ++r) {
      const int y = ybeg + r;
      if (y >= H) break;
      if (x0 + 3 < W) {   // W % 4 == 0 on this path
        stg_stream4(label_map + ((size_t)b * H + y) * W + x0, lab[r][0], lab[r][1], lab[r][2], lab[r][3]);
        if (SCORE) stg_stream4f(score_map + ((size_t)b * H + y) * W + x0, best[r][0], best[r][1], best[r][2], best[r][3]);
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
    __syncwarp();   // my_hit is rewritten by the next tile's culling
    tx += kTmaGroups;
    while (tx >= tilesX) { tx -= tilesX; if (++ty == tilesY) { ty = 0; ++b; } }
  }
}

// ---- host side -------------------------------------------------------------------------------
typedef CUresult (*isg_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline isg_encode_tiled_fn get_encode_tiled() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess) return nullptr;
  if (q != cudaDriverEntryPointSuccess) return nullptr;
  return reinterpret_cast<isg_encode_tiled_fn>(fn);
}

inline size_t dense_tma_smem_bytes(int Nmax, int nstages) {
  return (size_t)nstages * kStageBytes + (size_t)kTmaGroups * Nmax * sizeof(SeedRec) +
         (((size_t)kTmaGroups * kTmaConsumers * Nmax * 4 + 15) & ~(size_t)15) + (2 * kTmaMaxStages + kTmaGroups) * sizeof(uint64_t);
}

// returns ISG_EUNSUPPORTED when the v2 path cannot be used (caller falls back to v1)
inline int launch_dense_tma(const float* kp, int64_t kp_img_stride, const float* ae, int64_t ae_img_stride,
                            int64_t ae_plane_stride, const uint32_t* thr_key, const uint32_t* seeds, const float* ghost,
                            const int32_t* n_seeds, int B, int Nmax, int H, int W, const float* ys, const float* xs,
                            int32_t* label_map, float* score_map, uint32_t* keepbits, int32_t* stats, cudaStream_t stream) {
  if (Nmax > 65535) return ISG_EUNSUPPORTED;
  int nstages = kTmaMaxStages;
  while (nstages > 1 && dense_tma_smem_bytes(Nmax, nstages) > 220 * 1024) --nstages;
  if (nstages < 2) return ISG_EUNSUPPORTED;
  static isg_encode_tiled_fn encode = get_encode_tiled();
  if (!encode) return ISG_EUNSUPPORTED;
  CUtensorMap tm_kp, tm_ae;
  {
    const cuuint64_t dim[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    const cuuint64_t str[2] = {(cuuint64_t)W * 4, (cuuint64_t)kp_img_stride * 4};
    const cuuint32_t box[3] = {kKpBoxW, kKpBoxH, 1};
    const cuuint32_t es[3] = {1, 1, 1};
    if (encode(&tm_kp, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(kp), dim, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return ISG_EUNSUPPORTED;
  }
  {
    const cuuint64_t dim[4] = {(cuuint64_t)W, (cuuint64_t)H, 4, (cuuint64_t)B};
    const cuuint64_t str[3] = {(cuuint64_t)W * 4, (cuuint64_t)ae_plane_stride * 4, (cuuint64_t)ae_img_stride * 4};
    const cuuint32_t box[4] = {kTmaTileW, kTmaTileH, 4, 1};
    const cuuint32_t es[4] = {1, 1, 1, 1};
    if (encode(&tm_ae, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(ae), dim, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return ISG_EUNSUPPORTED;
  }
  const int tilesX = cdiv(W, kTmaTileW), tilesY = cdiv(H, kTmaTileH);
  const long long T = (long long)B * tilesX * tilesY;
  int dev = 0, sms = kSMs;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = (int)std::min<long long>(T, sms);
  const size_t smem = dense_tma_smem_bytes(Nmax, nstages);
  const int Wwords = cdiv(W, 32);
  if (score_map) {
    ISG_CUDA(cudaFuncSetAttribute(assign_dense_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    assign_dense_tma_kernel<true><<<grid, 32 * (kTmaConsumers * kTmaGroups + 1), smem, stream>>>(
        tm_kp, tm_ae, thr_key, reinterpret_cast<const SeedRec*>(seeds), reinterpret_cast<const float4*>(ghost), n_seeds, Nmax,
        B, H, W, Wwords, tilesX, tilesY, nstages, ys, xs, label_map, score_map, keepbits, stats);
  } else {
    ISG_CUDA(cudaFuncSetAttribute(assign_dense_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    assign_dense_tma_kernel<false><<<grid, 32 * (kTmaConsumers * kTmaGroups + 1), smem, stream>>>(
        tm_kp, tm_ae, thr_key, reinterpret_cast<const SeedRec*>(seeds), reinterpret_cast<const float4*>(ghost), n_seeds, Nmax,
        B, H, W, Wwords, tilesX, tilesY, nstages, ys, xs, label_map, score_map, keepbits, stats);
  }
  ISG_LAUNCH_CHECK();
  return ISG_OK;
}

}  // namespace isg
