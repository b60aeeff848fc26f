// Fused dense kernel (dense_v4_kernel + tile_lists_kernel): embedding + membership + assignment + 3x3 peak test.
//
// What one launch does for every pixel of the batch (reference: select_points / nms_hm, utils/decode.py:42-48,71-85,
// and the core of group_kp, :303-328):
//   keep bit  = selected (kp >= k-th largest) and 3x3 maximum of the thresholded map,
//   embedding = tanh(ae[0:2]) + grid, sigma = exp(ae[2:4]),
//   label     = first seed (in index order) with the largest membership exp(-q) among the seeds whose box
//               contains the pixel, q = (e_y-c_y)^2*s_y + (e_x-c_x)^2*s_x; 0 when no membership is > 0,
//   optionally the per-instance count / bbox of the keep pixels that pass the ghost filter.
// Every input pixel is read from HBM once (kp halo rows/columns are re-read from L2), the label is written once.
//
// Structure
//   * PRE-PASS (tile_lists_kernel, one warp per tile): the ordered list (ascending seed index, so that the first-index
//     tie rule survives) of the seed records whose boxes overlap the tile, as a 32-byte header + the first `cap`
//     records in the workspace; all hit indices also go to an overflow array read only by tiles with more hits.  The
//     lists depend on the seeds only: the engine builds them on the box branch while the top-k threshold is computed.
//   * MAIN KERNEL: persistent, one CTA per SM (minus Tuning::dense_spare).  The last warp's lane 0 is the PRODUCER: per
//     tile it issues three async copies into a shared-memory ring on one mbarrier - the kp box with a 1-pixel halo
//     (out-of-image elements NaN-filled = the reference's -inf padding because fmax ignores NaN), the four ae planes
//     as one 4-D box, and the tile's list (1-D bulk copy).  Tiles are assigned statically (t = cta, cta + grid, ...)
//     except the last `dyn_tail` per CTA, which come from a global counter so that early finishers absorb imbalance;
//   * the other warps are CONSUMERS in G groups of WG warps; group g takes the stages k = g (mod G).  Inside a
//     tile a warp owns RW consecutive rows and walks down them with a rolling window of the separable 3x3 maximum; a
//     lane owns 4 consecutive pixels of each row;
//   * consumers are stateless with respect to images: everything they need is in the stage (header + list);
//   * EARLY RELEASE: a warp hands its stage slot back as soon as the kp window and the ae values of its rows are in
//     registers, before the membership loop; the tile lists live in their own ring, G slots deeper than the stage ring
//     (D4Smem), so they stay valid until the group has finished the tile.  Measured: 3 / 4 / 5 stages 76.7 / 71.6 / 70.4 us
//     with release at the end of the tile, 71.7 / 70.2 / 70.4 us with early release - at 5 stages the ring depth is not
//     the bound, the early release makes the kernel insensitive to it (ISG_DENSE_DEBUG bit 2 restores the late release);
//   * the membership loop tracks the SMALLEST exponent q instead of the largest exp(-q): exp is monotone, so the
//     winner is the same seed whenever two memberships differ as fp32 numbers, and the transcendental per
//     (pixel, seed) pair disappears.  q >= ln(2^150) is where exp(-q) rounds to 0 in fp32 (label stays 0).
#pragma once
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include "keep.cuh"

// sigma = exp(ae[2:4]) of the dense kernel: the compensated form (<= 2 ulp) unless ISG_D4_BARE_SIGMA is defined
#ifdef ISG_D4_BARE_SIGMA
#define D4_SIGMA_EXP exp_bare_ftz
#else
#define D4_SIGMA_EXP exp_fast_ftz
#endif

namespace isg {

constexpr int kD4TileW = 128;
constexpr int kD4KpW = kD4TileW + 8;          // 4 columns of padding on each side keep the box 16-byte granular
constexpr int kD4MaxStages = 8;
constexpr float kQZero = 103.97207708f;       // ln(2^150): exp(-q) == 0 in fp32 (round to nearest, subnormals kept)

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (libcuda is not linked)
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

// KP = false: the labels-only form (isg_assign_labels) - no kp box in the stage, the keep bits come from
// isg_keep_from_candidates
template <int RW, int WG, bool KP = true>
struct D4Geom {
  static constexpr int TH = RW * WG;                                  // tile rows
  static constexpr int kKpRows = TH + 2;
  static constexpr int kKpBytes = KP ? kD4KpW * kKpRows * 4 : 0;
  static constexpr int kKpStage = (kKpBytes + 127) / 128 * 128;
  static constexpr int kAePlane = kD4TileW * TH;                      // floats
  static constexpr int kAeBytes = kAePlane * 16;
  static constexpr int kStage = kKpStage + kAeBytes;
  static constexpr uint32_t kTx = kKpBytes + kAeBytes;
};

// Per-tile work list, built by tile_lists_kernel and copied next to the tile by the producer (one 1-D bulk copy):
// a 32-byte header followed by the first `cap` seed records overlapping the tile, in ascending seed index.
struct __align__(16) TileHdr {
  int b, y0, x0, nhit;            // image, first row / column of the tile, seeds overlapping the tile (< 0: end marker)
  uint32_t thr_key;               // unused (the threshold comes from thr_key[b]: the lists do not depend on the top-k)
  int n_seeds;                    // seeds of image b
  int tile;                       // tile index (locates the overflow list when nhit > cap)
  int pad;
};
static_assert(sizeof(TileHdr) == 32, "tile header layout");

struct RowK {
  float h[4];      // horizontal 3-max of v centred on the lane's 4 pixels
  float v[4];      // thresholded value (0 when not selected, NaN outside the image)
  uint32_t sel;    // bit i: pixel i is selected
};

// One staged kp row -> RowK.  `row` points at the first float of the box row; the lane's pixels start at
// row[4 + 4*lane]; row[3] / row[132] are the halo pixels of lane 0 / lane 31 (`halo_off` selects one per lane).
template <bool USE_INT>
__device__ __forceinline__ RowK d4_prep_row(const float* __restrict__ row, int lane, int halo_off, const Thr& thr) {
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
__device__ __forceinline__ RowK d4_prep(const float* __restrict__ row, int lane, int halo_off, const Thr& thr) {
  return thr.use_int ? d4_prep_row<true>(row, lane, halo_off, thr) : d4_prep_row<false>(row, lane, halo_off, thr);
}
// keep nibble of the middle row's 4 pixels
__device__ __forceinline__ uint32_t d4_keep(const RowK& up, const RowK& mid, const RowK& dn) {
  uint32_t nib = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float m = fmaxf(fmaxf(up.h[i], mid.h[i]), dn.h[i]);
    if (mid.v[i] >= m) nib |= 1u << i;
  }
  return nib & mid.sel;
}

__device__ __forceinline__ void mbar_arrive_plain(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
// Waits used by this kernel.  With -DISG_D4_WATCHDOG a wait that does not complete within ~2 s prints where it is
// stuck and traps (debug builds only; the production build spins without a bound).
__device__ __forceinline__ bool d4_try_wait(uint64_t* bar, uint32_t phase) {
  uint32_t done;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}" : "=r"(done) : "r"(smem_u32(bar)), "r"(phase) : "memory");
  return done != 0;
}
// producer-side wait: the hardware parks the thread until the phase completes or the hint (ns) expires, so a
// waiting producer costs its scheduler almost no issue slots
__device__ __forceinline__ bool d4_try_wait_hint(uint64_t* bar, uint32_t phase, uint32_t hint_ns) {
  uint32_t done;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}" : "=r"(done) : "r"(smem_u32(bar)), "r"(phase), "r"(hint_ns) : "memory");
  return done != 0;
}
__device__ __forceinline__ void d4_wait_parked(uint64_t* bar, uint32_t phase) {
  while (!d4_try_wait_hint(bar, phase, 20000u)) {}
}
__device__ __forceinline__ void d4_wait(uint64_t* bar, uint32_t phase, unsigned sleep_ns, int site, int a0, int a1) {
#ifdef ISG_D4_WATCHDOG
  const long long t0 = clock64();
  bool said = false;
  while (!d4_try_wait(bar, phase)) {
    if (sleep_ns) __nanosleep(sleep_ns);
    const long long dt = clock64() - t0;
    if (dt > 4000000000ll && !said) {
      said = true;
      if ((threadIdx.x & 31) == 0 || site == 1)
        printf("[d4 watchdog] site=%d block=%d warp=%d lane=%d phase=%u a0=%d a1=%d\n", site, blockIdx.x, threadIdx.x >> 5,
               threadIdx.x & 31, phase, a0, a1);
    }
    if (dt > 8000000000ll) __trap();
  }
#else
  (void)site; (void)a0; (void)a1;
  while (!d4_try_wait(bar, phase)) {
    if (sleep_ns) __nanosleep(sleep_ns);
  }
#endif
}
__device__ __forceinline__ void d4_tma_3d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      :: "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void d4_tma_4d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
      :: "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
      : "memory");
}

// shared-memory layout (byte offsets from the dynamic shared base)
template <int RW, int WG, bool KP = true>
struct D4Smem {
  int stage, list, list_stride, bars, thr, total;
  // The list ring is G slots deeper than the stage ring: consumers hand a stage back as soon as its pixels are in
  // registers (early release) but keep reading the tile's list; the list slot of tile s is reused by tile s + nstages + G,
  // which the producer issues only after the same group released tile s + G - i.e. after it finished tile s.
  __host__ __device__ D4Smem(int cap, int nstages, int G) {
    using Geo = D4Geom<RW, WG, KP>;
    int o = 0;
    stage = o; o += nstages * Geo::kStage;
    list_stride = (int)sizeof(TileHdr) + cap * (int)sizeof(SeedRec);
    list = o;  o += (nstages + G) * list_stride;
    bars = o;  o += 2 * kD4MaxStages * 8;
    thr = o;   o += kD4MaxStages * 4;           // selection threshold key of the tile in each slot (written by the producer)
    total = o;
  }
};

constexpr size_t kDenseSchedBytes = 256;     // scheduler words at the head of the dense workspace

// Pre-pass: one warp per tile collects, in ascending seed index, the seeds of the tile's image whose boxes overlap
// the tile.  lists: [T] x (TileHdr + cap records); ovf: [T][Nmax] seed indices of ALL hits (read by the dense kernel
// only for the rare tiles with more than cap hits).
__global__ void __launch_bounds__(256)
tile_lists_kernel(const SeedRec* __restrict__ seeds, const int32_t* __restrict__ n_seeds,
                  int Nmax, int B, int H, int W, int TH, int tilesX, int tilesY,
                  int cap, unsigned char* __restrict__ lists, uint16_t* __restrict__ ovf, unsigned int* __restrict__ sched) {
  // programmatic dependent launch: let the dense kernel's CTAs start (barrier init, first kp/ae loads) while this
  // grid is still running; its producer waits on the grid dependency before it touches a list
  asm volatile("griddepcontrol.launch_dependents;");
  pdl_wait();         // the seed records come from the kernel launched just before (no-op for an ordinary launch)
  // the tile counter of the dynamic scheduler starts every dense launch at zero: reset here, together with the lists
  // (the dense kernel behind this grid reads it only after its own grid-dependency wait)
  if (blockIdx.x == 0 && threadIdx.x < kDenseSchedBytes / 4) sched[threadIdx.x] = 0u;
  const int lane = threadIdx.x & 31;
  const long long T = (long long)B * tilesX * tilesY;
  const long long t = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (t >= T) return;
  const int tiles_per_img = tilesX * tilesY;
  const int b = (int)(t / tiles_per_img);
  const int rem = (int)(t - (long long)b * tiles_per_img);
  const int ty = rem / tilesX, tx = rem - ty * tilesX;
  const int y0 = ty * TH, x0 = tx * kD4TileW;
  const int y1 = min(y0 + TH - 1, H - 1), x1 = min(x0 + kD4TileW - 1, W - 1);
  const int n = min(n_seeds[b], Nmax);
  const size_t stride = sizeof(TileHdr) + (size_t)cap * sizeof(SeedRec);
  unsigned char* out = lists + (size_t)t * stride;
  int4* recs = reinterpret_cast<int4*>(out + sizeof(TileHdr));
  uint16_t* ov = ovf + (size_t)t * Nmax;
  const int4* src = reinterpret_cast<const int4*>(seeds + (size_t)b * Nmax);
  int cnt = 0;
  for (int j0 = 0; j0 < n; j0 += 128) {            // 4 independent loads per lane in flight (the loop is latency-bound)
    int4 lo[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = j0 + u * 32 + lane;
      lo[u] = (j < n) ? __ldg(src + 2 * j) : make_int4(1, 0, 1, 0);     // empty box
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = j0 + u * 32 + lane;
      const bool hit = lo[u].x <= y1 && lo[u].y >= y0 && lo[u].z <= x1 && lo[u].w >= x0 && lo[u].x <= lo[u].y && lo[u].z <= lo[u].w;
      const unsigned bal = __ballot_sync(0xffffffffu, hit);
      if (hit) {
        const int pos = cnt + __popc(bal & ((1u << lane) - 1u));
        ov[pos] = (uint16_t)j;
        if (pos < cap) { recs[2 * pos] = lo[u]; recs[2 * pos + 1] = __ldg(src + 2 * j + 1); }
      }
      cnt += __popc(bal);
    }
  }
  if (lane == 0) {
    TileHdr h;
    h.b = b; h.y0 = y0; h.x0 = x0; h.nhit = cnt; h.thr_key = 0u; h.n_seeds = n; h.tile = (int)t; h.pad = 0;
    *reinterpret_cast<TileHdr*>(out) = h;
  }
}

template <int RW, int WG, int G, bool SCORE, bool KP>
__global__ void __launch_bounds__(32 * (WG * G + 1), 1)
dense_v4_kernel(const __grid_constant__ CUtensorMap tm_kp, const __grid_constant__ CUtensorMap tm_ae,
                const uint32_t* __restrict__ thr_key, const SeedRec* __restrict__ seeds, const float4* __restrict__ ghost,
                int Nmax, int B, int H, int W,
                int Wwords, int tilesX, int tilesY, int nstages, int cap, const unsigned char* __restrict__ lists,
                const uint16_t* __restrict__ ovf, const float* __restrict__ ys, const float* __restrict__ xs,
                int32_t* __restrict__ label_map, float* __restrict__ score_map, uint32_t* __restrict__ keepbits,
                int32_t* __restrict__ stats, unsigned int* __restrict__ sched, int dyn_tail, int dbg_flags, int skip_ae) {
  static_assert(RW % 2 == 0, "rows are processed in pairs");
  using Geo = D4Geom<RW, WG, KP>;
  constexpr int kConsumers = WG * G;
  const bool early_release = !(dbg_flags & 4);        // measurement aid: bit 2 = release stages only at the end of a tile
  extern __shared__ __align__(1024) unsigned char smem[];
  const D4Smem<RW, WG, KP> L(cap, nstages, G);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L.bars);           // [kD4MaxStages]
  uint64_t* empty = full + kD4MaxStages;                                 // [kD4MaxStages]
  uint32_t* s_thr = reinterpret_cast<uint32_t*>(smem + L.thr);           // [kD4MaxStages]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // a kernel launched behind this one with programmatic stream serialisation (isg_instance_polygons) may be scheduled as
  // soon as SMs free up at the tail; it waits for this grid's completion (griddepcontrol.wait) before reading anything
  asm volatile("griddepcontrol.launch_dependents;");
  if (threadIdx.x == 0) {
    for (int s = 0; s < nstages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], WG); }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp == kConsumers) {
    // ============================ producer: one thread feeds the ring ============================
    if (lane != 0) return;
    const int tiles_per_img = tilesX * tilesY;
    const unsigned T = (unsigned)B * (unsigned)tiles_per_img;
    const uint32_t list_bytes = (uint32_t)L.list_stride;
    const int lslots = nstages + G;
    int slot = 0, round = 0, ls = 0;                       // ls: slot of the list ring (tile sequence number mod lslots)
    // Tile order: the first n_static tiles of a CTA are c, c + P, c + 2P, ... (P = grid size; neighbouring CTAs work
    // on neighbouring tiles, no scheduling traffic); the last `dyn_tail` tiles per CTA (and the remainder of T / P)
    // come from a global counter (zeroed by the host before the launch) so that CTAs that finish early take over work
    // from the slow ones.  The counter value is always requested one tile ahead: an atomic round trip under full memory load is ~2 us, a CTA that
    // waited for it on every tile would be latency-bound.
    const unsigned P = gridDim.x;
    const unsigned n_static = (T / P > (unsigned)dyn_tail) ? (T / P - (unsigned)dyn_tail) : 0u;
    const unsigned T_s = n_static * P;
    unsigned i_static = 0;
    // the tile lists and the zeroed counter come from the preceding grid (tile_lists_kernel): wait for it (no-op
    // without PDL) before the first request
    asm volatile("griddepcontrol.wait;" ::: "memory");
    unsigned dyn = T_s + atomicAdd(&sched[0], 1u);
    unsigned t;
    if (n_static > 0) { t = blockIdx.x; i_static = 1; }
    else { t = dyn; dyn = (t < T) ? T_s + atomicAdd(&sched[0], 1u) : T; }
    // Seeds overlapping the tile (header word 3 of its list), fetched one tile ahead so that the load is in flight
    // while the producer waits for a free slot.  A tile that no seed box overlaps needs no embedding at all - every
    // pixel outside all boxes gets label 0 (utils/decode.py:325-328) - so its four ae planes (3/4 of a tile's bytes)
    // are not loaded.
    auto nhit_of = [&](unsigned tt) -> int {
      return tt < T ? __ldg(reinterpret_cast<const int*>(lists + (size_t)tt * list_bytes) + 3) : 0;
    };
    int nh = nhit_of(t);
    while (t < T) {
      const int b = (int)(t / (unsigned)tiles_per_img);
      const int rem = (int)t - b * tiles_per_img;
      const int ty = rem / tilesX, tx = rem - ty * tilesX;
      const int y0t = ty * Geo::TH, x0t = tx * kD4TileW;
      // the slot must have been released by the consumers of its previous tile
      if (round >= 1) d4_wait_parked(&empty[slot], (uint32_t)((round - 1) & 1));
      unsigned char* st = smem + L.stage + (size_t)slot * Geo::kStage;
      if (KP) s_thr[slot] = __ldg(thr_key + b);  // ordinary store: released to the consumers by the arrive below
      const bool with_ae = nh > 0 || skip_ae == 0;
      mbar_expect_tx(&full[slot], (with_ae ? Geo::kTx : (uint32_t)Geo::kKpBytes) + list_bytes);
      bulk_g2s(smem + L.list + (size_t)ls * L.list_stride, lists + (size_t)t * list_bytes, list_bytes, &full[slot]);
      if (++ls == lslots) ls = 0;
      if (KP) d4_tma_3d(st, &tm_kp, x0t - 4, y0t - 1, b, &full[slot]);
      if (with_ae) d4_tma_4d(st + Geo::kKpStage, &tm_ae, x0t, y0t, 0, b, &full[slot]);
      if (++slot == nstages) { slot = 0; ++round; }
      if (i_static < n_static) {
        t = blockIdx.x + (i_static++) * P;
      } else {
        t = dyn;
        dyn = (t < T) ? T_s + atomicAdd(&sched[0], 1u) : T;
      }
      nh = nhit_of(t);
    }
    // one end marker per consumer group (the next G stages cover every group once)
    for (int g = 0; g < G; ++g) {
      if (round >= 1) d4_wait_parked(&empty[slot], (uint32_t)((round - 1) & 1));
      reinterpret_cast<TileHdr*>(smem + L.list + (size_t)ls * L.list_stride)->nhit = -1;
      if (++ls == lslots) ls = 0;
      mbar_arrive_plain(&full[slot]);
      if (++slot == nstages) { slot = 0; ++round; }
    }
    // the last CTA to finish leaves the scheduler words zeroed, so a second lists_prebuilt launch on the same lists
    // also starts from zero (the host additionally zeroes them whenever it rebuilds the lists)
    __threadfence();
    const unsigned done = atomicAdd(&sched[1], 1u);
    if (done == gridDim.x - 1) { sched[0] = 0u; sched[1] = 0u; __threadfence(); }
    return;
  }

  // ========================================= consumers =========================================
  const int grp = warp / WG, wl = warp % WG;
  const int halo_off = (lane == 31) ? (4 + kD4TileW) : 3;
  const int lslots = nstages + G;
  int slot = grp, rnd = 0, ls = grp;
  while (true) {
    bool released = false;                                                  // warp-uniform: stage already handed back
    // Slots are shared between groups over time and TMA completions are not ordered, so the previous fill of this
    // slot (another group's tile) may not have landed when this group gets here; a parity wait only tells "this
    // phase" from "the one before".  The previous occupant's RELEASE (empty phase rnd-1) implies its fill completed,
    // and it cannot be more than one phase away (the release of fill rnd is ours): wait for it first.
    if (rnd > 0) d4_wait(&empty[slot], (uint32_t)((rnd - 1) & 1), 0, 6, slot, rnd);
    d4_wait(&full[slot], (uint32_t)(rnd & 1), 0, 5, slot, rnd);
    const unsigned char* lst = smem + L.list + (size_t)ls * L.list_stride;
    const int4 h0 = *reinterpret_cast<const int4*>(lst);                    // b, y0, x0, nhit
    const int nhit = h0.w;
    if (nhit < 0) break;
    const int b = h0.x;
    const int ybeg = h0.y + wl * RW;
    if (ybeg < H && !(dbg_flags & 1)) {                                     // warp-uniform (ragged bottom)
      const int4 h1 = *reinterpret_cast<const int4*>(lst + 16);             // -, n_seeds, tile, pad
      const Thr thr = make_thr(KP ? s_thr[slot] : 0xffffffffu);
      const int x0 = h0.z + lane * 4;
      const bool colvalid = x0 < W;                                         // W % 4 == 0: a lane is all in or all out
      float xs4[4] = {0.f, 0.f, 0.f, 0.f};
      if (colvalid) { const float4 v = __ldg(reinterpret_cast<const float4*>(xs + x0)); xs4[0] = v.x; xs4[1] = v.y; xs4[2] = v.z; xs4[3] = v.w; }
      const unsigned char* st = smem + L.stage + (size_t)slot * Geo::kStage;
      const float* krow = reinterpret_cast<const float*>(st) + (wl * RW) * kD4KpW;     // box row = tile row + 1
      const float* arow = reinterpret_cast<const float*>(st + Geo::kKpStage) + (wl * RW) * kD4TileW + lane * 4;
      const unsigned char* recs = lst + sizeof(TileHdr);
      const int nsm = min(nhit, cap);
      const size_t pix0 = ((size_t)b * H + ybeg) * W + x0;
      uint32_t* kbrow = keepbits + ((size_t)b * H + ybeg) * Wwords + (x0 >> 5);
      const bool kbwriter = ((lane & 7) == 0) && colvalid;

#pragma unroll
      for (int rp = 0; rp < RW; rp += 2) {
        const int y = ybeg + rp;
        if (y >= H) break;                                                  // warp-uniform
        const bool row1 = y + 1 < H;
        // --- keep bits of rows y, y+1 ---
        // A row segment without a selected pixel cannot hold a keep pixel, and ~1 % of the pixels are selected: test the
        // maximum of the lane's four raw values first (3 FMNMX + 1 compare per row) and run the thresholded 3x3 window
        // only for the rows of this warp that do hold a selected pixel.  (+-0 / NaN thresholds order by key: always run.)
        uint32_t nib0 = 0, nib1 = 0;
        if (KP) {
        bool a0 = true, a1 = true;
        if (!thr.use_int) {
          const float4 q0 = *reinterpret_cast<const float4*>(krow + (rp + 1) * kD4KpW + 4 + lane * 4);
          const float4 q1 = *reinterpret_cast<const float4*>(krow + (rp + 2) * kD4KpW + 4 + lane * 4);
          a0 = __any_sync(0xffffffffu, fmaxf(fmaxf(q0.x, q0.y), fmaxf(q0.z, q0.w)) >= thr.f);
          a1 = __any_sync(0xffffffffu, fmaxf(fmaxf(q1.x, q1.y), fmaxf(q1.z, q1.w)) >= thr.f);
        }
        if (a0 || a1) {                                                      // warp-uniform
          const RowK r1 = d4_prep(krow + (rp + 1) * kD4KpW, lane, halo_off, thr);
          const RowK r2 = d4_prep(krow + (rp + 2) * kD4KpW, lane, halo_off, thr);
          if (a0) nib0 = d4_keep(d4_prep(krow + rp * kD4KpW, lane, halo_off, thr), r1, r2);
          if (a1) nib1 = d4_keep(r1, r2, d4_prep(krow + (rp + 3) * kD4KpW, lane, halo_off, thr));
        }
        if (!colvalid) { nib0 = 0; nib1 = 0; }
        if (!row1) nib1 = 0;
        {
          const uint32_t w0 = nibbles_to_word(nib0, lane), w1 = nibbles_to_word(nib1, lane);
          if (kbwriter) {
            kbrow[(size_t)rp * Wwords] = w0;
            if (row1) kbrow[(size_t)(rp + 1) * Wwords] = w1;
          }
        }
        }   // KP

        float bq[2][4];
        int lab[2][4];
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
          for (int i = 0; i < 4; ++i) { bq[r][i] = kQZero; lab[r][i] = 0; }
        if (nhit > 0 || skip_ae == 0) {                                      // tile-uniform
#ifndef ISG_D4_SCALAR
        // --- embedding of the lane's 2 x 4 pixels, two pixels per instruction (FADD2 / FMUL2 / FFMA2, common.cuh) ---
        f32x2 ey[2][2], ex[2][2], sy[2][2], sx[2][2];
        {
          float4 a0[2], a1[2];
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            const float* ar = arow + (rp + r) * kD4TileW;
            a0[r] = *reinterpret_cast<const float4*>(ar);
            a1[r] = *reinterpret_cast<const float4*>(ar + Geo::kAePlane);
            const float4 a2 = *reinterpret_cast<const float4*>(ar + 2 * Geo::kAePlane);
            const float4 a3 = *reinterpret_cast<const float4*>(ar + 3 * Geo::kAePlane);
            sy[r][0] = exp_fast_ftz2(pack2(a2.x, a2.y)); sy[r][1] = exp_fast_ftz2(pack2(a2.z, a2.w));
            sx[r][0] = exp_fast_ftz2(pack2(a3.x, a3.y)); sx[r][1] = exp_fast_ftz2(pack2(a3.z, a3.w));
          }
          // Early release: with the last row pair in registers nothing of the stage is read again (the list lives in
          // its own, deeper ring), so the producer can refill the slot while this warp computes.
          if (rp + 2 >= RW && early_release) {
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[slot]);
            released = true;
          }
          float amax = 0.0f;
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            amax = fmaxf(amax, fmaxf(fmaxf(fabsf(a0[r].x), fabsf(a0[r].y)), fmaxf(fabsf(a0[r].z), fabsf(a0[r].w))));
            amax = fmaxf(amax, fmaxf(fmaxf(fabsf(a1[r].x), fabsf(a1[r].y)), fmaxf(fabsf(a1[r].z), fabsf(a1[r].w))));
          }
          const bool small = __all_sync(0xffffffffu, amax < 0.55f);       // warp-uniform: polynomial branch only
          const float yv[2] = {__ldg(ys + y), __ldg(ys + min(y + 1, H - 1))};
          const f32x2 xs01 = pack2(xs4[0], xs4[1]), xs23 = pack2(xs4[2], xs4[3]);
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            const f32x2 yy = pack2(yv[r], yv[r]);
            if (small) {
              ey[r][0] = add2(tanh_poly2(pack2(a0[r].x, a0[r].y)), yy); ey[r][1] = add2(tanh_poly2(pack2(a0[r].z, a0[r].w)), yy);
              ex[r][0] = add2(tanh_poly2(pack2(a1[r].x, a1[r].y)), xs01); ex[r][1] = add2(tanh_poly2(pack2(a1[r].z, a1[r].w)), xs23);
            } else {
              ey[r][0] = add2(pack2(tanh_fast(a0[r].x), tanh_fast(a0[r].y)), yy); ey[r][1] = add2(pack2(tanh_fast(a0[r].z), tanh_fast(a0[r].w)), yy);
              ex[r][0] = add2(pack2(tanh_fast(a1[r].x), tanh_fast(a1[r].y)), xs01); ex[r][1] = add2(pack2(tanh_fast(a1[r].z), tanh_fast(a1[r].w)), xs23);
            }
          }
        }

        // --- membership: smallest exponent among the seeds whose box contains the pixel, ascending seed index ---
        auto visit = [&](const int4 bx, const int4 cw) {                      // y0,y1,x0,x1 | cy,cx,id,pad (warp-uniform)
          const bool in0 = (y >= bx.x) && (y <= bx.y), in1 = (y + 1 >= bx.x) && (y + 1 <= bx.y);
          if (!(in0 || in1)) return;                                           // row test
          const float cy = __int_as_float(cw.x), cx = __int_as_float(cw.y);
          const f32x2 cy2 = pack2(cy, cy), cx2 = pack2(cx, cx);
          const int id = cw.z;
          const int lo = bx.z - x0, hi = bx.w - x0;                            // pixel i is inside iff lo <= i <= hi
          bool pin[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) pin[i] = (lo <= i) && (hi >= i);
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            if (r == 0 ? in0 : in1) {
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                // (e - c)^2 * sigma per axis with separately rounded operations (:326-327); the two products are added
                // with scalar FADDs (ptxas would contract a packed multiply + add into FFMA2)
                const f32x2 dy = sub2(ey[r][h], cy2), dx = sub2(ex[r][h], cx2);
                const f32x2 ty = mul2(mul2(dy, dy), sy[r][h]), tx = mul2(mul2(dx, dx), sx[r][h]);
                float ty0, ty1, tx0, tx1;
                unpack2(ty, ty0, ty1); unpack2(tx, tx0, tx1);
                const float q0 = __fadd_rn(ty0, tx0), q1 = __fadd_rn(ty1, tx1);
                if (pin[2 * h] && q0 < bq[r][2 * h]) { bq[r][2 * h] = q0; lab[r][2 * h] = id; }   // strict: first index wins ties (:328)
                if (pin[2 * h + 1] && q1 < bq[r][2 * h + 1]) { bq[r][2 * h + 1] = q1; lab[r][2 * h + 1] = id; }
              }
            }
          }
        };
#else
        // --- embedding of the lane's 2 x 4 pixels ---
        float ey[2][4], ex[2][4], sy[2][4], sx[2][4];
        {
          float t0[2][4], t1[2][4];
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            const float* ar = arow + (rp + r) * kD4TileW;
            const float4 a0 = *reinterpret_cast<const float4*>(ar);
            const float4 a1 = *reinterpret_cast<const float4*>(ar + Geo::kAePlane);
            const float4 a2 = *reinterpret_cast<const float4*>(ar + 2 * Geo::kAePlane);
            const float4 a3 = *reinterpret_cast<const float4*>(ar + 3 * Geo::kAePlane);
            t0[r][0] = a0.x; t0[r][1] = a0.y; t0[r][2] = a0.z; t0[r][3] = a0.w;
            t1[r][0] = a1.x; t1[r][1] = a1.y; t1[r][2] = a1.z; t1[r][3] = a1.w;
            sy[r][0] = D4_SIGMA_EXP(a2.x); sy[r][1] = D4_SIGMA_EXP(a2.y); sy[r][2] = D4_SIGMA_EXP(a2.z); sy[r][3] = D4_SIGMA_EXP(a2.w);
            sx[r][0] = D4_SIGMA_EXP(a3.x); sx[r][1] = D4_SIGMA_EXP(a3.y); sx[r][2] = D4_SIGMA_EXP(a3.z); sx[r][3] = D4_SIGMA_EXP(a3.w);
          }
          float amax = 0.0f;
#pragma unroll
          for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int i = 0; i < 4; ++i) amax = fmaxf(amax, fmaxf(fabsf(t0[r][i]), fabsf(t1[r][i])));
          const bool small = __all_sync(0xffffffffu, amax < 0.55f);       // warp-uniform: polynomial branch only
          const float yv0 = __ldg(ys + y), yv1 = __ldg(ys + min(y + 1, H - 1));
          if (small) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              ey[0][i] = __fadd_rn(tanh_poly(t0[0][i]), yv0); ex[0][i] = __fadd_rn(tanh_poly(t1[0][i]), xs4[i]);
              ey[1][i] = __fadd_rn(tanh_poly(t0[1][i]), yv1); ex[1][i] = __fadd_rn(tanh_poly(t1[1][i]), xs4[i]);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              ey[0][i] = __fadd_rn(tanh_fast(t0[0][i]), yv0); ex[0][i] = __fadd_rn(tanh_fast(t1[0][i]), xs4[i]);
              ey[1][i] = __fadd_rn(tanh_fast(t0[1][i]), yv1); ex[1][i] = __fadd_rn(tanh_fast(t1[1][i]), xs4[i]);
            }
          }
        }

        // --- membership: smallest exponent among the seeds whose box contains the pixel, ascending seed index ---
        auto visit = [&](const int4 bx, const int4 cw) {                      // y0,y1,x0,x1 | cy,cx,id,pad (warp-uniform)
          const bool in0 = (y >= bx.x) && (y <= bx.y), in1 = (y + 1 >= bx.x) && (y + 1 <= bx.y);
          if (!(in0 || in1)) return;                                           // row test
          const float cy = __int_as_float(cw.x), cx = __int_as_float(cw.y);
          const int id = cw.z;
          const int lo = bx.z - x0, hi = bx.w - x0;                            // pixel i is inside iff lo <= i <= hi
          bool pin[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) pin[i] = (lo <= i) && (hi >= i);
          if (in0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float dy = __fsub_rn(ey[0][i], cy), dx = __fsub_rn(ex[0][i], cx);
              const float qq = __fadd_rn(__fmul_rn(__fmul_rn(dy, dy), sy[0][i]), __fmul_rn(__fmul_rn(dx, dx), sx[0][i]));
              if (pin[i] && qq < bq[0][i]) { bq[0][i] = qq; lab[0][i] = id; }   // strict: first index wins ties (:328)
            }
          }
          if (in1) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float dy = __fsub_rn(ey[1][i], cy), dx = __fsub_rn(ex[1][i], cx);
              const float qq = __fadd_rn(__fmul_rn(__fmul_rn(dy, dy), sy[1][i]), __fmul_rn(__fmul_rn(dx, dx), sx[1][i]));
              if (pin[i] && qq < bq[1][i]) { bq[1][i] = qq; lab[1][i] = id; }
            }
          }
        };
#endif
        for (int q = 0; q < nsm; ++q)
          visit(*reinterpret_cast<const int4*>(recs + q * (int)sizeof(SeedRec)),
                *reinterpret_cast<const int4*>(recs + q * (int)sizeof(SeedRec) + 16));
        for (int q = nsm; q < nhit; ++q) {                                     // rare: more hits than staged records
          const int id = __ldg(ovf + (size_t)h1.z * Nmax + q);
          const int4* g = reinterpret_cast<const int4*>(seeds + (size_t)b * Nmax + id);
          visit(__ldg(g), __ldg(g + 1));
        }
        }   // nhit > 0

        // --- stores + statistics of the keep pixels ---
        if (colvalid) {
          int32_t* lp = label_map + pix0 + (size_t)rp * W;
          stg_stream4(lp, lab[0][0], lab[0][1], lab[0][2], lab[0][3]);
          if (row1) stg_stream4(lp + W, lab[1][0], lab[1][1], lab[1][2], lab[1][3]);
          if (SCORE) {
            float p[2][4];
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
              for (int i = 0; i < 4; ++i) p[r][i] = (bq[r][i] < kQZero) ? exp_fast(-bq[r][i]) : 0.0f;
            float* sp = score_map + pix0 + (size_t)rp * W;
            stg_stream4f(sp, p[0][0], p[0][1], p[0][2], p[0][3]);
            if (row1) stg_stream4f(sp + W, p[1][0], p[1][1], p[1][2], p[1][3]);
          }
        }
        uint32_t nib = nib0 | (nib1 << 4);
        if (nib && stats && h1.y > 0) {
          while (nib) {
            const int bit = __ffs(nib) - 1;
            nib &= nib - 1;
            const int r = bit >> 2, i = bit & 3;
            const int l = (r == 0) ? ((i == 0) ? lab[0][0] : (i == 1) ? lab[0][1] : (i == 2) ? lab[0][2] : lab[0][3])
                                   : ((i == 0) ? lab[1][0] : (i == 1) ? lab[1][1] : (i == 2) ? lab[1][2] : lab[1][3]);
            if (ghost_pass(ghost[(size_t)b * Nmax + l], y + r, x0 + i))
              stats_add(stats + ((size_t)b * Nmax + l) * ISG_STAT_WORDS, y + r, x0 + i);
          }
        }
      }
    }
    // every read of the stage and of its list is done: hand the slot back to the producer (unless already done)
    if (!released) {
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[slot]);
    }
    slot += G;
    if (slot >= nstages) { slot -= nstages; ++rnd; }
    ls += G;
    if (ls >= lslots) ls -= lslots;
  }
}

// ---- host side -------------------------------------------------------------------------------
// Workspace: [0,256) scheduler words | tile lists [T_max] x (32 + cap*32) | overflow indices [T_max][Nmax] u16.
// T_max is the tile count of the smallest compiled tile (8 rows), so that every geometry fits.
constexpr int kD4MinTileRows = 8;
inline int dense_list_cap(int Nmax) { return Nmax <= 16 ? 16 : (Nmax <= 160 ? 16 : 32); }
inline size_t dense_lists_bytes(long long T, int Nmax) {
  const size_t stride = sizeof(TileHdr) + (size_t)dense_list_cap(Nmax) * sizeof(SeedRec);
  return ((size_t)T * stride + 255) & ~(size_t)255;
}
inline size_t dense_workspace_bytes(int B, int Nmax, int H, int W) {
  const long long T = (long long)B * cdiv(W, kD4TileW) * cdiv(H, kD4MinTileRows);
  return kDenseSchedBytes + dense_lists_bytes(T, Nmax) + (((size_t)T * Nmax * 2 + 255) & ~(size_t)255);
}

template <int RW, int WG, int G, bool KP>
inline int launch_dense_v4_cfg(const float* kp, int64_t kp_img_stride, const float* ae, int64_t ae_img_stride,
                               int64_t ae_plane_stride, const uint32_t* thr_key, const uint32_t* seeds, const float* ghost,
                               const int32_t* n_seeds, int B, int Nmax, int H, int W, const float* ys, const float* xs,
                               int32_t* label_map, float* score_map, uint32_t* keepbits, int32_t* stats, void* workspace,
                               size_t workspace_bytes, int max_stages, int mode, cudaStream_t stream) {
  // mode 0: tile lists + dense kernel; 1: tile lists only (isg_build_tile_lists); 2: dense kernel only (lists prebuilt)
  using Geo = D4Geom<RW, WG, KP>;
  static_assert(Geo::TH >= kD4MinTileRows, "workspace is sized for tiles of at least kD4MinTileRows rows");
  if (Nmax > 65535) return ISG_EUNSUPPORTED;
  const int cap = dense_list_cap(Nmax);
  int nstages = std::min(max_stages, kD4MaxStages);
  while (nstages > 1 && D4Smem<RW, WG, KP>(cap, nstages, G).total > 227 * 1024) --nstages;
  if (nstages < G + 1) return ISG_EUNSUPPORTED;
  const int tilesX = cdiv(W, kD4TileW), tilesY = cdiv(H, Geo::TH);
  const long long T = (long long)B * tilesX * tilesY;
  if (T >= (1ll << 31) - 65536) return ISG_EUNSUPPORTED;
  if (workspace_bytes < dense_workspace_bytes(B, Nmax, H, W)) return ISG_EINVAL;
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  unsigned int* sched = reinterpret_cast<unsigned int*>(ws);
  unsigned char* lists = ws + kDenseSchedBytes;
  const long long T_max = (long long)B * tilesX * cdiv(H, kD4MinTileRows);
  uint16_t* ovf = reinterpret_cast<uint16_t*>(lists + dense_lists_bytes(T_max, Nmax));
  const SeedRec* srec = reinterpret_cast<const SeedRec*>(seeds);
  if (mode != 2) {
    if (mode == 1) {      // behind the seeds kernel of the pipeline: overlap the launch
      ISG_CUDA(launch_pdl(tile_lists_kernel, dim3((unsigned)cdiv64(T, 8)), dim3(256), 0, stream, srec, n_seeds, Nmax, B, H, W,
                          (int)Geo::TH, tilesX, tilesY, cap, lists, ovf, sched));
    } else {
      tile_lists_kernel<<<(unsigned)cdiv64(T, 8), 256, 0, stream>>>(srec, n_seeds, Nmax, B, H, W, Geo::TH, tilesX, tilesY, cap, lists,
                                                                    ovf, sched);
    }
    ISG_LAUNCH_CHECK();
    if (mode == 1) return ISG_OK;
  }
  static isg_encode_tiled_fn encode = get_encode_tiled();
  if (!encode) return ISG_EUNSUPPORTED;
  CUtensorMap tm_kp, tm_ae;
  if (KP) {
    const cuuint64_t dim[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    const cuuint64_t str[2] = {(cuuint64_t)W * 4, (cuuint64_t)kp_img_stride * 4};
    const cuuint32_t box[3] = {kD4KpW, (cuuint32_t)Geo::kKpRows, 1};
    const cuuint32_t es[3] = {1, 1, 1};
    if (encode(&tm_kp, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(kp), dim, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA) != CUDA_SUCCESS)
      return ISG_EUNSUPPORTED;
  }
  {
    const cuuint64_t dim[4] = {(cuuint64_t)W, (cuuint64_t)H, 4, (cuuint64_t)B};
    const cuuint64_t str[3] = {(cuuint64_t)W * 4, (cuuint64_t)ae_plane_stride * 4, (cuuint64_t)ae_img_stride * 4};
    const cuuint32_t box[4] = {kD4TileW, (cuuint32_t)Geo::TH, 4, 1};
    const cuuint32_t es[4] = {1, 1, 1, 1};
    if (encode(&tm_ae, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(ae), dim, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return ISG_EUNSUPPORTED;
  }
  if (!KP) tm_kp = tm_ae;                               // never dereferenced by the labels-only kernel
  int dev = 0, sms = kSMs;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const Tuning& tn = tuning();
  // one CTA per SM; `dense_spare` SMs stay free for the small kernels of neighbouring pipeline steps (engine.py)
  const int grid = (int)std::min<long long>(T, std::max(1, sms - tn.dense_spare));
  const size_t smem = (size_t)D4Smem<RW, WG, KP>(cap, nstages, G).total;
  const int Wwords = cdiv(W, 32);
  const int threads = 32 * (WG * G + 1);
  const int dyn_tail = tn.dense_tail;                 // tiles per CTA left to the dynamic scheduler
  const int dbg_flags = tn.dense_debug;               // measurement aid: bit 0 = consumers release tiles without computing
  const int skip_ae = tn.dense_skip_ae;               // seedless tiles: label 0 without loading the ae planes
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)threads); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  // programmatic dependent launch only behind our own pre-pass (mode 0); with prebuilt lists the predecessor is unknown
  cfg.attrs = attr; cfg.numAttrs = (mode != 0 || !tn.dense_pdl) ? 0 : 1;
  const float4* ghost4 = reinterpret_cast<const float4*>(ghost);
  const unsigned char* clists = lists;
  const uint16_t* covf = ovf;
  if (score_map) {
    ISG_CUDA(cudaFuncSetAttribute(dense_v4_kernel<RW, WG, G, true, KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ISG_CUDA(cudaLaunchKernelEx(&cfg, dense_v4_kernel<RW, WG, G, true, KP>, tm_kp, tm_ae, thr_key, srec, ghost4, Nmax, B, H, W, Wwords, tilesX,
                                tilesY, nstages, cap, clists, covf, ys, xs, label_map, score_map, keepbits, stats, sched, dyn_tail,
                                dbg_flags, skip_ae));
  } else {
    ISG_CUDA(cudaFuncSetAttribute(dense_v4_kernel<RW, WG, G, false, KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ISG_CUDA(cudaLaunchKernelEx(&cfg, dense_v4_kernel<RW, WG, G, false, KP>, tm_kp, tm_ae, thr_key, srec, ghost4, Nmax, B, H, W, Wwords, tilesX,
                                tilesY, nstages, cap, clists, covf, ys, xs, label_map, score_map, keepbits, stats, sched, dyn_tail,
                                dbg_flags, skip_ae));
  }
  ISG_LAUNCH_CHECK();
  return ISG_OK;
}

// Tuning knob for experiments: ISG_DENSE_CFG = "<RW>x<WG>x<G>[:stages]" selects one of the compiled geometries.
inline int launch_dense_v4(const float* kp, int64_t kp_img_stride, const float* ae, int64_t ae_img_stride,
                           int64_t ae_plane_stride, const uint32_t* thr_key, const uint32_t* seeds, const float* ghost,
                           const int32_t* n_seeds, int B, int Nmax, int H, int W, const float* ys, const float* xs,
                           int32_t* label_map, float* score_map, uint32_t* keepbits, int32_t* stats, void* workspace,
                           size_t workspace_bytes, int mode, cudaStream_t stream, bool with_kp = true) {
  const Tuning& tn = tuning();
  const int rw = tn.dense_rw, wg = tn.dense_wg, g = tn.dense_g, st = tn.dense_stages > 0 ? tn.dense_stages : kD4MaxStages;
#define ISG_V4_CASE(RW_, WG_, G_)                                                                                      \
  if (rw == RW_ && wg == WG_ && g == G_)                                                                               \
    return with_kp ? launch_dense_v4_cfg<RW_, WG_, G_, true>(kp, kp_img_stride, ae, ae_img_stride, ae_plane_stride, thr_key, seeds, ghost, \
                                             n_seeds, B, Nmax, H, W, ys, xs, label_map, score_map, keepbits, stats, workspace, workspace_bytes, st, mode, stream) \
                   : launch_dense_v4_cfg<RW_, WG_, G_, false>(kp, kp_img_stride, ae, ae_img_stride, ae_plane_stride, thr_key, seeds, ghost, \
                                             n_seeds, B, Nmax, H, W, ys, xs, label_map, score_map, keepbits, stats, workspace, workspace_bytes, st, mode, stream);
  ISG_V4_CASE(2, 8, 2)
  ISG_V4_CASE(4, 4, 3)
  ISG_V4_CASE(4, 4, 4)
  ISG_V4_CASE(4, 2, 6)
  ISG_V4_CASE(2, 4, 4)
  ISG_V4_CASE(2, 4, 6)
  ISG_V4_CASE(2, 8, 3)
#undef ISG_V4_CASE
  return ISG_EUNSUPPORTED;
}

}  // namespace isg
