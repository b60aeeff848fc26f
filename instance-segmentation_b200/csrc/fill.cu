// Polygon rasteriser on the device: poly_to_mask (reference utils/image.py:180-185, i.e.
// cv2.fillPoly(mask, [poly.astype(int32)], 1)) for many polygons at once, bit-packed output.
//
// One CTA per polygon.  The result of fillPoly for integer vertices is the union of
//   (1) the 8-connected Bresenham line of every edge, walked from its left end point to its right one, and
//   (2) the scan-line interior: an edge (x0,y0)-(x1,y1) with y0 < y1 is active on rows y0 <= y < y1 at the 16.16
//       fixed-point abscissa x0 + (y - y0) * dx, dx = ((x1 - x0) << 16) / (y1 - y0) (truncating division); on a row
//       the sorted abscissae are paired and the integer pixels ceil(xa) .. floor(xb) of each pair are set.
// (2) is evaluated without sorting: pixel p lies inside a pair iff an odd number of active edges lie strictly left of
// p, or an edge passes exactly through p.  Every edge toggles one bit per row (at floor(x) + 1) in a shared-memory bit
// row; a prefix XOR over the row turns the toggles into the inside mask.  oracle/ref_fill.py states both forms and
// tests/test_fill_oracle.py pins them against OpenCV.
//
// Output: for polygon i a block of rows x words uint32 (bit k of word w = pixel x0 + 32 w + k, row y0 + r) inside
// `words`, described by desc[i] = {status, x0, y0, rows, words_per_row, offset_lo, offset_hi, n_vertices}.  In
// full-frame mode the block is the whole H x ceil(W/32) frame at offset i * H * ceil(W/32) - the layout isg_mask_nms
// and isg_mask_pair_counts read; otherwise it is the polygon's bounding box (word aligned in x), allocated from
// `words` in completion order.
#include "common.cuh"

namespace isg {
namespace {

constexpr int kFillThreads = 256;
constexpr int kFillWarps = kFillThreads / 32;
constexpr int kFillTileWords = 4096;          // words per bit plane held in shared memory (2 planes, 32 KB)
constexpr int kFillThreadRowWords = 16;       // boxes up to this many words a row resolve with one thread per row

__global__ void __launch_bounds__(kFillThreads)
fill_polygons_kernel(const float2* __restrict__ pts, const int32_t* __restrict__ poly_start,
                     const int32_t* __restrict__ poly_count, int H, int W, int Wwords, int full_frame,
                     uint32_t* __restrict__ words, unsigned long long cap_words, int32_t* __restrict__ desc,
                     unsigned long long* __restrict__ total) {
  __shared__ uint32_t togg[kFillTileWords];   // edge crossings: bit p toggles the inside state from pixel p on
  __shared__ uint32_t cover[kFillTileWords];  // pixels set directly (edge lines, exact crossings)
  __shared__ int s_red[4][kFillWarps];
  __shared__ unsigned long long s_off;
  const int i = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int K = poly_count[i];
  int32_t* d = desc + (size_t)i * ISG_FILL_DESC_WORDS;
  if (K <= 0) {
    if (tid < ISG_FILL_DESC_WORDS) d[tid] = (tid == 0) ? ISG_FILL_EMPTY : 0;
    return;
  }
  const float2* v = pts + poly_start[i];
  // ---- bounding box of the (truncated) vertices ----
  int xmin = 0x7fffffff, ymin = 0x7fffffff, xmax = -0x7fffffff - 1, ymax = -0x7fffffff - 1;
  for (int k = tid; k < K; k += kFillThreads) {
    const float2 p = __ldg(v + k);
    const int x = (int)p.x, y = (int)p.y;                        // astype(np.int32): truncation
    xmin = min(xmin, x); xmax = max(xmax, x); ymin = min(ymin, y); ymax = max(ymax, y);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    xmin = min(xmin, __shfl_xor_sync(0xffffffffu, xmin, o)); ymin = min(ymin, __shfl_xor_sync(0xffffffffu, ymin, o));
    xmax = max(xmax, __shfl_xor_sync(0xffffffffu, xmax, o)); ymax = max(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
  }
  if (lane == 0) { s_red[0][warp] = xmin; s_red[1][warp] = ymin; s_red[2][warp] = xmax; s_red[3][warp] = ymax; }
  __syncthreads();
#pragma unroll
  for (int w = 0; w < kFillWarps; ++w) {
    xmin = min(xmin, s_red[0][w]); ymin = min(ymin, s_red[1][w]); xmax = max(xmax, s_red[2][w]); ymax = max(ymax, s_red[3][w]);
  }
  if (xmin < 0 || ymin < 0 || xmax >= W || ymax >= H) {           // OpenCV clips such edges; not restated
    if (tid < ISG_FILL_DESC_WORDS) d[tid] = (tid == 0) ? ISG_FILL_OUTSIDE : (tid == 7 ? K : 0);
    return;
  }
  const int w_lo = full_frame ? 0 : (xmin >> 5);
  const int Wb = full_frame ? Wwords : ((xmax >> 5) - w_lo + 1);
  const int Y0 = full_frame ? 0 : ymin;
  const int rows = full_frame ? H : (ymax - ymin + 1);
  const unsigned long long need = (unsigned long long)rows * (unsigned long long)Wb;
  if (tid == 0) s_off = full_frame ? (unsigned long long)i * need : atomicAdd(total, need);
  __syncthreads();
  const unsigned long long off = s_off;
  const bool fits = off + need <= cap_words;
  if (tid == 0) {
    d[0] = fits ? ISG_FILL_OK : ISG_FILL_OVERFLOW;
    d[1] = w_lo << 5; d[2] = Y0; d[3] = rows; d[4] = Wb;
    d[5] = (int32_t)(uint32_t)(off & 0xffffffffull); d[6] = (int32_t)(uint32_t)(off >> 32); d[7] = K;
  }
  if (!fits) return;
  uint32_t* out = words + off;

  const int chunk = max(1, min(rows, kFillTileWords / Wb));       // rows per pass (Wb <= kFillTileWords, host checked)
  for (int ya = 0; ya < rows; ya += chunk) {
    const int nr = min(chunk, rows - ya);
    const int lo = Y0 + ya, hi = lo + nr;                          // image rows [lo, hi) of this pass
    for (int k = tid; k < nr * Wb; k += kFillThreads) { togg[k] = 0u; cover[k] = 0u; }
    __syncthreads();
    // Every loop below has a warp-uniform trip count (the maximum over the warp, lanes past their own end are
    // predicated off): lanes with edges of different length would otherwise drift apart between the line loop and the
    // crossing loop and the warp would execute them as many separate fragments.
    for (int e0 = 0; e0 < K; e0 += kFillThreads) {
      const int e = e0 + tid;
      bool valid = e < K;
      int ax = 0, ay = 0, bx = 0, by = 0;
      if (valid) {
        const float2 pa = __ldg(v + (e == 0 ? K - 1 : e - 1)), pb = __ldg(v + e);
        ax = (int)pa.x; ay = (int)pa.y; bx = (int)pb.x; by = (int)pb.y;
        valid = !(max(ay, by) < lo || min(ay, by) >= hi);
      }
      // (1) the edge line, left end point first
      int x = ax, y = ay, sy = 1, major = -1, minor = 0, err = 0;
      bool steep = false;
      if (valid) {
        int x1 = bx, y1 = by;
        if (x1 < x) { x = bx; y = by; x1 = ax; y1 = ay; }
        const int dx = x1 - x;
        int dy = y1 - y;
        sy = dy < 0 ? -1 : 1;
        dy = dy < 0 ? -dy : dy;
        steep = dy > dx;
        major = steep ? dy : dx; minor = steep ? dx : dy;
        err = major - 2 * minor;
      }
      const int n_line = __reduce_max_sync(0xffffffffu, major) + 1;
      for (int s = 0; s < n_line; ++s) {
        if (s <= major) {
          if (y >= lo && y < hi) atomicOr(&cover[(y - lo) * Wb + (x >> 5) - w_lo], 1u << (x & 31));
          const bool neg = err < 0;
          err += -2 * minor + (neg ? 2 * major : 0);
          if (steep) { y += sy; x += neg ? 1 : 0; }
          else { x += 1; y += neg ? sy : 0; }
        }
      }
      // (2) scan-line crossings on rows [top, bottom) of the edge
      int n_rows = 0, yc = 0;
      long long X = 0, dxf = 0;
      if (valid && ay != by) {
        dxf = ((long long)(bx - ax) * 65536ll) / (long long)(by - ay);
        const int yt = min(ay, by), yb = max(ay, by);
        yc = max(yt, lo);
        n_rows = max(0, min(yb, hi) - yc);
        X = ((long long)(ay < by ? ax : bx) << 16) + (long long)(yc - yt) * dxf;
      }
      const int n_scan = __reduce_max_sync(0xffffffffu, n_rows);
      for (int k = 0; k < n_scan; ++k) {
        if (k < n_rows) {
          const int xi = (int)(X >> 16);
          const int rb = (yc + k - lo) * Wb - w_lo;
          if ((X & 0xffffll) == 0) atomicOr(&cover[rb + (xi >> 5)], 1u << (xi & 31));
          const int t = xi + 1;
          if ((t >> 5) - w_lo < Wb) atomicXor(&togg[rb + (t >> 5)], 1u << (t & 31));
          X += dxf;
        }
      }
    }
    __syncthreads();
    // ---- toggles -> inside mask (prefix XOR along the row) ----
    if (Wb <= kFillThreadRowWords) {
      // narrow boxes (the usual case): one thread per row, the carry between words stays in a register
      for (int r = tid; r < nr; r += kFillThreads) {
        uint32_t carry = 0u;
        for (int w = 0; w < Wb; ++w) {
          uint32_t t = togg[r * Wb + w];
          t ^= t << 1; t ^= t << 2; t ^= t << 4; t ^= t << 8; t ^= t << 16;
          t ^= carry;
          carry = (uint32_t)((int32_t)t >> 31);                        // all ones when the row is inside at the word's end
          out[(size_t)(ya + r) * Wb + w] = t | cover[r * Wb + w];
        }
      }
    } else
    // wide boxes / full frames: one warp per row, 32 words at a time
    for (int r = warp; r < nr; r += kFillWarps) {
      uint32_t carry = 0u;
      for (int wb = 0; wb < Wb; wb += 32) {
        const int w = wb + lane;
        uint32_t t = (w < Wb) ? togg[r * Wb + w] : 0u;
        t ^= t << 1; t ^= t << 2; t ^= t << 4; t ^= t << 8; t ^= t << 16;     // inclusive prefix XOR inside the word
        const uint32_t par = t >> 31;
        uint32_t incl = par;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl ^= u; }
        if ((incl ^ par ^ carry) & 1u) t = ~t;
        carry ^= __shfl_sync(0xffffffffu, incl, 31);
        if (w < Wb) out[(size_t)(ya + r) * Wb + w] = t | cover[r * Wb + w];
      }
    }
    __syncthreads();
  }
}

}  // namespace
}  // namespace isg

using namespace isg;

extern "C" int isg_fill_polygons(const float* points, const int32_t* poly_start, const int32_t* poly_count, int n, int H,
                                 int W, int full_frame, uint32_t* words, size_t cap_words, int32_t* desc,
                                 unsigned long long* total, isg_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!points || !poly_start || !poly_count || !words || !desc || !total || n <= 0 || H <= 0 || W <= 0) return ISG_EINVAL;
  if (((uintptr_t)points & 7u) != 0) return ISG_EINVAL;
  const int Wwords = cdiv(W, 32);
  if (Wwords > kFillTileWords) return ISG_EUNSUPPORTED;
  ISG_CUDA(cudaMemsetAsync(total, 0, sizeof(unsigned long long), stream));
  fill_polygons_kernel<<<(unsigned)n, kFillThreads, 0, stream>>>(reinterpret_cast<const float2*>(points), poly_start, poly_count, H,
                                                                 W, Wwords, full_frame ? 1 : 0, words,
                                                                 (unsigned long long)cap_words, desc, total);
  ISG_LAUNCH_CHECK();
  return ISG_OK;
}
