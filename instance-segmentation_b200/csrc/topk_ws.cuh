// Layout of the top-k workspace (isg_topk_workspace_bytes), shared by the selection kernels (select.cu) and the dense
// kernel's one-pass mode (dense_v4.cuh), which appends the candidate pixels itself.
#pragma once
#include "common.cuh"

namespace isg {

constexpr int kHistBins = 2048;          // 11-bit digits (11 + 11 + 10 = 32)

struct TopkWs {            // per-image views
  uint32_t* hist;          // [3][2048]   (legacy multi-CTA path)
  uint32_t* lower;         // [0] lower bound key of the candidates, [2] rank inside the bin, [3] bin mode
  uint32_t* ncand;         // [1]  candidates appended so far (keeps counting past the capacity)
  uint32_t* cand;          // [cap_c] candidate keys
  uint32_t* pos;           // [cap_c] their pixel indices y*W + x (one-pass mode)
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
  v.cand = (uint32_t*)p; p += topk_cand_cap(npx, k) * sizeof(uint32_t);
  v.pos = (uint32_t*)p;
  return v;
}
// the candidate list provably contains the k-th largest key: at least k candidates, none lost to the capacity
__host__ __device__ inline bool topk_cand_usable(uint32_t ncand, uint32_t rank, int npx, int k) {
  return ncand >= rank && rank >= 1u && (size_t)ncand <= topk_cand_cap(npx, k);
}

}  // namespace isg
