// K4 — seeded k-means with per-cluster allowed distance.
// Reference: kmeans / pairwise_distance / pairwise_cosine (utils/kmeans.py:16-130).
//
// One Lloyd iteration = assign (thread per point, centres in shared memory) + update (one warp per
// cluster, fp64 shuffle reduction over the label list: deterministic, no atomics) + finish (centre
// shift in cluster order, convergence flag).  Iterations are enqueued in batches; every kernel is a
// no-op once the device-side `done` flag is set, so the host only reads the flag once per batch.
#include <algorithm>
#include "common.cuh"

namespace isg {

constexpr int kMaxD = 16;

struct KmeansState {   // lives in the workspace
  int done;
  int iters;
  float shift;
  int pad;
};

template <int METRIC>
__device__ __forceinline__ float point_dist(const float* x, const float* c, int D, float xnorm, float cnorm) {
  if (METRIC == ISG_KMEANS_EUCLIDEAN) {
    // dis = ((A - B) ** 2.0).sum(-1).sqrt()          (utils/kmeans.py:106-109)
    float s = 0.0f;
    for (int d = 0; d < D; ++d) {
      const float t = __fsub_rn(x[d], c[d]);
      const float q = __fmul_rn(t, t);
      s = (d == 0) ? q : __fadd_rn(s, q);
    }
    return __fsqrt_rn(s);
  } else {
    // 1 - sum(A/|A| * B/|B|)                          (utils/kmeans.py:123-130)
    float s = 0.0f;
    for (int d = 0; d < D; ++d) {
      const float q = __fmul_rn(__fdiv_rn(x[d], xnorm), __fdiv_rn(c[d], cnorm));
      s = (d == 0) ? q : __fadd_rn(s, q);
    }
    return __fsub_rn(1.0f, s);
  }
}

__device__ __forceinline__ float vec_norm(const float* v, int D) {
  float s = 0.0f;
  for (int d = 0; d < D; ++d) s = (d == 0) ? __fmul_rn(v[d], v[d]) : __fadd_rn(s, __fmul_rn(v[d], v[d]));
  return __fsqrt_rn(s);
}

template <int METRIC>
__global__ void __launch_bounds__(256)
kmeans_assign_kernel(const float* __restrict__ X, int M, int D, const float* __restrict__ centers,
                     const float* __restrict__ allow, int N, int32_t* __restrict__ labels,
                     const KmeansState* __restrict__ st) {
  if (st->done) return;
  extern __shared__ float sm[];
  float* sc = sm;             // [N*D]
  float* sn = sm + N * D;     // [N] centre norms (cosine)
  for (int i = threadIdx.x; i < N * D; i += blockDim.x) sc[i] = centers[i];
  __syncthreads();
  if (METRIC == ISG_KMEANS_COSINE) {
    for (int i = threadIdx.x; i < N; i += blockDim.x) sn[i] = vec_norm(sc + i * D, D);
    __syncthreads();
  }
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  float x[kMaxD];
  for (int d = 0; d < D; ++d) x[d] = X[(size_t)m * D + d];
  const float xn = (METRIC == ISG_KMEANS_COSINE) ? vec_norm(x, D) : 0.0f;
  float best = 0.0f;
  int arg = 0;
  for (int j = 0; j < N; ++j) {
    const float dj = point_dist<METRIC>(x, sc + j * D, D, xn, (METRIC == ISG_KMEANS_COSINE) ? sn[j] : 0.0f);
    if (j == 0 || dj < best) { best = dj; arg = j; }   // torch.min: first index on ties (:57)
  }
  // cluster id = num_clusters unless min distance < allowed distance (strict, :60-61)
  labels[m] = (best < allow[arg]) ? arg : N;
}

constexpr int kUpdateWarps = 8;

__global__ void __launch_bounds__(32 * kUpdateWarps)
kmeans_update_kernel(const float* __restrict__ X, int M, int D, const float* __restrict__ centers, int N,
                     const int32_t* __restrict__ labels, float* __restrict__ cnew, float* __restrict__ shift,
                     int32_t* __restrict__ nonempty, const KmeansState* __restrict__ st) {
  if (st->done) return;
  const int k = blockIdx.x * kUpdateWarps + (threadIdx.x >> 5);
  if (k >= N) return;
  const int lane = threadIdx.x & 31;
  double acc[kMaxD];
  for (int d = 0; d < D; ++d) acc[d] = 0.0;
  int cnt = 0;
  for (int m = lane; m < M; m += 32) {
    if (labels[m] == k) {
      ++cnt;
      for (int d = 0; d < D; ++d) acc[d] += (double)X[(size_t)m * D + d];
    }
  }
  cnt = warp_sum(cnt);
  for (int d = 0; d < D; ++d) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[d] += __shfl_xor_sync(0xffffffffu, acc[d], o);
  }
  if (lane == 0) {
    float sh = 0.0f;
    if (cnt > 0) {
      // new centre = mean of the members (:70-72); shift = sqrt(sum((new - old)^2)) (:72)
      float s = 0.0f;
      for (int d = 0; d < D; ++d) {
        const float nc = (float)(acc[d] / (double)cnt);
        cnew[(size_t)k * D + d] = nc;
        const float t = __fsub_rn(nc, centers[(size_t)k * D + d]);
        s = (d == 0) ? __fmul_rn(t, t) : __fadd_rn(s, __fmul_rn(t, t));
      }
      sh = __fsqrt_rn(s);
    } else {
      for (int d = 0; d < D; ++d) cnew[(size_t)k * D + d] = centers[(size_t)k * D + d];   // unchanged (:74)
    }
    shift[k] = sh;
    nonempty[k] = cnt > 0;
  }
}

__global__ void __launch_bounds__(256)
kmeans_finish_kernel(float* __restrict__ centers, const float* __restrict__ cnew, int N, int D,
                     const float* __restrict__ shift, const int32_t* __restrict__ nonempty, float tol,
                     KmeansState* __restrict__ st) {
  if (st->done) return;   // uniform: every thread reads the flag before thread 0 may set it (sync below)
  __syncthreads();
  for (int i = threadIdx.x; i < N * D; i += blockDim.x) centers[i] = cnew[i];
  if (threadIdx.x == 0) {
    float cs = 0.0f;   // center_shift accumulates in cluster order, fp32 (:64,72)
    for (int k = 0; k < N; ++k)
      if (nonempty[k]) cs = __fadd_rn(cs, shift[k]);
    st->shift = cs;
    st->iters += 1;
    if (__fmul_rn(cs, cs) < tol) st->done = 1;   // center_shift ** 2 < tol (:90)
  }
}

template <int METRIC>
__global__ void pairwise_kernel(const float* __restrict__ X, int M, const float* __restrict__ Y, int N, int D,
                                float* __restrict__ out) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long long)M * N) return;
  const int m = (int)(e / N), j = (int)(e - (long long)m * N);
  float x[kMaxD], y[kMaxD];
  for (int d = 0; d < D; ++d) { x[d] = X[(size_t)m * D + d]; y[d] = Y[(size_t)j * D + d]; }
  const float xn = (METRIC == ISG_KMEANS_COSINE) ? vec_norm(x, D) : 0.0f;
  const float yn = (METRIC == ISG_KMEANS_COSINE) ? vec_norm(y, D) : 0.0f;
  out[e] = point_dist<METRIC>(x, y, D, xn, yn);
}

}  // namespace isg

using namespace isg;

static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

extern "C" size_t isg_kmeans_workspace_bytes(int M, int N, int D) {
  if (M < 0 || N <= 0 || D <= 0) return 0;
  return align256(sizeof(KmeansState)) + align256((size_t)N * D * 4) + align256((size_t)N * 4) * 2;
}

extern "C" int isg_kmeans(const float* X, int M, int D, float* centers, const float* allow, int N, float tol,
                          int metric, int max_iter, int32_t* labels, int* iters_host, void* ws, size_t ws_bytes,
                          isg_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!X || !centers || !allow || !labels || M <= 0 || N <= 0 || D <= 0) return ISG_EINVAL;
  if (D > kMaxD) return ISG_EUNSUPPORTED;
  if (metric != ISG_KMEANS_EUCLIDEAN && metric != ISG_KMEANS_COSINE) return ISG_EINVAL;
  if (!ws || ws_bytes < isg_kmeans_workspace_bytes(M, N, D) || ((uintptr_t)ws & 255)) return ISG_EWORKSPACE;
  const size_t smem = ((size_t)N * D + N) * sizeof(float);
  if (smem > 200 * 1024) return ISG_EUNSUPPORTED;
  char* p = (char*)ws;
  KmeansState* st = (KmeansState*)p; p += align256(sizeof(KmeansState));
  float* cnew = (float*)p;           p += align256((size_t)N * D * 4);
  float* shift = (float*)p;          p += align256((size_t)N * 4);
  int32_t* nonempty = (int32_t*)p;
  ISG_CUDA(cudaMemsetAsync(st, 0, sizeof(KmeansState), stream));
  if (metric == ISG_KMEANS_EUCLIDEAN)
    ISG_CUDA(cudaFuncSetAttribute(kmeans_assign_kernel<ISG_KMEANS_EUCLIDEAN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  else
    ISG_CUDA(cudaFuncSetAttribute(kmeans_assign_kernel<ISG_KMEANS_COSINE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int batch = 4;
  int it = 0;
  KmeansState h;
  h.done = 0; h.iters = 0; h.shift = 0.f; h.pad = 0;
  while (max_iter <= 0 || it < max_iter) {
    const int nb = (max_iter > 0) ? std::min(batch, max_iter - it) : batch;
    for (int r = 0; r < nb; ++r) {
      if (metric == ISG_KMEANS_EUCLIDEAN)
        kmeans_assign_kernel<ISG_KMEANS_EUCLIDEAN><<<cdiv(M, 256), 256, smem, stream>>>(X, M, D, centers, allow, N, labels, st);
      else
        kmeans_assign_kernel<ISG_KMEANS_COSINE><<<cdiv(M, 256), 256, smem, stream>>>(X, M, D, centers, allow, N, labels, st);
      kmeans_update_kernel<<<cdiv(N, kUpdateWarps), 32 * kUpdateWarps, 0, stream>>>(X, M, D, centers, N, labels, cnew,
                                                                                 shift, nonempty, st);
      kmeans_finish_kernel<<<1, 256, 0, stream>>>(centers, cnew, N, D, shift, nonempty, tol, st);
    }
    ISG_LAUNCH_CHECK();
    it += nb;
    ISG_CUDA(cudaMemcpyAsync(&h, st, sizeof(KmeansState), cudaMemcpyDeviceToHost, stream));
    ISG_CUDA(cudaStreamSynchronize(stream));
    if (h.done) break;
  }
  if (iters_host) *iters_host = h.iters;
  return h.done ? ISG_OK : ISG_ENOTCONVERGED;
}

extern "C" int isg_pairwise(const float* X, int M, const float* Y, int N, int D, int metric, float* out,
                            isg_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!X || !Y || !out || M <= 0 || N <= 0 || D <= 0) return ISG_EINVAL;
  if (D > kMaxD) return ISG_EUNSUPPORTED;
  const long long tot = (long long)M * N;
  const int blocks = (int)((tot + 255) / 256);
  if (metric == ISG_KMEANS_EUCLIDEAN) pairwise_kernel<ISG_KMEANS_EUCLIDEAN><<<blocks, 256, 0, stream>>>(X, M, Y, N, D, out);
  else if (metric == ISG_KMEANS_COSINE) pairwise_kernel<ISG_KMEANS_COSINE><<<blocks, 256, 0, stream>>>(X, M, Y, N, D, out);
  else return ISG_EINVAL;
  ISG_LAUNCH_CHECK();
  return ISG_OK;
}
