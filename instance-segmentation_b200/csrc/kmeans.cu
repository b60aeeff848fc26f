// K4 — seeded k-means with per-cluster allowed distance.
// Reference: kmeans / pairwise_distance / pairwise_cosine (utils/kmeans.py:16-130).
//
// The whole Lloyd loop is ONE cooperative launch (kmeans_loop_kernel): the host never looks at the convergence flag
// mid-run.  Per iteration:
//   assign   four threads per point, each against a quarter of the centres staged in shared memory, merged in index
//            order -> label (first index on ties, :57; outlier label N unless min distance < allow, :60-61);
//   update   per CTA, without atomics and in a fixed order: inside a warp the lowest lane of every label group
//            (match.any) adds its members in lane order, then the point warps add their groups to the CTA's fp64
//            accumulators in shared memory one warp after the other.  The CTA's partial sums / counts go to the
//            workspace;                                                                           O(M) work
//   -- grid barrier --
//   finish   warp per cluster: partials of all CTAs in CTA order (lane-strided + a fixed shuffle tree) -> mean ->
//            new centre (unchanged when empty, :74) and its shift (:72);                           O(N * CTAs) work
//   -- grid barrier --
//   every CTA adds the shifts in cluster order in fp32 (:64,72) and takes the same `center_shift^2 < tol` decision (:90).
// The additions of one cluster happen in a fixed order (lane order inside a warp, warp order inside a CTA, CTA order
// across CTAs): results are reproducible run to run.  The mean is accumulated in fp64 (the reference's fp32 `mean` differs by ~1e-7 relative).
#include <algorithm>
#include "common.cuh"

namespace isg {

constexpr int kMaxD = 16;
constexpr int kKmPoints = 128;       // points per CTA chunk
constexpr int kKmSlices = 4;         // threads per point: each scans a quarter of the centres
constexpr int kKmThreads = kKmPoints * kKmSlices;
constexpr int kKmMaxCtas = 2 * kSMs;

struct KmeansState {   // lives in the workspace
  unsigned int barrier;   // monotonic arrival counter of the grid barrier
  int done;
  int iters;
  float shift;
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

// all CTAs of a cooperative launch (co-resident by construction); `epoch` counts this CTA's barriers
__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int ctas, unsigned int& epoch) {
  __syncthreads();
  epoch += 1;
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    const unsigned int target = epoch * ctas;
    unsigned int seen;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
    } while (seen < target);
  }
  __syncthreads();
}

// nearest centre of one point among the centres [j0, j1): (distance, index) with torch.min's first-index rule (:57);
// an empty range returns (+inf, N)
template <int METRIC, int DT>
__device__ __forceinline__ void nearest(const float* x, const float* sc, const float* sn, int j0, int j1, int N, int D,
                                        float& best, int& arg) {
  const int DD = DT > 0 ? DT : D;
  best = INFINITY; arg = N;
  if (METRIC == ISG_KMEANS_EUCLIDEAN) {
    // sqrt is monotone: a centre whose SQUARED distance is not below the best one cannot win the strict `<` on the
    // rounded square roots, so the square root is only taken for improving candidates
    float best_s = INFINITY;
    for (int j = j0; j < j1; ++j) {
      const float* c = sc + j * DD;
      float s = 0.0f;
#pragma unroll
      for (int d = 0; d < DD; ++d) {
        const float t = __fsub_rn(x[d], c[d]);
        const float q = __fmul_rn(t, t);
        s = (d == 0) ? q : __fadd_rn(s, q);
      }
      if (j == j0 || s < best_s) {
        const float r = __fsqrt_rn(s);
        if (j == j0 || r < best) { best = r; arg = j; }
        best_s = s;
      }
    }
  } else {
    const float xn = vec_norm(x, DD);
    for (int j = j0; j < j1; ++j) {
      const float dj = point_dist<METRIC>(x, sc + j * DD, DD, xn, sn[j]);
      if (j == j0 || dj < best) { best = dj; arg = j; }
    }
  }
}

// A CTA works on chunks of kKmPoints points with kKmSlices threads per point: thread (slice s, point p) = s * kKmPoints
// + p scans the contiguous centre range of its slice (all lanes of a warp read the same centre: broadcast), the slices
// are merged in ascending order with a strict `<` (first index on ties).  More warps per SM hide the latency of the
// dependent compare chain; the loop itself is too short to fill the machine with one thread per point.
template <int METRIC, int DT>
__global__ void __launch_bounds__(kKmThreads)
kmeans_loop_kernel(const float* __restrict__ X, int M, int D, float* __restrict__ centers, const float* __restrict__ allow,
                   int N, float tol, int max_iter, int32_t* __restrict__ labels, double* __restrict__ psum /*[G][N][D]*/,
                   int32_t* __restrict__ pcnt /*[G][N]*/, float* __restrict__ shift /*[N], < 0 = empty*/,
                   KmeansState* __restrict__ st) {
  const int DD = DT > 0 ? DT : D;
  extern __shared__ __align__(16) unsigned char km_smem[];
  double* acc = reinterpret_cast<double*>(km_smem);                 // [N*DD] fp64 sums of this CTA's members
  float* sc = reinterpret_cast<float*>(acc + (size_t)N * DD);       // [N*DD] centres
  float* sn = sc + (size_t)N * DD;                                  // [N]    centre norms (cosine) / shifts (finish)
  int32_t* cnt = reinterpret_cast<int32_t*>(sn + N);                // [N]
  __shared__ float s_best[kKmSlices][kKmPoints];
  __shared__ int s_arg[kKmSlices][kKmPoints];
  __shared__ int s_done;
  const int t = threadIdx.x, G = gridDim.x, cta = blockIdx.x;
  const int lane = t & 31, warp = t >> 5;
  const int slice = t / kKmPoints, p = t - slice * kKmPoints;
  const int per = (N + kKmSlices - 1) / kKmSlices;
  const int j0 = min(slice * per, N), j1 = min(j0 + per, N);
  unsigned int epoch = 0;
  int it = 0;
  for (;;) {
    for (int i = t; i < N * DD; i += kKmThreads) { sc[i] = centers[i]; acc[i] = 0.0; }
    for (int i = t; i < N; i += kKmThreads) cnt[i] = 0;
    __syncthreads();
    if (METRIC == ISG_KMEANS_COSINE) {
      for (int i = t; i < N; i += kKmThreads) sn[i] = vec_norm(sc + i * DD, DD);
      __syncthreads();
    }
    for (int base = cta * kKmPoints; base < M; base += G * kKmPoints) {
      const int m = base + p;
      float x[DT > 0 ? DT : kMaxD];
      float best = INFINITY;
      int arg = N;
      if (m < M) {
#pragma unroll
        for (int d = 0; d < DD; ++d) x[d] = X[(size_t)m * DD + d];
        nearest<METRIC, DT>(x, sc, sn, j0, j1, N, D, best, arg);
      }
      s_best[slice][p] = best; s_arg[slice][p] = arg;
      __syncthreads();
      int lab = -1;
      if (slice == 0 && m < M) {
#pragma unroll
        for (int s2 = 1; s2 < kKmSlices; ++s2) {
          const float b2 = s_best[s2][p];
          if (b2 < best) { best = b2; arg = s_arg[s2][p]; }       // strict: the earlier slice (lower indices) keeps ties
        }
        lab = (best < allow[arg]) ? arg : N;                       // strict (:60-61)
        labels[m] = lab;
      }
      // per-cluster sums of the chunk, deterministic and without atomics: inside a warp the lowest lane of every label
      // group adds its members in lane order (shuffles); the point warps then update the CTA's accumulators one after
      // the other (leaders of one warp hold distinct labels)
      if (slice == 0) {
        const bool valid = lab >= 0 && lab < N;
        const unsigned peers = __match_any_sync(0xffffffffu, valid ? lab : -1 - lane);
        const bool leader = valid && lane == __ffs(peers) - 1;
        double a[DT > 0 ? DT : kMaxD];
#pragma unroll
        for (int d = 0; d < DD; ++d) a[d] = 0.0;
        for (int src = 0; src < 32; ++src) {
#pragma unroll
          for (int d = 0; d < DD; ++d) {
            const float xv = __shfl_sync(0xffffffffu, (m < M) ? x[d] : 0.0f, src);
            if (leader && ((peers >> src) & 1u)) a[d] += (double)xv;
          }
        }
        for (int w = 0; w < kKmPoints / 32; ++w) {
          if (warp == w && leader) {
            cnt[lab] += __popc(peers);
#pragma unroll
            for (int d = 0; d < DD; ++d) acc[lab * DD + d] += a[d];
          }
          asm volatile("bar.sync 1, %0;" :: "n"(kKmPoints) : "memory");   // the point warps only
        }
      }
      __syncthreads();
    }
    for (int i = t; i < N * DD; i += kKmThreads) psum[(size_t)cta * N * DD + i] = acc[i];
    for (int i = t; i < N; i += kKmThreads) pcnt[(size_t)cta * N + i] = cnt[i];
    grid_barrier(&st->barrier, G, epoch);

    // finish: warp per cluster over the CTAs' partials
    for (int k = cta * (kKmThreads / 32) + warp; k < N; k += G * (kKmThreads / 32)) {
      double a[DT > 0 ? DT : kMaxD];
#pragma unroll
      for (int d = 0; d < DD; ++d) a[d] = 0.0;
      int c = 0;
      for (int g = lane; g < G; g += 32) {
        c += pcnt[(size_t)g * N + k];
#pragma unroll
        for (int d = 0; d < DD; ++d) a[d] += psum[((size_t)g * N + k) * DD + d];
      }
      c = warp_sum(c);
#pragma unroll
      for (int d = 0; d < DD; ++d) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a[d] += __shfl_xor_sync(0xffffffffu, a[d], o);
      }
      if (lane == 0) {
        float sh = -1.0f;                           // empty cluster: centre unchanged, no shift term (:73-74)
        if (c > 0) {
          float s = 0.0f;
#pragma unroll
          for (int d = 0; d < DD; ++d) {
            const float nc = (float)(a[d] / (double)c);                     // mean of the members (:70-71)
            const float dlt = __fsub_rn(nc, sc[k * DD + d]);
            s = (d == 0) ? __fmul_rn(dlt, dlt) : __fadd_rn(s, __fmul_rn(dlt, dlt));
            centers[(size_t)k * DD + d] = nc;
          }
          sh = __fsqrt_rn(s);                                               // (:72)
        }
        shift[k] = sh;
      }
    }
    grid_barrier(&st->barrier, G, epoch);

    for (int i = t; i < N; i += kKmThreads) sn[i] = shift[i];
    __syncthreads();
    ++it;
    if (t == 0) {
      float cs = 0.0f;                              // center_shift accumulates in cluster order, fp32 (:64,72)
      for (int k = 0; k < N; ++k)
        if (sn[k] >= 0.0f) cs = __fadd_rn(cs, sn[k]);
      const int done = __fmul_rn(cs, cs) < tol;     // center_shift ** 2 < tol (:90)
      s_done = done;
      if (cta == 0) { st->shift = cs; st->iters = it; st->done = done; }
    }
    __syncthreads();
    if (s_done || (max_iter > 0 && it >= max_iter)) break;
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

static int km_max_ctas(int M) { return std::max(1, std::min(cdiv(M, kKmPoints), kKmMaxCtas)); }
static size_t km_smem_bytes(int N, int D) {
  return (size_t)N * D * 8 + (size_t)N * D * 4 + (size_t)N * 8;
}

extern "C" size_t isg_kmeans_workspace_bytes(int M, int N, int D) {
  if (M < 0 || N <= 0 || D <= 0) return 0;
  const size_t G = (size_t)km_max_ctas(M);
  return align256(sizeof(KmeansState)) + align256(G * N * D * 8) + align256(G * N * 4) + align256((size_t)N * 4);
}

template <int METRIC, int DT>
static int km_launch(const float* X, int M, int D, float* centers, const float* allow, int N, float tol, int max_iter,
                     int32_t* labels, double* psum, int32_t* pcnt, float* shift, KmeansState* st, cudaStream_t stream) {
  auto kern = kmeans_loop_kernel<METRIC, DT>;
  const size_t smem = km_smem_bytes(N, D);
  ISG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int dev = 0, sms = 0, per_sm = 0;
  ISG_CUDA(cudaGetDevice(&dev));
  ISG_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  ISG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kKmThreads, smem));
  if (per_sm < 1) return ISG_EUNSUPPORTED;
  const int G = std::min(km_max_ctas(M), per_sm * sms);   // cooperative launch: every CTA resident
  void* args[] = {(void*)&X, (void*)&M, (void*)&D, (void*)&centers, (void*)&allow, (void*)&N, (void*)&tol, (void*)&max_iter,
                  (void*)&labels, (void*)&psum, (void*)&pcnt, (void*)&shift, (void*)&st};
  ISG_CUDA(cudaLaunchCooperativeKernel((const void*)kern, dim3(G), dim3(kKmThreads), args, smem, stream));
  return ISG_OK;
}

extern "C" int isg_kmeans(const float* X, int M, int D, float* centers, const float* allow, int N, float tol,
                          int metric, int max_iter, int32_t* labels, int* iters_host, void* ws, size_t ws_bytes,
                          isg_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!X || !centers || !allow || !labels || M <= 0 || N <= 0 || D <= 0) return ISG_EINVAL;
  if (D > kMaxD) return ISG_EUNSUPPORTED;
  if (metric != ISG_KMEANS_EUCLIDEAN && metric != ISG_KMEANS_COSINE) return ISG_EINVAL;
  if (!ws || ws_bytes < isg_kmeans_workspace_bytes(M, N, D) || ((uintptr_t)ws & 255)) return ISG_EWORKSPACE;
  if (km_smem_bytes(N, D) > 200 * 1024) return ISG_EUNSUPPORTED;
  const size_t G = (size_t)km_max_ctas(M);
  char* p = (char*)ws;
  KmeansState* st = (KmeansState*)p; p += align256(sizeof(KmeansState));
  double* psum = (double*)p;         p += align256(G * N * D * 8);
  int32_t* pcnt = (int32_t*)p;       p += align256(G * N * 4);
  float* shift = (float*)p;
  ISG_CUDA(cudaMemsetAsync(st, 0, sizeof(KmeansState), stream));
  int rc;
  if (metric == ISG_KMEANS_EUCLIDEAN)
    rc = (D == 2) ? km_launch<ISG_KMEANS_EUCLIDEAN, 2>(X, M, D, centers, allow, N, tol, max_iter, labels, psum, pcnt, shift, st, stream)
                  : km_launch<ISG_KMEANS_EUCLIDEAN, 0>(X, M, D, centers, allow, N, tol, max_iter, labels, psum, pcnt, shift, st, stream);
  else
    rc = (D == 2) ? km_launch<ISG_KMEANS_COSINE, 2>(X, M, D, centers, allow, N, tol, max_iter, labels, psum, pcnt, shift, st, stream)
                  : km_launch<ISG_KMEANS_COSINE, 0>(X, M, D, centers, allow, N, tol, max_iter, labels, psum, pcnt, shift, st, stream);
  if (rc != ISG_OK) return rc;
  ISG_LAUNCH_CHECK();
  KmeansState h;
  ISG_CUDA(cudaMemcpyAsync(&h, st, sizeof(KmeansState), cudaMemcpyDeviceToHost, stream));
  ISG_CUDA(cudaStreamSynchronize(stream));
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
