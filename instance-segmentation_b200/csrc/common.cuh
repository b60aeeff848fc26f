// Shared device/host helpers for libisg.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/isg.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libisg is written for sm_100a (B200) only"
#endif

#define ISG_LAUNCH_CHECK()                                   \
  do {                                                       \
    cudaError_t e__ = cudaGetLastError();                    \
    if (e__ != cudaSuccess) return (int)e__;                 \
  } while (0)

#define ISG_CUDA(call)                                       \
  do {                                                       \
    cudaError_t e__ = (call);                                \
    if (e__ != cudaSuccess) return (int)e__;                 \
  } while (0)

namespace isg {

constexpr int kSMs = 148;  // B200: 2 dies x 74 SMs

// Order-preserving map fp32 -> uint32 (larger float <=> larger key; -0 < +0, NaN above +inf).
__host__ __device__ __forceinline__ uint32_t float_key(float f) {
#ifdef __CUDA_ARCH__
  uint32_t u = __float_as_uint(f);
#else
  union { float f; uint32_t u; } c; c.f = f; uint32_t u = c.u;
#endif
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// streaming 128-bit load that does not allocate in L1 (data touched once)
__device__ __forceinline__ float4 ldg_stream4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ float ldg_stream1(const float* p) {
  float r;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
  return r;
}
// streaming 128-bit store (evict-first: written once, never re-read by this kernel)
__device__ __forceinline__ void stg_stream4(int32_t* p, int a, int b, int c, int d) {
  asm volatile("st.global.cs.v4.s32 [%0], {%1,%2,%3,%4};" :: "l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void stg_stream4f(float* p, float a, float b, float c, float d) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// ---- mbarrier + 1-D bulk async copy (TMA unit, SASS UBLKCP) -------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" :: "r"(smem_u32(bar)), "r"(phase) : "memory");
}
// bytes must be a multiple of 16, src/dst 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ---- fast, accurate transcendentals ----------------------------------------------------------
// exp(x) = 2^(x*log2e): the product is split into its rounded value t and the exact rounding residual e
// (one FMA), so the only remaining error is ex2.approx's (<= 2 ulp); subnormal results are kept.
__device__ __forceinline__ float ex2_approx(float t) {
  float r;
  asm("ex2.approx.f32 %0, %1;" : "=f"(r) : "f"(t));
  return r;
}
// Finite arguments only: for |x| beyond ~2e38 (or an overflowing result) the residual term turns the result
// into NaN instead of 0 / +inf.  In the membership that is harmless: a NaN probability never wins the
// `P > best` comparison, exactly like the 0 the reference would produce.
__device__ __forceinline__ float exp_fast(float x) {
  const float L2E_HI = 1.44269502162933349609375f, L2E_LO = 1.925963033500011e-8f;
  const float t = x * L2E_HI;
  float e = fmaf(x, L2E_HI, -t);
  e = fmaf(x, L2E_LO, e);
  const float r = ex2_approx(t);
  return fmaf(r * 0.693147182464599609375f, e, r);   // 2^(t+e) ~ 2^t * (1 + e*ln2)
}
// Same, with ex2.approx.ftz: one MUFU instead of the range-scaled sequence nvcc emits for the non-ftz form (3 more
// instructions per call).  Results below 2^-126 are flushed to 0 - used where such a value cannot matter (the
// per-pixel sigma: a subnormal sigma contributes < 1e-37 to the exponent).
__device__ __forceinline__ float exp_fast_ftz(float x) {
  const float L2E_HI = 1.44269502162933349609375f, L2E_LO = 1.925963033500011e-8f;
  const float t = x * L2E_HI;
  float e = fmaf(x, L2E_HI, -t);
  e = fmaf(x, L2E_LO, e);
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t));
  return fmaf(r * 0.693147182464599609375f, e, r);
}
// Bare ex2.approx.ftz(x * log2 e): the rounding of the product adds |x| * log2e * 2^-24 * ln2 (3e-7 relative at
// |x| = 5.3, the sigma of a 200-unit-per-px^2 Gaussian) to the unit's 2 ulp.  Used for the per-pixel sigma of the DENSE
// kernel only, where sigma scales the exponent q whose ORDER decides the label: a relative error d of sigma moves
// exp(-q) by q*d - inside the 1e-5 membership tolerance up to q ~ 30, where the membership itself is below 1e-13.
__device__ __forceinline__ float exp_bare_ftz(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x * 1.44269502162933349609375f));
  return r;
}
// tanh: |x| < 0.55 -> x + x^3 * P(x^2) (degree-4 minimax fit, < 1 ulp); otherwise 1 - 2/(1 + e^{2|x|})
__device__ __forceinline__ float tanh_poly(float x) {
  const float u = x * x;
  float p = -0.00661575747653842f;
  p = fmaf(p, u, 0.02131274715065956f);
  p = fmaf(p, u, -0.053910065442323685f);
  p = fmaf(p, u, 0.13333117961883545f);
  p = fmaf(p, u, -0.3333333134651184f);
  return fmaf(x * u, p, x);
}
// |x| >= 0.55: 1 - 2/(e^{2|x|} + 1) with e^{2|x|} = ex2.approx(2|x| log2 e).  A relative error d of the exponential moves
// the result by 2e/(e+1)^2 * d <= 0.38 d here, so neither the compensated argument of exp_fast nor an exact division
// is needed: ex2.approx (2 ulp) + the rounding of the argument + rcp.approx stay below 1 ulp of the result (measured:
// tests/test_gpu_parity.py::test_fast_tanh_exp_accuracy).  Overflow: e = +inf -> 2/inf = 0 -> 1.
__device__ __forceinline__ float tanh_large(float x) {
  const float a = fabsf(x);
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(a * 2.88539008177792681472f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
  return copysignf(fmaf(-2.0f, r, 1.0f), x);
}
__device__ __forceinline__ float tanh_fast(float x) { return fabsf(x) < 0.55f ? tanh_poly(x) : tanh_large(x); }

// ---- packed fp32x2 arithmetic (sm_100: FADD2 / FMUL2 / FFMA2) -----------------------------------------------------
// Two IEEE round-to-nearest fp32 operations per issued instruction; every component is rounded exactly like the scalar
// __fadd_rn / __fmul_rn / fmaf.  NOTE: ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 (unlike the scalar .rn
// forms), so wherever the reference needs a separately rounded product and sum the final addition is done with scalar
// __fadd_rn on the unpacked halves.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) { f32x2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
// tanh_poly on two values (the same operations as the scalar form, component by component)
__device__ __forceinline__ f32x2 tanh_poly2(f32x2 x) {
  const f32x2 u = mul2(x, x);
  f32x2 p = pack2(-0.00661575747653842f, -0.00661575747653842f);
  p = fma2(p, u, pack2(0.02131274715065956f, 0.02131274715065956f));
  p = fma2(p, u, pack2(-0.053910065442323685f, -0.053910065442323685f));
  p = fma2(p, u, pack2(0.13333117961883545f, 0.13333117961883545f));
  p = fma2(p, u, pack2(-0.3333333134651184f, -0.3333333134651184f));
  return fma2(mul2(x, u), p, x);
}
// exp_fast_ftz on two values
__device__ __forceinline__ f32x2 exp_fast_ftz2(f32x2 x) {
  const f32x2 hi = pack2(1.44269502162933349609375f, 1.44269502162933349609375f);
  const f32x2 t = mul2(x, hi);
  f32x2 e = fma2(x, hi, t ^ 0x8000000080000000ull);                       // x * L2E_HI - t (exact residual)
  e = fma2(x, pack2(1.925963033500011e-8f, 1.925963033500011e-8f), e);
  float t0, t1, r0, r1;
  unpack2(t, t0, t1);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(t0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(t1));
  const f32x2 r = pack2(r0, r1);
  return fma2(mul2(r, pack2(0.693147182464599609375f, 0.693147182464599609375f)), e, r);
}

__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- programmatic dependent launch (PDL) --------------------------------------------------------------------------
// Kernels of one stream that consume each other's output are launched with programmatic stream serialisation: the
// consumer's CTAs may be scheduled while the producer grid is still draining (hides the launch latency, ~2 us per
// kernel boundary), and the consumer calls pdl_wait() before it touches anything the producer wrote - the wait returns
// when the whole preceding grid has completed and its writes are visible.  pdl_trigger() lets the next grid in.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Tuning knobs for experiments.  Read from the environment ONCE (first use) - never on the launch path; the only way to
// change them afterwards is isg_debug_reload_tuning() (tests and tools/sweep_dense.py).  Defaults are the shipped path.
struct Tuning {
  int dense_rw = 2, dense_wg = 8, dense_g = 2, dense_stages = 0;   // ISG_DENSE_CFG = "<RW>x<WG>x<G>[:stages]"
  int dense_tail = 4;        // ISG_DENSE_TAIL: tiles per CTA left to the dynamic scheduler
  int dense_debug = 0;       // ISG_DENSE_DEBUG: bit 0 = consumers release tiles without computing (loads-only ceiling)
  int dense_pdl = 1;         // ISG_DENSE_PDL=0: no programmatic dependent launch behind the pre-pass
  int dense_v1 = 0;          // ISG_DENSE_V1=1: plain-LDG kernel (also serves W % 4 != 0)
  int dense_v1_rw = 4;       // ISG_DENSE_RW: rows per warp of the v1 kernel
  int dense_spare = 0;       // ISG_DENSE_SPARE: SMs the persistent dense kernel leaves to concurrently running kernels
  int dense_skip_ae = 1;     // ISG_DENSE_SKIP_AE=0: load the ae planes of tiles no seed box overlaps too (A/B measurements)
  int topk_radix = 0;        // ISG_TOPK_PATH=radix: sampling-free two-level radix select
  int topk_cluster_sample = 0;   // ISG_TOPK_SAMPLE=cluster
  int topk_cluster_select = 0;   // ISG_TOPK_SELECT=cluster: the 8-CTA cluster form of the select step (64 CTAs per batch of 8)
  int nms_rounds = 32;       // ISG_NMS_ROUNDS: rounds of the parallel suppression scan before the sequential fall-back (0 = sequential only)
};
const Tuning& tuning();      // api.cu

inline int cdiv(int a, int b) { return (a + b - 1) / b; }
inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace isg
