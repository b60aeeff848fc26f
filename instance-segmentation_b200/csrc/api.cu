// Library-level entry points of libisg.so.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include "common.cuh"

namespace isg {
static Tuning read_tuning() {
  Tuning t;
  if (const char* e = getenv("ISG_DENSE_CFG")) {
    int a = 0, b = 0, c = 0, d = 0;
    const int got = sscanf(e, "%dx%dx%d:%d", &a, &b, &c, &d);
    if (got >= 3) { t.dense_rw = a; t.dense_wg = b; t.dense_g = c; }
    if (got >= 4 && d > 0) t.dense_stages = d;
  }
  if (const char* e = getenv("ISG_DENSE_TAIL")) t.dense_tail = atoi(e) > 0 ? atoi(e) : 0;
  if (const char* e = getenv("ISG_DENSE_DEBUG")) t.dense_debug = atoi(e);
  if (const char* e = getenv("ISG_DENSE_PDL")) t.dense_pdl = e[0] != '0';
  if (const char* e = getenv("ISG_DENSE_V1")) t.dense_v1 = e[0] == '1';
  if (const char* e = getenv("ISG_DENSE_RW")) t.dense_v1_rw = atoi(e);
  if (const char* e = getenv("ISG_DENSE_SPARE")) t.dense_spare = atoi(e) > 0 ? atoi(e) : 0;
  if (const char* e = getenv("ISG_DENSE_SKIP_AE")) t.dense_skip_ae = e[0] != '0';
  if (const char* e = getenv("ISG_TOPK_PATH")) t.topk_radix = e[0] == 'r';
  if (const char* e = getenv("ISG_NMS_ROUNDS")) t.nms_rounds = atoi(e) > 0 ? atoi(e) : 0;
  if (const char* e = getenv("ISG_TOPK_SAMPLE")) t.topk_cluster_sample = e[0] == 'c';
  if (const char* e = getenv("ISG_TOPK_SELECT")) t.topk_cluster_select = e[0] == 'c';
  return t;
}
static std::mutex g_tuning_mutex;
static std::atomic<const Tuning*> g_tuning{nullptr};
const Tuning& tuning() {
  const Tuning* t = g_tuning.load(std::memory_order_acquire);
  if (t) return *t;
  std::lock_guard<std::mutex> lock(g_tuning_mutex);
  t = g_tuning.load(std::memory_order_acquire);
  if (!t) { t = new Tuning(read_tuning()); g_tuning.store(t, std::memory_order_release); }
  return *t;
}
}  // namespace isg

extern "C" void isg_debug_reload_tuning(void) {
  std::lock_guard<std::mutex> lock(isg::g_tuning_mutex);
  isg::g_tuning.store(new isg::Tuning(isg::read_tuning()), std::memory_order_release);   // the old record is leaked on purpose (readers may hold it)
}

extern "C" int isg_abi_version(void) { return ISG_ABI_VERSION; }

extern "C" const char* isg_strerror(int code) {
  switch (code) {
    case ISG_OK: return "ok";
    case ISG_EINVAL: return "invalid argument";
    case ISG_EWORKSPACE: return "workspace too small or misaligned";
    case ISG_EUNSUPPORTED: return "size outside the supported range";
    case ISG_ENOTCONVERGED: return "k-means did not converge within max_iter";
    default: break;
  }
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "unknown error";
}

extern "C" int isg_device_supported(int device) {
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return 0;
  return prop.major == 10 ? 1 : 0;
}

extern "C" int isg_host_device_pointer(const void* host_ptr, void** device_ptr) {
  if (!host_ptr || !device_ptr) return ISG_EINVAL;
  void* d = nullptr;
  const cudaError_t e = cudaHostGetDevicePointer(&d, const_cast<void*>(host_ptr), 0);
  if (e != cudaSuccess) { cudaGetLastError(); return (int)e; }
  *device_ptr = d;
  return ISG_OK;
}
