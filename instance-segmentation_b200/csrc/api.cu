// Library-level entry points of libisg.so.
#include "common.cuh"

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
