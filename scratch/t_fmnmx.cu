__global__ void k(const float* a, float* o, float thr) {
  float x = a[threadIdx.x], y = a[threadIdx.x+32], z = a[threadIdx.x+64];
  float m = fmaxf(fmaxf(x, y), z);
  float v = (x < thr) ? 0.f : x;
  o[threadIdx.x] = m + v;
}
