// exhaustive check: sqrt_rn_normal == sqrtf and rcp_rn_normal == __frcp_rn on the ranges used
#include <cstdio>
#include <cuda_fp16.h>
__device__ __forceinline__ float sqrt_rn_normal(float x) {
  float y, s, h, r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  asm("mul.rn.ftz.f32 %0, %1, %2;" : "=f"(s) : "f"(x), "f"(y));
  asm("mul.rn.ftz.f32 %0, %1, 0f3F000000;" : "=f"(h) : "f"(y));
  r = __fmaf_rn(-s, s, x);
  return __fmaf_rn(r, h, s);
}
__device__ __forceinline__ float rcp_rn_normal(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  const float e = __fmaf_rn(x, y, -1.0f);
  return __fmaf_rn(y, -e, y);
}
__global__ void chk(unsigned lo, unsigned hi, unsigned long long* bad_s, unsigned long long* bad_r) {
  for (unsigned long long u = lo + blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; u <= hi;
       u += (unsigned long long)gridDim.x * blockDim.x) {
    float x = __uint_as_float((unsigned)u);
    if (__float_as_uint(sqrt_rn_normal(x)) != __float_as_uint(sqrtf(x))) atomicAdd(bad_s, 1ull);
  }
  // every positive normal fp16 value
  for (unsigned h = 0x0400 + blockIdx.x * blockDim.x + threadIdx.x; h < 0x7c00; h += gridDim.x * blockDim.x) {
    float x = __half2float(__ushort_as_half((unsigned short)h));
    if (__float_as_uint(rcp_rn_normal(x)) != __float_as_uint(__frcp_rn(x))) atomicAdd(bad_r, 1ull);
  }
}
int main() {
  unsigned long long *d, h[2];
  cudaMalloc(&d, 16); cudaMemset(d, 0, 16);
  // x in [2^-40, 2^40]
  chk<<<148 * 8, 256>>>(0x2b800000u, 0x53800000u, d, d + 1);
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("sqrt mismatches %llu  rcp mismatches %llu  (%s)\n", h[0], h[1], cudaGetErrorString(cudaGetLastError()));
  return (h[0] || h[1]) ? 1 : 0;
}
