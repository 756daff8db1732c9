// Host-side fp32 -> fp16 cast of the query batch (search_on_device casts on the host too,
// fast_plaid.py:241).  Done here, single-threaded with F16C, instead of through a torch CPU op: ATen
// parallelises even this 1 MB cast over its intra-op pool, and on a box whose cgroup grants fewer
// cores than `nproc` reports the pool's wake-up costs milliseconds at random (measured: 7 -> 18 ms
// outliers of the host-buffer search path).  Round-to-nearest-even, same values as ATen's cast.
#include <immintrin.h>
#include <stdint.h>
#include <string.h>

#include "common.cuh"

namespace {

uint16_t f32_to_f16_rn(float f) {  // portable fallback
  uint32_t x;
  memcpy(&x, &f, 4);
  const uint32_t sign = (x >> 16) & 0x8000u;
  x &= 0x7fffffffu;
  if (x >= 0x7f800000u) return uint16_t(sign | (x > 0x7f800000u ? 0x7e00u : 0x7c00u));  // NaN / inf
  if (x >= 0x477ff000u) return uint16_t(sign | 0x7c00u);                                  // rounds to inf
  if (x < 0x33000001u) return uint16_t(sign);                                             // rounds to zero
  if (x < 0x38800000u) {                                                                   // fp16 subnormal
    const int shift = 126 - int(x >> 23);                                                  // 14 .. 24
    const uint32_t m = (x & 0x7fffffu) | 0x800000u;
    uint32_t h = m >> shift;
    const uint32_t rem = m & ((1u << shift) - 1u), half = 1u << (shift - 1);
    if (rem > half || (rem == half && (h & 1u))) ++h;
    return uint16_t(sign | h);
  }
  uint32_t h = ((x >> 23) - 112u) << 10 | ((x >> 13) & 0x3ffu);
  const uint32_t rem = x & 0x1fffu;
  if (rem > 0x1000u || (rem == 0x1000u && (h & 1u))) ++h;  // a carry into the exponent is correct
  return uint16_t(sign | h);
}

__attribute__((target("avx,f16c"))) void cast_f16c(const float* src, uint16_t* dst, size_t n) {
  size_t i = 0;
  for (; i + 8 <= n; i += 8) {
    const __m256 v = _mm256_loadu_ps(src + i);
    _mm_storeu_si128(reinterpret_cast<__m128i*>(dst + i), _mm256_cvtps_ph(v, _MM_FROUND_TO_NEAREST_INT | _MM_FROUND_NO_EXC));
  }
  for (; i < n; ++i) dst[i] = f32_to_f16_rn(src[i]);
}

}  // namespace

extern "C" int fpb_cast_f32_to_f16_host(const float* h_src, void* h_dst, size_t n) {
  if ((!h_src || !h_dst) && n > 0) {
    fpb_set_error("fpb_cast_f32_to_f16_host: NULL pointer");
    return FPB_ERR_INVALID;
  }
  uint16_t* dst = static_cast<uint16_t*>(h_dst);
  static const bool has_f16c = __builtin_cpu_supports("f16c") && __builtin_cpu_supports("avx");
  if (has_f16c) {
    cast_f16c(h_src, dst, n);
  } else {
    for (size_t i = 0; i < n; ++i) dst[i] = f32_to_f16_rn(h_src[i]);
  }
  return FPB_OK;
}

// exposed for the tests: the portable path regardless of the CPU
extern "C" int fpb_cast_f32_to_f16_host_portable(const float* h_src, void* h_dst, size_t n) {
  uint16_t* dst = static_cast<uint16_t*>(h_dst);
  for (size_t i = 0; i < n; ++i) dst[i] = f32_to_f16_rn(h_src[i]);
  return FPB_OK;
}
