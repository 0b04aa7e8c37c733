// f16_convert.cpp -- float32 -> float16 (IEEE round-to-nearest-even) conversion of feature rows on the HOST, for the
// extraction job's reader (include/xvec_job.h, option feats_f16).  The first thing the device does with a feature is round
// it to fp16 (pack_im2col_kernel: __floats2half2_rn), so converting while the payload is copied from the page cache into the
// page-locked batch buffer gives bit-identical x-vectors and halves the bytes written to that buffer, read back by the DMA
// engine and sent over PCIe -- what bounds a multi-GPU job that is fed from one host's memory.
// Plain C++ (compiled by the host compiler, not cudafe): F16C + AVX2 when the CPU has them (runtime check), else the scalar
// routine below, which is exact too.
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#define XV_HAVE_X86 1
#else
#define XV_HAVE_X86 0
#endif

namespace {

inline uint16_t f32_to_f16_rn(float f) {
  uint32_t x;
  memcpy(&x, &f, 4);
  const uint32_t sign = (x >> 16) & 0x8000u;
  x &= 0x7fffffffu;
  if (x >= 0x7f800000u) return uint16_t(sign | 0x7c00u | (x > 0x7f800000u ? 0x0200u | ((x >> 13) & 0x3ffu) : 0u));   // inf / nan
  if (x >= 0x477ff000u) return uint16_t(sign | 0x7c00u);                       // rounds to a magnitude >= 65520: inf
  if (x < 0x33000001u) return uint16_t(sign);                                  // below half of the smallest subnormal: +-0
  if (x < 0x38800000u) {                                                       // subnormal half
    const int shift = 126 - int(x >> 23);                                      // 14 .. 24 bits to drop from the 24-bit significand
    const uint32_t mant = (x & 0x7fffffu) | 0x800000u;
    uint32_t h = mant >> shift;
    const uint32_t rem = mant & ((1u << shift) - 1u), half = 1u << (shift - 1);
    if (rem > half || (rem == half && (h & 1u))) ++h;
    return uint16_t(sign | h);
  }
  uint32_t h = ((x - 0x38000000u) >> 13);                                      // rebias exponent 127 -> 15, keep 10 mantissa bits
  const uint32_t rem = x & 0x1fffu;
  if (rem > 0x1000u || (rem == 0x1000u && (h & 1u))) ++h;                      // a carry into the exponent is the right answer
  return uint16_t(sign | h);
}

void convert_scalar(const float* src, uint16_t* dst, size_t n) {
  for (size_t i = 0; i < n; ++i) dst[i] = f32_to_f16_rn(src[i]);
}

#if XV_HAVE_X86
__attribute__((target("avx2,f16c"))) void convert_f16c(const float* src, uint16_t* dst, size_t n) {
  size_t i = 0;
  // head: up to the first 16-byte aligned destination
  while (i < n && (reinterpret_cast<uintptr_t>(dst + i) & 15u) != 0) { dst[i] = f32_to_f16_rn(src[i]); ++i; }
  // body: unaligned 32-byte loads, aligned 16-byte NON-TEMPORAL stores (the buffer is only read by the DMA engine: no
  // read-for-ownership, no cache pollution)
  for (; i + 8 <= n; i += 8) {
    const __m128i h = _mm256_cvtps_ph(_mm256_loadu_ps(src + i), _MM_FROUND_TO_NEAREST_INT | _MM_FROUND_NO_EXC);
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i), h);
  }
  for (; i < n; ++i) dst[i] = f32_to_f16_rn(src[i]);
  _mm_sfence();
}
#endif

}  // namespace

// n float32 values at src (any alignment of 4) -> n float16 bit patterns at dst (any alignment of 2).
extern "C" void xv_convert_f32_to_f16_host(const float* src, uint16_t* dst, size_t n) {
#if XV_HAVE_X86
  static const bool fast = __builtin_cpu_supports("f16c") && __builtin_cpu_supports("avx2");
  if (fast) { convert_f16c(src, dst, n); return; }
#endif
  convert_scalar(src, dst, n);
}

// The scalar routine alone (tests compare both against numpy's float16).
extern "C" void xv_convert_f32_to_f16_host_scalar(const float* src, uint16_t* dst, size_t n) { convert_scalar(src, dst, n); }
