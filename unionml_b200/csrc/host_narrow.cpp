// Host-side float64 -> float32 narrowing of one contiguous run, used by the gather threads of predict_host when a
// float64 frame is exactly representable in fp32 (integer / pixel domains): the chunk then crosses PCIe at half the bytes.
// Every value is checked (float -> double round trip, NaN counts as lossy); the caller re-sends a lossy chunk as float64.
//
// Plain C++ (g++ through nvcc -c): AVX2 body selected at run time, scalar otherwise.  Streaming stores: the destination
// is a pinned bounce buffer that the DMA engine reads next, the CPU never does.
#include <immintrin.h>
#include <stddef.h>
#include <stdint.h>

namespace {

int narrow_scalar(const double* src, float* dst, size_t n) {
  int lossy = 0;
  for (size_t k = 0; k < n; ++k) {
    const float f = static_cast<float>(src[k]);
    dst[k] = f;
    lossy |= static_cast<double>(f) != src[k];
  }
  return lossy;
}

__attribute__((target("avx2"))) int narrow_avx2(const double* src, float* dst, size_t n) {
  int lossy = 0;
  size_t k = 0;
  // head: up to the first 32-byte boundary of the destination
  while (k < n && (reinterpret_cast<uintptr_t>(dst + k) & 31u)) {
    const float f = static_cast<float>(src[k]);
    dst[k] = f;
    lossy |= static_cast<double>(f) != src[k];
    ++k;
  }
  __m256d bad0 = _mm256_setzero_pd(), bad1 = _mm256_setzero_pd();
  for (; k + 16 <= n; k += 16) {
    const __m256d a0 = _mm256_loadu_pd(src + k), a1 = _mm256_loadu_pd(src + k + 4);
    const __m256d a2 = _mm256_loadu_pd(src + k + 8), a3 = _mm256_loadu_pd(src + k + 12);
    const __m128 f0 = _mm256_cvtpd_ps(a0), f1 = _mm256_cvtpd_ps(a1), f2 = _mm256_cvtpd_ps(a2), f3 = _mm256_cvtpd_ps(a3);
    bad0 = _mm256_or_pd(bad0, _mm256_cmp_pd(_mm256_cvtps_pd(f0), a0, _CMP_NEQ_UQ));
    bad1 = _mm256_or_pd(bad1, _mm256_cmp_pd(_mm256_cvtps_pd(f1), a1, _CMP_NEQ_UQ));
    bad0 = _mm256_or_pd(bad0, _mm256_cmp_pd(_mm256_cvtps_pd(f2), a2, _CMP_NEQ_UQ));
    bad1 = _mm256_or_pd(bad1, _mm256_cmp_pd(_mm256_cvtps_pd(f3), a3, _CMP_NEQ_UQ));
    _mm256_stream_ps(dst + k, _mm256_set_m128(f1, f0));
    _mm256_stream_ps(dst + k + 8, _mm256_set_m128(f3, f2));
  }
  lossy |= _mm256_movemask_pd(_mm256_or_pd(bad0, bad1)) != 0;
  for (; k < n; ++k) {
    const float f = static_cast<float>(src[k]);
    dst[k] = f;
    lossy |= static_cast<double>(f) != src[k];
  }
  _mm_sfence();
  return lossy;
}

}  // namespace

namespace uml {

// dst[k] = (float)src[k] for k < n; returns 1 when some value does not survive the round trip (or is NaN)
int narrow_f64_to_f32(const double* src, float* dst, size_t n) {
  static const bool have_avx2 = __builtin_cpu_supports("avx2");
  return have_avx2 ? narrow_avx2(src, dst, n) : narrow_scalar(src, dst, n);
}

}  // namespace uml
