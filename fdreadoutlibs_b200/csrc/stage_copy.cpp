// Copy of one payload (frame / superchunk) into the pinned staging slot of the streaming path (swtpg_submit).
// The destination is written once by the CPU and read once by the GPU's copy engine, so it should neither be pulled into the
// cache first (read-for-ownership) nor stay there: non-temporal stores. Plain C++ (g++), because nvcc's host front end does
// not accept the AVX intrinsics headers.
#include <cstddef>
#include <cstdint>
#include <cstring>

#if defined(__x86_64__)
#include <immintrin.h>

__attribute__((target("avx2"))) static void
stage_copy_avx2(char* d, const char* s, size_t bytes)
{
  size_t i = 0;
  for (; i + 128 <= bytes; i += 128) {
    const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + i));
    const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + i + 32));
    const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + i + 64));
    const __m256i e = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + i + 96));
    _mm256_stream_si256(reinterpret_cast<__m256i*>(d + i), a);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(d + i + 32), b);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(d + i + 64), c);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(d + i + 96), e);
  }
  for (; i + 32 <= bytes; i += 32)
    _mm256_stream_si256(reinterpret_cast<__m256i*>(d + i), _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + i)));
  if (i < bytes)
    memcpy(d + i, s + i, bytes - i);
  _mm_sfence(); // the slot may be handed to the copy engine by another thread right after
}
#endif

extern "C" __attribute__((visibility("hidden"))) void
swtpg_stage_copy(void* dst, const void* src, size_t bytes)
{
#if defined(__x86_64__)
  static const bool avx2 = __builtin_cpu_supports("avx2");
  if (avx2 && (reinterpret_cast<uintptr_t>(dst) & 31u) == 0) {
    stage_copy_avx2(static_cast<char*>(dst), static_cast<const char*>(src), bytes);
    return;
  }
#endif
  memcpy(dst, src, bytes);
}
