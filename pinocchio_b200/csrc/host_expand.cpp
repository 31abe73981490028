// host_expand.cpp — host side of a host-pointer crba call with packed transfer (capi.cu: run_crba_expand): the device sends only
// the entries inside the structural pattern of M (a third of nv * nv for a humanoid) and `threads` host threads — the
// num_threads the reference's parallel API hands to its callee (parallel/rnea.hpp:38) — rebuild the caller's dense
// column-major matrices.  Every thread assembles a matrix in a buffer that stays in its L1 (the pattern is the same for every
// configuration: the zeros are written once) and streams it out with non-temporal stores, so that the destination lines are not
// read first.  Measured on this pool's hosts (scripts/micro/host_expand2.cpp, 16 threads): 113 GB/s of dense output against
// 56 GB/s for the DMA of the dense block and 65 GB/s with ordinary stores.
// No algorithm runs here: values are copied, bit for bit, from the device's result.
#include <cstdint>
#include <cstring>
#include <vector>
#if defined(__x86_64__)
#include <emmintrin.h>
#endif

namespace brbd
{
namespace
{
inline void stream_out(double * d, const double * s, int64_t n)
{
#if defined(__x86_64__)
  int64_t k = 0;
  if (k < n && (reinterpret_cast<uintptr_t>(d) & 15)) { _mm_stream_si64(reinterpret_cast<long long *>(d), *reinterpret_cast<const long long *>(s)); ++k; }
  for (; k + 2 <= n; k += 2) _mm_stream_pd(d + k, _mm_loadu_pd(s + k));
  if (k < n) _mm_stream_si64(reinterpret_cast<long long *>(d + k), *reinterpret_cast<const long long *>(s + k));
#else
  std::memcpy(d, s, (size_t)n * sizeof(double));
#endif
}
inline void stream_out(float * d, const float * s, int64_t n) { std::memcpy(d, s, (size_t)n * sizeof(float)); }
} // namespace

// dst: the caller's block (configuration c at dst + c * ld, nn = nv * nv elements each); src: packed, nnz per configuration;
// idx[k] = position of packed entry k inside a matrix
template<class T>
void expand_packed(T * dst, int64_t ld, const T * src, int64_t nnz, const int32_t * idx, int nn, int64_t count, int threads, int64_t src_ld)
{
  if (src_ld <= 0) src_ld = nnz;
  if (threads < 1) threads = 1;
#pragma omp parallel num_threads(threads)
  {
    std::vector<T> buf((size_t)nn + 8, T(0));
#pragma omp for schedule(static)
    for (int64_t c = 0; c < count; ++c)
    {
      const T * s = src + c * src_ld;
      for (int64_t k = 0; k < nnz; ++k) buf[idx[k]] = s[k];
      stream_out(dst + c * ld, buf.data(), nn);
    }
#if defined(__x86_64__)
    _mm_sfence();
#endif
  }
}
template void expand_packed<double>(double *, int64_t, const double *, int64_t, const int32_t *, int, int64_t, int, int64_t);
template void expand_packed<float>(float *, int64_t, const float *, int64_t, const int32_t *, int, int64_t, int, int64_t);
} // namespace brbd
