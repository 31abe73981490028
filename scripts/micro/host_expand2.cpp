// Host-side expansion, second look: T host threads rebuild B dense nv x nv matrices (zeros + E packed entries each) in the
// caller's memory.  Variants: 0 = memset + scatter in place (host_expand.cpp), 1 = assemble the matrix in a thread-local buffer
// (L1 resident) and memcpy it out, 2 = the same with non-temporal stores (no read-for-ownership of the destination lines).
//   g++ -O3 -march=native -fopenmp scripts/micro/host_expand2.cpp -o scripts/micro/host_expand2
#include <immintrin.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <omp.h>
static inline void stream_copy(double * d, const double * s, long n)
{
  long k = 0;
  while (k < n && ((uintptr_t)(d + k) & 31)) { _mm_stream_si64((long long *)(d + k), *(const long long *)(s + k)); ++k; }
  for (; k + 4 <= n; k += 4) _mm256_stream_pd(d + k, _mm256_loadu_pd(s + k));
  for (; k < n; ++k) _mm_stream_si64((long long *)(d + k), *(const long long *)(s + k));
}
int main(int argc, char ** argv)
{
  const long B = argc > 1 ? atol(argv[1]) : 65536;
  const int nv = 35, nn = nv * nv, E = 356;
  std::vector<int> idx(E);
  for (int k = 0; k < E; ++k) idx[k] = (int)((long)k * nn / E);
  double * dst = (double *)aligned_alloc(4096, (size_t)B * nn * 8);
  double * src = (double *)aligned_alloc(4096, (size_t)B * E * 8);
  memset(dst, 1, (size_t)B * nn * 8);
  for (long k = 0; k < B * E; ++k) src[k] = (double)k;
  for (int variant = 0; variant < 3; ++variant)
    for (int T = 1; T <= omp_get_max_threads(); T *= 2)
    {
      double best = 1e9;
      for (int rep = 0; rep < 5; ++rep)
      {
        auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel num_threads(T)
        {
          alignas(64) double buf[nn + 8];
          memset(buf, 0, sizeof buf);
#pragma omp for schedule(static)
          for (long c = 0; c < B; ++c)
          {
            double * d = dst + c * nn;
            const double * s = src + c * E;
            if (variant == 0)
            {
              memset(d, 0, nn * 8);
              for (int k = 0; k < E; ++k) d[idx[k]] = s[k];
            }
            else
            {
              for (int k = 0; k < E; ++k) buf[idx[k]] = s[k]; // the pattern is the same for every matrix: the zeros stay
              if (variant == 1) memcpy(d, buf, nn * 8);
              else stream_copy(d, buf, nn);
            }
          }
          if (variant == 2) _mm_sfence();
        }
        double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (dt < best) best = dt;
      }
      printf("variant %d threads %2d: %.2f ms for %ld matrices -> %.1f GB/s of dense output\n", variant, T, best * 1e3, B, (double)B * nn * 8 / best / 1e9);
    }
  return 0;
}
