// Host-side expansion micro-benchmark: how fast can T host threads rebuild B dense nv x nv matrices (zeros + E packed
// entries each) in the caller's memory?  Decides whether a packed D2H of CRBA's M (PCIe bytes / 3) can beat the plain
// 55 GB/s DMA of the dense matrix.   g++ -O3 -march=native -fopenmp host_expand.cpp -o host_expand
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <omp.h>
int main(int argc, char ** argv)
{
  const long B = argc > 1 ? atol(argv[1]) : 65536;
  const int nv = 35, nn = nv * nv, E = 391; // simple_humanoid: 356 + 35 structural non-zeros
  std::vector<int> idx(E);
  for (int k = 0; k < E; ++k) idx[k] = (int)((long)k * nn / E);
  double * dst = (double *)aligned_alloc(4096, (size_t)B * nn * 8);
  double * src = (double *)aligned_alloc(4096, (size_t)B * E * 8);
  memset(dst, 1, (size_t)B * nn * 8);
  for (long k = 0; k < B * E; ++k) src[k] = (double)k;
  for (int T = 1; T <= omp_get_max_threads(); T *= 2)
  {
    double best = 1e9;
    for (int rep = 0; rep < 5; ++rep)
    {
      auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for num_threads(T) schedule(static)
      for (long c = 0; c < B; ++c)
      {
        double * d = dst + c * nn;
        const double * s = src + c * E;
        memset(d, 0, nn * 8);
        for (int k = 0; k < E; ++k) d[idx[k]] = s[k];
      }
      double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      if (dt < best) best = dt;
    }
    printf("threads %2d: %.2f ms for %ld matrices -> %.1f GB/s of dense output (DMA of the dense block: ~55 GB/s)\n", T, best * 1e3, B,
           (double)B * nn * 8 / best / 1e9);
  }
  return 0;
}
