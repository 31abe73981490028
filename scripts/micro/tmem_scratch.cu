// Micro-test: TMEM (tcgen05.alloc/st/ld) as per-thread scratch for FP64 state.
// Each of 128 threads (4 warps) owns one TMEM lane; a double occupies two 32-bit columns.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void tmem_st_f64(uint32_t taddr, double v)
{
  const uint32_t lo = (uint32_t)__double2loint(v), hi = (uint32_t)__double2hiint(v);
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};\n" ::"r"(taddr), "r"(lo), "r"(hi) : "memory");
}
__device__ __forceinline__ double tmem_ld_f64(uint32_t taddr)
{
  uint32_t lo, hi;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];\n" : "=r"(lo), "=r"(hi) : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
  return __hiloint2double((int)hi, (int)lo);
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }

__global__ void __launch_bounds__(128, 1) tmem_test(double * out, int * bad, int iters, long long * cycles)
{
  __shared__ uint32_t tbase_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0)
  {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"((uint32_t)__cvta_generic_to_shared(&tbase_s)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n");
  const uint32_t tbase = tbase_s;
  // lane field = bits 31..16; a warp may only touch the 32 lanes of its own sub-partition (warp % 4)
  const uint32_t mine = tbase + ((uint32_t)(warp & 3) * 32u << 16);
  int nbad = 0;
  // write 256 doubles (512 columns) with a dynamic (loop) column index, read them back in reverse
  for (int k = 0; k < 256; ++k) tmem_st_f64(mine + 2 * k, 1000.0 * blockIdx.x + tid + 1e-3 * k);
  tmem_wait_st();
  for (int k = 255; k >= 0; --k)
  {
    const double v = tmem_ld_f64(mine + 2 * k);
    if (v != 1000.0 * blockIdx.x + tid + 1e-3 * k) ++nbad;
  }
  // read-modify-write chain (accumulate) to time latency / check ordering st -> ld on the same cell
  long long t0 = clock64();
  double acc = 0;
  for (int it = 0; it < iters; ++it)
  {
    const int k = (it * 7) & 255;
    double v = tmem_ld_f64(mine + 2 * k);
    v += 1.0;
    tmem_st_f64(mine + 2 * k, v);
    tmem_wait_st();
    acc += v;
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + tid] = acc;
  if (nbad) atomicAdd(bad, nbad);
  if (tid == 0 && blockIdx.x == 0) *cycles = t1 - t0;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tbase), "n"(512));
}

int main()
{
  double * out; int * bad; long long * cyc;
  cudaMalloc(&out, 148 * 128 * 8); cudaMalloc(&bad, 4); cudaMalloc(&cyc, 8);
  cudaMemset(bad, 0, 4);
  const int iters = 4096;
  tmem_test<<<148, 128>>>(out, bad, iters, cyc);
  cudaError_t e = cudaDeviceSynchronize();
  int hbad = -1; long long hc = 0; double h0[128];
  cudaMemcpy(&hbad, bad, 4, cudaMemcpyDeviceToHost); cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(h0, out, sizeof(h0), cudaMemcpyDeviceToHost);
  // expected acc: sum over iterations of the incremented value
  double exp0 = 0; { double cell[256]; for (int k = 0; k < 256; ++k) cell[k] = 0 + 1e-3 * k; for (int it = 0; it < iters; ++it) { int k = (it * 7) & 255; cell[k] += 1.0; exp0 += cell[k]; } }
  printf("status=%s mismatches=%d rmw_cycles_per_iter=%.1f acc[0]=%.6f expected=%.6f %s\n", cudaGetErrorString(e), hbad, (double)hc / iters,
         h0[0], exp0, (h0[0] == exp0) ? "OK" : "DIFF");
  return (e != cudaSuccess || hbad != 0 || h0[0] != exp0) ? 1 : 0;
}
