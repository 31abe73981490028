// Micro-benchmark: store throughput of an SM for the CRBA output pattern — every lane owns one "configuration" whose
// matrix is `stride` bytes of global memory, and repeatedly writes a run of `S` bytes of it —
//   mode 0: one cp.async.bulk.global.shared::cta (1-D TMA bulk copy) of S bytes per lane (source: the lane's shared-memory row)
//   mode 1: the warp writes the 32 runs cooperatively with coalesced 8-byte STG (lane l writes element l, l + 32, ... of a run)
//   mode 2: as 1 with 16-byte STG (runs 16-byte aligned)
// over a persistent grid of 148 CTAs x W warps walking B configurations.  Reports GB/s and bytes / clock / SM.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/micro/bulk_store_bw scripts/micro/bulk_store_bw.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(1024, 1) store_kernel(double * out, long long stride_e, int S_e, int B, int mode, int runs_per_cfg, int depth)
{
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double * smem = reinterpret_cast<double *>(smem_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
  const int pitch = (S_e + 1) & ~1; // 16-byte aligned rows
  double * tile = smem + (size_t)warp * 32 * pitch;
  for (int k = lane; k < 32 * pitch; k += 32) tile[k] = (double)k;
  __syncwarp();
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  const long long per_round = (long long)gridDim.x * warps * 32;
  for (long long cfg0 = ((long long)blockIdx.x * warps + warp) * 32; cfg0 < B; cfg0 += per_round)
  {
    for (int r = 0; r < runs_per_cfg; ++r)
    {
      if (mode == 0)
      {
        double * dst = out + (cfg0 + lane) * stride_e + (long long)r * S_e;
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"((unsigned)__cvta_generic_to_shared(tile + lane * pitch)),
                     "r"(S_e * 8)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        if (depth == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        else if (depth == 1) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        else if (depth == 2) asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
        else asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory");
      }
      else if (mode == 1)
      {
        for (int c = 0; c < 32; ++c)
        {
          double * dst = out + (cfg0 + c) * stride_e + (long long)r * S_e;
          for (int e = lane; e < S_e; e += 32) dst[e] = tile[c * pitch + e];
        }
      }
      else
      {
        for (int c = 0; c < 32; ++c)
        {
          double2 * dst = reinterpret_cast<double2 *>(out + (cfg0 + c) * stride_e + (long long)r * S_e);
          const double2 * src = reinterpret_cast<const double2 *>(tile + c * pitch);
          for (int e = lane; e < S_e / 2; e += 32) dst[e] = src[e];
        }
      }
    }
  }
  if (mode == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main(int argc, char ** argv)
{
  const int B = argc > 1 ? atoi(argv[1]) : 262144;
  const long long stride_e = argc > 2 ? atoll(argv[2]) : 1444; // talos: 38 x 38 doubles, 16-byte aligned matrices
  double * out = nullptr;
  cudaMalloc(&out, (size_t)B * stride_e * 8 + 4096);
  cudaMemset(out, 0, (size_t)B * stride_e * 8);
  cudaFuncSetAttribute(store_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  printf("B = %d configurations, %lld doubles apart; 148 CTAs\n", B, stride_e);
  const int sizes[] = {38, 76, 114, 152, 228, 304, 722, 1444};
  for (int mode = 0; mode < 3; ++mode)
    for (int W : {4, 8, 16})
      for (int depth : {0, 2})
      {
        if (mode != 0 && depth != 0) continue;
        for (int S_e : sizes)
        {
          if (S_e > stride_e) continue;
          const int pitch = (S_e + 1) & ~1;
          const size_t smem = (size_t)W * 32 * pitch * 8;
          if (smem > 220 * 1024) continue;
          const int runs = (int)(stride_e / S_e);
          store_kernel<<<148, W * 32, smem>>>(out, stride_e, S_e, B, mode, runs, depth);
          cudaEventRecord(e0);
          store_kernel<<<148, W * 32, smem>>>(out, stride_e, S_e, B, mode, runs, depth);
          cudaEventRecord(e1);
          cudaError_t err = cudaDeviceSynchronize();
          if (err != cudaSuccess) { printf("mode %d W %d S %d: %s\n", mode, W, S_e, cudaGetErrorString(err)); return 1; }
          float ms = 0;
          cudaEventElapsedTime(&ms, e0, e1);
          const double bytes = (double)B * runs * S_e * 8;
          printf("mode %d (%s) warps %2d depth %d run %5d B: %8.3f ms  %7.1f GB/s  %5.1f B/clk/SM\n", mode,
                 mode == 0 ? "bulk copy per lane" : (mode == 1 ? "coalesced STG.64" : "coalesced STG.128"), W, depth, S_e * 8, ms, bytes / ms / 1e6,
                 bytes / (ms * 1e-3) / 1.965e9 / 148);
        }
      }
  return 0;
}
