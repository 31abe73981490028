// launch_integrate.cu — launch of integrate / the semi-implicit Euler step, and the DFMA peak measurement
#include "host_ctx.hpp"
#include "integrate.cuh"

namespace brbd
{
// integrate / Euler step: one configuration per thread, warp tiles staged through shared memory
template<class T, bool EULER>
brbd_status launch_integrate(brbd_pool * p, DeviceCtx & d, const T * q, int64_t ldq, const T * v, int64_t ldv, const T * a,
                             int64_t lda, T dt, T * qout, int64_t ldqo, T * vout, int64_t ldvo, int64_t B)
{
  const ModelPOD<double> & M = p->model.pd;
  const size_t per_warp = (size_t)32 * ((M.nq | 1) + 2 * (M.nv | 1)) * sizeof(T);
  const Geometry g = pick_geometry(d, per_warp, sizeof(ModelPOD<T>), B, 8, 16);
  brbd_status st = set_smem(integrate_kernel<T, EULER>, g.dyn_bytes);
  if (st != BRBD_OK) return st;
  integrate_kernel<T, EULER><<<g.grid, g.warps_per_cta * 32, g.dyn_bytes, d.s()>>>(dev_model<T>(d), q, ldq, v, ldv, a, lda, dt, qout,
                                                                                    ldqo, vout, ldvo, B);
  p->launches += 1;
  CUDA_TRY(cudaGetLastError());
  return BRBD_OK;
}
__global__ void fp64_peak_kernel(double * out, int iters)
{
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double b = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i)
  {
    a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
    a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
brbd_status measure_fp64_peak(brbd_pool * p, double * flops_per_s, double * elapsed_ms)
{
  if (!p || p->devs.empty()) return fail(BRBD_EINVAL, "null pool");
  DeviceCtx & d = p->devs[0];
  CUDA_TRY(cudaSetDevice(d.dev));
  const int threads = 256, blocks = d.sm_count * 8, iters = 1 << 16;
  brbd_status st = ensure_work(d, (size_t)threads * blocks * sizeof(double));
  if (st != BRBD_OK) return st;
  fp64_peak_kernel<<<blocks, threads, 0, d.stream>>>((double *)d.work, 1 << 10); // warm-up
  CUDA_TRY(cudaEventRecord(d.ev0, d.stream));
  fp64_peak_kernel<<<blocks, threads, 0, d.stream>>>((double *)d.work, iters);
  CUDA_TRY(cudaEventRecord(d.ev1, d.stream));
  CUDA_TRY(cudaStreamSynchronize(d.stream));
  p->launches += 2;
  float ms = 0.f;
  CUDA_TRY(cudaEventElapsedTime(&ms, d.ev0, d.ev1));
  const double flops = 2.0 * 8.0 * (double)iters * threads * blocks;
  if (flops_per_s) *flops_per_s = flops / (ms * 1e-3);
  if (elapsed_ms) *elapsed_ms = ms;
  return BRBD_OK;
}
template brbd_status launch_integrate<double, false>(brbd_pool *, DeviceCtx &, const double *, int64_t, const double *, int64_t, const double *, int64_t, double, double *, int64_t, double *, int64_t, int64_t);
template brbd_status launch_integrate<float, false>(brbd_pool *, DeviceCtx &, const float *, int64_t, const float *, int64_t, const float *, int64_t, float, float *, int64_t, float *, int64_t, int64_t);
template brbd_status launch_integrate<double, true>(brbd_pool *, DeviceCtx &, const double *, int64_t, const double *, int64_t, const double *, int64_t, double, double *, int64_t, double *, int64_t, int64_t);
template brbd_status launch_integrate<float, true>(brbd_pool *, DeviceCtx &, const float *, int64_t, const float *, int64_t, const float *, int64_t, float, float *, int64_t, float *, int64_t, int64_t);
} // namespace brbd
