// launch_v1.cu — the generic one-configuration-per-thread kernels (rnea.cuh, aba.cuh, crba.cuh, rnea_derivatives.cuh): all
// per-configuration state in thread-local arrays, the model in shared memory.  They fit ANY model the engine accepts (up to
// MAXJ joints, any depth, any branching), so they are the last resort of every launch_* when a model does not fit the on-chip
// layouts of the tuned kernels; BRBD_*_V=v1 forces them (tests/test_gpu_parity.py::test_forced_paths).
#include "host_ctx.hpp"
#include "rnea.cuh"
#include "aba.cuh"
#include "crba.cuh"
#include "rnea_derivatives.cuh"

namespace brbd
{
template<class T>
brbd_status launch_rnea_v1(brbd_pool * p, DeviceCtx & d, const T * q, int64_t ldq, const T * v, int64_t ldv, const T * a,
                           int64_t lda, T * tau, int64_t ldtau, int64_t B)
{
  const ModelPOD<double> & M = p->model.pd;
  const size_t per_warp = (size_t)32 * ((M.nq | 1) + 2 * (M.nv | 1)) * sizeof(T);
  const Geometry g = pick_geometry(d, per_warp, sizeof(ModelPOD<T>), B, 16, 16);
  brbd_status st = set_smem(rnea_kernel<T>, g.dyn_bytes);
  if (st != BRBD_OK) return st;
  rnea_kernel<T><<<g.grid, g.warps_per_cta * 32, g.dyn_bytes, d.s()>>>(dev_model<T>(d), q, ldq, v, ldv, a, lda, tau, ldtau, B);
  p->launches += 1;
  CUDA_TRY(cudaGetLastError());
  return BRBD_OK;
}

template<class T>
brbd_status launch_aba_v1(brbd_pool * p, DeviceCtx & d, const T * q, int64_t ldq, const T * v, int64_t ldv, const T * tau,
                          int64_t ldtau, T * a, int64_t lda, int64_t B)
{
  const ModelPOD<double> & M = p->model.pd;
  const size_t per_warp = (size_t)32 * ((M.nq | 1) + 2 * (M.nv | 1)) * sizeof(T);
  const Geometry g = pick_geometry(d, per_warp, sizeof(ModelPOD<T>), B, 16, 16);
  brbd_status st = set_smem(aba_kernel<T>, g.dyn_bytes);
  if (st != BRBD_OK) return st;
  aba_kernel<T><<<g.grid, g.warps_per_cta * 32, g.dyn_bytes, d.s()>>>(dev_model<T>(d), q, ldq, v, ldv, tau, ldtau, a, lda, B);
  p->launches += 1;
  CUDA_TRY(cudaGetLastError());
  return BRBD_OK;
}

template<class T>
brbd_status launch_crba_v1(brbd_pool * p, DeviceCtx & d, const T * q, int64_t ldq, T * Mout, int64_t ldM, int64_t B)
{
  const ModelPOD<double> & M = p->model.pd;
  const size_t per_warp = (size_t)32 * ((M.nq | 1) + (M.nv | 1)) * sizeof(T);
  const Geometry g = pick_geometry(d, per_warp, sizeof(ModelPOD<T>), B, 16, 16);
  brbd_status st = set_smem(crba_kernel<T>, g.dyn_bytes);
  if (st != BRBD_OK) return st;
  crba_kernel<T><<<g.grid, g.warps_per_cta * 32, g.dyn_bytes, d.s()>>>(dev_model<T>(d), q, ldq, Mout, ldM, B);
  p->launches += 1;
  CUDA_TRY(cudaGetLastError());
  return BRBD_OK;
}

template<class T>
brbd_status launch_rnea_derivs_v1(brbd_pool * p, DeviceCtx & d, const T * q, int64_t ldq, const T * v, int64_t ldv,
                                  const T * a, int64_t lda, T * dq, int64_t ld_dq, T * dv, int64_t ld_dv, T * da,
                                  int64_t ld_da, T * tau, int64_t ldtau, int64_t B)
{
  const ModelPOD<double> & M = p->model.pd;
  const size_t per_warp = (size_t)32 * ((M.nq | 1) + 5 * (M.nv | 1)) * sizeof(T);
  const Geometry g = pick_geometry(d, per_warp, sizeof(ModelPOD<T>), B, 16, 16);
  brbd_status st = set_smem(rnea_derivatives_kernel<T>, g.dyn_bytes);
  if (st != BRBD_OK) return st;
  rnea_derivatives_kernel<T><<<g.grid, g.warps_per_cta * 32, g.dyn_bytes, d.s()>>>(
    dev_model<T>(d), q, ldq, v, ldv, a, lda, dq, ld_dq, dv, ld_dv, da, ld_da, tau, ldtau, B);
  p->launches += 1;
  CUDA_TRY(cudaGetLastError());
  return BRBD_OK;
}

#define BRBD_INST(T)                                                                                                              \
  template brbd_status launch_rnea_v1<T>(brbd_pool *, DeviceCtx &, const T *, int64_t, const T *, int64_t, const T *, int64_t,   \
                                         T *, int64_t, int64_t);                                                                  \
  template brbd_status launch_aba_v1<T>(brbd_pool *, DeviceCtx &, const T *, int64_t, const T *, int64_t, const T *, int64_t,    \
                                        T *, int64_t, int64_t);                                                                   \
  template brbd_status launch_crba_v1<T>(brbd_pool *, DeviceCtx &, const T *, int64_t, T *, int64_t, int64_t);                    \
  template brbd_status launch_rnea_derivs_v1<T>(brbd_pool *, DeviceCtx &, const T *, int64_t, const T *, int64_t, const T *,     \
                                                int64_t, T *, int64_t, T *, int64_t, T *, int64_t, T *, int64_t, int64_t);
BRBD_INST(double)
BRBD_INST(float)
#undef BRBD_INST
} // namespace brbd
