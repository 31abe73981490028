// launch_aba_derivs.cu — launch of batched computeABADerivatives (aba-derivatives.hpp:52-66)
#include "host_ctx.hpp"
#include "aba_derivatives.cuh"

namespace brbd
{
template<class T>
brbd_status launch_aba_derivs(brbd_pool * p, DeviceCtx & d, const T * q, int64_t ldq, const T * v, int64_t ldv,
                              const T * tau, int64_t ldtau, T * dq, int64_t ld_dq, T * dv, int64_t ld_dv, T * dtau,
                              int64_t ld_dtau, T * ddq, int64_t ldddq, int64_t B)
{
  const ModelPOD<double> & M = p->model.pd;
  if (!std::getenv("BRBD_DABA_V") && use_generated<T>(p, BRBD_GEN_ABA_DERIVATIVES, B))
    return launch_generated_derivs<T>(p, d, BRBD_GEN_ABA_DERIVATIVES, q, ldq, v, ldv, tau, ldtau, dq, ld_dq, dv, ld_dv, dtau, ld_dtau, ddq, ldddq, B);
  // preferred: one kernel, G lanes per configuration, everything in shared memory (aba_deriv_coop.cuh)
  {
    const int G = coop_group_size(M.nv);
    const AbaCoopLayout L = aba_coop_layout(M.nq, M.nv, M.njoints, G);
    const size_t static_bytes = sizeof(ModelPOD<T>) + sizeof(CoopTables) + 1024;
    const GeometryCoop g = pick_geometry_coop(d, (size_t)L.per_group * sizeof(T), G, static_bytes, B);
    if (!forced_path("BRBD_DABA_V", "v1") && p->model.coop.nbranch <= A_MAXBRANCH && g.dyn_bytes + static_bytes <= (size_t)d.max_smem_optin + 1024)
    {
      brbd_status st = BRBD_OK;
#define BRBD_LAUNCH_COOP(GG)                                                                                     \
  {                                                                                                              \
    st = set_smem(aba_derivatives_coop_kernel<T, GG>, g.dyn_bytes);                                              \
    if (st != BRBD_OK) return st;                                                                                \
    aba_derivatives_coop_kernel<T, GG><<<g.grid, g.warps * 32, g.dyn_bytes, d.s()>>>(                            \
      dev_model<T>(d), p->model.coop, L, q, ldq, v, ldv, tau, ldtau, dq, ld_dq, dv, ld_dv, dtau, ld_dtau, ddq, ldddq, B); \
  }
      if (G == 8) BRBD_LAUNCH_COOP(8)
      else if (G == 16) BRBD_LAUNCH_COOP(16)
      else BRBD_LAUNCH_COOP(32)
#undef BRBD_LAUNCH_COOP
      p->launches += 1;
      CUDA_TRY(cudaGetLastError());
      return BRBD_OK;
    }
  }
  // fallback (very large models, more than A_MAXBRANCH branching joints): v1, one configuration per thread
  const size_t per_warp = (size_t)32 * ((M.nq | 1) + 5 * (M.nv | 1)) * sizeof(T);
  const Geometry g = pick_geometry(d, per_warp, sizeof(ModelPOD<T>), B, 16, 16);
  brbd_status st = set_smem(aba_derivatives_sweep_kernel<T>, g.dyn_bytes);
  if (st != BRBD_OK) return st;
  // thread-private workspace [entry][thread]: Minv (nv*nv) + Fcrb per tree depth ((maxdepth+1)*nv*6)
  const size_t nthreads = (size_t)g.grid * g.warps_per_cta * 32;
  const size_t ws_elems = ((size_t)M.nv * M.nv + (size_t)(M.maxdepth + 1) * M.nv * 6) * nthreads;
  st = ensure_work(d, ws_elems * sizeof(T));
  if (st != BRBD_OK) return st;
  // pass A: sweeps -> Minv into `dtau`, dtau_dq / dtau_dv into `dq` / `dv` (all in the caller's layout)
  aba_derivatives_sweep_kernel<T><<<g.grid, g.warps_per_cta * 32, g.dyn_bytes, d.s()>>>(
    dev_model<T>(d), q, ldq, v, ldv, tau, ldtau, dq, ld_dq, dv, ld_dv, dtau, ld_dtau, ddq, ldddq, (T *)d.work, B);
  p->launches += 1;
  CUDA_TRY(cudaGetLastError());
  // pass B: dq <- -Minv * dq, dv <- -Minv * dv, one warp per configuration (aba-derivatives.hxx:451-452)
  {
    const int nv = M.nv;
    const size_t per_warp_gemm = (size_t)3 * nv * (nv + 1) * sizeof(T);
    const int warps = (int)std::max<size_t>(1, std::min<size_t>(8, (size_t)d.max_smem_optin / per_warp_gemm));
    const size_t dyn = (size_t)warps * per_warp_gemm;
    st = set_smem(aba_derivatives_gemm_kernel<T>, dyn);
    if (st != BRBD_OK) return st;
    const int64_t ctas = (B + warps - 1) / warps;
    const int grid = (int)std::min<int64_t>(ctas, (int64_t)d.sm_count * 8);
    aba_derivatives_gemm_kernel<T><<<grid, warps * 32, dyn, d.s()>>>(nv, dq, ld_dq, dv, ld_dv, dtau, ld_dtau, B);
    p->launches += 1;
    CUDA_TRY(cudaGetLastError());
  }
  return BRBD_OK;
}
template brbd_status launch_aba_derivs<double>(brbd_pool *, DeviceCtx &, const double *, int64_t, const double *, int64_t, const double *, int64_t, double *, int64_t, double *, int64_t, double *, int64_t, double *, int64_t, int64_t);
template brbd_status launch_aba_derivs<float>(brbd_pool *, DeviceCtx &, const float *, int64_t, const float *, int64_t, const float *, int64_t, float *, int64_t, float *, int64_t, float *, int64_t, float *, int64_t, int64_t);
} // namespace brbd
