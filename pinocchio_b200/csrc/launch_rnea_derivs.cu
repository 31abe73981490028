// launch_rnea_derivs.cu — launch of batched computeRNEADerivatives (rnea-derivatives.hpp:110-128)
#include "host_ctx.hpp"

namespace brbd
{
template<class T>
brbd_status launch_rnea_derivs(brbd_pool * p, DeviceCtx & d, const T * q, int64_t ldq, const T * v, int64_t ldv,
                               const T * a, int64_t lda, T * dq, int64_t ld_dq, T * dv, int64_t ld_dv, T * da,
                               int64_t ld_da, T * tau, int64_t ldtau, int64_t B)
{
  const ModelPOD<double> & M = p->model.pd;
  if (!std::getenv("BRBD_DRNEA_V") && use_generated<T>(p, BRBD_GEN_RNEA_DERIVATIVES, B))
    return launch_generated_derivs<T>(p, d, BRBD_GEN_RNEA_DERIVATIVES, q, ldq, v, ldv, a, lda, dq, ld_dq, dv, ld_dv, da, ld_da, tau, ldtau, B);
  const CoopLayout L = coop_layout(M.nq, M.nv, M.njoints);
  const int G = coop_group_size(M.nv);
  const size_t static_bytes = sizeof(ModelPOD<T>) + sizeof(CoopTables) + 1024;
  const GeometryCoop g = pick_geometry_coop(d, (size_t)L.per_group * sizeof(T), G, static_bytes, B);
  // models whose per-configuration state does not fit the cooperative layout run the generic kernel
  if (g.dyn_bytes + static_bytes > (size_t)d.max_smem_optin + 1024 || forced_path("BRBD_DRNEA_V", "v1"))
    return launch_rnea_derivs_v1<T>(p, d, q, ldq, v, ldv, a, lda, dq, ld_dq, dv, ld_dv, da, ld_da, tau, ldtau, B);
  brbd_status st = BRBD_OK;
#define BRBD_LAUNCH_COOP(GG)                                                                                     \
  {                                                                                                              \
    st = set_smem(rnea_derivatives_coop_kernel<T, GG>, g.dyn_bytes);                                             \
    if (st != BRBD_OK) return st;                                                                                \
    rnea_derivatives_coop_kernel<T, GG><<<g.grid, g.warps * 32, g.dyn_bytes, d.s()>>>(                           \
      dev_model<T>(d), p->model.coop, L, q, ldq, v, ldv, a, lda, dq, ld_dq, dv, ld_dv, da, ld_da, tau, ldtau, B); \
  }
  if (G == 8) BRBD_LAUNCH_COOP(8)
  else if (G == 16) BRBD_LAUNCH_COOP(16)
  else BRBD_LAUNCH_COOP(32)
#undef BRBD_LAUNCH_COOP
  p->launches += 1;
  CUDA_TRY(cudaGetLastError());
  return BRBD_OK;
}
template brbd_status launch_rnea_derivs<double>(brbd_pool *, DeviceCtx &, const double *, int64_t, const double *, int64_t, const double *, int64_t, double *, int64_t, double *, int64_t, double *, int64_t, double *, int64_t, int64_t);
template brbd_status launch_rnea_derivs<float>(brbd_pool *, DeviceCtx &, const float *, int64_t, const float *, int64_t, const float *, int64_t, float *, int64_t, float *, int64_t, float *, int64_t, float *, int64_t, int64_t);
} // namespace brbd
