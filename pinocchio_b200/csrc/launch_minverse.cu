// launch_minverse.cu — launch of batched computeMinverse (algorithm/aba.hpp:106): MODE 1 of the cooperative computeABADerivatives kernel
#include "host_ctx.hpp"

namespace brbd
{
// computeMinverse: the Minv phases of the warp-cooperative computeABADerivatives kernel (MODE 1)
template<class T>
brbd_status launch_minverse(brbd_pool * p, DeviceCtx & d, const T * q, int64_t ldq, T * Minv, int64_t ldM, int64_t B)
{
  const ModelPOD<double> & M = p->model.pd;
  const int G = coop_group_size(M.nv);
  const AbaCoopLayout L = aba_coop_layout(M.nq, M.nv, M.njoints, G);
  const size_t static_bytes = sizeof(ModelPOD<T>) + sizeof(CoopTables) + 1024;
  const GeometryCoop g = pick_geometry_coop(d, (size_t)L.per_group * sizeof(T), G, static_bytes, B);
  if (p->model.coop.nbranch > A_MAXBRANCH || g.dyn_bytes + static_bytes > (size_t)d.max_smem_optin + 1024)
    return fail(BRBD_EINVAL, "computeMinverse: model too large for the shared-memory state of one configuration");
  brbd_status st = BRBD_OK;
#define BRBD_LAUNCH_COOP(GG)                                                                                     \
  {                                                                                                              \
    st = set_smem(aba_derivatives_coop_kernel<T, GG, 1>, g.dyn_bytes);                                           \
    if (st != BRBD_OK) return st;                                                                                \
    aba_derivatives_coop_kernel<T, GG, 1><<<g.grid, g.warps * 32, g.dyn_bytes, d.s()>>>(                         \
      dev_model<T>(d), p->model.coop, L, q, ldq, (const T *)nullptr, 0, (const T *)nullptr, 0, (T *)nullptr, 0, (T *)nullptr, 0, Minv, \
      ldM, (T *)nullptr, 0, B);                                                                                  \
  }
  if (G == 8) BRBD_LAUNCH_COOP(8)
  else if (G == 16) BRBD_LAUNCH_COOP(16)
  else BRBD_LAUNCH_COOP(32)
#undef BRBD_LAUNCH_COOP
  p->launches += 1;
  CUDA_TRY(cudaGetLastError());
  return BRBD_OK;
}
template brbd_status launch_minverse<double>(brbd_pool *, DeviceCtx &, const double *, int64_t, double *, int64_t, int64_t);
template brbd_status launch_minverse<float>(brbd_pool *, DeviceCtx &, const float *, int64_t, float *, int64_t, int64_t);
} // namespace brbd
