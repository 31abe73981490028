// launch_aba_coop.cu — small-batch ABA: the A1 / A2 / A4 phases of the warp-cooperative computeABADerivatives kernel (MODE 2)
#include "host_ctx.hpp"

namespace brbd
{
template<class T>
brbd_status launch_aba_coop(brbd_pool * p, DeviceCtx & d, const T * q, int64_t ldq, const T * v, int64_t ldv, const T * tau,
                            int64_t ldtau, T * a, int64_t lda, int64_t B, bool * done)
{
  brbd_status st = BRBD_OK;
  *done = false;
  if (B <= coop_max_batch(true, p->model.pd.nv) && p->model.coop.nbranch <= A_MAXBRANCH)
  {
    const ModelPOD<double> & M = p->model.pd;
    const int G = coop_group_size(M.nv);
    const AbaCoopLayout L = aba_coop_layout(M.nq, M.nv, M.njoints, G);
    const size_t static_bytes = sizeof(ModelPOD<T>) + sizeof(CoopTables) + 1024;
    const GeometryCoop g = pick_geometry_coop(d, (size_t)L.per_group * sizeof(T), G, static_bytes, B);
    if (g.dyn_bytes + static_bytes <= (size_t)d.max_smem_optin + 1024)
    {
#define BRBD_LAUNCH_COOP(GG)                                                                                     \
  {                                                                                                              \
    st = set_smem(aba_derivatives_coop_kernel<T, GG, 2>, g.dyn_bytes);                                           \
    if (st != BRBD_OK) return st;                                                                                \
    aba_derivatives_coop_kernel<T, GG, 2><<<g.grid, g.warps * 32, g.dyn_bytes, d.s()>>>(                         \
      dev_model<T>(d), p->model.coop, L, q, ldq, v, ldv, tau, ldtau, (T *)nullptr, 0, (T *)nullptr, 0, (T *)nullptr, 0, a, lda, B); \
  }
      if (G == 8) BRBD_LAUNCH_COOP(8)
      else if (G == 16) BRBD_LAUNCH_COOP(16)
      else BRBD_LAUNCH_COOP(32)
#undef BRBD_LAUNCH_COOP
      p->launches += 1;
      CUDA_TRY(cudaGetLastError());
      *done = true;
      return BRBD_OK;
    }
  }
  return st;
}

template brbd_status launch_aba_coop<double>(brbd_pool *, DeviceCtx &, const double *, int64_t, const double *, int64_t, const double *, int64_t, double *, int64_t, int64_t, bool *);
template brbd_status launch_aba_coop<float>(brbd_pool *, DeviceCtx &, const float *, int64_t, const float *, int64_t, const float *, int64_t, float *, int64_t, int64_t, bool *);
} // namespace brbd
