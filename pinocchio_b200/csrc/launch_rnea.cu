// launch_rnea.cu — launch of batched RNEA (rneaInParallel, parallel/rnea.hpp:38-83): cooperative kernel for small batches, rnea_dfs_kernel otherwise
#include "host_ctx.hpp"
#include "rnea_dfs.cuh"

namespace brbd
{
template<class T>
brbd_status launch_rnea(brbd_pool * p, DeviceCtx & d, const T * q, int64_t ldq, const T * v, int64_t ldv, const T * a,
                        int64_t lda, T * tau, int64_t ldtau, int64_t B)
{
  const TreePOD<T> & t = tree_of<T>(p);
  if (forced_path("BRBD_RNEA_V", "v1")) return launch_rnea_v1<T>(p, d, q, ldq, v, ldv, a, lda, tau, ldtau, B);
  if (!std::getenv("BRBD_RNEA_V") && use_generated<T>(p, BRBD_GEN_RNEA, B))
    return launch_generated<T>(p, d, BRBD_GEN_RNEA, q, ldq, v, ldv, a, lda, tau, ldtau, B);
  if (B <= coop_max_batch(false, p->model.pd.nv))
  {
    const ModelPOD<double> & M = p->model.pd;
    const CoopLayout L = coop_layout(M.nq, M.nv, M.njoints);
    const int G = coop_group_size(M.nv);
    const size_t static_bytes = sizeof(ModelPOD<T>) + sizeof(CoopTables) + 1024;
    const GeometryCoop g = pick_geometry_coop(d, (size_t)L.per_group * sizeof(T), G, static_bytes, B);
    if (g.dyn_bytes + static_bytes <= (size_t)d.max_smem_optin + 1024)
    {
      brbd_status st = BRBD_OK;
#define BRBD_LAUNCH_COOP(GG)                                                                                     \
  {                                                                                                              \
    st = set_smem(rnea_coop_kernel<T, GG>, g.dyn_bytes);                                                         \
    if (st != BRBD_OK) return st;                                                                                \
    rnea_coop_kernel<T, GG><<<g.grid, g.warps * 32, g.dyn_bytes, d.s()>>>(dev_model<T>(d), p->model.coop, L, q, ldq, v, ldv, a, lda, \
                                                                            tau, ldtau, B);                      \
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
  const RneaLayout L = rnea_layout(t.maxdepth, t.nbranch);
  // one CTA per SM, up to 8 warps, chosen by the number of rounds (as CRBA)
  const size_t per_warp = (size_t)32 * L.nstate * sizeof(T);
  if (per_warp > (size_t)d.max_smem_optin) // deeper / more branched than one warp's shared-memory state allows
    return launch_rnea_v1<T>(p, d, q, ldq, v, ldv, a, lda, tau, ldtau, B);
  int warps = (int)std::max<size_t>(1, std::min<size_t>(8, (size_t)d.max_smem_optin / per_warp));
  warps = pick_warps_by_rounds(d, B, warps);
  if (const char * e = std::getenv("BRBD_RNEA_WARPS")) warps = std::max(1, std::min(warps, std::atoi(e)));
  const size_t dyn_bytes = (size_t)warps * per_warp;
  const int64_t ctas_needed = (B + warps * 32 - 1) / (warps * 32);
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ctas_needed, (int64_t)d.sm_count));
  brbd_status st = BRBD_OK;
#define BRBD_LAUNCH(NT)                                                                              \
  {                                                                                                  \
    st = set_smem(rnea_dfs_kernel<T, NT>, dyn_bytes);                                                \
    if (st != BRBD_OK) return st;                                                                    \
    rnea_dfs_kernel<T, NT><<<grid, NT, dyn_bytes, d.s()>>>(t, L, q, ldq, v, ldv, a, lda, tau, ldtau, B); \
  }
  switch (warps)
  {
  case 1: BRBD_LAUNCH(32) break;
  case 2: BRBD_LAUNCH(64) break;
  case 3: BRBD_LAUNCH(96) break;
  case 4: BRBD_LAUNCH(128) break;
  case 5: BRBD_LAUNCH(160) break;
  case 6: BRBD_LAUNCH(192) break;
  case 7: BRBD_LAUNCH(224) break;
  default: BRBD_LAUNCH(256) break;
  }
#undef BRBD_LAUNCH
  p->launches += 1;
  CUDA_TRY(cudaGetLastError());
  return BRBD_OK;
}
template brbd_status launch_rnea<double>(brbd_pool *, DeviceCtx &, const double *, int64_t, const double *, int64_t, const double *, int64_t, double *, int64_t, int64_t);
template brbd_status launch_rnea<float>(brbd_pool *, DeviceCtx &, const float *, int64_t, const float *, int64_t, const float *, int64_t, float *, int64_t, int64_t);
} // namespace brbd
