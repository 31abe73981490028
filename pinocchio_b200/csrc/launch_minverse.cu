// launch_minverse.cu — launch of batched computeMinverse (algorithm/aba.hpp:106): the batched CRBA followed by the dense
// Cholesky inversion of minv_chol.cuh (default), or MODE 1 of the cooperative computeABADerivatives kernel (BRBD_MINV_V=coop)
#include "host_ctx.hpp"
#include "minv_chol.cuh"

namespace brbd
{
// computeMinverse: the Minv phases of the warp-cooperative computeABADerivatives kernel (MODE 1)
// M (crba, whatever kernel the pool runs for it) into the device's aux buffer, chunk by chunk, then M^-1 by Cholesky
template<class T>
brbd_status launch_minverse_chol(brbd_pool * p, DeviceCtx & d, const T * q, int64_t ldq, T * Minv, int64_t ldM, int64_t B)
{
  const int nv = p->model.pd.nv;
  const int64_t nn = (int64_t)nv * nv;
  const int G = coop_group_size(nv);
  const MinvCholLayout L = minv_chol_layout(nv);
  const MinvCholBlockedLayout LB = minv_chol_blocked_layout(nv); // 32 lanes per configuration: the 4 x 4 blocked kernel
  const size_t per_warp = G == 32 ? (size_t)LB.per_group * sizeof(T) : (size_t)L.per_group * sizeof(T) * (32 / G);
  int warps = (int)std::max<size_t>(1, std::min<size_t>(16, ((size_t)d.max_smem_optin - 1024) / per_warp));
  if (const char * e = std::getenv("BRBD_MINV_WARPS")) warps = std::max(1, std::min(warps, std::atoi(e))); // experiments
  // small batches: spread the configurations over all SMs
  while (warps > 1 && (int64_t)(warps - 1) * (32 / G) * d.sm_count >= B) --warps;
  const size_t dyn = (size_t)warps * per_warp;
  const int64_t chunk = std::min<int64_t>(B, 32768); // M of a chunk: 321 MB for a 35-dof humanoid in FP64
  brbd_status st = ensure_aux(d, (size_t)chunk * nn * sizeof(T));
  if (st != BRBD_OK) return st;
  T * Mbuf = (T *)d.aux;
  for (int64_t c0 = 0; c0 < B; c0 += chunk)
  {
    const int64_t bc = std::min<int64_t>(chunk, B - c0);
    st = launch_crba<T>(p, d, q + c0 * ldq, ldq, Mbuf, nn, bc);
    if (st != BRBD_OK) return st;
    const int64_t per_cta = (int64_t)warps * (32 / G);
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((bc + per_cta - 1) / per_cta, (int64_t)d.sm_count));
#define BRBD_LAUNCH_CHOL(GG, RR)                                                                                 \
  {                                                                                                              \
    st = set_smem(minv_chol_kernel<T, GG, RR>, dyn);                                                             \
    if (st != BRBD_OK) return st;                                                                                \
    minv_chol_kernel<T, GG, RR><<<grid, warps * 32, dyn, d.s()>>>(Mbuf, nn, Minv + c0 * ldM, ldM, nv, L, bc);    \
  }
#define BRBD_LAUNCH_BLOCKED(RR)                                                                                  \
  {                                                                                                              \
    st = set_smem(minv_chol_blocked_kernel<T, RR>, dyn);                                                         \
    if (st != BRBD_OK) return st;                                                                                \
    minv_chol_blocked_kernel<T, RR><<<grid, warps * 32, dyn, d.s()>>>(Mbuf, nn, Minv + c0 * ldM, ldM, nv, LB, bc); \
  }
    if (G == 8) BRBD_LAUNCH_CHOL(8, 1)
    else if (G == 16) BRBD_LAUNCH_CHOL(16, 1)
    else if (LB.nvp <= 32) BRBD_LAUNCH_BLOCKED(1)
    else BRBD_LAUNCH_BLOCKED(2)
#undef BRBD_LAUNCH_BLOCKED
#undef BRBD_LAUNCH_CHOL
    p->launches += 1;
    CUDA_TRY(cudaGetLastError());
  }
  return BRBD_OK;
}

template<class T>
brbd_status launch_minverse(brbd_pool * p, DeviceCtx & d, const T * q, int64_t ldq, T * Minv, int64_t ldM, int64_t B)
{
  const ModelPOD<double> & M = p->model.pd;
  // crba + Cholesky (minv_chol.cuh) is the default — 65 536 configurations: 6-dof manipulator 0.209 -> 0.097 ms, 35-dof humanoid
  // 4.08 -> 2.06 ms, talos 4.9 -> 2.5 ms (profiles/r2_minv_chol.txt); BRBD_MINV_V=coop keeps the articulated-body kernel
  const bool chol = !forced_path("BRBD_MINV_V", "coop");
  if (chol) return launch_minverse_chol<T>(p, d, q, ldq, Minv, ldM, B);
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
