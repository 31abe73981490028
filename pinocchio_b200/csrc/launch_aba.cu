// launch_aba.cu — launch of batched ABA (abaInParallel, parallel/aba.hpp:40-84), one configuration per thread: aba_rr (v4), aba_tmem (v3), aba_dfs
#include "host_ctx.hpp"
#include "aba_dfs.cuh"
#include "aba_rr.cuh"

namespace brbd
{
template<class T>
brbd_status launch_aba(brbd_pool * p, DeviceCtx & d, const T * q, int64_t ldq, const T * v, int64_t ldv, const T * tau,
                       int64_t ldtau, T * a, int64_t lda, int64_t B)
{
  const TreePOD<T> & t = tree_of<T>(p);
  brbd_status st = BRBD_OK;
  const char * ver = std::getenv("BRBD_ABA_V"); // "v3" (aba_tmem_kernel), "dfs" (aba_dfs_kernel), "v1" (aba_kernel)
  if (ver && std::strcmp(ver, "v1") == 0) return launch_aba_v1<T>(p, d, q, ldq, v, ldv, tau, ldtau, a, lda, B);
  if (!ver && use_generated<T>(p, BRBD_GEN_ABA, B)) return launch_generated<T>(p, d, BRBD_GEN_ABA, q, ldq, v, ldv, tau, ldtau, a, lda, B);
  {
    bool done = false;
    st = launch_aba_coop<T>(p, d, q, ldq, v, ldv, tau, ldtau, a, lda, B, &done);
    if (st != BRBD_OK || done) return st;
  }
  // preferred (v4): the backward sweep recomputes the per-depth quantities; (sin, cos, v) per depth and the branch slots in
  // tensor memory, only the pass-3 record ring in shared memory -> up to 8 warps per SM
  if (!ver)
  {
    const int wpv = (int)(sizeof(T) / 4);
    AbaRRLayout L = aba_rr_layout<T>(t.maxdepth, t.nbranch, 4);
    const int cols_per_slice = L.tvals * wpv;
    const int max_warps_tmem = cols_per_slice <= 256 ? 8 : (cols_per_slice <= 512 ? 4 : 0);
    if (max_warps_tmem > 0)
    {
      const size_t per_warp = (size_t)32 * L.nstate * sizeof(T);
      int warps = (int)std::max<size_t>(1, std::min<size_t>((size_t)max_warps_tmem, (size_t)d.max_smem_optin / per_warp));
      warps = pick_warps_by_rounds(d, B, warps);
      if (const char * e = std::getenv("BRBD_ABA_WARPS")) warps = std::max(1, std::min(warps, std::atoi(e)));
      const size_t dyn_bytes = (size_t)warps * per_warp;
      const int64_t ctas_needed = (B + warps * 32 - 1) / (warps * 32);
      const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ctas_needed, (int64_t)d.sm_count));
      L = aba_rr_layout<T>(t.maxdepth, t.nbranch, warps);
      st = ensure_work(d, (size_t)grid * warps * 32 * (size_t)t.pslots * sizeof(T));
      if (st != BRBD_OK) return st;
#define BRBD_LAUNCH(NT)                                                                              \
  {                                                                                                  \
    st = set_smem(aba_rr_kernel<T, NT>, dyn_bytes);                                                  \
    if (st != BRBD_OK) return st;                                                                    \
    aba_rr_kernel<T, NT><<<grid, NT, dyn_bytes, d.s()>>>(t, L, q, ldq, v, ldv, tau, ldtau, a, lda, (T *)d.work, B); \
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
  }
  // v3: per-depth (Y, f, a_bias) in tensor memory, J of the root path + branch slots in shared memory
  if (!ver || std::strcmp(ver, "v3") == 0)
  {
    AbaTmemLayout L = aba_tmem_layout<T>(t.maxpathdof, t.maxdepth, t.nbranch, 4);
    if (L.tvals * (int)(sizeof(T) / 4) <= 512 && (size_t)32 * L.nstate * sizeof(T) <= (size_t)d.max_smem_optin)
    {
      const Geometry2 g = pick_geometry2(d, (size_t)L.nstate * sizeof(T), 0, B, 4, 1);
      L = aba_tmem_layout<T>(t.maxpathdof, t.maxdepth, t.nbranch, g.warps);
      st = ensure_work(d, (size_t)g.grid * g.warps * 32 * (size_t)t.pslots * sizeof(T));
      if (st != BRBD_OK) return st;
#define BRBD_LAUNCH(NT)                                                                              \
  {                                                                                                  \
    st = set_smem(aba_tmem_kernel<T, NT>, g.dyn_bytes);                                              \
    if (st != BRBD_OK) return st;                                                                    \
    aba_tmem_kernel<T, NT><<<g.grid, NT, g.dyn_bytes, d.s()>>>(t, L, q, ldq, v, ldv, tau, ldtau, a, lda, (T *)d.work, B); \
  }
      BRBD_SWITCH_WARPS(g.warps)
#undef BRBD_LAUNCH
      p->launches += 1;
      CUDA_TRY(cudaGetLastError());
      return BRBD_OK;
    }
  }
  // fallback for very deep trees: per-depth state in shared memory
  const AbaLayout L = aba_layout(t.maxdepth, t.nbranch);
  if ((size_t)32 * L.nstate * sizeof(T) > (size_t)d.max_smem_optin) // not even one warp fits: generic kernel
    return launch_aba_v1<T>(p, d, q, ldq, v, ldv, tau, ldtau, a, lda, B);
  const Geometry2 g = pick_geometry2(d, (size_t)L.nstate * sizeof(T), 0, B, 4, 2);
  // per-thread persistent store (J, a_bias, U Dinv, Dinv, u of every joint), [slot][thread]
  st = ensure_work(d, (size_t)g.grid * g.warps * 32 * (size_t)t.pslots * sizeof(T));
  if (st != BRBD_OK) return st;
#define BRBD_LAUNCH(NT)                                                                              \
  {                                                                                                  \
    st = set_smem(aba_dfs_kernel<T, NT>, g.dyn_bytes);                                               \
    if (st != BRBD_OK) return st;                                                                    \
    aba_dfs_kernel<T, NT><<<g.grid, NT, g.dyn_bytes, d.s()>>>(t, L, q, ldq, v, ldv, tau, ldtau, a, lda, (T *)d.work, B); \
  }
  BRBD_SWITCH_WARPS(g.warps)
#undef BRBD_LAUNCH
  p->launches += 1;
  CUDA_TRY(cudaGetLastError());
  return BRBD_OK;
}
template brbd_status launch_aba<double>(brbd_pool *, DeviceCtx &, const double *, int64_t, const double *, int64_t, const double *, int64_t, double *, int64_t, int64_t);
template brbd_status launch_aba<float>(brbd_pool *, DeviceCtx &, const float *, int64_t, const float *, int64_t, const float *, int64_t, float *, int64_t, int64_t);
} // namespace brbd
