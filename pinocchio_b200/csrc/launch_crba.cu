// launch_crba.cu — launch of batched CRBA (crba, algorithm/crba.hpp:47-51): TMA tensor-store kernel, LSU emitter, shared-memory fallback
#include "host_ctx.hpp"
#include "crba_dfs.cuh"

namespace brbd
{
template<class T>
brbd_status launch_crba(brbd_pool * p, DeviceCtx & d, const T * q, int64_t ldq, T * Mout, int64_t ldM, int64_t B)
{
  const TreePOD<T> & t = tree_of<T>(p);
  if (ldM >= (int64_t(1) << 25)) return fail(BRBD_EINVAL, "crba: leading dimension of M too large");
  brbd_status st = BRBD_OK;
  const char * ver = std::getenv("BRBD_CRBA_V"); // "tmem" (LSU emitter), "dfs" (crba_dfs_kernel), "v1" (crba_kernel)
  if (ver && std::strcmp(ver, "v1") == 0) return launch_crba_v1<T>(p, d, q, ldq, Mout, ldM, B);
  // a pool specialised for its model runs the generated CRBA (launch_gen.cu): 6-dof manipulator 0.034 -> 0.016 ms, 35-dof
  // humanoid 0.225 -> 0.160 ms at 65 536 configurations; BRBD_CRBA_V=gen forces it below the specialised minimum batch too
  if ((!ver && use_generated<T>(p, BRBD_GEN_CRBA, B)) || (ver && std::strncmp(ver, "gen", 3) == 0 && p->gen[BRBD_GEN_CRBA][sizeof(T) == 4 ? 1 : 0].nvar > 0))
    return launch_generated<T>(p, d, BRBD_GEN_CRBA, q, ldq, (const T *)nullptr, 0, (const T *)nullptr, 0, Mout, ldM, B);
  // preferred: oYcrb / oMi stacks in tensor memory (<= 8 warps per CTA, one CTA per SM)
  if (!(ver && std::strcmp(ver, "dfs") == 0))
  {
    const int wpv = (int)(sizeof(T) / 4);
    CrbaTmemLayout L = crba_tmem_layout<T>(t.maxpathdof, t.maxdepth, t.nbranch, t.nv, 4, t.ffroot);
    const int cols_per_slice = L.tvals * wpv;
    const int max_warps_tmem = cols_per_slice <= 256 ? 8 : (cols_per_slice <= 512 ? 4 : 0);
    if (max_warps_tmem > 0 && (size_t)32 * (L.nstate + L.epad + 2) * sizeof(T) + 128 * (size_t)t.nv + 64 <= (size_t)d.max_smem_optin)
    {
      // default: column blocks leave through TMA tensor stores (crba_tma_kernel) where the caller's layout allows a tensor
      // map (see crba_dfs.cuh); BRBD_CRBA_V=tmem keeps the LSU emitter
      CrbaTmaGeom G{0, 0, 0};
      CUtensorMap map0, map1;
      const bool tma = !(ver && std::strcmp(ver, "tmem") == 0) && crba_tma_setup<T>(Mout, ldM, B, t.nv, G, map0, map1);
      if (tma) L.epad = G.bx;
      const int epad = L.epad;
      const size_t tab_bytes = tma ? 0 : 128 * (size_t)t.nv;
      Geometry2 g = pick_geometry2(d, (size_t)L.nstate * sizeof(T), (size_t)32 * L.epad * sizeof(T), B, max_warps_tmem, 1);
      // Per-SM throughput is flat from 5 warps up (measured, profiles/r1_v5_crba_warps.txt), so what counts is the number of
      // rounds the persistent grid needs: fewest rounds first, then the fewest warps that reach it (65536 configurations of
      // simple_humanoid: 7 warps -> 1.98 rounds, 8 -> 1.73 rounds of which the second is 73 % full, 6 -> 2.3 i.e. 3 rounds).
      g.warps = pick_warps_by_rounds(d, B, g.warps);
      if (const char * e = std::getenv("BRBD_CRBA_WARPS")) // experiments: cap the warps per SM
        g.warps = std::max(1, std::min(g.warps, std::atoi(e)));
      // the element -> global offset table of the emitter (32 * nv ints) sits after the warp regions
      while (g.warps > 1 && (size_t)g.warps * (32 * (size_t)L.nstate * sizeof(T) + 32 * (size_t)L.epad * sizeof(T)) + tab_bytes + 64 > (size_t)d.max_smem_optin) --g.warps;
      g.dyn_bytes = (size_t)g.warps * (32 * (size_t)L.nstate * sizeof(T) + 32 * (size_t)L.epad * sizeof(T)) + tab_bytes;
      const int64_t ctas_needed = (B + g.warps * 32 - 1) / (g.warps * 32);
      g.grid = (int)std::max<int64_t>(1, std::min<int64_t>(ctas_needed, (int64_t)d.sm_count));
      L = crba_tmem_layout<T>(t.maxpathdof, t.maxdepth, t.nbranch, t.nv, g.warps, t.ffroot);
      L.epad = epad;
#define BRBD_LAUNCH(NT)                                                                              \
  {                                                                                                  \
    if (tma && G.odd)                                                                                \
    {                                                                                                \
      st = set_smem(crba_tma_kernel<T, NT, true>, g.dyn_bytes);                                      \
      if (st != BRBD_OK) return st;                                                                  \
      crba_tma_kernel<T, NT, true><<<g.grid, NT, g.dyn_bytes, d.s()>>>(t, L, G, map0, map1, q, ldq, Mout, ldM, B); \
    }                                                                                                \
    else if (tma)                                                                                    \
    {                                                                                                \
      st = set_smem(crba_tma_kernel<T, NT, false>, g.dyn_bytes);                                     \
      if (st != BRBD_OK) return st;                                                                  \
      crba_tma_kernel<T, NT, false><<<g.grid, NT, g.dyn_bytes, d.s()>>>(t, L, G, map0, map1, q, ldq, Mout, ldM, B); \
    }                                                                                                \
    else                                                                                             \
    {                                                                                                \
      st = set_smem(crba_tmem_kernel<T, NT>, g.dyn_bytes);                                           \
      if (st != BRBD_OK) return st;                                                                  \
      crba_tmem_kernel<T, NT><<<g.grid, NT, g.dyn_bytes, d.s()>>>(t, L, q, ldq, Mout, ldM, B);       \
    }                                                                                                \
  }
      switch (g.warps)
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
  // fallback for very deep trees: all state in shared memory
  const CrbaLayout L = crba_layout(t.maxpathdof, t.maxdepth, t.nbranch, t.nv);
  if ((size_t)32 * (L.nstate + L.epad) * sizeof(T) + 128 * (size_t)t.nv > (size_t)d.max_smem_optin)
    return launch_crba_v1<T>(p, d, q, ldq, Mout, ldM, B);
  const Geometry2 g = pick_geometry2(d, (size_t)L.nstate * sizeof(T), (size_t)32 * L.epad * sizeof(T) + 128 * t.nv, B, 4, 2);
#define BRBD_LAUNCH(NT)                                                                              \
  {                                                                                                  \
    st = set_smem(crba_dfs_kernel<T, NT>, g.dyn_bytes);                                              \
    if (st != BRBD_OK) return st;                                                                    \
    crba_dfs_kernel<T, NT><<<g.grid, NT, g.dyn_bytes, d.s()>>>(t, L, q, ldq, Mout, ldM, B);          \
  }
  BRBD_SWITCH_WARPS(g.warps)
#undef BRBD_LAUNCH
  p->launches += 1;
  CUDA_TRY(cudaGetLastError());
  return BRBD_OK;
}
template brbd_status launch_crba<double>(brbd_pool *, DeviceCtx &, const double *, int64_t, double *, int64_t, int64_t);
template brbd_status launch_crba<float>(brbd_pool *, DeviceCtx &, const float *, int64_t, float *, int64_t, int64_t);
} // namespace brbd
