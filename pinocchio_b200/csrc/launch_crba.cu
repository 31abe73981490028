// launch_crba.cu — launch of batched CRBA (crba, algorithm/crba.hpp:47-51): TMA tensor-store kernel, LSU emitter, shared-memory fallback
#include "host_ctx.hpp"
#include "crba_dfs.cuh"

namespace brbd
{
// ---- TMA tensor maps over the caller's (nv*nv x B, leading dimension ldM) matrix block: see crba_tma_kernel ----------
typedef CUresult (*brbd_encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                         const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                         CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static brbd_encode_tiled_fn encode_tiled_fn()
{
  static brbd_encode_tiled_fn fn = [] {
    void * p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess)
      p = nullptr;
    return (brbd_encode_tiled_fn)p;
  }();
  return fn;
}
template<class T>
bool crba_tma_setup(T * Mout, int64_t ldM, int64_t B, int nv, CrbaTmaGeom & G, CUtensorMap & map0, CUtensorMap & map1)
{
  const brbd_encode_tiled_fn enc = encode_tiled_fn();
  constexpr int E = (int)sizeof(T), K = 16 / E;
  if (!enc || (reinterpret_cast<uintptr_t>(Mout) & 15) || nv > 255 || ldM < (int64_t)nv * nv) return false;
  const bool even = (nv % K) == 0 && (ldM % K) == 0;
  const bool odd = E == 8 && (nv & 1) && nv >= 3;
  if (!even && !odd) return false;
  G.odd = even ? 0 : 1;
  G.pairs = (!even && (ldM & 1)) ? 1 : 0;
  G.bx = even ? nv : nv + 1;
  const CUtensorMapDataType dt = E == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  const cuuint32_t es[2] = {1, 1};
  auto make = [&](CUtensorMap & mp, T * base, cuuint64_t inner, cuuint64_t outer, cuuint64_t stride_elems, cuuint32_t rows) {
    const cuuint64_t gd[2] = {inner, outer > 0 ? outer : 1};
    const cuuint64_t gs[1] = {stride_elems * (cuuint64_t)E};
    const cuuint32_t bd[2] = {(cuuint32_t)G.bx, rows};
    return enc(&mp, dt, 2, (void *)base, gd, gs, bd, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
               CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
  };
  if (!G.pairs)
  {
    if (!make(map0, Mout, (cuuint64_t)ldM, (cuuint64_t)B, (cuuint64_t)ldM, 32)) return false;
    map1 = map0;
    return true;
  }
  if (!make(map0, Mout, (cuuint64_t)ldM, (cuuint64_t)((B + 1) / 2), (cuuint64_t)(2 * ldM), 16)) return false;
  if (B < 2) { map1 = map0; return true; } // the kernel issues no odd-half store for a single configuration
  return make(map1, Mout + (ldM - 1), (cuuint64_t)(ldM + 1), (cuuint64_t)(B / 2), (cuuint64_t)(2 * ldM), 16);
}

template<class T>
brbd_status launch_crba(brbd_pool * p, DeviceCtx & d, const T * q, int64_t ldq, T * Mout, int64_t ldM, int64_t B)
{
  const TreePOD<T> & t = tree_of<T>(p);
  if (ldM >= (int64_t(1) << 25)) return fail(BRBD_EINVAL, "crba: leading dimension of M too large");
  brbd_status st = BRBD_OK;
  const char * ver = std::getenv("BRBD_CRBA_V"); // "tmem" (LSU emitter), "dfs" (crba_dfs_kernel), "v1" (crba_kernel)
  if (ver && std::strcmp(ver, "v1") == 0) return launch_crba_v1<T>(p, d, q, ldq, Mout, ldM, B);
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
