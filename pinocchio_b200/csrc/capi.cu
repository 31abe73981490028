// capi.cu — host side of the engine behind the C ABI declared in include/pinocchio_b200.h.
//
// brbd_model : validated copy of the flattened model + derived topology tables
//              (nvSubtree / parents_fromRow: reference multibody/data.hxx:197-315; the model
//              checks the reference runs at algorithm entry — parents[i] < i, CRBAChecker
//              crba.hxx:573-595 — run once here).
// brbd_pool  : device analogue of ModelPoolTpl (multibody/pool/model.hpp:19-165): one staged model
//              replica, one stream and one grow-only staging arena per device.
// *_batch    : the drop-in for rneaInParallel / abaInParallel (algorithm/parallel/rnea.hpp:38-83,
//              parallel/aba.hpp:40-84) and their crba / derivative analogues.  The batch is split in
//              contiguous column ranges over the pool's devices; there is no inter-device exchange.
// There is no CPU fallback anywhere in this file.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/pinocchio_b200.h"
#include "aba.cuh"
#include "aba_derivatives.cuh"
#include "aba_deriv_coop.cuh"
#include "aba_dfs.cuh"
#include "aba_rr.cuh"
#include "crba.cuh"
#include "crba_dfs.cuh"
#include "engine.cuh"
#include "model_build.hpp"
#include "rnea.cuh"
#include "rnea_derivatives.cuh"
#include "deriv_coop.cuh"
#include "integrate.cuh"
#include "rnea_dfs.cuh"
#include "tree.cuh"

using namespace brbd;

namespace
{
thread_local std::string g_err;
brbd_status fail(brbd_status s, const std::string & msg)
{
  g_err = msg;
  return s;
}
#define CUDA_TRY(expr)                                                                              \
  do                                                                                                \
  {                                                                                                 \
    cudaError_t e__ = (expr);                                                                       \
    if (e__ != cudaSuccess)                                                                         \
      return fail(BRBD_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));                 \
  } while (0)

} // namespace

struct brbd_model
{
  ModelPOD<double> pd;
  ModelPOD<float> pf;
  TreePOD<double> td; // v2 kernels: passed by value as a __grid_constant__ kernel parameter
  TreePOD<float> tf;
  CoopTables coop; // warp-cooperative derivative kernels: level lists, ancestor masks
};

namespace
{
struct DeviceCtx
{
  int dev = -1;
  int sm_count = 0;
  int max_smem_optin = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t user_stream = nullptr;
  bool use_user_stream = false;
  ModelPOD<double> * d_pd = nullptr;
  ModelPOD<float> * d_pf = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // host-pointer calls: copy-in / copy-out streams beside the compute stream, and grow-only staging,
  // double-buffered (slot = 2 * argument + buffer) so that chunk k+1 uploads and chunk k-1 downloads
  // while chunk k computes
  cudaStream_t s_in = nullptr, s_out = nullptr;
  cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_k[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
  void * stage[16] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                      nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  size_t stage_bytes[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  // grow-only device workspace (intermediates of aba-derivatives)
  void * work = nullptr;
  size_t work_bytes = 0;
  // second grow-only buffer: the ABA result between the two kernels of the Euler step
  void * aux = nullptr;
  size_t aux_bytes = 0;
  // MAXNV zeros: the `v` / `a` operand (leading dimension 0) of nonLinearEffects / computeGeneralizedGravity
  void * zeros = nullptr;
  cudaStream_t s() const { return use_user_stream ? user_stream : stream; }
};
} // namespace

struct brbd_pool
{
  brbd_model model;
  std::vector<DeviceCtx> devs;
  int64_t launches = 0;
  double last_ms = 0.0;
};

namespace
{
brbd_status ensure_stage(DeviceCtx & d, int slot, size_t bytes)
{
  if (d.stage_bytes[slot] >= bytes) return BRBD_OK;
  if (d.stage[slot]) CUDA_TRY(cudaFree(d.stage[slot]));
  d.stage[slot] = nullptr;
  d.stage_bytes[slot] = 0;
  CUDA_TRY(cudaMalloc(&d.stage[slot], bytes));
  d.stage_bytes[slot] = bytes;
  return BRBD_OK;
}
brbd_status ensure_work(DeviceCtx & d, size_t bytes)
{
  if (d.work_bytes >= bytes) return BRBD_OK;
  if (d.work) CUDA_TRY(cudaFree(d.work));
  d.work = nullptr;
  d.work_bytes = 0;
  CUDA_TRY(cudaMalloc(&d.work, bytes));
  d.work_bytes = bytes;
  return BRBD_OK;
}

brbd_status ensure_aux(DeviceCtx & d, size_t bytes)
{
  if (d.aux_bytes >= bytes) return BRBD_OK;
  if (d.aux) CUDA_TRY(cudaFree(d.aux));
  d.aux = nullptr;
  d.aux_bytes = 0;
  CUDA_TRY(cudaMalloc(&d.aux, bytes));
  d.aux_bytes = bytes;
  return BRBD_OK;
}

// Launch geometry for the warp-tile kernels: `per_warp` bytes of dynamic shared memory per warp,
// `static_bytes` of static shared memory per CTA. Picks the CTA size that maximises resident
// warps per SM, and a persistent grid (CTAs loop over tiles).
struct Geometry
{
  int warps_per_cta, ctas_per_sm, grid;
  size_t dyn_bytes;
};
Geometry pick_geometry(const DeviceCtx & d, size_t per_warp, size_t static_bytes, int64_t batch, int max_warps_per_cta,
                       int max_warps_per_sm)
{
  const size_t sm_total = 227 * 1024; // usable shared memory per SM on sm_100
  Geometry best{1, 1, 1, per_warp};
  int best_warps = 0;
  for (int ctas = 1; ctas <= 8; ++ctas)
  {
    const size_t per_cta = sm_total / ctas;
    if (per_cta < static_bytes + 1024 + per_warp) break;
    int w = (int)((per_cta - static_bytes - 1024) / per_warp);
    w = std::min(w, max_warps_per_cta);
    w = std::min(w, std::max(1, max_warps_per_sm / ctas));
    if (w < 1) break;
    if ((size_t)w * per_warp + static_bytes > (size_t)d.max_smem_optin + 0) w = (int)((d.max_smem_optin - static_bytes) / per_warp);
    if (w < 1) break;
    if (w * ctas > best_warps)
    {
      best_warps = w * ctas;
      best.warps_per_cta = w;
      best.ctas_per_sm = ctas;
    }
  }
  best.dyn_bytes = (size_t)best.warps_per_cta * per_warp;
  const int64_t ntiles = (batch + 31) / 32;
  const int64_t ctas_needed = (ntiles + best.warps_per_cta - 1) / best.warps_per_cta;
  best.grid = (int)std::max<int64_t>(1, std::min<int64_t>(ctas_needed, (int64_t)d.sm_count * best.ctas_per_sm));
  return best;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is raised only when a launch needs more than what was set before for that
// kernel on that device (the call costs microseconds, which is what a small batch is made of)
template<class K> brbd_status set_smem(K kernel, size_t dyn_bytes)
{
  static std::mutex mu;
  static std::map<std::pair<int, const void *>, size_t> done;
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  const std::pair<int, const void *> key(dev, reinterpret_cast<const void *>(kernel));
  {
    std::lock_guard<std::mutex> lock(mu);
    auto it = done.find(key);
    if (it != done.end() && it->second >= dyn_bytes) return BRBD_OK;
  }
  CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_bytes));
  std::lock_guard<std::mutex> lock(mu);
  done[key] = dyn_bytes;
  return BRBD_OK;
}

template<class T> const ModelPOD<T> * dev_model(const DeviceCtx & d);
template<> const ModelPOD<double> * dev_model<double>(const DeviceCtx & d) { return d.d_pd; }
template<> const ModelPOD<float> * dev_model<float>(const DeviceCtx & d) { return d.d_pf; }

// ------------------------------------------------------------------------------------------------
// Device-pointer launches (one device)
// ------------------------------------------------------------------------------------------------
template<class T> const TreePOD<T> & tree_of(const brbd_pool * p);
template<> const TreePOD<double> & tree_of<double>(const brbd_pool * p) { return p->model.td; }
template<> const TreePOD<float> & tree_of<float>(const brbd_pool * p) { return p->model.tf; }

// Launch geometry of the v2 (DFS-interleaved) kernels: `state_bytes` of shared memory per thread plus
// `warp_bytes` per warp; as many warps per CTA as fit (<= max_warps), as many CTAs per SM as fit
// (<= max_ctas), persistent grid.
struct Geometry2
{
  int warps, ctas_per_sm, grid;
  size_t dyn_bytes;
};
Geometry2 pick_geometry2(const DeviceCtx & d, size_t state_bytes, size_t warp_bytes, int64_t batch, int max_warps, int max_ctas)
{
  const size_t per_warp = 32 * state_bytes + warp_bytes;
  const size_t cap = (size_t)d.max_smem_optin;
  Geometry2 g;
  g.warps = (int)std::max<size_t>(1, std::min<size_t>((size_t)max_warps, cap / per_warp));
  g.dyn_bytes = (size_t)g.warps * per_warp;
  const size_t sm_total = 228 * 1024; // per SM; every resident CTA also reserves 1 KB
  g.ctas_per_sm = (int)std::max<size_t>(1, std::min<size_t>((size_t)max_ctas, sm_total / (g.dyn_bytes + 1024)));
  const int64_t ctas_needed = (batch + g.warps * 32 - 1) / (g.warps * 32);
  g.grid = (int)std::max<int64_t>(1, std::min<int64_t>(ctas_needed, (int64_t)d.sm_count * g.ctas_per_sm));
  return g;
}

// Warps per SM of a persistent one-CTA-per-SM kernel whose per-SM throughput does not grow with occupancy any more: fewest
// rounds of the grid over the batch first, then the fewest warps that reach that number of rounds.
inline int pick_warps_by_rounds(const DeviceCtx & d, int64_t B, int wmax)
{
  const int64_t tiles = (B + 31) / 32;
  int best_w = wmax;
  int64_t best_rounds = (tiles + (int64_t)d.sm_count * wmax - 1) / ((int64_t)d.sm_count * wmax);
  for (int w = wmax - 1; w >= 1; --w)
  {
    const int64_t rounds = (tiles + (int64_t)d.sm_count * w - 1) / ((int64_t)d.sm_count * w);
    if (rounds <= best_rounds) { best_rounds = rounds; best_w = w; }
  }
  return best_w;
}

// kernels are instantiated for 1..4 warps per CTA (NT = threads per CTA is a template parameter)
#define BRBD_SWITCH_WARPS(w)              \
  switch (w)                              \
  {                                       \
  case 1: BRBD_LAUNCH(32) break;          \
  case 2: BRBD_LAUNCH(64) break;          \
  case 3: BRBD_LAUNCH(96) break;          \
  default: BRBD_LAUNCH(128) break;        \
  }

// Launch geometry of the warp-cooperative kernels: G lanes per configuration, `per_group` elements of shared
// memory per configuration, one CTA per SM with as many warps as fit (<= 8), persistent grid.
struct GeometryCoop
{
  int warps, grid;
  size_t dyn_bytes;
};
inline int coop_group_size(int nv) { return nv <= 8 ? 8 : (nv <= 16 ? 16 : 32); }
GeometryCoop pick_geometry_coop(const DeviceCtx & d, size_t group_bytes, int G, size_t static_bytes, int64_t batch)
{
  const size_t per_warp = group_bytes * (size_t)(32 / G);
  const size_t cap = (size_t)d.max_smem_optin - static_bytes;
  GeometryCoop g;
  g.warps = (int)std::max<size_t>(1, std::min<size_t>(8, cap / per_warp));
  const int64_t per_cta = (int64_t)g.warps * (32 / G);
  g.grid = (int)std::max<int64_t>(1, std::min<int64_t>((batch + per_cta - 1) / per_cta, (int64_t)d.sm_count));
  // small batches: spread the configurations over all SMs instead of filling a few CTAs
  while (g.warps > 1 && (int64_t)(g.warps - 1) * (32 / G) * d.sm_count >= batch) --g.warps;
  g.dyn_bytes = (size_t)g.warps * per_warp;
  const int64_t per_cta2 = (int64_t)g.warps * (32 / G);
  g.grid = (int)std::max<int64_t>(1, std::min<int64_t>((batch + per_cta2 - 1) / per_cta2, (int64_t)d.sm_count));
  return g;
}
// Below this many configurations per device the one-configuration-per-thread kernels cannot fill the GPU and their latency
// (one thread walking the whole tree) dominates: rneaInParallel / abaInParallel switch to the cooperative kernels
// (G lanes per configuration).  Measured crossover: profiles/r1_v5_small_batch.txt.
inline int64_t coop_max_batch(bool aba, int nv)
{
  if (const char * e = std::getenv("BRBD_COOP_MAX_BATCH")) return std::atoll(e);
  if (nv <= 8) return 4096; // 4 configurations per warp
  return aba ? 2048 : 4096;
}

template<class T>
brbd_status launch_rnea(brbd_pool * p, DeviceCtx & d, const T * q, int64_t ldq, const T * v, int64_t ldv, const T * a,
                        int64_t lda, T * tau, int64_t ldtau, int64_t B)
{
  const TreePOD<T> & t = tree_of<T>(p);
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

template<class T>
brbd_status launch_aba(brbd_pool * p, DeviceCtx & d, const T * q, int64_t ldq, const T * v, int64_t ldv, const T * tau,
                       int64_t ldtau, T * a, int64_t lda, int64_t B)
{
  const TreePOD<T> & t = tree_of<T>(p);
  brbd_status st = BRBD_OK;
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
      return BRBD_OK;
    }
  }
  // preferred (v4): the backward sweep recomputes the per-depth quantities; (sin, cos, v) per depth and the branch slots in
  // tensor memory, only the pass-3 record ring in shared memory -> up to 8 warps per SM
  if (!std::getenv("BRBD_ABA_V3"))
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
  {
    AbaTmemLayout L = aba_tmem_layout<T>(t.maxpathdof, t.maxdepth, t.nbranch, 4);
    if (L.tvals * (int)(sizeof(T) / 4) <= 512)
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
  // preferred: oYcrb / oMi stacks in tensor memory (<= 8 warps per CTA, one CTA per SM)
  {
    const int wpv = (int)(sizeof(T) / 4);
    CrbaTmemLayout L = crba_tmem_layout<T>(t.maxpathdof, t.maxdepth, t.nbranch, t.nv, 4, t.ffroot);
    const int cols_per_slice = L.tvals * wpv;
    const int max_warps_tmem = cols_per_slice <= 256 ? 8 : (cols_per_slice <= 512 ? 4 : 0);
    if (max_warps_tmem > 0)
    {
      // default: column blocks leave through TMA tensor stores (crba_tma_kernel) where the caller's layout allows a tensor
      // map (see crba_dfs.cuh); BRBD_CRBA_V=tmem keeps the LSU emitter
      const char * ver = std::getenv("BRBD_CRBA_V");
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

template<class T>
brbd_status launch_rnea_derivs(brbd_pool * p, DeviceCtx & d, const T * q, int64_t ldq, const T * v, int64_t ldv,
                               const T * a, int64_t lda, T * dq, int64_t ld_dq, T * dv, int64_t ld_dv, T * da,
                               int64_t ld_da, T * tau, int64_t ldtau, int64_t B)
{
  const ModelPOD<double> & M = p->model.pd;
  const CoopLayout L = coop_layout(M.nq, M.nv, M.njoints);
  const int G = coop_group_size(M.nv);
  const size_t static_bytes = sizeof(ModelPOD<T>) + sizeof(CoopTables) + 1024;
  const GeometryCoop g = pick_geometry_coop(d, (size_t)L.per_group * sizeof(T), G, static_bytes, B);
  if (g.dyn_bytes + static_bytes > (size_t)d.max_smem_optin + 1024)
    return fail(BRBD_EINVAL, "computeRNEADerivatives: model too large for the shared-memory state of one configuration");
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

template<class T>
brbd_status launch_aba_derivs(brbd_pool * p, DeviceCtx & d, const T * q, int64_t ldq, const T * v, int64_t ldv,
                              const T * tau, int64_t ldtau, T * dq, int64_t ld_dq, T * dv, int64_t ld_dv, T * dtau,
                              int64_t ld_dtau, T * ddq, int64_t ldddq, int64_t B)
{
  const ModelPOD<double> & M = p->model.pd;
  // preferred: one kernel, G lanes per configuration, everything in shared memory (aba_deriv_coop.cuh)
  {
    const int G = coop_group_size(M.nv);
    const AbaCoopLayout L = aba_coop_layout(M.nq, M.nv, M.njoints, G);
    const size_t static_bytes = sizeof(ModelPOD<T>) + sizeof(CoopTables) + 1024;
    const GeometryCoop g = pick_geometry_coop(d, (size_t)L.per_group * sizeof(T), G, static_bytes, B);
    if (p->model.coop.nbranch <= A_MAXBRANCH && g.dyn_bytes + static_bytes <= (size_t)d.max_smem_optin + 1024)
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

// integrate / Euler step: one configuration per thread, warp tiles staged through shared memory
template<class T, bool EULER>
brbd_status launch_integrate(brbd_pool * p, DeviceCtx & d, const T * q, int64_t ldq, const T * v, int64_t ldv, const T * a,
                             int64_t lda, T dt, T * qout, int64_t ldqo, T * vout, int64_t ldvo, int64_t B)
{
  const ModelPOD<double> & M = p->model.pd;
  const size_t per_warp = (size_t)32 * ((M.nq | 1) + 2 * (M.nv | 1)) * sizeof(T);
  const Geometry g = pick_geometry(d, per_warp, sizeof(ModelPOD<T>), B, 8, 16);
  brbd_status st = set_smem(integrate_kernel<T, EULER>, g.dyn_bytes);
  if (st != BRBD_OK) return st;
  integrate_kernel<T, EULER><<<g.grid, g.warps_per_cta * 32, g.dyn_bytes, d.s()>>>(dev_model<T>(d), q, ldq, v, ldv, a, lda, dt, qout,
                                                                                    ldqo, vout, ldvo, B);
  p->launches += 1;
  CUDA_TRY(cudaGetLastError());
  return BRBD_OK;
}

// ------------------------------------------------------------------------------------------------
// Generic call wrapper: argument checks, device/host pointer handling, sharding over devices.
// ------------------------------------------------------------------------------------------------
struct Arg
{
  const void * in;  // non-null for inputs
  void * out;       // non-null for outputs
  int64_t ld;
  int64_t rows;
  bool optional;
};

template<class T, class F>
brbd_status run_call(brbd_pool * p, std::vector<Arg> & args, int64_t B, int flags, F && launch)
{
  if (!p) return fail(BRBD_EINVAL, "null pool");
  if (p->devs.empty()) return fail(BRBD_EINVAL, "The pool should have at least one element"); // parallel/rnea.hpp:52
  if (B < 0) return fail(BRBD_EINVAL, "negative batch size");
  for (size_t k = 0; k < args.size(); ++k)
  {
    const Arg & a = args[k];
    const bool present = a.in || a.out;
    if (!present && !a.optional) return fail(BRBD_EINVAL, "null pointer argument #" + std::to_string(k));
    if (present && a.ld < a.rows)
      return fail(BRBD_EINVAL, "argument #" + std::to_string(k) + ": leading dimension " + std::to_string(a.ld)
                                 + " smaller than the expected number of rows " + std::to_string(a.rows));
  }
  if (B == 0) return BRBD_OK;
  const bool device_ptrs = (flags & BRBD_PTR_DEVICE) != 0;
  if (device_ptrs)
  {
    if (p->devs.size() != 1) return fail(BRBD_EINVAL, "device pointers require a single-device pool");
    DeviceCtx & d = p->devs[0];
    CUDA_TRY(cudaSetDevice(d.dev));
    std::vector<void *> ptrs(args.size());
    for (size_t k = 0; k < args.size(); ++k) ptrs[k] = args[k].in ? const_cast<void *>(args[k].in) : args[k].out;
    CUDA_TRY(cudaEventRecord(d.ev0, d.s()));
    brbd_status st = launch(d, ptrs, B);
    if (st != BRBD_OK) return st;
    CUDA_TRY(cudaEventRecord(d.ev1, d.s()));
    if (!(flags & BRBD_ASYNC))
    {
      CUDA_TRY(cudaStreamSynchronize(d.s()));
      float ms = 0.f;
      CUDA_TRY(cudaEventElapsedTime(&ms, d.ev0, d.ev1));
      p->last_ms = ms;
    }
    return BRBD_OK;
  }
  // host pointers: shard columns contiguously over the devices; every shard moves through the device in
  // chunks on three streams (upload | kernels | download) with double-buffered staging
  const int nd = (int)p->devs.size();
  const int64_t per = (B + nd - 1) / nd;
  if (args.size() > 8) return fail(BRBD_EINVAL, "too many arguments");
  size_t bytes_per_col = 0;
  for (const Arg & a : args)
    if (a.in || a.out) bytes_per_col += (size_t)a.rows * sizeof(T);
  // chunk: about 48 MB of traffic, a multiple of 1024 columns, at least 4096 columns
  int64_t chunk = (int64_t)((48u << 20) / std::max<size_t>(bytes_per_col, 1));
  chunk = std::max<int64_t>(4096, (chunk / 1024) * 1024);
  const std::vector<Arg> saved = args;
  brbd_status result = BRBD_OK;
  for (int g = 0; g < nd && result == BRBD_OK; ++g)
  {
    const int64_t c0 = (int64_t)g * per, c1 = std::min<int64_t>(B, c0 + per);
    if (c0 >= c1) break;
    DeviceCtx & d = p->devs[g];
    CUDA_TRY(cudaSetDevice(d.dev));
    const bool user = d.use_user_stream;
    d.use_user_stream = false; // kernels of host-pointer calls run on the pool's own stream
    const int64_t cw = std::min<int64_t>(chunk, c1 - c0);
    for (size_t k = 0; k < args.size(); ++k)
      if (args[k].in || args[k].out)
        for (int b = 0; b < 2; ++b)
        {
          brbd_status st = ensure_stage(d, (int)(2 * k + b), (size_t)args[k].rows * cw * sizeof(T));
          if (st != BRBD_OK) { d.use_user_stream = user; return st; }
        }
    int it = 0;
    for (int64_t b0 = c0; b0 < c1; b0 += cw, ++it)
    {
      const int buf = it & 1;
      const int64_t nb = std::min<int64_t>(cw, c1 - b0);
      std::vector<void *> ptrs(args.size(), nullptr);
      // upload may start once the kernel that read this input buffer two chunks ago is done
      if (it >= 2) CUDA_TRY(cudaStreamWaitEvent(d.s_in, d.ev_k[buf], 0));
      for (size_t k = 0; k < args.size(); ++k)
      {
        const Arg & a = saved[k];
        if (!a.in && !a.out) continue;
        ptrs[k] = d.stage[2 * k + buf];
        if (!a.in) continue;
        const T * src = static_cast<const T *>(a.in) + b0 * a.ld;
        // dense blocks (ld == rows, the Eigen::MatrixXd case) move as ONE copy: a pitched copy of many short
        // rows runs at a fraction of the link bandwidth
        if (a.ld == a.rows) CUDA_TRY(cudaMemcpyAsync(ptrs[k], src, (size_t)a.rows * nb * sizeof(T), cudaMemcpyHostToDevice, d.s_in));
        else
          CUDA_TRY(cudaMemcpy2DAsync(ptrs[k], a.rows * sizeof(T), src, a.ld * sizeof(T), a.rows * sizeof(T), nb,
                                     cudaMemcpyHostToDevice, d.s_in));
      }
      CUDA_TRY(cudaEventRecord(d.ev_in[buf], d.s_in));
      CUDA_TRY(cudaStreamWaitEvent(d.stream, d.ev_in[buf], 0));
      // the kernel overwrites the output buffer whose download was queued two chunks ago
      if (it >= 2) CUDA_TRY(cudaStreamWaitEvent(d.stream, d.ev_out[buf], 0));
      for (size_t k = 0; k < args.size(); ++k) args[k].ld = args[k].rows; // staged blocks are dense
      result = launch(d, ptrs, nb);
      args = saved;
      if (result != BRBD_OK) break;
      CUDA_TRY(cudaEventRecord(d.ev_k[buf], d.stream));
      CUDA_TRY(cudaStreamWaitEvent(d.s_out, d.ev_k[buf], 0));
      for (size_t k = 0; k < args.size(); ++k)
      {
        const Arg & a = saved[k];
        if (!a.out) continue;
        T * dst = static_cast<T *>(a.out) + b0 * a.ld;
        if (a.ld == a.rows) CUDA_TRY(cudaMemcpyAsync(dst, ptrs[k], (size_t)a.rows * nb * sizeof(T), cudaMemcpyDeviceToHost, d.s_out));
        else
          CUDA_TRY(cudaMemcpy2DAsync(dst, a.ld * sizeof(T), ptrs[k], a.rows * sizeof(T), a.rows * sizeof(T), nb,
                                     cudaMemcpyDeviceToHost, d.s_out));
      }
      CUDA_TRY(cudaEventRecord(d.ev_out[buf], d.s_out));
    }
    d.use_user_stream = user;
  }
  for (int g = 0; g < nd; ++g)
  {
    CUDA_TRY(cudaSetDevice(p->devs[g].dev));
    CUDA_TRY(cudaStreamSynchronize(p->devs[g].s_in));
    CUDA_TRY(cudaStreamSynchronize(p->devs[g].stream));
    CUDA_TRY(cudaStreamSynchronize(p->devs[g].s_out));
  }
  return result;
}

__global__ void fp64_peak_kernel(double * out, int iters)
{
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double b = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i)
  {
    a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
    a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
} // namespace

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

const char * brbd_last_error_string(void) { return g_err.c_str(); }
const char * brbd_version(void) { return "pinocchio_b200 0.1 (sm_100a)"; }
int brbd_device_count(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

brbd_status brbd_model_create(const brbd_flat_model * f, brbd_model ** out)
{
  if (!f || !out) return fail(BRBD_EINVAL, "null argument");
  *out = nullptr;
  brbd_model * m = new brbd_model();
  std::string err;
  const brbd_status bst = build_model_pod(f, m->pd, err);
  if (bst != BRBD_OK)
  {
    delete m;
    return fail(bst, err);
  }
  const ModelPOD<double> & P = m->pd;
  fill_pod(m->pf, P);
  build_tree(P, m->td);
  build_tree(P, m->tf);
  build_coop_tables(P, m->coop);
  if (m->td.maxpathdof > MAXPATH)
  {
    delete m;
    return fail(BRBD_ETOPOLOGY, "more than " + std::to_string(MAXPATH) + " degrees of freedom on one root path");
  }
  *out = m;
  return BRBD_OK;
}
void brbd_model_destroy(brbd_model * m) { delete m; }
int brbd_model_nq(const brbd_model * m) { return m ? m->pd.nq : -1; }
int brbd_model_nv(const brbd_model * m) { return m ? m->pd.nv : -1; }
int brbd_model_njoints(const brbd_model * m) { return m ? m->pd.njoints : -1; }

static brbd_status upload_model(brbd_pool * p)
{
  for (DeviceCtx & d : p->devs)
  {
    CUDA_TRY(cudaSetDevice(d.dev));
    CUDA_TRY(cudaMemcpy(d.d_pd, &p->model.pd, sizeof(ModelPOD<double>), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(d.d_pf, &p->model.pf, sizeof(ModelPOD<float>), cudaMemcpyHostToDevice));
  }
  return BRBD_OK;
}

void brbd_pool_destroy(brbd_pool * p)
{
  if (!p) return;
  for (DeviceCtx & d : p->devs)
  {
    if (cudaSetDevice(d.dev) != cudaSuccess) continue;
    if (d.stream) cudaStreamSynchronize(d.stream);
    for (int k = 0; k < 16; ++k) if (d.stage[k]) cudaFree(d.stage[k]);
    for (int b = 0; b < 2; ++b)
    {
      if (d.ev_in[b]) cudaEventDestroy(d.ev_in[b]);
      if (d.ev_k[b]) cudaEventDestroy(d.ev_k[b]);
      if (d.ev_out[b]) cudaEventDestroy(d.ev_out[b]);
    }
    if (d.s_in) cudaStreamDestroy(d.s_in);
    if (d.s_out) cudaStreamDestroy(d.s_out);
    if (d.work) cudaFree(d.work);
    if (d.aux) cudaFree(d.aux);
    if (d.zeros) cudaFree(d.zeros);
    if (d.d_pd) cudaFree(d.d_pd);
    if (d.d_pf) cudaFree(d.d_pf);
    if (d.ev0) cudaEventDestroy(d.ev0);
    if (d.ev1) cudaEventDestroy(d.ev1);
    if (d.stream) cudaStreamDestroy(d.stream);
  }
  delete p;
}

brbd_status brbd_pool_create(const brbd_model * m, const int * device_ids, int n_devices, brbd_pool ** out)
{
  if (!m || !out) return fail(BRBD_EINVAL, "null argument");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(BRBD_ECUDA, std::string("no CUDA device available (there is no CPU fallback): ") + cudaGetErrorString(e));
  std::vector<int> ids;
  if (!device_ids || n_devices <= 0) ids.push_back(0);
  else ids.assign(device_ids, device_ids + n_devices);
  brbd_pool * p = new brbd_pool();
  p->model = *m;
  for (int id : ids)
  {
    if (id < 0 || id >= ndev)
    {
      brbd_pool_destroy(p);
      return fail(BRBD_EINVAL, "device id " + std::to_string(id) + " out of range");
    }
    DeviceCtx d;
    d.dev = id;
    p->devs.push_back(d);
  }
  for (DeviceCtx & d : p->devs)
  {
    cudaDeviceProp prop;
    brbd_status st = BRBD_OK;
    auto tryc = [&](cudaError_t err, const char * what) {
      if (err != cudaSuccess && st == BRBD_OK) st = fail(BRBD_ECUDA, std::string(what) + ": " + cudaGetErrorString(err));
    };
    tryc(cudaSetDevice(d.dev), "cudaSetDevice");
    tryc(cudaGetDeviceProperties(&prop, d.dev), "cudaGetDeviceProperties");
    if (st == BRBD_OK)
    {
      d.sm_count = prop.multiProcessorCount;
      d.max_smem_optin = (int)prop.sharedMemPerBlockOptin;
      tryc(cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking), "cudaStreamCreate");
      tryc(cudaStreamCreateWithFlags(&d.s_in, cudaStreamNonBlocking), "cudaStreamCreate");
      tryc(cudaStreamCreateWithFlags(&d.s_out, cudaStreamNonBlocking), "cudaStreamCreate");
      for (int b = 0; b < 2; ++b)
      {
        tryc(cudaEventCreateWithFlags(&d.ev_in[b], cudaEventDisableTiming), "cudaEventCreate");
        tryc(cudaEventCreateWithFlags(&d.ev_k[b], cudaEventDisableTiming), "cudaEventCreate");
        tryc(cudaEventCreateWithFlags(&d.ev_out[b], cudaEventDisableTiming), "cudaEventCreate");
      }
      tryc(cudaEventCreate(&d.ev0), "cudaEventCreate");
      tryc(cudaEventCreate(&d.ev1), "cudaEventCreate");
      tryc(cudaMalloc(&d.d_pd, sizeof(ModelPOD<double>)), "cudaMalloc");
      tryc(cudaMalloc(&d.d_pf, sizeof(ModelPOD<float>)), "cudaMalloc");
      tryc(cudaMalloc(&d.zeros, MAXNV * sizeof(double)), "cudaMalloc");
      if (st == BRBD_OK) tryc(cudaMemset(d.zeros, 0, MAXNV * sizeof(double)), "cudaMemset");
    }
    if (st != BRBD_OK)
    {
      brbd_pool_destroy(p);
      return st;
    }
  }
  brbd_status st = upload_model(p);
  if (st != BRBD_OK)
  {
    brbd_pool_destroy(p);
    return st;
  }
  *out = p;
  return BRBD_OK;
}
int brbd_pool_size(const brbd_pool * p) { return p ? (int)p->devs.size() : 0; }
brbd_status brbd_pool_update(brbd_pool * p, const brbd_model * m)
{
  if (!p || !m) return fail(BRBD_EINVAL, "null argument");
  brbd_status st = brbd_pool_synchronize(p);
  if (st != BRBD_OK) return st;
  p->model = *m;
  return upload_model(p);
}
brbd_status brbd_pool_set_stream(brbd_pool * p, void * cuda_stream)
{
  if (!p) return fail(BRBD_EINVAL, "null pool");
  if (p->devs.size() != 1) return fail(BRBD_EINVAL, "external streams require a single-device pool");
  p->devs[0].user_stream = static_cast<cudaStream_t>(cuda_stream);
  p->devs[0].use_user_stream = true;
  return BRBD_OK;
}
brbd_status brbd_pool_synchronize(brbd_pool * p)
{
  if (!p) return fail(BRBD_EINVAL, "null pool");
  for (DeviceCtx & d : p->devs)
  {
    CUDA_TRY(cudaSetDevice(d.dev));
    CUDA_TRY(cudaStreamSynchronize(d.s()));
    if (d.use_user_stream) CUDA_TRY(cudaStreamSynchronize(d.stream));
  }
  return BRBD_OK;
}
int64_t brbd_pool_launch_count(const brbd_pool * p) { return p ? p->launches : 0; }
double brbd_pool_last_kernel_ms(const brbd_pool * p) { return p ? p->last_ms : 0.0; }

#define DISPATCH(flags, CALL)                              \
  if ((flags)&BRBD_FP32) { typedef float T; return CALL; } \
  else { typedef double T; return CALL; }

brbd_status brbd_rnea_batch(brbd_pool * p, const void * q, int64_t ldq, const void * v, int64_t ldv, const void * a,
                            int64_t lda, void * tau, int64_t ldtau, int64_t batch, int flags)
{
  if (!p) return fail(BRBD_EINVAL, "null pool");
  const int nq = p->model.pd.nq, nv = p->model.pd.nv;
  std::vector<Arg> args = {{q, nullptr, ldq, nq, false}, {v, nullptr, ldv, nv, false}, {a, nullptr, lda, nv, false}, {nullptr, tau, ldtau, nv, false}};
  DISPATCH(flags, (run_call<T>(p, args, batch, flags, [&](DeviceCtx & d, std::vector<void *> & P, int64_t B) {
             return launch_rnea<T>(p, d, (const T *)P[0], args[0].ld, (const T *)P[1], args[1].ld, (const T *)P[2], args[2].ld,
                                   (T *)P[3], args[3].ld, B);
           })));
}

brbd_status brbd_aba_batch(brbd_pool * p, const void * q, int64_t ldq, const void * v, int64_t ldv, const void * tau,
                           int64_t ldtau, void * a, int64_t lda, int64_t batch, int flags)
{
  if (!p) return fail(BRBD_EINVAL, "null pool");
  const int nq = p->model.pd.nq, nv = p->model.pd.nv;
  std::vector<Arg> args = {{q, nullptr, ldq, nq, false}, {v, nullptr, ldv, nv, false}, {tau, nullptr, ldtau, nv, false}, {nullptr, a, lda, nv, false}};
  DISPATCH(flags, (run_call<T>(p, args, batch, flags, [&](DeviceCtx & d, std::vector<void *> & P, int64_t B) {
             return launch_aba<T>(p, d, (const T *)P[0], args[0].ld, (const T *)P[1], args[1].ld, (const T *)P[2], args[2].ld,
                                  (T *)P[3], args[3].ld, B);
           })));
}

brbd_status brbd_crba_batch(brbd_pool * p, const void * q, int64_t ldq, void * M, int64_t ldM, int64_t batch, int flags)
{
  if (!p) return fail(BRBD_EINVAL, "null pool");
  const int nq = p->model.pd.nq, nv = p->model.pd.nv;
  std::vector<Arg> args = {{q, nullptr, ldq, nq, false}, {nullptr, M, ldM, (int64_t)nv * nv, false}};
  DISPATCH(flags, (run_call<T>(p, args, batch, flags, [&](DeviceCtx & d, std::vector<void *> & P, int64_t B) {
             return launch_crba<T>(p, d, (const T *)P[0], args[0].ld, (T *)P[1], args[1].ld, B);
           })));
}

brbd_status brbd_rnea_derivatives_batch(brbd_pool * p, const void * q, int64_t ldq, const void * v, int64_t ldv,
                                        const void * a, int64_t lda, void * dtau_dq, int64_t ld_dq, void * dtau_dv,
                                        int64_t ld_dv, void * dtau_da, int64_t ld_da, void * tau, int64_t ldtau,
                                        int64_t batch, int flags)
{
  if (!p) return fail(BRBD_EINVAL, "null pool");
  const int nq = p->model.pd.nq, nv = p->model.pd.nv;
  const int64_t nn = (int64_t)nv * nv;
  std::vector<Arg> args = {{q, nullptr, ldq, nq, false},   {v, nullptr, ldv, nv, false},     {a, nullptr, lda, nv, false},
                           {nullptr, dtau_dq, ld_dq, nn, false}, {nullptr, dtau_dv, ld_dv, nn, false}, {nullptr, dtau_da, ld_da, nn, false},
                           {nullptr, tau, ldtau, nv, true}};
  DISPATCH(flags, (run_call<T>(p, args, batch, flags, [&](DeviceCtx & d, std::vector<void *> & P, int64_t B) {
             return launch_rnea_derivs<T>(p, d, (const T *)P[0], args[0].ld, (const T *)P[1], args[1].ld, (const T *)P[2],
                                          args[2].ld, (T *)P[3], args[3].ld, (T *)P[4], args[4].ld, (T *)P[5], args[5].ld,
                                          (T *)P[6], args[6].ld, B);
           })));
}

brbd_status brbd_aba_derivatives_batch(brbd_pool * p, const void * q, int64_t ldq, const void * v, int64_t ldv,
                                       const void * tau, int64_t ldtau, void * ddq_dq, int64_t ld_dq, void * ddq_dv,
                                       int64_t ld_dv, void * ddq_dtau, int64_t ld_dtau, void * ddq, int64_t ldddq,
                                       int64_t batch, int flags)
{
  if (!p) return fail(BRBD_EINVAL, "null pool");
  const int nq = p->model.pd.nq, nv = p->model.pd.nv;
  const int64_t nn = (int64_t)nv * nv;
  std::vector<Arg> args = {{q, nullptr, ldq, nq, false},   {v, nullptr, ldv, nv, false},     {tau, nullptr, ldtau, nv, false},
                           {nullptr, ddq_dq, ld_dq, nn, false}, {nullptr, ddq_dv, ld_dv, nn, false}, {nullptr, ddq_dtau, ld_dtau, nn, false},
                           {nullptr, ddq, ldddq, nv, true}};
  DISPATCH(flags, (run_call<T>(p, args, batch, flags, [&](DeviceCtx & d, std::vector<void *> & P, int64_t B) {
             return launch_aba_derivs<T>(p, d, (const T *)P[0], args[0].ld, (const T *)P[1], args[1].ld, (const T *)P[2],
                                         args[2].ld, (T *)P[3], args[3].ld, (T *)P[4], args[4].ld, (T *)P[5], args[5].ld,
                                         (T *)P[6], args[6].ld, B);
           })));
}

brbd_status brbd_nle_batch(brbd_pool * p, const void * q, int64_t ldq, const void * v, int64_t ldv, void * nle, int64_t ldn,
                           int64_t batch, int flags)
{
  if (!p) return fail(BRBD_EINVAL, "null pool");
  const int nq = p->model.pd.nq, nv = p->model.pd.nv;
  std::vector<Arg> args = {{q, nullptr, ldq, nq, false}, {v, nullptr, ldv, nv, false}, {nullptr, nle, ldn, nv, false}};
  DISPATCH(flags, (run_call<T>(p, args, batch, flags, [&](DeviceCtx & d, std::vector<void *> & P, int64_t B) {
             return launch_rnea<T>(p, d, (const T *)P[0], args[0].ld, (const T *)P[1], args[1].ld, (const T *)d.zeros, 0,
                                   (T *)P[2], args[2].ld, B);
           })));
}

brbd_status brbd_gravity_batch(brbd_pool * p, const void * q, int64_t ldq, void * g, int64_t ldg, int64_t batch, int flags)
{
  if (!p) return fail(BRBD_EINVAL, "null pool");
  const int nq = p->model.pd.nq, nv = p->model.pd.nv;
  std::vector<Arg> args = {{q, nullptr, ldq, nq, false}, {nullptr, g, ldg, nv, false}};
  DISPATCH(flags, (run_call<T>(p, args, batch, flags, [&](DeviceCtx & d, std::vector<void *> & P, int64_t B) {
             return launch_rnea<T>(p, d, (const T *)P[0], args[0].ld, (const T *)d.zeros, 0, (const T *)d.zeros, 0, (T *)P[1],
                                   args[1].ld, B);
           })));
}

brbd_status brbd_minverse_batch(brbd_pool * p, const void * q, int64_t ldq, void * Minv, int64_t ldM, int64_t batch, int flags)
{
  if (!p) return fail(BRBD_EINVAL, "null pool");
  const int nq = p->model.pd.nq, nv = p->model.pd.nv;
  std::vector<Arg> args = {{q, nullptr, ldq, nq, false}, {nullptr, Minv, ldM, (int64_t)nv * nv, false}};
  DISPATCH(flags, (run_call<T>(p, args, batch, flags, [&](DeviceCtx & d, std::vector<void *> & P, int64_t B) {
             return launch_minverse<T>(p, d, (const T *)P[0], args[0].ld, (T *)P[1], args[1].ld, B);
           })));
}

brbd_status brbd_integrate_batch(brbd_pool * p, const void * q, int64_t ldq, const void * v, int64_t ldv, void * qout, int64_t ldqo,
                                 int64_t batch, int flags)
{
  if (!p) return fail(BRBD_EINVAL, "null pool");
  const int nq = p->model.pd.nq, nv = p->model.pd.nv;
  std::vector<Arg> args = {{q, nullptr, ldq, nq, false}, {v, nullptr, ldv, nv, false}, {nullptr, qout, ldqo, nq, false}};
  DISPATCH(flags, (run_call<T>(p, args, batch, flags, [&](DeviceCtx & d, std::vector<void *> & P, int64_t B) {
             return launch_integrate<T, false>(p, d, (const T *)P[0], args[0].ld, (const T *)P[1], args[1].ld, (const T *)nullptr, 0,
                                               T(1), (T *)P[2], args[2].ld, (T *)nullptr, 0, B);
           })));
}

brbd_status brbd_aba_euler_step_batch(brbd_pool * p, const void * q, int64_t ldq, const void * v, int64_t ldv, const void * tau,
                                      int64_t ldtau, double dt, void * q_next, int64_t ldqn, void * v_next, int64_t ldvn,
                                      int64_t batch, int flags)
{
  if (!p) return fail(BRBD_EINVAL, "null pool");
  const int nq = p->model.pd.nq, nv = p->model.pd.nv;
  std::vector<Arg> args = {{q, nullptr, ldq, nq, false},          {v, nullptr, ldv, nv, false},         {tau, nullptr, ldtau, nv, false},
                           {nullptr, q_next, ldqn, nq, false}, {nullptr, v_next, ldvn, nv, false}};
  DISPATCH(flags, (run_call<T>(p, args, batch, flags, [&](DeviceCtx & d, std::vector<void *> & P, int64_t B) {
             brbd_status st = ensure_aux(d, (size_t)B * nv * sizeof(T));
             if (st != BRBD_OK) return st;
             st = launch_aba<T>(p, d, (const T *)P[0], args[0].ld, (const T *)P[1], args[1].ld, (const T *)P[2], args[2].ld,
                                (T *)d.aux, nv, B);
             if (st != BRBD_OK) return st;
             return launch_integrate<T, true>(p, d, (const T *)P[0], args[0].ld, (const T *)P[1], args[1].ld, (const T *)d.aux, nv,
                                              (T)dt, (T *)P[3], args[3].ld, (T *)P[4], args[4].ld, B);
           })));
}

brbd_status brbd_host_register(void * ptr, uint64_t bytes)
{
  if (!ptr || bytes == 0) return fail(BRBD_EINVAL, "null host block");
  CUDA_TRY(cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterPortable));
  return BRBD_OK;
}
brbd_status brbd_host_unregister(void * ptr)
{
  if (!ptr) return fail(BRBD_EINVAL, "null host block");
  CUDA_TRY(cudaHostUnregister(ptr));
  return BRBD_OK;
}

brbd_status brbd_measure_fp64_peak(brbd_pool * p, double * flops_per_s, double * elapsed_ms)
{
  if (!p || p->devs.empty()) return fail(BRBD_EINVAL, "null pool");
  DeviceCtx & d = p->devs[0];
  CUDA_TRY(cudaSetDevice(d.dev));
  const int threads = 256, blocks = d.sm_count * 8, iters = 1 << 16;
  brbd_status st = ensure_work(d, (size_t)threads * blocks * sizeof(double));
  if (st != BRBD_OK) return st;
  fp64_peak_kernel<<<blocks, threads, 0, d.stream>>>((double *)d.work, 1 << 10); // warm-up
  CUDA_TRY(cudaEventRecord(d.ev0, d.stream));
  fp64_peak_kernel<<<blocks, threads, 0, d.stream>>>((double *)d.work, iters);
  CUDA_TRY(cudaEventRecord(d.ev1, d.stream));
  CUDA_TRY(cudaStreamSynchronize(d.stream));
  p->launches += 2;
  float ms = 0.f;
  CUDA_TRY(cudaEventElapsedTime(&ms, d.ev0, d.ev1));
  const double flops = 2.0 * 8.0 * (double)iters * threads * blocks;
  if (flops_per_s) *flops_per_s = flops / (ms * 1e-3);
  if (elapsed_ms) *elapsed_ms = ms;
  return BRBD_OK;
}

} // extern "C"
