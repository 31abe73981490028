// host_ctx.hpp — host-side state shared by the translation units of the engine: the validated model, the per-device
// context of a pool (streams, staging, workspaces) and the launch-geometry helpers.  Every algorithm's launch code sits in
// its own launch_*.cu so that the library builds in parallel; the C ABI itself is capi.cu.
#pragma once
#include <cuda.h>
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
#include "engine.cuh"
#include "tree.cuh"
#include "deriv_coop.cuh"
#include "aba_deriv_coop.cuh"

namespace brbd
{
brbd_status fail(brbd_status s, const std::string & msg); // sets brbd_last_error_string() of the calling thread (capi.cu)
}
#define CUDA_TRY(expr)                                                                              \
  do                                                                                                \
  {                                                                                                 \
    cudaError_t e__ = (expr);                                                                       \
    if (e__ != cudaSuccess)                                                                         \
      return brbd::fail(BRBD_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));           \
  } while (0)

struct brbd_model
{
  brbd::ModelPOD<double> pd;
  brbd::ModelPOD<float> pf;
  brbd::TreePOD<double> td; // v2 kernels: passed by value as a __grid_constant__ kernel parameter
  brbd::TreePOD<float> tf;
  brbd::CoopTables coop; // warp-cooperative derivative kernels: level lists, ancestor masks
  // the model as the caller gave it (brbd_model_get_flat)
  std::vector<int32_t> f_parents, f_type, f_idx_q, f_idx_v;
  std::vector<double> f_placement, f_inertia, f_armature, f_axis;
  double f_gravity[3] = {0, 0, 0};
};

namespace brbd
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
} // namespace brbd

namespace brbd
{
// kernels generated for the pool's model (codegen.cu) and compiled at brbd_pool_specialize (launch_gen.cu)
struct GenKernel
{
  void * lib = nullptr;    // cudaLibrary_t
  void * kernel = nullptr; // cudaKernel_t
  void * kernel_tma = nullptr; // CRBA: the variant whose column blocks leave through TMA tensor stores
  bool compact = false;        // CRBA: compact staging (dense and packed output modes, see codegen.cu)
  int nnz = 0;                 // CRBA: entries of the structural pattern (rows of the packed result)
  int nt = 0, nrec = 0;
  size_t smem_bytes = 0;
};
// One algorithm / precision: up to three variants that differ in threads per CTA.  The generated code is hundreds of KB of
// straight-line instructions and runs at the rate the SM can FETCH them, whatever the number of resident warps (measured:
// scripts/gen_fetch_probe.py) — so a pass of the persistent grid costs ~ (c0 + warps) and the launch picks the variant with
// the cheapest rounds * (c0 + warps) for the batch at hand.
struct GenSet
{
  GenKernel var[3];
  int nvar = 0;
};
} // namespace brbd

struct brbd_pool
{
  brbd::GenSet gen[5][2]; // [BRBD_GEN_*][fp64, fp32]
  brbd::GenKernel crba_packed[2]; // the generated CRBA with compact staging (brbd_crba_packed_batch), built at its first call
  // host-pointer crba with packed transfer (brbd_pool_set_host_threads): position of every packed entry inside a matrix, pinned
  // landing buffers for two chunks in flight
  int host_threads = 0;
  std::vector<int32_t> crba_idx;
  void * host_stage[2] = {nullptr, nullptr};
  size_t host_stage_bytes[2] = {0, 0};
  bool packed_unavailable = false; // the compact-staging kernel could not be built (no NVRTC): dense transfers
  int64_t gen_min_batch = -1; // batches at least this large use a specialised kernel when there is one; -1: per algorithm (use_generated)
  brbd_model model;
  std::vector<brbd::DeviceCtx> devs;
  int64_t launches = 0;
  double last_ms = 0.0;
};

namespace brbd
{
inline brbd_status ensure_stage(DeviceCtx & d, int slot, size_t bytes)
{
  if (d.stage_bytes[slot] >= bytes) return BRBD_OK;
  if (d.stage[slot]) CUDA_TRY(cudaFree(d.stage[slot]));
  d.stage[slot] = nullptr;
  d.stage_bytes[slot] = 0;
  CUDA_TRY(cudaMalloc(&d.stage[slot], bytes));
  d.stage_bytes[slot] = bytes;
  return BRBD_OK;
}
inline brbd_status ensure_work(DeviceCtx & d, size_t bytes)
{
  if (d.work_bytes >= bytes) return BRBD_OK;
  if (d.work) CUDA_TRY(cudaFree(d.work));
  d.work = nullptr;
  d.work_bytes = 0;
  CUDA_TRY(cudaMalloc(&d.work, bytes));
  d.work_bytes = bytes;
  return BRBD_OK;
}

inline brbd_status ensure_aux(DeviceCtx & d, size_t bytes)
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
inline Geometry pick_geometry(const DeviceCtx & d, size_t per_warp, size_t static_bytes, int64_t batch, int max_warps_per_cta,
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
template<> inline const ModelPOD<double> * dev_model<double>(const DeviceCtx & d) { return d.d_pd; }
template<> inline const ModelPOD<float> * dev_model<float>(const DeviceCtx & d) { return d.d_pf; }

// ------------------------------------------------------------------------------------------------
// Device-pointer launches (one device)
// ------------------------------------------------------------------------------------------------
template<class T> const TreePOD<T> & tree_of(const brbd_pool * p);
template<> inline const TreePOD<double> & tree_of<double>(const brbd_pool * p) { return p->model.td; }
template<> inline const TreePOD<float> & tree_of<float>(const brbd_pool * p) { return p->model.tf; }

// Launch geometry of the v2 (DFS-interleaved) kernels: `state_bytes` of shared memory per thread plus
// `warp_bytes` per warp; as many warps per CTA as fit (<= max_warps), as many CTAs per SM as fit
// (<= max_ctas), persistent grid.
struct Geometry2
{
  int warps, ctas_per_sm, grid;
  size_t dyn_bytes;
};
inline Geometry2 pick_geometry2(const DeviceCtx & d, size_t state_bytes, size_t warp_bytes, int64_t batch, int max_warps, int max_ctas)
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
inline GeometryCoop pick_geometry_coop(const DeviceCtx & d, size_t group_bytes, int G, size_t static_bytes, int64_t batch)
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

// ---- TMA tensor maps over the caller's (nv*nv x B, leading dimension ldM) matrix block: see crba_tma_kernel ----------
typedef CUresult (*brbd_encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                         const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                         CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline brbd_encode_tiled_fn encode_tiled_fn()
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


// ---- one launch function per algorithm, each instantiated for double and float in its own translation unit ----------
template<class T>
brbd_status launch_rnea(brbd_pool * p, DeviceCtx & d, const T * q, int64_t ldq, const T * v, int64_t ldv, const T * a,
                        int64_t lda, T * tau, int64_t ldtau, int64_t B);
template<class T>
brbd_status launch_aba(brbd_pool * p, DeviceCtx & d, const T * q, int64_t ldq, const T * v, int64_t ldv, const T * tau,
                       int64_t ldtau, T * a, int64_t lda, int64_t B);
template<class T>
brbd_status launch_aba_coop(brbd_pool * p, DeviceCtx & d, const T * q, int64_t ldq, const T * v, int64_t ldv, const T * tau,
                            int64_t ldtau, T * a, int64_t lda, int64_t B, bool * done);
template<class T>
brbd_status launch_crba(brbd_pool * p, DeviceCtx & d, const T * q, int64_t ldq, T * Mout, int64_t ldM, int64_t B);
template<class T>
brbd_status launch_rnea_derivs(brbd_pool * p, DeviceCtx & d, const T * q, int64_t ldq, const T * v, int64_t ldv,
                               const T * a, int64_t lda, T * dq, int64_t ld_dq, T * dv, int64_t ld_dv, T * da,
                               int64_t ld_da, T * tau, int64_t ldtau, int64_t B);
template<class T>
brbd_status launch_aba_derivs(brbd_pool * p, DeviceCtx & d, const T * q, int64_t ldq, const T * v, int64_t ldv,
                              const T * tau, int64_t ldtau, T * dq, int64_t ld_dq, T * dv, int64_t ld_dv, T * dtau,
                              int64_t ld_dtau, T * ddq, int64_t ldddq, int64_t B);
template<class T>
brbd_status launch_minverse(brbd_pool * p, DeviceCtx & d, const T * q, int64_t ldq, T * Minv, int64_t ldM, int64_t B);
template<class T, bool EULER>
brbd_status launch_integrate(brbd_pool * p, DeviceCtx & d, const T * q, int64_t ldq, const T * v, int64_t ldv, const T * a,
                             int64_t lda, T dt, T * qout, int64_t ldqo, T * vout, int64_t ldvo, int64_t B);
// generic thread-local-state kernels (launch_v1.cu): fit any accepted model
template<class T>
brbd_status launch_rnea_v1(brbd_pool * p, DeviceCtx & d, const T * q, int64_t ldq, const T * v, int64_t ldv, const T * a,
                           int64_t lda, T * tau, int64_t ldtau, int64_t B);
template<class T>
brbd_status launch_aba_v1(brbd_pool * p, DeviceCtx & d, const T * q, int64_t ldq, const T * v, int64_t ldv, const T * tau,
                          int64_t ldtau, T * a, int64_t lda, int64_t B);
template<class T>
brbd_status launch_crba_v1(brbd_pool * p, DeviceCtx & d, const T * q, int64_t ldq, T * Mout, int64_t ldM, int64_t B);
template<class T>
brbd_status launch_rnea_derivs_v1(brbd_pool * p, DeviceCtx & d, const T * q, int64_t ldq, const T * v, int64_t ldv,
                                  const T * a, int64_t lda, T * dq, int64_t ld_dq, T * dv, int64_t ld_dv, T * da,
                                  int64_t ld_da, T * tau, int64_t ldtau, int64_t B);
brbd_status measure_fp64_peak(brbd_pool * p, double * flops_per_s, double * elapsed_ms);
// specialised kernels (launch_gen.cu)
brbd_status specialize_one(brbd_pool * p, int algo, bool fp32, int flags);
void release_generated(brbd_pool * p);
template<class T>
brbd_status launch_generated(brbd_pool * p, DeviceCtx & d, int algo, const T * q, int64_t ldq, const T * v, int64_t ldv, const T * x,
                             int64_t ldx, T * out, int64_t ldo, int64_t B);
// packed CRBA through the generated kernel with compact staging (specialises CRBA first if need be)
template<class T>
brbd_status launch_crba_packed(brbd_pool * p, DeviceCtx & d, const T * q, int64_t ldq, T * P, int64_t ldP, int64_t B);
int crba_pattern_nnz(const brbd_model & m);
void crba_pattern_index(const brbd_model & m, std::vector<int32_t> & idx); // idx[k] = cols[k] * nv + rows[k]
// host_expand.cpp
template<class T>
void expand_packed(T * dst, int64_t ld, const T * src, int64_t nnz, const int32_t * idx, int nn, int64_t count, int threads, int64_t src_ld = 0);
// generated computeRNEADerivatives / computeABADerivatives (small models): q, v, x -> three nv*nv blocks + an nv block
template<class T>
brbd_status launch_generated_derivs(brbd_pool * p, DeviceCtx & d, int algo, const T * q, int64_t ldq, const T * v, int64_t ldv, const T * x,
                                    int64_t ldx, T * o0, int64_t ld0, T * o1, int64_t ld1, T * o2, int64_t ld2, T * o3, int64_t ld3, int64_t B);
// generated CRBA, bulk-copy variant: elements between the staging rows of two lanes for `group` columns (room for the alignment
// shift, 16-byte aligned rows, an odd number of 16-byte units so that the lanes' row writes spread over the banks)
inline int crba_bulk_pitch(int nv, int group, bool fp32)
{
  const int A = fp32 ? 4 : 2;
  int pitch = (nv * group + (A - 1) + (A - 1)) / A * A;
  if ((pitch / A) % 2 == 0) pitch += A;
  return pitch;
}
// From which batch size a specialised pool runs the generated kernel.  brbd_pool_set_specialized_min_batch overrides it; by default
// (-1) per algorithm, from the measured crossover against the small-batch paths (profiles/r2_gen_small_batch.txt, time per
// device-resident call): CRBA — always (35-dof humanoid: 34 us against 71 us from 128 configurations on); RNEA — from 4096
// (24-38 us flat up to 16 384 against the cooperative kernel's 20 us up to 1024 and the generic thread kernel's 46-48 us from
// 4096); ABA — from 8192 (86-117 us against 36-83 us below); a model of at most 8 dofs — always (6-dof arm: 17 / 20 / 14 us
// against 19 / 24 / 20 us).
template<class T> inline bool use_generated(const brbd_pool * p, int algo, int64_t B)
{
  if (p->gen[algo][sizeof(T) == 4 ? 1 : 0].nvar == 0) return false;
  if (p->gen_min_batch >= 0) return B >= p->gen_min_batch;
  if (algo == BRBD_GEN_CRBA) return true;
  if (p->model.pd.nv <= 8 && algo <= BRBD_GEN_ABA) return true;
  return B >= (algo == BRBD_GEN_RNEA ? 4096 : 8192);
}
// BRBD_<ALGO>_V=<name> forces one device path of an algorithm (tests, experiments)
inline bool forced_path(const char * var, const char * name)
{
  const char * e = std::getenv(var);
  return e && std::strcmp(e, name) == 0;
}
} // namespace brbd
