// launch_gen.cu — pool specialisation: the generated source of codegen.cu is compiled with NVRTC for sm_100a, loaded with
// the runtime's library API and launched in place of the generic kernels for large batches.
// NVRTC is opened lazily with dlopen: the engine loads and works without it (brbd_pool_specialize then fails with
// BRBD_ECUDA and the pool keeps its generic kernels).
#include <dlfcn.h>

#include "host_ctx.hpp"

namespace brbd
{
namespace
{
struct Nvrtc
{
  void * h = nullptr;
  int (*CreateProgram)(void **, const char *, const char *, int, const char * const *, const char * const *) = nullptr;
  int (*CompileProgram)(void *, int, const char * const *) = nullptr;
  int (*GetCUBINSize)(void *, size_t *) = nullptr;
  int (*GetCUBIN)(void *, char *) = nullptr;
  int (*GetProgramLogSize)(void *, size_t *) = nullptr;
  int (*GetProgramLog)(void *, char *) = nullptr;
  int (*DestroyProgram)(void **) = nullptr;
  bool ok = false;
};
const Nvrtc & nvrtc()
{
  static Nvrtc n = [] {
    Nvrtc r;
    const char * names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so"};
    for (const char * nm : names)
      if ((r.h = dlopen(nm, RTLD_NOW | RTLD_LOCAL))) break;
    if (!r.h) return r;
#define BRBD_SYM(field, sym) *(void **)(&r.field) = dlsym(r.h, sym)
    BRBD_SYM(CreateProgram, "nvrtcCreateProgram");
    BRBD_SYM(CompileProgram, "nvrtcCompileProgram");
    BRBD_SYM(GetCUBINSize, "nvrtcGetCUBINSize");
    BRBD_SYM(GetCUBIN, "nvrtcGetCUBIN");
    BRBD_SYM(GetProgramLogSize, "nvrtcGetProgramLogSize");
    BRBD_SYM(GetProgramLog, "nvrtcGetProgramLog");
    BRBD_SYM(DestroyProgram, "nvrtcDestroyProgram");
#undef BRBD_SYM
    r.ok = r.CreateProgram && r.CompileProgram && r.GetCUBINSize && r.GetCUBIN && r.GetProgramLogSize && r.GetProgramLog && r.DestroyProgram;
    return r;
  }();
  return n;
}

brbd_status compile_cubin(const char * source, const char * name, std::vector<char> & cubin)
{
  const Nvrtc & N = nvrtc();
  if (!N.ok) return fail(BRBD_ECUDA, "specialisation needs NVRTC (libnvrtc.so.12), which could not be loaded");
  void * prog = nullptr;
  if (N.CreateProgram(&prog, source, name, 0, nullptr, nullptr) != 0) return fail(BRBD_ECUDA, "nvrtcCreateProgram failed");
  const char * opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo", "--fmad=true"};
  const int rc = N.CompileProgram(prog, 4, opts);
  if (rc != 0)
  {
    size_t n = 0;
    N.GetProgramLogSize(prog, &n);
    std::string log(n, '\0');
    if (n) N.GetProgramLog(prog, &log[0]);
    N.DestroyProgram(&prog);
    return fail(BRBD_ECUDA, "NVRTC compilation of the specialised kernel failed: " + log.substr(0, 2000));
  }
  size_t n = 0;
  N.GetCUBINSize(prog, &n);
  cubin.resize(n);
  N.GetCUBIN(prog, cubin.data());
  N.DestroyProgram(&prog);
  return BRBD_OK;
}
} // namespace

void release_generated(brbd_pool * p)
{
  for (int a = 0; a < 5; ++a)
    for (int f = 0; f < 2; ++f)
    {
      GenSet & g = p->gen[a][f];
      for (int k = 0; k < g.nvar; ++k)
        if (g.var[k].lib) cudaLibraryUnload((cudaLibrary_t)g.var[k].lib);
      g = GenSet();
    }
  for (int f = 0; f < 2; ++f)
  {
    if (p->crba_packed[f].lib) cudaLibraryUnload((cudaLibrary_t)p->crba_packed[f].lib);
    p->crba_packed[f] = GenKernel();
  }
  p->crba_idx.clear(); // the pattern belongs to the model the kernels were generated for
  p->packed_unavailable = false;
}

namespace
{
const char * kAlgoNames[] = {"rnea", "aba", "crba", "rnea_derivatives", "aba_derivatives"};

// generate + compile + load one variant; BRBD_OK with k.kernel == nullptr when it does not fit the SM (too much shared memory)
// crba_mode (CRBA only): 0 = default (BRBD_GEN_CRBA_MODE or the built-in choice), 1 = compact staging (packed output)
brbd_status build_variant(brbd_pool * p, int algo, bool fp32, int nt, bool direct, bool slots, GenKernel & k, int crba_mode = 0)
{
  char * src = nullptr;
  brbd_codegen_info info;
  // CRBA: adjacent columns per flush — as many as keep a staging row under ~1.5 KB (the whole matrix of a 6-dof arm).
  // Measured on the 35-dof humanoid (profiles/r2_gen_crba_experiments.txt): 1 / 3 / 5 columns 4.11 / 3.00 / - ms at 2^20.
  int group = 1, nbuf = 1;
  bool compact = false, bulk = false;
  if (algo == BRBD_GEN_CRBA)
  {
    const int nvm = p->model.pd.nv;
    // how the columns leave (codegen.cu): "bulk" — every lane hands `group` adjacent columns of its configuration to the copy
    // engine (asynchronous, one run of group * nv elements per copy); "lsu" — coalesced stores by the warp (+ the TMA
    // tensor-store kernel for single columns); "compact" — only the structural pattern staged (the packed result).
    // Measured (profiles/r2_crba_bulk_sweep.txt): the bulk variant is the faster the FEWER warps run and the MORE columns a copy
    // takes — 3 warps per SM with one staging row per lane as long as shared memory allows (8 columns of a 35-dof humanoid):
    // 65 536 configurations 0.225 -> 0.160 ms, 2^20: 3.13 -> 2.22 ms (73 % of the copy bandwidth); small models (a 6-dof arm's
    // whole matrix is 288 bytes) stay with the warp's coalesced stores.
    const char * mode = std::getenv("BRBD_GEN_CRBA_MODE");
    bulk = mode ? std::strcmp(mode, "bulk") == 0 : nvm > 24;
    compact = crba_mode == 1 || (mode && std::strcmp(mode, "compact") == 0);
    if (compact) bulk = false;
    group = std::max(1, std::min(std::min(nvm, 31), 192 / std::max(1, nvm)));
    if (bulk)
    { // as many columns as the CTA's staging rows leave room for
      group = 1;
      while (group < std::min(nvm, 31) &&
             (size_t)nt * crba_bulk_pitch(nvm, group + 1, fp32) * (fp32 ? 4 : 8) + 2048 <= (size_t)p->devs[0].max_smem_optin - 4096)
        ++group;
    }
    if (compact) group = 8; // entries of the pattern per flush, in units of 8
    if (crba_mode == 0)
    {
      if (const char * e = std::getenv("BRBD_GEN_CRBA_K")) group = std::max(1, std::min(31, std::atoi(e)));
      if (const char * e = std::getenv("BRBD_GEN_CRBA_NBUF")) nbuf = std::max(1, std::min(4, std::atoi(e)));
    }
  }
  const int gflags = (direct ? BRBD_GEN_DIRECT_IO : 0) | (slots ? BRBD_GEN_EXPLICIT_SLOTS : 0) | (fp32 ? BRBD_GEN_FP32 : 0) | (nt << 8) | (1 << 20) | (group << 24) |
                     (compact ? BRBD_GEN_CRBA_COMPACT : 0) | (bulk ? (BRBD_GEN_CRBA_BULK | ((nbuf - 1) << 29)) : 0);
  brbd_status st = brbd_codegen_source(&p->model, algo, gflags, &src, &info);
  if (st != BRBD_OK) return st;
  if ((size_t)info.dynamic_smem_bytes + 2048 + (compact ? 2 * (size_t)p->model.pd.nv * p->model.pd.nv : 0) > (size_t)p->devs[0].max_smem_optin)
  {
    brbd_codegen_free(src);
    return BRBD_OK;
  }
  std::vector<char> cubin;
  st = compile_cubin(src, (std::string("brbd_gen_") + kAlgoNames[algo] + ".cu").c_str(), cubin);
  brbd_codegen_free(src);
  if (st != BRBD_OK) return st;
  cudaLibrary_t lib = nullptr;
  CUDA_TRY(cudaLibraryLoadData(&lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
  cudaKernel_t kern = nullptr;
  const cudaError_t e = cudaLibraryGetKernel(&kern, lib, (std::string("brbd_gen_") + kAlgoNames[algo] + "_0").c_str());
  if (e != cudaSuccess)
  {
    cudaLibraryUnload(lib);
    return fail(BRBD_ECUDA, std::string("cudaLibraryGetKernel: ") + cudaGetErrorString(e));
  }
  k.smem_bytes = (size_t)info.dynamic_smem_bytes;
  cudaKernel_t kern_tma = nullptr;
  // (only the single-column warp-store source carries the tensor-store variant)
  if (algo == BRBD_GEN_CRBA && !bulk && !compact && group == 1 && cudaLibraryGetKernel(&kern_tma, lib, "brbd_gen_crba_tma") != cudaSuccess)
  { // grouped columns: the source has no tensor-store variant
    kern_tma = nullptr;
    (void)cudaGetLastError();
  }
  if (k.smem_bytes + 8 * 1024 > 48 * 1024) // (the compact CRBA also holds its position table, nv * nv shorts, in static shared memory)
    for (const DeviceCtx & d : p->devs)
    {
      CUDA_TRY(cudaKernelSetAttributeForDevice(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k.smem_bytes, d.dev));
      if (kern_tma) CUDA_TRY(cudaKernelSetAttributeForDevice(kern_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k.smem_bytes, d.dev));
    }
  k.kernel_tma = kern_tma;
  k.compact = compact;
  k.nnz = algo == BRBD_GEN_CRBA ? crba_pattern_nnz(p->model) : 0;
  k.lib = lib; k.kernel = kern; k.nt = nt; k.nrec = info.record_slots;
  return BRBD_OK;
}
} // namespace

brbd_status specialize_one(brbd_pool * p, int algo, bool fp32, int flags)
{
  GenSet & g = p->gen[algo][fp32 ? 1 : 0];
  if (g.nvar > 0) return BRBD_OK;
  // RNEA: warp tiles of q / v / a in shared memory (coalesced), as many warps (<= 8) as the tiles leave room for — its state
  // stays in registers.  ABA: long-lived values in explicit tensor-memory slots (168 registers and no spills at 12 warps for a
  // 35-dof humanoid), no tiles, 14 and 16 warps per SM: see GenSet.
  std::vector<int> nts;
  bool direct = algo == BRBD_GEN_ABA, slots = algo == BRBD_GEN_ABA || (flags & BRBD_GEN_EXPLICIT_SLOTS);
  if (const char * e = std::getenv("BRBD_GEN_SLOTS")) slots = std::atoi(e) != 0;
  if (const char * e = std::getenv("BRBD_GEN_DIRECT")) direct = std::atoi(e) != 0;
  const char * nt_env = std::getenv(algo == BRBD_GEN_CRBA ? "BRBD_GEN_CRBA_NT" : "BRBD_GEN_NT");
  if (nt_env) nts.push_back(std::max(32, std::min(1024, std::atoi(nt_env) / 32 * 32)));
  else if (algo >= BRBD_GEN_RNEA_DERIVATIVES)
  { // every result stays alive (registers limit the warps); as many warps (<= 8) as the per-warp result tiles leave room for
    const int nvm = p->model.pd.nv;
    const size_t tile_bytes = (size_t)32 * (3 * crba_bulk_pitch(nvm * nvm, 1, fp32) + crba_bulk_pitch(nvm, 1, fp32)) * (fp32 ? 4 : 8);
    const int w = (int)std::min<size_t>(8, (220 * 1024) / tile_bytes);
    nts = {w >= 4 ? 32 * w : 256}; // fewer than 4 warps: no tiles, every lane stores its own results
  }
  else if (algo == BRBD_GEN_CRBA) nts = p->model.pd.nv > 24 ? std::vector<int>{96, 64, 32} : std::vector<int>{512, 384, 256, 128}; // see build_variant
  else if (direct) nts = {448, 512, 256};
  else
  {
    const size_t tile_bytes = (size_t)32 * ((p->model.pd.nq | 1) + 2 * (p->model.pd.nv | 1)) * (fp32 ? 4 : 8);
    nts = {32 * (int)std::max<size_t>(1, std::min<size_t>(8, ((size_t)p->devs[0].max_smem_optin - 2048) / tile_bytes))};
  }
  for (int nt : nts)
  {
    if (g.nvar >= 2 && nt == 256) break; // 8 warps only when the larger variants do not fit
    if (algo == BRBD_GEN_CRBA && g.nvar >= 1) break;
    GenKernel k;
    brbd_status st = build_variant(p, algo, fp32, nt, direct, slots, k);
    if (st != BRBD_OK) return st;
    if (k.kernel) g.var[g.nvar++] = k;
  }
  if (g.nvar == 0) return fail(BRBD_EINVAL, "specialisation: the model's long-lived state does not fit the SM; the generic kernels remain in use");
  return BRBD_OK;
}

// one generated kernel: q, v, x (tau or a) -> out
template<class T>
brbd_status launch_generated(brbd_pool * p, DeviceCtx & d, int algo, const T * q, int64_t ldq, const T * v, int64_t ldv, const T * x,
                             int64_t ldx, T * out, int64_t ldo, int64_t B)
{
  const GenSet & g = p->gen[algo][sizeof(T) == 4 ? 1 : 0];
  // cheapest rounds * (c0 + warps); c0 = 9 measured on the 35-dof humanoid ABA (one pass: 0.124 ms at 8 warps, 0.178 at 16)
  int best = 0;
  int64_t best_cost = -1;
  for (int k = 0; k < g.nvar; ++k)
  {
    const int64_t per_round = (int64_t)d.sm_count * g.var[k].nt;
    const int64_t cost = ((B + per_round - 1) / per_round) * (9 + g.var[k].nt / 32);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = k; }
  }
  const GenKernel & k = g.var[best];
  const int64_t ctas_needed = (B + k.nt - 1) / k.nt;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ctas_needed, (int64_t)d.sm_count));
  if (algo == BRBD_GEN_CRBA && k.kernel_tma && !forced_path("BRBD_CRBA_V", "gen-lsu"))
  { // the caller's layout allows tensor maps: column blocks leave through TMA tensor stores (as crba_tma_kernel)
    CrbaTmaGeom G{0, 0, 0};
    CUtensorMap map0, map1;
    if (crba_tma_setup<T>(out, ldo, B, p->model.pd.nv, G, map0, map1))
    {
      long long ldq_ = ldq, ldo_ = ldo, B_ = B;
      int odd = G.odd, pairs = G.pairs;
      void * args[] = {(void *)&q, &ldq_, (void *)&out, &ldo_, &B_, &odd, &pairs, &map0, &map1};
      CUDA_TRY(cudaLaunchKernel((const void *)k.kernel_tma, dim3(grid), dim3(k.nt), args, k.smem_bytes, d.s()));
      p->launches += 1;
      return BRBD_OK;
    }
  }
  brbd_status st = ensure_work(d, std::max<size_t>(8, (size_t)k.nrec * grid * k.nt * sizeof(T)));
  if (st != BRBD_OK) return st;
  long long ldq_ = ldq, ldv_ = ldv, ldx_ = ldx, ldo_ = ldo, B_ = B;
  T * rec = (T *)d.work;
  void * args[] = {(void *)&q, &ldq_, (void *)&v, &ldv_, (void *)&x, &ldx_, (void *)&out, &ldo_, (void *)&rec, &B_};
  CUDA_TRY(cudaLaunchKernel((const void *)k.kernel, dim3(grid), dim3(k.nt), args, k.smem_bytes, d.s()));
  p->launches += 1;
  return BRBD_OK;
}
template<class T>
brbd_status launch_crba_packed(brbd_pool * p, DeviceCtx & d, const T * q, int64_t ldq, T * P, int64_t ldP, int64_t B)
{
  const bool fp32 = sizeof(T) == 4;
  GenKernel & k = p->crba_packed[fp32 ? 1 : 0];
  if (!k.kernel)
  { // first call: generate + compile the compact-staging kernel, with as many warps as its staging rows leave room for
    for (int nt : {480, 384, 256, 128, 64, 32})
    {
      brbd_status st = build_variant(p, BRBD_GEN_CRBA, fp32, nt, false, false, k, 1);
      if (st != BRBD_OK) return st;
      if (k.kernel) break;
    }
    if (!k.kernel) return fail(BRBD_EINVAL, "packed crba: the staging rows of the model's pattern do not fit the SM");
  }
  const int64_t ctas_needed = (B + k.nt - 1) / k.nt;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ctas_needed, (int64_t)d.sm_count));
  // the kernel's v / x / record arguments are unused by CRBA; ldx carries the output mode (1 = packed)
  const T * nul = nullptr;
  T * rec = nullptr;
  long long ldq_ = ldq, zero = 0, one = 1, ldo_ = ldP, B_ = B;
  void * args[] = {(void *)&q, &ldq_, (void *)&nul, &zero, (void *)&nul, &one, (void *)&P, &ldo_, (void *)&rec, &B_};
  CUDA_TRY(cudaLaunchKernel((const void *)k.kernel, dim3(grid), dim3(k.nt), args, k.smem_bytes, d.s()));
  p->launches += 1;
  return BRBD_OK;
}
template brbd_status launch_crba_packed<double>(brbd_pool *, DeviceCtx &, const double *, int64_t, double *, int64_t, int64_t);
template brbd_status launch_crba_packed<float>(brbd_pool *, DeviceCtx &, const float *, int64_t, float *, int64_t, int64_t);

template<class T>
brbd_status launch_generated_derivs(brbd_pool * p, DeviceCtx & d, int algo, const T * q, int64_t ldq, const T * v, int64_t ldv, const T * x,
                                    int64_t ldx, T * o0, int64_t ld0, T * o1, int64_t ld1, T * o2, int64_t ld2, T * o3, int64_t ld3, int64_t B)
{
  const GenKernel & k = p->gen[algo][sizeof(T) == 4 ? 1 : 0].var[0];
  const int64_t ctas_needed = (B + k.nt - 1) / k.nt;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ctas_needed, (int64_t)d.sm_count * (k.smem_bytes > 110 * 1024 ? 1 : 2)));
  long long ldq_ = ldq, ldv_ = ldv, ldx_ = ldx, l0 = ld0, l1 = ld1, l2 = ld2, l3 = ld3, B_ = B;
  void * args[] = {(void *)&q, &ldq_, (void *)&v, &ldv_, (void *)&x, &ldx_, (void *)&o0, &l0, (void *)&o1, &l1, (void *)&o2, &l2, (void *)&o3, &l3, &B_};
  CUDA_TRY(cudaLaunchKernel((const void *)k.kernel, dim3(grid), dim3(k.nt), args, k.smem_bytes, d.s()));
  p->launches += 1;
  return BRBD_OK;
}
template brbd_status launch_generated_derivs<double>(brbd_pool *, DeviceCtx &, int, const double *, int64_t, const double *, int64_t, const double *,
                                                     int64_t, double *, int64_t, double *, int64_t, double *, int64_t, double *, int64_t, int64_t);
template brbd_status launch_generated_derivs<float>(brbd_pool *, DeviceCtx &, int, const float *, int64_t, const float *, int64_t, const float *,
                                                    int64_t, float *, int64_t, float *, int64_t, float *, int64_t, float *, int64_t, int64_t);
template brbd_status launch_generated<double>(brbd_pool *, DeviceCtx &, int, const double *, int64_t, const double *, int64_t, const double *,
                                              int64_t, double *, int64_t, int64_t);
template brbd_status launch_generated<float>(brbd_pool *, DeviceCtx &, int, const float *, int64_t, const float *, int64_t, const float *,
                                             int64_t, float *, int64_t, int64_t);
} // namespace brbd
