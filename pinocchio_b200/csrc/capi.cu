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
#include "host_ctx.hpp"
#include <nvtx3/nvToolsExt.h>
#include "model_build.hpp"

using namespace brbd;

namespace
{
thread_local std::string g_err;
}
namespace brbd
{
brbd_status fail(brbd_status s, const std::string & msg)
{
  g_err = msg;
  return s;
}
} // namespace brbd

namespace
{
// NVTX range around every batched entry point of the C ABI (tracing): shows up by name in Nsight Systems / as an
// ncu --nvtx-include filter; a no-op when no tool is attached.
struct NvtxRange
{
  explicit NvtxRange(const char * name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};
#define BRBD_NVTX(name) NvtxRange nvtx_range__(name)
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
  // host pointers: shard columns contiguously over the devices; every shard moves through its device in chunks on three
  // streams (upload | kernels | download) with double-buffered staging.  The chunk loop is the OUTER loop and the devices the
  // inner one, so that every device has work queued before the host blocks in a copy from pageable memory (with pinned
  // blocks — brbd_host_register — nothing blocks and the devices overlap fully either way).
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
  // whatever happens below, every device gets its stream selection back and is idle when the call returns
  struct Restore
  {
    brbd_pool * p;
    std::vector<char> user;
    explicit Restore(brbd_pool * pool) : p(pool)
    {
      for (DeviceCtx & d : p->devs) { user.push_back(d.use_user_stream ? 1 : 0); d.use_user_stream = false; }
    }
    ~Restore()
    {
      for (size_t g = 0; g < p->devs.size(); ++g)
      {
        DeviceCtx & d = p->devs[g];
        d.use_user_stream = user[g] != 0;
        if (cudaSetDevice(d.dev) != cudaSuccess) continue;
        cudaStreamSynchronize(d.s_in);
        cudaStreamSynchronize(d.stream);
        cudaStreamSynchronize(d.s_out);
      }
    }
  } restore(p); // kernels of host-pointer calls run on the pool's own streams
  struct Shard { int64_t c0, c1, cw, next; int it; };
  std::vector<Shard> shards(nd);
  for (int g = 0; g < nd; ++g)
  {
    Shard & sh = shards[g];
    sh.c0 = std::min<int64_t>(B, (int64_t)g * per);
    sh.c1 = std::min<int64_t>(B, sh.c0 + per);
    sh.cw = std::max<int64_t>(1, std::min<int64_t>(chunk, sh.c1 - sh.c0));
    sh.next = sh.c0;
    sh.it = 0;
    if (sh.c0 >= sh.c1) continue;
    DeviceCtx & d = p->devs[g];
    CUDA_TRY(cudaSetDevice(d.dev));
    for (size_t k = 0; k < args.size(); ++k)
      if (args[k].in || args[k].out)
        for (int b = 0; b < 2; ++b)
        {
          brbd_status st = ensure_stage(d, (int)(2 * k + b), (size_t)args[k].rows * sh.cw * sizeof(T));
          if (st != BRBD_OK) return st;
        }
  }
  for (bool any = true; any;)
  {
    any = false;
    for (int g = 0; g < nd; ++g)
    {
      Shard & sh = shards[g];
      if (sh.next >= sh.c1) continue;
      any = true;
      DeviceCtx & d = p->devs[g];
      CUDA_TRY(cudaSetDevice(d.dev));
      const int64_t b0 = sh.next;
      const int it = sh.it;
      const int buf = it & 1;
      const int64_t nb = std::min<int64_t>(sh.cw, sh.c1 - b0);
      sh.next += nb;
      sh.it += 1;
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
      const brbd_status result = launch(d, ptrs, nb);
      args = saved;
      if (result != BRBD_OK) return result;
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
  }
  // errors of the asynchronous work surface here (the guard synchronises again, harmlessly, on the way out)
  for (int g = 0; g < nd; ++g)
  {
    CUDA_TRY(cudaSetDevice(p->devs[g].dev));
    CUDA_TRY(cudaStreamSynchronize(p->devs[g].s_in));
    CUDA_TRY(cudaStreamSynchronize(p->devs[g].stream));
    CUDA_TRY(cudaStreamSynchronize(p->devs[g].s_out));
  }
  return BRBD_OK;
}

// Host-pointer crba with packed transfer (brbd_pool_set_host_threads): per chunk, q up | the generated CRBA with compact staging
// | the nnz pattern entries down into a pinned landing buffer | host threads rebuild the caller's dense matrices (host_expand.cpp)
// while the next chunk is in flight.  Single-device pools.  *done = false when the packed kernel is not available (no NVRTC):
// the caller then takes the dense path.
template<class T>
brbd_status run_crba_expand(brbd_pool * p, const T * q, int64_t ldq, T * M, int64_t ldM, int64_t B, bool * done)
{
  *done = false;
  DeviceCtx & d = p->devs[0];
  const int nq = p->model.pd.nq, nv = p->model.pd.nv, nn = nv * nv;
  if (p->crba_idx.empty()) crba_pattern_index(p->model, p->crba_idx);
  const int64_t nnz = (int64_t)p->crba_idx.size();
  if (2 * nnz > nn) return BRBD_OK; // nothing to gain from packing (a chain: the upper triangle is the pattern)
  CUDA_TRY(cudaSetDevice(d.dev));
  // chunks of ~24 MB of packed entries (8192 configurations of a 35-dof humanoid; measured with 65 536 of them, 16 threads:
  // 4096 / 8192 / 16 384 / 32 768 per chunk -> 7.05 / 5.85 / 6.14 / 7.32 ms, the dense copy 11.4 ms; profiles/r2_crba_host_sweep.txt)
  int64_t chunk = std::min<int64_t>(B, std::max<int64_t>(4096, (((int64_t)(24u << 20) / (int64_t)(nnz * sizeof(T))) / 1024) * 1024));
  if (const char * e = std::getenv("BRBD_EXPAND_CHUNK")) chunk = std::min<int64_t>(B, std::max<int64_t>(1024, std::atoll(e))); // experiments
  brbd_status st = BRBD_OK;
  for (int b = 0; b < 2 && st == BRBD_OK; ++b)
  {
    st = ensure_stage(d, b, (size_t)nq * chunk * sizeof(T));
    if (st == BRBD_OK) st = ensure_stage(d, 2 + b, (size_t)nnz * chunk * sizeof(T));
    const size_t need = (size_t)nnz * chunk * sizeof(T);
    if (st == BRBD_OK && p->host_stage_bytes[b] < need)
    {
      if (p->host_stage[b]) cudaFreeHost(p->host_stage[b]);
      p->host_stage[b] = nullptr;
      p->host_stage_bytes[b] = 0;
      CUDA_TRY(cudaHostAlloc(&p->host_stage[b], need, cudaHostAllocDefault));
      p->host_stage_bytes[b] = need;
    }
  }
  if (st != BRBD_OK) return st;
  const bool user = d.use_user_stream;
  d.use_user_stream = false; // the kernels of host-pointer calls run on the pool's own streams
  struct Guard
  {
    DeviceCtx & d; bool user;
    ~Guard()
    {
      d.use_user_stream = user;
      cudaStreamSynchronize(d.s_in); cudaStreamSynchronize(d.stream); cudaStreamSynchronize(d.s_out);
    }
  } guard{d, user};
  const int64_t nchunks = (B + chunk - 1) / chunk;
  auto expand = [&](int64_t it) -> brbd_status {
    const int buf = (int)(it & 1);
    const int64_t b0 = it * chunk, nb = std::min<int64_t>(chunk, B - b0);
    CUDA_TRY(cudaEventSynchronize(d.ev_out[buf]));
    expand_packed<T>(M + b0 * ldM, ldM, static_cast<const T *>(p->host_stage[buf]), nnz, p->crba_idx.data(), nn, nb, p->host_threads);
    return BRBD_OK;
  };
  for (int64_t it = 0; it < nchunks; ++it)
  {
    const int buf = (int)(it & 1);
    const int64_t b0 = it * chunk, nb = std::min<int64_t>(chunk, B - b0);
    if (it >= 2) CUDA_TRY(cudaStreamWaitEvent(d.s_in, d.ev_k[buf], 0)); // the kernel that read this q buffer two chunks ago
    const T * src = q + b0 * ldq;
    if (ldq == nq) CUDA_TRY(cudaMemcpyAsync(d.stage[buf], src, (size_t)nq * nb * sizeof(T), cudaMemcpyHostToDevice, d.s_in));
    else
      CUDA_TRY(cudaMemcpy2DAsync(d.stage[buf], nq * sizeof(T), src, ldq * sizeof(T), nq * sizeof(T), nb, cudaMemcpyHostToDevice, d.s_in));
    CUDA_TRY(cudaEventRecord(d.ev_in[buf], d.s_in));
    CUDA_TRY(cudaStreamWaitEvent(d.stream, d.ev_in[buf], 0));
    if (it >= 2) CUDA_TRY(cudaStreamWaitEvent(d.stream, d.ev_out[buf], 0)); // the download of this P buffer two chunks ago
    st = launch_crba_packed<T>(p, d, (const T *)d.stage[buf], nq, (T *)d.stage[2 + buf], nnz, nb);
    if (st != BRBD_OK)
    {
      if (it == 0)
      { // no NVRTC on this host: remember it and let the caller copy the dense block
        p->packed_unavailable = true;
        return BRBD_OK;
      }
      return st;
    }
    CUDA_TRY(cudaEventRecord(d.ev_k[buf], d.stream));
    CUDA_TRY(cudaStreamWaitEvent(d.s_out, d.ev_k[buf], 0));
    // (the landing buffer of this parity was expanded before this chunk was queued: see below)
    CUDA_TRY(cudaMemcpyAsync(p->host_stage[buf], d.stage[2 + buf], (size_t)nnz * nb * sizeof(T), cudaMemcpyDeviceToHost, d.s_out));
    CUDA_TRY(cudaEventRecord(d.ev_out[buf], d.s_out));
    if (it >= 1)
    {
      st = expand(it - 1);
      if (st != BRBD_OK) return st;
    }
  }
  st = expand(nchunks - 1);
  if (st != BRBD_OK) return st;
  *done = true;
  return BRBD_OK;
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
  const int n = f->njoints;
  m->f_parents.assign(f->parents, f->parents + n);
  m->f_type.assign(f->joint_type, f->joint_type + n);
  m->f_idx_q.assign(f->idx_q, f->idx_q + n);
  m->f_idx_v.assign(f->idx_v, f->idx_v + n);
  m->f_placement.assign(f->placement, f->placement + 12 * n);
  m->f_inertia.assign(f->inertia, f->inertia + 10 * n);
  m->f_armature.assign(f->nv, 0.0);
  if (f->armature) m->f_armature.assign(f->armature, f->armature + f->nv);
  m->f_axis.assign(3 * n, 0.0);
  if (f->axis) m->f_axis.assign(f->axis, f->axis + 3 * n);
  for (int k = 0; k < 3; ++k) m->f_gravity[k] = f->gravity[k];
  *out = m;
  return BRBD_OK;
}
void brbd_model_destroy(brbd_model * m) { delete m; }
brbd_status brbd_model_get_flat(const brbd_model * m, brbd_flat_model * out)
{
  if (!m || !out) return fail(BRBD_EINVAL, "null argument");
  out->njoints = m->pd.njoints; out->nq = m->pd.nq; out->nv = m->pd.nv;
  out->parents = m->f_parents.data(); out->joint_type = m->f_type.data();
  out->idx_q = m->f_idx_q.data(); out->idx_v = m->f_idx_v.data();
  out->placement = m->f_placement.data(); out->inertia = m->f_inertia.data();
  out->armature = m->f_armature.data(); out->axis = m->f_axis.data();
  for (int k = 0; k < 3; ++k) out->gravity[k] = m->f_gravity[k];
  return BRBD_OK;
}
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
  release_generated(p);
  for (int b = 0; b < 2; ++b)
    if (p->host_stage[b]) cudaFreeHost(p->host_stage[b]);
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
  for (size_t a = 0; a < ids.size(); ++a)
    for (size_t b = a + 1; b < ids.size(); ++b)
      if (ids[a] == ids[b]) return fail(BRBD_EINVAL, "device id " + std::to_string(ids[a]) + " listed twice");
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
const brbd_model * brbd_pool_model(const brbd_pool * p) { return p ? &p->model : nullptr; }
int brbd_pool_device_id(const brbd_pool * p, int index)
{
  return (p && index >= 0 && index < (int)p->devs.size()) ? p->devs[index].dev : -1;
}
uint64_t brbd_pool_workspace_bytes(const brbd_pool * p, int index)
{
  if (!p || index < 0 || index >= (int)p->devs.size()) return 0;
  const DeviceCtx & d = p->devs[index];
  uint64_t n = d.work_bytes + d.aux_bytes + sizeof(ModelPOD<double>) + sizeof(ModelPOD<float>) + MAXNV * sizeof(double);
  for (int k = 0; k < 16; ++k) n += d.stage_bytes[k];
  return n;
}
brbd_status brbd_pool_resize(brbd_pool * p, const int * device_ids, int n_devices)
{
  if (!p) return fail(BRBD_EINVAL, "null pool");
  brbd_status st = brbd_pool_synchronize(p);
  if (st != BRBD_OK) return st;
  brbd_pool * fresh = nullptr;
  st = brbd_pool_create(&p->model, device_ids, n_devices, &fresh);
  if (st != BRBD_OK) return st;
  std::swap(p->devs, fresh->devs); // the old replicas (streams, arenas, staged models) go away with `fresh`
  brbd_pool_destroy(fresh);
  return BRBD_OK;
}
brbd_status brbd_pool_update(brbd_pool * p, const brbd_model * m)
{
  if (!p || !m) return fail(BRBD_EINVAL, "null argument");
  brbd_status st = brbd_pool_synchronize(p);
  if (st != BRBD_OK) return st;
  release_generated(p); // kernels generated for the previous model
  p->model = *m;
  return upload_model(p);
}
brbd_status brbd_pool_specialize(brbd_pool * p, int algo_mask, int flags)
{
  BRBD_NVTX("brbd_pool_specialize");
  if (!p) return fail(BRBD_EINVAL, "null pool");
  brbd_status st = brbd_pool_synchronize(p);
  if (st != BRBD_OK) return st;
  if (!p->devs.empty()) CUDA_TRY(cudaSetDevice(p->devs[0].dev));
  for (int algo = 0; algo < 5; ++algo)
    if (algo_mask & (1 << algo))
    {
      st = specialize_one(p, algo, (flags & BRBD_GEN_FP32) != 0, flags);
      if (st != BRBD_OK) return st;
    }
  return BRBD_OK;
}
int brbd_pool_specialized(const brbd_pool * p)
{
  int mask = 0;
  if (p)
    for (int algo = 0; algo < 5; ++algo)
      if (p->gen[algo][0].nvar || p->gen[algo][1].nvar) mask |= 1 << algo;
  return mask;
}
brbd_status brbd_pool_set_specialized_min_batch(brbd_pool * p, int64_t min_batch)
{
  if (!p) return fail(BRBD_EINVAL, "null pool");
  p->gen_min_batch = min_batch;
  return BRBD_OK;
}
brbd_status brbd_pool_set_stream(brbd_pool * p, void * cuda_stream)
{
  if (!p) return fail(BRBD_EINVAL, "null pool");
  if (p->devs.size() != 1) return fail(BRBD_EINVAL, "external streams require a single-device pool");
  // NULL restores the pool's own (non-blocking) stream, as include/pinocchio_b200.h documents; the legacy default stream
  // is selected with its CUDA handle cudaStreamLegacy ((cudaStream_t)0x1), the per-thread one with cudaStreamPerThread
  p->devs[0].user_stream = static_cast<cudaStream_t>(cuda_stream);
  p->devs[0].use_user_stream = cuda_stream != nullptr;
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
  BRBD_NVTX("brbd_rnea_batch");
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
  BRBD_NVTX("brbd_aba_batch");
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
  BRBD_NVTX("brbd_crba_batch");
  if (!p) return fail(BRBD_EINVAL, "null pool");
  const int nq = p->model.pd.nq, nv = p->model.pd.nv;
  std::vector<Arg> args = {{q, nullptr, ldq, nq, false}, {nullptr, M, ldM, (int64_t)nv * nv, false}};
  // host pointers + host threads granted (brbd_pool_set_host_threads): packed transfer, dense matrices rebuilt by the threads
  if (!(flags & BRBD_PTR_DEVICE) && p->host_threads >= 2 && p->devs.size() == 1 && !p->packed_unavailable && q && M && batch >= 4096 &&
      ldq >= nq && ldM >= (int64_t)nv * nv)
  {
    bool done = false;
    brbd_status st = (flags & BRBD_FP32) ? run_crba_expand<float>(p, (const float *)q, ldq, (float *)M, ldM, batch, &done)
                                         : run_crba_expand<double>(p, (const double *)q, ldq, (double *)M, ldM, batch, &done);
    if (st != BRBD_OK || done) return st;
  }
  DISPATCH(flags, (run_call<T>(p, args, batch, flags, [&](DeviceCtx & d, std::vector<void *> & P, int64_t B) {
             return launch_crba<T>(p, d, (const T *)P[0], args[0].ld, (T *)P[1], args[1].ld, B);
           })));
}

brbd_status brbd_crba_packed_batch(brbd_pool * p, const void * q, int64_t ldq, void * P, int64_t ldP, int64_t batch, int flags)
{
  BRBD_NVTX("brbd_crba_packed_batch");
  if (!p) return fail(BRBD_EINVAL, "null pool");
  const int nq = p->model.pd.nq;
  if (p->crba_idx.empty()) crba_pattern_index(p->model, p->crba_idx); // computed once per model (brbd_pool_update clears it)
  const int64_t nnz = (int64_t)p->crba_idx.size();
  std::vector<Arg> args = {{q, nullptr, ldq, nq, false}, {nullptr, P, ldP, nnz, false}};
  DISPATCH(flags, (run_call<T>(p, args, batch, flags, [&](DeviceCtx & d, std::vector<void *> & PP, int64_t B) {
             return launch_crba_packed<T>(p, d, (const T *)PP[0], args[0].ld, (T *)PP[1], args[1].ld, B);
           })));
}

brbd_status brbd_rnea_derivatives_batch(brbd_pool * p, const void * q, int64_t ldq, const void * v, int64_t ldv,
                                        const void * a, int64_t lda, void * dtau_dq, int64_t ld_dq, void * dtau_dv,
                                        int64_t ld_dv, void * dtau_da, int64_t ld_da, void * tau, int64_t ldtau,
                                        int64_t batch, int flags)
{
  BRBD_NVTX("brbd_rnea_derivatives_batch");
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
  BRBD_NVTX("brbd_aba_derivatives_batch");
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
  BRBD_NVTX("brbd_nle_batch");
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
  BRBD_NVTX("brbd_gravity_batch");
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
  BRBD_NVTX("brbd_minverse_batch");
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
  BRBD_NVTX("brbd_integrate_batch");
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
  BRBD_NVTX("brbd_aba_euler_step_batch");
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

brbd_status brbd_crba_expand_packed(const brbd_model * m, const void * P, int64_t ldP, void * M, int64_t ldM, int64_t batch, int threads,
                                    int flags)
{
  if (!m || !P || !M) return fail(BRBD_EINVAL, "null argument");
  if (flags & BRBD_PTR_DEVICE) return fail(BRBD_EINVAL, "brbd_crba_expand_packed works on host blocks");
  std::vector<int32_t> idx;
  crba_pattern_index(*m, idx);
  const int nv = m->pd.nv;
  const int64_t nnz = (int64_t)idx.size(), nn = (int64_t)nv * nv;
  if (batch < 0) return fail(BRBD_EINVAL, "negative batch size");
  if (ldP < nnz) return fail(BRBD_EINVAL, "P: leading dimension " + std::to_string(ldP) + " smaller than the expected number of rows " + std::to_string(nnz));
  if (ldM < nn) return fail(BRBD_EINVAL, "M: leading dimension " + std::to_string(ldM) + " smaller than the expected number of rows " + std::to_string(nn));
  if (batch == 0) return BRBD_OK;
  if (flags & BRBD_FP32) expand_packed<float>((float *)M, ldM, (const float *)P, nnz, idx.data(), (int)nn, batch, threads, ldP);
  else expand_packed<double>((double *)M, ldM, (const double *)P, nnz, idx.data(), (int)nn, batch, threads, ldP);
  return BRBD_OK;
}
brbd_status brbd_pool_set_host_threads(brbd_pool * p, int n)
{
  if (!p) return fail(BRBD_EINVAL, "null pool");
  if (n < 0) return fail(BRBD_EINVAL, "negative number of host threads");
  p->host_threads = n;
  return BRBD_OK;
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
  return measure_fp64_peak(p, flops_per_s, elapsed_ms);
}

} // extern "C"
