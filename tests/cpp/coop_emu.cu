// coop_emu.cu — TEST INFRASTRUCTURE, not part of the product: a CPU lane emulator for the warp-cooperative kernels
// (pinocchio_b200/csrc/deriv_coop.cuh, aba_deriv_coop.cuh).  The per-configuration device code is compiled for the
// host (BRBD_DI = __host__ __device__), one std::thread per lane, __syncwarp() replaced by a pthread barrier, the
// group's shared-memory region by a heap block pre-filled with NaN (so a read of never-written state shows up).
// It lets the lane/phase logic of a kernel be checked against the oracle in a container without a GPU
// (tests/test_coop_emulator.py); the real parity tests remain the `-m gpu` ones through the C ABI.
#include <pthread.h>

#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <thread>
#include <vector>

#define BRBD_DI __host__ __device__ __forceinline__
namespace emu
{
static thread_local pthread_barrier_t * bar = nullptr;
inline void sync() { pthread_barrier_wait(bar); }
} // namespace emu
__host__ __device__ inline void brbd_emu_syncwarp()
{
#ifdef __CUDA_ARCH__
  __syncwarp();
#else
  emu::sync();
#endif
}
#define BRBD_SYNCWARP() brbd_emu_syncwarp()

#include "../../pinocchio_b200/csrc/model_build.hpp"
#include "../../pinocchio_b200/csrc/aba_deriv_coop.cuh"
#include "../../pinocchio_b200/csrc/minv_chol.cuh"

using namespace brbd;

namespace
{
std::string g_err;

template<class F> void run_lanes(int G, F && body)
{
  pthread_barrier_t b;
  pthread_barrier_init(&b, nullptr, (unsigned)G);
  std::vector<std::thread> th;
  for (int gl = 0; gl < G; ++gl)
    th.emplace_back([&, gl]() {
      emu::bar = &b;
      body(gl);
    });
  for (auto & t : th) t.join();
  pthread_barrier_destroy(&b);
}

double * alloc_region(size_t n)
{
  void * p = nullptr;
  if (posix_memalign(&p, 64, (n + 16) * sizeof(double)) != 0) return nullptr;
  double * d = static_cast<double *>(p);
  for (size_t k = 0; k < n + 16; ++k) d[k] = std::numeric_limits<double>::quiet_NaN();
  return d;
}

template<int G, int MODE>
void emu_aba_derivs(const ModelPOD<double> & m, const CoopTables & tb, const double * q, const double * v, const double * tau,
                    double * dq, double * dv, double * dtau, double * ddq, int64_t B)
{
  const AbaCoopLayout L = aba_coop_layout(m.nq, m.nv, m.njoints, G);
  const int64_t nn = (int64_t)m.nv * m.nv;
  for (int64_t cfg = 0; cfg < B; ++cfg)
  {
    double * base = alloc_region((size_t)L.per_group);
    for (int k = 0; k < m.nq; ++k) base[L.oq + k] = q[cfg * m.nq + k];
    for (int k = 0; k < m.nv; ++k)
    {
      base[L.ov + k] = MODE != 1 ? v[cfg * m.nv + k] : 0.0;
      base[L.ou + k] = MODE != 1 ? tau[cfg * m.nv + k] : 0.0;
    }
    run_lanes(G, [&](int gl) {
      aba_derivatives_coop_config<double, G, MODE>(m, tb, L, base, gl, dq + cfg * nn, dv + cfg * nn, dtau + cfg * nn, ddq + cfg * m.nv,
                                                   true);
    });
    free(base);
  }
}

template<int G>
void emu_rnea_derivs(const ModelPOD<double> & m, const CoopTables & tb, const double * q, const double * v, const double * a,
                     double * dq, double * dv, double * da, double * tau, int64_t B)
{
  const CoopLayout L = coop_layout(m.nq, m.nv, m.njoints);
  const int64_t nn = (int64_t)m.nv * m.nv;
  for (int64_t cfg = 0; cfg < B; ++cfg)
  {
    double * base = alloc_region((size_t)L.per_group);
    double * sq = base + L.oq, * sv = base + L.ov, * sa = base + L.oa, * jr = base + L.ojr, * cb = base + L.ocb;
    for (int k = 0; k < m.nq; ++k) sq[k] = q[cfg * m.nq + k];
    for (int k = 0; k < m.nv; ++k) { sv[k] = v[cfg * m.nv + k]; sa[k] = a[cfg * m.nv + k]; }
    run_lanes(G, [&](int gl) {
      int oa_off = JR_OA;
      const int xoff = coop_forward<double, G, true>(m, tb, sq, sv, sa, jr, cb, gl, &oa_off);
      coop_joint_quantities<double, G>(m, jr, gl, xoff, oa_off);
      coop_subtree_sums<double, G>(m, jr, gl);
      coop_columns<double, G>(m, jr, cb, sa, gl);
      coop_entries<double, G, true>(m, tb, cb, dq + cfg * nn, dv + cfg * nn, da + cfg * nn, gl, true);
      brbd_emu_syncwarp();
      for (int k = gl; k < m.nv; k += G) tau[cfg * m.nv + k] = sa[k];
    });
    free(base);
  }
}
template<int G>
void emu_rnea(const ModelPOD<double> & m, const CoopTables & tb, const double * q, const double * v, const double * a, double * tau, int64_t B)
{
  const CoopLayout L = coop_layout(m.nq, m.nv, m.njoints);
  for (int64_t cfg = 0; cfg < B; ++cfg)
  {
    double * base = alloc_region((size_t)L.per_group);
    for (int k = 0; k < m.nq; ++k) base[L.oq + k] = q[cfg * m.nq + k];
    for (int k = 0; k < m.nv; ++k) { base[L.ov + k] = v[cfg * m.nv + k]; base[L.oa + k] = a[cfg * m.nv + k]; }
    run_lanes(G, [&](int gl) { rnea_coop_config<double, G>(m, tb, L, base, gl); });
    for (int k = 0; k < m.nv; ++k) tau[cfg * m.nv + k] = base[L.oa + k];
    free(base);
  }
}
// computeMinverse by Cholesky (minv_chol.cuh): Min = the upper triangles of crba, column-major, one matrix per column
template<int G, int R>
void emu_minv_chol(int nv, const double * Min, double * out, int64_t B)
{
  const MinvCholLayout L = minv_chol_layout(nv);
  const int64_t nn = (int64_t)nv * nv;
  for (int64_t cfg = 0; cfg < B; ++cfg)
  {
    double * base = alloc_region((size_t)L.per_group);
    run_lanes(G, [&](int gl) { minv_chol_config<double, G, R>(nv, L, base, gl, Min + cfg * nn, out + cfg * nn, true); });
    free(base);
  }
}
template<int R>
void emu_minv_chol_blocked(int nv, const double * Min, double * out, int64_t B)
{
  const MinvCholBlockedLayout L = minv_chol_blocked_layout(nv);
  const int64_t nn = (int64_t)nv * nv;
  for (int64_t cfg = 0; cfg < B; ++cfg)
  {
    double * base = alloc_region((size_t)L.per_group);
    run_lanes(32, [&](int gl) { minv_chol_blocked_config<double, R>(nv, L, base, gl, Min + cfg * nn, out + cfg * nn, true); });
    free(base);
  }
}
} // namespace

extern "C" {

const char * emu_last_error(void) { return g_err.c_str(); }

int emu_minv_chol_blocked_run(int nv, const double * Min, double * out, int64_t B)
{
  if (nv <= 32) emu_minv_chol_blocked<1>(nv, Min, out, B);
  else if (nv <= 64) emu_minv_chol_blocked<2>(nv, Min, out, B);
  else return -1;
  return 0;
}

int emu_minv_chol_run(int nv, const double * Min, double * out, int64_t B)
{
  if (nv <= 8) emu_minv_chol<8, 1>(nv, Min, out, B);
  else if (nv <= 16) emu_minv_chol<16, 1>(nv, Min, out, B);
  else if (nv <= 32) emu_minv_chol<32, 1>(nv, Min, out, B);
  else if (nv <= 64) emu_minv_chol<32, 2>(nv, Min, out, B);
  else return -1;
  return 0;
}

// algo: 0 = computeRNEADerivatives (third input = a; outputs dtau_dq, dtau_dv, dtau_da, tau),
//       1 = computeABADerivatives  (third input = tau; outputs ddq_dq, ddq_dv, ddq_dtau, ddq),
//       2 = computeMinverse        (only q is read; output o3 = upper triangle of Minv + zeros),
//       3 = rnea, small-batch path (third input = a; output ovec = tau),
//       4 = aba, small-batch path  (third input = tau; output ovec = ddq).
// Dense column-major blocks, one configuration per column (ld == rows).  G = 0 picks the group size the engine uses.
int emu_derivatives(int algo, const brbd_flat_model * f, const double * q, const double * v, const double * x, double * o1, double * o2,
                    double * o3, double * ovec, int64_t B, int G)
{
  ModelPOD<double> m;
  CoopTables tb;
  const brbd_status st = build_model_pod(f, m, g_err);
  if (st != BRBD_OK) return (int)st;
  build_coop_tables(m, tb);
  if (G == 0) G = m.nv <= 8 ? 8 : (m.nv <= 16 ? 16 : 32);
  if ((algo == 1 || algo == 2 || algo == 4) && tb.nbranch > A_MAXBRANCH)
  {
    g_err = "more branching joints than save slots";
    return -1;
  }
#define EMU_RUN(GG)                                                                \
  {                                                                                \
    if (algo == 0) emu_rnea_derivs<GG>(m, tb, q, v, x, o1, o2, o3, ovec, B);       \
    else if (algo == 1) emu_aba_derivs<GG, 0>(m, tb, q, v, x, o1, o2, o3, ovec, B); \
    else if (algo == 2) emu_aba_derivs<GG, 1>(m, tb, q, v, x, o1, o2, o3, ovec, B); \
    else if (algo == 3) emu_rnea<GG>(m, tb, q, v, x, ovec, B);                     \
    else emu_aba_derivs<GG, 2>(m, tb, q, v, x, o1, o2, o3, ovec, B);               \
  }
  if (G == 8) EMU_RUN(8)
  else if (G == 16) EMU_RUN(16)
  else if (G == 32) EMU_RUN(32)
  else
  {
    g_err = "G must be 8, 16 or 32";
    return -1;
  }
#undef EMU_RUN
  return 0;
}

} // extern "C"
