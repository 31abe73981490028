// oracle/oracle_capi.cpp — C entry points of the CPU oracle (TEST INFRASTRUCTURE, see
// rbd_oracle.hpp).  The batch drivers restate rneaInParallel / abaInParallel
// (reference: include/pinocchio/algorithm/parallel/rnea.hpp:69-82, parallel/aba.hpp:70-83,
// parallel/omp.hpp:12-16): one private (Model, Data) workspace per OpenMP thread and
// `#pragma omp parallel for schedule(static)` over the batch columns.
#define ORACLE_WITH_QUAD 1
#include <quadmath.h>

#include "rbd_oracle.hpp"

#include <omp.h>

#include <memory>

using namespace rbdo;

struct oracle_model
{
  Model<double> md;
  Model<long double> ml;
  Model<Counted> mc;
  Model<__float128> mq;
  explicit oracle_model(const brbd_flat_model & f) : md(f), ml(f), mc(f), mq(f) {}
};

namespace
{
template<class S> struct Pool
{
  std::vector<std::unique_ptr<Data<S>>> datas;
  Pool(const Model<S> & m, int n)
  {
    for (int i = 0; i < n; ++i) datas.emplace_back(new Data<S>(m));
  }
};
// set_default_omp_options, parallel/omp.hpp:12-16
inline int setup_threads(int num_threads)
{
  if (num_threads <= 0) num_threads = omp_get_max_threads();
  omp_set_num_threads(num_threads);
  omp_set_dynamic(0);
  return num_threads;
}
template<class S> inline void load(const double * src, int n, std::vector<S> & dst)
{
  dst.resize(n);
  for (int k = 0; k < n; ++k) dst[k] = S(src[k]);
}
template<class S> inline const S * in_ptr(const double * src, int n, std::vector<S> & tmp)
{
  load(src, n, tmp);
  return tmp.data();
}
template<> inline const double * in_ptr<double>(const double * src, int, std::vector<double> &) { return src; }

template<class S>
void rnea_batch(const Model<S> & m, const double * q, const double * v, const double * a, double * tau, int64_t B, int nt)
{
  nt = setup_threads(nt);
  Pool<S> pool(m, nt);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < B; ++i)
  {
    Data<S> & d = *pool.datas[omp_get_thread_num()];
    std::vector<S> tq, tv, ta;
    rnea(m, d, in_ptr(q + i * m.nq, m.nq, tq), in_ptr(v + i * m.nv, m.nv, tv), in_ptr(a + i * m.nv, m.nv, ta));
    for (int k = 0; k < m.nv; ++k) tau[i * m.nv + k] = to_double(d.tau[k]);
  }
}
template<class S>
void aba_batch(const Model<S> & m, const double * q, const double * v, const double * tau, double * a, int64_t B, int nt)
{
  nt = setup_threads(nt);
  Pool<S> pool(m, nt);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < B; ++i)
  {
    Data<S> & d = *pool.datas[omp_get_thread_num()];
    std::vector<S> tq, tv, ta;
    abaWorld(m, d, in_ptr(q + i * m.nq, m.nq, tq), in_ptr(v + i * m.nv, m.nv, tv), in_ptr(tau + i * m.nv, m.nv, ta));
    for (int k = 0; k < m.nv; ++k) a[i * m.nv + k] = to_double(d.ddq[k]);
  }
}
template<class S>
void crba_batch(const Model<S> & m, const double * q, double * M, int64_t B, int nt, int world)
{
  nt = setup_threads(nt);
  Pool<S> pool(m, nt);
  const int64_t nn = (int64_t)m.nv * m.nv;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < B; ++i)
  {
    Data<S> & d = *pool.datas[omp_get_thread_num()];
    std::vector<S> tq;
    if (world) crbaWorld(m, d, in_ptr(q + i * m.nq, m.nq, tq));
    else crbaLocal(m, d, in_ptr(q + i * m.nq, m.nq, tq));
    for (int64_t k = 0; k < nn; ++k) M[i * nn + k] = to_double(d.M[k]);
  }
}
template<class S>
void rnea_derivs_batch(const Model<S> & m, const double * q, const double * v, const double * a, double * dq,
                       double * dv, double * da, double * tau, int64_t B, int nt)
{
  nt = setup_threads(nt);
  Pool<S> pool(m, nt);
  const int64_t nn = (int64_t)m.nv * m.nv;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < B; ++i)
  {
    Data<S> & d = *pool.datas[omp_get_thread_num()];
    std::vector<S> tq, tv, ta, odq(nn, S(0)), odv(nn, S(0)), oda(nn, S(0));
    rneaDerivatives(m, d, in_ptr(q + i * m.nq, m.nq, tq), in_ptr(v + i * m.nv, m.nv, tv),
                    in_ptr(a + i * m.nv, m.nv, ta), odq.data(), odv.data(), oda.data());
    for (int64_t k = 0; k < nn; ++k)
    {
      dq[i * nn + k] = to_double(odq[k]);
      dv[i * nn + k] = to_double(odv[k]);
      da[i * nn + k] = to_double(oda[k]);
    }
    if (tau) for (int k = 0; k < m.nv; ++k) tau[i * m.nv + k] = to_double(d.tau[k]);
  }
}
template<class S>
void aba_derivs_batch(const Model<S> & m, const double * q, const double * v, const double * tau, double * dq,
                      double * dv, double * dtau, double * ddq, int64_t B, int nt)
{
  nt = setup_threads(nt);
  Pool<S> pool(m, nt);
  const int64_t nn = (int64_t)m.nv * m.nv;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < B; ++i)
  {
    Data<S> & d = *pool.datas[omp_get_thread_num()];
    std::vector<S> tq, tv, ta, odq(nn, S(0)), odv(nn, S(0)), odt(nn, S(0));
    abaDerivatives(m, d, in_ptr(q + i * m.nq, m.nq, tq), in_ptr(v + i * m.nv, m.nv, tv),
                   in_ptr(tau + i * m.nv, m.nv, ta), odq.data(), odv.data(), odt.data());
    for (int64_t k = 0; k < nn; ++k)
    {
      dq[i * nn + k] = to_double(odq[k]);
      dv[i * nn + k] = to_double(odv[k]);
      dtau[i * nn + k] = to_double(odt[k]);
    }
    if (ddq) for (int k = 0; k < m.nv; ++k) ddq[i * m.nv + k] = to_double(d.ddq[k]);
  }
}
} // namespace

extern "C" {

oracle_model * oracle_model_create(const brbd_flat_model * f) { return new oracle_model(*f); }
void oracle_model_destroy(oracle_model * m) { delete m; }
int oracle_max_threads(void) { return omp_get_max_threads(); }

// precision: 0 = double, 1 = long double (x87 80-bit), 2 = __float128 (IEEE binary128); inputs/outputs are double
void oracle_rnea(const oracle_model * m, const double * q, const double * v, const double * a, double * tau,
                 int64_t B, int nthreads, int precision)
{
  if (precision == 2) rnea_batch(m->mq, q, v, a, tau, B, nthreads);
  else if (precision) rnea_batch(m->ml, q, v, a, tau, B, nthreads);
  else rnea_batch(m->md, q, v, a, tau, B, nthreads);
}
void oracle_aba(const oracle_model * m, const double * q, const double * v, const double * tau, double * a,
                int64_t B, int nthreads, int precision)
{
  if (precision == 2) aba_batch(m->mq, q, v, tau, a, B, nthreads);
  else if (precision) aba_batch(m->ml, q, v, tau, a, B, nthreads);
  else aba_batch(m->md, q, v, tau, a, B, nthreads);
}
// world: 1 = crbaWorldConvention, 0 = crbaLocalConvention (the default of crba(), crba.hpp:51)
void oracle_crba(const oracle_model * m, const double * q, double * M, int64_t B, int nthreads, int precision,
                 int world)
{
  if (precision == 2) crba_batch(m->mq, q, M, B, nthreads, world);
  else if (precision) crba_batch(m->ml, q, M, B, nthreads, world);
  else crba_batch(m->md, q, M, B, nthreads, world);
}
void oracle_rnea_derivatives(const oracle_model * m, const double * q, const double * v, const double * a,
                             double * dq, double * dv, double * da, double * tau, int64_t B, int nthreads,
                             int precision)
{
  if (precision == 2) rnea_derivs_batch(m->mq, q, v, a, dq, dv, da, tau, B, nthreads);
  else if (precision) rnea_derivs_batch(m->ml, q, v, a, dq, dv, da, tau, B, nthreads);
  else rnea_derivs_batch(m->md, q, v, a, dq, dv, da, tau, B, nthreads);
}
void oracle_aba_derivatives(const oracle_model * m, const double * q, const double * v, const double * tau,
                            double * dq, double * dv, double * dtau, double * ddq, int64_t B, int nthreads,
                            int precision)
{
  if (precision == 2) aba_derivs_batch(m->mq, q, v, tau, dq, dv, dtau, ddq, B, nthreads);
  else if (precision) aba_derivs_batch(m->ml, q, v, tau, dq, dv, dtau, ddq, B, nthreads);
  else aba_derivs_batch(m->md, q, v, tau, dq, dv, dtau, ddq, B, nthreads);
}

// qout[:, i] = integrate(model, q[:, i], v[:, i])  (algorithm/joint-configuration.hpp:49-74)
void oracle_integrate(const oracle_model * m, const double * q, const double * v, double * qout, int64_t B, int precision)
{
  for (int64_t i = 0; i < B; ++i)
  {
    if (precision)
    {
      const Model<long double> & ml = m->ml;
      std::vector<long double> tq(ml.nq), tv(ml.nv), to(ml.nq);
      for (int k = 0; k < ml.nq; ++k) tq[k] = q[i * ml.nq + k];
      for (int k = 0; k < ml.nv; ++k) tv[k] = v[i * ml.nv + k];
      integrate(ml, tq.data(), tv.data(), to.data());
      for (int k = 0; k < ml.nq; ++k) qout[i * ml.nq + k] = (double)to[k];
    }
    else
      integrate(m->md, q + i * m->md.nq, v + i * m->md.nv, qout + i * m->md.nq);
  }
}

// Exact algorithmic operation counts of one evaluation: out = {add, mul, div, sqrt, sincos}.
// algo: 0 rnea, 1 aba(world), 2 crba(world), 3 crba(local), 4 rnea-derivatives, 5 aba-derivatives
void oracle_count_flops(const oracle_model * m, int algo, const double * q, const double * v, const double * a,
                        uint64_t out[5])
{
  const Model<Counted> & mc = m->mc;
  Data<Counted> d(mc);
  std::vector<Counted> tq, tv, ta;
  load(q, mc.nq, tq);
  load(v, mc.nv, tv);
  load(a, mc.nv, ta);
  const size_t nn = (size_t)mc.nv * mc.nv;
  std::vector<Counted> o1(nn), o2(nn), o3(nn);
  flop_counter() = FlopCounter();
  switch (algo)
  {
  case 0: rnea(mc, d, tq.data(), tv.data(), ta.data()); break;
  case 1: abaWorld(mc, d, tq.data(), tv.data(), ta.data()); break;
  case 2: crbaWorld(mc, d, tq.data()); break;
  case 3: crbaLocal(mc, d, tq.data()); break;
  case 4: rneaDerivatives(mc, d, tq.data(), tv.data(), ta.data(), o1.data(), o2.data(), o3.data()); break;
  case 5: abaDerivatives(mc, d, tq.data(), tv.data(), ta.data(), o1.data(), o2.data(), o3.data()); break;
  default: break;
  }
  const FlopCounter c = flop_counter();
  out[0] = c.add; out[1] = c.mul; out[2] = c.div; out[3] = c.sqrt_; out[4] = c.sincos;
}

} // extern "C"
