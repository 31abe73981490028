// codegen/trace_derivs.hpp — computeRNEADerivatives / computeABADerivatives for the code generator.
//
// The algorithm definitions are the engine's generic one-configuration-per-thread functions, rnea_derivatives_thread
// (rnea_derivatives.cuh; reference impl::computeRNEADerivatives, rnea-derivatives.hxx:472-541) and aba_derivatives_thread
// (aba_derivatives.cuh; impl::computeABADerivatives, aba-derivatives.hxx:380-453), instantiated with the recording scalar:
// the same code that runs as the v1 fallback kernels is executed once on the host and leaves its expression graph.  The
// ColumnEmitter they write through is specialised for cg::Sym as a plain column buffer.  For computeABADerivatives the two
// products -Minv dtau_dq, -Minv dtau_dv (aba-derivatives.hxx:451-452) are traced too, so the kernel is one straight-line program.
// All 3 nv^2 + nv results are outputs of that program and stay alive until it ends: meant for SMALL models (the 6-dof
// manipulator of BASELINE configs[2]); the launch code uses it for nv <= 12.
#pragma once

#include <memory>
#include <vector>

#include "trace.hpp"
#include "../rnea_derivatives.cuh"
#include "../aba_derivatives.cuh"

namespace brbd
{
// column buffer standing in for the warp-cooperative emitter (engine.cuh): put / add / get on one column, flush copies it out
template<> struct ColumnEmitter<cg::Sym>
{
  std::vector<cg::Sym> col;
  int nv = 0;
  void init(int nv_) { nv = nv_; col.assign(nv_, cg::Sym(0.0)); }
  void put(int row, cg::Sym val) { col[row] = val; }
  void add(int row, cg::Sym val) { col[row] += val; }
  cg::Sym get(int row) const { return col[row]; }
  void flush(cg::Sym * g, int64_t, int)
  {
    for (int r = 0; r < nv; ++r) { g[r] = col[r]; col[r] = cg::Sym(0.0); }
  }
};

namespace cg
{
inline void sym_model(const ModelPOD<double> & D, ModelPOD<Sym> & P)
{
  P.njoints = D.njoints; P.nq = D.nq; P.nv = D.nv; P.maxdepth = D.maxdepth;
  for (int i = 0; i < MAXJ; ++i)
  {
    P.parent[i] = D.parent[i]; P.type[i] = D.type[i]; P.idx_q[i] = D.idx_q[i]; P.idx_v[i] = D.idx_v[i];
    P.nvj[i] = D.nvj[i]; P.nvsub[i] = D.nvsub[i]; P.depth[i] = D.depth[i]; P.unb[i] = D.unb[i];
    for (int k = 0; k < 12; ++k) P.placement[i][k] = Sym(D.placement[i][k]);
    for (int k = 0; k < 10; ++k) P.inertia[i][k] = Sym(D.inertia[i][k]);
  }
  for (int k = 0; k < MAXNV; ++k)
  {
    P.dof_joint[k] = D.dof_joint[k]; P.parent_row[k] = D.parent_row[k]; P.armature[k] = Sym(D.armature[k]);
  }
  for (int k = 0; k < 3; ++k) P.gravity[k] = Sym(D.gravity[k]);
}

// outputs: array 0 dtau_dq, 1 dtau_dv, 2 dtau_da (each nv*nv, col-major), 3 tau
inline void trace_rnea_derivatives(Tracer & T)
{
  const ModelPOD<double> & M = T.M;
  const int nv = M.nv, nn = nv * nv;
  std::unique_ptr<ModelPOD<Sym>> P(new ModelPOD<Sym>());
  sym_model(M, *P);
  std::vector<Sym> q, v, a, gq(nn, Sym(0.0)), gv(nn, Sym(0.0)), ga(nn, Sym(0.0));
  for (int k = 0; k < M.nq; ++k) q.push_back(T.in(IN_Q, k));
  for (int k = 0; k < nv; ++k) v.push_back(T.in(IN_V, k));
  for (int k = 0; k < nv; ++k) a.push_back(T.in(IN_X, k));
  ColumnEmitter<Sym> eq, ev, ea;
  eq.init(nv); ev.init(nv); ea.init(nv);
  rnea_derivatives_thread(*P, q.data(), v.data(), a.data(), eq, ev, ea, gq.data(), (int64_t)nn, gv.data(), (int64_t)nn, ga.data(), (int64_t)nn, 1);
  for (int k = 0; k < nn; ++k) { T.output(0, k, gq[k]); T.output(1, k, gv[k]); T.output(2, k, ga[k]); }
  for (int k = 0; k < nv; ++k) T.output(3, k, a[k]); // a_tau: tau on return
}

// outputs: array 0 ddq_dq, 1 ddq_dv, 2 ddq_dtau (= Minv, full symmetric), 3 ddq
inline void trace_aba_derivatives(Tracer & T)
{
  const ModelPOD<double> & M = T.M;
  const int nv = M.nv, nn = nv * nv;
  std::unique_ptr<ModelPOD<Sym>> P(new ModelPOD<Sym>());
  sym_model(M, *P);
  std::vector<Sym> q, v, u, gq(nn, Sym(0.0)), gv(nn, Sym(0.0)), gm(nn, Sym(0.0));
  for (int k = 0; k < M.nq; ++k) q.push_back(T.in(IN_Q, k));
  for (int k = 0; k < nv; ++k) v.push_back(T.in(IN_V, k));
  for (int k = 0; k < nv; ++k) u.push_back(T.in(IN_X, k));
  std::vector<Sym> minv((size_t)nn, Sym(0.0)), fd((size_t)(M.maxdepth + 1) * nv * 6, Sym(0.0));
  ColumnEmitter<Sym> eq, ev, em;
  eq.init(nv); ev.init(nv); em.init(nv);
  const ThreadWS<Sym> wm{minv.data(), 1}, wf{fd.data(), 1};
  aba_derivatives_thread(*P, q.data(), v.data(), u.data(), eq, ev, em, gq.data(), (int64_t)nn, gv.data(), (int64_t)nn, gm.data(), (int64_t)nn, wm, wf, 1);
  // ddq_dq = -Minv dtau_dq, ddq_dv = -Minv dtau_dv (aba-derivatives.hxx:451-452); entries are col-major: (r, c) at c * nv + r
  for (int c = 0; c < nv; ++c)
    for (int r = 0; r < nv; ++r)
    {
      Sym aq(0.0), av(0.0);
      for (int k = 0; k < nv; ++k)
      {
        aq += gm[k * nv + r] * gq[c * nv + k];
        av += gm[k * nv + r] * gv[c * nv + k];
      }
      T.output(0, c * nv + r, -aq);
      T.output(1, c * nv + r, -av);
    }
  for (int k = 0; k < nn; ++k) T.output(2, k, gm[k]);
  for (int k = 0; k < nv; ++k) T.output(3, k, u[k]); // data.ddq
}

} // namespace cg
} // namespace brbd
