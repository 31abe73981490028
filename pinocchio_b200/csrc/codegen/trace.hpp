// codegen/trace.hpp — the algorithms of the hot path written once more for the recording scalar (sym.hpp), in the order and
// with the data flow the generated kernel should have.  They restate the same reference steps as the generic kernels:
//   ABA  impl::abaWorldConvention, include/pinocchio/algorithm/aba.hxx:242-293 (steps :101-138, :152-192, :206-231)
//   RNEA impl::rnea, algorithm/rnea.hxx:117-161 (steps :45-79, :92-107)
//   CRBA impl::crbaWorldConvention, algorithm/crba.hxx:498-548 (steps :35-58, :80-99)
// using the engine's own spatial algebra (spatial.cuh, engine.cuh, aba.cuh instantiated with cg::Sym).
//
// ABA follows aba_rr.cuh (v4): depth-first interleaving of passes 1 and 2, the backward sweep recomputes the joint's world
// quantities from (oMi, ov) walked back up the chain, so that the long-lived state is (sin q, cos q, v) per tree depth and one
// slot per open branching joint; pass 3 recomputes the forward kinematics too and reads 10 values per 1-dof joint
// (sin, cos, U Dinv, Dinv, u) from the per-thread record store.  `Tracer::park / fetch` mark what is long-lived: the emitter
// decides where it lives (registers + local memory under the compiler's control, or explicit on-chip slots).
#pragma once
#include <cstdlib>

#include <string>
#include <vector>

#include "sym.hpp"
#include "../engine.cuh"
#include "../aba.cuh"

namespace brbd
{
namespace cg
{

enum InputArray { IN_Q = 0, IN_V = 1, IN_X = 2 };         // x = tau (ABA) or a (RNEA)
enum OutputArray { OUT_MAIN = 0 };                         // ddq / tau / M

struct Tracer
{
  Graph g;
  const ModelPOD<double> & M;
  bool explicit_parking;  // false: park / fetch are the identity (the compiler places long-lived values)
  int nrec = 0;           // record slots per configuration
  explicit Tracer(const ModelPOD<double> & m, bool explicit_park) : M(m), explicit_parking(explicit_park) { current_graph() = &g; }
  ~Tracer() { if (current_graph() == &g) current_graph() = nullptr; }

  // inputs: `gen` distinguishes re-reads of the same element (a later pass re-loads instead of keeping the value alive)
  Sym in(int array, int index, int gen = 0) { return Sym::node(g.intern(OP_INPUT, array, index, (double)gen)); }
  Sym c(double v) { return Sym(v); }

  struct Parked { int handle = -1; int fetches = 0; std::vector<Sym> vals; };
  Parked park(const std::vector<Sym> & vals, const char * what)
  {
    Parked p;
    p.vals = vals;
    p.handle = g.n_handles++;
    g.handle_size.push_back((int)vals.size());
    if (explicit_parking)
    {
      Effect e;
      e.kind = Effect::PARK; e.at = (int)g.nodes.size(); e.handle = p.handle; e.index = 0; e.text = what;
      for (const Sym & s : vals) e.vals.push_back(s.id);
      g.effects.push_back(e);
    }
    return p;
  }
  std::vector<Sym> fetch(Parked & p)
  {
    if (!explicit_parking) return p.vals;
    std::vector<Sym> out;
    const double gen = (double)p.fetches++; // every fetch re-loads: the values do not stay alive between two readers
    for (int k = 0; k < (int)p.vals.size(); ++k)
    {
      // a parked constant stays a constant
      if (g.is_const(p.vals[k].id)) out.push_back(p.vals[k]);
      else out.push_back(Sym::node(g.intern(OP_FETCH, p.handle, k, gen)));
    }
    return out;
  }
  void release(Parked & p)
  {
    if (explicit_parking && p.handle >= 0)
    {
      Effect e;
      e.kind = Effect::RELEASE; e.at = (int)g.nodes.size(); e.handle = p.handle; e.index = 0;
      g.effects.push_back(e);
    }
    p.handle = -1;
    p.vals.clear();
  }
  // per-thread record store in global memory ([slot][thread], coalesced): written in pass 2, read in pass 3
  int record_store(const std::vector<Sym> & vals)
  {
    Effect e;
    e.kind = Effect::RECORD_ST; e.at = (int)g.nodes.size(); e.handle = -1; e.index = nrec;
    for (const Sym & s : vals) e.vals.push_back(s.id);
    g.effects.push_back(e);
    const int first = nrec;
    nrec += (int)vals.size();
    return first;
  }
  std::vector<Sym> record_load(int first, int n)
  {
    std::vector<Sym> out;
    for (int k = 0; k < n; ++k) out.push_back(Sym::node(g.intern(OP_INPUT, 100, first + k, 0.0))); // array 100 = the record store
    return out;
  }
  void output(int array, int row, const Sym & v)
  {
    Effect e;
    e.kind = Effect::OUTPUT; e.at = (int)g.nodes.size(); e.handle = array; e.index = row; e.vals.push_back(v.id);
    g.effects.push_back(e);
  }
  // matrix output through the wrapper's staging row (CRBA): colbegin(col), clear(row) for what the previous column left,
  // output(row, value) for this column's entries, flush(col)
  void col_effect(int what, int index)
  {
    Effect e;
    e.kind = Effect::FLUSH; e.at = (int)g.nodes.size(); e.handle = what; e.index = index;
    g.effects.push_back(e);
  }
  void colbegin(int col) { col_effect(0, col); }
  void clear(int row) { col_effect(2, row); }
  void flush(int col) { col_effect(1, col); }
  void comment(const std::string & s)
  {
    Effect e;
    e.kind = Effect::COMMENT; e.at = (int)g.nodes.size(); e.handle = -1; e.index = 0; e.text = s;
    g.effects.push_back(e);
  }
};

// ---- model constants as recording scalars ---------------------------------------------------------------------------
inline SE3<Sym> sym_placement(const ModelPOD<double> & M, int i)
{
  const double * P = M.placement[i];
  SE3<Sym> X;
  X.R.c0 = Vec3<Sym>(Sym(P[0]), Sym(P[1]), Sym(P[2]));
  X.R.c1 = Vec3<Sym>(Sym(P[3]), Sym(P[4]), Sym(P[5]));
  X.R.c2 = Vec3<Sym>(Sym(P[6]), Sym(P[7]), Sym(P[8]));
  X.p = Vec3<Sym>(Sym(P[9]), Sym(P[10]), Sym(P[11]));
  return X;
}
inline Inertia<Sym> sym_inertia(const ModelPOD<double> & M, int i)
{
  const double * Y = M.inertia[i];
  Inertia<Sym> I;
  I.m = Sym(Y[0]);
  I.c = Vec3<Sym>(Sym(Y[1]), Sym(Y[2]), Sym(Y[3]));
  I.I.xx = Sym(Y[4]); I.I.xy = Sym(Y[5]); I.I.yy = Sym(Y[6]); I.I.xz = Sym(Y[7]); I.I.yz = Sym(Y[8]); I.I.zz = Sym(Y[9]);
  return I;
}

// liMi = jointPlacements[i] * M_J(q) with the structural zeros of M_J dropped (engine.cuh joint_liMi); 1-dof joints take
// (sin q, cos q) / q as `s`, `c` so that a later pass can feed stored values; multi-dof joints take their q segment
inline SE3<Sym> sym_liMi(const ModelPOD<double> & M, int i, const Sym & s, const Sym & c, const std::vector<Sym> & qj)
{
  const int type = M.type[i];
  const SE3<Sym> P = sym_placement(M, i);
  SE3<Sym> X;
  if (type <= J_RZ)
  {
    X.p = P.p;
    if (type == J_RX) { X.R.c0 = P.R.c0; X.R.c1 = c * P.R.c1 + s * P.R.c2; X.R.c2 = c * P.R.c2 - s * P.R.c1; }
    else if (type == J_RY) { X.R.c1 = P.R.c1; X.R.c2 = c * P.R.c2 + s * P.R.c0; X.R.c0 = c * P.R.c0 - s * P.R.c2; }
    else { X.R.c2 = P.R.c2; X.R.c0 = c * P.R.c0 + s * P.R.c1; X.R.c1 = c * P.R.c1 - s * P.R.c0; }
  }
  else if (type <= J_PZ)
  {
    X.R = P.R;
    X.p = P.p + s * P.R.col(type - J_PX);
  }
  else if (type == J_FF)
  {
    SE3<Sym> MJ;
    MJ.R = quat_to_mat(qj[3], qj[4], qj[5], qj[6]);
    MJ.p = Vec3<Sym>(qj[0], qj[1], qj[2]);
    X = P * MJ;
  }
  else if (type == J_SPH)
  {
    X.R = P.R * quat_to_mat(qj[0], qj[1], qj[2], qj[3]);
    X.p = P.p;
  }
  else
  { // planar: q = (x, y, cos, sin)
    X.R.c0 = qj[2] * P.R.c0 + qj[3] * P.R.c1;
    X.R.c1 = qj[2] * P.R.c1 - qj[3] * P.R.c0;
    X.R.c2 = P.R.c2;
    X.p = P.p + qj[0] * P.R.c0 + qj[1] * P.R.c1;
  }
  return X;
}
inline Mat3<Sym> sym_mul_bt(const Mat3<Sym> & A, const Mat3<Sym> & B) // A B^T
{
  Mat3<Sym> r;
  r.c0 = A * Vec3<Sym>(B.c0.x, B.c1.x, B.c2.x);
  r.c1 = A * Vec3<Sym>(B.c0.y, B.c1.y, B.c2.y);
  r.c2 = A * Vec3<Sym>(B.c0.z, B.c1.z, B.c2.z);
  return r;
}
inline std::vector<Sym> pack_se3_motion(const SE3<Sym> & X, const Motion<Sym> & m)
{
  return {X.R.c0.x, X.R.c0.y, X.R.c0.z, X.R.c1.x, X.R.c1.y, X.R.c1.z, X.R.c2.x, X.R.c2.y, X.R.c2.z, X.p.x, X.p.y, X.p.z,
          m.lin.x, m.lin.y, m.lin.z, m.ang.x, m.ang.y, m.ang.z};
}
inline void unpack_se3_motion(const std::vector<Sym> & x, SE3<Sym> & X, Motion<Sym> & m)
{
  X.R.c0 = Vec3<Sym>(x[0], x[1], x[2]); X.R.c1 = Vec3<Sym>(x[3], x[4], x[5]); X.R.c2 = Vec3<Sym>(x[6], x[7], x[8]);
  X.p = Vec3<Sym>(x[9], x[10], x[11]);
  m.lin = Vec3<Sym>(x[12], x[13], x[14]); m.ang = Vec3<Sym>(x[15], x[16], x[17]);
}

struct JointTopo
{
  std::vector<int> nchild, stop;
  explicit JointTopo(const ModelPOD<double> & M) : nchild(M.njoints, 0), stop(M.njoints, 0)
  {
    for (int i = 1; i < M.njoints; ++i) nchild[M.parent[i]]++;
    for (int i = 1; i < M.njoints; ++i) stop[i] = i + 1 < M.njoints ? M.parent[i + 1] : 0;
  }
};

// ======================================================================================================================
// ABA (WORLD convention)
// ======================================================================================================================
inline void trace_aba(Tracer & T)
{
  const ModelPOD<double> & M = T.M;
  const int nj = M.njoints;
  const JointTopo topo(M);
  struct Branch { Tracer::Parked acc; bool open = false; };   // the children's (oYaba, of) of a branching joint so far
  std::vector<Branch> branch(nj);
  std::vector<Tracer::Parked> depth_scv(M.maxdepth + 2);             // (s, c, v) of the joint at each depth of the current path
  struct Rec { int first = -1, n = 0; };
  std::vector<Rec> rec(nj);
  // values of the last joint of pass 2 (a root) can stay in registers for pass 3: kept as plain values
  SE3<Sym> X;
  Motion<Sym> ov = mzero<Sym>();

  auto qseg = [&](int i, int gen) {
    std::vector<Sym> q;
    const int nqj = (i + 1 < nj ? M.idx_q[i + 1] : M.nq) - M.idx_q[i];
    for (int k = 0; k < nqj; ++k) q.push_back(T.in(IN_Q, M.idx_q[i] + k, gen));
    return q;
  };
  auto vseg = [&](int i, int gen) {
    std::vector<Sym> v;
    for (int k = 0; k < M.nvj[i]; ++k) v.push_back(T.in(IN_V, M.idx_v[i] + k, gen));
    return v;
  };
  auto joint_vel = [&](int i, const std::vector<Sym> & vj) {
    std::vector<Sym> tmp = vj;
    return joint_velocity(M.type[i], tmp.data());
  };

  // Loads are volatile and stay where the trace issues them, so they are issued AHEAD of their use: q / v of joint i + 1 while
  // joint i is computed, tau of the next joint of an unwind one step early, the pass-3 records two joints ahead (first version:
  // each load sat right in front of its consumer and `long_scoreboard` was 48 % of the stall samples of the kernel).
  std::vector<std::vector<Sym>> q_pf(nj + 1), v_pf(nj + 1), tau_pf(nj + 1);
  auto tauseg = [&](int j) {
    std::vector<Sym> t;
    for (int k = 0; k < M.nvj[j]; ++k) t.push_back(T.in(IN_X, M.idx_v[j] + k, 0));
    return t;
  };
  if (nj > 1) { q_pf[1] = qseg(1, 0); v_pf[1] = vseg(1, 0); }
  for (int i = 1; i < nj; ++i)
  {
    const int type = M.type[i], parent = M.parent[i], nvj = M.nvj[i];
    T.comment("pass 1, joint " + std::to_string(i));
    // ---- pass 1, joint i: kinematics only (aba.hxx:101-131) ----
    Sym si(0.0), ci(0.0);
    std::vector<Sym> qj = q_pf[i], vj = v_pf[i];
    if (i + 1 < nj) { q_pf[i + 1] = qseg(i + 1, 0); v_pf[i + 1] = vseg(i + 1, 0); }
    if (topo.stop[i] != i) tau_pf[i] = tauseg(i); // a leaf: its backward step follows at once
    if (type <= J_RZ && M.unb[i]) { ci = qj[0]; si = qj[1]; }
    else if (type <= J_RZ) sincos_t(qj[0], &si, &ci);
    else if (type <= J_PZ) si = qj[0];
    const SE3<Sym> Xl = sym_liMi(M, i, si, ci, qj);
    if (parent > 0)
    {
      X = X * Xl;
      if (nvj == 1)
      {
        const Motion<Sym> J0 = act_S_col(X, type, 0);
        ov.lin += vj[0] * J0.lin;
        ov.ang += vj[0] * J0.ang;
      }
      else
        ov += X.act(joint_vel(i, vj));
    }
    else
    {
      X = Xl;
      if (nvj == 1)
      {
        const Motion<Sym> J0 = act_S_col(X, type, 0);
        ov.lin = vj[0] * J0.lin;
        ov.ang = vj[0] * J0.ang;
      }
      else
        ov = X.act(joint_vel(i, vj));
    }
    if (topo.nchild[i] > 0 && nvj == 1) depth_scv[M.depth[i]] = T.park({si, ci, vj[0]}, "scv");
    if (topo.nchild[i] >= 2) branch[i].open = false;
    // ---- pass 2 for every joint whose subtree is now complete (aba.hxx:152-192) ----
    const int stop = topo.stop[i];
    bool have_child = false;              // an only child's contribution handed over as plain values
    Sym Ac[21];
    Force<Sym> fc = fzero<Sym>();
    for (int j = i; j != stop; j = M.parent[j])
    {
      T.comment("pass 2, joint " + std::to_string(j));
      const int tj = M.type[j], pj = M.parent[j], nv_j = M.nvj[j], iv = M.idx_v[j];
      if (pj != stop && pj > 0) tau_pf[pj] = tauseg(pj); // the unwind goes on with the parent
      Sym sj = si, cj = ci;
      std::vector<Sym> vjj = vj;
      std::vector<Sym> qjj = qj;
      if (j != i)
      {
        if (nv_j == 1)
        {
          std::vector<Sym> s3 = T.fetch(depth_scv[M.depth[j]]);
          T.release(depth_scv[M.depth[j]]);
          sj = s3[0]; cj = s3[1]; vjj = {s3[2]};
        }
        else
        {
          qjj = qseg(j, 1);
          vjj = vseg(j, 1);
        }
      }
      const Inertia<Sym> Y = act(X, sym_inertia(M, j));
      Force<Sym> fi = fcross(ov, Y * ov);
      Sym A[21];
      inertia_to_sym6(Y, A);
      if (j != i)
      {
        if (topo.nchild[j] >= 2)
        { // the children's contributions sit in the accumulator
          std::vector<Sym> acc = T.fetch(branch[j].acc);
          T.release(branch[j].acc);
          for (int k = 0; k < 21; ++k) A[k] += acc[k];
          fi.lin += Vec3<Sym>(acc[21], acc[22], acc[23]);
          fi.ang += Vec3<Sym>(acc[24], acc[25], acc[26]);
        }
        else if (have_child)
        {
          for (int k = 0; k < 21; ++k) A[k] += Ac[k];
          fi += fc;
        }
      }
      // joint columns, parent's velocity, bias acceleration
      Motion<Sym> Jc[6];
      for (int k = 0; k < nv_j; ++k) Jc[k] = act_S_col(X, tj, k);
      Motion<Sym> ovp = mzero<Sym>(), abm = mzero<Sym>();
      if (pj > 0)
      {
        if (nv_j == 1)
        {
          ovp.lin = ov.lin - vjj[0] * Jc[0].lin;
          ovp.ang = ov.ang - vjj[0] * Jc[0].ang;
        }
        else
          ovp = ov - X.act(joint_vel(j, vjj));
        abm = mcross(ovp, ov);
      }
      // U = Ia J, StU = J^T U + armature, Dinv, UDinv, u
      Sym U[6][6], StU[6][6], Di[6][6], UD[6][6], uj[6];
      for (int k = 0; k < nv_j; ++k)
      {
        uj[k] = tau_pf[j][k] - dot6(Jc[k], fi);
        Sym Jv[6], Uk[6];
        m2a(Jc[k], Jv);
        sym6_mul(A, Jv, Uk);
        for (int r = 0; r < 6; ++r) U[r][k] = Uk[r];
      }
      for (int a = 0; a < nv_j; ++a)
      {
        Sym Jv[6];
        m2a(Jc[a], Jv);
        for (int b = 0; b < nv_j; ++b)
        {
          Sym acc = Jv[0] * U[0][b];
          for (int r = 1; r < 6; ++r) acc += Jv[r] * U[r][b];
          StU[a][b] = acc;
        }
        StU[a][a] += Sym(M.armature[iv + a]);
      }
      if (nv_j == 1) Di[0][0] = Sym(1.0) / StU[0][0];
      else llt_inverse(nv_j, StU, Di);
      for (int r = 0; r < 6; ++r)
        for (int k = 0; k < nv_j; ++k)
        {
          Sym acc = U[r][0] * Di[0][k];
          for (int cc = 1; cc < nv_j; ++cc) acc += U[r][cc] * Di[cc][k];
          UD[r][k] = acc;
        }
      { // pass-3 record: (sin, cos) of a 1-dof joint (pass 3 rebuilds the kinematics from them), UDinv (6 per dof), Dinv u
        std::vector<Sym> r;
        if (nv_j == 1) { r.push_back(sj); r.push_back(cj); }
        for (int k = 0; k < nv_j; ++k)
          for (int rr = 0; rr < 6; ++rr) r.push_back(UD[rr][k]);
        for (int a = 0; a < nv_j; ++a)
        {
          Sym acc = Di[a][0] * uj[0];
          for (int b = 1; b < nv_j; ++b) acc += Di[a][b] * uj[b];
          r.push_back(acc);
        }
        rec[j].first = T.record_store(r);
        rec[j].n = (int)r.size();
      }
      if (pj > 0)
      {
        for (int r = 0; r < 6; ++r)
          for (int cc = r; cc < 6; ++cc)
          {
            Sym acc = UD[r][0] * U[cc][0];
            for (int k = 1; k < nv_j; ++k) acc += UD[r][k] * U[cc][k];
            A[r * 6 - (r * (r - 1)) / 2 + (cc - r)] -= acc;
          }
        Sym ab[6], Iab[6], fa[6];
        m2a(abm, ab);
        sym6_mul(A, ab, Iab);
        f2a(fi, fa);
        for (int r = 0; r < 6; ++r)
        {
          Sym acc = UD[r][0] * uj[0];
          for (int k = 1; k < nv_j; ++k) acc += UD[r][k] * uj[k];
          fa[r] += Iab[r] + acc;
        }
        if (topo.nchild[pj] >= 2)
        {
          std::vector<Sym> acc(27);
          if (branch[pj].open)
          {
            acc = T.fetch(branch[pj].acc);
            T.release(branch[pj].acc);
            for (int k = 0; k < 21; ++k) acc[k] += A[k];
            for (int k = 0; k < 6; ++k) acc[21 + k] += fa[k];
          }
          else
          {
            for (int k = 0; k < 21; ++k) acc[k] = A[k];
            for (int k = 0; k < 6; ++k) acc[21 + k] = fa[k];
          }
          branch[pj].acc = T.park(acc, "branch acc");
          branch[pj].open = true;
          have_child = false;
        }
        else
        {
          for (int k = 0; k < 21; ++k) Ac[k] = A[k];
          fc.lin = Vec3<Sym>(fa[0], fa[1], fa[2]);
          fc.ang = Vec3<Sym>(fa[3], fa[4], fa[5]);
          have_child = true;
        }
        { // back to the parent: oMi_parent = oMi_j liMi_j^-1, ov_parent = ov_j - J_j v_j.  Also when the parent is the branching
          // joint the next joint hangs off (pj == stop): its (oMi, ov) come back for free instead of being kept per branch
          const SE3<Sym> Xlj = sym_liMi(M, j, sj, cj, qjj);
          SE3<Sym> Xp;
          Xp.R = sym_mul_bt(X.R, Xlj.R);
          Xp.p = X.p - Xp.R * Xlj.p;
          X = Xp;
          ov = ovp;
        }
      }
    }
  }

  // ---- pass 3 (aba.hxx:206-226) with the forward kinematics rebuilt from the recorded (sin, cos) -------------------
  std::vector<Tracer::Parked> bstate(nj);   // (oMi, ov, oa_gf) of a branching joint for its later children
  Motion<Sym> ag = mzero<Sym>();
  struct P3 { std::vector<Sym> R, qj, vj; };
  std::vector<P3> p3(nj + 1);
  auto p3_load = [&](int i) {
    p3[i].R = T.record_load(rec[i].first, rec[i].n);
    if (M.nvj[i] > 1) p3[i].qj = qseg(i, 2);
    p3[i].vj = vseg(i, 2);
  };
  int P3_AHEAD = 1; // joints (measured, BRBD_GEN_ABA_P3AHEAD = 0 .. 4: 0.1583 / 0.1556 / 0.1575 / 0.1628 / 0.1667 ms at 65 536 x simple_humanoid)
  if (const char * e = std::getenv("BRBD_GEN_ABA_P3AHEAD")) P3_AHEAD = std::max(0, std::min(8, std::atoi(e)));
  for (int i = 1; i < nj && i <= P3_AHEAD; ++i) p3_load(i);
  for (int i = 1; i < nj; ++i)
  {
    T.comment("pass 3, joint " + std::to_string(i));
    const int type = M.type[i], parent = M.parent[i], nvj = M.nvj[i], iv = M.idx_v[i];
    if (i + P3_AHEAD < nj) p3_load(i + P3_AHEAD);
    const std::vector<Sym> R = p3[i].R;
    int o = 0;
    Sym si(0.0), ci(0.0);
    const std::vector<Sym> qj = p3[i].qj;
    if (nvj == 1) { si = R[0]; ci = R[1]; o = 2; }
    const std::vector<Sym> vj = p3[i].vj;
    Motion<Sym> ovp = mzero<Sym>(), agp = mzero<Sym>();
    if (parent == 0)
      agp.lin = Vec3<Sym>(Sym(-M.gravity[0]), Sym(-M.gravity[1]), Sym(-M.gravity[2])); // data.oa_gf[0] = -gravity (aba.hxx:260)
    else if (parent != i - 1)
    {
      std::vector<Sym> x = T.fetch(bstate[parent]);
      unpack_se3_motion(x, X, ovp);
      agp.lin = Vec3<Sym>(x[18], x[19], x[20]);
      agp.ang = Vec3<Sym>(x[21], x[22], x[23]);
    }
    else
    {
      ovp = ov;
      agp = ag;
    }
    const SE3<Sym> Xl = sym_liMi(M, i, si, ci, qj);
    X = parent > 0 ? X * Xl : Xl;
    Motion<Sym> Jc[6];
    for (int k = 0; k < nvj; ++k) Jc[k] = act_S_col(X, type, k);
    ov = ovp;
    for (int k = 0; k < nvj; ++k)
    {
      ov.lin += vj[k] * Jc[k].lin;
      ov.ang += vj[k] * Jc[k].ang;
    }
    Motion<Sym> agv = agp;
    if (parent > 0) agv += mcross(ovp, ov);
    Sym av[6];
    m2a(agv, av);
    // ddq = Dinv u - UDinv^T oa_gf ; oa_gf += J ddq
    const Sym * UD = &R[o];                 // [k][6]
    const Sym * Du = &R[o + 6 * nvj];       // (Dinv u)[k]
    Sym dd[6];
    for (int k = 0; k < nvj; ++k)
    {
      const Sym t1 = Du[k];
      Sym t2 = UD[6 * k] * av[0];
      for (int r = 1; r < 6; ++r) t2 += UD[6 * k + r] * av[r];
      dd[k] = t1 - t2;
      T.output(OUT_MAIN, iv + k, dd[k]);
    }
    ag = agv;
    for (int k = 0; k < nvj; ++k)
    {
      ag.lin += dd[k] * Jc[k].lin;
      ag.ang += dd[k] * Jc[k].ang;
    }
    if (topo.nchild[i] >= 2)
    {
      std::vector<Sym> x = pack_se3_motion(X, ov);
      x.insert(x.end(), {ag.lin.x, ag.lin.y, ag.lin.z, ag.ang.x, ag.ang.y, ag.ang.z});
      bstate[i] = T.park(x, "branch pass 3");
    }
  }
}

// ======================================================================================================================
// RNEA
// ======================================================================================================================
inline void trace_rnea(Tracer & T)
{
  const ModelPOD<double> & M = T.M;
  const int nj = M.njoints;
  const JointTopo topo(M);
  // local-frame recursion as the reference: v_i = liMi^-1 v_parent + S qd, a_gf likewise, f_i = Y a + v x* (Y v); the
  // backward sweep runs depth-first interleaved, f of an only child travels as plain values, a branching joint accumulates
  struct Open { Tracer::Parked kin, f; bool has_f = false; };
  std::vector<Open> open(nj);
  std::vector<Tracer::Parked> depth_x(M.maxdepth + 2);   // liMi (as its generating values) and own f of the joints on the current path
  Motion<Sym> v = mzero<Sym>(), a = mzero<Sym>();
  std::vector<SE3<Sym>> liMi(nj);
  std::vector<Force<Sym>> fown(nj);
  std::vector<Tracer::Parked> fpark(nj);
  for (int i = 1; i < nj; ++i)
  {
    T.comment("forward, joint " + std::to_string(i));
    const int type = M.type[i], parent = M.parent[i], nvj = M.nvj[i], iv = M.idx_v[i];
    std::vector<Sym> qj, vj, aj;
    const int nqj = (i + 1 < nj ? M.idx_q[i + 1] : M.nq) - M.idx_q[i];
    for (int k = 0; k < nqj; ++k) qj.push_back(T.in(IN_Q, M.idx_q[i] + k));
    for (int k = 0; k < nvj; ++k) vj.push_back(T.in(IN_V, iv + k));
    for (int k = 0; k < nvj; ++k) aj.push_back(T.in(IN_X, iv + k));
    Sym si(0.0), ci(0.0);
    if (type <= J_RZ && M.unb[i]) { ci = qj[0]; si = qj[1]; }
    else if (type <= J_RZ) sincos_t(qj[0], &si, &ci);
    else if (type <= J_PZ) si = qj[0];
    const SE3<Sym> Xl = sym_liMi(M, i, si, ci, qj);
    liMi[i] = Xl;
    Motion<Sym> vp = mzero<Sym>(), ap = mzero<Sym>();
    if (parent == 0) ap.lin = Vec3<Sym>(Sym(-M.gravity[0]), Sym(-M.gravity[1]), Sym(-M.gravity[2]));
    else if (parent != i - 1)
    {
      std::vector<Sym> x = T.fetch(open[parent].kin);
      vp.lin = Vec3<Sym>(x[0], x[1], x[2]); vp.ang = Vec3<Sym>(x[3], x[4], x[5]);
      ap.lin = Vec3<Sym>(x[6], x[7], x[8]); ap.ang = Vec3<Sym>(x[9], x[10], x[11]);
    }
    else { vp = v; ap = a; }
    std::vector<Sym> tmpv = vj;
    Motion<Sym> vi = joint_velocity(type, tmpv.data());
    if (parent > 0) vi += Xl.actInv(vp);
    Motion<Sym> ai = cross_joint_velocity(vi, type, tmpv.data());
    for (int k = 0; k < nvj; ++k)
    {
      const int row = joint_S_row(type, k);
      if (row < 3) ai.lin.set(row, ai.lin.get(row) + aj[k]); else ai.ang.set(row - 3, ai.ang.get(row - 3) + aj[k]);
    }
    ai += Xl.actInv(ap);
    v = vi; a = ai;
    if (topo.nchild[i] >= 2)
      open[i].kin = T.park({vi.lin.x, vi.lin.y, vi.lin.z, vi.ang.x, vi.ang.y, vi.ang.z, ai.lin.x, ai.lin.y, ai.lin.z, ai.ang.x, ai.ang.y, ai.ang.z}, "rnea branch kin");
    const Inertia<Sym> Y = sym_inertia(M, i);
    Force<Sym> f = Y * ai;
    f += fcross(vi, Y * vi);
    // a joint with children waits for them: its own force and its liMi generators are parked per depth
    if (topo.nchild[i] > 0)
    {
      std::vector<Sym> x = {f.lin.x, f.lin.y, f.lin.z, f.ang.x, f.ang.y, f.ang.z, si, ci};
      x.insert(x.end(), aj.begin(), aj.end());
      fpark[i] = T.park(x, "rnea own f");
    }
    // ---- backward steps of every joint whose subtree is complete ----
    const int stop = topo.stop[i];
    Force<Sym> fc = f; // force of joint j including its subtree
    for (int j = i; j != stop; j = M.parent[j])
    {
      T.comment("backward, joint " + std::to_string(j));
      const int tj = M.type[j], pj = M.parent[j], nv_j = M.nvj[j], ivj = M.idx_v[j];
      Sym sj = si, cj = ci;
      std::vector<Sym> ajj = aj;
      if (j != i)
      {
        std::vector<Sym> x = T.fetch(fpark[j]);
        T.release(fpark[j]);
        Force<Sym> fo;
        fo.lin = Vec3<Sym>(x[0], x[1], x[2]); fo.ang = Vec3<Sym>(x[3], x[4], x[5]);
        sj = x[6]; cj = x[7];
        ajj.assign(x.begin() + 8, x.end());
        if (topo.nchild[j] >= 2)
        {
          std::vector<Sym> acc = T.fetch(open[j].f);
          T.release(open[j].f);
          fc.lin = fo.lin + Vec3<Sym>(acc[0], acc[1], acc[2]);
          fc.ang = fo.ang + Vec3<Sym>(acc[3], acc[4], acc[5]);
        }
        else
          fc = fo + fc; // fc holds the only child's force, already in this joint's frame
      }
      for (int k = 0; k < nv_j; ++k)
        T.output(OUT_MAIN, ivj + k, get6(fc, joint_S_row(tj, k)) + Sym(M.armature[ivj + k]) * ajj[k]);
      if (pj > 0)
      {
        std::vector<Sym> qjj;
        if (M.nvj[j] > 1 || tj > J_PZ)
        {
          const int nq2 = (j + 1 < nj ? M.idx_q[j + 1] : M.nq) - M.idx_q[j];
          for (int k = 0; k < nq2; ++k) qjj.push_back(T.in(IN_Q, M.idx_q[j] + k, j == i ? 0 : 1));
        }
        const SE3<Sym> Xlj = (j == i) ? Xl : sym_liMi(M, j, sj, cj, qjj);
        const Force<Sym> fp = Xlj.act(fc);
        if (topo.nchild[pj] >= 2)
        {
          std::vector<Sym> acc = {fp.lin.x, fp.lin.y, fp.lin.z, fp.ang.x, fp.ang.y, fp.ang.z};
          if (open[pj].has_f)
          {
            std::vector<Sym> old = T.fetch(open[pj].f);
            T.release(open[pj].f);
            for (int k = 0; k < 6; ++k) acc[k] += old[k];
          }
          open[pj].f = T.park(acc, "rnea branch f");
          open[pj].has_f = true;
        }
        else
          fc = fp;
      }
    }
  }
}

// ======================================================================================================================
// CRBA — column by column in joint-local frames.
// The reference's default convention is LOCAL (crba.hpp:51; crbaLocalConvention, crba.hxx:454-491 with the steps :234-309):
// Ycrb[parent] += liMi.act(Ycrb[i]), F[:, i] = Ycrb[i] S_i, M[i, subtree(i)] = S_i^T F[:, subtree(i)] after carrying the
// subtree's force columns into frame i.  Carrying ALL subtree columns up level by level needs 6 nvSubtree live values; here
// each column is finished at once instead: when joint j's composite inertia is complete, F = Ycrb_j S_j is walked up the root
// path (f <- liMi.act(f) per level) and leaves one entry per ancestor dof — the same products in another order, and the only
// long-lived state is (sin q, cos q) per tree depth plus one composite inertia per open branching joint.
// Output: row r of the current column through Tracer::output (the wrapper's staging row of this configuration), then
// Tracer::flush(column); rows outside the tree sparsity are zeros (a fresh Data, data.hxx:43).
// ======================================================================================================================
// `group`: up to that many ADJACENT columns (the unwind of a chain emits columns c, c - 1, c - 2, ...) share one flush, so that
// the wrapper writes group * nv contiguous elements per configuration at a time instead of nv (DRAM page locality of the store
// stream); staging offsets and flush ranges are then relative to the group's first column.
//
// The structural pattern of the result (static: the dofs of the column's own joint and of its ancestors) and the column groups.
// COMPACT staging keeps only the pattern's entries of a group in the staging row, in column-major order; the wrapper's flush
// puts the zeros back from the `pos` table (dense result) or writes the rows as they are (packed result, brbd_crba_packed_batch:
// entry k of a configuration = M[rows[k], cols[k]], column-major over the pattern whatever the grouping).
struct CrbaPattern
{
  int nv = 0, nnz = 0, maxrow = 0;
  std::vector<int> order;               // the columns in the order the sweep finishes them
  std::vector<int> g_lo, g_hi, g_last;  // per column: first / last column of its group, "closes the group"
  std::vector<int> pos;                 // [col * nv + row]: position inside the group's compact row, -1 outside the pattern
  std::vector<int> gbase, gnnz;         // per group, indexed by its first column: packed index of its first entry, entries
  std::vector<int> rows, cols;          // the packed order
  std::vector<std::vector<char>> staged; // per group in the order the sweep closes them: the positions (column - lo) * nv + row it stages
};
// `budget` > 0: a group also closes before its pattern would exceed that many entries (compact staging: the rows of all groups
// then have about the same length, which is what sizes the warp's staging tile)
inline CrbaPattern crba_pattern(const ModelPOD<double> & M, int group, int budget = 0)
{
  CrbaPattern P;
  const int nj = M.njoints, nv = M.nv;
  const JointTopo topo(M);
  P.nv = nv;
  P.g_lo.assign(nv, 0); P.g_hi.assign(nv, 0); P.g_last.assign(nv, 0);
  P.pos.assign((size_t)nv * nv, -1); P.gbase.assign(nv, 0); P.gnnz.assign(nv, 0);
  for (int i = 1; i < nj; ++i)
    for (int j = i; j != topo.stop[i]; j = M.parent[j])
      for (int k = 0; k < M.nvj[j]; ++k) P.order.push_back(M.idx_v[j] + k);
  // pattern of a column: the full diagonal block of its joint (M.block(idx_v, idx_v, nv, nvSubtree) = S^T F writes it whole)
  // and the dofs of every ancestor
  std::vector<char> nz((size_t)nv * nv, 0);
  std::vector<int> colnnz(nv, 0);
  for (int j = 1; j < nj; ++j)
    for (int k = 0; k < M.nvj[j]; ++k)
    {
      const int col = M.idx_v[j] + k;
      for (int kk = 0; kk < M.nvj[j]; ++kk) nz[(size_t)col * nv + M.idx_v[j] + kk] = 1;
      for (int a = j; M.parent[a] > 0; a = M.parent[a])
        for (int kk = 0; kk < M.nvj[M.parent[a]]; ++kk) nz[(size_t)col * nv + M.idx_v[M.parent[a]] + kk] = 1;
      for (int r = 0; r < nv; ++r) colnnz[col] += nz[(size_t)col * nv + r];
    }
  for (size_t a = 0; a < P.order.size();)
  {
    int lo = P.order[a], hi = P.order[a], entries = colnnz[P.order[a]];
    size_t b = a + 1;
    while (b < P.order.size() && (int)(b - a) < group && (P.order[b] == lo - 1 || P.order[b] == hi + 1) &&
           (budget <= 0 || entries + colnnz[P.order[b]] <= budget))
    {
      lo = std::min(lo, P.order[b]); hi = std::max(hi, P.order[b]);
      entries += colnnz[P.order[b]];
      ++b;
    }
    for (size_t c = a; c < b; ++c) { P.g_lo[P.order[c]] = lo; P.g_hi[P.order[c]] = hi; P.g_last[P.order[c]] = (c + 1 == b); }
    a = b;
  }
  for (int c : P.order)
    if (P.g_last[c])
    {
      std::vector<char> rows((size_t)nv * std::max(1, group), 0);
      for (int cc = P.g_lo[c]; cc <= P.g_hi[c]; ++cc)
        for (int r = 0; r < nv; ++r)
          if (nz[(size_t)cc * nv + r]) rows[(size_t)(cc - P.g_lo[c]) * nv + r] = 1;
      P.staged.push_back(rows);
    }
  for (int col = 0; col < nv;)
  { // groups are disjoint runs of columns; walk them in column order: the packed order is column-major over the pattern
    const int lo = P.g_lo[col], hi = P.g_hi[col];
    int n = 0;
    for (int c = lo; c <= hi; ++c)
      for (int r = 0; r < nv; ++r)
        if (nz[(size_t)c * nv + r])
        {
          P.pos[(size_t)c * nv + r] = n++;
          P.rows.push_back(r); P.cols.push_back(c);
        }
    P.gbase[lo] = P.nnz; P.gnnz[lo] = n;
    P.nnz += n;
    P.maxrow = std::max(P.maxrow, n);
    col = hi + 1;
  }
  return P;
}

inline void trace_crba(Tracer & T, int nbuf = 1, int group = 1, bool compact = false, int budget = 0)
{
  const ModelPOD<double> & M = T.M;
  const int nj = M.njoints;
  const JointTopo topo(M);
  struct SC { Sym s, c; std::vector<Sym> q; };
  std::vector<SC> sc(nj);
  std::vector<Inertia<Sym>> Yacc(nj);
  std::vector<char> has_acc(nj, 0);
  // the wrapper may rotate over `nbuf` staging rows: the one a column is assembled in still holds the column of nbuf flushes ago
  std::vector<std::vector<char>> prev_rows(nbuf, std::vector<char>((size_t)M.nv * group, 0));
  int ncol = 0;
  // the column order is static: partition it into groups of adjacent columns before tracing
  const CrbaPattern pat = crba_pattern(M, group, budget);
  const std::vector<int> & g_lo = pat.g_lo, & g_hi = pat.g_hi, & g_last = pat.g_last;
  // the body runs once per round of the persistent grid: when it starts again, the staging rows still hold the LAST nbuf groups
  // of the previous configuration (in the first round they hold zeros and the clears are idle)
  for (int k = 0; k < nbuf && !compact; ++k)
  {
    const int n = (int)pat.staged.size();
    const int g = ((n - nbuf + k) % n + n) % n;
    if (n > 0) prev_rows[k] = pat.staged[g];
  }
  std::vector<char> cur_rows((size_t)M.nv * group, 0);
  bool group_open = false;
  auto liMi_of = [&](int a) { return sym_liMi(M, a, sc[a].s, sc[a].c, sc[a].q); };
  auto S_col = [&](int type, int k) { return joint_S_col<Sym>(type, k); };
  for (int i = 1; i < nj; ++i)
  {
    T.comment("joint " + std::to_string(i));
    {
      const int type = M.type[i];
      const int nqj = (i + 1 < nj ? M.idx_q[i + 1] : M.nq) - M.idx_q[i];
      sc[i].q.clear();
      for (int k = 0; k < nqj; ++k) sc[i].q.push_back(T.in(IN_Q, M.idx_q[i] + k));
      sc[i].s = Sym(0.0); sc[i].c = Sym(0.0);
      if (type <= J_RZ && M.unb[i]) { sc[i].c = sc[i].q[0]; sc[i].s = sc[i].q[1]; }
      else if (type <= J_RZ) sincos_t(sc[i].q[0], &sc[i].s, &sc[i].c);
      else if (type <= J_PZ) sc[i].s = sc[i].q[0];
    }
    const int stop = topo.stop[i];
    for (int j = i; j != stop; j = M.parent[j])
    {
      const int tj = M.type[j], pj = M.parent[j], nvj = M.nvj[j], iv = M.idx_v[j];
      Inertia<Sym> Y = sym_inertia(M, j);
      if (has_acc[j]) Y += Yacc[j];
      for (int k = 0; k < nvj; ++k)
      {
        const int col = iv + k;
        const int off = (col - g_lo[col]) * M.nv; // position of this column in the group's staging row
        if (!group_open)
        {
          T.colbegin(g_lo[col]);
          if (!compact) // compact rows hold entries of the pattern only, every one of them rewritten
            for (int r = 0; r < M.nv * group; ++r)
              if (prev_rows[ncol % nbuf][r]) T.clear(r);
          std::fill(cur_rows.begin(), cur_rows.end(), 0);
          group_open = true;
        }
        std::vector<char> & rows = cur_rows;
        Force<Sym> f = Y * S_col(tj, k);
        for (int kk = 0; kk < nvj; ++kk)
        { // the joint's own diagonal block (full, as M.block(idx_v, idx_v, nv, nvSubtree) = S^T F writes it)
          Sym val = dot6(S_col(tj, kk), f);
          if (kk == k) val += Sym(M.armature[col]);
          T.output(OUT_MAIN, compact ? pat.pos[(size_t)col * M.nv + iv + kk] : off + iv + kk, val);
          rows[off + iv + kk] = 1;
        }
        for (int a = j; M.parent[a] > 0; a = M.parent[a])
        {
          const int pa = M.parent[a];
          f = liMi_of(a).act(f); // into the parent's frame
          for (int kk = 0; kk < M.nvj[pa]; ++kk)
          {
            T.output(OUT_MAIN, compact ? pat.pos[(size_t)col * M.nv + M.idx_v[pa] + kk] : off + M.idx_v[pa] + kk, dot6(S_col(M.type[pa], kk), f));
            rows[off + M.idx_v[pa] + kk] = 1;
          }
        }
        if (g_last[col])
        {
          T.flush(g_lo[col] | ((g_hi[col] - g_lo[col] + 1) << 16)); // first column | number of columns << 16
          prev_rows[ncol % nbuf] = rows;
          ++ncol;
          group_open = false;
        }
      }
      if (pj > 0)
      {
        const Inertia<Sym> Yp = act(liMi_of(j), Y);
        if (has_acc[pj]) Yacc[pj] += Yp;
        else { Yacc[pj] = Yp; has_acc[pj] = 1; }
      }
    }
  }
}

} // namespace cg
} // namespace brbd
