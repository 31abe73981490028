// tree.cuh — the device-side flattened Model of the v2 ("DFS-interleaved") kernels.
//
// TreePOD is passed to every kernel BY VALUE as a __grid_constant__ parameter, so the topology and the
// constants (parents, joint type tags, placements, inertias) sit in the constant bank and every access is
// a warp-uniform LDC — no shared memory is spent on the model.  It replaces, for the batched path,
// ModelTpl (reference: include/pinocchio/multibody/model.hpp:97-205) and the index tables of DataTpl
// (multibody/data.hxx:197-315).
//
// The sweeps of all algorithms are DFS-interleaved: joints are numbered depth-first (CRBAChecker,
// algorithm/crba.hxx:573-595), so after the forward step of joint i the backward steps of every joint
// whose subtree is now complete (i, parent(i), ... up to but excluding parent(i+1)) can run at once.
// Live per-configuration state is then proportional to the tree DEPTH, not to the number of joints, and
// fits in shared memory, laid out slot-major ([slot][thread]) so that a warp's access to one slot is one
// conflict-free 256-byte row.
#pragma once

#include <stdint.h>

#include "engine.cuh"

namespace brbd
{

constexpr int MAXPATH = 32; // dofs on one root path

struct JointRec
{
  short type;    // JointTag
  short parent;  // 0 = universe
  short idx_q, idx_v, nvj;
  short depth;   // universe = 0, root joints = 1
  short pdof;    // number of dofs owned by the strict ancestors (offset of this joint on the root path)
  short bslot;   // >= 0: this joint has >= 2 children and owns branch slot `bslot`; -1 otherwise
  short stop;    // parent[i + 1] (0 for the last joint): the unwind after fwd(i) stops there
  short nchild;
  short poff;    // offset of this joint in the per-configuration persistent store (ABA)
  short unb;     // 1: unbounded revolute joint, q = (cos, sin)
};

template<class T> struct TreePOD
{
  int njoints, nq, nv, maxdepth;
  int maxpathdof; // max over joints of pdof + nvj
  int nbranch;    // number of branch slots (max number of branching joints on one root path)
  int pslots;     // persistent-store slots per configuration (ABA)
  int ffroot;     // 1: joint 1 is a free-flyer and the only root joint (every root path starts with its 6 dofs)
  JointRec j[MAXJ];
  unsigned long long anc_mask[MAXNV]; // bit r set: row r belongs to an ancestor-or-self of the joint owning this column
  unsigned char path_row[MAXJ][MAXPATH]; // path_row[j][t]: tangent row of path dof t on the root path of joint j
  T placement[MAXJ][12];              // R by columns then p
  T inertia[MAXJ][10];                // m, c, (xx,xy,yy,xz,yz,zz)
  T armature[MAXNV];
  T gravity[3];
};

// host: derive the v2 tables from the validated v1 POD
template<class T> inline void build_tree(const ModelPOD<double> & M, TreePOD<T> & P)
{
  memset(&P, 0, sizeof(P));
  P.njoints = M.njoints; P.nq = M.nq; P.nv = M.nv; P.maxdepth = M.maxdepth;
  int nchild[MAXJ] = {0};
  for (int i = 1; i < M.njoints; ++i) nchild[M.parent[i]]++;
  int bdepth[MAXJ] = {0}; // branching strict ancestors
  int pdof[MAXJ] = {0};
  int poff = 0;
  for (int i = 1; i < M.njoints; ++i)
  {
    const int p = M.parent[i];
    JointRec & r = P.j[i];
    r.type = (short)M.type[i]; r.parent = (short)p; r.idx_q = (short)M.idx_q[i]; r.idx_v = (short)M.idx_v[i];
    r.nvj = (short)M.nvj[i]; r.depth = (short)M.depth[i]; r.unb = (short)M.unb[i];
    pdof[i] = p > 0 ? pdof[p] + M.nvj[p] : 0;
    r.pdof = (short)pdof[i];
    bdepth[i] = p > 0 ? bdepth[p] + (nchild[p] >= 2 ? 1 : 0) : 0;
    r.bslot = (short)(nchild[i] >= 2 ? bdepth[i] : -1);
    r.stop = (short)(i + 1 < M.njoints ? M.parent[i + 1] : 0);
    r.nchild = (short)nchild[i];
    r.poff = (short)poff;
    const int n = M.nvj[i];
    poff += 6 * n + 6 + 6 * n + n * n + n; // J, ab, UDinv, Dinv, u
    if (pdof[i] + n > P.maxpathdof) P.maxpathdof = pdof[i] + n;
    if (nchild[i] >= 2 && bdepth[i] + 1 > P.nbranch) P.nbranch = bdepth[i] + 1;
  }
  P.pslots = poff;
  P.ffroot = (M.njoints > 1 && M.type[1] == J_FF && M.parent[1] == 0) ? 1 : 0;
  for (int i = 2; i < M.njoints; ++i)
    if (M.parent[i] == 0) P.ffroot = 0;
  for (int i = 1; i < M.njoints; ++i)
  {
    unsigned long long mask = 0;
    for (int a = i; a > 0; a = M.parent[a])
      for (int k = 0; k < M.nvj[a]; ++k) mask |= 1ull << (M.idx_v[a] + k);
    for (int k = 0; k < M.nvj[i]; ++k) P.anc_mask[M.idx_v[i] + k] = mask;
    for (int a = i; a > 0; a = M.parent[a])
      for (int k = 0; k < M.nvj[a]; ++k)
        if (pdof[a] + k < MAXPATH) P.path_row[i][pdof[a] + k] = (unsigned char)(M.idx_v[a] + k);
  }
  for (int i = 0; i < MAXJ; ++i)
  {
    for (int k = 0; k < 12; ++k) P.placement[i][k] = (T)M.placement[i][k];
    for (int k = 0; k < 10; ++k) P.inertia[i][k] = (T)M.inertia[i][k];
  }
  for (int k = 0; k < MAXNV; ++k) P.armature[k] = (T)M.armature[k];
  for (int k = 0; k < 3; ++k) P.gravity[k] = (T)M.gravity[k];
}

// geometry of the TMA tensor stores of a (nv*nv x B) matrix block, see crba_dfs.cuh (crba_tma_kernel)
struct CrbaTmaGeom
{
  int bx;    // box inner extent = emitter row length (elements): nv, or nv + 1 for odd nv
  int odd;   // odd nv (FP64): shifted boxes
  int pairs; // odd nv and odd ldM: even / odd configurations over two maps
};

// ---- per-thread view of the slot-major shared state ------------------------------------------------
// NT = threads per CTA is a template parameter so that slot offsets fold into the LDS/STS immediates.
template<class T, int NT> struct Slots
{
  T * p; // &smem[threadIdx.x]
  BRBD_DI T & operator[](int slot) const { return p[slot * NT]; }
};
template<class T, class S> BRBD_DI void put3(const S & s, int o, const Vec3<T> & v) { s[o] = v.x; s[o + 1] = v.y; s[o + 2] = v.z; }
template<class T, class S> BRBD_DI Vec3<T> get3(const S & s, int o) { return Vec3<T>(s[o], s[o + 1], s[o + 2]); }
template<class T, class S> BRBD_DI void put_se3(const S & s, int o, const SE3<T> & X)
{
  put3<T>(s, o, X.R.c0); put3<T>(s, o + 3, X.R.c1); put3<T>(s, o + 6, X.R.c2); put3<T>(s, o + 9, X.p);
}
template<class T, class S> BRBD_DI SE3<T> get_se3(const S & s, int o)
{
  SE3<T> X;
  X.R.c0 = get3<T>(s, o); X.R.c1 = get3<T>(s, o + 3); X.R.c2 = get3<T>(s, o + 6); X.p = get3<T>(s, o + 9);
  return X;
}
template<class T, class S> BRBD_DI void put_motion(const S & s, int o, const Motion<T> & m) { put3<T>(s, o, m.lin); put3<T>(s, o + 3, m.ang); }
template<class T, class S> BRBD_DI Motion<T> get_motion(const S & s, int o) { Motion<T> m; m.lin = get3<T>(s, o); m.ang = get3<T>(s, o + 3); return m; }
template<class T, class S> BRBD_DI void put_force(const S & s, int o, const Force<T> & m) { put3<T>(s, o, m.lin); put3<T>(s, o + 3, m.ang); }
template<class T, class S> BRBD_DI Force<T> get_force(const S & s, int o) { Force<T> m; m.lin = get3<T>(s, o); m.ang = get3<T>(s, o + 3); return m; }
template<class T, class S> BRBD_DI void put_inertia(const S & s, int o, const Inertia<T> & Y)
{
  s[o] = Y.m; put3<T>(s, o + 1, Y.c);
  s[o + 4] = Y.I.xx; s[o + 5] = Y.I.xy; s[o + 6] = Y.I.yy; s[o + 7] = Y.I.xz; s[o + 8] = Y.I.yz; s[o + 9] = Y.I.zz;
}
template<class T, class S> BRBD_DI Inertia<T> get_inertia(const S & s, int o)
{
  Inertia<T> Y;
  Y.m = s[o]; Y.c = get3<T>(s, o + 1);
  Y.I.xx = s[o + 4]; Y.I.xy = s[o + 5]; Y.I.yy = s[o + 6]; Y.I.xz = s[o + 7]; Y.I.yz = s[o + 8]; Y.I.zz = s[o + 9];
  return Y;
}

// ---- per-thread asynchronous prefetch: one element global -> this thread's shared slot (LDGSTS) -------------
// Prefetching into registers does not survive the register pressure of the sweeps (the compiler spills the
// loaded value at once, which waits for the load); cp.async needs no register and no scoreboard wait.
template<class T> BRBD_DI void async_fetch(T * smem_dst, const T * __restrict__ gsrc)
{
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  if constexpr (sizeof(T) == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
  else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
}
BRBD_DI void async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
BRBD_DI void async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
template<int N> BRBD_DI void async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- constants of joint i out of the parameter bank ---------------------------------------------------
template<class T> BRBD_DI SE3<T> tree_placement(const TreePOD<T> & m, int i)
{
  const T * P = m.placement[i];
  SE3<T> X;
  X.R.c0 = Vec3<T>(P[0], P[1], P[2]);
  X.R.c1 = Vec3<T>(P[3], P[4], P[5]);
  X.R.c2 = Vec3<T>(P[6], P[7], P[8]);
  X.p = Vec3<T>(P[9], P[10], P[11]);
  return X;
}
template<class T> BRBD_DI Inertia<T> tree_inertia(const TreePOD<T> & m, int i)
{
  const T * Y = m.inertia[i];
  Inertia<T> I;
  I.m = Y[0];
  I.c = Vec3<T>(Y[1], Y[2], Y[3]);
  I.I.xx = Y[4]; I.I.xy = Y[5]; I.I.yy = Y[6]; I.I.xz = Y[7]; I.I.yz = Y[8]; I.I.zz = Y[9];
  return I;
}

// liMi = jointPlacements[i] * M_J(q) with the structural zeros of M_J dropped (same arithmetic as
// engine.cuh joint_liMi; reference rnea.hxx:61, aba.hxx:117, crba.hxx:47 + the joints' calc()).
// qj points at this configuration's q segment in GLOBAL memory.
// liMi of joint i from (s, c): revolute (s, c) = sincos(q); prismatic s = q; multi-dof joints read q from global memory
// (s = first coordinate).  Kernels that need liMi again in their backward sweep keep (s, c) instead of the 12 values.
template<class T> BRBD_DI SE3<T> tree_liMi_sc(const TreePOD<T> & m, int i, int type, const T * __restrict__ qj, T s, T c)
{
  const SE3<T> P = tree_placement(m, i);
  SE3<T> X;
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
    SE3<T> MJ;
    MJ.R = quat_to_mat(__ldg(qj + 3), __ldg(qj + 4), __ldg(qj + 5), __ldg(qj + 6));
    MJ.p = Vec3<T>(s, __ldg(qj + 1), __ldg(qj + 2));
    X = P * MJ;
  }
  else if (type == J_SPH)
  {
    X.R = P.R * quat_to_mat(s, __ldg(qj + 1), __ldg(qj + 2), __ldg(qj + 3));
    X.p = P.p;
  }
  else
  { // planar: q = (x, y, cos, sin)
    const T cc = __ldg(qj + 2), ss = __ldg(qj + 3);
    X.R.c0 = cc * P.R.c0 + ss * P.R.c1;
    X.R.c1 = cc * P.R.c1 - ss * P.R.c0;
    X.R.c2 = P.R.c2;
    X.p = P.p + s * P.R.c0 + __ldg(qj + 1) * P.R.c1;
  }
  return X;
}
// (s, c) of joint i for tree_liMi_sc
template<class T> BRBD_DI void tree_sc(int type, T q0, T * s, T * c)
{
  if (type <= J_RZ) sincos_t(q0, s, c);
  else { *s = q0; *c = T(0); }
}
// (s, c) of joint i, also for the unbounded revolute joints whose configuration IS (cos, sin)
template<class T> BRBD_DI void tree_sc_joint(const TreePOD<T> & m, int i, int type, const T * __restrict__ qj, T q0, T * s, T * c)
{
  if (m.j[i].unb) { *c = q0; *s = __ldg(qj + 1); }
  else tree_sc(type, q0, s, c);
}
template<class T> BRBD_DI SE3<T> tree_liMi(const TreePOD<T> & m, int i, int type, const T * __restrict__ qj, T q0)
{
  T s, c;
  tree_sc_joint(m, i, type, qj, q0, &s, &c);
  return tree_liMi_sc(m, i, type, qj, s, c);
}
template<class T> BRBD_DI SE3<T> tree_liMi(const TreePOD<T> & m, int i, int type, const T * __restrict__ qj)
{
  return tree_liMi(m, i, type, qj, __ldg(qj));
}

// joint velocity v_J = S qdot with qdot read from global memory
template<class T> BRBD_DI Motion<T> tree_joint_velocity(int type, const T * __restrict__ vj)
{
  Motion<T> v = mzero<T>();
  if (type <= J_RZ) v.ang.set(type - J_RX, __ldg(vj));
  else if (type <= J_PZ) v.lin.set(type - J_PX, __ldg(vj));
  else if (type == J_FF) { v.lin = Vec3<T>(__ldg(vj), __ldg(vj + 1), __ldg(vj + 2)); v.ang = Vec3<T>(__ldg(vj + 3), __ldg(vj + 4), __ldg(vj + 5)); }
  else if (type == J_SPH) v.ang = Vec3<T>(__ldg(vj), __ldg(vj + 1), __ldg(vj + 2));
  else { v.lin = Vec3<T>(__ldg(vj), __ldg(vj + 1), T(0)); v.ang = Vec3<T>(T(0), T(0), __ldg(vj + 2)); }
  return v;
}

template<class T> BRBD_DI void add6(Motion<T> & m, int row, T val)
{
  if (row < 3) m.lin.set(row, m.lin.get(row) + val); else m.ang.set(row - 3, m.ang.get(row - 3) + val);
}

} // namespace brbd
