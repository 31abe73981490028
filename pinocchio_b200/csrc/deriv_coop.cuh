// deriv_coop.cuh — batched computeRNEADerivatives, warp-cooperative: G lanes work on ONE configuration.
//
// Why not one configuration per thread (rnea_derivatives.cuh, v1): the backward sweep of
// impl::computeRNEADerivatives (reference: include/pinocchio/algorithm/rnea-derivatives.hxx:378-459) needs, for
// every column, the J / dFda / dYtJ columns of the whole subtree and the root path — about 25 KB of FP64 state
// per configuration for a humanoid.  Per thread that is thread-local memory (v1: 6.9 ms for 65 536
// configurations of simple_humanoid, 4 % of the HBM roofline); per warp it fits in shared memory, and the
// work inside one configuration is wide enough for 32 lanes:
//
//   phase 1  forward kinematics by pointer jumping along the root paths (lanes = joints, log2(depth) rounds):
//            oMi, J = oMi.act(S), ov, oa_gf                       rnea-derivatives.hxx:280-317 (world-frame form)
//   phase 2  lanes = joints: oYcrb = oMi.act(I), oh, of, doYcrb    :303-317, :335-351
//   phase 3  lanes = components: subtree sums of (oYcrb, doYcrb, of) — the `[parent] +=` of :452-456; the
//            inertia is kept in its linear form (m, m c, I about the origin) so that the sum is a plain sum
//   phase 4  lanes = tangent columns: dVdq, dAdq, dAdv (:319-332), dFda, dFdq, dFdv, dYtJ (:412-431), tau (:409)
//   phase 5  lanes = matrix rows: every entry of dtau_dq / dtau_dv / dtau_da is one or two 6-D dot products of
//            a row record and a column record (:420-451); rows of a column are contiguous in the caller's
//            col-major matrix, so each column leaves with one coalesced store per matrix, zeros included
//            (the reference requires pre-zeroed outputs, rnea-derivatives.hpp:104-106).
//
// Small models use sub-warp groups (G = 8 or 16 lanes per configuration, 4 or 2 configurations per warp).
#pragma once

#include "engine.cuh"
#include "rnea.cuh"
#include "rnea_derivatives.cuh"

namespace brbd
{

// lane-varying lookups (kept in shared memory; the constant bank serialises divergent indices)
struct CoopTables
{
  short lvl_start[MAXDEPTH + 2]; // joints of depth l are lvl_joint[lvl_start[l] .. lvl_start[l + 1])
  short lvl_joint[MAXJ];
  unsigned long long anc_mask[MAXNV]; // bit r: row r belongs to an ancestor-or-self joint of the joint owning this column
  unsigned int colinfo[MAXNV];        // idx_v | (idx_v + nv) << 8 | (idx_v + nvSubtree) << 16 of the joint owning this column
  short anc[5][MAXJ];                 // anc[s][i] = 2^s-th ancestor of joint i (0 = none): pointer jumping along the root path
  int nsteps;                         // number of pointer-jumping rounds = ceil(log2(maxdepth))
  short jlast[MAXJ];                  // last joint of the subtree of i (depth-first numbering: subtree = [i, jlast[i]])
  short bslot[MAXJ];                  // save slot of a joint with two or more children (-1 otherwise), see aba_deriv_coop.cuh
  int nbranch;                        // number of such joints
  unsigned lvl_multi;                 // bit l: depth l has a joint with more than one degree of freedom
  unsigned jinfo[MAXJ];               // idx_v | nv_joint << 8 | (idx_v + nvSubtree) << 16 | (parent > 0) << 24
  unsigned rinfo[MAXNV];              // per tangent row: idx_v | own_end << 8 | sub_end << 16 | flags << 24 (1 first row of its joint, 2 last, 4 parent > 0)
  unsigned short rslots[MAXNV];       // per tangent row: load slot (first row, parent is not the previous joint) | save slot << 8 (last row); 0xff = none
};
inline void build_coop_tables(const ModelPOD<double> & M, CoopTables & C)
{
  memset(&C, 0, sizeof(C));
  int n = 0;
  for (int l = 1; l <= M.maxdepth; ++l)
  {
    C.lvl_start[l] = (short)n;
    for (int i = 1; i < M.njoints; ++i)
      if (M.depth[i] == l) C.lvl_joint[n++] = (short)i;
  }
  for (int l = M.maxdepth + 1; l < MAXDEPTH + 2; ++l) C.lvl_start[l] = (short)n;
  for (int i = 1; i < M.njoints; ++i)
  {
    unsigned long long mask = 0;
    for (int a = i; a > 0; a = M.parent[a])
      for (int k = 0; k < M.nvj[a]; ++k) mask |= 1ull << (M.idx_v[a] + k);
    for (int k = 0; k < M.nvj[i]; ++k)
    {
      C.anc_mask[M.idx_v[i] + k] = mask;
      C.colinfo[M.idx_v[i] + k] = (unsigned)M.idx_v[i] | ((unsigned)(M.idx_v[i] + M.nvj[i]) << 8) | ((unsigned)(M.idx_v[i] + M.nvsub[i]) << 16);
    }
    C.anc[0][i] = (short)M.parent[i];
  }
  for (int i = 0; i < M.njoints; ++i) C.jlast[i] = (short)i;
  for (int i = M.njoints - 1; i > 0; --i)
    if (C.jlast[i] > C.jlast[M.parent[i]]) C.jlast[M.parent[i]] = C.jlast[i];
  C.nbranch = 0;
  for (int i = 0; i < M.njoints; ++i)
  {
    int nchild = 0;
    for (int k = i + 1; k < M.njoints; ++k) nchild += (M.parent[k] == i);
    C.bslot[i] = (short)((i > 0 && nchild >= 2) ? C.nbranch++ : -1);
  }
  C.lvl_multi = 0;
  for (int i = 1; i < M.njoints; ++i)
  {
    const int iv = M.idx_v[i], nvj = M.nvj[i], p = M.parent[i];
    if (nvj > 1) C.lvl_multi |= 1u << M.depth[i];
    C.jinfo[i] = (unsigned)iv | ((unsigned)nvj << 8) | ((unsigned)(iv + M.nvsub[i]) << 16) | ((unsigned)(p > 0) << 24);
    for (int k = 0; k < nvj; ++k)
    {
      const unsigned flags = (k == 0 ? 1u : 0u) | (k == nvj - 1 ? 2u : 0u) | (p > 0 ? 4u : 0u);
      C.rinfo[iv + k] = (unsigned)iv | ((unsigned)(iv + nvj) << 8) | ((unsigned)(iv + M.nvsub[i]) << 16) | (flags << 24);
      const unsigned lslot = (k == 0 && p > 0 && p != i - 1) ? (unsigned)C.bslot[p] : 0xffu;
      const unsigned sslot = (k == nvj - 1 && C.bslot[i] >= 0) ? (unsigned)C.bslot[i] : 0xffu;
      C.rslots[iv + k] = (unsigned short)(lslot | (sslot << 8));
    }
  }
  C.nsteps = 0;
  while ((1 << C.nsteps) < M.maxdepth) ++C.nsteps;
  for (int s = 1; s < 5; ++s)
    for (int i = 1; i < M.njoints; ++i) C.anc[s][i] = C.anc[s - 1][C.anc[s - 1][i]];
}

// Per-joint record (JR_STRIDE values, odd => lanes = joints are conflict-free):
//   [0] m  [1..3] h = m c  [4..9] I about the world origin (xx,xy,yy,xz,yz,zz)   <- summed over the subtree
//   [10..36] doYcrb (LA, AL, AA)                                                 <- summed (oMi sits here until phase 2)
//   [37..42] of                                                                  <- summed
//   [43..48] ov   [49..54] oa_gf
constexpr int JR_STRIDE = 55, JR_DY = 10, JR_OF = 37, JR_OV = 43, JR_OA = 49, JR_NSUM = 43;
// Per-column record (CB_STRIDE values, even => 16-byte aligned pairs):
constexpr int CB_STRIDE = 54, CB_DFDQ = 0, CB_DFDQP = 6, CB_DFDV = 12, CB_DFDA = 18, CB_DADQ = 24, CB_DVDQ = 30,
              CB_DADV = 36, CB_J = 42, CB_DYTJ = 48;

struct CoopLayout
{
  int oq, ov, oa, ojr, ocb; // offsets (elements) inside one group's region
  int per_group;            // elements, even
};
inline CoopLayout coop_layout(int nq, int nv, int nj)
{
  CoopLayout L;
  L.ocb = 0;
  L.ojr = L.ocb + CB_STRIDE * nv;
  L.oq = L.ojr + JR_STRIDE * nj;
  L.ov = L.oq + nq;
  L.oa = L.ov + nv;
  L.per_group = (L.oa + nv + 1) & ~1;
  return L;
}

template<class T> struct Pair;
template<> struct Pair<double> { typedef double2 type; };
template<> struct Pair<float> { typedef float2 type; };
// 6 values from an address that is aligned to 2 elements
template<class T> BRBD_DI void ld6(const T * p, T * x)
{
  typedef typename Pair<T>::type P;
  const P * q = reinterpret_cast<const P *>(p);
  const P a = q[0], b = q[1], c = q[2];
  x[0] = a.x; x[1] = a.y; x[2] = b.x; x[3] = b.y; x[4] = c.x; x[5] = c.y;
}
template<class T> BRBD_DI void ld4(const T * p, T * x)
{
  typedef typename Pair<T>::type P;
  const P * q = reinterpret_cast<const P *>(p);
  const P a = q[0], b = q[1];
  x[0] = a.x; x[1] = a.y; x[2] = b.x; x[3] = b.y;
}
template<class T> BRBD_DI void st6(T * p, const T * x)
{
  typedef typename Pair<T>::type P;
  P * q = reinterpret_cast<P *>(p);
  P a, b, c;
  a.x = x[0]; a.y = x[1]; b.x = x[2]; b.y = x[3]; c.x = x[4]; c.y = x[5];
  q[0] = a; q[1] = b; q[2] = c;
}
template<class T> BRBD_DI void st6(T * p, const Motion<T> & m) { const T x[6] = {m.lin.x, m.lin.y, m.lin.z, m.ang.x, m.ang.y, m.ang.z}; st6(p, x); }
template<class T> BRBD_DI void st6(T * p, const Force<T> & m) { const T x[6] = {m.lin.x, m.lin.y, m.lin.z, m.ang.x, m.ang.y, m.ang.z}; st6(p, x); }
template<class T> BRBD_DI Motion<T> ld6m(const T * p) { T x[6]; ld6(p, x); Motion<T> m; m.lin = Vec3<T>(x[0], x[1], x[2]); m.ang = Vec3<T>(x[3], x[4], x[5]); return m; }
template<class T> BRBD_DI T dot6a(const T * a, const T * b)
{
  return a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3] + a[4] * b[4] + a[5] * b[5];
}

// composite inertia in linear form: (m, h = m c, Io = I_c - m [c]x^2)
template<class T> struct LinInertia
{
  T m;
  Vec3<T> h;
  Sym3<T> Io;
  BRBD_DI Force<T> operator*(const Motion<T> & v) const
  {
    Force<T> f;
    f.lin = m * v.lin + cross(v.ang, h);
    f.ang = Io.mul(v.ang) + cross(h, v.lin);
    return f;
  }
};
template<class T> BRBD_DI void store_lin_inertia(T * d, const Inertia<T> & Y)
{
  const T m = Y.m, cx = Y.c.x, cy = Y.c.y, cz = Y.c.z;
  d[0] = m; d[1] = m * cx; d[2] = m * cy; d[3] = m * cz;
  d[4] = Y.I.xx + m * (cy * cy + cz * cz);
  d[5] = Y.I.xy - m * cx * cy;
  d[6] = Y.I.yy + m * (cx * cx + cz * cz);
  d[7] = Y.I.xz - m * cx * cz;
  d[8] = Y.I.yz - m * cy * cz;
  d[9] = Y.I.zz + m * (cx * cx + cy * cy);
}
template<class T> BRBD_DI LinInertia<T> load_lin_inertia(const T * d)
{
  LinInertia<T> Y;
  Y.m = d[0]; Y.h = Vec3<T>(d[1], d[2], d[3]);
  Y.Io.xx = d[4]; Y.Io.xy = d[5]; Y.Io.yy = d[6]; Y.Io.xz = d[7]; Y.Io.yz = d[8]; Y.Io.zz = d[9];
  return Y;
}

// ---- phase 1: forward kinematics by pointer jumping along the root paths ---------------------------------
// oMi_i = prod_{a on path(i)} liMi_a, ov_i = sum_path J_a v_a, oa_gf_i = -g + sum_path (J_a a_a + ov_a x (J_a v_a))
// (the world-frame form of a_i = S a + v_i x v_J + liMi^-1 a_parent, oa_gf = oMi.act(a_i) - g, rnea-derivatives.hxx
// :295-317).  A level-by-level walk would issue the per-joint code once per tree level (11 times for a humanoid)
// with a handful of active lanes; pointer jumping composes every joint with its 2^s-th ancestor's partial
// product, all lanes active, in ceil(log2(depth)) rounds.  Buffers ping-pong inside the joint record:
//   oMi: [JR_DY, JR_DY+12) <-> [JR_DY+12, JR_DY+24);  ov: JR_OV <-> JR_OF;  oa: JR_OA <-> JR_OF.
// On return: oMi at jr[i] + xoff (returned), ov final at JR_OV, oa WITHOUT gravity at jr[i] + *oa_off.
template<class T, int G>
BRBD_DI void coop_scan6(const CoopTables & tb, T * jr, int nj, int gl, int oA, int oB)
{
  // prefix sums along the root path of a 6-vector; source oA, result ends in (nsteps even ? oA : oB)
  for (int s = 0; s < tb.nsteps; ++s)
  {
    const int src = (s & 1) ? oB : oA, dst = (s & 1) ? oA : oB;
    for (int i = 1 + gl; i < nj; i += G)
    {
      const int a = tb.anc[s][i];
      const T * ri = jr + i * JR_STRIDE + src;
      T x[6] = {ri[0], ri[1], ri[2], ri[3], ri[4], ri[5]};
      if (a > 0)
      {
        const T * ra = jr + a * JR_STRIDE + src;
#pragma unroll
        for (int k = 0; k < 6; ++k) x[k] += ra[k];
      }
      T * d = jr + i * JR_STRIDE + dst;
#pragma unroll
      for (int k = 0; k < 6; ++k) d[k] = x[k];
    }
    BRBD_SYNCWARP();
  }
}

template<class T, int G, bool WITH_ACC>
BRBD_DI int coop_forward(const ModelPOD<T> & m, const CoopTables & tb, const T * sq, const T * sv, const T * sa, T * jr, T * cb, int gl,
                         int * oa_off)
{
  const int nj = m.njoints, nsteps = tb.nsteps;
  // 1a: liMi
  for (int i = 1 + gl; i < nj; i += G)
    store_se3(jr + i * JR_STRIDE + JR_DY, joint_liMi(m, i, m.type[i], sq + m.idx_q[i]));
  BRBD_SYNCWARP();
  // 1b: oMi
  for (int s = 0; s < nsteps; ++s)
  {
    const int src = JR_DY + ((s & 1) ? 12 : 0), dst = JR_DY + ((s & 1) ? 0 : 12);
    for (int i = 1 + gl; i < nj; i += G)
    {
      const int a = tb.anc[s][i];
      SE3<T> X = load_se3(jr + i * JR_STRIDE + src);
      if (a > 0) X = load_se3(jr + a * JR_STRIDE + src) * X;
      store_se3(jr + i * JR_STRIDE + dst, X);
    }
    BRBD_SYNCWARP();
  }
  const int xoff = JR_DY + ((nsteps & 1) ? 12 : 0);
  // 1c: J = oMi.act(S), w = J v, u = J a
  for (int i = 1 + gl; i < nj; i += G)
  {
    const int type = m.type[i], iv = m.idx_v[i], nvj = m.nvj[i];
    T * r = jr + i * JR_STRIDE;
    const SE3<T> X = load_se3(r + xoff);
    Motion<T> w = mzero<T>(), u = mzero<T>();
    for (int kk = 0; kk < nvj; ++kk)
    {
      const Motion<T> J = act_S_col(X, type, kk);
      st6(cb + (iv + kk) * CB_STRIDE + CB_J, J);
      const T vk = sv[iv + kk];
      w.lin += vk * J.lin; w.ang += vk * J.ang;
      if (WITH_ACC)
      {
        const T ak = sa[iv + kk];
        u.lin += ak * J.lin; u.ang += ak * J.ang;
      }
    }
    store6(r, w);
    store6(r + JR_OV, w);
    if (WITH_ACC) store6(r + JR_OA, u);
  }
  BRBD_SYNCWARP();
  // 1d: ov
  coop_scan6<T, G>(tb, jr, nj, gl, JR_OV, JR_OF);
  const int ov_res = (nsteps & 1) ? JR_OF : JR_OV;
  if (WITH_ACC || ov_res != JR_OV)
  {
    // 1e: t = u + ov x w
    for (int i = 1 + gl; i < nj; i += G)
    {
      T * r = jr + i * JR_STRIDE;
      const Motion<T> ov = load_motion(r + ov_res);
      if (ov_res != JR_OV) store6(r + JR_OV, ov);
      if (WITH_ACC)
      {
        Motion<T> t = load_motion(r + JR_OA);
        t += mcross(ov, load_motion(r));
        store6(r + JR_OA, t);
      }
    }
    BRBD_SYNCWARP();
  }
  if (WITH_ACC)
  {
    coop_scan6<T, G>(tb, jr, nj, gl, JR_OA, JR_OF);
    *oa_off = (nsteps & 1) ? JR_OF : JR_OA;
  }
  return xoff;
}

// ---- phase 2: per-joint world inertia, momentum, force, inertia variation --------------------------------
template<class T, int G, bool SUB_GRAVITY = true>
BRBD_DI void coop_joint_quantities(const ModelPOD<T> & m, T * jr, int gl, int xoff, int oa_off)
{
  const int nj = m.njoints;
  for (int i = 1 + gl; i < nj; i += G)
  {
    T * r = jr + i * JR_STRIDE;
    const SE3<T> X = load_se3(r + xoff);
    const Motion<T> ov = load_motion(r + JR_OV);
    Motion<T> oa = load_motion(r + oa_off);
    if (SUB_GRAVITY)
    {
      oa.lin -= Vec3<T>(m.gravity[0], m.gravity[1], m.gravity[2]); // oa_gf[0] = -gravity (:504)
      store6(r + JR_OA, oa);
    }
    const Inertia<T> Y = act(X, model_inertia(m, i));
    const Force<T> oh = Y * ov;
    Force<T> of = Y * oa;
    of += fcross(ov, oh);
    store_lin_inertia(r, Y);
    store_dy(r + JR_DY, inertia_variation(Y, ov, oh));
    store6(r + JR_OF, of);
  }
  BRBD_SYNCWARP();
}

// ---- phase 3: subtree sums, lanes = components ------------------------------------------------------------
// Joints are numbered depth-first, so walking i = nj-1 .. 1 and adding record i into record parent(i) completes
// every subtree sum; a lane owns NC components for the whole walk (no cross-lane dependency, no barrier), and the
// running sum of a chain (parent(i) == i - 1) stays in registers.  (Tried: per-branching-joint accumulators with the own
// values loaded four joints ahead — more instructions, no shorter: 4.2 % -> 5.6 % of computeABADerivatives.)
template<class T, int G, int FIRST = 0, int COUNT = JR_NSUM>
BRBD_DI void coop_subtree_sums(const ModelPOD<T> & m, T * jr, int gl)
{
  constexpr int NC = (COUNT + G - 1) / G;
  const int nj = m.njoints;
  T carry[NC];
  bool on[NC];
#pragma unroll
  for (int t = 0; t < NC; ++t) { on[t] = gl + t * G < COUNT; carry[t] = T(0); }
  T * col = jr + FIRST + gl;
  int carry_idx = -1;
  for (int i = nj - 1; i > 0; --i)
  {
    const int p = m.parent[i];
    if (p > 0)
    {
      const bool have = carry_idx == i;
#pragma unroll
      for (int t = 0; t < NC; ++t)
        if (on[t])
        {
          const T x = have ? carry[t] : col[i * JR_STRIDE + t * G];
          carry[t] = col[p * JR_STRIDE + t * G] + x;
          col[p * JR_STRIDE + t * G] = carry[t];
        }
      carry_idx = p;
    }
  }
  BRBD_SYNCWARP();
}

// ---- phase 4: column records ------------------------------------------------------------------------------
// tau_io: in = a (RNEA derivatives) and out = tau (rnea-derivatives.hxx:409 + armature :538-539); may be null.
template<class T, int G>
BRBD_DI void coop_columns(const ModelPOD<T> & m, const T * jr, T * cb, T * tau_io, int gl)
{
  const int nv = m.nv;
  for (int c = gl; c < nv; c += G)
  {
    const int j = m.dof_joint[c], p = m.parent[j];
    const T * r = jr + j * JR_STRIDE;
    T * P = cb + c * CB_STRIDE;
    const Motion<T> J = ld6m(P + CB_J);
    const LinInertia<T> Y = load_lin_inertia(r);
    const DY<T> dY = load_dy(r + JR_DY);
    const Force<T> of = load_force(r + JR_OF);
    const Motion<T> ovj = load_motion(r + JR_OV);
    Motion<T> dVdq = mzero<T>(), dAdq;
    if (p > 0)
    {
      const T * pr = jr + p * JR_STRIDE;
      const Motion<T> ovp = load_motion(pr + JR_OV), oap = load_motion(pr + JR_OA);
      dVdq = mcross(ovp, J);
      dAdq = mcross(oap, J);
      dAdq += mcross(ovp, dVdq);
    }
    else
    {
      Motion<T> g0 = mzero<T>();
      g0.lin = Vec3<T>(-m.gravity[0], -m.gravity[1], -m.gravity[2]);
      dAdq = mcross(g0, J);
    }
    Motion<T> dAdv = mcross(ovj, J);
    dAdv += dVdq;
    const Force<T> dFda = Y * J;
    Force<T> dFdq = Y * dAdq;
    if (p > 0) dFdq += dY.mul(dVdq);
    Force<T> dFdv = dY.mul(J);
    dFdv += Y * dAdv;
    if (tau_io) tau_io[c] = dot6(J, of) + m.armature[c] * tau_io[c];
    st6(P + CB_DFDQ, dFdq);
    dFdq += fcross(J, of); // motionSet::act<ADDTO>(J_cols, of[i], dFdq_cols) (:440)
    st6(P + CB_DFDQP, dFdq);
    st6(P + CB_DFDV, dFdv);
    st6(P + CB_DFDA, dFda);
    st6(P + CB_DADQ, dAdq);
    st6(P + CB_DVDQ, dVdq);
    st6(P + CB_DADV, dAdv);
    st6(P + CB_DYTJ, dY.tmul(J));
  }
  BRBD_SYNCWARP();
}

// ---- phase 5: matrix entries ------------------------------------------------------------------------------
// Row block of R <= G rows x C = G / R column slices; a lane keeps its row record (J_r, dFda_r, dYtJ_r) in
// registers and walks the columns of its slice.  Entry (r, c):
//   r on the root path of joint(c):  J_r . dFdq_c (own joint) | J_r . dFdq_c+ (strict ancestor)        (:437-440)
//                                    J_r . dFdv_c (:450-451),  J_r . dFda_c (:420-421, + armature on the diagonal)
//   r in the strict subtree:         dFda_r . dAdq_c + dYtJ_r . dVdq_c (:433-435),  dFda_r . dAdv_c + dYtJ_r . J_c (:446-448)
//   otherwise 0.
// any_up / any_low: whether ANY row of the caller's row block is an ancestor row / a subtree row of column c (the same for all
// lanes of the warp: every group of the warp evaluates the same model) — a part no row needs is skipped altogether.
template<class T, bool WITH_DA>
BRBD_DI void coop_entry(const ModelPOD<T> & m, const CoopTables & tb, const T * cb, const T * Jr, const T * Fd, const T * Yd, int r, int c,
                        T & vq, T & vv, T & va, bool any_up = true, bool any_low = true)
{
  const unsigned info = tb.colinfo[c];
  const int ivc = info & 0xff, own_end = (info >> 8) & 0xff, sub_end = info >> 16;
  const bool own = r >= ivc && r < own_end;
  const bool up = (tb.anc_mask[c] >> r) & 1ull;
  const bool low = r >= own_end && r < sub_end;
  const T * P = cb + c * CB_STRIDE;
  T x[6], y[6];
  T Aq = T(0), Av = T(0), Bq = T(0), Bv = T(0), Aa = T(0);
  if (any_up)
  {
    ld6(P + (own ? CB_DFDQ : CB_DFDQP), x);
    Aq = dot6a(Jr, x);
    ld6(P + CB_DFDV, x);
    Av = dot6a(Jr, x);
    if (WITH_DA)
    {
      ld6(P + CB_DFDA, x);
      Aa = dot6a(Jr, x);
    }
  }
  if (any_low)
  {
    ld6(P + CB_DADQ, x); ld6(P + CB_DVDQ, y);
    Bq = dot6a(Fd, x) + dot6a(Yd, y);
    ld6(P + CB_DADV, x); ld6(P + CB_J, y);
    Bv = dot6a(Fd, x) + dot6a(Yd, y);
  }
  vq = up ? Aq : (low ? Bq : T(0));
  vv = up ? Av : (low ? Bv : T(0));
  if (WITH_DA)
  {
    va = up ? Aa : T(0);
    if (r == c) va += m.armature[c];
  }
}
// row-block flags of column c for rows [rb, rb + R)
BRBD_DI void coop_block_flags(const CoopTables & tb, int c, int rb, int R, bool & any_up, bool & any_low)
{
  const unsigned info = tb.colinfo[c];
  const int own_end = (info >> 8) & 0xff, sub_end = info >> 16;
  const unsigned long long blk = (R >= 64 ? ~0ull : ((1ull << R) - 1ull)) << rb;
  any_up = (tb.anc_mask[c] & blk) != 0ull;
  const int lo = own_end > rb ? own_end : rb, hi = sub_end < rb + R ? sub_end : rb + R;
  any_low = lo < hi;
}

template<class T, int G, bool WITH_DA>
BRBD_DI void coop_entries(const ModelPOD<T> & m, const CoopTables & tb, const T * cb, T * __restrict__ gq, T * __restrict__ gv,
                          T * __restrict__ ga, int gl, bool active)
{
  const int nv = m.nv;
  for (int rb = 0; rb < nv; rb += G)
  {
    const int R = (nv - rb) < G ? (nv - rb) : G;
    T Jr[6], Fd[6], Yd[6];
    if (R == G)
    {
      // full row block: one column per step, the column record is a broadcast
      const int r = rb + gl;
      const T * Pr = cb + r * CB_STRIDE;
      ld6(Pr + CB_J, Jr); ld6(Pr + CB_DFDA, Fd); ld6(Pr + CB_DYTJ, Yd);
      T * pq = gq + r, * pv = gv + r, * pa = ga + r;
      for (int c = 0; c < nv; ++c)
      {
        T vq, vv, va;
        coop_entry<T, WITH_DA>(m, tb, cb, Jr, Fd, Yd, r, c, vq, vv, va);
        if (active)
        {
          *pq = vq; *pv = vv;
          if (WITH_DA) *pa = va;
        }
        pq += nv; pv += nv; pa += nv;
      }
    }
    else
    {
      // tail rows: R rows x C = G / R column slices
      const int C = G / R;
      const int rl = gl % R, slice = gl / R;
      const int r = rb + rl;
      const bool lane_on = slice < C;
      const T * Pr = cb + r * CB_STRIDE;
      ld6(Pr + CB_J, Jr); ld6(Pr + CB_DFDA, Fd); ld6(Pr + CB_DYTJ, Yd);
      for (int c0 = 0; c0 < nv; c0 += C)
      {
        const int c = c0 + slice;
        const bool valid = lane_on && c < nv;
        const int cc = valid ? c : 0;
        T vq, vv, va;
        coop_entry<T, WITH_DA>(m, tb, cb, Jr, Fd, Yd, r, cc, vq, vv, va);
        if (valid && active)
        {
          gq[cc * nv + r] = vq; gv[cc * nv + r] = vv;
          if (WITH_DA) ga[cc * nv + r] = va;
        }
      }
    }
  }
}

template<class T, int G>
__global__ void __launch_bounds__(256, 1)
rnea_derivatives_coop_kernel(const ModelPOD<T> * __restrict__ gm, const __grid_constant__ CoopTables gtb, const CoopLayout L,
                             const T * __restrict__ q, int64_t ldq, const T * __restrict__ v, int64_t ldv,
                             const T * __restrict__ a, int64_t lda, T * __restrict__ dq, int64_t ld_dq, T * __restrict__ dv,
                             int64_t ld_dv, T * __restrict__ da, int64_t ld_da, T * __restrict__ tau, int64_t ldtau, int64_t B)
{
  __shared__ ModelPOD<T> m;
  __shared__ CoopTables tb;
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  copy_model_to_smem(&m, gm);
  copy_words_to_smem(reinterpret_cast<int *>(&tb), reinterpret_cast<const int *>(&gtb), (int)(sizeof(CoopTables) / 4));
  __syncthreads();
  constexpr int GPW = 32 / G; // configurations per warp
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int gl = lane % G, grp = lane / G;
  T * base = reinterpret_cast<T *>(dyn_smem) + (size_t)(warp * GPW + grp) * L.per_group;
  T * sq = base + L.oq, * sv = base + L.ov, * sa = base + L.oa, * jr = base + L.ojr, * cb = base + L.ocb;
  const int nq = m.nq, nv = m.nv;
  const int64_t ntiles = (B + GPW - 1) / GPW;
  // the inputs of the next configuration are fetched into registers while the current one is evaluated (nq <= 2 G)
  T rq[2], rv[2], ra[2];
  auto fetch = [&](int64_t tile) {
    int64_t c = tile * GPW + grp;
    if (c >= B) c = B - 1;
    const T * gq_in = q + c * ldq, * gv_in = v + c * ldv, * ga_in = a + c * lda;
#pragma unroll
    for (int t = 0; t < 2; ++t)
    {
      const int k = gl + t * G;
      rq[t] = k < nq ? gq_in[k] : T(0);
      rv[t] = k < nv ? gv_in[k] : T(0);
      ra[t] = k < nv ? ga_in[k] : T(0);
    }
  };
  const int64_t tile0 = (int64_t)blockIdx.x * nw + warp, tstep = (int64_t)gridDim.x * nw;
  if (tile0 < ntiles) fetch(tile0);
  for (int64_t tile = tile0; tile < ntiles; tile += tstep)
  {
    int64_t cfg = tile * GPW + grp;
    const bool active = cfg < B;
    if (!active) cfg = B - 1; // idle groups shadow the last configuration (stores suppressed)
#pragma unroll
    for (int t = 0; t < 2; ++t)
    {
      const int k = gl + t * G;
      if (k < nq) sq[k] = rq[t];
      if (k < nv) { sv[k] = rv[t]; sa[k] = ra[t]; }
    }
    if (tile + tstep < ntiles) fetch(tile + tstep);
    BRBD_SYNCWARP();
    int oa_off = JR_OA;
    const int xoff = coop_forward<T, G, true>(m, tb, sq, sv, sa, jr, cb, gl, &oa_off);
    coop_joint_quantities<T, G>(m, jr, gl, xoff, oa_off);
    coop_subtree_sums<T, G>(m, jr, gl);
    coop_columns<T, G>(m, jr, cb, sa, gl);
    coop_entries<T, G, true>(m, tb, cb, dq + cfg * ld_dq, dv + cfg * ld_dv, da + cfg * ld_da, gl, active);
    if (tau && active)
      for (int k = gl; k < nv; k += G) tau[cfg * ldtau + k] = sa[k];
    BRBD_SYNCWARP();
  }
}

// ---- small batches: RNEA, G lanes per configuration ------------------------------------------------------------------
// One configuration per thread needs >= 33 000 configurations to fill a B200 and its latency is that of one thread walking
// the whole tree (~90 us for a humanoid).  Below that, the batch is given to the cooperative phases above: forward kinematics
// by pointer jumping, f_i = Y_i a_i + v_i x* (Y_i v_i) per joint (rnea.hxx:71-73 in the world frame), subtree sums of f
// (the `f[parent] += liMi.act(f)` of rnea.hxx:105-106), tau = J^T f + armature o a (rnea.hxx:103, :158).
template<class T, int G>
BRBD_DI void rnea_coop_config(const ModelPOD<T> & m, const CoopTables & tb, const CoopLayout & L, T * base, int gl)
{
  T * sq = base + L.oq, * sv = base + L.ov, * sa = base + L.oa, * jr = base + L.ojr, * cb = base + L.ocb;
  int oa_off = JR_OA;
  const int xoff = coop_forward<T, G, true>(m, tb, sq, sv, sa, jr, cb, gl, &oa_off);
  const int nj = m.njoints, nv = m.nv;
  for (int i = 1 + gl; i < nj; i += G)
  {
    T * r = jr + i * JR_STRIDE;
    const SE3<T> X = load_se3(r + xoff);
    const Motion<T> ov = load_motion(r + JR_OV);
    Motion<T> oa = load_motion(r + oa_off);
    oa.lin -= Vec3<T>(m.gravity[0], m.gravity[1], m.gravity[2]); // a_gf[0] = -gravity (rnea.hxx:138)
    const Inertia<T> Y = act(X, model_inertia(m, i));
    Force<T> of = Y * oa;
    of += fcross(ov, Y * ov);
    store6(r + JR_OF, of);
  }
  BRBD_SYNCWARP();
  coop_subtree_sums<T, G, JR_OF, 6>(m, jr, gl);
  for (int c = gl; c < nv; c += G)
  {
    T Jv[6];
    ld6(cb + c * CB_STRIDE + CB_J, Jv);
    const T * f = jr + m.dof_joint[c] * JR_STRIDE + JR_OF;
    sa[c] = Jv[0] * f[0] + Jv[1] * f[1] + Jv[2] * f[2] + Jv[3] * f[3] + Jv[4] * f[4] + Jv[5] * f[5] + m.armature[c] * sa[c];
  }
  BRBD_SYNCWARP();
}

template<class T, int G>
__global__ void __launch_bounds__(256, 1)
rnea_coop_kernel(const ModelPOD<T> * __restrict__ gm, const __grid_constant__ CoopTables gtb, const CoopLayout L,
                 const T * __restrict__ q, int64_t ldq, const T * __restrict__ v, int64_t ldv, const T * __restrict__ a, int64_t lda,
                 T * __restrict__ tau, int64_t ldtau, int64_t B)
{
  __shared__ ModelPOD<T> m;
  __shared__ CoopTables tb;
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  copy_model_to_smem(&m, gm);
  copy_words_to_smem(reinterpret_cast<int *>(&tb), reinterpret_cast<const int *>(&gtb), (int)(sizeof(CoopTables) / 4));
  __syncthreads();
  constexpr int GPW = 32 / G;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int gl = lane % G, grp = lane / G;
  T * base = reinterpret_cast<T *>(dyn_smem) + (size_t)(warp * GPW + grp) * L.per_group;
  const int nq = m.nq, nv = m.nv;
  const int64_t ntiles = (B + GPW - 1) / GPW;
  for (int64_t tile = (int64_t)blockIdx.x * nw + warp; tile < ntiles; tile += (int64_t)gridDim.x * nw)
  {
    int64_t cfg = tile * GPW + grp;
    const bool active = cfg < B;
    if (!active) cfg = B - 1;
    const T * gq_in = q + cfg * ldq, * gv_in = v + cfg * ldv, * ga_in = a + cfg * lda;
    for (int k = gl; k < nq; k += G) base[L.oq + k] = gq_in[k];
    for (int k = gl; k < nv; k += G) { base[L.ov + k] = gv_in[k]; base[L.oa + k] = ga_in[k]; }
    BRBD_SYNCWARP();
    rnea_coop_config<T, G>(m, tb, L, base, gl);
    if (active)
      for (int k = gl; k < nv; k += G) tau[cfg * ldtau + k] = base[L.oa + k];
    BRBD_SYNCWARP();
  }
}

} // namespace brbd
