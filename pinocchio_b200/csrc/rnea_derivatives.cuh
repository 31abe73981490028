// rnea_derivatives.cuh — batched computeRNEADerivatives, one configuration per thread.
//
// Restates impl::computeRNEADerivatives (reference: include/pinocchio/algorithm/rnea-derivatives.hxx
// :472-541) with ComputeRNEADerivativesForwardStep (:263-351) and ...BackwardStep (:378-459).
//
// The reference writes, at joint i, the row block (i, subtree(i)) and the column block
// (subtree+(i), i) of dtau_dq / dtau_dv and the row block of dtau_da.  Here every output matrix is
// produced column by column: when the backward sweep reaches joint i all quantities of column i are
// known —
//   upper part  (a, i), a ancestor-or-self :  J_a^T dFd*_i            (rnea-derivatives.hxx:437-438,450-451,420-421)
//   lower part  (d, i), d strict descendant:  dFda_d^T dAd*_i + dYtJ_d^T dV*_i   (:433-435, :446-448)
// — and a column of the caller's col-major matrix is contiguous, so it leaves through the
// warp-cooperative ColumnEmitter with coalesced stores.  The dot products are the reference's.
// Entries outside the tree sparsity are written as zeros (the reference requires pre-zeroed
// outputs, rnea-derivatives.hpp:104-106).
#pragma once

#include "crba.cuh"
#include "engine.cuh"
#include "rnea.cuh"

namespace brbd
{

// general 3x3, row-major
template<class T> struct M33
{
  T a[9];
  BRBD_DI T & operator()(int r, int c) { return a[3 * r + c]; }
  BRBD_DI const T & operator()(int r, int c) const { return a[3 * r + c]; }
};
template<class T> BRBD_DI Vec3<T> mul(const M33<T> & A, const Vec3<T> & v)
{
  return Vec3<T>(A.a[0] * v.x + A.a[1] * v.y + A.a[2] * v.z, A.a[3] * v.x + A.a[4] * v.y + A.a[5] * v.z,
                 A.a[6] * v.x + A.a[7] * v.y + A.a[8] * v.z);
}
template<class T> BRBD_DI Vec3<T> tmul(const M33<T> & A, const Vec3<T> & v)
{
  return Vec3<T>(A.a[0] * v.x + A.a[3] * v.y + A.a[6] * v.z, A.a[1] * v.x + A.a[4] * v.y + A.a[7] * v.z,
                 A.a[2] * v.x + A.a[5] * v.y + A.a[8] * v.z);
}
// skewSquare(u, v) = v u^T - (u.v) 1 — skew.hpp:182-197
template<class T> BRBD_DI M33<T> skew_square(const Vec3<T> & u, const Vec3<T> & v)
{
  M33<T> C;
  C.a[0] = v.x * u.x; C.a[1] = v.x * u.y; C.a[2] = v.x * u.z;
  C.a[3] = v.y * u.x; C.a[4] = v.y * u.y; C.a[5] = v.y * u.z;
  C.a[6] = v.z * u.x; C.a[7] = v.z * u.y; C.a[8] = v.z * u.z;
  const T d = dot(u, v);
  C.a[0] -= d; C.a[4] -= d; C.a[8] -= d;
  return C;
}
// M += skew(v) — skew.hpp:68-84
template<class T> BRBD_DI void add_skew(const Vec3<T> & v, M33<T> & M)
{
  M(0, 1) -= v.z; M(0, 2) += v.y;
  M(1, 0) += v.z; M(1, 2) -= v.x;
  M(2, 0) -= v.y; M(2, 1) += v.x;
}

// d(oYcrb)/dt-like 6x6 of the reference: blocks LL = 0, LA, AL, AA (27 numbers).
template<class T> struct DY
{
  M33<T> LA, AL, AA;
  // doY * m
  BRBD_DI Force<T> mul(const Motion<T> & m) const
  {
    Force<T> f;
    f.lin = brbd::mul(LA, m.ang);
    f.ang = brbd::mul(AL, m.lin) + brbd::mul(AA, m.ang);
    return f;
  }
  // doY^T * m  (J_cols^T * doYcrb, rnea-derivatives.hxx:432)
  BRBD_DI Force<T> tmul(const Motion<T> & m) const
  {
    Force<T> f;
    f.lin = brbd::tmul(AL, m.ang);
    f.ang = brbd::tmul(LA, m.lin) + brbd::tmul(AA, m.ang);
    return f;
  }
};
template<class T> BRBD_DI void store_dy(T * d, const DY<T> & D)
{
#pragma unroll
  for (int k = 0; k < 9; ++k) { d[k] = D.LA.a[k]; d[9 + k] = D.AL.a[k]; d[18 + k] = D.AA.a[k]; }
}
template<class T> BRBD_DI DY<T> load_dy(const T * d)
{
  DY<T> D;
#pragma unroll
  for (int k = 0; k < 9; ++k) { D.LA.a[k] = d[k]; D.AL.a[k] = d[9 + k]; D.AA.a[k] = d[18 + k]; }
  return D;
}

// Inertia::variation(v) (inertia.hpp:749-776) followed by addForceCrossMatrix(h, .) (rnea-derivatives.hxx:340-351)
template<class T> BRBD_DI DY<T> inertia_variation(const Inertia<T> & Y, const Motion<T> & v, const Force<T> & h)
{
  DY<T> D;
  const Vec3<T> ml = Y.m * v.lin, mw = Y.m * v.ang, c = Y.c;
  // LA = -skew(mv.lin) - skewSquare(mv.ang, c) + skewSquare(c, mv.ang)
  const M33<T> A = skew_square(mw, c), Bm = skew_square(c, mw);
  M33<T> LA;
#pragma unroll
  for (int k = 0; k < 9; ++k) LA.a[k] = -A.a[k];
  // -skew(ml): skew = [[0,-z,y],[z,0,-x],[-y,x,0]]
  LA(0, 1) = ml.z - A(0, 1); LA(0, 2) = -ml.y - A(0, 2);
  LA(1, 0) = -ml.z - A(1, 0); LA(1, 2) = ml.x - A(1, 2);
  LA(2, 0) = ml.y - A(2, 0); LA(2, 1) = -ml.x - A(2, 1);
#pragma unroll
  for (int k = 0; k < 9; ++k) LA.a[k] += Bm.a[k];
  D.LA = LA;
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int cc = 0; cc < 3; ++cc) D.AL(r, cc) = LA(cc, r);
  // AA = -skewSquare(mv.lin, c) - skewSquare(c, mv.lin)
  const M33<T> C1 = skew_square(ml, c), C2 = skew_square(c, ml);
  M33<T> AA;
#pragma unroll
  for (int k = 0; k < 9; ++k) AA.a[k] = -C1.a[k] - C2.a[k];
  // S = I_c - m [c]x^2 as a matrix (AlphaSkewSquare)
  const T cx = c.x, cy = c.y, cz = c.z, m = Y.m;
  const T sxx = Y.I.xx + m * (cy * cy + cz * cz), sxy = Y.I.xy - m * cx * cy, sxz = Y.I.xz - m * cx * cz;
  const T syy = Y.I.yy + m * (cx * cx + cz * cz), syz = Y.I.yz - m * cy * cz, szz = Y.I.zz + m * (cx * cx + cy * cy);
  M33<T> S;
  S.a[0] = sxx; S.a[1] = sxy; S.a[2] = sxz; S.a[3] = sxy; S.a[4] = syy; S.a[5] = syz; S.a[6] = sxz; S.a[7] = syz; S.a[8] = szz;
  // AA -= S * skew(w);  skew(w) columns: col0 = (0, wz, -wy), col1 = (-wz, 0, wx), col2 = (wy, -wx, 0)
  const Vec3<T> w = v.ang;
#pragma unroll
  for (int r = 0; r < 3; ++r)
  {
    AA(r, 0) -= S(r, 1) * w.z - S(r, 2) * w.y;
    AA(r, 1) -= S(r, 2) * w.x - S(r, 0) * w.z;
    AA(r, 2) -= S(r, 0) * w.y - S(r, 1) * w.x;
  }
  // AA += [w]x S  (cross(v.angular(), S), skew.hpp:228-245)
#pragma unroll
  for (int j = 0; j < 3; ++j)
  {
    AA(0, j) += w.y * S(2, j) - w.z * S(1, j);
    AA(1, j) += w.z * S(0, j) - w.x * S(2, j);
    AA(2, j) += w.x * S(1, j) - w.y * S(0, j);
  }
  D.AA = AA;
  // addForceCrossMatrix(h): addSkew(-h.lin) on LA and AL, addSkew(-h.ang) on AA
  add_skew(-h.lin, D.LA);
  add_skew(-h.lin, D.AL);
  add_skew(-h.ang, D.AA);
  return D;
}

template<class T> BRBD_DI T dotff(const Force<T> & a, const Motion<T> & b) { return dot(a.lin, b.lin) + dot(a.ang, b.ang); }

// State shared by the derivative sweeps (thread-local memory).
template<class T> struct DerivState
{
  T J[MAXNV][6], dVdq[MAXNV][6], dAdq[MAXNV][6], dAdv[MAXNV][6];
  T dFda[MAXNV][6], dYtJ[MAXNV][6];
  T Y[MAXJ][10];   // oYcrb
  T dY[MAXJ][27];  // doYcrb
  T of[MAXJ][6];
};

// dJ, dVdq, dAdq, dAdv columns of joint i (rnea-derivatives.hxx:319-332 == aba-derivatives.hxx:241-251)
template<class T>
BRBD_DI void deriv_columns(const ModelPOD<T> & m, DerivState<T> & st, int i, const Motion<T> & ov, const Motion<T> & ov_parent,
                           const Motion<T> & oa_gf_parent)
{
  const int parent = m.parent[i], iv = m.idx_v[i], nvj = m.nvj[i];
  for (int k = 0; k < nvj; ++k)
  {
    const Motion<T> Jk = load_motion(st.J[iv + k]);
    const Motion<T> dJ = mcross(ov, Jk);
    Motion<T> dAdq = mcross(oa_gf_parent, Jk);
    Motion<T> dAdv = dJ;
    Motion<T> dVdq = mzero<T>();
    if (parent > 0)
    {
      dVdq = mcross(ov_parent, Jk);
      dAdq += mcross(ov_parent, dVdq);
      dAdv += dVdq;
    }
    store6(st.dVdq[iv + k], dVdq);
    store6(st.dAdq[iv + k], dAdq);
    store6(st.dAdv[iv + k], dAdv);
  }
}

// Backward sweep shared by computeRNEADerivatives (WITH_DA = true: also dtau_da and tau) and the last
// pass of computeABADerivatives (aba-derivatives.hxx:283-367, WITH_DA = false).
// gq / gv / ga point at configuration 0 of the warp tile; tau_row is this thread's staged `a` row,
// overwritten by tau (rnea-derivatives.hxx:409-410, 538-539).
template<class T, bool WITH_DA>
BRBD_DI void deriv_backward(const ModelPOD<T> & m, DerivState<T> & st, ColumnEmitter<T> & eq, ColumnEmitter<T> & ev,
                            ColumnEmitter<T> & ea, T * gq, int64_t ldq_, T * gv, int64_t ldv_, T * ga, int64_t lda_,
                            T * tau_row, int nc)
{
  const int nj = m.njoints, nv = m.nv;
  for (int i = nj - 1; i > 0; --i)
  {
    const int parent = m.parent[i], iv = m.idx_v[i], nvj = m.nvj[i], nsub = m.nvsub[i];
    const Inertia<T> Y = load_inertia(st.Y[i]);
    const DY<T> dY = load_dy(st.dY[i]);
    const Force<T> of = load_force(st.of[i]);
    for (int k = 0; k < nvj; ++k)
    {
      const int col = iv + k;
      const Motion<T> Jc = load_motion(st.J[col]);
      const Motion<T> dVdq = load_motion(st.dVdq[col]), dAdq = load_motion(st.dAdq[col]), dAdv = load_motion(st.dAdv[col]);
      if (WITH_DA) tau_row[col] = dot6(Jc, of) + m.armature[col] * tau_row[col];
      const Force<T> dFda = Y * Jc;
      Force<T> dFdq = Y * dAdq;
      if (parent > 0) dFdq += dY.mul(dVdq);
      const Force<T> dYtJ = dY.tmul(Jc);
      Force<T> dFdv = dY.mul(Jc);
      dFdv += Y * dAdv;
      store6(st.dFda[col], dFda);
      store6(st.dYtJ[col], dYtJ);
      Force<T> dFdq_post = dFdq;
      dFdq_post += fcross(Jc, of); // motionSet::act<ADDTO>(J_cols, of[i], dFdq_cols) (:440)
      for (int a = i; a > 0; a = m.parent[a])
      {
        const int ia = m.idx_v[a], na = m.nvj[a];
        for (int r = 0; r < na; ++r)
        {
          const Motion<T> Jr = load_motion(st.J[ia + r]);
          eq.put(ia + r, dot6(Jr, a == i ? dFdq : dFdq_post));
          ev.put(ia + r, dot6(Jr, dFdv));
          if (WITH_DA)
          {
            T val = dot6(Jr, dFda);
            if (ia + r == col) val += m.armature[col];
            ea.put(ia + r, val);
          }
        }
      }
      for (int d = iv + nvj; d < iv + nsub; ++d)
      {
        const Force<T> Fd = load_force(st.dFda[d]), Yd = load_force(st.dYtJ[d]);
        eq.put(d, dotff(Fd, dAdq) + dotff(Yd, dVdq));
        ev.put(d, dotff(Fd, dAdv) + dotff(Yd, Jc));
      }
      eq.flush(gq + (int64_t)col * nv, ldq_, nc);
      ev.flush(gv + (int64_t)col * nv, ldv_, nc);
      if (WITH_DA) ea.flush(ga + (int64_t)col * nv, lda_, nc);
    }
    if (parent > 0)
    {
      Inertia<T> Yp = load_inertia(st.Y[parent]);
      Yp += Y;
      store_inertia(st.Y[parent], Yp);
      for (int k = 0; k < 27; ++k) st.dY[parent][k] += st.dY[i][k];
      Force<T> fp = load_force(st.of[parent]);
      fp += of;
      store6(st.of[parent], fp);
    }
  }
}

template<class T>
BRBD_DI void rnea_derivatives_thread(const ModelPOD<T> & m, const T * q, const T * v, T * a_tau, ColumnEmitter<T> & eq,
                                     ColumnEmitter<T> & ev, ColumnEmitter<T> & ea, T * gq, int64_t ld_q, T * gv,
                                     int64_t ld_v, T * ga, int64_t ld_a, int nc)
{
  DerivState<T> st;
  T oMi_d[MAXDEPTH][12], vl_d[MAXDEPTH][6], al_d[MAXDEPTH][6], ov_d[MAXDEPTH][6], oag_d[MAXDEPTH][6];
  const Vec3<T> g(m.gravity[0], m.gravity[1], m.gravity[2]);
  {
    Motion<T> g0 = mzero<T>();
    g0.lin = -g; // data.oa_gf[0] = -model.gravity (:504)
    store6(oag_d[0], g0);
  }
  const int nj = m.njoints;
  for (int i = 1; i < nj; ++i)
  {
    const int type = m.type[i], parent = m.parent[i], iq = m.idx_q[i], iv = m.idx_v[i], d = m.depth[i], nvj = m.nvj[i];
    const SE3<T> X = joint_liMi(m, i, type, q + iq);
    Motion<T> vi = joint_velocity(type, v + iv);
    SE3<T> oMi = X;
    if (parent > 0)
    {
      oMi = load_se3(oMi_d[d - 1]) * X;
      vi += X.actInv(load_motion(vl_d[d - 1]));
    }
    // a_i = S*a + c + (v_i x v_J) (+ liMi^-1 a_parent)   (:295-300)
    Motion<T> ai = cross_joint_velocity(vi, type, v + iv);
    for (int k = 0; k < nvj; ++k)
    {
      const int row = joint_S_row(type, k);
      const T ak = a_tau[iv + k];
      if (row < 3) ai.lin.set(row, ai.lin.get(row) + ak); else ai.ang.set(row - 3, ai.ang.get(row - 3) + ak);
    }
    if (parent > 0) ai += X.actInv(load_motion(al_d[d - 1]));
    store_se3(oMi_d[d], oMi);
    store6(vl_d[d], vi);
    store6(al_d[d], ai);
    const Inertia<T> Y = act(oMi, model_inertia(m, i));
    const Motion<T> ov = oMi.act(vi);
    Motion<T> oa_gf = oMi.act(ai);
    oa_gf.lin -= g;
    const Force<T> oh = Y * ov;
    Force<T> of = Y * oa_gf;
    of += fcross(ov, oh);
    store_inertia(st.Y[i], Y);
    store6(st.of[i], of);
    for (int k = 0; k < nvj; ++k) store6(st.J[iv + k], act_S_col(oMi, type, k));
    Motion<T> ovp = mzero<T>();
    if (parent > 0) ovp = load_motion(ov_d[d - 1]);
    deriv_columns(m, st, i, ov, ovp, load_motion(oag_d[d - 1]));
    store6(ov_d[d], ov);
    store6(oag_d[d], oa_gf);
    store_dy(st.dY[i], inertia_variation(Y, ov, oh));
  }
  deriv_backward<T, true>(m, st, eq, ev, ea, gq, ld_q, gv, ld_v, ga, ld_a, a_tau, nc);
}

template<class T>
__global__ void __launch_bounds__(512)
rnea_derivatives_kernel(const ModelPOD<T> * __restrict__ gm, const T * __restrict__ q, int64_t ldq,
                        const T * __restrict__ v, int64_t ldv, const T * __restrict__ a, int64_t lda,
                        T * __restrict__ dq, int64_t ld_dq, T * __restrict__ dv, int64_t ld_dv, T * __restrict__ da,
                        int64_t ld_da, T * __restrict__ tau, int64_t ldtau, int64_t B)
{
  __shared__ ModelPOD<T> m;
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  copy_model_to_smem(&m, gm);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int qpad = m.nq | 1, vpad = m.nv | 1;
  T * sq = reinterpret_cast<T *>(dyn_smem) + (size_t)warp * 32 * (qpad + 5 * vpad);
  T * sv = sq + 32 * qpad;
  T * sa = sv + 32 * vpad;
  ColumnEmitter<T> eq, ev, ea;
  eq.init(sa + 32 * vpad, vpad, m.nv, lane);
  ev.init(sa + 64 * vpad, vpad, m.nv, lane);
  ea.init(sa + 96 * vpad, vpad, m.nv, lane);
  const int64_t ntiles = (B + 31) / 32;
  for (int64_t tile = (int64_t)blockIdx.x * nw + warp; tile < ntiles; tile += (int64_t)gridDim.x * nw)
  {
    const int64_t c0 = tile * 32;
    const int nc = (int)((B - c0) < 32 ? (B - c0) : 32);
    tile_load(sq, qpad, q + c0 * ldq, ldq, m.nq, nc, lane);
    tile_load(sv, vpad, v + c0 * ldv, ldv, m.nv, nc, lane);
    tile_load(sa, vpad, a + c0 * lda, lda, m.nv, nc, lane);
    BRBD_SYNCWARP();
    pad_tile_rows(sq, qpad, m.nq, nc, lane);
    pad_tile_rows(sv, vpad, m.nv, nc, lane);
    pad_tile_rows(sa, vpad, m.nv, nc, lane);
    BRBD_SYNCWARP();
    rnea_derivatives_thread(m, sq + lane * qpad, sv + lane * vpad, sa + lane * vpad, eq, ev, ea, dq + c0 * ld_dq, ld_dq,
                            dv + c0 * ld_dv, ld_dv, da + c0 * ld_da, ld_da, nc);
    BRBD_SYNCWARP();
    if (tau) tile_store(tau + c0 * ldtau, ldtau, sa, vpad, m.nv, nc, lane);
    BRBD_SYNCWARP();
  }
}

} // namespace brbd
