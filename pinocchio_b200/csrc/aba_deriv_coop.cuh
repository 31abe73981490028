// aba_deriv_coop.cuh — batched computeABADerivatives, warp-cooperative: G lanes work on ONE configuration and the
// whole evaluation (ABA sweeps, Minv, the RNEA-derivative columns and the two dense products) stays in shared memory.
//
// Restates impl::computeABADerivatives (reference: include/pinocchio/algorithm/aba-derivatives.hxx:380-453).  The
// reference interleaves four sweeps over the joints; the data flow of each quantity is kept, the sweeps are
// re-cut by what can run across lanes:
//
//   A1  forward kinematics by pointer jumping (deriv_coop.cuh): oMi, J, ov                ForwardStep1 (:38-79)
//   A2  articulated inertia, leaf -> root, LEVEL-parallel (lanes = joints of one depth; a parent gathers the
//       contributions its children left behind): U, Dinv, UDinv, u, pa                   BackwardStep1 (:97-170), ABA part
//   A3  upper rows of Minv, lanes = COLUMNS: column c only needs U / Dinv / J of the joints on its root path and
//       its own 6-vector Fcrb[0](:, c), which lives in registers                          BackwardStep1, Minv part (:131-166)
//   A4  ddq and oa_gf, root -> leaf, level-parallel                                       ForwardStep2 (:188-219)
//   A5  completion of Minv, lanes = columns: Minv(i, c) -= UDinv_i^T Fcrb[parent](:, c); Fcrb[i](:, c) of the
//       current chain in registers, saved only at joints with several children            ForwardStep2 (:220-234)
//       then the lower triangle is mirrored (:448-449)
//   B   the RNEA-derivative phases 2-4 of deriv_coop.cuh with a := ddq                    ForwardStep2 (:236-256), BackwardStep2 (:283-367)
//   C   blocks of NB columns of dtau_dq / dtau_dv are produced into shared memory (phase 5 of deriv_coop.cuh) and
//       immediately multiplied: ddq_dq = -Minv dtau_dq, ddq_dv = -Minv dtau_dv (:451-452) with a register-tiled
//       product (lane tile R x 4), so the two intermediates never exist in global memory
//   D   ddq_dtau = Minv, ddq
//
// Shared memory per configuration (simple_humanoid: 39 KB -> 5 warps per SM): column records (54 nv), joint
// records (55 nj, reused as the product's D block), Minv (nv x (nv|1)), q / v / tau.
#pragma once

#include "aba.cuh"
#include "deriv_coop.cuh"

namespace brbd
{

// ABA quantities kept in the column records until phase B overwrites them (CB_J = 42..47 stays):
//   per column:  U (6) | U Dinv (6) | row of Dinv (6)
//   per joint, in the record of its first column: articulated inertia left for the parent (21), pa (6)
//   A5: Fcrb save slots of column c, 6 values per branching joint, reuse [18, 42)
constexpr int A_U = 0, A_UD = 6, A_DINV = 12, A_IACC = 18, A_PA = 48, A_FD = 18, A_MAXBRANCH = 4;

struct AbaCoopLayout
{
  int oq, ov, ou, ojr, ocb, ominv; // offsets (elements) inside one group's region
  int mld;                         // leading dimension of Minv (odd)
  int per_group;                   // elements, even
};
template<int G> struct CoopGemmShape
{
  static constexpr int RG = (G == 32) ? 8 : 4; // lane grid: RG row groups x CG column groups
  static constexpr int CG = G / RG;
  static constexpr int NB = 2 * CG;            // columns of dtau_dq (and of dtau_dv) per block
  static constexpr int DLD = 4 * CG + 1;       // leading dimension of the D block (2 NB columns), odd
};
inline AbaCoopLayout aba_coop_layout(int nq, int nv, int nj, int G)
{
  const int dld = (G == 32 || G == 16) ? 17 : 9;
  AbaCoopLayout L;
  L.ocb = 0;
  L.ojr = L.ocb + CB_STRIDE * nv;
  int jr = JR_STRIDE * nj;
  if (jr < dld * nv) jr = dld * nv;
  jr = (jr + 1) & ~1;
  L.ominv = L.ojr + jr;
  L.mld = nv | 1;
  L.oq = (L.ominv + L.mld * nv + 8 + 1) & ~1; // + 8: the product reads up to RG - 1 rows past the matrix
  L.ov = L.oq + nq;
  L.ou = L.ov + nv;
  L.per_group = (L.ou + nv + 1) & ~1;
  return L;
}

// ---- A2: articulated-body inertia, leaf -> root -----------------------------------------------------------
template<class T, int G>
BRBD_DI void coop_aba_backward(const ModelPOD<T> & m, const CoopTables & tb, T * jr, T * cb, T * su, int gl, int xoff)
{
  for (int l = m.maxdepth; l >= 1; --l)
  {
    for (int s = tb.lvl_start[l] + gl; s < tb.lvl_start[l + 1]; s += G)
    {
      const int i = tb.lvl_joint[s];
      const int parent = m.parent[i], iv = m.idx_v[i], nvj = m.nvj[i];
      T * r = jr + i * JR_STRIDE;
      T * P0 = cb + iv * CB_STRIDE;
      const SE3<T> X = load_se3(r + xoff);
      const Motion<T> ov = load_motion(r + JR_OV);
      const Inertia<T> Y = act(X, model_inertia(m, i));
      T Ia[21], f[6];
      inertia_to_sym6(Y, Ia);
      f2a(fcross(ov, Y * ov), f);
      // contributions of the children (aba-derivatives.hxx:168-169: oYaba[parent] += Ia, of[parent] += pa)
      const int last = tb.jlast[i];
      for (int c = i + 1; c <= last; c = tb.jlast[c] + 1)
      {
        const T * pc = cb + m.idx_v[c] * CB_STRIDE;
#pragma unroll
        for (int k = 0; k < 21; ++k) Ia[k] += pc[A_IACC + k];
        T pa[6];
        ld6(pc + A_PA, pa);
#pragma unroll
        for (int k = 0; k < 6; ++k) f[k] += pa[k];
      }
      Motion<T> ab = mzero<T>();
      if (parent > 0) ab = mcross(load_motion(jr + parent * JR_STRIDE + JR_OV), ov); // a_gf bias (:66)
      store6(r + JR_OA, ab);
      if (nvj == 1)
      {
        T Jv[6], U[6], UD[6];
        ld6(P0 + CB_J, Jv);
        sym6_mul(Ia, Jv, U);
        const T D = dot6a(Jv, U) + m.armature[iv];
        const T Dinv = T(1) / D;
#pragma unroll
        for (int k = 0; k < 6; ++k) UD[k] = U[k] * Dinv;
        const T u = su[iv] - dot6a(Jv, f);
        su[iv] = u;
        st6(P0 + A_U, U);
        st6(P0 + A_UD, UD);
        P0[A_DINV] = Dinv;
        if (parent > 0)
        {
#pragma unroll
          for (int a = 0; a < 6; ++a)
#pragma unroll
            for (int b = a; b < 6; ++b) Ia[a * 6 - (a * (a - 1)) / 2 + (b - a)] -= UD[a] * U[b];
          T abv[6], Iab[6];
          m2a(ab, abv);
          sym6_mul(Ia, abv, Iab);
#pragma unroll
          for (int k = 0; k < 6; ++k) f[k] += Iab[k] + UD[k] * u;
#pragma unroll
          for (int k = 0; k < 21; ++k) P0[A_IACC + k] = Ia[k];
          st6(P0 + A_PA, f);
        }
      }
      else
      {
        // multi-dof joint (free-flyer, spherical, planar): Dinv by Cholesky (joint-common-operations.hpp:23-33)
        T U[6][6], StU[6][6], Di[6][6], UD[6][6], uj[6];
        for (int k = 0; k < nvj; ++k)
        {
          T Jv[6], Uk[6];
          ld6(P0 + k * CB_STRIDE + CB_J, Jv);
          sym6_mul(Ia, Jv, Uk);
          uj[k] = su[iv + k] - dot6a(Jv, f);
          su[iv + k] = uj[k];
          for (int a = 0; a < 6; ++a) U[a][k] = Uk[a];
          st6(P0 + k * CB_STRIDE + A_U, Uk);
        }
        for (int a = 0; a < nvj; ++a)
        {
          T Jv[6];
          ld6(P0 + a * CB_STRIDE + CB_J, Jv);
          for (int b = 0; b < nvj; ++b)
          {
            T acc = Jv[0] * U[0][b];
            for (int k = 1; k < 6; ++k) acc += Jv[k] * U[k][b];
            StU[a][b] = acc;
          }
          StU[a][a] += m.armature[iv + a];
        }
        llt_inverse(nvj, StU, Di);
        for (int k = 0; k < nvj; ++k)
        {
          T UDk[6], Dk[6];
          for (int a = 0; a < 6; ++a)
          {
            T acc = U[a][0] * Di[0][k];
            for (int c = 1; c < nvj; ++c) acc += U[a][c] * Di[c][k];
            UD[a][k] = acc;
            UDk[a] = acc;
            Dk[a] = a < nvj ? Di[k][a] : T(0);
          }
          st6(P0 + k * CB_STRIDE + A_UD, UDk);
          st6(P0 + k * CB_STRIDE + A_DINV, Dk);
        }
        if (parent > 0)
        {
          for (int a = 0; a < 6; ++a)
            for (int b = a; b < 6; ++b)
            {
              T acc = UD[a][0] * U[b][0];
              for (int k = 1; k < nvj; ++k) acc += UD[a][k] * U[b][k];
              Ia[a * 6 - (a * (a - 1)) / 2 + (b - a)] -= acc;
            }
          T abv[6], Iab[6];
          m2a(ab, abv);
          sym6_mul(Ia, abv, Iab);
          for (int a = 0; a < 6; ++a)
          {
            T acc = UD[a][0] * uj[0];
            for (int k = 1; k < nvj; ++k) acc += UD[a][k] * uj[k];
            f[a] += Iab[a] + acc;
          }
          for (int k = 0; k < 21; ++k) P0[A_IACC + k] = Ia[k];
          st6(P0 + A_PA, f);
        }
      }
    }
    BRBD_SYNCWARP();
  }
}

// ---- A3: upper rows of Minv, lanes = columns -------------------------------------------------------------
// For column c and every joint i on the root path of joint(c), leaf to root (aba-derivatives.hxx:131-166):
//   own block:      Minv(i, c) = Dinv
//   ancestors:      Minv(i, c) = -(J_i Dinv_i)^T F,   then   F += U_i Minv(i, c),   F = Fcrb[0](:, c)
template<class T, int G>
BRBD_DI void coop_minv_upper(const ModelPOD<T> & m, const T * cb, T * Minv, int mld, int gl)
{
  const int nv = m.nv, nj = m.njoints;
  for (int cblk = 0; cblk < nv; cblk += G)
  {
    const int c = cblk + gl;
    T F[6] = {T(0), T(0), T(0), T(0), T(0), T(0)};
    for (int i = nj - 1; i > 0; --i)
    {
      const int iv = m.idx_v[i], nvj = m.nvj[i], nsub = m.nvsub[i], parent = m.parent[i];
      if (iv + nsub <= cblk || iv >= cblk + G) continue; // no column of this block below joint i
      if (c < iv || c >= iv + nsub) continue;
      const T * P0 = cb + iv * CB_STRIDE;
      const bool own = c < iv + nvj;
      if (nvj == 1)
      {
        T Jv[6], U[6];
        const T Dinv = P0[A_DINV];
        T mk = Dinv;
        if (!own)
        {
          ld6(P0 + CB_J, Jv);
          mk = -(Dinv * dot6a(Jv, F));
        }
        Minv[iv * mld + c] = mk;
        if (parent > 0)
        {
          ld6(P0 + A_U, U);
#pragma unroll
          for (int k = 0; k < 6; ++k) F[k] += U[k] * mk;
        }
      }
      else
      {
        T jf[6], mk[6];
        if (!own)
          for (int a = 0; a < nvj; ++a)
          {
            T Jv[6];
            ld6(P0 + a * CB_STRIDE + CB_J, Jv);
            jf[a] = dot6a(Jv, F);
          }
        for (int k = 0; k < nvj; ++k)
        {
          const T * Dk = P0 + k * CB_STRIDE + A_DINV; // row k of Dinv (symmetric)
          T val;
          if (own) val = Dk[c - iv];
          else
          {
            val = Dk[0] * jf[0];
            for (int a = 1; a < nvj; ++a) val += Dk[a] * jf[a];
            val = -val;
          }
          mk[k] = val;
          Minv[(iv + k) * mld + c] = val;
        }
        if (parent > 0)
          for (int k = 0; k < nvj; ++k)
          {
            T U[6];
            ld6(P0 + k * CB_STRIDE + A_U, U);
            for (int a = 0; a < 6; ++a) F[a] += U[a] * mk[k];
          }
      }
    }
  }
  BRBD_SYNCWARP();
}

// ---- A4: ddq and oa_gf, root -> leaf (aba-derivatives.hxx:206-219) ---------------------------------------
template<class T, int G>
BRBD_DI void coop_aba_forward2(const ModelPOD<T> & m, const CoopTables & tb, T * jr, const T * cb, T * su, int gl)
{
  for (int l = 1; l <= m.maxdepth; ++l)
  {
    for (int s = tb.lvl_start[l] + gl; s < tb.lvl_start[l + 1]; s += G)
    {
      const int i = tb.lvl_joint[s];
      const int parent = m.parent[i], iv = m.idx_v[i], nvj = m.nvj[i];
      T * r = jr + i * JR_STRIDE;
      const T * P0 = cb + iv * CB_STRIDE;
      T ag[6], ab[6];
      if (parent > 0)
      {
        const T * pr = jr + parent * JR_STRIDE + JR_OA;
#pragma unroll
        for (int k = 0; k < 6; ++k) ag[k] = pr[k];
      }
      else
      {
        ag[0] = -m.gravity[0]; ag[1] = -m.gravity[1]; ag[2] = -m.gravity[2]; // data.oa_gf[0] = -gravity (:410)
        ag[3] = ag[4] = ag[5] = T(0);
      }
#pragma unroll
      for (int k = 0; k < 6; ++k) { ab[k] = r[JR_OA + k]; ag[k] += ab[k]; }
      if (nvj == 1)
      {
        T UD[6], Jv[6];
        ld6(P0 + A_UD, UD);
        ld6(P0 + CB_J, Jv);
        const T dd = P0[A_DINV] * su[iv] - dot6a(UD, ag);
        su[iv] = dd;
#pragma unroll
        for (int k = 0; k < 6; ++k) ag[k] += Jv[k] * dd;
      }
      else
      {
        T dd[6];
        for (int k = 0; k < nvj; ++k)
        {
          T UD[6];
          ld6(P0 + k * CB_STRIDE + A_UD, UD);
          const T * Dk = P0 + k * CB_STRIDE + A_DINV;
          T t1 = Dk[0] * su[iv];
          for (int c = 1; c < nvj; ++c) t1 += Dk[c] * su[iv + c];
          dd[k] = t1 - dot6a(UD, ag);
        }
        for (int k = 0; k < nvj; ++k)
        {
          T Jv[6];
          ld6(P0 + k * CB_STRIDE + CB_J, Jv);
          su[iv + k] = dd[k];
          for (int a = 0; a < 6; ++a) ag[a] += Jv[a] * dd[k];
        }
      }
#pragma unroll
      for (int k = 0; k < 6; ++k) r[JR_OA + k] = ag[k];
    }
    BRBD_SYNCWARP();
  }
}

// ---- A5: completion of Minv, lanes = columns (aba-derivatives.hxx:220-234), then the mirror (:448-449) -----
template<class T, int G>
BRBD_DI void coop_minv_complete(const ModelPOD<T> & m, const CoopTables & tb, T * cb, T * Minv, int mld, int gl)
{
  const int nv = m.nv, nj = m.njoints;
  for (int cblk = 0; cblk < nv; cblk += G)
  {
    const int c = cblk + gl;
    const int cs = c < nv ? c : nv - 1; // idle lanes shadow the last column's save slots (never read back as active)
    T * save = cb + cs * CB_STRIDE + A_FD;
    T Fd[6] = {T(0), T(0), T(0), T(0), T(0), T(0)};
    for (int i = 1; i < nj; ++i)
    {
      const int iv = m.idx_v[i], nvj = m.nvj[i], parent = m.parent[i];
      if (iv >= cblk + G) break;
      const bool active = c >= iv && c < nv;
      if (active)
      {
        if (parent > 0 && parent != i - 1) ld6(save + 6 * tb.bslot[parent], Fd);
        const T * P0 = cb + iv * CB_STRIDE;
        const bool in_sub = c < iv + m.nvsub[i]; // right of the subtree the upper triangle starts at zero (:414)
        if (nvj == 1)
        {
          T Jv[6];
          T mv = in_sub ? Minv[iv * mld + c] : T(0);
          if (parent > 0)
          {
            T UD[6];
            ld6(P0 + A_UD, UD);
            mv -= dot6a(UD, Fd);
          }
          Minv[iv * mld + c] = mv;
          ld6(P0 + CB_J, Jv);
#pragma unroll
          for (int k = 0; k < 6; ++k) Fd[k] = (parent > 0 ? Fd[k] : T(0)) + Jv[k] * mv;
        }
        else
        {
          T acc[6];
          for (int a = 0; a < 6; ++a) acc[a] = parent > 0 ? Fd[a] : T(0);
          for (int k = 0; k < nvj; ++k)
          {
            T mv = in_sub ? Minv[(iv + k) * mld + c] : T(0);
            if (parent > 0)
            {
              T UD[6];
              ld6(P0 + k * CB_STRIDE + A_UD, UD);
              mv -= dot6a(UD, Fd);
            }
            Minv[(iv + k) * mld + c] = mv;
            T Jv[6];
            ld6(P0 + k * CB_STRIDE + CB_J, Jv);
            for (int a = 0; a < 6; ++a) acc[a] += Jv[a] * mv;
          }
          for (int a = 0; a < 6; ++a) Fd[a] = acc[a];
        }
        if (tb.bslot[i] >= 0) st6(save + 6 * tb.bslot[i], Fd);
      }
    }
  }
  BRBD_SYNCWARP();
  // lower triangle := upper triangle
  for (int r = 0; r < nv; ++r)
    for (int c = r + 1 + gl; c < nv; c += G) Minv[c * mld + r] = Minv[r * mld + c];
  BRBD_SYNCWARP();
}

// ---- C: blocks of dtau_dq / dtau_dv columns -> shared memory -> -Minv * block -> global ---------------------
template<class T, int G>
BRBD_DI void coop_dblock_fill(const ModelPOD<T> & m, const CoopTables & tb, const T * cb, T * Dblk, int c0, int gl)
{
  typedef CoopGemmShape<G> S;
  const int nv = m.nv;
  for (int rb = 0; rb < nv; rb += G)
  {
    const int R = (nv - rb) < G ? (nv - rb) : G;
    T Jr[6], Fd[6], Yd[6];
    if (R == G)
    {
      const int r = rb + gl;
      const T * Pr = cb + r * CB_STRIDE;
      ld6(Pr + CB_J, Jr); ld6(Pr + CB_DFDA, Fd); ld6(Pr + CB_DYTJ, Yd);
      for (int cc = 0; cc < S::NB; ++cc)
      {
        const int c = c0 + cc;
        T vq = T(0), vv = T(0), va;
        if (c < nv) coop_entry<T, false>(m, tb, cb, Jr, Fd, Yd, r, c, vq, vv, va);
        Dblk[r * S::DLD + cc] = vq;
        Dblk[r * S::DLD + S::NB + cc] = vv;
      }
    }
    else
    {
      const int C = G / R;
      const int rl = gl % R, slice = gl / R;
      const int r = rb + rl;
      const bool lane_on = slice < C;
      const T * Pr = cb + r * CB_STRIDE;
      ld6(Pr + CB_J, Jr); ld6(Pr + CB_DFDA, Fd); ld6(Pr + CB_DYTJ, Yd);
      for (int cc0 = 0; cc0 < S::NB; cc0 += C)
      {
        const int cc = cc0 + slice;
        const int c = c0 + cc;
        if (lane_on && cc < S::NB)
        {
          T vq = T(0), vv = T(0), va;
          if (c < nv) coop_entry<T, false>(m, tb, cb, Jr, Fd, Yd, r, c, vq, vv, va);
          Dblk[r * S::DLD + cc] = vq;
          Dblk[r * S::DLD + S::NB + cc] = vv;
        }
      }
    }
  }
  BRBD_SYNCWARP();
}

// out(:, block) = -Minv * Dblk; lane (rg, cg) accumulates rows rg + RG i (i < R), block columns cg + CG j (j < 4)
template<class T, int G, int R>
BRBD_DI void coop_dblock_product(int nv, const T * Minv, int mld, const T * Dblk, int c0, T * __restrict__ gq, T * __restrict__ gv, int gl,
                                 bool active)
{
  typedef CoopGemmShape<G> S;
  const int rg = gl % S::RG, cg = gl / S::RG;
  T acc[R][4];
#pragma unroll
  for (int i = 0; i < R; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = T(0);
  const T * pa = Minv + rg;
  const T * pb = Dblk + cg;
  for (int k = 0; k < nv; ++k)
  {
    T a[R], b[4];
#pragma unroll
    for (int i = 0; i < R; ++i) a[i] = pa[S::RG * i];
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = pb[S::CG * j];
#pragma unroll
    for (int i = 0; i < R; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * b[j];
    pa += mld;
    pb += S::DLD;
  }
  if (active)
  {
#pragma unroll
    for (int j = 0; j < 4; ++j)
    {
      const int dc = cg + S::CG * j;
      const int col = c0 + (dc < S::NB ? dc : dc - S::NB);
      T * dst = (dc < S::NB ? gq : gv) + col * nv + rg;
      if (col < nv)
      {
#pragma unroll
        for (int i = 0; i < R; ++i)
          if (rg + S::RG * i < nv) dst[S::RG * i] = -acc[i][j];
      }
    }
  }
}

template<class T, int G>
BRBD_DI void coop_dblock_product_dispatch(int nv, const T * Minv, int mld, const T * Dblk, int c0, T * gq, T * gv, int gl, bool active)
{
  typedef CoopGemmShape<G> S;
  const int R = (nv + S::RG - 1) / S::RG;
  if (G == 32)
  {
    switch (R)
    {
    case 1: case 2: case 3: coop_dblock_product<T, G, 3>(nv, Minv, mld, Dblk, c0, gq, gv, gl, active); break;
    case 4: coop_dblock_product<T, G, 4>(nv, Minv, mld, Dblk, c0, gq, gv, gl, active); break;
    case 5: coop_dblock_product<T, G, 5>(nv, Minv, mld, Dblk, c0, gq, gv, gl, active); break;
    default: coop_dblock_product<T, G, 6>(nv, Minv, mld, Dblk, c0, gq, gv, gl, active); break;
    }
  }
  else if (G == 16)
  {
    if (R <= 3) coop_dblock_product<T, G, 3>(nv, Minv, mld, Dblk, c0, gq, gv, gl, active);
    else coop_dblock_product<T, G, 4>(nv, Minv, mld, Dblk, c0, gq, gv, gl, active);
  }
  else
  {
    if (R <= 1) coop_dblock_product<T, G, 1>(nv, Minv, mld, Dblk, c0, gq, gv, gl, active);
    else coop_dblock_product<T, G, 2>(nv, Minv, mld, Dblk, c0, gq, gv, gl, active);
  }
}

// ---- one configuration, all phases --------------------------------------------------------------------------
// `base` = this group's region (AbaCoopLayout), q / v / tau already staged at oq / ov / ou.
template<class T, int G>
BRBD_DI void aba_derivatives_coop_config(const ModelPOD<T> & m, const CoopTables & tb, const AbaCoopLayout & L, T * base, int gl,
                                         T * __restrict__ gq, T * __restrict__ gv, T * __restrict__ gm, T * __restrict__ gddq, bool active)
{
  typedef CoopGemmShape<G> S;
  T * sq = base + L.oq, * sv = base + L.ov, * su = base + L.ou, * jr = base + L.ojr, * cb = base + L.ocb, * Minv = base + L.ominv;
  const int nv = m.nv, mld = L.mld;
  int oa_unused = JR_OA;
  const int xoff = coop_forward<T, G, false>(m, tb, sq, sv, (const T *)nullptr, jr, cb, gl, &oa_unused);
  coop_aba_backward<T, G>(m, tb, jr, cb, su, gl, xoff);
  coop_minv_upper<T, G>(m, cb, Minv, mld, gl);
  coop_aba_forward2<T, G>(m, tb, jr, cb, su, gl);
  coop_minv_complete<T, G>(m, tb, cb, Minv, mld, gl);
  coop_joint_quantities<T, G, false>(m, jr, gl, xoff, JR_OA);
  coop_subtree_sums<T, G>(m, jr, gl);
  coop_columns<T, G>(m, jr, cb, (T *)nullptr, gl);
  T * Dblk = jr; // the joint records are dead from here on
  for (int c0 = 0; c0 < nv; c0 += S::NB)
  {
    coop_dblock_fill<T, G>(m, tb, cb, Dblk, c0, gl);
    coop_dblock_product_dispatch<T, G>(nv, Minv, mld, Dblk, c0, gq, gv, gl, active);
    BRBD_SYNCWARP();
  }
  if (active)
  {
    for (int c = 0; c < nv; ++c)
      for (int r = gl; r < nv; r += G) gm[c * nv + r] = Minv[c * mld + r];
    if (gddq)
      for (int k = gl; k < nv; k += G) gddq[k] = su[k];
  }
}

template<class T, int G>
__global__ void __launch_bounds__(256, 1)
aba_derivatives_coop_kernel(const ModelPOD<T> * __restrict__ gmod, const __grid_constant__ CoopTables gtb, const AbaCoopLayout L,
                            const T * __restrict__ q, int64_t ldq, const T * __restrict__ v, int64_t ldv,
                            const T * __restrict__ tau, int64_t ldtau, T * __restrict__ dq, int64_t ld_dq, T * __restrict__ dv,
                            int64_t ld_dv, T * __restrict__ dtau, int64_t ld_dtau, T * __restrict__ ddq, int64_t ldddq, int64_t B)
{
  __shared__ ModelPOD<T> m;
  __shared__ CoopTables tb;
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  copy_model_to_smem(&m, gmod);
  {
    const int n = (int)(sizeof(CoopTables) / 4);
    const int * s = reinterpret_cast<const int *>(&gtb);
    int * d = reinterpret_cast<int *>(&tb);
    for (int k = threadIdx.x; k < n; k += blockDim.x) d[k] = s[k];
  }
  __syncthreads();
  constexpr int GPW = 32 / G; // configurations per warp
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int gl = lane % G, grp = lane / G;
  T * base = reinterpret_cast<T *>(dyn_smem) + (size_t)(warp * GPW + grp) * L.per_group;
  const int nq = m.nq, nv = m.nv;
  const int64_t ntiles = (B + GPW - 1) / GPW;
  for (int64_t tile = (int64_t)blockIdx.x * nw + warp; tile < ntiles; tile += (int64_t)gridDim.x * nw)
  {
    int64_t cfg = tile * GPW + grp;
    const bool active = cfg < B;
    if (!active) cfg = B - 1; // idle groups shadow the last configuration (stores suppressed)
    const T * gq_in = q + cfg * ldq, * gv_in = v + cfg * ldv, * gt_in = tau + cfg * ldtau;
    for (int k = gl; k < nq; k += G) base[L.oq + k] = gq_in[k];
    for (int k = gl; k < nv; k += G) { base[L.ov + k] = gv_in[k]; base[L.ou + k] = gt_in[k]; }
    BRBD_SYNCWARP();
    aba_derivatives_coop_config<T, G>(m, tb, L, base, gl, dq + cfg * ld_dq, dv + cfg * ld_dv, dtau + cfg * ld_dtau,
                                      ddq ? ddq + cfg * ldddq : (T *)nullptr, active);
    BRBD_SYNCWARP();
  }
}

} // namespace brbd
