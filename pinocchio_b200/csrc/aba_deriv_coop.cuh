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
constexpr int SC_SIZE = 128; // scratch of the multi-dof joint step

struct AbaCoopLayout
{
  int oq, ov, ou, ojr, ocb, ominv; // offsets (elements) inside one group's region
  int osc;                         // scratch of the multi-dof joint step (aliases Minv, which is written later, when large enough)
  int mld;                         // leading dimension of Minv (odd)
  int per_group;                   // elements, even
};
template<int G> struct CoopGemmShape
{
  static constexpr int RG = (G == 32) ? 8 : 4; // lane grid: RG row groups x CG column groups
  static constexpr int CG = G / RG;
  static constexpr int NB = 2 * CG;            // columns of dtau_dq (and of dtau_dv) per block
  static constexpr int DLD = 4 * CG + 2;       // leading dimension of the D block (2 NB columns): rows 16-byte aligned
};
inline AbaCoopLayout aba_coop_layout(int nq, int nv, int nj, int G)
{
  const int dld = (G == 32 || G == 16) ? 18 : 10;
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
  if (L.mld * nv + 8 >= SC_SIZE) L.osc = L.ominv;
  else
  {
    L.osc = L.per_group;
    L.per_group += SC_SIZE;
  }
  return L;
}

// ---- A2: articulated-body inertia, leaf -> root -----------------------------------------------------------
// Multi-dof joint (free-flyer, spherical, planar): all lanes of the group work on the one joint — the 6 x 6 / nvj x nvj
// pieces live in a 128-value scratch block, Dinv by Cholesky (PerformStYSInversion, joint-common-operations.hpp:23-33):
// factor by lane 0, the columns of the inverse by nvj lanes.
constexpr int SC_IA = 0, SC_F = 36, SC_S = 42, SC_D = 78, SC_INVD = 114, SC_AB = 120;
BRBD_DI int sym6_index(int a, int b)
{
  const int lo = a < b ? a : b, hi = a < b ? b : a;
  return lo * 6 - (lo * (lo - 1)) / 2 + (hi - lo);
}
template<class T, int G>
BRBD_DI void coop_aba_multidof(const ModelPOD<T> & m, const CoopTables & tb, T * jr, T * cb, T * su, T * sc, int gl, int i, int xoff)
{
  const int parent = m.parent[i], iv = m.idx_v[i], nvj = m.nvj[i];
  T * r = jr + i * JR_STRIDE;
  T * P0 = cb + iv * CB_STRIDE;
  if (gl == 0)
  {
    const SE3<T> X = load_se3(r + xoff);
    const Motion<T> ov = load_motion(r + JR_OV);
    const Inertia<T> Y = act(X, model_inertia(m, i));
    T Ia[21], f[6];
    inertia_to_sym6(Y, Ia);
    f2a(fcross(ov, Y * ov), f);
#pragma unroll
    for (int a = 0; a < 6; ++a)
#pragma unroll
      for (int b = 0; b < 6; ++b) sc[SC_IA + a * 6 + b] = Ia[a <= b ? a * 6 - (a * (a - 1)) / 2 + (b - a) : b * 6 - (b * (b - 1)) / 2 + (a - b)];
    Motion<T> ab = mzero<T>();
    if (parent > 0) ab = mcross(load_motion(jr + parent * JR_STRIDE + JR_OV), ov);
    store6(r + JR_OA, ab);
    store6(sc + SC_AB, ab);
#pragma unroll
    for (int k = 0; k < 6; ++k) sc[SC_F + k] = f[k];
  }
  BRBD_SYNCWARP();
  // contributions of the children
  const int last = tb.jlast[i];
  for (int e = gl; e < 42; e += G)
  {
    const int off = e < 36 ? A_IACC + sym6_index(e / 6, e % 6) : A_PA + (e - 36);
    T * dst = sc + (e < 36 ? SC_IA + e : SC_F + (e - 36));
    T acc = *dst;
    for (int c = i + 1; c <= last; c = tb.jlast[c] + 1) acc += cb[m.idx_v[c] * CB_STRIDE + off];
    *dst = acc;
  }
  BRBD_SYNCWARP();
  // U = Ia J
  for (int e = gl; e < 6 * nvj; e += G)
  {
    const int a = e % 6, k = e / 6;
    const T * Jk = P0 + k * CB_STRIDE + CB_J;
    const T * Ir = sc + SC_IA + a * 6;
    T acc = Ir[0] * Jk[0];
#pragma unroll
    for (int b = 1; b < 6; ++b) acc += Ir[b] * Jk[b];
    P0[k * CB_STRIDE + A_U + a] = acc;
  }
  BRBD_SYNCWARP();
  // StU = J^T U + armature;  u = tau - J^T f
  for (int e = gl; e < nvj * nvj + nvj; e += G)
  {
    if (e < nvj * nvj)
    {
      const int a = e / nvj, b = e % nvj;
      const T * Ja = P0 + a * CB_STRIDE + CB_J, * Ub = P0 + b * CB_STRIDE + A_U;
      T acc = dot6a(Ja, Ub);
      if (a == b) acc += m.armature[iv + a];
      sc[SC_S + a * 6 + b] = acc;
    }
    else
    {
      const int k = e - nvj * nvj;
      su[iv + k] -= dot6a(P0 + k * CB_STRIDE + CB_J, sc + SC_F);
    }
  }
  BRBD_SYNCWARP();
  if (gl == 0)
  {
    T * S = sc + SC_S;
    for (int a = 0; a < nvj; ++a)
      for (int b = 0; b <= a; ++b)
      {
        T acc = S[a * 6 + b];
        for (int k = 0; k < b; ++k) acc -= S[a * 6 + k] * S[b * 6 + k];
        if (a == b)
        {
          const T d = sqrt_t(acc);
          S[a * 6 + a] = d;
          sc[SC_INVD + a] = T(1) / d;
        }
        else
          S[a * 6 + b] = acc * sc[SC_INVD + b];
      }
  }
  BRBD_SYNCWARP();
  if (gl < nvj)
  {
    const T * S = sc + SC_S;
    T * D = sc + SC_D + gl; // column gl of the inverse, D[a * 6]
    for (int a = 0; a < nvj; ++a)
    {
      T acc = (a == gl) ? T(1) : T(0);
      for (int k = 0; k < a; ++k) acc -= S[a * 6 + k] * D[k * 6];
      D[a * 6] = acc * sc[SC_INVD + a];
    }
    for (int a = nvj - 1; a >= 0; --a)
    {
      T acc = D[a * 6];
      for (int k = a + 1; k < nvj; ++k) acc -= S[k * 6 + a] * D[k * 6];
      D[a * 6] = acc * sc[SC_INVD + a];
    }
  }
  BRBD_SYNCWARP();
  // U Dinv, rows of Dinv (zero-padded to 6)
  for (int e = gl; e < 6 * nvj; e += G)
  {
    const int a = e % 6, k = e / 6;
    T acc = T(0);
    for (int c = 0; c < nvj; ++c) acc += P0[c * CB_STRIDE + A_U + a] * sc[SC_D + c * 6 + k];
    P0[k * CB_STRIDE + A_UD + a] = acc;
    P0[k * CB_STRIDE + A_DINV + a] = a < nvj ? sc[SC_D + k * 6 + a] : T(0);
  }
  BRBD_SYNCWARP();
  if (parent > 0)
  {
    // Ia -= U Dinv U^T, left for the parent together with pa = f + Ia a_bias + U Dinv u
    for (int e = gl; e < 36; e += G)
    {
      const int a = e / 6, b = e % 6;
      T acc = T(0);
      for (int k = 0; k < nvj; ++k) acc += P0[k * CB_STRIDE + A_UD + a] * P0[k * CB_STRIDE + A_U + b];
      const T val = sc[SC_IA + e] - acc;
      sc[SC_IA + e] = val;
      if (a <= b) P0[A_IACC + sym6_index(a, b)] = val;
    }
    BRBD_SYNCWARP();
    for (int a = gl; a < 6; a += G)
    {
      T acc = sc[SC_F + a];
#pragma unroll
      for (int b = 0; b < 6; ++b) acc += sc[SC_IA + a * 6 + b] * sc[SC_AB + b];
      for (int k = 0; k < nvj; ++k) acc += P0[k * CB_STRIDE + A_UD + a] * su[iv + k];
      P0[A_PA + a] = acc;
    }
  }
  BRBD_SYNCWARP();
}

template<class T, int G>
BRBD_DI void coop_aba_backward(const ModelPOD<T> & m, const CoopTables & tb, T * jr, T * cb, T * su, T * sc, int gl, int xoff)
{
  for (int l = m.maxdepth; l >= 1; --l)
  {
    // 1-dof joints of this depth: one lane each
    for (int s = tb.lvl_start[l] + gl; s < tb.lvl_start[l + 1]; s += G)
    {
      const int i = tb.lvl_joint[s];
      const int parent = m.parent[i], iv = m.idx_v[i];
      if (m.nvj[i] != 1) continue;
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
    // multi-dof joints of this depth: all lanes together, one joint after the other
    if (tb.lvl_multi & (1u << l))
      for (int s = tb.lvl_start[l]; s < tb.lvl_start[l + 1]; ++s)
      {
        const int i = tb.lvl_joint[s];
        if (m.nvj[i] > 1) coop_aba_multidof<T, G>(m, tb, jr, cb, su, sc, gl, i, xoff);
      }
    BRBD_SYNCWARP();
  }
}

// ---- A3: upper rows of Minv, lanes = columns (NCB columns per lane, interleaved) -------------------------
// For column c and every joint i on the root path of joint(c), leaf to root (aba-derivatives.hxx:131-166):
//   own block:      Minv(i, c) = Dinv
//   ancestors:      Minv(i, c) = -(J_i Dinv_i)^T F,   then   F += U_i Minv(i, c),   F = Fcrb[0](:, c)
template<class T, int G, int NCB>
BRBD_DI void coop_minv_upper(const ModelPOD<T> & m, const CoopTables & tb, const T * cb, T * Minv, int mld, int gl)
{
  const int nj = m.njoints;
  T F[NCB][6];
#pragma unroll
  for (int b = 0; b < NCB; ++b)
#pragma unroll
    for (int k = 0; k < 6; ++k) F[b][k] = T(0);
  for (int i = nj - 1; i > 0; --i)
  {
    const unsigned info = tb.jinfo[i]; // idx_v | nv_joint << 8 | (idx_v + nvSubtree) << 16 | (parent > 0) << 24
    const int iv = info & 0xff, nvj = (info >> 8) & 0xff, end = (info >> 16) & 0xff;
    const bool hp = (info >> 24) != 0;
    const T * P0 = cb + iv * CB_STRIDE;
    if (nvj == 1)
    {
      T Jv[6], U[6];
      const T Dinv = P0[A_DINV];
      ld6(P0 + CB_J, Jv);
      if (hp) ld6(P0 + A_U, U);
#pragma unroll
      for (int b = 0; b < NCB; ++b)
      {
        const int c = gl + b * G;
        if (c >= iv && c < end)
        {
          const T mk = (c == iv) ? Dinv : -(Dinv * dot6a(Jv, F[b]));
          Minv[iv * mld + c] = mk;
          if (hp)
          {
#pragma unroll
            for (int k = 0; k < 6; ++k) F[b][k] += U[k] * mk;
          }
        }
      }
    }
    else
    {
#pragma unroll
      for (int b = 0; b < NCB; ++b)
      {
        const int c = gl + b * G;
        if (c >= iv && c < end)
        {
          const bool own = c < iv + nvj;
          T jf[6], mk[6];
#pragma unroll
          for (int a = 0; a < 6; ++a)
          {
            jf[a] = T(0);
            if (a < nvj && !own)
            {
              T Jv[6];
              ld6(P0 + a * CB_STRIDE + CB_J, Jv);
              jf[a] = dot6a(Jv, F[b]);
            }
          }
#pragma unroll
          for (int k = 0; k < 6; ++k)
          {
            mk[k] = T(0);
            if (k < nvj)
            {
              const T * Dk = P0 + k * CB_STRIDE + A_DINV; // row k of Dinv (symmetric), zero-padded to 6
              T Dr[6];
              ld6(Dk, Dr);
              const T val = own ? Dk[c - iv] : -dot6a(Dr, jf);
              mk[k] = val;
              Minv[(iv + k) * mld + c] = val;
            }
          }
          if (hp)
          {
#pragma unroll
            for (int k = 0; k < 6; ++k)
              if (k < nvj)
              {
                T U[6];
                ld6(P0 + k * CB_STRIDE + A_U, U);
#pragma unroll
                for (int a = 0; a < 6; ++a) F[b][a] += U[a] * mk[k];
              }
          }
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
        T dd[6], uu[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) uu[k] = k < nvj ? su[iv + k] : T(0);
#pragma unroll
        for (int k = 0; k < 6; ++k)
        {
          dd[k] = T(0);
          if (k < nvj)
          {
            T UD[6], Dr[6];
            ld6(P0 + k * CB_STRIDE + A_UD, UD);
            ld6(P0 + k * CB_STRIDE + A_DINV, Dr);
            dd[k] = dot6a(Dr, uu) - dot6a(UD, ag);
          }
        }
#pragma unroll
        for (int k = 0; k < 6; ++k)
          if (k < nvj)
          {
            T Jv[6];
            ld6(P0 + k * CB_STRIDE + CB_J, Jv);
            su[iv + k] = dd[k];
#pragma unroll
            for (int a = 0; a < 6; ++a) ag[a] += Jv[a] * dd[k];
          }
      }
#pragma unroll
      for (int k = 0; k < 6; ++k) r[JR_OA + k] = ag[k];
    }
    BRBD_SYNCWARP();
  }
}

// ---- A5: completion of Minv, lanes = columns, one tangent row per step (aba-derivatives.hxx:220-234) -------
//   Minv(r, c) -= UDinv_r^T Fcrb[parent](:, c)   for every column c >= idx_v of the joint owning row r
//   Fcrb[i](:, c) = Fcrb[parent](:, c) + sum over the joint's rows of J_r Minv(r, c)
// Fcrb of the current chain stays in registers; it is saved at joints with several children and reloaded when the
// depth-first order returns to them.  Entries right of the joint's own block are mirrored into the lower triangle
// as they become final (:448-449); the own blocks of multi-dof joints are mirrored at the end.
template<class T, int G, int NCB>
BRBD_DI void coop_minv_complete(const ModelPOD<T> & m, const CoopTables & tb, T * cb, T * Minv, int mld, int gl)
{
  const int nv = m.nv, nj = m.njoints;
  T Fd[NCB][6], acc[NCB][6];
  T * save[NCB];
#pragma unroll
  for (int b = 0; b < NCB; ++b)
  {
    const int c = gl + b * G;
    save[b] = cb + (c < nv ? c : nv - 1) * CB_STRIDE + A_FD;
#pragma unroll
    for (int k = 0; k < 6; ++k) { Fd[b][k] = T(0); acc[b][k] = T(0); }
  }
  for (int r = 0; r < nv; ++r)
  {
    const unsigned info = tb.rinfo[r]; // idx_v | own_end << 8 | sub_end << 16 | flags << 24 (1 first row, 2 last row, 4 parent > 0)
    const int slots = tb.rslots[r];    // load slot | save slot << 8 (0xff = none)
    const int iv = info & 0xff, own_end = (info >> 8) & 0xff, sub_end = (info >> 16) & 0xff;
    const bool first = (info >> 24) & 1, last = (info >> 24) & 2, hp = (info >> 24) & 4;
    const int lslot = slots & 0xff, sslot = (slots >> 8) & 0xff;
    const T * Pr = cb + r * CB_STRIDE;
    T Jv[6], UD[6];
    ld6(Pr + CB_J, Jv);
    T * Mrow = Minv + r * mld;
    if (first && last)
    {
      // 1-dof joint: Fcrb[i] = Fcrb[parent] + J Minv(r, c)
      if (hp)
      {
        ld6(Pr + A_UD, UD);
#pragma unroll
        for (int b = 0; b < NCB; ++b)
        {
          const int c = gl + b * G;
          if (c >= iv && c < nv)
          {
            if (lslot != 0xff) ld6(save[b] + 6 * lslot, Fd[b]);
            T mv = c < sub_end ? Mrow[c] : T(0); // right of the subtree the upper triangle starts at zero (:414)
            mv -= dot6a(UD, Fd[b]);
            Mrow[c] = mv;
            if (c >= own_end) Minv[c * mld + r] = mv;
#pragma unroll
            for (int k = 0; k < 6; ++k) Fd[b][k] += Jv[k] * mv;
            if (sslot != 0xff) st6(save[b] + 6 * sslot, Fd[b]);
          }
        }
      }
      else
      {
#pragma unroll
        for (int b = 0; b < NCB; ++b)
        {
          const int c = gl + b * G;
          if (c >= iv && c < nv)
          {
            const T mv = c < sub_end ? Mrow[c] : T(0);
            Mrow[c] = mv;
            if (c >= own_end) Minv[c * mld + r] = mv;
#pragma unroll
            for (int k = 0; k < 6; ++k) Fd[b][k] = Jv[k] * mv;
            if (sslot != 0xff) st6(save[b] + 6 * sslot, Fd[b]);
          }
        }
      }
      continue;
    }
    if (hp) ld6(Pr + A_UD, UD);
#pragma unroll
    for (int b = 0; b < NCB; ++b)
    {
      const int c = gl + b * G;
      if (c >= iv && c < nv)
      {
        if (first)
        {
          if (lslot != 0xff) ld6(save[b] + 6 * lslot, Fd[b]);
#pragma unroll
          for (int k = 0; k < 6; ++k) acc[b][k] = hp ? Fd[b][k] : T(0);
        }
        T mv = c < sub_end ? Mrow[c] : T(0);
        if (hp) mv -= dot6a(UD, Fd[b]);
        Mrow[c] = mv;
        if (c >= own_end) Minv[c * mld + r] = mv;
#pragma unroll
        for (int k = 0; k < 6; ++k) acc[b][k] += Jv[k] * mv;
        if (last)
        {
#pragma unroll
          for (int k = 0; k < 6; ++k) Fd[b][k] = acc[b][k];
          if (sslot != 0xff) st6(save[b] + 6 * sslot, Fd[b]);
        }
      }
    }
  }
  BRBD_SYNCWARP();
  if (tb.lvl_multi)
  {
    for (int i = 1; i < nj; ++i)
    {
      const unsigned info = tb.jinfo[i];
      const int iv = info & 0xff, nvj = (info >> 8) & 0xff;
      if (nvj > 1)
        for (int e = gl; e < nvj * nvj; e += G)
        {
          const int a = e / nvj, b = e % nvj;
          if (a > b) Minv[(iv + a) * mld + iv + b] = Minv[(iv + b) * mld + iv + a];
        }
    }
    BRBD_SYNCWARP();
  }
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

// out(:, block) = -Minv * Dblk; lane (rg, cg) accumulates rows rg + RG i (i < R), block columns 4 cg + j (j < 4)
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
  const T * pb = Dblk + 4 * cg;
#pragma unroll 5
  for (int k = 0; k < nv; ++k)
  {
    T a[R], b[4];
#pragma unroll
    for (int i = 0; i < R; ++i) a[i] = pa[S::RG * i];
    ld4(pb, b);
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
      const int dc = 4 * cg + j;
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
    case 6: coop_dblock_product<T, G, 6>(nv, Minv, mld, Dblk, c0, gq, gv, gl, active); break;
    default: coop_dblock_product<T, G, 8>(nv, Minv, mld, Dblk, c0, gq, gv, gl, active); break; // nv <= 64 = MAXNV
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
// MODE 0: computeABADerivatives.  MODE 1: computeMinverse (reference: algorithm/aba.hxx:613-902) — phases A1, A2, A3, A5 with
// v = tau = 0; like the reference's data.Minv only the upper triangle is meaningful, the strictly-lower part is written as zeros.
// MODE 2: ABA alone (phases A1, A2, A4) — the small-batch path of abaInParallel, see rnea_coop_kernel in deriv_coop.cuh.
template<class T, int G, int MODE = 0>
BRBD_DI void aba_derivatives_coop_config(const ModelPOD<T> & m, const CoopTables & tb, const AbaCoopLayout & L, T * base, int gl,
                                         T * __restrict__ gq, T * __restrict__ gv, T * __restrict__ gm, T * __restrict__ gddq, bool active)
{
  typedef CoopGemmShape<G> S;
  T * sq = base + L.oq, * sv = base + L.ov, * su = base + L.ou, * jr = base + L.ojr, * cb = base + L.ocb, * Minv = base + L.ominv;
  const int nv = m.nv, mld = L.mld;
  int oa_unused = JR_OA;
  const int xoff = coop_forward<T, G, false>(m, tb, sq, sv, (const T *)nullptr, jr, cb, gl, &oa_unused);
  coop_aba_backward<T, G>(m, tb, jr, cb, su, base + L.osc, gl, xoff);
  if (MODE == 2)
  {
    coop_aba_forward2<T, G>(m, tb, jr, cb, su, gl);
    if (active)
      for (int k = gl; k < nv; k += G) gddq[k] = su[k];
    return;
  }
  if (G == 32 && nv > G)
  {
    coop_minv_upper<T, G, (G == 32 ? 2 : 1)>(m, tb, cb, Minv, mld, gl);
    if (MODE == 0) coop_aba_forward2<T, G>(m, tb, jr, cb, su, gl);
    coop_minv_complete<T, G, (G == 32 ? 2 : 1)>(m, tb, cb, Minv, mld, gl);
  }
  else
  {
    coop_minv_upper<T, G, 1>(m, tb, cb, Minv, mld, gl);
    if (MODE == 0) coop_aba_forward2<T, G>(m, tb, jr, cb, su, gl);
    coop_minv_complete<T, G, 1>(m, tb, cb, Minv, mld, gl);
  }
  if (MODE == 1)
  {
    if (active)
    {
      int c = 0, r = gl;
      while (r >= nv) { r -= nv; ++c; }
      for (int e = gl; e < nv * nv; e += G)
      {
        gm[e] = r <= c ? Minv[c * mld + r] : T(0);
        r += G;
        while (r >= nv) { r -= nv; ++c; }
      }
    }
    return;
  }
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
    int c = 0, r = gl;
    while (r >= nv) { r -= nv; ++c; }
    for (int e = gl; e < nv * nv; e += G)
    {
      gm[e] = Minv[c * mld + r];
      r += G;
      while (r >= nv) { r -= nv; ++c; }
    }
    if (gddq)
      for (int k = gl; k < nv; k += G) gddq[k] = su[k];
  }
}

template<class T, int G, int MODE = 0>
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
  copy_words_to_smem(reinterpret_cast<int *>(&tb), reinterpret_cast<const int *>(&gtb), (int)(sizeof(CoopTables) / 4));
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
    if (MODE != 1)
      for (int k = gl; k < nv; k += G) { base[L.ov + k] = gv_in[k]; base[L.ou + k] = gt_in[k]; }
    else
      for (int k = gl; k < nv; k += G) { base[L.ov + k] = T(0); base[L.ou + k] = T(0); }
    BRBD_SYNCWARP();
    aba_derivatives_coop_config<T, G, MODE>(m, tb, L, base, gl, MODE == 0 ? dq + cfg * ld_dq : (T *)nullptr,
                                            MODE == 0 ? dv + cfg * ld_dv : (T *)nullptr, MODE == 2 ? (T *)nullptr : dtau + cfg * ld_dtau,
                                            (MODE != 1 && ddq) ? ddq + cfg * ldddq : (T *)nullptr, active);
    BRBD_SYNCWARP();
  }
}

} // namespace brbd
