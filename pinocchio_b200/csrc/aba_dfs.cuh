// aba_dfs.cuh — batched ABA (WORLD convention), v2: one configuration per thread, pass 1 and pass 2
// DFS-interleaved with their state in shared memory, the per-joint quantities pass 3 needs
// (J, a_bias, U Dinv, Dinv, u) in a coalesced, L2-resident per-thread store.
//
// Restates impl::abaWorldConvention (reference: include/pinocchio/algorithm/aba.hxx:242-293), which is
// what abaInParallel evaluates (algorithm/parallel/aba.hpp:82):
//   pass 1  AbaWorldConventionForwardStep1  (aba.hxx:101-138)
//   pass 2  AbaWorldConventionBackwardStep  (aba.hxx:152-192)
//   pass 3  AbaWorldConventionForwardStep2  (aba.hxx:206-226)
// The "consistent output" tail (aba.hxx:228-230, 286-290: data.oa, data.of) is not returned by the batched
// entry point and is skipped.  oYaba is a packed symmetric 6x6 (21 numbers) instead of the reference's
// dense Matrix6.  Dinv = StU^-1 by Cholesky for multi-dof joints (PerformStYSInversion,
// joint-common-operations.hpp:23-33), a reciprocal for 1-dof joints.
//
// State (tree.cuh): per depth the joint's own world inertia (10) and bias force (6); per open branching
// joint oMi (12), ov / oa_gf (6) and the accumulated children contribution to (oYaba, of) (21 + 6).  The
// contribution of an only child travels to its parent in registers.
#pragma once

#include "aba.cuh"
#include "tmem.cuh"
#include "tree.cuh"

namespace brbd
{

struct AbaLayout
{
  int oY, oF, oB, nstate; // per depth: Y (10), f (6); per branch slot: 45
};
constexpr int ABA_BR = 45; // oMi 12 | ov / oa_gf 6 | Ia acc 21 | f acc 6
inline AbaLayout aba_layout(int maxdepth, int nbranch)
{
  AbaLayout L;
  L.oY = 0;
  L.oF = L.oY + 10 * maxdepth;
  L.oB = L.oF + 6 * maxdepth;
  L.nstate = L.oB + ABA_BR * (nbranch > 0 ? nbranch : 1);
  return L;
}

// per-thread persistent store in global memory: element k of this thread at P[k * stride]
// (one [pslots][NT] block per CTA, so slot offsets are compile-time multiples of NT)
template<class T, int NT> struct PStore
{
  T * p;
  BRBD_DI T & operator[](int k) const { return p[k * NT]; }
};
template<class T, class PS> BRBD_DI void pput6(const PS & P, int o, const Motion<T> & m)
{
  P[o] = m.lin.x; P[o + 1] = m.lin.y; P[o + 2] = m.lin.z; P[o + 3] = m.ang.x; P[o + 4] = m.ang.y; P[o + 5] = m.ang.z;
}
template<class T, class PS> BRBD_DI Motion<T> pget6(const PS & P, int o)
{
  Motion<T> m;
  m.lin = Vec3<T>(P[o], P[o + 1], P[o + 2]);
  m.ang = Vec3<T>(P[o + 3], P[o + 4], P[o + 5]);
  return m;
}

template<class T, int NT>
__global__ void __launch_bounds__(NT, 1)
aba_dfs_kernel(const __grid_constant__ TreePOD<T> m, const AbaLayout L, const T * __restrict__ q, int64_t ldq,
               const T * __restrict__ v, int64_t ldv, const T * __restrict__ tau, int64_t ldtau, T * __restrict__ ddq,
               int64_t ldddq, T * __restrict__ pstore, int64_t B)
{
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  T * sm = reinterpret_cast<T *>(dyn_smem);
  const int tid = threadIdx.x;
  const Slots<T, NT> st{sm + tid};
  const int64_t nthreads = (int64_t)gridDim.x * NT;
  const PStore<T, NT> P{pstore + (int64_t)blockIdx.x * m.pslots * NT + tid};
  const int nj = m.njoints;
  for (int64_t cfg = (int64_t)blockIdx.x * NT + tid; cfg < B; cfg += nthreads)
  {
    const T * __restrict__ qc = q + cfg * ldq;
    const T * __restrict__ vc = v + cfg * ldv;
    const T * __restrict__ tc = tau + cfg * ldtau;
    T * __restrict__ out = ddq + cfg * ldddq;
    SE3<T> X;      // oMi of the joint visited last
    Motion<T> ov = mzero<T>();  // its spatial velocity
    T cI[21];      // contribution (oYaba, of) of an only child, on its way to the parent
    Force<T> cf = fzero<T>();
#pragma unroll
    for (int k = 0; k < 21; ++k) cI[k] = T(0);
    for (int i = 1; i < nj; ++i)
    {
      // ---- pass 1, joint i (aba.hxx:101-138) --------------------------------------------------------
      {
        const JointRec r = m.j[i];
        const SE3<T> Xl = tree_liMi(m, i, r.type, qc + r.idx_q);
        Motion<T> ovp = ov;
        if (r.parent > 0)
        {
          if (r.parent != i - 1)
          {
            const int b = L.oB + ABA_BR * m.j[r.parent].bslot;
            X = get_se3<T>(st, b);
            ovp = get_motion<T>(st, b + 12);
          }
          X = X * Xl;
        }
        else
          X = Xl;
        const int po = r.poff;
        if (r.nvj == 1)
        {
          const Motion<T> J0 = act_S_col(X, r.type, 0);
          pput6(P, po, J0);
          const T vq = __ldg(vc + r.idx_v);
          ov.lin = vq * J0.lin;
          ov.ang = vq * J0.ang;
        }
        else
        {
          for (int k = 0; k < r.nvj; ++k) pput6(P, po + 6 * k, act_S_col(X, r.type, k));
          ov = X.act(tree_joint_velocity(r.type, vc + r.idx_v));
        }
        Motion<T> ab = mzero<T>();
        if (r.parent > 0)
        {
          ov += ovp;
          ab = mcross(ovp, ov);
        }
        pput6(P, po + 6 * r.nvj, ab);
        if (r.bslot >= 0)
        {
          const int b = L.oB + ABA_BR * r.bslot;
          put_se3(st, b, X);
          put_motion(st, b + 12, ov);
        }
        const Inertia<T> Y = act(X, tree_inertia(m, i));
        put_inertia(st, L.oY + 10 * (r.depth - 1), Y);
        put_force(st, L.oF + 6 * (r.depth - 1), fcross(ov, Y * ov));
      }
      // ---- pass 2 for every joint whose subtree is now complete (aba.hxx:152-192) -------------------
      const int stop = m.j[i].stop;
      for (int j = i; j != stop; j = m.j[j].parent)
      {
        const JointRec r = m.j[j];
        const int po = r.poff, nvj = r.nvj, iv = r.idx_v;
        T Ia[21];
        inertia_to_sym6(get_inertia<T>(st, L.oY + 10 * (r.depth - 1)), Ia);
        Force<T> fi = get_force<T>(st, L.oF + 6 * (r.depth - 1));
        if (r.bslot >= 0)
        {
          const int b = L.oB + ABA_BR * r.bslot + 18;
#pragma unroll
          for (int k = 0; k < 21; ++k) Ia[k] += st[b + k];
          fi += get_force<T>(st, b + 21);
        }
        else if (r.nchild == 1)
        {
#pragma unroll
          for (int k = 0; k < 21; ++k) Ia[k] += cI[k];
          fi += cf;
        }
        T fa[6];
        if (nvj == 1)
        {
          const Motion<T> J = pget6<T>(P, po);
          const T ui = __ldg(tc + iv) - dot6(J, fi);
          T Jv[6], U[6];
          m2a(J, Jv);
          sym6_mul(Ia, Jv, U);
          T D = Jv[0] * U[0];
#pragma unroll
          for (int rr = 1; rr < 6; ++rr) D += Jv[rr] * U[rr];
          D += m.armature[iv];
          const T Dinv = T(1) / D;
          T UD[6];
#pragma unroll
          for (int rr = 0; rr < 6; ++rr) { UD[rr] = U[rr] * Dinv; P[po + 12 + rr] = UD[rr]; }
          P[po + 18] = Dinv;
          P[po + 19] = ui;
          if (r.parent > 0)
          {
#pragma unroll
            for (int rr = 0; rr < 6; ++rr)
#pragma unroll
              for (int c = rr; c < 6; ++c) Ia[rr * 6 - (rr * (rr - 1)) / 2 + (c - rr)] -= UD[rr] * U[c];
            T ab[6], Iab[6];
            m2a(pget6<T>(P, po + 6), ab);
            sym6_mul(Ia, ab, Iab);
            f2a(fi, fa);
#pragma unroll
            for (int rr = 0; rr < 6; ++rr) fa[rr] += Iab[rr] + UD[rr] * ui;
          }
        }
        else
        {
          // multi-dof joints (free-flyer, spherical, planar)
          T U[6][6], StU[6][6], Di[6][6], UD[6][6], uj[6];
          for (int k = 0; k < nvj; ++k)
          {
            const Motion<T> J = pget6<T>(P, po + 6 * k);
            uj[k] = __ldg(tc + iv + k) - dot6(J, fi);
            T Jv[6], Uk[6];
            m2a(J, Jv);
            sym6_mul(Ia, Jv, Uk);
            for (int rr = 0; rr < 6; ++rr) U[rr][k] = Uk[rr];
          }
          for (int a = 0; a < nvj; ++a)
          {
            T Jv[6];
            m2a(pget6<T>(P, po + 6 * a), Jv);
            for (int b = 0; b < nvj; ++b)
            {
              T acc = Jv[0] * U[0][b];
              for (int rr = 1; rr < 6; ++rr) acc += Jv[rr] * U[rr][b];
              StU[a][b] = acc;
            }
            StU[a][a] += m.armature[iv + a];
          }
          llt_inverse(nvj, StU, Di);
          for (int rr = 0; rr < 6; ++rr)
            for (int k = 0; k < nvj; ++k)
            {
              T acc = U[rr][0] * Di[0][k];
              for (int c = 1; c < nvj; ++c) acc += U[rr][c] * Di[c][k];
              UD[rr][k] = acc;
            }
          const int oUD = po + 6 * nvj + 6, oD = oUD + 6 * nvj, oU = oD + nvj * nvj;
          for (int k = 0; k < nvj; ++k)
          {
            for (int rr = 0; rr < 6; ++rr) P[oUD + 6 * k + rr] = UD[rr][k];
            for (int c = 0; c < nvj; ++c) P[oD + k * nvj + c] = Di[k][c];
            P[oU + k] = uj[k];
          }
          if (r.parent > 0)
          {
            for (int rr = 0; rr < 6; ++rr)
              for (int c = rr; c < 6; ++c)
              {
                T acc = UD[rr][0] * U[c][0];
                for (int k = 1; k < nvj; ++k) acc += UD[rr][k] * U[c][k];
                Ia[rr * 6 - (rr * (rr - 1)) / 2 + (c - rr)] -= acc;
              }
            T ab[6], Iab[6];
            m2a(pget6<T>(P, po + 6 * nvj), ab);
            sym6_mul(Ia, ab, Iab);
            f2a(fi, fa);
            for (int rr = 0; rr < 6; ++rr)
            {
              T acc = UD[rr][0] * uj[0];
              for (int k = 1; k < nvj; ++k) acc += UD[rr][k] * uj[k];
              fa[rr] += Iab[rr] + acc;
            }
          }
        }
        if (r.parent > 0)
        {
          const JointRec rp = m.j[r.parent];
          if (rp.bslot >= 0)
          {
            const int b = L.oB + ABA_BR * rp.bslot + 18;
            if (j == r.parent + 1)
            { // first child opens the accumulator
#pragma unroll
              for (int k = 0; k < 21; ++k) st[b + k] = Ia[k];
#pragma unroll
              for (int k = 0; k < 6; ++k) st[b + 21 + k] = fa[k];
            }
            else
            {
#pragma unroll
              for (int k = 0; k < 21; ++k) st[b + k] += Ia[k];
#pragma unroll
              for (int k = 0; k < 6; ++k) st[b + 21 + k] += fa[k];
            }
          }
          else
          {
#pragma unroll
            for (int k = 0; k < 21; ++k) cI[k] = Ia[k];
            cf.lin = Vec3<T>(fa[0], fa[1], fa[2]);
            cf.ang = Vec3<T>(fa[3], fa[4], fa[5]);
          }
        }
      }
    }
    // ---- pass 3 (aba.hxx:206-226) -------------------------------------------------------------------
    Motion<T> ag; // oa_gf of the joint visited last
    for (int i = 1; i < nj; ++i)
    {
      const JointRec r = m.j[i];
      const int po = r.poff, nvj = r.nvj, iv = r.idx_v;
      Motion<T> agp;
      if (r.parent == 0)
      {
        agp = mzero<T>();
        agp.lin = Vec3<T>(-m.gravity[0], -m.gravity[1], -m.gravity[2]); // data.oa_gf[0] = -gravity (aba.hxx:260)
      }
      else if (r.parent != i - 1)
        agp = get_motion<T>(st, L.oB + ABA_BR * m.j[r.parent].bslot + 12);
      else
        agp = ag;
      ag = pget6<T>(P, po + 6 * nvj);
      ag += agp;
      T agv[6];
      m2a(ag, agv);
      if (nvj == 1)
      {
        T t2 = P[po + 12] * agv[0];
#pragma unroll
        for (int rr = 1; rr < 6; ++rr) t2 += P[po + 12 + rr] * agv[rr];
        const T dd = P[po + 18] * P[po + 19] - t2;
        out[iv] = dd;
        const Motion<T> J = pget6<T>(P, po);
        ag.lin += dd * J.lin;
        ag.ang += dd * J.ang;
      }
      else
      {
        const int oUD = po + 6 * nvj + 6, oD = oUD + 6 * nvj, oU = oD + nvj * nvj;
        T dd[6];
        for (int k = 0; k < nvj; ++k)
        {
          T t1 = P[oD + k * nvj] * P[oU];
          for (int c = 1; c < nvj; ++c) t1 += P[oD + k * nvj + c] * P[oU + c];
          T t2 = P[oUD + 6 * k] * agv[0];
          for (int rr = 1; rr < 6; ++rr) t2 += P[oUD + 6 * k + rr] * agv[rr];
          dd[k] = t1 - t2;
        }
        for (int k = 0; k < nvj; ++k)
        {
          out[iv + k] = dd[k];
          const Motion<T> J = pget6<T>(P, po + 6 * k);
          ag.lin += dd[k] * J.lin;
          ag.ang += dd[k] * J.ang;
        }
      }
      if (r.bslot >= 0) put_motion(st, L.oB + ABA_BR * r.bslot + 12, ag);
    }
  }
}


// ------------------------------------------------------------------------------------------------------
// v3: pass 1 / pass 2 entirely on chip.  Shared memory holds the J columns of the root path and the branch
// slots; tensor memory (tmem.cuh) holds, per depth, the joint's own world inertia (10), bias force (6) and
// bias acceleration (6).  The per-joint record pass 3 needs (J, a_bias, U Dinv, Dinv, u) is written to the
// per-thread global store once, at the backward step; pass 3 streams it back through a 4-deep per-thread
// cp.async ring.  The first coordinates of the next joint (q, v) and the tau of the next backward step
// also arrive through cp.async slots: prefetching into registers does not survive the register pressure
// of the sweeps (the compiler spills the value at once, which waits for the load).
// ------------------------------------------------------------------------------------------------------
struct AbaTmemLayout
{
  int oJ, oB, oP, nstate;     // shared memory: J (6 x maxpathdof) / record ring, branch slots (45 each), prefetch slots (4)
  int tY, tF, tA, tvals;      // TMEM value offsets, per depth: Y (10), f (6), a_bias (6)
  int tcols;
};
constexpr int ABA_RING = 4;   // pass-3 records in flight per thread
template<class T> inline AbaTmemLayout aba_tmem_layout(int maxpathdof, int maxdepth, int nbranch, int warps)
{
  AbaTmemLayout L;
  L.oJ = 0;
  L.oB = 6 * maxpathdof > 20 * ABA_RING ? 6 * maxpathdof : 20 * ABA_RING; // the J region doubles as the record ring
  L.oP = L.oB + ABA_BR * (nbranch > 0 ? nbranch : 1);
  L.nstate = L.oP + 4; // q, v of the next joint; tau of the next backward step (two alternating slots)
  L.tY = 0;
  L.tF = 10 * maxdepth;
  L.tA = L.tF + 6 * maxdepth;
  L.tvals = L.tA + 6 * maxdepth;
  L.tcols = tmem_round_cols(L.tvals * (int)(sizeof(T) / 4) * ((warps + 3) / 4));
  return L;
}

template<class T> struct AbaContribution { T A[21]; T fa[6]; };

// inverse of an N x N SPD matrix through its Cholesky factor, fully unrolled (registers only)
// — PerformStYSInversion, joint-common-operations.hpp:23-33
template<class T, int N> BRBD_DI void llt_inverse_n(const T (&S)[N][N], T (&Sinv)[N][N])
{
  T Lm[N][N], dinv[N];
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j <= i; ++j)
    {
      T s = S[i][j];
#pragma unroll
      for (int k = 0; k < j; ++k) s -= Lm[i][k] * Lm[j][k];
      if (i == j) { Lm[i][i] = sqrt_t(s); dinv[i] = T(1) / Lm[i][i]; }
      else Lm[i][j] = s * dinv[j];
    }
#pragma unroll
  for (int c = 0; c < N; ++c)
  {
    T y[N];
#pragma unroll
    for (int i = 0; i < N; ++i)
    {
      if (i < c) { y[i] = T(0); continue; }
      T s = (i == c) ? T(1) : T(0);
#pragma unroll
      for (int k = c; k < i; ++k) s -= Lm[i][k] * y[k];
      y[i] = s * dinv[i];
    }
#pragma unroll
    for (int i = N - 1; i >= 0; --i)
    {
      T s = y[i];
#pragma unroll
      for (int k = i + 1; k < N; ++k) s -= Lm[k][i] * Sinv[k][c];
      Sinv[i][c] = s * dinv[i];
    }
  }
}

// Backward step of a multi-dof joint (free-flyer NVJ = 6; spherical, planar NVJ = 3), kept out of line: it
// runs once or twice per configuration and would otherwise dominate the register budget of the 1-dof hot
// path.  By-value in / out so that the caller's oYaba stays in registers; NVJ is a template parameter so that
// every small array has compile-time indices.  in.A = oYaba, in.fa = of; the result holds (if the joint has a
// parent) the contribution to the parent.
template<class T, int NT, int NVJ>
__device__ __noinline__ AbaContribution<T> aba_backward_multidof(const TreePOD<T> & m, const JointRec r, const Slots<T, NT> st, const int oJ,
                                                                 const PStore<T, NT> P, const T * __restrict__ tc, const bool live,
                                                                 const AbaContribution<T> in, const Motion<T> abm)
{
  AbaContribution<T> io = in;
  T (&A)[21] = io.A;
  T (&fa)[6] = io.fa;
  Force<T> fi;
  fi.lin = Vec3<T>(in.fa[0], in.fa[1], in.fa[2]);
  fi.ang = Vec3<T>(in.fa[3], in.fa[4], in.fa[5]);
  const int po = r.poff, iv = r.idx_v;
  T Jm[NVJ][6], U[6][NVJ], StU[NVJ][NVJ], Di[NVJ][NVJ], UD[6][NVJ], uj[NVJ];
#pragma unroll
  for (int k = 0; k < NVJ; ++k)
  {
    const Motion<T> J = get_motion<T>(st, oJ + 6 * (r.pdof + k));
    uj[k] = __ldg(tc + iv + k) - dot6(J, fi);
    T Uk[6];
    m2a(J, Jm[k]);
    sym6_mul(A, Jm[k], Uk);
#pragma unroll
    for (int rr = 0; rr < 6; ++rr) U[rr][k] = Uk[rr];
    if (live) pput6(P, po + 6 * k, J);
  }
#pragma unroll
  for (int a = 0; a < NVJ; ++a)
  {
#pragma unroll
    for (int b = 0; b < NVJ; ++b)
    {
      T acc = Jm[a][0] * U[0][b];
#pragma unroll
      for (int rr = 1; rr < 6; ++rr) acc += Jm[a][rr] * U[rr][b];
      StU[a][b] = acc;
    }
    StU[a][a] += m.armature[iv + a];
  }
  llt_inverse_n<T, NVJ>(StU, Di);
#pragma unroll
  for (int rr = 0; rr < 6; ++rr)
#pragma unroll
    for (int k = 0; k < NVJ; ++k)
    {
      T acc = U[rr][0] * Di[0][k];
#pragma unroll
      for (int c = 1; c < NVJ; ++c) acc += U[rr][c] * Di[c][k];
      UD[rr][k] = acc;
    }
  constexpr int oUDr = 6 * NVJ + 6, oDr = oUDr + 6 * NVJ, oUr = oDr + NVJ * NVJ;
  if (live)
  {
    pput6(P, po + 6 * NVJ, abm);
#pragma unroll
    for (int k = 0; k < NVJ; ++k)
    {
#pragma unroll
      for (int rr = 0; rr < 6; ++rr) P[po + oUDr + 6 * k + rr] = UD[rr][k];
#pragma unroll
      for (int c = 0; c < NVJ; ++c) P[po + oDr + k * NVJ + c] = Di[k][c];
      P[po + oUr + k] = uj[k];
    }
  }
  if (r.parent > 0)
  {
#pragma unroll
    for (int rr = 0; rr < 6; ++rr)
#pragma unroll
      for (int c = rr; c < 6; ++c)
      {
        T acc = UD[rr][0] * U[c][0];
#pragma unroll
        for (int k = 1; k < NVJ; ++k) acc += UD[rr][k] * U[c][k];
        A[rr * 6 - (rr * (rr - 1)) / 2 + (c - rr)] -= acc;
      }
    T ab[6], Iab[6];
    m2a(abm, ab);
    sym6_mul(A, ab, Iab);
#pragma unroll
    for (int rr = 0; rr < 6; ++rr)
    {
      T acc = UD[rr][0] * uj[0];
#pragma unroll
      for (int k = 1; k < NVJ; ++k) acc += UD[rr][k] * uj[k];
      fa[rr] += Iab[rr] + acc;
    }
  }
  return io;
}

// pass 3 of a multi-dof joint, reading its record straight from the per-thread store
template<class T, int NT, int NVJ>
__device__ __noinline__ Motion<T> aba_forward2_multidof(const JointRec r, const PStore<T, NT> P, const Motion<T> agp, T * __restrict__ out)
{
  const int po = r.poff, iv = r.idx_v;
  constexpr int oUDr = 6 * NVJ + 6, oDr = oUDr + 6 * NVJ, oUr = oDr + NVJ * NVJ;
  T Jm[NVJ][6], UD[NVJ][6], Di[NVJ][NVJ], u[NVJ], agv[6];
  const Motion<T> ab = pget6<T>(P, po + 6 * NVJ);
#pragma unroll
  for (int k = 0; k < NVJ; ++k)
  {
#pragma unroll
    for (int rr = 0; rr < 6; ++rr) { Jm[k][rr] = P[po + 6 * k + rr]; UD[k][rr] = P[po + oUDr + 6 * k + rr]; }
#pragma unroll
    for (int c = 0; c < NVJ; ++c) Di[k][c] = P[po + oDr + k * NVJ + c];
    u[k] = P[po + oUr + k];
  }
  Motion<T> ag = ab;
  ag += agp;
  m2a(ag, agv);
  T dd[NVJ];
#pragma unroll
  for (int k = 0; k < NVJ; ++k)
  {
    T t1 = Di[k][0] * u[0];
#pragma unroll
    for (int c = 1; c < NVJ; ++c) t1 += Di[k][c] * u[c];
    T t2 = UD[k][0] * agv[0];
#pragma unroll
    for (int rr = 1; rr < 6; ++rr) t2 += UD[k][rr] * agv[rr];
    dd[k] = t1 - t2;
  }
#pragma unroll
  for (int k = 0; k < NVJ; ++k)
  {
    if (out) out[iv + k] = dd[k];
#pragma unroll
    for (int rr = 0; rr < 6; ++rr) agv[rr] += dd[k] * Jm[k][rr];
  }
  ag.lin = Vec3<T>(agv[0], agv[1], agv[2]);
  ag.ang = Vec3<T>(agv[3], agv[4], agv[5]);
  return ag;
}

template<class T, int NT>
__global__ void __launch_bounds__(NT, 1)
aba_tmem_kernel(const __grid_constant__ TreePOD<T> m, const AbaTmemLayout L, const T * __restrict__ q, int64_t ldq,
                const T * __restrict__ v, int64_t ldv, const T * __restrict__ tau, int64_t ldtau, T * __restrict__ ddq,
                int64_t ldddq, T * __restrict__ pstore, int64_t B)
{
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  __shared__ uint32_t tmem_base_slot;
  T * sm = reinterpret_cast<T *>(dyn_smem);
  const int tid = threadIdx.x, warp = tid >> 5;
  const Slots<T, NT> st{sm + tid};
  const int64_t nthreads = (int64_t)gridDim.x * NT;
  const PStore<T, NT> P{pstore + (int64_t)blockIdx.x * m.pslots * NT + tid};
  const uint32_t tbase = tmem_alloc_cta(L.tcols, &tmem_base_slot);
  const TmemSlots<T> tm{tbase + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)((warp >> 2) * L.tvals * (int)(sizeof(T) / 4))};
  const int nj = m.njoints;
  // every warp runs the same number of rounds (idle lanes shadow the last configuration and do not store)
  const int64_t rounds = (B + nthreads - 1) / nthreads;
  for (int64_t rd = 0; rd < rounds; ++rd)
  {
    const int64_t cfg_raw = rd * nthreads + (int64_t)blockIdx.x * NT + tid;
    const bool live = cfg_raw < B;
    const int64_t cfg = live ? cfg_raw : B - 1;
    const T * __restrict__ qc = q + cfg * ldq;
    const T * __restrict__ vc = v + cfg * ldv;
    const T * __restrict__ tc = tau + cfg * ldtau;
    T * __restrict__ out = live ? ddq + cfg * ldddq : nullptr;
    SE3<T> X;
    Motion<T> ov = mzero<T>();
    // (oYaba, of) of the joint in its backward step; what the step leaves in (A, fA) is the contribution to
    // the parent, which an only child hands over in these very registers
    T A[21];
    Force<T> fA = fzero<T>();
#pragma unroll
    for (int k = 0; k < 21; ++k) A[k] = T(0);
    async_fetch(&st[L.oP], qc + m.j[1].idx_q);
    async_fetch(&st[L.oP + 1], vc + m.j[1].idx_v);
    for (int i = 1; i < nj; ++i)
    {
      Inertia<T> Yown;
      Force<T> fown;
      Motion<T> abown;
      // ---- pass 1, joint i (aba.hxx:101-138) --------------------------------------------------------
      {
        const JointRec r = m.j[i];
        async_wait_all();
        const T q0 = st[L.oP], v0 = st[L.oP + 1];
        if (i + 1 < nj)
        {
          async_fetch(&st[L.oP], qc + m.j[i + 1].idx_q);
          async_fetch(&st[L.oP + 1], vc + m.j[i + 1].idx_v);
        }
        async_fetch(&st[L.oP + 2], tc + r.idx_v); // tau of this joint, for its backward step (a leaf's follows at once)
        const SE3<T> Xl = tree_liMi(m, i, r.type, qc + r.idx_q, q0);
        // (X, ov) are those of the parent: the joint visited last or, after an unwind, the branching joint
        // reloaded at the end of that unwind
        const Motion<T> ovp = ov;
        if (r.parent > 0) X = X * Xl;
        else X = Xl;
        if (r.nvj == 1)
        {
          const Motion<T> J0 = act_S_col(X, r.type, 0);
          put_motion(st, L.oJ + 6 * r.pdof, J0);
          ov.lin = v0 * J0.lin;
          ov.ang = v0 * J0.ang;
        }
        else
        {
          for (int k = 0; k < r.nvj; ++k) put_motion(st, L.oJ + 6 * (r.pdof + k), act_S_col(X, r.type, k));
          ov = X.act(tree_joint_velocity(r.type, vc + r.idx_v));
        }
        abown = mzero<T>();
        if (r.parent > 0)
        {
          ov += ovp;
          abown = mcross(ovp, ov);
        }
        if (r.bslot >= 0)
        {
          const int b = L.oB + ABA_BR * r.bslot;
          put_se3(st, b, X);
          put_motion(st, b + 12, ov);
        }
        Yown = act(X, tree_inertia(m, i));
        fown = fcross(ov, Yown * ov);
        if (r.nchild > 0)
        { // a leaf's pass 2 follows at once: its (Y, f, a_bias) stay in registers
          const T y[10] = {Yown.m, Yown.c.x, Yown.c.y, Yown.c.z, Yown.I.xx, Yown.I.xy, Yown.I.yy, Yown.I.xz, Yown.I.yz, Yown.I.zz};
          const T f6[6] = {fown.lin.x, fown.lin.y, fown.lin.z, fown.ang.x, fown.ang.y, fown.ang.z};
          const T a6[6] = {abown.lin.x, abown.lin.y, abown.lin.z, abown.ang.x, abown.ang.y, abown.ang.z};
          tm.template store<10>(L.tY + 10 * (r.depth - 1), y);
          tm.template store<6>(L.tF + 6 * (r.depth - 1), f6);
          tm.template store<6>(L.tA + 6 * (r.depth - 1), a6);
        }
      }
      // ---- pass 2 for every joint whose subtree is now complete (aba.hxx:152-192) -------------------
      const int stop = m.j[i].stop;
      int tslot = 0;
      for (int j = i; j != stop; j = m.j[j].parent)
      {
        const JointRec r = m.j[j];
        const int po = r.poff, nvj = r.nvj, iv = r.idx_v;
        async_wait_all();
        const T tau0 = st[L.oP + 2 + tslot];
        tslot ^= 1;
        if (r.parent != stop) async_fetch(&st[L.oP + 2 + tslot], tc + m.j[r.parent].idx_v);
        Force<T> fi;
        Motion<T> abm;
        if (j == i)
        {
          inertia_to_sym6(Yown, A);
          fi = fown;
          abm = abown;
        }
        else
        {
          T y[10], f6[6], a6[6], own[21];
          tmem_wait_st();
          tm.template load<10>(L.tY + 10 * (r.depth - 1), y);
          tm.template load<6>(L.tF + 6 * (r.depth - 1), f6);
          tm.template load<6>(L.tA + 6 * (r.depth - 1), a6);
          Inertia<T> Y;
          Y.m = y[0]; Y.c = Vec3<T>(y[1], y[2], y[3]);
          Y.I.xx = y[4]; Y.I.xy = y[5]; Y.I.yy = y[6]; Y.I.xz = y[7]; Y.I.yz = y[8]; Y.I.zz = y[9];
          inertia_to_sym6(Y, own);
          fi.lin = Vec3<T>(f6[0], f6[1], f6[2]); fi.ang = Vec3<T>(f6[3], f6[4], f6[5]);
          abm.lin = Vec3<T>(a6[0], a6[1], a6[2]); abm.ang = Vec3<T>(a6[3], a6[4], a6[5]);
          if (r.bslot >= 0)
          {
            const int b = L.oB + ABA_BR * r.bslot + 18;
#pragma unroll
            for (int k = 0; k < 21; ++k) A[k] = own[k] + st[b + k];
            fi += get_force<T>(st, b + 21);
          }
          else
          { // only child: its contribution is what the previous step left in (A, fA)
#pragma unroll
            for (int k = 0; k < 21; ++k) A[k] += own[k];
            fi += fA;
          }
        }
        T fa[6];
        if (nvj == 1)
        {
          const Motion<T> J = get_motion<T>(st, L.oJ + 6 * r.pdof);
          const T ui = tau0 - dot6(J, fi);
          T Jv[6], U[6];
          m2a(J, Jv);
          sym6_mul(A, Jv, U);
          T D = Jv[0] * U[0];
#pragma unroll
          for (int rr = 1; rr < 6; ++rr) D += Jv[rr] * U[rr];
          D += m.armature[iv];
          const T Dinv = T(1) / D;
          T UD[6];
#pragma unroll
          for (int rr = 0; rr < 6; ++rr) UD[rr] = U[rr] * Dinv;
          if (live)
          {
            pput6(P, po, J);
            pput6(P, po + 6, abm);
#pragma unroll
            for (int rr = 0; rr < 6; ++rr) P[po + 12 + rr] = UD[rr];
            P[po + 18] = Dinv;
            P[po + 19] = ui;
          }
          if (r.parent > 0)
          {
#pragma unroll
            for (int rr = 0; rr < 6; ++rr)
#pragma unroll
              for (int c = rr; c < 6; ++c) A[rr * 6 - (rr * (rr - 1)) / 2 + (c - rr)] -= UD[rr] * U[c];
            T ab[6], Iab[6];
            m2a(abm, ab);
            sym6_mul(A, ab, Iab);
            f2a(fi, fa);
#pragma unroll
            for (int rr = 0; rr < 6; ++rr) fa[rr] += Iab[rr] + UD[rr] * ui;
          }
        }
        else
        {
          AbaContribution<T> io;
#pragma unroll
          for (int k = 0; k < 21; ++k) io.A[k] = A[k];
          f2a(fi, io.fa);
          if (nvj == 6) io = aba_backward_multidof<T, NT, 6>(m, r, st, L.oJ, P, tc, live, io, abm);
          else io = aba_backward_multidof<T, NT, 3>(m, r, st, L.oJ, P, tc, live, io, abm);
#pragma unroll
          for (int k = 0; k < 21; ++k) A[k] = io.A[k];
#pragma unroll
          for (int k = 0; k < 6; ++k) fa[k] = io.fa[k];
        }
        if (r.parent > 0)
        {
          const JointRec rp = m.j[r.parent];
          if (rp.bslot >= 0)
          {
            const int b = L.oB + ABA_BR * rp.bslot + 18;
            if (j == r.parent + 1)
            { // first child opens the accumulator
#pragma unroll
              for (int k = 0; k < 21; ++k) st[b + k] = A[k];
#pragma unroll
              for (int k = 0; k < 6; ++k) st[b + 21 + k] = fa[k];
            }
            else
            {
#pragma unroll
              for (int k = 0; k < 21; ++k) st[b + k] += A[k];
#pragma unroll
              for (int k = 0; k < 6; ++k) st[b + 21 + k] += fa[k];
            }
          }
          else
          {
            fA.lin = Vec3<T>(fa[0], fa[1], fa[2]);
            fA.ang = Vec3<T>(fa[3], fa[4], fa[5]);
          }
        }
      }
      if (stop != i && stop > 0)
      { // the next joint hangs off the branching joint `stop`: fetch its (oMi, ov) now, so that the
        // registers of (X, ov) are free during the unwind above
        const int b = L.oB + ABA_BR * m.j[stop].bslot;
        X = get_se3<T>(st, b);
        ov = get_motion<T>(st, b + 12);
      }
    }
    // ---- pass 3 (aba.hxx:206-226) -------------------------------------------------------------------
    // 1-dof records stream global -> shared through a per-thread cp.async ring that reuses the J region (dead
    // by now).  Shadow lanes read whatever their store slot holds and write nothing.
    {
      Motion<T> ag = mzero<T>();
#pragma unroll
      for (int d = 0; d < ABA_RING; ++d)
      {
        const int i = 1 + d;
        if (i < nj && m.j[i].nvj == 1)
        {
          const int pn = m.j[i].poff, ro = L.oJ + 20 * d;
#pragma unroll
          for (int k = 0; k < 20; ++k) async_fetch(&st[ro + k], &P[pn + k]);
        }
        async_commit();
      }
      for (int i = 1; i < nj; ++i)
      {
        async_wait_group<ABA_RING - 1>();
        const JointRec r = m.j[i];
        const int ro = L.oJ + 20 * ((i - 1) & (ABA_RING - 1));
        Motion<T> agp;
        if (r.parent == 0)
        {
          agp = mzero<T>();
          agp.lin = Vec3<T>(-m.gravity[0], -m.gravity[1], -m.gravity[2]); // data.oa_gf[0] = -gravity (aba.hxx:260)
        }
        else if (r.parent != i - 1)
          agp = get_motion<T>(st, L.oB + ABA_BR * m.j[r.parent].bslot + 12);
        else
          agp = ag;
        if (r.nvj == 1)
        {
          T c[20];
#pragma unroll
          for (int k = 0; k < 20; ++k) c[k] = st[ro + k];
          const T agv[6] = {c[6] + agp.lin.x, c[7] + agp.lin.y, c[8] + agp.lin.z, c[9] + agp.ang.x, c[10] + agp.ang.y, c[11] + agp.ang.z};
          T t2 = c[12] * agv[0];
#pragma unroll
          for (int rr = 1; rr < 6; ++rr) t2 += c[12 + rr] * agv[rr];
          const T dd = c[18] * c[19] - t2;
          if (out) out[r.idx_v] = dd;
          ag.lin = Vec3<T>(agv[0] + dd * c[0], agv[1] + dd * c[1], agv[2] + dd * c[2]);
          ag.ang = Vec3<T>(agv[3] + dd * c[3], agv[4] + dd * c[4], agv[5] + dd * c[5]);
        }
        else if (r.nvj == 6)
          ag = aba_forward2_multidof<T, NT, 6>(r, P, agp, out);
        else
          ag = aba_forward2_multidof<T, NT, 3>(r, P, agp, out);
        if (r.bslot >= 0) put_motion(st, L.oB + ABA_BR * r.bslot + 12, ag);
        {
          const int in = i + ABA_RING; // refill the slot just consumed
          if (in < nj && m.j[in].nvj == 1)
          {
            const int pn = m.j[in].poff;
#pragma unroll
            for (int k = 0; k < 20; ++k) async_fetch(&st[ro + k], &P[pn + k]);
          }
          async_commit();
        }
      }
      async_wait_all();
    }
  }
  tmem_wait_st();
  tmem_free_cta(tbase, L.tcols);
}

} // namespace brbd
