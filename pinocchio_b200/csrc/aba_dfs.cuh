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

} // namespace brbd
