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
// per-thread global store once, at the backward step, and pass 3 streams it back one joint ahead of use.
// ------------------------------------------------------------------------------------------------------
struct AbaTmemLayout
{
  int oJ, oB, nstate;         // shared memory: J (6 x maxpathdof), branch slots (45 each)
  int tY, tF, tA, tvals;      // TMEM value offsets, per depth: Y (10), f (6), a_bias (6)
  int tcols;
};
template<class T> inline AbaTmemLayout aba_tmem_layout(int maxpathdof, int maxdepth, int nbranch, int warps)
{
  AbaTmemLayout L;
  L.oJ = 0;
  L.oB = 6 * maxpathdof;
  L.nstate = L.oB + ABA_BR * (nbranch > 0 ? nbranch : 1);
  L.tY = 0;
  L.tF = 10 * maxdepth;
  L.tA = L.tF + 6 * maxdepth;
  L.tvals = L.tA + 6 * maxdepth;
  L.tcols = tmem_round_cols(L.tvals * (int)(sizeof(T) / 4) * ((warps + 3) / 4));
  return L;
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
    T * __restrict__ out = ddq + cfg * ldddq;
    SE3<T> X;
    Motion<T> ov = mzero<T>();
    T cI[21];
    Force<T> cf = fzero<T>();
#pragma unroll
    for (int k = 0; k < 21; ++k) cI[k] = T(0);
    T qnext = __ldg(qc + m.j[1].idx_q), vnext = __ldg(vc + m.j[1].idx_v);
    for (int i = 1; i < nj; ++i)
    {
      Inertia<T> Yown;
      Force<T> fown;
      Motion<T> abown;
      // ---- pass 1, joint i (aba.hxx:101-138) --------------------------------------------------------
      {
        const JointRec r = m.j[i];
        const T q0 = qnext, v0 = vnext;
        if (i + 1 < nj) { qnext = __ldg(qc + m.j[i + 1].idx_q); vnext = __ldg(vc + m.j[i + 1].idx_v); }
        const SE3<T> Xl = tree_liMi(m, i, r.type, qc + r.idx_q, q0);
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
      for (int j = i; j != stop; j = m.j[j].parent)
      {
        const JointRec r = m.j[j];
        const int po = r.poff, nvj = r.nvj, iv = r.idx_v;
        const T tau0 = __ldg(tc + iv);
        T Ia[21];
        Force<T> fi;
        Motion<T> abm;
        if (j == i)
        {
          inertia_to_sym6(Yown, Ia);
          fi = fown;
          abm = abown;
        }
        else
        {
          T y[10], f6[6], a6[6];
          tmem_wait_st();
          tm.template load<10>(L.tY + 10 * (r.depth - 1), y);
          tm.template load<6>(L.tF + 6 * (r.depth - 1), f6);
          tm.template load<6>(L.tA + 6 * (r.depth - 1), a6);
          Inertia<T> Y;
          Y.m = y[0]; Y.c = Vec3<T>(y[1], y[2], y[3]);
          Y.I.xx = y[4]; Y.I.xy = y[5]; Y.I.yy = y[6]; Y.I.xz = y[7]; Y.I.yz = y[8]; Y.I.zz = y[9];
          inertia_to_sym6(Y, Ia);
          fi.lin = Vec3<T>(f6[0], f6[1], f6[2]); fi.ang = Vec3<T>(f6[3], f6[4], f6[5]);
          abm.lin = Vec3<T>(a6[0], a6[1], a6[2]); abm.ang = Vec3<T>(a6[3], a6[4], a6[5]);
        }
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
          const Motion<T> J = get_motion<T>(st, L.oJ + 6 * r.pdof);
          const T ui = tau0 - dot6(J, fi);
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
              for (int c = rr; c < 6; ++c) Ia[rr * 6 - (rr * (rr - 1)) / 2 + (c - rr)] -= UD[rr] * U[c];
            T ab[6], Iab[6];
            m2a(abm, ab);
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
            const Motion<T> J = get_motion<T>(st, L.oJ + 6 * (r.pdof + k));
            uj[k] = __ldg(tc + iv + k) - dot6(J, fi);
            T Jv[6], Uk[6];
            m2a(J, Jv);
            sym6_mul(Ia, Jv, Uk);
            for (int rr = 0; rr < 6; ++rr) U[rr][k] = Uk[rr];
            if (live) pput6(P, po + 6 * k, J);
          }
          for (int a = 0; a < nvj; ++a)
          {
            T Jv[6];
            m2a(get_motion<T>(st, L.oJ + 6 * (r.pdof + a)), Jv);
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
          if (live)
          {
            pput6(P, po + 6 * nvj, abm);
            for (int k = 0; k < nvj; ++k)
            {
              for (int rr = 0; rr < 6; ++rr) P[oUD + 6 * k + rr] = UD[rr][k];
              for (int c = 0; c < nvj; ++c) P[oD + k * nvj + c] = Di[k][c];
              P[oU + k] = uj[k];
            }
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
            m2a(abm, ab);
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
            {
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
    // ---- pass 3 (aba.hxx:206-226): the record of a 1-dof joint is fetched one joint ahead -------------
    if (live)
    {
      Motion<T> ag = mzero<T>();
      T rec[20];
      bool have = false;
      if (m.j[1].nvj == 1)
      {
#pragma unroll
        for (int k = 0; k < 20; ++k) rec[k] = P[m.j[1].poff + k];
        have = true;
      }
      for (int i = 1; i < nj; ++i)
      {
        const JointRec r = m.j[i];
        const int po = r.poff, nvj = r.nvj, iv = r.idx_v;
        T cur[20];
        const bool cur_ok = have;
#pragma unroll
        for (int k = 0; k < 20; ++k) cur[k] = rec[k];
        have = false;
        if (i + 1 < nj && m.j[i + 1].nvj == 1)
        {
          const int pn = m.j[i + 1].poff;
#pragma unroll
          for (int k = 0; k < 20; ++k) rec[k] = P[pn + k];
          have = true;
        }
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
        if (nvj == 1 && cur_ok)
        {
          T agv[6] = {cur[6] + agp.lin.x, cur[7] + agp.lin.y, cur[8] + agp.lin.z, cur[9] + agp.ang.x, cur[10] + agp.ang.y, cur[11] + agp.ang.z};
          T t2 = cur[12] * agv[0];
#pragma unroll
          for (int rr = 1; rr < 6; ++rr) t2 += cur[12 + rr] * agv[rr];
          const T dd = cur[18] * cur[19] - t2;
          out[iv] = dd;
          ag.lin = Vec3<T>(agv[0] + dd * cur[0], agv[1] + dd * cur[1], agv[2] + dd * cur[2]);
          ag.ang = Vec3<T>(agv[3] + dd * cur[3], agv[4] + dd * cur[4], agv[5] + dd * cur[5]);
        }
        else
        {
          ag = pget6<T>(P, po + 6 * nvj);
          ag += agp;
          T agv[6];
          m2a(ag, agv);
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
  tmem_wait_st();
  tmem_free_cta(tbase, L.tcols);
}

} // namespace brbd
