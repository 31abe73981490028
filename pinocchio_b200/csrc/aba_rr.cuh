// aba_rr.cuh — batched ABA (WORLD convention), v4: the backward sweep RECOMPUTES what v3 kept per tree depth.
//
// Same algorithm and pass structure as aba_tmem_kernel (aba_dfs.cuh; reference impl::abaWorldConvention,
// include/pinocchio/algorithm/aba.hxx:242-293): DFS-interleaved pass 1 / pass 2 on chip, pass 3 from the per-thread record
// store.  v3 parked, per depth, the joint's world inertia, bias force and bias acceleration (22 values) in tensor memory and
// the J columns of the root path in shared memory: 484 TMEM columns and 190 shared-memory values per configuration, i.e.
// 4 warps per SM.  But when the backward sweep reaches joint j it holds (oMi_j, ov_j) in registers, from which all of that
// follows — and the parent's pair follows from the child's:
//     oMi_parent = oMi_j liMi_j^-1        (liMi_j from (sin q_j, cos q_j) and the constant placement)
//     ov_parent  = ov_j - J_j v_j
//     Y_j = oMi_j.act(I_j),  f_j = ov_j x* (Y_j ov_j),  a_bias_j = ov_parent x ov_j,  J_j = oMi_j.act(S_j)
// so a depth only keeps (sin q, cos q, v) — 3 values — and a leaf-to-root unwind costs ~85 extra flops per joint (+11 %).
// Everything per-depth and per-branch now fits in 246 TMEM columns, shared memory only holds the pass-3 record ring:
// 8 warps per SM fit (7 are used for 65 536 configurations: 1.98 rounds of the persistent grid instead of 3.46).
// The recomputed oMi / ov differ from the forward ones by rounding only (parity tolerance unaffected).
#pragma once

#include "aba_dfs.cuh"

namespace brbd
{

#ifndef BRBD_ABA_RR_RING
#define BRBD_ABA_RR_RING 2
#endif
constexpr int ABA_RR_RING = BRBD_ABA_RR_RING; // pass-3 records in flight per thread
static_assert(20 * ABA_RR_RING >= 27, "the hand-over slots alias the record ring");
struct AbaRRLayout
{
  int oR, oP, oA, nstate; // oA: hand-over of an only child's (oYaba 21 | of 6) to its parent's step, aliases the ring (idle in passes 1-2)
                          // shared memory: pass-3 record ring (20 x ABA_RR_RING), prefetch (4).  Kept small on purpose: the kernel
                          // runs at 255 registers with ~70 local-memory spill accesses per joint, and what shared memory
                          // does not take is L1 for them (ring of 4 -> 2: 0.367 -> 0.294 ms; 3: 0.304 ms)
  int tS, tB, tvals;      // TMEM values: per depth (s, c, v) (3), per branch slot oMi 12 | ov 6 | Ia acc 21 | f acc 6 (45)
  int tcols;
};
template<class T> inline AbaRRLayout aba_rr_layout(int maxdepth, int nbranch, int warps)
{
  AbaRRLayout L;
  const int nb = nbranch > 0 ? nbranch : 1;
  L.oR = 0;
  L.oP = 20 * ABA_RR_RING;
  L.oA = L.oR;
  L.nstate = L.oP + 4;
  L.tS = 0;
  L.tB = 3 * maxdepth;
  L.tvals = L.tB + ABA_BR * nb;
  L.tcols = tmem_round_cols(L.tvals * (int)(sizeof(T) / 4) * ((warps + 3) / 4));
  return L;
}

template<class T> BRBD_DI Mat3<T> mul_bt(const Mat3<T> & A, const Mat3<T> & B) // A B^T
{
  Mat3<T> r;
  r.c0 = A * Vec3<T>(B.c0.x, B.c1.x, B.c2.x);
  r.c1 = A * Vec3<T>(B.c0.y, B.c1.y, B.c2.y);
  r.c2 = A * Vec3<T>(B.c0.z, B.c1.z, B.c2.z);
  return r;
}

// multi-dof backward step with the J columns rebuilt from oMi (see aba_backward_multidof in aba_dfs.cuh).  Out of line: it runs
// once or twice per configuration and would otherwise dominate the register budget of the 1-dof path.  (oYaba, of) come in and
// the contribution to the parent goes out through the 27 hand-over slots st[oA ..]: by-value structs travel through the stack.
template<class T, int NT, int NVJ>
__device__ __noinline__ void aba_rr_backward_multidof(const TreePOD<T> & m, const JointRec r, const SE3<T> X, const Slots<T, NT> st, const int oA,
                                                      const PStore<T, NT> P, const T * __restrict__ tc, const bool live, const Motion<T> abm)
{
  T A[21], fa[6];
#pragma unroll
  for (int k = 0; k < 21; ++k) A[k] = st[oA + k];
#pragma unroll
  for (int k = 0; k < 6; ++k) fa[k] = st[oA + 21 + k];
  Force<T> fi;
  fi.lin = Vec3<T>(fa[0], fa[1], fa[2]);
  fi.ang = Vec3<T>(fa[3], fa[4], fa[5]);
  const int po = r.poff, iv = r.idx_v;
  T Jm[NVJ][6], U[6][NVJ], StU[NVJ][NVJ], Di[NVJ][NVJ], UD[6][NVJ], uj[NVJ];
#pragma unroll
  for (int k = 0; k < NVJ; ++k)
  {
    const Motion<T> J = act_S_col(X, r.type, k);
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
#pragma unroll
    for (int k = 0; k < 21; ++k) st[oA + k] = A[k];
#pragma unroll
    for (int k = 0; k < 6; ++k) st[oA + 21 + k] = fa[k];
  }
}

template<class T, int NT>
__global__ void __launch_bounds__(NT, 1)
aba_rr_kernel(const __grid_constant__ TreePOD<T> m, const AbaRRLayout L, const T * __restrict__ q, int64_t ldq,
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
    // (oYaba, of) of a joint only live in registers inside its backward step; what the step leaves for the parent goes through
    // the branching joint's TMEM accumulator or, for an only child, through 27 shared-memory slots
    async_fetch(&st[L.oP], qc + m.j[1].idx_q);
    async_fetch(&st[L.oP + 1], vc + m.j[1].idx_v);
    for (int i = 1; i < nj; ++i)
    {
      T si, ci, vi; // (sin q, cos q, v) of joint i (1-dof joints)
      // ---- pass 1, joint i: kinematics only (aba.hxx:101-131) ---------------------------------------
      {
        const JointRec r = m.j[i];
        async_wait_all();
        const T q0 = st[L.oP];
        vi = st[L.oP + 1];
        if (i + 1 < nj)
        {
          async_fetch(&st[L.oP], qc + m.j[i + 1].idx_q);
          async_fetch(&st[L.oP + 1], vc + m.j[i + 1].idx_v);
        }
        async_fetch(&st[L.oP + 2], tc + r.idx_v); // tau of this joint, for its backward step (a leaf's follows at once)
        tree_sc_joint(m, i, r.type, qc + r.idx_q, q0, &si, &ci);
        const SE3<T> Xl = tree_liMi_sc(m, i, r.type, qc + r.idx_q, si, ci);
        if (r.parent > 0)
        {
          X = X * Xl;
          if (r.nvj == 1)
          {
            const Motion<T> J0 = act_S_col(X, r.type, 0);
            ov.lin += vi * J0.lin;
            ov.ang += vi * J0.ang;
          }
          else
            ov += X.act(tree_joint_velocity(r.type, vc + r.idx_v));
        }
        else
        {
          X = Xl;
          if (r.nvj == 1)
          {
            const Motion<T> J0 = act_S_col(X, r.type, 0);
            ov.lin = vi * J0.lin;
            ov.ang = vi * J0.ang;
          }
          else
            ov = X.act(tree_joint_velocity(r.type, vc + r.idx_v));
        }
        if (r.nchild > 0)
        { // a leaf's pass 2 follows at once with (s, c, v) still in registers
          const T s3[3] = {si, ci, vi};
          tm.template store<3>(L.tS + 3 * (r.depth - 1), s3);
        }
        if (r.bslot >= 0)
        {
          const T x[18] = {X.R.c0.x, X.R.c0.y, X.R.c0.z, X.R.c1.x, X.R.c1.y, X.R.c1.z, X.R.c2.x, X.R.c2.y, X.R.c2.z,
                           X.p.x, X.p.y, X.p.z, ov.lin.x, ov.lin.y, ov.lin.z, ov.ang.x, ov.ang.y, ov.ang.z};
          tm.template store<18>(L.tB + ABA_BR * r.bslot, x);
        }
      }
      // ---- pass 2 for every joint whose subtree is now complete (aba.hxx:152-192) -------------------
      // (X, ov) are those of joint j throughout; they move to the parent at the end of each step
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
        T A[21];
        T sj = si, cj = ci, vj = vi;
        if (j != i)
        {
          T s3[3];
          tmem_wait_st();
          tm.template load<3>(L.tS + 3 * (r.depth - 1), s3);
          sj = s3[0]; cj = s3[1]; vj = s3[2];
        }
        // the joint's own quantities, from (oMi_j, ov_j)
        const Inertia<T> Y = act(X, tree_inertia(m, j));
        Force<T> fi = fcross(ov, Y * ov);
        Motion<T> ovp = mzero<T>(), abm = mzero<T>();
        Motion<T> J0 = mzero<T>();
        if (nvj == 1) J0 = act_S_col(X, r.type, 0);
        if (r.parent > 0)
        {
          if (nvj == 1)
          {
            ovp.lin = ov.lin - vj * J0.lin;
            ovp.ang = ov.ang - vj * J0.ang;
          }
          else
            ovp = ov - X.act(tree_joint_velocity(r.type, vc + iv));
          abm = mcross(ovp, ov);
        }
        inertia_to_sym6(Y, A);
        if (j != i)
        {
          if (r.bslot >= 0)
          { // the children's contributions sit in the accumulator; added in chunks (register pressure)
            const int b = L.tB + ABA_BR * r.bslot + 18;
#pragma unroll
            for (int ch = 0; ch < 3; ++ch)
            {
              T acc[9];
              tm.template load<9>(b + 9 * ch, acc);
#pragma unroll
              for (int k = 0; k < 9; ++k)
                if (9 * ch + k < 21) A[9 * ch + k] += acc[k];
              if (ch == 2)
              {
                fi.lin += Vec3<T>(acc[3], acc[4], acc[5]);
                fi.ang += Vec3<T>(acc[6], acc[7], acc[8]);
              }
            }
          }
          else
          { // only child: its contribution is what the previous step left in the hand-over slots
#pragma unroll
            for (int k = 0; k < 21; ++k) A[k] += st[L.oA + k];
            fi += get_force<T>(st, L.oA + 21);
          }
        }
        T fa[6];
        if (nvj == 1)
        {
          const T ui = tau0 - dot6(J0, fi);
          T Jv[6], U[6];
          m2a(J0, Jv);
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
            pput6(P, po, J0);
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
#pragma unroll
          for (int k = 0; k < 21; ++k) st[L.oA + k] = A[k];
          put_force(st, L.oA + 21, fi);
          if (nvj == 6) aba_rr_backward_multidof<T, NT, 6>(m, r, X, st, L.oA, P, tc, live, abm);
          else aba_rr_backward_multidof<T, NT, 3>(m, r, X, st, L.oA, P, tc, live, abm);
          if (r.parent > 0)
          {
#pragma unroll
            for (int k = 0; k < 21; ++k) A[k] = st[L.oA + k];
#pragma unroll
            for (int k = 0; k < 6; ++k) fa[k] = st[L.oA + 21 + k];
          }
        }
        if (r.parent > 0)
        {
          const JointRec rp = m.j[r.parent];
          if (rp.bslot >= 0)
          {
            const int b = L.tB + ABA_BR * rp.bslot + 18;
            if (j == r.parent + 1)
            { // first child opens the accumulator
              tm.template store<21>(b, A);
              tm.template store<6>(b + 21, fa);
            }
            else
            {
              tmem_wait_st();
#pragma unroll
              for (int ch = 0; ch < 3; ++ch)
              {
                T acc[9];
                tm.template load<9>(b + 9 * ch, acc);
#pragma unroll
                for (int k = 0; k < 9; ++k)
                {
                  const int e = 9 * ch + k;
                  acc[k] += e < 21 ? A[e] : fa[e - 21];
                }
                tm.template store<9>(b + 9 * ch, acc);
              }
            }
          }
          else
          {
#pragma unroll
            for (int k = 0; k < 21; ++k) st[L.oA + k] = A[k];
#pragma unroll
            for (int k = 0; k < 6; ++k) st[L.oA + 21 + k] = fa[k];
          }
          if (r.parent != stop)
          { // the unwind goes on with the parent: oMi_parent = oMi_j liMi_j^-1, ov_parent = ov_j - J_j v_j
            const SE3<T> Xl = tree_liMi_sc(m, j, r.type, qc + r.idx_q, sj, cj);
            SE3<T> Xp;
            Xp.R = mul_bt(X.R, Xl.R);
            Xp.p = X.p - Xp.R * Xl.p;
            X = Xp;
            ov = ovp;
          }
        }
      }
      if (stop != i && stop > 0)
      { // the next joint hangs off the branching joint `stop`: its (oMi, ov) as the forward step left them
        T x[18];
        tmem_wait_st();
        tm.template load<18>(L.tB + ABA_BR * m.j[stop].bslot, x);
        X.R.c0 = Vec3<T>(x[0], x[1], x[2]); X.R.c1 = Vec3<T>(x[3], x[4], x[5]); X.R.c2 = Vec3<T>(x[6], x[7], x[8]);
        X.p = Vec3<T>(x[9], x[10], x[11]);
        ov.lin = Vec3<T>(x[12], x[13], x[14]); ov.ang = Vec3<T>(x[15], x[16], x[17]);
      }
    }
    // ---- pass 3 (aba.hxx:206-226), as aba_tmem_kernel: records stream back through a per-thread cp.async ring ----------
    {
      Motion<T> ag = mzero<T>();
      int rslot = 0;
#pragma unroll
      for (int d = 0; d < ABA_RR_RING; ++d)
      {
        const int i = 1 + d;
        if (i < nj && m.j[i].nvj == 1)
        {
          const int pn = m.j[i].poff, ro = L.oR + 20 * d;
#pragma unroll
          for (int k = 0; k < 20; ++k) async_fetch(&st[ro + k], &P[pn + k]);
        }
        async_commit();
      }
      for (int i = 1; i < nj; ++i)
      {
        async_wait_group<ABA_RR_RING - 1>();
        const JointRec r = m.j[i];
        const int ro = L.oR + 20 * rslot;
        rslot = rslot + 1 == ABA_RR_RING ? 0 : rslot + 1;
        Motion<T> agp;
        if (r.parent == 0)
        {
          agp = mzero<T>();
          agp.lin = Vec3<T>(-m.gravity[0], -m.gravity[1], -m.gravity[2]); // data.oa_gf[0] = -gravity (aba.hxx:260)
        }
        else if (r.parent != i - 1)
        {
          T g6[6];
          tmem_wait_st();
          tm.template load<6>(L.tB + ABA_BR * m.j[r.parent].bslot + 12, g6);
          agp.lin = Vec3<T>(g6[0], g6[1], g6[2]);
          agp.ang = Vec3<T>(g6[3], g6[4], g6[5]);
        }
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
        if (r.bslot >= 0)
        { // oa_gf of an open branching joint: in its TMEM slot, where ov sat during passes 1 and 2
          const T g6[6] = {ag.lin.x, ag.lin.y, ag.lin.z, ag.ang.x, ag.ang.y, ag.ang.z};
          tm.template store<6>(L.tB + ABA_BR * r.bslot + 12, g6);
        }
        {
          const int in = i + ABA_RR_RING; // refill the slot just consumed
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
