// rnea_dfs.cuh — batched RNEA, v2: one configuration per thread, forward and backward sweeps
// DFS-interleaved, live state in shared memory: per depth the joint's (sin q, cos q) — liMi is rebuilt from it and the
// constant placement in the backward step instead of being kept (12 values) — and f; per open branching joint v, a_gf, f.
// 124 instead of 234 values per configuration for the humanoids: 7 instead of 3 resident warps per SM.
//
// Restates impl::rnea (reference: include/pinocchio/algorithm/rnea.hxx:117-161) with RneaForwardStep
// (rnea.hxx:45-79) and RneaBackwardStep (rnea.hxx:92-107); `tau += armature o a` (rnea.hxx:158).
// Batch driver semantics: rneaInParallel (algorithm/parallel/rnea.hpp:38-83).
#pragma once

#include "tree.cuh"

namespace brbd
{

struct RneaLayout
{
  int oX, oF, oB, nstate; // per depth: (s, c) of the joint (2), f (6); per branch slot: v (6) | a_gf (6) | f acc (6)
};
constexpr int RNEA_BR = 18;
inline RneaLayout rnea_layout(int maxdepth, int nbranch)
{
  RneaLayout L;
  L.oX = 0;
  L.oF = L.oX + 2 * maxdepth;
  L.oB = L.oF + 6 * maxdepth;
  L.nstate = L.oB + RNEA_BR * (nbranch > 0 ? nbranch : 1);
  return L;
}

template<class T, int NT>
__global__ void __launch_bounds__(NT, 1)
rnea_dfs_kernel(const __grid_constant__ TreePOD<T> m, const RneaLayout L, const T * __restrict__ q, int64_t ldq,
                const T * __restrict__ v, int64_t ldv, const T * __restrict__ a, int64_t lda, T * __restrict__ tau,
                int64_t ldtau, int64_t B)
{
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  T * sm = reinterpret_cast<T *>(dyn_smem);
  const int tid = threadIdx.x;
  const Slots<T, NT> st{sm + tid};
  const int64_t nthreads = (int64_t)gridDim.x * NT;
  const int nj = m.njoints;
  for (int64_t cfg = (int64_t)blockIdx.x * NT + tid; cfg < B; cfg += nthreads)
  {
    const T * __restrict__ qc = q + cfg * ldq;
    const T * __restrict__ vc = v + cfg * ldv;
    const T * __restrict__ ac = a + cfg * lda;
    T * __restrict__ out = tau + cfg * ldtau;
    Motion<T> vi = mzero<T>(), ai = mzero<T>(); // v, a_gf of the joint visited last
    Force<T> cf = fzero<T>();                   // force of an only child, already in its parent's frame
    // first coordinates of the next joint, fetched one joint ahead: with 217 KB of shared memory the L1 is ~20 KB and every
    // global load is an L2 round trip (ncu: long_scoreboard 3.7 of 7 cycles per instruction before this)
    T qn = __ldg(qc + m.j[1].idx_q), vn = __ldg(vc + m.j[1].idx_v), an = __ldg(ac + m.j[1].idx_v);
    for (int i = 1; i < nj; ++i)
    {
      // ---- forward step (rnea.hxx:45-79) --------------------------------------------------------------
      {
        const JointRec r = m.j[i];
        const T q0 = qn, v0 = vn, a0 = an;
        if (i + 1 < nj)
        {
          qn = __ldg(qc + m.j[i + 1].idx_q);
          vn = __ldg(vc + m.j[i + 1].idx_v);
          an = __ldg(ac + m.j[i + 1].idx_v);
        }
        T sj, cj;
        tree_sc_joint(m, i, r.type, qc + r.idx_q, q0, &sj, &cj);
        const SE3<T> X = tree_liMi_sc(m, i, r.type, qc + r.idx_q, sj, cj);
        Motion<T> vp = vi, ap = ai;
        if (r.parent == 0)
        { // data.v[0] = 0; data.a_gf[0] = -gravity (rnea.hxx:137-138)
          vp = mzero<T>();
          ap = mzero<T>();
          ap.lin = Vec3<T>(-m.gravity[0], -m.gravity[1], -m.gravity[2]);
        }
        else if (r.parent != i - 1)
        {
          const int b = L.oB + RNEA_BR * m.j[r.parent].bslot;
          vp = get_motion<T>(st, b);
          ap = get_motion<T>(st, b + 6);
        }
        if (r.type <= J_RZ) { vi = mzero<T>(); vi.ang.set(r.type - J_RX, v0); }
        else if (r.type <= J_PZ) { vi = mzero<T>(); vi.lin.set(r.type - J_PX, v0); }
        else vi = tree_joint_velocity(r.type, vc + r.idx_v);
        if (r.parent > 0) vi += X.actInv(vp);
        // a_i = c_J (= 0) + v_i x v_J + S a_J + liMi^-1 a_parent   (rnea.hxx:67-69)
        if (r.type <= J_RZ)
        {
          ai.lin = cross_axis(vi.lin, r.type - J_RX, v0);
          ai.ang = cross_axis(vi.ang, r.type - J_RX, v0);
        }
        else if (r.type <= J_PZ)
        {
          ai.lin = cross_axis(vi.ang, r.type - J_PX, v0);
          ai.ang = Vec3<T>::zero();
        }
        else
          ai = mcross(vi, tree_joint_velocity(r.type, vc + r.idx_v));
        add6(ai, joint_S_row(r.type, 0), a0);
        for (int k = 1; k < r.nvj; ++k) add6(ai, joint_S_row(r.type, k), __ldg(ac + r.idx_v + k));
        ai += X.actInv(ap);
        if (r.bslot >= 0)
        {
          const int b = L.oB + RNEA_BR * r.bslot;
          put_motion(st, b, vi);
          put_motion(st, b + 6, ai);
        }
        const Inertia<T> Y = tree_inertia(m, i);
        Force<T> f = Y * ai;
        f += fcross(vi, Y * vi);
        st[L.oX + 2 * (r.depth - 1)] = sj;
        st[L.oX + 2 * (r.depth - 1) + 1] = cj;
        put_force(st, L.oF + 6 * (r.depth - 1), f);
      }
      // ---- backward steps of the joints whose subtree is complete (rnea.hxx:92-107) --------------------
      const int stop = m.j[i].stop;
      for (int j = i; j != stop; j = m.j[j].parent)
      {
        const JointRec r = m.j[j];
        Force<T> f = get_force<T>(st, L.oF + 6 * (r.depth - 1));
        if (r.bslot >= 0) f += get_force<T>(st, L.oB + RNEA_BR * r.bslot + 12);
        else if (r.nchild == 1) f += cf;
        for (int k = 0; k < r.nvj; ++k)
        {
          // tau = S^T f + armature o a (rnea.hxx:103, :158); the reload of a is skipped where the armature is zero (the default)
          const T arm = m.armature[r.idx_v + k];
          T t = get6(f, joint_S_row(r.type, k));
          if (arm != T(0)) t += arm * __ldg(ac + r.idx_v + k);
          out[r.idx_v + k] = t;
        }
        if (r.parent > 0)
        {
          const Force<T> fp = tree_liMi_sc(m, j, r.type, qc + r.idx_q, st[L.oX + 2 * (r.depth - 1)], st[L.oX + 2 * (r.depth - 1) + 1]).act(f);
          const JointRec rp = m.j[r.parent];
          if (rp.bslot >= 0)
          {
            const int b = L.oB + RNEA_BR * rp.bslot + 12;
            if (j == r.parent + 1) put_force(st, b, fp);
            else put_force(st, b, get_force<T>(st, b) + fp);
          }
          else
            cf = fp;
        }
      }
    }
  }
}

} // namespace brbd
