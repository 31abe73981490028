// crba.cuh — batched CRBA (joint-space inertia matrix), one configuration per thread.
//
// Restates impl::crbaWorldConvention (reference: include/pinocchio/algorithm/crba.hxx:498-548) with
// CrbaWorldConventionForwardStep (crba.hxx:35-58) and CrbaWorldConventionBackwardStep (crba.hxx:80-99).
// The reference fills M row-block by row-block: M[i, subtree(i)] = J_i^T Ag[:, subtree(i)] with
// Ag_j = oYcrb_j J_j.  The same dot products are evaluated here column by column — when joint j's
// composite inertia is complete, column j = { J_a^T Ag_j : a ancestor-or-self of j } — because one
// column of the caller's col-major matrix is contiguous in memory and can leave the SM with
// coalesced stores (ColumnEmitter).  Entries outside the tree sparsity and the strictly-lower
// triangle are written as zeros (crba.hpp:15-22; a fresh Data holds zeros there, data.hxx:43);
// armature goes on the diagonal (crba.hxx:532).  No batched crba exists in the reference; the
// single-configuration entry point is crba(model, data, q) (crba.hpp:47-51).
#pragma once

#include "engine.cuh"
#include "rnea.cuh"

namespace brbd
{

template<class T> BRBD_DI void store_inertia(T * d, const Inertia<T> & Y)
{
  d[0] = Y.m; d[1] = Y.c.x; d[2] = Y.c.y; d[3] = Y.c.z;
  d[4] = Y.I.xx; d[5] = Y.I.xy; d[6] = Y.I.yy; d[7] = Y.I.xz; d[8] = Y.I.yz; d[9] = Y.I.zz;
}
template<class T> BRBD_DI Inertia<T> load_inertia(const T * d)
{
  Inertia<T> Y;
  Y.m = d[0]; Y.c = Vec3<T>(d[1], d[2], d[3]);
  Y.I.xx = d[4]; Y.I.xy = d[5]; Y.I.yy = d[6]; Y.I.xz = d[7]; Y.I.yz = d[8]; Y.I.zz = d[9];
  return Y;
}

// All 32 lanes of the warp must call this together (the emitter is warp-cooperative).
template<class T>
BRBD_DI void crba_thread(const ModelPOD<T> & m, const T * q, ColumnEmitter<T> & em, T * gM, int64_t ldM, int nc)
{
  T J_s[MAXNV][6];
  T Y_s[MAXJ][10];
  T oMi_d[MAXDEPTH][12];
  const int nj = m.njoints, nv = m.nv;
  for (int i = 1; i < nj; ++i)
  {
    const int type = m.type[i], parent = m.parent[i], iq = m.idx_q[i], iv = m.idx_v[i], d = m.depth[i], nvj = m.nvj[i];
    SE3<T> X = joint_liMi(m, i, type, q + iq);
    if (parent > 0) X = load_se3(oMi_d[d - 1]) * X;
    store_se3(oMi_d[d], X);
    for (int k = 0; k < nvj; ++k) store6(J_s[iv + k], act_S_col(X, type, k));
    store_inertia(Y_s[i], act(X, model_inertia(m, i)));
  }
  for (int i = nj - 1; i > 0; --i)
  {
    const int parent = m.parent[i], iv = m.idx_v[i], nvj = m.nvj[i];
    const Inertia<T> Y = load_inertia(Y_s[i]);
    for (int k = 0; k < nvj; ++k)
    {
      const int col = iv + k;
      const Force<T> F = Y * load_motion(J_s[col]); // Ag column (crba.hxx:91)
      for (int a = i; a > 0; a = m.parent[a])
      {
        const int ia = m.idx_v[a], na = m.nvj[a];
        for (int r = 0; r < na; ++r)
        {
          T val = dot6(load_motion(J_s[ia + r]), F);
          if (ia + r == col) val += m.armature[col];
          em.put(ia + r, val);
        }
      }
      em.flush(gM + (int64_t)col * nv, ldM, nc);
    }
    if (parent > 0)
    {
      Inertia<T> Yp = load_inertia(Y_s[parent]);
      Yp += Y;
      store_inertia(Y_s[parent], Yp);
    }
  }
}

// lanes >= nc adopt a copy of the last valid staged row so that the whole warp can run the
// warp-cooperative sweeps uniformly
template<class T> BRBD_DI void pad_tile_rows(T * s, int pad, int rows, int nc, int lane)
{
  if (lane >= nc)
    for (int r = 0; r < rows; ++r) s[lane * pad + r] = s[(nc - 1) * pad + r];
}

template<class T>
__global__ void __launch_bounds__(512)
crba_kernel(const ModelPOD<T> * __restrict__ gm, const T * __restrict__ q, int64_t ldq, T * __restrict__ Mout,
            int64_t ldM, int64_t B)
{
  __shared__ ModelPOD<T> m;
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  copy_model_to_smem(&m, gm);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int qpad = m.nq | 1, vpad = m.nv | 1;
  T * sq = reinterpret_cast<T *>(dyn_smem) + (size_t)warp * 32 * (qpad + vpad);
  T * se = sq + 32 * qpad;
  ColumnEmitter<T> em;
  em.init(se, vpad, m.nv, lane);
  const int64_t ntiles = (B + 31) / 32;
  for (int64_t tile = (int64_t)blockIdx.x * nw + warp; tile < ntiles; tile += (int64_t)gridDim.x * nw)
  {
    const int64_t c0 = tile * 32;
    const int nc = (int)((B - c0) < 32 ? (B - c0) : 32);
    tile_load(sq, qpad, q + c0 * ldq, ldq, m.nq, nc, lane);
    BRBD_SYNCWARP();
    pad_tile_rows(sq, qpad, m.nq, nc, lane);
    BRBD_SYNCWARP();
    crba_thread(m, sq + lane * qpad, em, Mout + c0 * ldM, ldM, nc);
    BRBD_SYNCWARP();
  }
}

} // namespace brbd
