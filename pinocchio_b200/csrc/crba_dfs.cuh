// crba_dfs.cuh — batched CRBA, v2: one configuration per thread, DFS-interleaved sweeps, all live state
// in shared memory, matrix columns leave the SM through a warp-cooperative coalesced emitter.
//
// Restates impl::crbaWorldConvention (reference: include/pinocchio/algorithm/crba.hxx:498-548):
//   forward step  CrbaWorldConventionForwardStep  (crba.hxx:35-58):  oMi, J_cols = oMi.act(S), oYcrb = oMi.act(I)
//   backward step CrbaWorldConventionBackwardStep (crba.hxx:80-99):  Ag_cols = oYcrb J_cols,
//                 M[i, subtree(i)] = J_i^T Ag[:, subtree(i)],  oYcrb[parent] += oYcrb[i]
//   armature on the diagonal (crba.hxx:532).
// The same dot products are produced column by column (column j = { J_a^T Ag_j : a ancestor-or-self of
// joint(j) }) because one column of the caller's col-major nv x nv matrix is contiguous in memory.  The
// backward step of joint j runs as soon as its subtree is complete (tree.cuh), so only the root path is
// live: J columns of the path (6 x maxpathdof), oYcrb per depth (10 x maxdepth) and oMi of the open
// branching joints (12 x nbranch).  Entries outside the tree sparsity (incl. the strictly-lower triangle)
// are written as zeros (crba.hpp:15-22; a fresh Data holds zeros there, data.hxx:43).
#pragma once

#include "tree.cuh"

namespace brbd
{

struct CrbaLayout
{
  int oJ, oY, oX, nstate; // slot offsets of the per-thread state, total slots
  int epad;               // emitter row length (>= nv, odd)
};
inline CrbaLayout crba_layout(int maxpathdof, int maxdepth, int nbranch, int nv)
{
  CrbaLayout L;
  L.oJ = 0;
  L.oY = L.oJ + 6 * maxpathdof;
  L.oX = L.oY + 10 * maxdepth;
  L.nstate = L.oX + 12 * (nbranch > 0 ? nbranch : 1);
  L.epad = nv | 1;
  return L;
}

template<class T, int NT>
__global__ void __launch_bounds__(NT, 1)
crba_dfs_kernel(const __grid_constant__ TreePOD<T> m, const CrbaLayout L, const T * __restrict__ q, int64_t ldq,
                T * __restrict__ Mout, int64_t ldM, int64_t B)
{
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  T * sm = reinterpret_cast<T *>(dyn_smem);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int nw = NT / 32;
  const Slots<T, NT> st{sm + tid};
  T * em = sm + (size_t)L.nstate * NT + (size_t)warp * 32 * L.epad; // [32][epad], row = configuration of the tile
  T * myrow = em + lane * L.epad;
  const int nj = m.njoints, nv = m.nv;
  // The emitter rows hold zeros outside the entries of the column being assembled: a joint zeroes its own
  // rows once its columns are flushed (it is never again an ancestor of a column of this tile).
  for (int k = lane; k < 32 * L.epad; k += 32) em[k] = T(0);
  __syncwarp();
  // flush iteration state: element e = lane + 32 t of the (nc x nv) column block -> (config c, row rr)
  const int rr0 = lane % nv, c0l = lane / nv, r32 = 32 % nv, q32 = 32 / nv;
  const int so0 = c0l * L.epad + rr0, ds = q32 * L.epad + r32, dsw = L.epad - nv;
  const int go0 = c0l * (int)ldM + rr0, dg = q32 * (int)ldM + r32, dgw = (int)ldM - nv;
  const int64_t ntiles = (B + 31) / 32;
  for (int64_t tile = (int64_t)blockIdx.x * nw + warp; tile < ntiles; tile += (int64_t)gridDim.x * nw)
  {
    const int64_t c0 = tile * 32;
    const int nc = (int)((B - c0) < 32 ? (B - c0) : 32);
    const int64_t cfg = c0 + (lane < nc ? lane : nc - 1); // idle lanes shadow the last configuration
    const T * __restrict__ qc = q + cfg * ldq;
    T * __restrict__ gtile = Mout + c0 * ldM;
    const int total = nc * nv;
    SE3<T> X; // oMi of the joint visited last
    for (int i = 1; i < nj; ++i)
    {
      // ---- forward step of joint i ---------------------------------------------------------------
      {
        const JointRec r = m.j[i];
        const SE3<T> Xl = tree_liMi(m, i, r.type, qc + r.idx_q);
        if (r.parent > 0)
        {
          if (r.parent != i - 1) X = get_se3<T>(st, L.oX + 12 * m.j[r.parent].bslot);
          X = X * Xl;
        }
        else
          X = Xl;
        if (r.bslot >= 0) put_se3(st, L.oX + 12 * r.bslot, X);
        for (int k = 0; k < r.nvj; ++k) put_motion(st, L.oJ + 6 * (r.pdof + k), act_S_col(X, r.type, k));
        put_inertia(st, L.oY + 10 * (r.depth - 1), act(X, tree_inertia(m, i)));
      }
      // ---- backward steps of every joint whose subtree is now complete ---------------------------
      const int stop = m.j[i].stop;
      for (int j = i; j != stop; j = m.j[j].parent)
      {
        const JointRec r = m.j[j];
        const Inertia<T> Y = get_inertia<T>(st, L.oY + 10 * (r.depth - 1));
        const int npath = r.pdof + r.nvj; // dofs of the ancestors-or-self = rows of these columns
        for (int k = 0; k < r.nvj; ++k)
        {
          const int col = r.idx_v + k;
          const Force<T> F = Y * get_motion<T>(st, L.oJ + 6 * (r.pdof + k)); // Ag column (crba.hxx:91)
#pragma unroll 2
          for (int t = 0; t < npath; ++t)
            myrow[m.path_row[j][t]] = dot6(get_motion<T>(st, L.oJ + 6 * t), F);
          myrow[col] += m.armature[col];
          // flush column `col` of the tile: nc segments of nv contiguous elements, coalesced
          __syncwarp();
          {
            T * __restrict__ g = gtile + (int64_t)col * nv;
            int rr = rr0, so = so0, go = go0;
            for (int e = lane; e < total; e += 32)
            {
              g[go] = em[so];
              rr += r32; so += ds; go += dg;
              if (rr >= nv) { rr -= nv; so += dsw; go += dgw; }
            }
          }
          __syncwarp();
        }
        for (int k = 0; k < r.nvj; ++k) myrow[r.idx_v + k] = T(0);
        if (r.parent > 0)
        {
          Inertia<T> Yp = get_inertia<T>(st, L.oY + 10 * (r.depth - 2));
          Yp += Y;
          put_inertia(st, L.oY + 10 * (r.depth - 2), Yp);
        }
      }
    }
    __syncwarp();
  }
}

} // namespace brbd
