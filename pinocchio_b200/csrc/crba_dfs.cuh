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

#include "tmem.cuh"
#include "tree.cuh"

namespace brbd
{

struct CrbaLayout
{
  int oJ, oY, oX, nstate; // slot offsets of the per-thread state, total slots
  int epad;               // emitter row length (= nv)
};
inline CrbaLayout crba_layout(int maxpathdof, int maxdepth, int nbranch, int nv)
{
  CrbaLayout L;
  L.oJ = 0;
  L.oY = L.oJ + 6 * maxpathdof;
  L.oX = L.oY + 10 * maxdepth;
  L.nstate = L.oX + 12 * (nbranch > 0 ? nbranch : 1);
  L.epad = nv; // unpadded: element e of the (32 x nv) column block sits at em[e]
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
  T * em = sm + (size_t)L.nstate * NT + (size_t)warp * 32 * L.epad; // [32][nv], row = configuration of the tile
  T * myrow = em + lane * L.epad;
  // ctab[t * 32 + lane] = configuration (row of em) of element e = lane + 32 t of a column block
  unsigned char * ctab = reinterpret_cast<unsigned char *>(sm + (size_t)L.nstate * NT + (size_t)nw * 32 * L.epad);
  const int nj = m.njoints, nv = m.nv;
  for (int e = tid; e < 32 * nv; e += NT) ctab[e] = (unsigned char)(e / nv);
  // The emitter rows hold zeros outside the entries of the column being assembled: a joint zeroes its own
  // rows once its columns are flushed (it is never again an ancestor of a column of this tile).
  for (int k = lane; k < 32 * L.epad; k += 32) em[k] = T(0);
  __syncthreads();
  const int dgc = (int)ldM - nv; // global offset of element e of a column block: e + c * (ldM - nv)
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
#pragma unroll 4
          for (int t = 0; t < npath; ++t)
            myrow[m.path_row[j][t]] = dot6(get_motion<T>(st, L.oJ + 6 * t), F);
          myrow[col] += m.armature[col];
          // flush column `col` of the tile: nc segments of nv contiguous elements, coalesced
          __syncwarp();
          {
            T * __restrict__ g = gtile + (int64_t)col * nv;
#pragma unroll 5
            for (int e = lane; e < total; e += 32) g[e + (int)ctab[e] * dgc] = em[e];
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


// ------------------------------------------------------------------------------------------------------
// Same algorithm with the rarely-touched part of the state — oYcrb per depth and oMi of the open branching
// joints — in tensor memory (tmem.cuh), which leaves only the J columns of the root path and the emitter
// row in shared memory and roughly doubles the number of resident configurations per SM.
// A leaf's own inertia never leaves the registers (its backward step follows its forward step at once),
// so the TMEM stack holds depths 1 .. maxdepth-1 only.
// ------------------------------------------------------------------------------------------------------
struct CrbaTmemLayout
{
  int oJ, nstate;   // shared memory: J (6 x maxpathdof)
  int epad;
  int tY, tX, tvals; // TMEM value offsets: Y (10 x (maxdepth - 1)), oMi (12 x nbranch); values per warp slice
  int tcols;         // columns allocated by the CTA (power of two)
};
// With a free-flyer root (TreePOD::ffroot) the 6 J columns of the root are not stored: its oMi (12 values) takes their
// place, the columns of the other path dofs move down by CRBA_FF_SAVED slots, and the six root rows of every column are
// oMi_root.actInv(F) — the force in the root frame — instead of six 6-D dot products (J_root = oMi_root's action matrix).
constexpr int CRBA_FF_SAVED = 24;
template<class T> inline CrbaTmemLayout crba_tmem_layout(int maxpathdof, int maxdepth, int nbranch, int nv, int warps, int ffroot)
{
  CrbaTmemLayout L;
  L.oJ = 0;
  L.nstate = 6 * maxpathdof - (ffroot ? CRBA_FF_SAVED : 0);
  L.epad = nv;
  L.tY = 0;
  L.tX = 10 * (maxdepth > 1 ? maxdepth - 1 : 1);
  L.tvals = L.tX + 12 * (nbranch > 0 ? nbranch : 1);
  const int cols_per_slice = L.tvals * (int)(sizeof(T) / 4);
  L.tcols = tmem_round_cols(cols_per_slice * ((warps + 3) / 4));
  return L;
}

template<class T, int NT>
__global__ void __launch_bounds__(NT, 1)
crba_tmem_kernel(const __grid_constant__ TreePOD<T> m, const CrbaTmemLayout L, const T * __restrict__ q, int64_t ldq,
                 T * __restrict__ Mout, int64_t ldM, int64_t B)
{
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  __shared__ uint32_t tmem_base_slot;
  T * sm = reinterpret_cast<T *>(dyn_smem);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int nw = NT / 32;
  const Slots<T, NT> st{sm + tid};
  T * em = sm + (size_t)L.nstate * NT + (size_t)warp * 32 * L.epad;
  T * myrow = em + lane * L.epad;
  unsigned char * ctab = reinterpret_cast<unsigned char *>(sm + (size_t)L.nstate * NT + (size_t)nw * 32 * L.epad);
  const int nj = m.njoints, nv = m.nv;
  for (int e = tid; e < 32 * nv; e += NT) ctab[e] = (unsigned char)(e / nv);
  for (int k = lane; k < 32 * L.epad; k += 32) em[k] = T(0);
  const uint32_t tbase = tmem_alloc_cta(L.tcols, &tmem_base_slot); // includes __syncthreads()
  // this warp's slice: lanes of its SM sub-partition, columns after those of warp - 4 (if any)
  const TmemSlots<T> tm{tbase + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)((warp >> 2) * L.tvals * (int)(sizeof(T) / 4))};
  const int dgc = (int)ldM - nv;
  const bool ffroot = m.ffroot != 0;
  const int joff = ffroot ? CRBA_FF_SAVED : 0;
  const int64_t ntiles = (B + 31) / 32;
  for (int64_t tile = (int64_t)blockIdx.x * nw + warp; tile < ntiles; tile += (int64_t)gridDim.x * nw)
  {
    const int64_t c0 = tile * 32;
    const int nc = (int)((B - c0) < 32 ? (B - c0) : 32);
    const int64_t cfg = c0 + (lane < nc ? lane : nc - 1);
    const T * __restrict__ qc = q + cfg * ldq;
    T * __restrict__ gtile = Mout + c0 * ldM;
    const int total = nc * nv;
    SE3<T> X;
    Inertia<T> Yown;
    T qnext = __ldg(qc + m.j[1].idx_q); // first coordinate of the next joint, fetched one joint ahead
    for (int i = 1; i < nj; ++i)
    {
      {
        const JointRec r = m.j[i];
        const T q0 = qnext;
        if (i + 1 < nj) qnext = __ldg(qc + m.j[i + 1].idx_q);
        const SE3<T> Xl = tree_liMi(m, i, r.type, qc + r.idx_q, q0);
        if (r.parent > 0)
        {
          if (r.parent != i - 1)
          {
            T x[12];
            tmem_wait_st();
            tm.template load<12>(L.tX + 12 * m.j[r.parent].bslot, x);
            X.R.c0 = Vec3<T>(x[0], x[1], x[2]); X.R.c1 = Vec3<T>(x[3], x[4], x[5]); X.R.c2 = Vec3<T>(x[6], x[7], x[8]);
            X.p = Vec3<T>(x[9], x[10], x[11]);
          }
          X = X * Xl;
        }
        else
          X = Xl;
        if (r.bslot >= 0)
        {
          const T x[12] = {X.R.c0.x, X.R.c0.y, X.R.c0.z, X.R.c1.x, X.R.c1.y, X.R.c1.z, X.R.c2.x, X.R.c2.y, X.R.c2.z, X.p.x, X.p.y, X.p.z};
          tm.template store<12>(L.tX + 12 * r.bslot, x);
        }
        if (ffroot && i == 1) put_se3(st, L.oJ, X);
        else
          for (int k = 0; k < r.nvj; ++k) put_motion(st, L.oJ + 6 * (r.pdof + k) - joff, act_S_col(X, r.type, k));
        Yown = act(X, tree_inertia(m, i));
        if (r.nchild > 0)
        {
          const T y[10] = {Yown.m, Yown.c.x, Yown.c.y, Yown.c.z, Yown.I.xx, Yown.I.xy, Yown.I.yy, Yown.I.xz, Yown.I.yz, Yown.I.zz};
          tm.template store<10>(L.tY + 10 * (r.depth - 1), y);
        }
      }
      const int stop = m.j[i].stop;
      Inertia<T> Y = Yown; // the first joint of the unwind is the leaf just visited
      for (int j = i; j != stop; j = m.j[j].parent)
      {
        const JointRec r = m.j[j];
        const int npath = r.pdof + r.nvj;
        for (int k = 0; k < r.nvj; ++k)
        {
          const int col = r.idx_v + k;
          Force<T> F;
          int t0 = 0;
          if (ffroot)
          {
            const SE3<T> Xr = get_se3<T>(st, L.oJ);
            F = Y * (j == 1 ? act_S_col(Xr, J_FF, k) : get_motion<T>(st, L.oJ + 6 * (r.pdof + k) - joff));
            // rows of the free-flyer root: J_root^T F = oMi_root.actInv(F)  (S = identity, joint-free-flyer.hpp:47-57)
            const Vec3<T> fl = tmul(Xr.R, F.lin), fa = tmul(Xr.R, F.ang - cross(Xr.p, F.lin));
            myrow[0] = fl.x; myrow[1] = fl.y; myrow[2] = fl.z; myrow[3] = fa.x; myrow[4] = fa.y; myrow[5] = fa.z;
            t0 = 6;
          }
          else
            F = Y * get_motion<T>(st, L.oJ + 6 * (r.pdof + k));
#pragma unroll 4
          for (int t = t0; t < npath; ++t)
            myrow[m.path_row[j][t]] = dot6(get_motion<T>(st, L.oJ + 6 * t - joff), F);
          myrow[col] += m.armature[col];
          __syncwarp();
          {
            T * __restrict__ g = gtile + (int64_t)col * nv;
#pragma unroll 5
            for (int e = lane; e < total; e += 32) g[e + (int)ctab[e] * dgc] = em[e];
          }
          __syncwarp();
        }
        for (int k = 0; k < r.nvj; ++k) myrow[r.idx_v + k] = T(0);
        if (r.parent > 0)
        {
          // oYcrb[parent] += oYcrb[j]; the sum is the parent's Y if the parent is next in this unwind
          T y[10];
          tmem_wait_st();
          tm.template load<10>(L.tY + 10 * (r.depth - 2), y);
          Inertia<T> Yp;
          Yp.m = y[0]; Yp.c = Vec3<T>(y[1], y[2], y[3]);
          Yp.I.xx = y[4]; Yp.I.xy = y[5]; Yp.I.yy = y[6]; Yp.I.xz = y[7]; Yp.I.yz = y[8]; Yp.I.zz = y[9];
          Yp += Y;
          Y = Yp;
          if (r.parent != stop)
            ; // consumed by the next iteration straight from registers
          else
          {
            const T z[10] = {Yp.m, Yp.c.x, Yp.c.y, Yp.c.z, Yp.I.xx, Yp.I.xy, Yp.I.yy, Yp.I.xz, Yp.I.yz, Yp.I.zz};
            tm.template store<10>(L.tY + 10 * (r.depth - 2), z);
          }
        }
      }
    }
    __syncwarp();
  }
  tmem_wait_st();
  tmem_free_cta(tbase, L.tcols);
}

} // namespace brbd
