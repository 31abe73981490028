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

#include <cuda.h> // CUtensorMap (type only; the encoder is fetched through cudaGetDriverEntryPoint, capi.cu)

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
  // gofs[e] = global offset of element e of a (32 configurations x nv) column block: e + (e / nv) * (ldM - nv)
  int * gofs = reinterpret_cast<int *>(sm + (size_t)L.nstate * NT + (size_t)nw * 32 * L.epad);
  const int nj = m.njoints, nv = m.nv;
  const int dgc = (int)ldM - nv;
  for (int e = tid; e < 32 * nv; e += NT) gofs[e] = e + (e / nv) * dgc;
  // The emitter rows hold zeros outside the entries of the column being assembled: a joint zeroes its own
  // rows once its columns are flushed (it is never again an ancestor of a column of this tile).
  for (int k = lane; k < 32 * L.epad; k += 32) em[k] = T(0);
  __syncthreads();
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
            for (int e = lane; e < total; e += 32) g[gofs[e]] = em[e];
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
constexpr int CRBA_PF = 4; // L2 prefetch distance of q, in joints 
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
  int * gofs = reinterpret_cast<int *>(sm + (size_t)L.nstate * NT + (size_t)nw * 32 * L.epad);
  const int nj = m.njoints, nv = m.nv;
  const int dgc = (int)ldM - nv;
  for (int e = tid; e < 32 * nv; e += NT) gofs[e] = e + (e / nv) * dgc;
  for (int k = lane; k < 32 * L.epad; k += 32) em[k] = T(0);
  const uint32_t tbase = tmem_alloc_cta(L.tcols, &tmem_base_slot); // includes __syncthreads()
  // this warp's slice: lanes of its SM sub-partition, columns after those of warp - 4 (if any)
  const TmemSlots<T> tm{tbase + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)((warp >> 2) * L.tvals * (int)(sizeof(T) / 4))};
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
    // ... and its cache line is asked into L2 CRBA_PF joints ahead (running into the next tile of this warp): under the
    // kernel's write stream a DRAM read takes longer than one joint step
    const int64_t tnext = tile + (int64_t)gridDim.x * nw;
    const int64_t cnext = tnext * 32 + lane;
    const T * __restrict__ qn = tnext < ntiles ? q + (cnext < B ? cnext : B - 1) * ldq : qc;
    for (int i = 1; i < nj; ++i)
    {
      {
        const JointRec r = m.j[i];
        const T q0 = qnext;
        if (i + 1 < nj) qnext = __ldg(qc + m.j[i + 1].idx_q);
        {
          const int ip = i + CRBA_PF;
          const T * pa = ip < nj ? qc + m.j[ip].idx_q : qn + m.j[ip - nj + 1 < nj ? ip - nj + 1 : nj - 1].idx_q;
          asm volatile("prefetch.global.L2 [%0];" ::"l"(pa));
        }
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
            for (int e = lane; e < total; e += 32) g[gofs[e]] = em[e];
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

// ------------------------------------------------------------------------------------------------------
// crba_tma_kernel: crba_tmem_kernel with the column blocks leaving the SM through TMA tensor stores.
//
// Measured on the kernel above (65 536 x simple_humanoid, temporary variants): 0.245 ms, of which
// 0.11 ms are the column flush; the LSU stores do not overlap with the other warps' arithmetic whatever the unrolling, the
// address table, the store width or the phase of the warps, and one cp.async.bulk per lane and column (256..304 bytes
// each) is bound by the copy engine's issue rate.  Here the warp's (32 configurations x nv) column block is ONE (or two)
// cp.async.bulk.tensor.2d stores issued by one lane: asynchronous, a few KB per instruction.
//
// The caller's matrix is seen as a 2-D tensor, inner dimension = one configuration's nv x nv matrix (col-major, so column
// `col` is the inner range [col nv, col nv + nv)), outer dimension = configurations with stride ldM.  A tensor store needs
// 16-byte aligned box starts (a box starting at an odd FP64 element raises "illegal instruction", scripts/tma_probe.py), a
// box whose inner extent is a multiple of 16 bytes and 16-byte strides.
//   * even nv, even ldM (talos 38, humanoid_random 32, humanoid 34): box = the column block itself (ODD = false).
//   * odd nv (FP64; simple_humanoid 35) — every other column segment starts 8 bytes off a 16-byte boundary (ODD = true):
//     the box is nv + 1 wide and starts EARLY, inside the previous column — one element early where the segment starts at
//     an odd element (the extra element is M[nv-1, col-1]), two where it starts at an even one (M[nv-2, col-1], M[nv-1, col-1];
//     the column's own last entry M[nv-1, col] is then left to the box of column col + 1, whose parity is the opposite).
//     All of these are entries of the strictly lower triangle, i.e. structural zeros of the result (crba.hpp:15-22), so the
//     extra elements always store the value that belongs there.  Column 0 (nothing in front of it) and column nv - 1 where
//     it would need the two-early box (M[nv-2, nv-2] is not a zero) are written by plain stores — 1.5 columns per tile.
//     With an odd ldM the parity also alternates with the configuration: two maps over PAIRS of configurations
//     (stride 2 ldM), the odd ones based at Mout + ldM - 1 with every inner coordinate shifted by one; the tile keeps the
//     even configurations in rows 0..15 and the odd ones in rows 16..31 and is stored in two halves.
// The emitter tile is rewritten only after cp.async.bulk.wait_group.read, lazily — just before the next column is
// assembled — so the wait hides behind the forward step / the Y J product in between.
// ------------------------------------------------------------------------------------------------------
// struct CrbaTmaGeom {bx, odd, pairs}: tree.cuh (shared with the generated CRBA kernel, codegen.cu)
BRBD_DI void tma_store_2d(const void * tmap, const void * ssrc, int x, int y)
{
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tmap),
               "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "r"(x), "r"(y)
               : "memory");
}
BRBD_DI void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
BRBD_DI void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
BRBD_DI void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
BRBD_DI void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template<class T, int NT, bool ODD>
__global__ void __launch_bounds__(NT, 1)
crba_tma_kernel(const __grid_constant__ TreePOD<T> m, const CrbaTmemLayout L, const CrbaTmaGeom G,
                const __grid_constant__ CUtensorMap map0, const __grid_constant__ CUtensorMap map1,
                const T * __restrict__ q, int64_t ldq, T * __restrict__ Mout, int64_t ldM, int64_t B)
{
  extern __shared__ __align__(128) unsigned char dyn_smem128[];
  __shared__ uint32_t tmem_base_slot;
  T * sm = reinterpret_cast<T *>(dyn_smem128);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int nw = NT / 32;
  // [tiles of the warps (128-byte aligned, TMA source) | J slots]
  T * em = sm + (size_t)warp * 32 * L.epad;
  const Slots<T, NT> st{sm + (size_t)nw * 32 * L.epad + tid};
  // ODD: with two maps the tile keeps the even configurations in rows 0..15 and the odd ones in rows 16..31
  const int half = (ODD && G.pairs) ? (lane & 1) : 0;
  T * myrow = em + ((ODD && G.pairs) ? (lane >> 1) + 16 * half : lane) * L.epad;
  T * row = myrow; // ODD: myrow + shift of the column being assembled
  const int nj = m.njoints, nv = m.nv;
  for (int k = 0; k < L.epad; ++k) myrow[k] = T(0);
  const uint32_t tbase = tmem_alloc_cta(L.tcols, &tmem_base_slot); // includes __syncthreads()
  const TmemSlots<T> tm{tbase + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)((warp >> 2) * L.tvals * (int)(sizeof(T) / 4))};
  const bool ffroot = m.ffroot != 0;
  const int joff = ffroot ? CRBA_FF_SAVED : 0;
  const int64_t ntiles = (B + 31) / 32;
  int pj = 0;            // joint of the column in flight (0: none)
  int ps = 0;            // ODD: shift of its entries in this lane's row
  bool inflight = false; // ODD: a tensor store of this warp may still be reading the tile
  for (int64_t tile = (int64_t)blockIdx.x * nw + warp; tile < ntiles; tile += (int64_t)gridDim.x * nw)
  {
    const int64_t c0 = tile * 32;
    const bool live = c0 + lane < B;
    const int64_t cfg = live ? c0 + lane : B - 1; // idle lanes shadow the last configuration
    const T * __restrict__ qc = q + cfg * ldq;
    SE3<T> X;
    Inertia<T> Yown;
    T qnext = __ldg(qc + m.j[1].idx_q);
    const int64_t tnext = tile + (int64_t)gridDim.x * nw;
    const int64_t cnext = tnext * 32 + lane;
    const T * __restrict__ qn = tnext < ntiles ? q + (cnext < B ? cnext : B - 1) * ldq : qc;
    for (int i = 1; i < nj; ++i)
    {
      {
        const JointRec r = m.j[i];
        const T q0 = qnext;
        if (i + 1 < nj) qnext = __ldg(qc + m.j[i + 1].idx_q);
        {
          const int ip = i + CRBA_PF;
          const T * pa = ip < nj ? qc + m.j[ip].idx_q : qn + m.j[ip - nj + 1 < nj ? ip - nj + 1 : nj - 1].idx_q;
          asm volatile("prefetch.global.L2 [%0];" ::"l"(pa));
        }
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
      Inertia<T> Y = Yown;
      for (int j = i; j != stop; j = m.j[j].parent)
      {
        const JointRec r = m.j[j];
        const int npath = r.pdof + r.nvj;
        for (int k = 0; k < r.nvj; ++k)
        {
          const int col = r.idx_v + k;
          Force<T> F;
          SE3<T> Xr;
          if (ffroot)
          {
            Xr = get_se3<T>(st, L.oJ);
            F = Y * (j == 1 ? act_S_col(Xr, J_FF, k) : get_motion<T>(st, L.oJ + 6 * (r.pdof + k) - joff));
          }
          else
            F = Y * get_motion<T>(st, L.oJ + 6 * (r.pdof + k));
          int sh = 0;          // ODD: shift of this lane's entries; plain: this lane stores the column itself
          bool plain = false;
          if constexpr (ODD)
          {
            // parity of the segment start (elements): nv is odd, an odd ldM moves the odd configurations by one more
            const int a = (half + col) & 1;
            plain = col == 0 || (col == nv - 1 && a == 0);
            sh = plain ? 0 : (a ? 1 : 2);
            if (inflight)
            {
              if (lane == 0) bulk_wait_read();
              __syncwarp();
            }
            if (pj)
            { // the shift changes from column to column: clear what the previous column left
              const int np = m.j[pj].pdof + m.j[pj].nvj;
              for (int t = 0; t < np; ++t) myrow[ps + m.path_row[pj][t]] = T(0);
            }
            row = myrow + sh;
          }
          else if (pj)
          { // the tile is free again once the engine has read the previous column block out of it
            if (lane == 0) bulk_wait_read();
            __syncwarp();
            if (pj != j) // the rows of a finished joint are never again on the path of a column of this tile
              for (int kk = 0; kk < m.j[pj].nvj; ++kk) row[m.j[pj].idx_v + kk] = T(0);
          }
          int t0 = 0;
          if (ffroot)
          {
            // rows of the free-flyer root: J_root^T F = oMi_root.actInv(F)  (S = identity, joint-free-flyer.hpp:47-57)
            const Vec3<T> fl = tmul(Xr.R, F.lin), fa = tmul(Xr.R, F.ang - cross(Xr.p, F.lin));
            row[0] = fl.x; row[1] = fl.y; row[2] = fl.z; row[3] = fa.x; row[4] = fa.y; row[5] = fa.z;
            t0 = 6;
          }
#pragma unroll 4
          for (int t = t0; t < npath; ++t)
            row[m.path_row[j][t]] = dot6(get_motion<T>(st, L.oJ + 6 * t - joff), F);
          row[col] += m.armature[col];
          fence_async_smem();
          __syncwarp();
          if constexpr (ODD)
          {
            // per half: box one element early (segment starts at an odd element) or two (even), see the header comment
            const int a0 = col & 1, a1 = (col + 1) & 1;
            const bool plain0 = col == 0 || (col == nv - 1 && a0 == 0), plain1 = col == 0 || (col == nv - 1 && a1 == 0);
            if (lane == 0)
            {
              if (G.pairs)
              {
                const int y = (int)(c0 >> 1);
                if (!plain0) tma_store_2d(&map0, em, col * nv - (a0 ? 1 : 2), y);
                if (!plain1 && c0 + 1 < B) tma_store_2d(&map1, em + 16 * L.epad, col * nv - (a1 ? 1 : 2) + 1, y);
              }
              else if (!plain0)
                tma_store_2d(&map0, em, col * nv - (a0 ? 1 : 2), (int)c0);
              bulk_commit();
            }
            inflight = G.pairs ? !(plain0 && plain1) : !plain0;
            if (plain && live)
            {
              T * __restrict__ g = Mout + cfg * ldM + (int64_t)col * nv;
              for (int e = 0; e < nv; ++e) g[e] = myrow[e];
            }
            ps = sh;
          }
          else if (lane == 0)
          {
            tma_store_2d(&map0, em, col * nv, (int)c0); // rows past the batch are clipped by the tensor map
            bulk_commit();
          }
          pj = j;
        }
        if (r.parent > 0)
        {
          T y[10];
          tmem_wait_st();
          tm.template load<10>(L.tY + 10 * (r.depth - 2), y);
          Inertia<T> Yp;
          Yp.m = y[0]; Yp.c = Vec3<T>(y[1], y[2], y[3]);
          Yp.I.xx = y[4]; Yp.I.xy = y[5]; Yp.I.yy = y[6]; Yp.I.xz = y[7]; Yp.I.yz = y[8]; Yp.I.zz = y[9];
          Yp += Y;
          Y = Yp;
          if (r.parent == stop)
          {
            const T z[10] = {Yp.m, Yp.c.x, Yp.c.y, Yp.c.z, Yp.I.xx, Yp.I.xy, Yp.I.yy, Yp.I.xz, Yp.I.yz, Yp.I.zz};
            tm.template store<10>(L.tY + 10 * (r.depth - 2), z);
          }
        }
      }
    }
  }
  if (lane == 0) bulk_wait_all();
  __syncwarp();
  tmem_wait_st();
  tmem_free_cta(tbase, L.tcols);
}

} // namespace brbd
