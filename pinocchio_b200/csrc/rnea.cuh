// rnea.cuh — batched RNEA, one configuration per thread, 32 configurations per warp tile.
//
// Restates impl::rnea (reference: include/pinocchio/algorithm/rnea.hxx:117-161) with
// RneaForwardStep (rnea.hxx:45-79) and RneaBackwardStep (rnea.hxx:92-107) specialised per joint
// type tag.  Batch driver semantics: rneaInParallel (algorithm/parallel/rnea.hpp:38-83).
//
// Per-thread state: (v, a_gf) are only needed by the children of a joint, so they live in a
// depth-indexed stack (valid because joints are numbered depth-first, CRBAChecker crba.hxx:573-595);
// (liMi, f) per joint go to thread-local memory (coalesced, L1/L2-resident) until the backward sweep.
#pragma once

#include "engine.cuh"

namespace brbd
{

template<class T> struct LocalSE3 { T d[12]; };
template<class T> BRBD_DI void store_se3(T * d, const SE3<T> & X)
{
  d[0] = X.R.c0.x; d[1] = X.R.c0.y; d[2] = X.R.c0.z; d[3] = X.R.c1.x; d[4] = X.R.c1.y; d[5] = X.R.c1.z;
  d[6] = X.R.c2.x; d[7] = X.R.c2.y; d[8] = X.R.c2.z; d[9] = X.p.x; d[10] = X.p.y; d[11] = X.p.z;
}
template<class T> BRBD_DI SE3<T> load_se3(const T * d)
{
  SE3<T> X;
  X.R.c0 = Vec3<T>(d[0], d[1], d[2]); X.R.c1 = Vec3<T>(d[3], d[4], d[5]); X.R.c2 = Vec3<T>(d[6], d[7], d[8]);
  X.p = Vec3<T>(d[9], d[10], d[11]);
  return X;
}
template<class T> BRBD_DI void store6(T * d, const Motion<T> & m) { d[0] = m.lin.x; d[1] = m.lin.y; d[2] = m.lin.z; d[3] = m.ang.x; d[4] = m.ang.y; d[5] = m.ang.z; }
template<class T> BRBD_DI void store6(T * d, const Force<T> & m) { d[0] = m.lin.x; d[1] = m.lin.y; d[2] = m.lin.z; d[3] = m.ang.x; d[4] = m.ang.y; d[5] = m.ang.z; }
template<class T> BRBD_DI Motion<T> load_motion(const T * d) { Motion<T> m; m.lin = Vec3<T>(d[0], d[1], d[2]); m.ang = Vec3<T>(d[3], d[4], d[5]); return m; }
template<class T> BRBD_DI Force<T> load_force(const T * d) { Force<T> m; m.lin = Vec3<T>(d[0], d[1], d[2]); m.ang = Vec3<T>(d[3], d[4], d[5]); return m; }

// One configuration. q/v/a point at this thread's staged rows; tau overwrites the `a` row.
template<class T> BRBD_DI void rnea_thread(const ModelPOD<T> & m, const T * q, const T * v, T * a_tau)
{
  T liMi_s[MAXJ][12];
  T f_s[MAXJ][6];
  T v_d[MAXDEPTH][6], a_d[MAXDEPTH][6];

  // data.v[0] = 0; data.a_gf[0] = -gravity (rnea.hxx:137-138)
  {
    Motion<T> z = mzero<T>();
    store6(v_d[0], z);
    z.lin = Vec3<T>(-m.gravity[0], -m.gravity[1], -m.gravity[2]);
    store6(a_d[0], z);
  }
  const int nj = m.njoints;
  for (int i = 1; i < nj; ++i)
  {
    const int type = m.type[i], parent = m.parent[i], iq = m.idx_q[i], iv = m.idx_v[i], d = m.depth[i];
    const SE3<T> X = joint_liMi(m, i, type, q + iq);
    store_se3(liMi_s[i], X);
    Motion<T> vi = joint_velocity(type, v + iv);
    if (parent > 0) vi += X.actInv(load_motion(v_d[d - 1]));
    Motion<T> ai = cross_joint_velocity(vi, type, v + iv); // jdata.c() == 0
    // += S * a_J, then overwrite the consumed slots with armature * a (rnea.hxx:68, :158)
    const int nvj = m.nvj[i];
    for (int k = 0; k < nvj; ++k)
    {
      const int row = joint_S_row(type, k);
      const T ak = a_tau[iv + k];
      if (row < 3) ai.lin.set(row, ai.lin.get(row) + ak); else ai.ang.set(row - 3, ai.ang.get(row - 3) + ak);
      a_tau[iv + k] = m.armature[iv + k] * ak;
    }
    ai += X.actInv(load_motion(a_d[d - 1]));
    store6(v_d[d], vi);
    store6(a_d[d], ai);
    const Inertia<T> Y = model_inertia(m, i);
    const Force<T> h = Y * vi;
    Force<T> f = Y * ai;
    f += fcross(vi, h);
    store6(f_s[i], f);
  }
  for (int i = nj - 1; i > 0; --i)
  {
    const int type = m.type[i], parent = m.parent[i], iv = m.idx_v[i];
    const Force<T> f = load_force(f_s[i]);
    const int nvj = m.nvj[i];
    for (int k = 0; k < nvj; ++k) a_tau[iv + k] += get6(f, joint_S_row(type, k));
    if (parent > 0)
    {
      const SE3<T> X = load_se3(liMi_s[i]);
      Force<T> fp = load_force(f_s[parent]);
      fp += X.act(f);
      store6(f_s[parent], fp);
    }
  }
}

template<class T>
__global__ void __launch_bounds__(512)
rnea_kernel(const ModelPOD<T> * __restrict__ gm, const T * __restrict__ q, int64_t ldq, const T * __restrict__ v,
            int64_t ldv, const T * __restrict__ a, int64_t lda, T * __restrict__ tau, int64_t ldtau, int64_t B)
{
  __shared__ ModelPOD<T> m;
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  copy_model_to_smem(&m, gm);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int qpad = m.nq | 1, vpad = m.nv | 1;
  T * sq = reinterpret_cast<T *>(dyn_smem) + (size_t)warp * 32 * (qpad + 2 * vpad);
  T * sv = sq + 32 * qpad;
  T * sa = sv + 32 * vpad;
  const int64_t ntiles = (B + 31) / 32;
  for (int64_t tile = (int64_t)blockIdx.x * nw + warp; tile < ntiles; tile += (int64_t)gridDim.x * nw)
  {
    const int64_t c0 = tile * 32;
    const int nc = (int)((B - c0) < 32 ? (B - c0) : 32);
    tile_load(sq, qpad, q + c0 * ldq, ldq, m.nq, nc, lane);
    tile_load(sv, vpad, v + c0 * ldv, ldv, m.nv, nc, lane);
    tile_load(sa, vpad, a + c0 * lda, lda, m.nv, nc, lane);
    BRBD_SYNCWARP();
    if (lane < nc) rnea_thread(m, sq + lane * qpad, sv + lane * vpad, sa + lane * vpad);
    BRBD_SYNCWARP();
    tile_store(tau + c0 * ldtau, ldtau, sa, vpad, m.nv, nc, lane);
    BRBD_SYNCWARP();
  }
}

} // namespace brbd
