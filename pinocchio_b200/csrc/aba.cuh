// aba.cuh — batched ABA in the WORLD convention (what abaInParallel evaluates,
// reference: include/pinocchio/algorithm/parallel/aba.hpp:82), one configuration per thread.
//
// Restates impl::abaWorldConvention (algorithm/aba.hxx:242-293):
//   pass 1  AbaWorldConventionForwardStep1  (aba.hxx:101-138)
//   pass 2  AbaWorldConventionBackwardStep  (aba.hxx:152-192)
//   pass 3  AbaWorldConventionForwardStep2  (aba.hxx:206-226)
// The "consistent output" tail (aba.hxx:228-230, 286-290: data.oa, data.of) is not part of what the
// batched entry point returns and is skipped.
// The articulated inertia oYaba is kept as a packed symmetric 6x6 (21 numbers) instead of the
// reference's dense Matrix6; U = Ia*J, StU = J^T U + armature, Dinv = StU^-1 (Cholesky for
// multi-dof joints as PerformStYSInversion, joint-common-operations.hpp:23-33; plain reciprocal for
// 1-dof joints).
#pragma once

#include "engine.cuh"
#include "rnea.cuh"

namespace brbd
{

// ---- packed symmetric 6x6 helpers (index (r,c), r<=c : r*6 - r(r-1)/2 + c - r) -------------------
template<class T> BRBD_DI T sym6_get(const T * a, int r, int c)
{
  return r <= c ? a[r * 6 - (r * (r - 1)) / 2 + (c - r)] : a[c * 6 - (c * (c - 1)) / 2 + (r - c)];
}
template<class T> BRBD_DI void sym6_mul(const T * a, const T * x, T * y)
{
#pragma unroll
  for (int r = 0; r < 6; ++r)
  {
    T acc = sym6_get(a, r, 0) * x[0];
#pragma unroll
    for (int c = 1; c < 6; ++c) acc += sym6_get(a, r, c) * x[c];
    y[r] = acc;
  }
}
// Inertia::matrix() — inertia.hpp:480-491, packed
template<class T> BRBD_DI void inertia_to_sym6(const Inertia<T> & Y, T * a)
{
  const T m = Y.m, cx = Y.c.x, cy = Y.c.y, cz = Y.c.z;
  // LL = m 1
  a[0] = m;  a[1] = T(0); a[2] = T(0);
  a[6] = m;  a[7] = T(0);
  a[11] = m;
  // LA = -m [c]x : rows linear, cols angular. [c]x = [[0,-cz,cy],[cz,0,-cx],[-cy,cx,0]]
  const T mx = m * cx, my = m * cy, mz = m * cz;
  a[3] = T(0); a[4] = mz;   a[5] = -my;   // row 0, cols 3..5
  a[8] = -mz;  a[9] = T(0); a[10] = mx;   // row 1
  a[12] = my;  a[13] = -mx; a[14] = T(0); // row 2
  // AA = I_c - m [c]x^2 (AlphaSkewSquare, symmetric3.hpp:259-271)
  a[15] = Y.I.xx + m * (cy * cy + cz * cz);
  a[16] = Y.I.xy - m * cx * cy;
  a[17] = Y.I.xz - m * cx * cz;
  a[18] = Y.I.yy + m * (cx * cx + cz * cz);
  a[19] = Y.I.yz - m * cy * cz;
  a[20] = Y.I.zz + m * (cx * cx + cy * cy);
}

template<class T> BRBD_DI void m2a(const Motion<T> & m, T * x) { x[0] = m.lin.x; x[1] = m.lin.y; x[2] = m.lin.z; x[3] = m.ang.x; x[4] = m.ang.y; x[5] = m.ang.z; }
template<class T> BRBD_DI void f2a(const Force<T> & m, T * x) { x[0] = m.lin.x; x[1] = m.lin.y; x[2] = m.lin.z; x[3] = m.ang.x; x[4] = m.ang.y; x[5] = m.ang.z; }

// In-place inverse of a small SPD matrix through its Cholesky factor (n <= 6), row-major A[6][6].
template<class T> BRBD_DI void llt_inverse(int n, T A[6][6], T Ainv[6][6])
{
  T L[6][6];
  for (int i = 0; i < n; ++i)
    for (int j = 0; j <= i; ++j)
    {
      T s = A[i][j];
      for (int k = 0; k < j; ++k) s -= L[i][k] * L[j][k];
      if (i == j) L[i][i] = sqrt_t(s);
      else L[i][j] = s / L[j][j];
    }
  for (int c = 0; c < n; ++c)
  {
    T y[6];
    for (int i = 0; i < n; ++i)
    {
      T s = (i == c) ? T(1) : T(0);
      for (int k = 0; k < i; ++k) s -= L[i][k] * y[k];
      y[i] = s / L[i][i];
    }
    for (int i = n - 1; i >= 0; --i)
    {
      T s = y[i];
      for (int k = i + 1; k < n; ++k) s -= L[k][i] * Ainv[k][c];
      Ainv[i][c] = s / L[i][i];
    }
  }
}

// Thread-local state of the ABA sweeps (local memory; coalesced across the warp by construction).
template<class T> struct AbaState
{
  T J[MAXNV][6];      // world-frame joint columns (data.J)
  T UD[MAXNV][6];     // UDinv columns
  T Dinv[MAXNV][6];   // row k of the joint's Dinv block
  T ab[MAXJ][6];      // oa_gf bias (pass 1) of each joint
  T Ia[MAXJ][21];     // oYaba, packed symmetric
  T f[MAXJ][6];       // of
  T oMi[MAXDEPTH][12];
  T ov[MAXDEPTH][6];
  T ag[MAXDEPTH][6];  // oa_gf along the current root path (pass 3)
};

// pass 1 for joint i. Returns nothing; fills st.J, st.ab, st.Ia, st.f and the depth stacks.
template<class T> BRBD_DI void aba_forward1(const ModelPOD<T> & m, AbaState<T> & st, int i, const T * q, const T * v)
{
  const int type = m.type[i], parent = m.parent[i], iq = m.idx_q[i], iv = m.idx_v[i], d = m.depth[i], nvj = m.nvj[i];
  SE3<T> X = joint_liMi(m, i, type, q + iq);
  if (parent > 0) X = load_se3(st.oMi[d - 1]) * X;
  store_se3(st.oMi[d], X);
  Motion<T> ov;
  if (nvj == 1)
  {
    const Motion<T> J0 = act_S_col(X, type, 0);
    store6(st.J[iv], J0);
    const T vq = v[iv];
    ov.lin = vq * J0.lin;
    ov.ang = vq * J0.ang;
  }
  else
  {
    for (int k = 0; k < nvj; ++k) store6(st.J[iv + k], act_S_col(X, type, k));
    ov = X.act(joint_velocity(type, v + iv));
  }
  Motion<T> ab = mzero<T>();
  if (parent > 0)
  {
    const Motion<T> ovp = load_motion(st.ov[d - 1]);
    ov += ovp;
    ab = mcross(ovp, ov);
  }
  store6(st.ov[d], ov);
  store6(st.ab[i], ab);
  const Inertia<T> Y = act(X, model_inertia(m, i));
  T Ia[21];
  inertia_to_sym6(Y, Ia);
#pragma unroll
  for (int k = 0; k < 21; ++k) st.Ia[i][k] = Ia[k];
  const Force<T> oh = Y * ov;
  store6(st.f[i], fcross(ov, oh));
}

// pass 2 for joint i; `u` is the thread's staged tau row (data.u), updated in place.
template<class T> BRBD_DI void aba_backward(const ModelPOD<T> & m, AbaState<T> & st, int i, T * u)
{
  const int parent = m.parent[i], iv = m.idx_v[i], nvj = m.nvj[i];
  T Ia[21];
#pragma unroll
  for (int k = 0; k < 21; ++k) Ia[k] = st.Ia[i][k];
  Force<T> fi = load_force(st.f[i]);
  if (nvj == 1)
  {
    const Motion<T> J = load_motion(st.J[iv]);
    const T ui = u[iv] - dot6(J, fi);
    u[iv] = ui;
    T Jv[6], U[6];
    m2a(J, Jv);
    sym6_mul(Ia, Jv, U);
    T D = Jv[0] * U[0];
#pragma unroll
    for (int r = 1; r < 6; ++r) D += Jv[r] * U[r];
    D += m.armature[iv];
    const T Dinv = T(1) / D;
    T UD[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) { UD[r] = U[r] * Dinv; st.UD[iv][r] = UD[r]; }
    st.Dinv[iv][0] = Dinv;
    if (parent > 0)
    {
#pragma unroll
      for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int c = r; c < 6; ++c) Ia[r * 6 - (r * (r - 1)) / 2 + (c - r)] -= UD[r] * U[c];
      T ab[6], Iab[6];
      m2a(load_motion(st.ab[i]), ab);
      sym6_mul(Ia, ab, Iab);
      fi.lin.x += Iab[0] + UD[0] * ui; fi.lin.y += Iab[1] + UD[1] * ui; fi.lin.z += Iab[2] + UD[2] * ui;
      fi.ang.x += Iab[3] + UD[3] * ui; fi.ang.y += Iab[4] + UD[4] * ui; fi.ang.z += Iab[5] + UD[5] * ui;
#pragma unroll
      for (int k = 0; k < 21; ++k) st.Ia[parent][k] += Ia[k];
      Force<T> fp = load_force(st.f[parent]);
      fp += fi;
      store6(st.f[parent], fp);
    }
    return;
  }
  // multi-dof joints (free-flyer, spherical, planar)
  T U[6][6], StU[6][6], Di[6][6], UD[6][6], uj[6];
  for (int k = 0; k < nvj; ++k)
  {
    const Motion<T> J = load_motion(st.J[iv + k]);
    uj[k] = u[iv + k] - dot6(J, fi);
    u[iv + k] = uj[k];
    T Jv[6], Uk[6];
    m2a(J, Jv);
    sym6_mul(Ia, Jv, Uk);
    for (int r = 0; r < 6; ++r) U[r][k] = Uk[r];
  }
  for (int a = 0; a < nvj; ++a)
  {
    T Jv[6];
    m2a(load_motion(st.J[iv + a]), Jv);
    for (int b = 0; b < nvj; ++b)
    {
      T acc = Jv[0] * U[0][b];
      for (int r = 1; r < 6; ++r) acc += Jv[r] * U[r][b];
      StU[a][b] = acc;
    }
    StU[a][a] += m.armature[iv + a];
  }
  llt_inverse(nvj, StU, Di);
  for (int r = 0; r < 6; ++r)
    for (int k = 0; k < nvj; ++k)
    {
      T acc = U[r][0] * Di[0][k];
      for (int c = 1; c < nvj; ++c) acc += U[r][c] * Di[c][k];
      UD[r][k] = acc;
    }
  for (int k = 0; k < nvj; ++k)
  {
    for (int r = 0; r < 6; ++r) st.UD[iv + k][r] = UD[r][k];
    for (int c = 0; c < nvj; ++c) st.Dinv[iv + k][c] = Di[k][c];
  }
  if (parent > 0)
  {
    for (int r = 0; r < 6; ++r)
      for (int c = r; c < 6; ++c)
      {
        T acc = UD[r][0] * U[c][0];
        for (int k = 1; k < nvj; ++k) acc += UD[r][k] * U[c][k];
        Ia[r * 6 - (r * (r - 1)) / 2 + (c - r)] -= acc;
      }
    T ab[6], Iab[6], fa[6];
    m2a(load_motion(st.ab[i]), ab);
    sym6_mul(Ia, ab, Iab);
    f2a(fi, fa);
    for (int r = 0; r < 6; ++r)
    {
      T acc = UD[r][0] * uj[0];
      for (int k = 1; k < nvj; ++k) acc += UD[r][k] * uj[k];
      fa[r] += Iab[r] + acc;
    }
    for (int k = 0; k < 21; ++k) st.Ia[parent][k] += Ia[k];
    Force<T> fp = load_force(st.f[parent]);
    fp.lin.x += fa[0]; fp.lin.y += fa[1]; fp.lin.z += fa[2];
    fp.ang.x += fa[3]; fp.ang.y += fa[4]; fp.ang.z += fa[5];
    store6(st.f[parent], fp);
  }
}

// pass 3 for joint i: ddq written over u[idx_v..]; returns the joint's oa_gf (also pushed on the stack)
template<class T> BRBD_DI Motion<T> aba_forward2(const ModelPOD<T> & m, AbaState<T> & st, int i, T * u)
{
  const int iv = m.idx_v[i], d = m.depth[i], nvj = m.nvj[i];
  Motion<T> ag = load_motion(st.ab[i]);
  ag += load_motion(st.ag[d - 1]);
  T agv[6];
  m2a(ag, agv);
  if (nvj == 1)
  {
    T t2 = st.UD[iv][0] * agv[0];
#pragma unroll
    for (int r = 1; r < 6; ++r) t2 += st.UD[iv][r] * agv[r];
    const T ddq = st.Dinv[iv][0] * u[iv] - t2;
    u[iv] = ddq;
    const Motion<T> J = load_motion(st.J[iv]);
    ag.lin += ddq * J.lin;
    ag.ang += ddq * J.ang;
  }
  else
  {
    T dd[6];
    for (int k = 0; k < nvj; ++k)
    {
      T t1 = st.Dinv[iv + k][0] * u[iv];
      for (int c = 1; c < nvj; ++c) t1 += st.Dinv[iv + k][c] * u[iv + c];
      T t2 = st.UD[iv + k][0] * agv[0];
      for (int r = 1; r < 6; ++r) t2 += st.UD[iv + k][r] * agv[r];
      dd[k] = t1 - t2;
    }
    for (int k = 0; k < nvj; ++k)
    {
      u[iv + k] = dd[k];
      const Motion<T> J = load_motion(st.J[iv + k]);
      ag.lin += dd[k] * J.lin;
      ag.ang += dd[k] * J.ang;
    }
  }
  store6(st.ag[d], ag);
  return ag;
}

template<class T> BRBD_DI void aba_thread(const ModelPOD<T> & m, const T * q, const T * v, T * tau_ddq)
{
  AbaState<T> st;
  const int nj = m.njoints;
  for (int i = 1; i < nj; ++i) aba_forward1(m, st, i, q, v);
  for (int i = nj - 1; i > 0; --i) aba_backward(m, st, i, tau_ddq);
  {
    Motion<T> g0 = mzero<T>();
    g0.lin = Vec3<T>(-m.gravity[0], -m.gravity[1], -m.gravity[2]); // data.oa_gf[0] = -gravity (aba.hxx:260)
    store6(st.ag[0], g0);
  }
  for (int i = 1; i < nj; ++i) aba_forward2(m, st, i, tau_ddq);
}

template<class T>
__global__ void __launch_bounds__(512)
aba_kernel(const ModelPOD<T> * __restrict__ gm, const T * __restrict__ q, int64_t ldq, const T * __restrict__ v,
           int64_t ldv, const T * __restrict__ tau, int64_t ldtau, T * __restrict__ a, int64_t lda, int64_t B)
{
  __shared__ ModelPOD<T> m;
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  copy_model_to_smem(&m, gm);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int qpad = m.nq | 1, vpad = m.nv | 1;
  T * sq = reinterpret_cast<T *>(dyn_smem) + (size_t)warp * 32 * (qpad + 2 * vpad);
  T * sv = sq + 32 * qpad;
  T * st = sv + 32 * vpad;
  const int64_t ntiles = (B + 31) / 32;
  for (int64_t tile = (int64_t)blockIdx.x * nw + warp; tile < ntiles; tile += (int64_t)gridDim.x * nw)
  {
    const int64_t c0 = tile * 32;
    const int nc = (int)((B - c0) < 32 ? (B - c0) : 32);
    tile_load(sq, qpad, q + c0 * ldq, ldq, m.nq, nc, lane);
    tile_load(sv, vpad, v + c0 * ldv, ldv, m.nv, nc, lane);
    tile_load(st, vpad, tau + c0 * ldtau, ldtau, m.nv, nc, lane);
    BRBD_SYNCWARP();
    if (lane < nc) aba_thread(m, sq + lane * qpad, sv + lane * vpad, st + lane * vpad);
    BRBD_SYNCWARP();
    tile_store(a + c0 * lda, lda, st, vpad, m.nv, nc, lane);
    BRBD_SYNCWARP();
  }
}

} // namespace brbd
