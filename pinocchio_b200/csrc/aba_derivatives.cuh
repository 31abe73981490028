// aba_derivatives.cuh — batched computeABADerivatives.
//
// Restates impl::computeABADerivatives (reference: include/pinocchio/algorithm/aba-derivatives.hxx
// :380-453):
//   pass 1  ComputeABADerivativesForwardStep1   (:38-79)    = ABA forward
//   pass 2  ComputeABADerivativesBackwardStep1  (:97-170)   = ABA backward + rows of Minv (upper part)
//   pass 3  ComputeABADerivativesForwardStep2   (:188-256)  = ABA forward 2 + Minv completion + derivative columns
//   pass 4  ComputeABADerivativesBackwardStep2  (:283-367)  = dtau_dq / dtau_dv (shared with rnea_derivatives.cuh)
//   tail    Minv symmetrisation (:448-449) and the two dense products -Minv*dtau_dq, -Minv*dtau_dv (:451-452)
//
// Kernel A (one configuration per thread) runs passes 1-4 and emits Minv, dtau_dq and dtau_dv column by
// column straight into the caller's ddq_dtau / ddq_dq / ddq_dv buffers; kernel B (one warp per
// configuration, rows across lanes) overwrites ddq_dq / ddq_dv with the two products.  The large
// per-thread tables (Minv, Fcrb per tree depth) live in an explicit global workspace laid out
// [entry][thread] so that every access is coalesced.
#pragma once

#include "aba.cuh"
#include "rnea_derivatives.cuh"

namespace brbd
{

template<class T> struct ThreadWS // thread-private strided view of a global workspace
{
  T * p;
  int64_t stride;
  BRBD_DI T & at(int k) const { return p[(int64_t)k * stride]; }
};

template<class T> struct AbaDerivState
{
  DerivState<T> ds;
  T U[MAXNV][6], UD[MAXNV][6], Dinv[MAXNV][6];
  T ab[MAXJ][6];  // oa_gf: bias after pass 1
  T Ia[MAXJ][21];
  T oh[MAXJ][6], ov[MAXJ][6];
  T F0[MAXNV][6]; // data.Fcrb[0]
  T oMi[MAXDEPTH][12];
  T ag[MAXDEPTH][6];
};

template<class T>
BRBD_DI void aba_derivatives_thread(const ModelPOD<T> & m, const T * q, const T * v, T * u, ColumnEmitter<T> & eq,
                                    ColumnEmitter<T> & ev, ColumnEmitter<T> & em, T * gq, int64_t ld_q, T * gv,
                                    int64_t ld_v, T * gm, int64_t ld_m, const ThreadWS<T> & minv,
                                    const ThreadWS<T> & fd, int nc)
{
  AbaDerivState<T> st;
  const int nj = m.njoints, nv = m.nv;
#define MINV(r, c) minv.at((r) * nv + (c))
#define FD(d, c, r) fd.at(((d) * nv + (c)) * 6 + (r))

  // ---- pass 1 (aba-derivatives.hxx:38-79) ----------------------------------------------------
  for (int i = 1; i < nj; ++i)
  {
    const int type = m.type[i], parent = m.parent[i], iq = m.idx_q[i], iv = m.idx_v[i], d = m.depth[i], nvj = m.nvj[i];
    SE3<T> X = joint_liMi(m, i, type, q + iq);
    if (parent > 0) X = load_se3(st.oMi[d - 1]) * X;
    store_se3(st.oMi[d], X);
    for (int k = 0; k < nvj; ++k) store6(st.ds.J[iv + k], act_S_col(X, type, k));
    Motion<T> ov = X.act(joint_velocity(type, v + iv));
    Motion<T> ab = mzero<T>();
    if (parent > 0)
    {
      const Motion<T> ovp = load_motion(st.ov[parent]);
      ov += ovp;
      ab = mcross(ovp, ov);
    }
    store6(st.ov[i], ov);
    store6(st.ab[i], ab);
    const Inertia<T> Y = act(X, model_inertia(m, i));
    store_inertia(st.ds.Y[i], Y);
    T Ia[21];
    inertia_to_sym6(Y, Ia);
    for (int k = 0; k < 21; ++k) st.Ia[i][k] = Ia[k];
    const Force<T> oh = Y * ov;
    store6(st.oh[i], oh);
    store6(st.ds.of[i], fcross(ov, oh));
  }

  // ---- pass 2 (aba-derivatives.hxx:97-170) ---------------------------------------------------
  for (int r = 0; r < nv; ++r)
    for (int c = r; c < nv; ++c) MINV(r, c) = T(0); // Minv_.triangularView<Upper>().setZero() (:414)
  for (int c = 0; c < nv; ++c)
    for (int r = 0; r < 6; ++r) st.F0[c][r] = T(0); // data.Fcrb[0].setZero() (:427)
  for (int i = nj - 1; i > 0; --i)
  {
    const int parent = m.parent[i], iv = m.idx_v[i], nvj = m.nvj[i], nsub = m.nvsub[i];
    T Ia[21];
    for (int k = 0; k < 21; ++k) Ia[k] = st.Ia[i][k];
    Force<T> fi = load_force(st.ds.of[i]);
    T U[6][6], StU[6][6], Di[6][6], UD[6][6], uj[6];
    for (int k = 0; k < nvj; ++k)
    {
      const Motion<T> J = load_motion(st.ds.J[iv + k]);
      uj[k] = u[iv + k] - dot6(J, fi);
      u[iv + k] = uj[k];
      T Jv[6], Uk[6];
      m2a(J, Jv);
      sym6_mul(Ia, Jv, Uk);
      for (int r = 0; r < 6; ++r) { U[r][k] = Uk[r]; st.U[iv + k][r] = Uk[r]; }
    }
    for (int a = 0; a < nvj; ++a)
    {
      T Jv[6];
      m2a(load_motion(st.ds.J[iv + a]), Jv);
      for (int b = 0; b < nvj; ++b)
      {
        T acc = Jv[0] * U[0][b];
        for (int r = 1; r < 6; ++r) acc += Jv[r] * U[r][b];
        StU[a][b] = acc;
      }
      StU[a][a] += m.armature[iv + a];
    }
    if (nvj == 1) Di[0][0] = T(1) / StU[0][0];
    else llt_inverse(nvj, StU, Di);
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
      for (int c = 0; c < nvj; ++c) { st.Dinv[iv + k][c] = Di[k][c]; MINV(iv + k, iv + c) = Di[k][c]; }
    }
    const int nvc = nsub - nvj;
    if (nvc > 0)
    {
      for (int k = 0; k < nvj; ++k)
      {
        // SDinv_k = J_cols * Dinv(:, k)
        T SD[6];
        for (int r = 0; r < 6; ++r)
        {
          T acc = st.ds.J[iv][r] * Di[0][k];
          for (int c = 1; c < nvj; ++c) acc += st.ds.J[iv + c][r] * Di[c][k];
          SD[r] = acc;
        }
        for (int c = 0; c < nvc; ++c)
        {
          const T * F = st.F0[iv + nvj + c];
          T acc = SD[0] * F[0];
          for (int r = 1; r < 6; ++r) acc += SD[r] * F[r];
          MINV(iv + k, iv + nvj + c) = -acc;
        }
      }
      if (parent > 0)
        for (int c = 0; c < nsub; ++c)
          for (int r = 0; r < 6; ++r)
          {
            T acc = U[r][0] * MINV(iv, iv + c);
            for (int k = 1; k < nvj; ++k) acc += U[r][k] * MINV(iv + k, iv + c);
            st.F0[iv + c][r] += acc;
          }
    }
    else
    {
      for (int c = 0; c < nsub; ++c)
        for (int r = 0; r < 6; ++r)
        {
          T acc = U[r][0] * MINV(iv, iv + c);
          for (int k = 1; k < nvj; ++k) acc += U[r][k] * MINV(iv + k, iv + c);
          st.F0[iv + c][r] = acc;
        }
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
      Force<T> fp = load_force(st.ds.of[parent]);
      fp.lin.x += fa[0]; fp.lin.y += fa[1]; fp.lin.z += fa[2];
      fp.ang.x += fa[3]; fp.ang.y += fa[4]; fp.ang.z += fa[5];
      store6(st.ds.of[parent], fp);
    }
  }

  // ---- pass 3 (aba-derivatives.hxx:188-256) --------------------------------------------------
  {
    Motion<T> g0 = mzero<T>();
    g0.lin = Vec3<T>(-m.gravity[0], -m.gravity[1], -m.gravity[2]); // data.oa_gf[0] = -gravity (:410)
    store6(st.ag[0], g0);
  }
  for (int i = 1; i < nj; ++i)
  {
    const int parent = m.parent[i], iv = m.idx_v[i], nvj = m.nvj[i], d = m.depth[i];
    const Motion<T> agp = load_motion(st.ag[d - 1]);
    Motion<T> ag = load_motion(st.ab[i]);
    ag += agp;
    T agv[6], dd[6];
    m2a(ag, agv);
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
      u[iv + k] = dd[k]; // data.ddq
      const Motion<T> J = load_motion(st.ds.J[iv + k]);
      ag.lin += dd[k] * J.lin;
      ag.ang += dd[k] * J.ang;
    }
    store6(st.ag[d], ag);
    const Motion<T> ov = load_motion(st.ov[i]);
    const Inertia<T> Y = load_inertia(st.ds.Y[i]);
    const Force<T> oh = load_force(st.oh[i]);
    Force<T> of = Y * ag;
    of += fcross(ov, oh);
    store6(st.ds.of[i], of);
    // Minv rows of this joint, all columns to the right (:220-234)
    const int nr = nv - iv;
    if (parent > 0)
      for (int k = 0; k < nvj; ++k)
        for (int c = 0; c < nr; ++c)
        {
          T acc = st.UD[iv + k][0] * FD(d - 1, iv + c, 0);
          for (int r = 1; r < 6; ++r) acc += st.UD[iv + k][r] * FD(d - 1, iv + c, r);
          MINV(iv + k, iv + c) -= acc;
        }
    for (int c = 0; c < nr; ++c)
      for (int r = 0; r < 6; ++r)
      {
        T acc = st.ds.J[iv][r] * MINV(iv, iv + c);
        for (int k = 1; k < nvj; ++k) acc += st.ds.J[iv + k][r] * MINV(iv + k, iv + c);
        if (parent > 0) acc += FD(d - 1, iv + c, r);
        FD(d, iv + c, r) = acc;
      }
    Motion<T> ovp = mzero<T>();
    if (parent > 0) ovp = load_motion(st.ov[parent]);
    deriv_columns(m, st.ds, i, ov, ovp, agp);
    store_dy(st.ds.dY[i], inertia_variation(Y, ov, oh));
  }

  // ---- pass 4 (aba-derivatives.hxx:283-367): dtau_dq, dtau_dv columns ------------------------------
  deriv_backward<T, false>(m, st.ds, eq, ev, em, gq, ld_q, gv, ld_v, (T *)nullptr, 0, (T *)nullptr, nc);

  // ---- Minv, symmetrised (:448-449), column by column -------------------------------------------
  for (int c = 0; c < nv; ++c)
  {
    for (int r = 0; r < c; ++r) em.put(r, MINV(r, c));
    for (int r = c; r < nv; ++r) em.put(r, MINV(c, r));
    em.flush(gm + (int64_t)c * nv, ld_m, nc);
  }
#undef MINV
#undef FD
}

template<class T>
__global__ void __launch_bounds__(512)
aba_derivatives_sweep_kernel(const ModelPOD<T> * __restrict__ gmod, const T * __restrict__ q, int64_t ldq,
                             const T * __restrict__ v, int64_t ldv, const T * __restrict__ tau, int64_t ldtau,
                             T * __restrict__ dq, int64_t ld_dq, T * __restrict__ dv, int64_t ld_dv,
                             T * __restrict__ dtau, int64_t ld_dtau, T * __restrict__ ddq, int64_t ldddq,
                             T * __restrict__ workspace, int64_t B)
{
  __shared__ ModelPOD<T> m;
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  copy_model_to_smem(&m, gmod);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int qpad = m.nq | 1, vpad = m.nv | 1;
  T * sq = reinterpret_cast<T *>(dyn_smem) + (size_t)warp * 32 * (qpad + 5 * vpad);
  T * sv = sq + 32 * qpad;
  T * su = sv + 32 * vpad;
  ColumnEmitter<T> eq, ev, em;
  eq.init(su + 32 * vpad, vpad, m.nv, lane);
  ev.init(su + 64 * vpad, vpad, m.nv, lane);
  em.init(su + 96 * vpad, vpad, m.nv, lane);
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  ThreadWS<T> minv{workspace + gtid, nthreads};
  ThreadWS<T> fd{workspace + (int64_t)m.nv * m.nv * nthreads + gtid, nthreads};
  const int64_t ntiles = (B + 31) / 32;
  for (int64_t tile = (int64_t)blockIdx.x * nw + warp; tile < ntiles; tile += (int64_t)gridDim.x * nw)
  {
    const int64_t c0 = tile * 32;
    const int nc = (int)((B - c0) < 32 ? (B - c0) : 32);
    tile_load(sq, qpad, q + c0 * ldq, ldq, m.nq, nc, lane);
    tile_load(sv, vpad, v + c0 * ldv, ldv, m.nv, nc, lane);
    tile_load(su, vpad, tau + c0 * ldtau, ldtau, m.nv, nc, lane);
    BRBD_SYNCWARP();
    pad_tile_rows(sq, qpad, m.nq, nc, lane);
    pad_tile_rows(sv, vpad, m.nv, nc, lane);
    pad_tile_rows(su, vpad, m.nv, nc, lane);
    BRBD_SYNCWARP();
    aba_derivatives_thread(m, sq + lane * qpad, sv + lane * vpad, su + lane * vpad, eq, ev, em, dq + c0 * ld_dq, ld_dq,
                           dv + c0 * ld_dv, ld_dv, dtau + c0 * ld_dtau, ld_dtau, minv, fd, nc);
    BRBD_SYNCWARP();
    if (ddq) tile_store(ddq + c0 * ldddq, ldddq, su, vpad, m.nv, nc, lane);
    BRBD_SYNCWARP();
  }
}

// Kernel B: D <- -Minv * D for D in {dq, dv}; one warp per configuration, in place.
template<class T>
__global__ void __launch_bounds__(256)
aba_derivatives_gemm_kernel(int nv, T * __restrict__ dq, int64_t ld_dq, T * __restrict__ dv, int64_t ld_dv,
                            const T * __restrict__ minv, int64_t ld_m, int64_t B)
{
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int pad = nv + 1;
  T * sM = reinterpret_cast<T *>(dyn_smem) + (size_t)warp * 3 * nv * pad;
  T * sA = sM + nv * pad;
  T * sB = sA + nv * pad;
  const int nn = nv * nv;
  for (int64_t cfg = (int64_t)blockIdx.x * nw + warp; cfg < B; cfg += (int64_t)gridDim.x * nw)
  {
    const T * gM = minv + cfg * ld_m;
    T * gA = dq + cfg * ld_dq;
    T * gB = dv + cfg * ld_dv;
    {
      int c = 0, r = lane;
      while (r >= nv) { r -= nv; ++c; }
      for (int k = lane; k < nn; k += 32)
      {
        sM[r * pad + c] = gM[k]; // row r of Minv contiguous in shared memory
        sA[c * pad + r] = gA[k];
        sB[c * pad + r] = gB[k];
        r += 32;
        while (r >= nv) { r -= nv; ++c; }
      }
    }
    BRBD_SYNCWARP();
    for (int rb = 0; rb < nv; rb += 32)
    {
      const int r = rb + lane;
      if (r < nv)
      {
        const T * mr = sM + r * pad;
        for (int c = 0; c < nv; ++c)
        {
          const T * a = sA + c * pad;
          const T * b = sB + c * pad;
          T accA = T(0), accB = T(0);
          for (int k = 0; k < nv; ++k)
          {
            const T mk = mr[k];
            accA += mk * a[k];
            accB += mk * b[k];
          }
          gA[c * nv + r] = -accA;
          gB[c * nv + r] = -accB;
        }
      }
    }
    BRBD_SYNCWARP();
  }
}

} // namespace brbd
