// engine.cuh — device-side flattened Model, warp-tile staging and the per-joint type-tag switch.
//
// Replaces, for the batched path only:
//   * ModelTpl / DataTpl (multibody/model.hpp:97-205, data.hxx:30-315) by one POD staged per GPU
//     and copied into shared memory by every CTA;
//   * the boost::variant visitor dispatch (multibody/visitor/joint-unary-visitor.hpp:25-247) by a
//     switch over the joint type tag;
//   * JointModel*::calc (joint-revolute.hpp:791-820, joint-prismatic.hpp:698-725,
//     joint-free-flyer.hpp:342-374, joint-spherical.hpp:524-570, joint-planar.hpp:600-647).
#pragma once

#include <stdint.h>

#include "spatial.cuh"

namespace brbd
{

constexpr int MAXJ = 64;     // joints including the universe
constexpr int MAXNV = 64;    // tangent dimension
constexpr int MAXDEPTH = 16; // tree depth (universe = 0)

enum JointTag : int { J_RX = 0, J_RY = 1, J_RZ = 2, J_PX = 3, J_PY = 4, J_PZ = 5, J_FF = 6, J_SPH = 7, J_PLANAR = 8 };

// Everything a kernel needs to know about the model; constants already converted to T.
template<class T> struct ModelPOD
{
  int njoints, nq, nv, maxdepth;
  int parent[MAXJ], type[MAXJ], idx_q[MAXJ], idx_v[MAXJ], nvj[MAXJ];
  int nvsub[MAXJ];       // nvSubtree (data.hxx:197-242)
  int depth[MAXJ];       // universe = 0
  int unb[MAXJ];         // 1: revolute joint with an unbounded configuration, q = (cos, sin) (joint-revolute-unbounded.hpp:154-179)
  int dof_joint[MAXNV];  // joint owning each tangent row
  int parent_row[MAXNV]; // parents_fromRow (data.hxx:246-315)
  T placement[MAXJ][12]; // R by columns (c0, c1, c2) then p
  T inertia[MAXJ][10];   // m, c, (xx,xy,yy,xz,yz,zz)
  T armature[MAXNV];
  T gravity[3];
};

// global -> shared copy of a POD by the whole CTA; loads are issued 8 at a time per thread so that their latencies overlap
// (a small-batch launch runs one warp per CTA: a plain loop would serialise ~100 round trips to L2)
BRBD_DI void copy_words_to_smem(int * d, const int * s, int n)
{
  constexpr int U = 8;
  const int nt = blockDim.x;
  for (int k0 = threadIdx.x; k0 < n; k0 += nt * U)
  {
    int tmp[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
    {
      const int k = k0 + u * nt;
      tmp[u] = k < n ? s[k] : 0;
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
    {
      const int k = k0 + u * nt;
      if (k < n) d[k] = tmp[u];
    }
  }
}
template<class T> BRBD_DI void copy_model_to_smem(ModelPOD<T> * dst, const ModelPOD<T> * src)
{
  copy_words_to_smem(reinterpret_cast<int *>(dst), reinterpret_cast<const int *>(src), (int)(sizeof(ModelPOD<T>) / 4));
}

template<class T> BRBD_DI SE3<T> model_placement(const ModelPOD<T> & m, int i)
{
  const T * P = m.placement[i];
  SE3<T> X;
  X.R.c0 = Vec3<T>(P[0], P[1], P[2]);
  X.R.c1 = Vec3<T>(P[3], P[4], P[5]);
  X.R.c2 = Vec3<T>(P[6], P[7], P[8]);
  X.p = Vec3<T>(P[9], P[10], P[11]);
  return X;
}
template<class T> BRBD_DI Inertia<T> model_inertia(const ModelPOD<T> & m, int i)
{
  const T * Y = m.inertia[i];
  Inertia<T> I;
  I.m = Y[0];
  I.c = Vec3<T>(Y[1], Y[2], Y[3]);
  I.I.xx = Y[4]; I.I.xy = Y[5]; I.I.yy = Y[6]; I.I.xz = Y[7]; I.I.yz = Y[8]; I.I.zz = Y[9];
  return I;
}

// liMi = jointPlacements[i] * M_J(q): the plain SE3 product of the reference with the structural
// zeros / ones of M_J dropped (rnea.hxx:61, aba.hxx:117, crba.hxx:47).
template<class T> BRBD_DI SE3<T> joint_liMi(const ModelPOD<T> & m, int i, int type, const T * qj)
{
  const SE3<T> P = model_placement(m, i);
  SE3<T> X;
  switch (type)
  {
  case J_RX: case J_RY: case J_RZ: {
    T s, c;
    if (m.unb[i]) { c = qj[0]; s = qj[1]; } // JointModelRevoluteUnbounded::calc: data.M.setValues(sa, ca) from q = (ca, sa)
    else sincos_t(qj[0], &s, &c);
    X.p = P.p;
    if (type == J_RX) { X.R.c0 = P.R.c0; X.R.c1 = c * P.R.c1 + s * P.R.c2; X.R.c2 = c * P.R.c2 - s * P.R.c1; }
    else if (type == J_RY) { X.R.c1 = P.R.c1; X.R.c2 = c * P.R.c2 + s * P.R.c0; X.R.c0 = c * P.R.c0 - s * P.R.c2; }
    else { X.R.c2 = P.R.c2; X.R.c0 = c * P.R.c0 + s * P.R.c1; X.R.c1 = c * P.R.c1 - s * P.R.c0; }
    break;
  }
  case J_PX: case J_PY: case J_PZ: {
    X.R = P.R;
    X.p = P.p + qj[0] * P.R.col(type - J_PX);
    break;
  }
  case J_FF: {
    SE3<T> MJ;
    MJ.R = quat_to_mat(qj[3], qj[4], qj[5], qj[6]);
    MJ.p = Vec3<T>(qj[0], qj[1], qj[2]);
    X = P * MJ;
    break;
  }
  case J_SPH: {
    X.R = P.R * quat_to_mat(qj[0], qj[1], qj[2], qj[3]);
    X.p = P.p;
    break;
  }
  default: { // planar: q = (x, y, cos, sin)
    const T c = qj[2], s = qj[3];
    X.R.c0 = c * P.R.c0 + s * P.R.c1;
    X.R.c1 = c * P.R.c1 - s * P.R.c0;
    X.R.c2 = P.R.c2;
    X.p = P.p + qj[0] * P.R.c0 + qj[1] * P.R.c1;
    break;
  }
  }
  return X;
}

// joint velocity v_J = S * qdot in the joint frame
template<class T> BRBD_DI Motion<T> joint_velocity(int type, const T * vj)
{
  Motion<T> v = mzero<T>();
  switch (type)
  {
  case J_RX: v.ang.x = vj[0]; break;
  case J_RY: v.ang.y = vj[0]; break;
  case J_RZ: v.ang.z = vj[0]; break;
  case J_PX: v.lin.x = vj[0]; break;
  case J_PY: v.lin.y = vj[0]; break;
  case J_PZ: v.lin.z = vj[0]; break;
  case J_FF: v.lin = Vec3<T>(vj[0], vj[1], vj[2]); v.ang = Vec3<T>(vj[3], vj[4], vj[5]); break;
  case J_SPH: v.ang = Vec3<T>(vj[0], vj[1], vj[2]); break;
  default: v.lin = Vec3<T>(vj[0], vj[1], T(0)); v.ang = Vec3<T>(T(0), T(0), vj[2]); break;
  }
  return v;
}

// vi x v_J with the zeros of v_J dropped (rnea.hxx:67)
template<class T> BRBD_DI Motion<T> cross_joint_velocity(const Motion<T> & vi, int type, const T * vj)
{
  Motion<T> r;
  if (type <= J_RZ)
  {
    r.lin = cross_axis(vi.lin, type - J_RX, vj[0]);
    r.ang = cross_axis(vi.ang, type - J_RX, vj[0]);
  }
  else if (type <= J_PZ)
  {
    r.lin = cross_axis(vi.ang, type - J_PX, vj[0]);
    r.ang = Vec3<T>::zero();
  }
  else
    r = mcross(vi, joint_velocity(type, vj));
  return r;
}

// column k of S (joint frame) as a Motion
template<class T> BRBD_DI Motion<T> joint_S_col(int type, int k)
{
  Motion<T> s = mzero<T>();
  int row;
  if (type <= J_RZ) row = 3 + (type - J_RX);
  else if (type <= J_PZ) row = type - J_PX;
  else if (type == J_FF) row = k;
  else if (type == J_SPH) row = 3 + k;
  else row = (k == 2) ? 5 : k;
  if (row < 3) s.lin.set(row, T(1)); else s.ang.set(row - 3, T(1));
  return s;
}
// row of the 6-vector picked by S^T for dof k
BRBD_DI int joint_S_row(int type, int k)
{
  if (type <= J_RZ) return 3 + (type - J_RX);
  if (type <= J_PZ) return type - J_PX;
  if (type == J_FF) return k;
  if (type == J_SPH) return 3 + k;
  return (k == 2) ? 5 : k;
}
template<class T> BRBD_DI T get6(const Force<T> & f, int row) { return row < 3 ? f.lin.get(row) : f.ang.get(row - 3); }
template<class T> BRBD_DI T get6(const Motion<T> & f, int row) { return row < 3 ? f.lin.get(row) : f.ang.get(row - 3); }

// X . S_k : world/parent-frame column of the joint (aba.hxx:123 `oMi.act(jdata.S())`)
template<class T> BRBD_DI Motion<T> act_S_col(const SE3<T> & X, int type, int k)
{
  const int row = joint_S_row(type, k);
  Motion<T> r;
  if (row >= 3)
  {
    r.ang = X.R.col(row - 3);
    r.lin = cross(X.p, r.ang);
  }
  else
  {
    r.lin = X.R.col(row);
    r.ang = Vec3<T>::zero();
  }
  return r;
}

// ------------------------------------------------------------------------------------------
// Warp-tile staging: a warp owns 32 consecutive configurations (columns).  Inputs arrive as a
// column-major (rows x B) block with leading dimension ld (Eigen layout: one configuration =
// `rows` contiguous elements).  The tile is copied with coalesced accesses into shared memory as
// s[col * pad + row] (pad odd => each lane then walks its own row conflict-free).
// ------------------------------------------------------------------------------------------
template<class T>
BRBD_DI void tile_load(T * s, int pad, const T * g, int64_t ld, int rows, int ncols, int lane)
{
  // loads are issued in batches of 8 per lane before the first shared-memory store, so that their latencies overlap
  constexpr int U = 8;
  if (ld == rows)
  {
    const int total = rows * ncols;
    int c = 0, r = lane;
    while (r >= rows) { r -= rows; ++c; }
    for (int k0 = lane; k0 < total; k0 += 32 * U)
    {
      T tmp[U];
#pragma unroll
      for (int u = 0; u < U; ++u)
      {
        const int k = k0 + 32 * u;
        tmp[u] = k < total ? g[k] : T(0);
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
      {
        if (k0 + 32 * u < total) s[c * pad + r] = tmp[u];
        r += 32;
        while (r >= rows) { r -= rows; ++c; }
      }
    }
  }
  else
  {
    for (int c0 = 0; c0 < ncols; c0 += U)
      for (int r = lane; r < rows; r += 32)
      {
        T tmp[U];
#pragma unroll
        for (int u = 0; u < U; ++u) tmp[u] = c0 + u < ncols ? g[(int64_t)(c0 + u) * ld + r] : T(0);
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (c0 + u < ncols) s[(c0 + u) * pad + r] = tmp[u];
      }
  }
}
template<class T>
BRBD_DI void tile_store(T * g, int64_t ld, const T * s, int pad, int rows, int ncols, int lane)
{
  if (ld == rows)
  {
    const int total = rows * ncols;
    int c = 0, r = lane;
    while (r >= rows) { r -= rows; ++c; }
    for (int k = lane; k < total; k += 32)
    {
      g[k] = s[c * pad + r];
      r += 32;
      while (r >= rows) { r -= rows; ++c; }
    }
  }
  else
  {
    for (int c = 0; c < ncols; ++c)
      for (int r = lane; r < rows; r += 32) g[(int64_t)c * ld + r] = s[c * pad + r];
  }
}

// Column emitter for (nv x nv per configuration) matrix outputs.  Each lane fills the entries of
// ONE column of ITS configuration into its shared-memory row; the warp then writes the 32 columns
// (one per configuration, `nv` contiguous elements each) with coalesced stores.  Rows a lane
// does not `put` are written as zeros (the reference relies on pre-zeroed outputs,
// rnea-derivatives.hpp:104-106; crba.hpp:15-22).
template<class T> struct ColumnEmitter
{
  T * s;   // [32][pad]
  int pad; // >= nv, odd
  int nv, lane;
  BRBD_DI void init(T * buf, int pad_, int nv_, int lane_)
  {
    s = buf; pad = pad_; nv = nv_; lane = lane_;
    for (int k = lane; k < 32 * pad; k += 32) s[k] = T(0);
    BRBD_SYNCWARP();
  }
  BRBD_DI void put(int row, T val) { s[lane * pad + row] = val; }
  BRBD_DI void add(int row, T val) { s[lane * pad + row] += val; }
  BRBD_DI T get(int row) const { return s[lane * pad + row]; }
  // g points at element (row 0, column `col`) of configuration 0 of the tile; ld = elements between
  // consecutive configurations.
  BRBD_DI void flush(T * g, int64_t ld, int ncols)
  {
    BRBD_SYNCWARP();
    for (int c = 0; c < ncols; ++c)
      for (int r = lane; r < nv; r += 32) g[(int64_t)c * ld + r] = s[c * pad + r];
    BRBD_SYNCWARP();
    for (int k = lane; k < 32 * pad; k += 32) s[k] = T(0);
    BRBD_SYNCWARP();
  }
};

} // namespace brbd
