// model_build.hpp — host side: brbd_flat_model (include/pinocchio_b200.h) -> ModelPOD<double>, with the checks the
// reference runs at algorithm entry (parents[i] < i, cumulative idx_q / idx_v, CRBAChecker crba.hxx:573-595) and
// the derived topology tables (nvSubtree / parents_fromRow: multibody/data.hxx:197-315).  Shared by capi.cu and by
// the CPU lane emulator of the warp-cooperative kernels (tests/cpp/coop_emu.cu).
#pragma once

#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/pinocchio_b200.h"
#include "engine.cuh"

namespace brbd
{
inline bool joint_is_unbounded(int t) { return t >= BRBD_JOINT_RUBX && t <= BRBD_JOINT_REVOLUTE_UNBOUNDED_UNALIGNED; }
inline bool joint_is_unaligned(int t)
{
  return t == BRBD_JOINT_REVOLUTE_UNALIGNED || t == BRBD_JOINT_PRISMATIC_UNALIGNED || t == BRBD_JOINT_REVOLUTE_UNBOUNDED_UNALIGNED;
}
inline int joint_nq_of(int t)
{
  if (joint_is_unbounded(t)) return 2;
  return (t <= BRBD_JOINT_PZ || joint_is_unaligned(t)) ? 1 : (t == BRBD_JOINT_FREEFLYER ? 7 : 4);
}
inline int joint_nv_of(int t) { return (t <= BRBD_JOINT_PZ || joint_is_unaligned(t) || joint_is_unbounded(t)) ? 1 : (t == BRBD_JOINT_FREEFLYER ? 6 : 3); }

template<class T> inline void fill_pod(ModelPOD<T> & P, const ModelPOD<double> & D)
{
  std::memset(&P, 0, sizeof(P));
  P.njoints = D.njoints; P.nq = D.nq; P.nv = D.nv; P.maxdepth = D.maxdepth;
  for (int i = 0; i < MAXJ; ++i)
  {
    P.parent[i] = D.parent[i]; P.type[i] = D.type[i]; P.idx_q[i] = D.idx_q[i]; P.idx_v[i] = D.idx_v[i];
    P.nvj[i] = D.nvj[i]; P.nvsub[i] = D.nvsub[i]; P.depth[i] = D.depth[i]; P.unb[i] = D.unb[i];
    for (int k = 0; k < 12; ++k) P.placement[i][k] = (T)D.placement[i][k];
    for (int k = 0; k < 10; ++k) P.inertia[i][k] = (T)D.inertia[i][k];
  }
  for (int k = 0; k < MAXNV; ++k)
  {
    P.dof_joint[k] = D.dof_joint[k]; P.parent_row[k] = D.parent_row[k]; P.armature[k] = (T)D.armature[k];
  }
  for (int k = 0; k < 3; ++k) P.gravity[k] = (T)D.gravity[k];
}

inline brbd_status build_model_pod(const brbd_flat_model * f, ModelPOD<double> & P, std::string & err)
{
  if (f->njoints < 1 || f->njoints > MAXJ)
    { err = "njoints must be in [1, " + std::to_string(MAXJ) + "]"; return BRBD_EINVAL; }
  if (f->nv > MAXNV || f->nv < 0) { err = "nv must be in [0, " + std::to_string(MAXNV) + "]"; return BRBD_EINVAL; }
  std::memset(&P, 0, sizeof(P));
  P.njoints = f->njoints; P.nq = f->nq; P.nv = f->nv;
  int nq = 0, nv = 0, maxdepth = 0;
  for (int i = 0; i < f->njoints; ++i)
  {
    P.parent[i] = f->parents[i];
    P.type[i] = f->joint_type[i];
    P.idx_q[i] = f->idx_q[i];
    P.idx_v[i] = f->idx_v[i];
    if (i == 0)
    {
      P.nvj[i] = 0; P.depth[i] = 0;
      continue;
    }
    if (P.type[i] < BRBD_JOINT_RX || P.type[i] > BRBD_JOINT_REVOLUTE_UNBOUNDED_UNALIGNED)
    { err = "joint " + std::to_string(i) + " has unsupported type tag " + std::to_string(f->joint_type[i]); return BRBD_EUNSUPPORTED_JOINT; }
    if (P.parent[i] < 0 || P.parent[i] >= i)
    { err = "parents[" + std::to_string(i) + "] must be < " + std::to_string(i); return BRBD_ETOPOLOGY; }
    if (P.idx_q[i] != nq || P.idx_v[i] != nv)
    { err = "idx_q / idx_v of joint " + std::to_string(i) + " are not cumulative"; return BRBD_EINVAL; }
    P.nvj[i] = joint_nv_of(P.type[i]);
    nq += joint_nq_of(P.type[i]);
    nv += P.nvj[i];
    P.depth[i] = P.depth[P.parent[i]] + 1;
    maxdepth = std::max(maxdepth, P.depth[i]);
    for (int k = 0; k < P.nvj[i]; ++k) P.dof_joint[P.idx_v[i] + k] = i;
  }
  if (nq != f->nq || nv != f->nv)
  { err = "nq / nv do not match the joint list"; return BRBD_EINVAL; }
  if (maxdepth >= MAXDEPTH)
  { err = "tree depth exceeds " + std::to_string(MAXDEPTH - 1); return BRBD_ETOPOLOGY; }
  P.maxdepth = maxdepth;
  // compact depth-first numbering (CRBAChecker, crba.hxx:573-595): the subtree of i is [i, last(i)]
  {
    std::vector<int> last(f->njoints);
    for (int i = 0; i < f->njoints; ++i) last[i] = i;
    for (int i = f->njoints - 1; i > 0; --i) last[P.parent[i]] = std::max(last[P.parent[i]], last[i]);
    for (int i = 1; i < f->njoints; ++i)
      for (int k = i + 1; k <= last[i]; ++k)
      {
        int a = k;
        while (a > i) a = P.parent[a];
        if (a != i)
        { err = "joints are not numbered depth-first (subtree of joint " + std::to_string(i) + " is not contiguous)"; return BRBD_ETOPOLOGY; }
      }
    for (int i = 0; i < f->njoints; ++i)
    {
      const int lc = last[i];
      P.nvsub[i] = (lc == 0) ? 0 : P.idx_v[lc] + P.nvj[lc] - (i == 0 ? 0 : P.idx_v[i]);
    }
  }
  for (int k = 0; k < MAXNV; ++k) P.parent_row[k] = -1;
  for (int j = 1; j < f->njoints; ++j)
  {
    const int parent = P.parent[j], iv = P.idx_v[j];
    P.parent_row[iv] = parent > 0 ? P.idx_v[parent] + P.nvj[parent] - 1 : -1;
    for (int r = 1; r < P.nvj[j]; ++r) P.parent_row[iv + r] = iv + r - 1;
  }
  // Working copies of the constants (R row-major, p; m, c, Symmetric3).  A joint about an arbitrary unit axis a
  // (JointModelRevoluteUnaligned / PrismaticUnaligned, reference joint-revolute-unaligned.hpp:668-672,
  // joint-prismatic-unaligned.hpp) is RE-FRAMED into an axis-aligned one: with Ra the rotation that takes z onto a,
  //   placement_i . Rot(a, q) = (placement_i . Ra) . Rot(z, q) . Ra^-1,
  // so the joint frame is turned by Ra — placement_i <- placement_i . Ra, the body's inertia and the placements of the
  // children re-expressed in the turned frame — and the tag becomes RZ / PZ.  q, v, a, tau and every joint-space result
  // (tau, ddq, M, Minv, the derivatives) do not depend on the choice of the joint frame, so the kernels never see an
  // unaligned joint.
  std::vector<double> plc(f->placement, f->placement + 12 * (size_t)f->njoints);
  std::vector<double> inr(f->inertia, f->inertia + 10 * (size_t)f->njoints);
  for (int i = 1; i < f->njoints; ++i)
  {
    if (!joint_is_unaligned(P.type[i])) continue;
    if (!f->axis) { err = "joint " + std::to_string(i) + " is unaligned but brbd_flat_model::axis is NULL"; return BRBD_EINVAL; }
    double a[3] = {f->axis[3 * i], f->axis[3 * i + 1], f->axis[3 * i + 2]};
    const double n = std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
    if (!(n > 1e-12)) { err = "joint " + std::to_string(i) + ": zero axis"; return BRBD_EINVAL; }
    for (double & x : a) x /= n;
    // Ra (row-major) with Ra z = a: rotation about u = z x a by the angle between z and a (math/rotation.hpp:26-55)
    double Ra[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    const double s = std::sqrt(a[0] * a[0] + a[1] * a[1]), c = a[2];
    if (s > 1e-14)
    {
      const double u[3] = {-a[1] / s, a[0] / s, 0.0};
      const double c1[3] = {(1 - c) * u[0], (1 - c) * u[1], 0.0}, su[3] = {s * u[0], s * u[1], 0.0};
      Ra[0] = c1[0] * u[0] + c; Ra[4] = c1[1] * u[1] + c; Ra[8] = c;
      Ra[1] = c1[0] * u[1] - su[2]; Ra[3] = c1[0] * u[1] + su[2];
      Ra[2] = su[1]; Ra[6] = -su[1];
      Ra[5] = -su[0]; Ra[7] = su[0];
    }
    else if (c < 0) { Ra[4] = -1; Ra[8] = -1; } // a = -z: half turn about x
    auto mul = [](const double * A, const double * B, double * C) { // C = A B, 3x3 row-major
      for (int r = 0; r < 3; ++r) for (int cc = 0; cc < 3; ++cc) C[3 * r + cc] = A[3 * r] * B[cc] + A[3 * r + 1] * B[3 + cc] + A[3 * r + 2] * B[6 + cc];
    };
    double RaT[9];
    for (int r = 0; r < 3; ++r) for (int cc = 0; cc < 3; ++cc) RaT[3 * r + cc] = Ra[3 * cc + r];
    double tmp[9];
    // placement_i <- placement_i . (Ra, 0)
    mul(&plc[12 * i], Ra, tmp);
    std::memcpy(&plc[12 * i], tmp, sizeof(tmp));
    // body inertia in the turned frame: c <- Ra^T c, I <- Ra^T I Ra
    {
      double * Y = &inr[10 * i];
      const double cc[3] = {Y[1], Y[2], Y[3]};
      for (int r = 0; r < 3; ++r) Y[1 + r] = RaT[3 * r] * cc[0] + RaT[3 * r + 1] * cc[1] + RaT[3 * r + 2] * cc[2];
      const double I[9] = {Y[4], Y[5], Y[7], Y[5], Y[6], Y[8], Y[7], Y[8], Y[9]}; // (xx,xy,yy,xz,yz,zz)
      double t1[9], t2[9];
      mul(RaT, I, t1);
      mul(t1, Ra, t2);
      Y[4] = t2[0]; Y[5] = t2[1]; Y[6] = t2[4]; Y[7] = t2[2]; Y[8] = t2[5]; Y[9] = t2[8];
    }
    // children: placement_k <- (Ra, 0)^-1 . placement_k
    for (int k = i + 1; k < f->njoints; ++k)
      if (P.parent[k] == i)
      {
        mul(RaT, &plc[12 * k], tmp);
        std::memcpy(&plc[12 * k], tmp, sizeof(tmp));
        const double pp[3] = {plc[12 * k + 9], plc[12 * k + 10], plc[12 * k + 11]};
        for (int r = 0; r < 3; ++r) plc[12 * k + 9 + r] = RaT[3 * r] * pp[0] + RaT[3 * r + 1] * pp[1] + RaT[3 * r + 2] * pp[2];
      }
    if (P.type[i] == BRBD_JOINT_REVOLUTE_UNBOUNDED_UNALIGNED) { P.type[i] = BRBD_JOINT_RZ; P.unb[i] = 1; }
    else P.type[i] = P.type[i] == BRBD_JOINT_REVOLUTE_UNALIGNED ? BRBD_JOINT_RZ : BRBD_JOINT_PZ;
  }
  // RUBX / RUBY / RUBZ: the revolute joint of the same axis whose configuration is (cos q, sin q); the kernels read the pair
  // where they would call sincos (JointModelRevoluteUnbounded::calc, joint-revolute-unbounded.hpp:154-162)
  for (int i = 1; i < f->njoints; ++i)
    if (P.type[i] >= BRBD_JOINT_RUBX && P.type[i] <= BRBD_JOINT_RUBZ)
    {
      P.type[i] = BRBD_JOINT_RX + (P.type[i] - BRBD_JOINT_RUBX);
      P.unb[i] = 1;
    }
  for (int i = 0; i < f->njoints; ++i)
  {
    const double * S = &plc[12 * i];         // R row-major, p
    double * D = P.placement[i];             // R by columns, p
    for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) D[3 * c + r] = S[3 * r + c];
    for (int k = 0; k < 3; ++k) D[9 + k] = S[9 + k];
    for (int k = 0; k < 10; ++k) P.inertia[i][k] = inr[10 * i + k];
  }
  for (int k = 0; k < f->nv; ++k) P.armature[k] = f->armature[k];
  for (int k = 0; k < 3; ++k) P.gravity[k] = f->gravity[k];
  return BRBD_OK;
}
} // namespace brbd
