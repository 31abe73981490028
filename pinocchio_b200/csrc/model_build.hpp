// model_build.hpp — host side: brbd_flat_model (include/pinocchio_b200.h) -> ModelPOD<double>, with the checks the
// reference runs at algorithm entry (parents[i] < i, cumulative idx_q / idx_v, CRBAChecker crba.hxx:573-595) and
// the derived topology tables (nvSubtree / parents_fromRow: multibody/data.hxx:197-315).  Shared by capi.cu and by
// the CPU lane emulator of the warp-cooperative kernels (tests/cpp/coop_emu.cu).
#pragma once

#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/pinocchio_b200.h"
#include "engine.cuh"

namespace brbd
{
inline int joint_nq_of(int t) { return t <= BRBD_JOINT_PZ ? 1 : (t == BRBD_JOINT_FREEFLYER ? 7 : 4); }
inline int joint_nv_of(int t) { return t <= BRBD_JOINT_PZ ? 1 : (t == BRBD_JOINT_FREEFLYER ? 6 : 3); }

template<class T> inline void fill_pod(ModelPOD<T> & P, const ModelPOD<double> & D)
{
  std::memset(&P, 0, sizeof(P));
  P.njoints = D.njoints; P.nq = D.nq; P.nv = D.nv; P.maxdepth = D.maxdepth;
  for (int i = 0; i < MAXJ; ++i)
  {
    P.parent[i] = D.parent[i]; P.type[i] = D.type[i]; P.idx_q[i] = D.idx_q[i]; P.idx_v[i] = D.idx_v[i];
    P.nvj[i] = D.nvj[i]; P.nvsub[i] = D.nvsub[i]; P.depth[i] = D.depth[i];
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
    if (P.type[i] < BRBD_JOINT_RX || P.type[i] > BRBD_JOINT_PLANAR)
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
  for (int i = 0; i < f->njoints; ++i)
  {
    const double * S = f->placement + 12 * i; // R row-major, p
    double * D = P.placement[i];             // R by columns, p
    for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) D[3 * c + r] = S[3 * r + c];
    for (int k = 0; k < 3; ++k) D[9 + k] = S[9 + k];
    for (int k = 0; k < 10; ++k) P.inertia[i][k] = f->inertia[10 * i + k];
  }
  for (int k = 0; k < f->nv; ++k) P.armature[k] = f->armature[k];
  for (int k = 0; k < 3; ++k) P.gravity[k] = f->gravity[k];
  return BRBD_OK;
}
} // namespace brbd
