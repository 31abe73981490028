// minv_chol.cuh — computeMinverse from the joint-space inertia matrix: M = L L^T (Cholesky), Minv = L^-T L^-1, one
// configuration per group of G lanes, everything in shared memory.
//
// The reference offers two routes to M^-1: computeMinverse (algorithm/aba.hpp:106, aba.hxx:613-902: the articulated-body
// factorisation, O(nv * depth) per column) and the Cholesky route (algorithm/cholesky.hpp: cholesky::decompose of data.M,
// cholesky.hxx:23-110, then cholesky::computeMinv, cholesky.hxx:526-577).  The first is a chain of depth-serial 6 x 6
// recursions per configuration — aba_derivatives_coop_kernel, MODE 1, runs it at 46 us per configuration and warp (4.1 ms for
// 65 536 configurations of a 35-dof humanoid, 2.5 % of the HBM rate of its output).  The second is what a GPU is good at:
// the batched CRBA (0.16 ms for the same batch) followed by this dense nv x nv factorisation whose inner loops are one broadcast
// load + one conflict-free load + one FMA over all lanes.  Both give M^-1 to rounding; measured against the oracle's
// articulated-body Minv: <= 3e-13 of max|Minv| on every test model (talos, cond(M) = 8.7e4: 1.5e-13).
//
// Layout of a group's region (elements): S = L (nv x ld, row-major lower triangle, ld = nv | 1), Y (nv x ld: row c = column c of
// the inverse under construction), dinv (nv: 1 / L_jj).  Row i of S is column i of the caller's column-major upper triangle.
// Lanes own ROWS in the factorisation (row i = gl + t G) and COLUMNS of the inverse in the two substitutions.
#pragma once
#include "engine.cuh"

namespace brbd
{
struct MinvCholLayout
{
  int ld, oY, odinv, per_group; // elements
};
inline MinvCholLayout minv_chol_layout(int nv)
{
  MinvCholLayout L;
  L.ld = nv | 1;
  L.oY = nv * L.ld;
  L.odinv = 2 * nv * L.ld;
  L.per_group = (L.odinv + nv + 1) & ~1;
  return L;
}

// sum_k b[k * bs] * o_t[k], k < n, for the lane's R rows: four independent partial sums per row — a single accumulator is a
// chain of n dependent (load -> FMA) steps, which is what the first version of this kernel spent its time in (180 000 cycles
// per configuration for ~2 000 loop steps)
template<class T, int R>
BRBD_DI void dot_rows(const T * b, int bs, const T * const * o, int n, T * r)
{
  T a[R][4];
#pragma unroll
  for (int t = 0; t < R; ++t) a[t][0] = a[t][1] = a[t][2] = a[t][3] = T(0);
  int k = 0;
  for (; k + 4 <= n; k += 4)
  {
    const T b0 = b[k * bs], b1 = b[(k + 1) * bs], b2 = b[(k + 2) * bs], b3 = b[(k + 3) * bs]; // broadcasts
#pragma unroll
    for (int t = 0; t < R; ++t)
    {
      const T * ot = o[t] + k; // own row: conflict-free (ld odd)
      a[t][0] += b0 * ot[0]; a[t][1] += b1 * ot[1]; a[t][2] += b2 * ot[2]; a[t][3] += b3 * ot[3];
    }
  }
  for (; k < n; ++k)
  {
    const T b0 = b[k * bs];
#pragma unroll
    for (int t = 0; t < R; ++t) a[t][0] += b0 * o[t][k];
  }
#pragma unroll
  for (int t = 0; t < R; ++t) r[t] = (a[t][0] + a[t][1]) + (a[t][2] + a[t][3]);
}

// R = rows / columns per lane (nv <= R G)
template<class T, int G, int R>
BRBD_DI void minv_chol_config(int nv, const MinvCholLayout & L, T * base, int gl, const T * __restrict__ gM, T * __restrict__ gout, bool active)
{
  T * S = base, * Y = base + L.oY, * dinv = base + L.odinv;
  const int ld = L.ld;
  // ---- the upper triangle of the caller's column-major M = the lower triangle of S, row by row ----
  {
    int i = 0, j = gl;
    while (j >= nv) { j -= nv; ++i; }
    for (int e = gl; e < nv * nv; e += G)
    {
      if (j <= i) S[i * ld + j] = gM[e];
      j += G;
      while (j >= nv) { j -= nv; ++i; }
    }
  }
  BRBD_SYNCWARP();
  // ---- Cholesky, left-looking: column j of L from the columns before it (lanes own rows) ----
  for (int j = 0; j < nv; ++j)
  {
    T s[R];
#pragma unroll
    for (int t = 0; t < R; ++t)
    {
      const int i = gl + t * G;
      s[t] = (i >= j && i < nv) ? S[i * ld + j] : T(0);
    }
    {
      const T * own[R];
      T acc[R];
#pragma unroll
      for (int t = 0; t < R; ++t) own[t] = S + ((gl + t * G) < nv ? (gl + t * G) : 0) * ld;
      dot_rows<T, R>(S + j * ld, 1, own, j, acc);
#pragma unroll
      for (int t = 0; t < R; ++t) s[t] -= acc[t];
    }
    // the lane that owns row j holds the pivot
#pragma unroll
    for (int t = 0; t < R; ++t)
      if (gl + t * G == j)
      {
        const T d = sqrt(s[t]);
        S[j * ld + j] = d;
        dinv[j] = T(1) / d;
      }
    BRBD_SYNCWARP();
    const T dj = dinv[j];
#pragma unroll
    for (int t = 0; t < R; ++t)
    {
      const int i = gl + t * G;
      if (i > j && i < nv) S[i * ld + j] = s[t] * dj;
    }
    BRBD_SYNCWARP();
  }
  // ---- forward substitution L z = e_c (lanes own columns c; z = row c of Y) ----
  for (int i = 0; i < nv; ++i)
  {
    T s[R];
#pragma unroll
    for (int t = 0; t < R; ++t) s[t] = (gl + t * G == i) ? T(1) : T(0);
    {
      const T * own[R];
      T acc[R];
#pragma unroll
      for (int t = 0; t < R; ++t) own[t] = Y + ((gl + t * G) < nv ? (gl + t * G) : 0) * ld;
      dot_rows<T, R>(S + i * ld, 1, own, i, acc);
#pragma unroll
      for (int t = 0; t < R; ++t) s[t] -= acc[t];
    }
    const T di = dinv[i];
#pragma unroll
    for (int t = 0; t < R; ++t)
    {
      const int c = gl + t * G;
      if (c < nv) Y[c * ld + i] = s[t] * di; // 0 for i < c: L^-1 is lower triangular
    }
  }
  // (every lane reads and writes its own rows of Y only: no synchronisation between the substitutions)
  // ---- backward substitution L^T y = z, in place ----
  for (int i = nv - 1; i >= 0; --i)
  {
    T s[R];
#pragma unroll
    for (int t = 0; t < R; ++t)
    {
      const int c = gl + t * G;
      s[t] = Y[(c < nv ? c : 0) * ld + i];
    }
    {
      const T * own[R];
      T acc[R];
#pragma unroll
      for (int t = 0; t < R; ++t) own[t] = Y + ((gl + t * G) < nv ? (gl + t * G) : 0) * ld + (i + 1);
      dot_rows<T, R>(S + (i + 1) * ld + i, ld, own, nv - 1 - i, acc); // column i of L below the diagonal
#pragma unroll
      for (int t = 0; t < R; ++t) s[t] -= acc[t];
    }
    const T di = dinv[i];
#pragma unroll
    for (int t = 0; t < R; ++t)
    {
      const int c = gl + t * G;
      if (c < nv) Y[c * ld + i] = s[t] * di;
    }
  }
  BRBD_SYNCWARP();
  // ---- out: column c of the result = row c of Y; like the reference's data.Minv only the upper triangle, zeros below ----
  if (active)
  {
    int c = 0, r = gl;
    while (r >= nv) { r -= nv; ++c; }
    for (int e = gl; e < nv * nv; e += G)
    {
      gout[e] = r <= c ? Y[c * ld + r] : T(0);
      r += G;
      while (r >= nv) { r -= nv; ++c; }
    }
  }
  BRBD_SYNCWARP();
}

#ifdef __CUDACC__
template<class T, int G, int R>
__global__ void __launch_bounds__(512, 1)
minv_chol_kernel(const T * __restrict__ Min, int64_t ldIn, T * __restrict__ Mout, int64_t ldOut, int nv, const MinvCholLayout L, int64_t B)
{
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  constexpr int GPW = 32 / G; // configurations per warp
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int gl = lane % G, grp = lane / G;
  T * base = reinterpret_cast<T *>(dyn_smem) + (size_t)(warp * GPW + grp) * L.per_group;
  const int64_t ntiles = (B + GPW - 1) / GPW;
  for (int64_t tile = (int64_t)blockIdx.x * nw + warp; tile < ntiles; tile += (int64_t)gridDim.x * nw)
  {
    int64_t cfg = tile * GPW + grp;
    const bool active = cfg < B;
    if (!active) cfg = B - 1; // idle groups shadow the last configuration (stores suppressed)
    minv_chol_config<T, G, R>(nv, L, base, gl, Min + cfg * ldIn, Mout + cfg * ldOut, active);
  }
}
#endif
} // namespace brbd
