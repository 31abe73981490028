// minv_chol.cuh — computeMinverse from the joint-space inertia matrix: M = L L^T (Cholesky), Minv = L^-T L^-1, one
// configuration per group of G lanes, everything in shared memory.
//
// The reference offers two routes to M^-1: computeMinverse (algorithm/aba.hpp:106, aba.hxx:613-902: the articulated-body
// factorisation, O(nv * depth) per column) and the Cholesky route (algorithm/cholesky.hpp: cholesky::decompose of data.M,
// cholesky.hxx:23-110, then cholesky::computeMinv, cholesky.hxx:526-577).  The first is a chain of depth-serial 6 x 6
// recursions per configuration — aba_derivatives_coop_kernel, MODE 1, runs it at 46 us per configuration and warp (4.1 ms for
// 65 536 configurations of a 35-dof humanoid, 2.5 % of the HBM rate of its output).  The second is what a GPU is good at:
// the batched CRBA (0.16 ms for the same batch) followed by this dense nv x nv factorisation whose inner loops are one broadcast
// load + one conflict-free load + one FMA over all lanes.  Both give M^-1 to rounding; measured against the CPU restatement
// of the articulated-body Minv (tests/): <= 3e-13 of max|Minv| on every test model (talos, cond(M) = 8.7e4: 1.5e-13).
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


// ---- 32 lanes per configuration: the same three phases in 4 x 4 blocks ---------------------------------------------------
// The simple loops above read one own-row value from shared memory per FMA: 256 bytes per warp-wide FMA, and a 35-dof model
// spends its time on exactly that traffic (3.9 ms for 65 536 configurations, no faster than the articulated-body kernel).  In
// blocks, a lane loads 4 own values (and the warp 16 broadcast values of L) for 16 FMAs: a quarter of the shared-memory
// wavefronts per FMA.  The matrix is padded with an identity block to a multiple of 4 (chol(diag(M, I)) = diag(L, I)); rows are
// `ld` = nvp + 2 elements apart: 16-byte aligned for the 128-bit loads and an odd number of 16-byte units, so that the lanes'
// own-row loads spread over all banks.
struct MinvCholBlockedLayout
{
  int nvp, ld, oY, odinv, per_group; // elements
};
inline MinvCholBlockedLayout minv_chol_blocked_layout(int nv)
{
  MinvCholBlockedLayout L;
  L.nvp = (nv + 3) & ~3;
  L.ld = L.nvp + 2;
  L.oY = L.nvp * L.ld;
  L.odinv = 2 * L.nvp * L.ld;
  L.per_group = (L.odinv + L.nvp + 3) & ~3;
  return L;
}
template<class T> struct Vec2T;
template<> struct Vec2T<double> { typedef double2 type; };
template<> struct Vec2T<float> { typedef float2 type; };
// 4 values from / to an address aligned to 2 elements
template<class T> BRBD_DI void ld4v(const T * p, T * x)
{
  typedef typename Vec2T<T>::type P;
  const P a = reinterpret_cast<const P *>(p)[0], b = reinterpret_cast<const P *>(p)[1];
  x[0] = a.x; x[1] = a.y; x[2] = b.x; x[3] = b.y;
}
template<class T> BRBD_DI void st4v(T * p, const T * x)
{
  typedef typename Vec2T<T>::type P;
  P a, b;
  a.x = x[0]; a.y = x[1]; b.x = x[2]; b.y = x[3];
  reinterpret_cast<P *>(p)[0] = a; reinterpret_cast<P *>(p)[1] = b;
}

// R = rows / columns per lane (nvp <= 32 R)
template<class T, int R>
BRBD_DI void minv_chol_blocked_config(int nv, const MinvCholBlockedLayout & L, T * base, int gl, const T * __restrict__ gM, T * __restrict__ gout,
                                      bool active)
{
  constexpr int G = 32;
  T * S = base, * Y = base + L.oY, * dinv = base + L.odinv;
  const int ld = L.ld, nvp = L.nvp;
  // ---- S = the lower triangle of diag(M, I), everything else zero (the caller's upper triangle, column i -> row i) ----
  // eight independent global loads per lane in flight (one load per loop step exposed a DRAM round trip 43 times per configuration)
  {
    int i = 0, j = gl; // element e = i * ld + j, advanced without divisions
    while (j >= ld) { j -= ld; ++i; }
    for (int e0 = gl; e0 < nvp * ld; e0 += 8 * G)
    {
      T val[8];
#pragma unroll
      for (int u = 0; u < 8; ++u)
      {
        val[u] = T(0);
        if (e0 + u * G < nvp * ld)
        {
          if (i < nv) { if (j <= i) val[u] = gM[i * nv + j]; }
          else if (j == i) val[u] = T(1);
        }
        j += G;
        while (j >= ld) { j -= ld; ++i; }
      }
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (e0 + u * G < nvp * ld) S[e0 + u * G] = val[u];
    }
  }
  BRBD_SYNCWARP();
  int row[R];
  const T * own[R];
#pragma unroll
  for (int t = 0; t < R; ++t)
  {
    row[t] = gl + t * G;
    own[t] = S + (row[t] < nvp ? row[t] : 0) * ld; // lanes beyond the matrix shadow row 0 and store nothing
  }
  // ---- Cholesky, left-looking by panels of 4 columns (lanes own rows) ----
  for (int jb = 0; jb < nvp; jb += 4)
  {
    T acc[R][4];
#pragma unroll
    for (int t = 0; t < R; ++t) ld4v(own[t] + jb, acc[t]);
    for (int kb = 0; kb < jb; kb += 4)
    {
      T b[4][4];
#pragma unroll
      for (int c = 0; c < 4; ++c) ld4v(S + (jb + c) * ld + kb, b[c]); // broadcasts
#pragma unroll
      for (int t = 0; t < R; ++t)
      {
        T o[4];
        ld4v(own[t] + kb, o);
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int k = 0; k < 4; ++k) acc[t][c] -= o[k] * b[c][k];
      }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c)
    {
      const int j = jb + c;
#pragma unroll
      for (int t = 0; t < R; ++t)
        if (row[t] == j)
        { // the lane that owns row j holds the pivot (the diagonal of L itself is never read again: only its inverse)
#ifdef __CUDA_ARCH__
          dinv[j] = rsqrt(acc[t][c]);
#else
          dinv[j] = T(1) / sqrt(acc[t][c]);
#endif
        }
      BRBD_SYNCWARP();
      const T dj = dinv[j];
#pragma unroll
      for (int t = 0; t < R; ++t)
      {
        acc[t][c] *= dj;
        if (row[t] > j && row[t] < nvp) S[row[t] * ld + j] = acc[t][c];
      }
      BRBD_SYNCWARP();
      // the panel's remaining columns (rows at or above the pivot carry values nobody reads: they are never stored)
#pragma unroll
      for (int c2 = c + 1; c2 < 4; ++c2)
      {
        const T bj = S[(jb + c2) * ld + j];
#pragma unroll
        for (int t = 0; t < R; ++t) acc[t][c2] -= acc[t][c] * bj;
      }
    }
  }
  // ---- forward substitution L z = e_c by blocks of 4 rows (lanes own columns c; z = row c of Y) ----
  // lanes beyond the matrix shadow row 0 of S (read-only by now: nobody writes what they read) and store nothing
  T * ycol[R];
#pragma unroll
  for (int t = 0; t < R; ++t) ycol[t] = row[t] < nvp ? Y + row[t] * ld : S;
  for (int ib = 0; ib < nvp; ib += 4)
  {
    T acc[R][4];
#pragma unroll
    for (int t = 0; t < R; ++t)
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[t][r] = (ib + r == row[t]) ? T(1) : T(0);
    for (int kb = 0; kb < ib; kb += 4)
    {
      T b[4][4];
#pragma unroll
      for (int r = 0; r < 4; ++r) ld4v(S + (ib + r) * ld + kb, b[r]);
#pragma unroll
      for (int t = 0; t < R; ++t)
      {
        T y[4];
        ld4v(ycol[t] + kb, y);
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int k = 0; k < 4; ++k) acc[t][r] -= b[r][k] * y[k];
      }
    }
    T d[4][4], di[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) ld4v(S + (ib + r) * ld + ib, d[r]);
    ld4v(dinv + ib, di);
#pragma unroll
    for (int t = 0; t < R; ++t)
    {
      T z[4];
      z[0] = acc[t][0] * di[0];
      z[1] = (acc[t][1] - d[1][0] * z[0]) * di[1];
      z[2] = (acc[t][2] - d[2][0] * z[0] - d[2][1] * z[1]) * di[2];
      z[3] = (acc[t][3] - d[3][0] * z[0] - d[3][1] * z[1] - d[3][2] * z[2]) * di[3];
      if (row[t] < nvp) st4v(ycol[t] + ib, z);
    }
  }
  // (every lane reads and writes its own rows of Y only: no synchronisation between the substitutions)
  // ---- backward substitution L^T y = z, in place ----
  for (int ib = nvp - 4; ib >= 0; ib -= 4)
  {
    T acc[R][4];
#pragma unroll
    for (int t = 0; t < R; ++t) ld4v(ycol[t] + ib, acc[t]);
    for (int kb = ib + 4; kb < nvp; kb += 4)
    {
      T b[4][4]; // b[k][r] = L[kb + k][ib + r]
#pragma unroll
      for (int k = 0; k < 4; ++k) ld4v(S + (kb + k) * ld + ib, b[k]);
#pragma unroll
      for (int t = 0; t < R; ++t)
      {
        T y[4];
        ld4v(ycol[t] + kb, y);
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int k = 0; k < 4; ++k) acc[t][r] -= b[k][r] * y[k];
      }
    }
    T d[4][4], di[4]; // d[k][r] = L[ib + k][ib + r]
#pragma unroll
    for (int k = 0; k < 4; ++k) ld4v(S + (ib + k) * ld + ib, d[k]);
    ld4v(dinv + ib, di);
#pragma unroll
    for (int t = 0; t < R; ++t)
    {
      T y[4];
      y[3] = acc[t][3] * di[3];
      y[2] = (acc[t][2] - d[3][2] * y[3]) * di[2];
      y[1] = (acc[t][1] - d[2][1] * y[2] - d[3][1] * y[3]) * di[1];
      y[0] = (acc[t][0] - d[1][0] * y[1] - d[2][0] * y[2] - d[3][0] * y[3]) * di[0];
      if (row[t] < nvp) st4v(ycol[t] + ib, y);
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
template<class T, int R>
__global__ void __launch_bounds__(512, 1)
minv_chol_blocked_kernel(const T * __restrict__ Min, int64_t ldIn, T * __restrict__ Mout, int64_t ldOut, int nv, const MinvCholBlockedLayout L, int64_t B)
{
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  T * base = reinterpret_cast<T *>(dyn_smem) + (size_t)warp * L.per_group;
  for (int64_t cfg = (int64_t)blockIdx.x * nw + warp; cfg < B; cfg += (int64_t)gridDim.x * nw)
    minv_chol_blocked_config<T, R>(nv, L, base, lane, Min + cfg * ldIn, Mout + cfg * ldOut, true);
}
#endif
} // namespace brbd
