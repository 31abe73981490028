// spatial.cuh — 6-D spatial algebra as inlined sm_100a device code (templated on the element type).
//
// Each operation restates the formula of the reference header it replaces;
// paths are relative to /root/reference/include/pinocchio/spatial/.  Conventions: 6-vectors are
// (linear[0:3], angular[3:6]) for motions and forces (force-tpl.hpp:27-28); X = (R, p) maps
// child -> parent coordinates; Symmetric3 packs [xx, xy, yy, xz, yz, zz] (symmetric3.hpp:46-51).
#pragma once

#include <cuda_runtime.h>

namespace brbd
{

#ifndef BRBD_DI
#define BRBD_DI __device__ __forceinline__
#endif
// warp barrier of the warp-cooperative kernels; the CPU lane emulator of tests/cpp/coop_emu.cu overrides it
#ifndef BRBD_SYNCWARP
#define BRBD_SYNCWARP() __syncwarp()
#endif

// max(a, b) as the reference's math::max; an overload for the code generator's recording scalar lives in codegen/sym.hpp
BRBD_DI double max_t(double a, double b) { return a > b ? a : b; }
BRBD_DI float max_t(float a, float b) { return a > b ? a : b; }

template<class T> struct Vec3
{
  T x, y, z;
  BRBD_DI Vec3() {}
  BRBD_DI Vec3(T a, T b, T c) : x(a), y(b), z(c) {}
  BRBD_DI static Vec3 zero() { return Vec3(T(0), T(0), T(0)); }
  BRBD_DI T get(int k) const { return k == 0 ? x : (k == 1 ? y : z); }
  BRBD_DI void set(int k, T v) { if (k == 0) x = v; else if (k == 1) y = v; else z = v; }
};
template<class T> BRBD_DI Vec3<T> operator+(const Vec3<T> & a, const Vec3<T> & b) { return Vec3<T>(a.x + b.x, a.y + b.y, a.z + b.z); }
template<class T> BRBD_DI Vec3<T> operator-(const Vec3<T> & a, const Vec3<T> & b) { return Vec3<T>(a.x - b.x, a.y - b.y, a.z - b.z); }
template<class T> BRBD_DI Vec3<T> operator-(const Vec3<T> & a) { return Vec3<T>(-a.x, -a.y, -a.z); }
template<class T> BRBD_DI Vec3<T> operator*(T s, const Vec3<T> & a) { return Vec3<T>(s * a.x, s * a.y, s * a.z); }
template<class T> BRBD_DI void operator+=(Vec3<T> & a, const Vec3<T> & b) { a.x += b.x; a.y += b.y; a.z += b.z; }
template<class T> BRBD_DI void operator-=(Vec3<T> & a, const Vec3<T> & b) { a.x -= b.x; a.y -= b.y; a.z -= b.z; }
template<class T> BRBD_DI T dot(const Vec3<T> & a, const Vec3<T> & b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template<class T> BRBD_DI Vec3<T> cross(const Vec3<T> & a, const Vec3<T> & b)
{
  return Vec3<T>(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
// a x (s e_k) without the structural zeros
template<class T> BRBD_DI Vec3<T> cross_axis(const Vec3<T> & a, int k, T s)
{
  if (k == 0) return Vec3<T>(T(0), a.z * s, -(a.y * s));
  if (k == 1) return Vec3<T>(-(a.z * s), T(0), a.x * s);
  return Vec3<T>(a.y * s, -(a.x * s), T(0));
}

// 3x3 rotation stored by columns (c0, c1, c2)
template<class T> struct Mat3
{
  Vec3<T> c0, c1, c2;
  BRBD_DI Vec3<T> col(int k) const { return k == 0 ? c0 : (k == 1 ? c1 : c2); }
  BRBD_DI static Mat3 identity()
  {
    Mat3 r;
    r.c0 = Vec3<T>(T(1), T(0), T(0)); r.c1 = Vec3<T>(T(0), T(1), T(0)); r.c2 = Vec3<T>(T(0), T(0), T(1));
    return r;
  }
};
template<class T> BRBD_DI Vec3<T> operator*(const Mat3<T> & A, const Vec3<T> & v) // A v
{
  return Vec3<T>(A.c0.x * v.x + A.c1.x * v.y + A.c2.x * v.z, A.c0.y * v.x + A.c1.y * v.y + A.c2.y * v.z,
                 A.c0.z * v.x + A.c1.z * v.y + A.c2.z * v.z);
}
template<class T> BRBD_DI Vec3<T> tmul(const Mat3<T> & A, const Vec3<T> & v) // A^T v
{
  return Vec3<T>(dot(A.c0, v), dot(A.c1, v), dot(A.c2, v));
}
template<class T> BRBD_DI Mat3<T> operator*(const Mat3<T> & A, const Mat3<T> & B)
{
  Mat3<T> r;
  r.c0 = A * B.c0; r.c1 = A * B.c1; r.c2 = A * B.c2;
  return r;
}

template<class T> struct Motion { Vec3<T> lin, ang; };
template<class T> struct Force { Vec3<T> lin, ang; };
template<class T> BRBD_DI Motion<T> mzero() { Motion<T> m; m.lin = Vec3<T>::zero(); m.ang = Vec3<T>::zero(); return m; }
template<class T> BRBD_DI Force<T> fzero() { Force<T> m; m.lin = Vec3<T>::zero(); m.ang = Vec3<T>::zero(); return m; }
template<class T> BRBD_DI Motion<T> operator+(const Motion<T> & a, const Motion<T> & b) { Motion<T> r; r.lin = a.lin + b.lin; r.ang = a.ang + b.ang; return r; }
template<class T> BRBD_DI Motion<T> operator-(const Motion<T> & a, const Motion<T> & b) { Motion<T> r; r.lin = a.lin - b.lin; r.ang = a.ang - b.ang; return r; }
template<class T> BRBD_DI Force<T> operator+(const Force<T> & a, const Force<T> & b) { Force<T> r; r.lin = a.lin + b.lin; r.ang = a.ang + b.ang; return r; }
template<class T> BRBD_DI void operator+=(Motion<T> & a, const Motion<T> & b) { a.lin += b.lin; a.ang += b.ang; }
template<class T> BRBD_DI void operator+=(Force<T> & a, const Force<T> & b) { a.lin += b.lin; a.ang += b.ang; }
template<class T> BRBD_DI T dot6(const Motion<T> & m, const Force<T> & f) { return dot(m.lin, f.lin) + dot(m.ang, f.ang); }

// m1 x m2 — motion-dense.hpp:222-227
template<class T> BRBD_DI Motion<T> mcross(const Motion<T> & v, const Motion<T> & m)
{
  Motion<T> r;
  r.lin = cross(v.lin, m.ang) + cross(v.ang, m.lin);
  r.ang = cross(v.ang, m.ang);
  return r;
}
// v x* f — force-dense.hpp:177-182
template<class T> BRBD_DI Force<T> fcross(const Motion<T> & v, const Force<T> & f)
{
  Force<T> r;
  r.lin = cross(v.ang, f.lin);
  r.ang = cross(v.ang, f.ang) + cross(v.lin, f.lin);
  return r;
}

template<class T> struct SE3
{
  Mat3<T> R;
  Vec3<T> p;
  // se3-tpl.hpp:314-317
  BRBD_DI SE3 operator*(const SE3 & b) const { SE3 r; r.R = R * b.R; r.p = p + R * b.p; return r; }
  // motion-dense.hpp:258-263
  BRBD_DI Motion<T> act(const Motion<T> & m) const
  {
    Motion<T> r;
    r.ang = R * m.ang;
    r.lin = R * m.lin + cross(p, r.ang);
    return r;
  }
  // motion-dense.hpp:273-279
  BRBD_DI Motion<T> actInv(const Motion<T> & m) const
  {
    Motion<T> r;
    r.lin = tmul(R, m.lin - cross(p, m.ang));
    r.ang = tmul(R, m.ang);
    return r;
  }
  // force-dense.hpp:213-219
  BRBD_DI Force<T> act(const Force<T> & f) const
  {
    Force<T> r;
    r.lin = R * f.lin;
    r.ang = R * f.ang + cross(p, r.lin);
    return r;
  }
};

// packed symmetric 3x3 — symmetric3.hpp
template<class T> struct Sym3
{
  T xx, xy, yy, xz, yz, zz;
  BRBD_DI static Sym3 zero() { Sym3 s; s.xx = s.xy = s.yy = s.xz = s.yz = s.zz = T(0); return s; }
  // symmetric3.hpp:490-503
  BRBD_DI Vec3<T> mul(const Vec3<T> & v) const
  {
    return Vec3<T>(xx * v.x + xy * v.y + xz * v.z, xy * v.x + yy * v.y + yz * v.z, xz * v.x + yz * v.y + zz * v.z);
  }
  // R S R^T, the reference's factorisation — symmetric3.hpp:561-601
  BRBD_DI Sym3 rotate(const Mat3<T> & R) const
  {
    const T L00 = xx - zz, L01 = xy, L10 = xy, L11 = yy - zz, L20 = T(2) * xz, L21 = yz + yz;
    // Y = R.block<2,3>(1,0) * L ;  R(i,j) = col(j).component(i)
    const T Y00 = R.c0.y * L00 + R.c1.y * L10 + R.c2.y * L20;
    const T Y01 = R.c0.y * L01 + R.c1.y * L11 + R.c2.y * L21;
    const T Y10 = R.c0.z * L00 + R.c1.z * L10 + R.c2.z * L20;
    const T Y11 = R.c0.z * L01 + R.c1.z * L11 + R.c2.z * L21;
    Sym3 r;
    r.xy = Y00 * R.c0.x + Y01 * R.c1.x;
    r.yy = Y00 * R.c0.y + Y01 * R.c1.y;
    r.xz = Y10 * R.c0.x + Y11 * R.c1.x;
    r.yz = Y10 * R.c0.y + Y11 * R.c1.y;
    r.zz = Y10 * R.c0.z + Y11 * R.c1.z;
    const T r0 = -R.c0.x * yz + R.c1.x * xz;
    const T r1 = -R.c0.y * yz + R.c1.y * xz;
    const T r2 = -R.c0.z * yz + R.c1.z * xz;
    r.xx = L00 + L11 - r.yy - r.zz;
    r.xx += zz;
    r.xy += r2;
    r.yy += zz;
    r.xz -= r1;
    r.yz += r0;
    r.zz += zz;
    return r;
  }
};

// rigid-body inertia (m, c, I_c) — inertia.hpp:286-291
template<class T> struct Inertia
{
  T m;
  Vec3<T> c;
  Sym3<T> I;
  BRBD_DI static Inertia zero() { Inertia y; y.m = T(0); y.c = Vec3<T>::zero(); y.I = Sym3<T>::zero(); return y; }
  // inertia.hpp:728-735
  BRBD_DI Force<T> operator*(const Motion<T> & v) const
  {
    Force<T> f;
    f.lin = m * (v.lin - cross(c, v.ang));
    f.ang = I.mul(v.ang) + cross(c, f.lin);
    return f;
  }
  // inertia.hpp:659-673 (__pequ__), Symmetric3 -= AlphaSkewSquare symmetric3.hpp:259-271
  BRBD_DI void operator+=(const Inertia & b)
  {
    const T mab = m + b.m;
    const T eps = sizeof(T) == 8 ? T(2.220446049250313e-16) : T(1.1920929e-07);
    const T mab_inv = T(1) / max_t(mab, eps);
    const Vec3<T> AB = c - b.c;
    const T k = m * b.m * mab_inv;
    c = (m * mab_inv) * c + (b.m * mab_inv) * b.c;
    const T x = AB.x, y = AB.y, z = AB.z;
    I.xx = I.xx + b.I.xx + k * (y * y + z * z);
    I.xy = I.xy + b.I.xy - k * x * y;
    I.yy = I.yy + b.I.yy + k * (x * x + z * z);
    I.xz = I.xz + b.I.xz - k * x * z;
    I.yz = I.yz + b.I.yz - k * y * z;
    I.zz = I.zz + b.I.zz + k * (x * x + y * y);
    m = mab;
  }
};
// X . I — inertia.hpp:872-880
template<class T> BRBD_DI Inertia<T> act(const SE3<T> & M, const Inertia<T> & Y)
{
  Inertia<T> r;
  r.m = Y.m;
  r.c = M.p + M.R * Y.c;
  r.I = Y.I.rotate(M.R);
  return r;
}

// Symmetric 6x6 (articulated-body inertia), blocks: LL (sym), LA (3x3 general: rows linear, cols angular), AA (sym).
// The reference keeps oYaba as a dense 6x6 (data.oYaba, aba.hxx:134); the matrix stays symmetric through
// Ia -= U Dinv U^T and the parent accumulation, so 21 unique numbers suffice.
template<class T> struct SymMat6
{
  T a[21]; // packed upper triangle, row-major: (r,c), c>=r -> idx(r,c)
  BRBD_DI static int idx(int r, int c) { return r * 6 - (r * (r - 1)) / 2 + (c - r); }
  BRBD_DI T get(int r, int c) const { return r <= c ? a[idx(r, c)] : a[idx(c, r)]; }
};

// Eigen Quaternion::toRotationMatrix (free-flyer / spherical calc, joint-free-flyer.hpp:342-351)
template<class T> BRBD_DI Mat3<T> quat_to_mat(T x, T y, T z, T w)
{
  const T tx = T(2) * x, ty = T(2) * y, tz = T(2) * z;
  const T twx = tx * w, twy = ty * w, twz = tz * w;
  const T txx = tx * x, txy = ty * x, txz = tz * x;
  const T tyy = ty * y, tyz = tz * y, tzz = tz * z;
  Mat3<T> R;
  R.c0 = Vec3<T>(T(1) - (tyy + tzz), txy + twz, txz - twy);
  R.c1 = Vec3<T>(txy - twz, T(1) - (txx + tzz), tyz + twx);
  R.c2 = Vec3<T>(txz + twy, tyz - twx, T(1) - (txx + tyy));
  return R;
}

BRBD_DI void sincos_t(double x, double * s, double * c) { sincos(x, s, c); }
BRBD_DI void sincos_t(float x, float * s, float * c) { sincosf(x, s, c); }
BRBD_DI double sqrt_t(double x) { return sqrt(x); }
BRBD_DI float sqrt_t(float x) { return sqrtf(x); }

} // namespace brbd
