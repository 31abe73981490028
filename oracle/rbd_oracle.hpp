// oracle/rbd_oracle.hpp
//
// TEST INFRASTRUCTURE — NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may build, link or call anything in oracle/.
//
// CPU restatement of the reference algorithms of pinocchio 3.7.0 for the batched-dynamics hot
// path.  The reference itself cannot be compiled in this environment (no Eigen, no Boost), so
// every function below follows the cited reference file:line, statement by statement, on plain
// fixed-size arrays.  Paths are relative to /root/reference/include/pinocchio/.
//
// PARITY STATUS: "parity unpinned" by golden vectors — the reference ships none for this path
// (unittest/rnea.cpp:5-9).  The oracle is pinned by the reference's own identity tests re-run
// in tests/test_oracle_identities.py.
//
// The joint variant dispatch (boost::apply_visitor) is restated with a dense motion subspace
// S (6 x nv_j) per joint type, so one generic code path covers RX/RY/RZ, PX/PY/PZ, FreeFlyer,
// Spherical and Planar; the revolute / prismatic placement shortcuts of the reference are kept.
//
// Templated on Scalar: double (the oracle), long double (independent higher-precision guard)
// and Counted (exact algorithmic FLOP counts for the roofline numerators).
#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

#include "../include/pinocchio_b200.h"

namespace rbdo
{

// ---------------------------------------------------------------------------------------------
// Counting scalar: mul/add/sub/div/sqrt = 1 flop each; sincos counted as calls.
// ---------------------------------------------------------------------------------------------
struct FlopCounter
{
  uint64_t add = 0, mul = 0, div = 0, sqrt_ = 0, sincos = 0;
  uint64_t flops() const { return add + mul + div + sqrt_; }
};
inline FlopCounter & flop_counter()
{
  static thread_local FlopCounter c;
  return c;
}
// `kind` marks the STRUCTURAL entries of a joint's motion subspace (0 and 1 of S): the reference's specialised joints never
// multiply by them (joint-revolute.hpp:290-340: S^T f = f.angular[axis], S x = one component) and neither do the kernels, so
// a product with a structural 0 / 1 and a sum with a structural 0 are not counted.  Data that merely happens to be 0 or 1
// (an identity rotation in a placement) is NOT marked: the reference multiplies by it.
struct Counted
{
  double x;
  unsigned char kind; // 0 general, 1 structural zero, 2 structural one
  Counted() : x(0), kind(0) {}
  Counted(double v) : x(v), kind(0) {}
  Counted(double v, unsigned char k) : x(v), kind(k) {}
  explicit operator double() const { return x; }
};
inline Counted operator+(Counted a, Counted b)
{
  if (a.kind == 1) return Counted(b.x);
  if (b.kind == 1) return Counted(a.x);
  flop_counter().add++;
  return Counted(a.x + b.x);
}
inline Counted operator-(Counted a, Counted b)
{
  if (b.kind == 1) return Counted(a.x);
  if (a.kind == 1) return Counted(-b.x);
  flop_counter().add++;
  return Counted(a.x - b.x);
}
inline Counted operator*(Counted a, Counted b)
{
  if (a.kind == 1 || b.kind == 1) return Counted(0.0, 1);
  if (a.kind == 2) return b;
  if (b.kind == 2) return a;
  flop_counter().mul++;
  return Counted(a.x * b.x);
}
inline Counted operator/(Counted a, Counted b) { flop_counter().div++; return Counted(a.x / b.x); }
inline Counted operator-(Counted a) { return Counted(-a.x, a.kind == 1 ? 1 : 0); }
inline Counted & operator+=(Counted & a, Counted b) { a = a + b; return a; }
inline Counted & operator-=(Counted & a, Counted b) { a = a - b; return a; }
inline Counted & operator*=(Counted & a, Counted b) { a = a * b; return a; }
inline bool operator<(Counted a, Counted b) { return a.x < b.x; }
inline bool operator>(Counted a, Counted b) { return a.x > b.x; }

inline void sincos_s(double x, double & s, double & c) { ::sincos(x, &s, &c); }
inline void sincos_s(long double x, long double & s, long double & c) { ::sincosl(x, &s, &c); }
inline void sincos_s(float x, float & s, float & c) { ::sincosf(x, &s, &c); }
inline void sincos_s(Counted x, Counted & s, Counted & c)
{
  flop_counter().sincos++;
  ::sincos(x.x, &s.x, &c.x);
}
#ifdef ORACLE_WITH_QUAD
// IEEE binary128 (libquadmath): the second, wider independent guard of the double-precision oracle
inline void sincos_s(__float128 x, __float128 & s, __float128 & c) { ::sincosq(x, &s, &c); }
inline __float128 sqrt_s(__float128 x) { return sqrtq(x); }
#endif
inline double sqrt_s(double x) { return std::sqrt(x); }
inline long double sqrt_s(long double x) { return sqrtl(x); }
inline float sqrt_s(float x) { return std::sqrt(x); }
inline Counted sqrt_s(Counted x) { flop_counter().sqrt_++; return Counted(std::sqrt(x.x)); }
template<class S> inline S eps_s() { return std::numeric_limits<S>::epsilon(); }
template<> inline Counted eps_s<Counted>() { return Counted(std::numeric_limits<double>::epsilon()); }
#ifdef ORACLE_WITH_QUAD
template<> inline __float128 eps_s<__float128>() { return FLT128_EPSILON; }
#endif
template<class S> inline S max_s(S a, S b) { return (a < b) ? b : a; }
// structural entries of a motion subspace (see Counted)
template<class S> inline S structural_zero() { return S(0); }
template<class S> inline S structural_one() { return S(1); }
template<> inline Counted structural_zero<Counted>() { return Counted(0.0, 1); }
template<> inline Counted structural_one<Counted>() { return Counted(1.0, 2); }
template<class S> inline double to_double(S x) { return (double)x; }

// ---------------------------------------------------------------------------------------------
// Fixed-size algebra
// ---------------------------------------------------------------------------------------------
template<class S> struct V3
{
  S v[3];
  V3() { v[0] = v[1] = v[2] = S(0); }
  V3(S a, S b, S c) { v[0] = a; v[1] = b; v[2] = c; }
  S & operator[](int i) { return v[i]; }
  const S & operator[](int i) const { return v[i]; }
};
template<class S> inline V3<S> operator+(const V3<S> & a, const V3<S> & b) { return V3<S>(a[0] + b[0], a[1] + b[1], a[2] + b[2]); }
template<class S> inline V3<S> operator-(const V3<S> & a, const V3<S> & b) { return V3<S>(a[0] - b[0], a[1] - b[1], a[2] - b[2]); }
template<class S> inline V3<S> operator-(const V3<S> & a) { return V3<S>(-a[0], -a[1], -a[2]); }
template<class S> inline V3<S> operator*(S s, const V3<S> & a) { return V3<S>(s * a[0], s * a[1], s * a[2]); }
template<class S> inline V3<S> & operator+=(V3<S> & a, const V3<S> & b) { a = a + b; return a; }
template<class S> inline V3<S> & operator-=(V3<S> & a, const V3<S> & b) { a = a - b; return a; }
template<class S> inline S dot(const V3<S> & a, const V3<S> & b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
template<class S> inline V3<S> cross(const V3<S> & a, const V3<S> & b)
{
  return V3<S>(a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]);
}

template<class S> struct M3
{
  S m[3][3]; // m[row][col]
  M3() { for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) m[i][j] = S(0); }
  static M3 Identity() { M3 r; r.m[0][0] = r.m[1][1] = r.m[2][2] = S(1); return r; }
  S & operator()(int i, int j) { return m[i][j]; }
  const S & operator()(int i, int j) const { return m[i][j]; }
  V3<S> col(int j) const { return V3<S>(m[0][j], m[1][j], m[2][j]); }
  void setCol(int j, const V3<S> & c) { m[0][j] = c[0]; m[1][j] = c[1]; m[2][j] = c[2]; }
  M3 transpose() const { M3 r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i][j] = m[j][i]; return r; }
};
template<class S> inline V3<S> operator*(const M3<S> & A, const V3<S> & x)
{
  return V3<S>(A(0, 0) * x[0] + A(0, 1) * x[1] + A(0, 2) * x[2],
               A(1, 0) * x[0] + A(1, 1) * x[1] + A(1, 2) * x[2],
               A(2, 0) * x[0] + A(2, 1) * x[1] + A(2, 2) * x[2]);
}
template<class S> inline V3<S> tmul(const M3<S> & A, const V3<S> & x) // A^T x
{
  return V3<S>(A(0, 0) * x[0] + A(1, 0) * x[1] + A(2, 0) * x[2],
               A(0, 1) * x[0] + A(1, 1) * x[1] + A(2, 1) * x[2],
               A(0, 2) * x[0] + A(1, 2) * x[1] + A(2, 2) * x[2]);
}
template<class S> inline M3<S> operator*(const M3<S> & A, const M3<S> & B)
{
  M3<S> r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      r.m[i][j] = A(i, 0) * B(0, j) + A(i, 1) * B(1, j) + A(i, 2) * B(2, j);
  return r;
}
template<class S> inline M3<S> operator+(const M3<S> & A, const M3<S> & B)
{
  M3<S> r;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i][j] = A(i, j) + B(i, j);
  return r;
}
template<class S> inline M3<S> operator-(const M3<S> & A, const M3<S> & B)
{
  M3<S> r;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i][j] = A(i, j) - B(i, j);
  return r;
}
template<class S> inline M3<S> operator-(const M3<S> & A)
{
  M3<S> r;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i][j] = -A(i, j);
  return r;
}

// spatial/skew.hpp:22-41
template<class S> inline M3<S> skew(const V3<S> & v)
{
  M3<S> M;
  M(0, 1) = -v[2]; M(0, 2) = v[1];
  M(1, 0) = v[2];  M(1, 2) = -v[0];
  M(2, 0) = -v[1]; M(2, 1) = v[0];
  return M;
}
// spatial/skew.hpp:68-84
template<class S> inline void addSkew(const V3<S> & v, M3<S> & M)
{
  M(0, 1) -= v[2]; M(0, 2) += v[1];
  M(1, 0) += v[2]; M(1, 2) -= v[0];
  M(2, 0) -= v[1]; M(2, 1) += v[0];
}
// spatial/skew.hpp:134-154
template<class S> inline M3<S> alphaSkew(S alpha, const V3<S> & v)
{
  M3<S> M;
  M(0, 1) = -v[2] * alpha;
  M(0, 2) = v[1] * alpha;
  M(1, 0) = -M(0, 1);
  M(1, 2) = -v[0] * alpha;
  M(2, 0) = -M(0, 2);
  M(2, 1) = -M(1, 2);
  return M;
}
// spatial/skew.hpp:182-197 : C = v u^T - (u.v) 1
template<class S> inline M3<S> skewSquare(const V3<S> & u, const V3<S> & v)
{
  M3<S> C;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) C(i, j) = v[i] * u[j];
  const S udotv = dot(u, v);
  for (int i = 0; i < 3; ++i) C(i, i) -= udotv;
  return C;
}
// spatial/skew.hpp:228-245 : Mout = [v]x Min
template<class S> inline M3<S> crossM(const V3<S> & v, const M3<S> & Min)
{
  M3<S> Mout;
  for (int j = 0; j < 3; ++j)
  {
    Mout(0, j) = v[1] * Min(2, j) - v[2] * Min(1, j);
    Mout(1, j) = v[2] * Min(0, j) - v[0] * Min(2, j);
    Mout(2, j) = v[0] * Min(1, j) - v[1] * Min(0, j);
  }
  return Mout;
}

// spatial/symmetric3.hpp:46-51 packing [xx, xy, yy, xz, yz, zz]
template<class S> struct Sym3
{
  S d[6];
  Sym3() { for (int i = 0; i < 6; ++i) d[i] = S(0); }
  Sym3(S a, S b, S c, S dd, S e, S f) { d[0] = a; d[1] = b; d[2] = c; d[3] = dd; d[4] = e; d[5] = f; }
  // symmetric3.hpp:304-318
  M3<S> matrix() const
  {
    M3<S> r;
    r(0, 0) = d[0]; r(0, 1) = d[1]; r(0, 2) = d[3];
    r(1, 0) = d[1]; r(1, 1) = d[2]; r(1, 2) = d[4];
    r(2, 0) = d[3]; r(2, 1) = d[4]; r(2, 2) = d[5];
    return r;
  }
  // symmetric3.hpp:490-503
  V3<S> rhsMult(const V3<S> & vin) const
  {
    return V3<S>(d[0] * vin[0] + d[1] * vin[1] + d[3] * vin[2],
                 d[1] * vin[0] + d[2] * vin[1] + d[4] * vin[2],
                 d[3] * vin[0] + d[4] * vin[1] + d[5] * vin[2]);
  }
  // symmetric3.hpp:561-601 (R S R^T in the reference's 28-multiplication factorisation)
  Sym3 rotate(const M3<S> & R) const
  {
    Sym3 r;
    // decomposeltI, symmetric3.hpp:552-558
    S L[3][2];
    L[0][0] = d[0] - d[5]; L[0][1] = d[1];
    L[1][0] = d[1];        L[1][1] = d[2] - d[5];
    L[2][0] = S(2) * d[3]; L[2][1] = d[4] + d[4];
    // Y = R.block<2,3>(1,0) * L
    S Y[2][2];
    for (int a = 0; a < 2; ++a)
      for (int b = 0; b < 2; ++b)
        Y[a][b] = R(1 + a, 0) * L[0][b] + R(1 + a, 1) * L[1][b] + R(1 + a, 2) * L[2][b];
    r.d[1] = Y[0][0] * R(0, 0) + Y[0][1] * R(0, 1);
    r.d[2] = Y[0][0] * R(1, 0) + Y[0][1] * R(1, 1);
    r.d[3] = Y[1][0] * R(0, 0) + Y[1][1] * R(0, 1);
    r.d[4] = Y[1][0] * R(1, 0) + Y[1][1] * R(1, 1);
    r.d[5] = Y[1][0] * R(2, 0) + Y[1][1] * R(2, 1);
    const V3<S> rr(-R(0, 0) * d[4] + R(0, 1) * d[3], -R(1, 0) * d[4] + R(1, 1) * d[3],
                   -R(2, 0) * d[4] + R(2, 1) * d[3]);
    r.d[0] = L[0][0] + L[1][1] - r.d[2] - r.d[5];
    r.d[0] += d[5];
    r.d[1] += rr[2];
    r.d[2] += d[5];
    r.d[3] -= rr[1];
    r.d[4] += rr[0];
    r.d[5] += d[5];
    return r;
  }
};
template<class S> inline Sym3<S> operator+(const Sym3<S> & a, const Sym3<S> & b)
{
  Sym3<S> r;
  for (int i = 0; i < 6; ++i) r.d[i] = a.d[i] + b.d[i];
  return r;
}
// symmetric3.hpp:188-220 : S -= alpha * SkewSquare(v)  /  S - AlphaSkewSquare(m, v) (259-271)
template<class S> inline Sym3<S> minusAlphaSkewSquare(const Sym3<S> & a, S m, const V3<S> & v)
{
  const S x = v[0], y = v[1], z = v[2];
  return Sym3<S>(a.d[0] + m * (y * y + z * z), a.d[1] - m * x * y, a.d[2] + m * (x * x + z * z),
                 a.d[3] - m * x * z, a.d[4] - m * y * z, a.d[5] + m * (x * x + y * y));
}

template<class S> struct Motion { V3<S> lin, ang; };
template<class S> struct Force { V3<S> lin, ang; };
template<class S> inline Motion<S> operator+(const Motion<S> & a, const Motion<S> & b) { return Motion<S>{a.lin + b.lin, a.ang + b.ang}; }
template<class S> inline Motion<S> operator-(const Motion<S> & a, const Motion<S> & b) { return Motion<S>{a.lin - b.lin, a.ang - b.ang}; }
template<class S> inline Force<S> operator+(const Force<S> & a, const Force<S> & b) { return Force<S>{a.lin + b.lin, a.ang + b.ang}; }
template<class S> inline Motion<S> & operator+=(Motion<S> & a, const Motion<S> & b) { a = a + b; return a; }
template<class S> inline Force<S> & operator+=(Force<S> & a, const Force<S> & b) { a = a + b; return a; }

// m1 ^ m2 : spatial/motion-dense.hpp:222-227  (m2.motionAction(m1))
template<class S> inline Motion<S> mcross(const Motion<S> & v, const Motion<S> & m)
{
  return Motion<S>{cross(v.lin, m.ang) + cross(v.ang, m.lin), cross(v.ang, m.ang)};
}
// v x* f : spatial/force-dense.hpp:177-182 (f.motionAction(v))
template<class S> inline Force<S> fcross(const Motion<S> & v, const Force<S> & f)
{
  Force<S> r;
  r.lin = cross(v.ang, f.lin);
  r.ang = cross(v.ang, f.ang) + cross(v.lin, f.lin);
  return r;
}

template<class S> struct SE3
{
  M3<S> R;
  V3<S> p;
  static SE3 Identity() { SE3 r; r.R = M3<S>::Identity(); return r; }
  // spatial/se3-tpl.hpp:314-317
  SE3 operator*(const SE3 & m2) const { return SE3{R * m2.R, p + R * m2.p}; }
  // spatial/motion-dense.hpp:258-263
  Motion<S> act(const Motion<S> & m) const
  {
    Motion<S> r;
    r.ang = R * m.ang;
    r.lin = R * m.lin + cross(p, r.ang);
    return r;
  }
  // spatial/motion-dense.hpp:273-279
  Motion<S> actInv(const Motion<S> & m) const
  {
    Motion<S> r;
    r.lin = tmul(R, m.lin - cross(p, m.ang));
    r.ang = tmul(R, m.ang);
    return r;
  }
  // spatial/force-dense.hpp:213-219
  Force<S> act(const Force<S> & f) const
  {
    Force<S> r;
    r.lin = R * f.lin;
    r.ang = R * f.ang + cross(p, r.lin);
    return r;
  }
  // spatial/force-dense.hpp:229-235
  Force<S> actInv(const Force<S> & f) const
  {
    Force<S> r;
    r.lin = tmul(R, f.lin);
    r.ang = tmul(R, f.ang - cross(p, f.lin));
    return r;
  }
};

template<class S> struct Mat6
{
  S m[6][6];
  Mat6() { setZero(); }
  void setZero() { for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) m[i][j] = S(0); }
  S & operator()(int i, int j) { return m[i][j]; }
  const S & operator()(int i, int j) const { return m[i][j]; }
  void setBlock(int r, int c, const M3<S> & B) { for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) m[r + i][c + j] = B(i, j); }
  M3<S> block(int r, int c) const { M3<S> B; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) B(i, j) = m[r + i][c + j]; return B; }
  Mat6 & operator+=(const Mat6 & o) { for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) m[i][j] += o.m[i][j]; return *this; }
};
enum { LINEAR = 0, ANGULAR = 3 }; // spatial/force-tpl.hpp:27-28

template<class S> struct Inertia
{
  S mass;
  V3<S> c;
  Sym3<S> I;
  Inertia() : mass(0) {}
  Inertia(S m, const V3<S> & c_, const Sym3<S> & I_) : mass(m), c(c_), I(I_) {}
  // spatial/inertia.hpp:728-735
  Force<S> operator*(const Motion<S> & v) const
  {
    Force<S> f;
    f.lin = mass * (v.lin - cross(c, v.ang));
    f.ang = I.rhsMult(v.ang);
    f.ang += cross(c, f.lin);
    return f;
  }
  // spatial/inertia.hpp:659-673
  Inertia & operator+=(const Inertia & Yb)
  {
    const S eps = eps_s<S>();
    const S mab = mass + Yb.mass;
    const S mab_inv = S(1) / max_s(mab, eps);
    const V3<S> AB = c - Yb.c;
    const S ma = mass;
    c = (mass * mab_inv) * c;
    c += (Yb.mass * mab_inv) * Yb.c;
    I = I + Yb.I;
    // inertia() -= (ma*mb*mab_inv) * SkewSquare(AB): Symmetric3::operator-=(AlphaSkewSquare),
    // symmetric3.hpp:259-271
    I = minusAlphaSkewSquare(I, ma * Yb.mass * mab_inv, AB);
    mass = mab;
    return *this;
  }
  // spatial/inertia.hpp:480-491
  Mat6<S> matrix() const
  {
    Mat6<S> M;
    for (int i = 0; i < 3; ++i) M(LINEAR + i, LINEAR + i) = mass;
    const M3<S> AL = alphaSkew(mass, c);
    M.setBlock(ANGULAR, LINEAR, AL);
    M.setBlock(LINEAR, ANGULAR, -AL);
    M.setBlock(ANGULAR, ANGULAR, minusAlphaSkewSquare(I, mass, c).matrix());
    return M;
  }
  // spatial/inertia.hpp:749-776
  Mat6<S> variation(const Motion<S> & v) const
  {
    Mat6<S> res;
    const Motion<S> mv{mass * v.lin, mass * v.ang};
    const M3<S> LA = -skew(mv.lin) - skewSquare(mv.ang, c) + skewSquare(c, mv.ang);
    res.setBlock(LINEAR, ANGULAR, LA);
    res.setBlock(ANGULAR, LINEAR, LA.transpose());
    M3<S> AA = -skewSquare(mv.lin, c) - skewSquare(c, mv.lin);
    const M3<S> LL = minusAlphaSkewSquare(I, mass, c).matrix();
    AA = AA - LL * skew(v.ang);
    AA = AA + crossM(v.ang, LL);
    res.setBlock(ANGULAR, ANGULAR, AA);
    return res;
  }
};
// spatial/inertia.hpp:872-880
template<class S> inline Inertia<S> act(const SE3<S> & M, const Inertia<S> & Y)
{
  return Inertia<S>(Y.mass, M.p + M.R * Y.c, Y.I.rotate(M.R));
}

// 6 x n column-major block of spatial columns (Data::Matrix6x)
template<class S> struct Mat6x
{
  int n = 0;
  std::vector<S> d;
  void resize(int n_) { n = n_; d.assign((size_t)6 * n, S(0)); }
  void setZero() { std::fill(d.begin(), d.end(), S(0)); }
  S & operator()(int r, int c) { return d[(size_t)6 * c + r]; }
  const S & operator()(int r, int c) const { return d[(size_t)6 * c + r]; }
  Motion<S> motion(int c) const { return Motion<S>{V3<S>((*this)(0, c), (*this)(1, c), (*this)(2, c)), V3<S>((*this)(3, c), (*this)(4, c), (*this)(5, c))}; }
  Force<S> force(int c) const { return Force<S>{V3<S>((*this)(0, c), (*this)(1, c), (*this)(2, c)), V3<S>((*this)(3, c), (*this)(4, c), (*this)(5, c))}; }
  void set(int c, const Motion<S> & m) { for (int k = 0; k < 3; ++k) { (*this)(k, c) = m.lin[k]; (*this)(3 + k, c) = m.ang[k]; } }
  void set(int c, const Force<S> & f) { for (int k = 0; k < 3; ++k) { (*this)(k, c) = f.lin[k]; (*this)(3 + k, c) = f.ang[k]; } }
  void add(int c, const Force<S> & f) { for (int k = 0; k < 3; ++k) { (*this)(k, c) += f.lin[k]; (*this)(3 + k, c) += f.ang[k]; } }
  void add(int c, const Motion<S> & f) { for (int k = 0; k < 3; ++k) { (*this)(k, c) += f.lin[k]; (*this)(3 + k, c) += f.ang[k]; } }
};
template<class S> inline void toVec(const Motion<S> & m, S * o) { for (int k = 0; k < 3; ++k) { o[k] = m.lin[k]; o[3 + k] = m.ang[k]; } }
template<class S> inline void toVec(const Force<S> & m, S * o) { for (int k = 0; k < 3; ++k) { o[k] = m.lin[k]; o[3 + k] = m.ang[k]; } }
template<class S> inline Force<S> mul6(const Mat6<S> & A, const Motion<S> & v)
{
  S x[6], y[6];
  toVec(v, x);
  for (int i = 0; i < 6; ++i)
  {
    S acc = A(i, 0) * x[0];
    for (int j = 1; j < 6; ++j) acc += A(i, j) * x[j];
    y[i] = acc;
  }
  return Force<S>{V3<S>(y[0], y[1], y[2]), V3<S>(y[3], y[4], y[5])};
}

// ---------------------------------------------------------------------------------------------
// Model (constants in Scalar) and Data (workspaces): multibody/model.hpp:97-205, data.hxx:30-315
// ---------------------------------------------------------------------------------------------
// unbounded revolute joints (URDF "continuous"): joint-revolute-unbounded.hpp:121-235, joint-revolute-unbounded-unaligned.hpp:133-260;
// configuration (cos q, sin q), nq = 2, nv = 1
inline bool joint_unbounded(int t) { return t >= BRBD_JOINT_RUBX && t <= BRBD_JOINT_REVOLUTE_UNBOUNDED_UNALIGNED; }
inline bool joint_unaligned(int t)
{
  return t == BRBD_JOINT_REVOLUTE_UNALIGNED || t == BRBD_JOINT_PRISMATIC_UNALIGNED || t == BRBD_JOINT_REVOLUTE_UNBOUNDED_UNALIGNED;
}
inline int joint_nq(int t)
{
  if (joint_unbounded(t)) return 2;
  return (t <= BRBD_JOINT_PZ || joint_unaligned(t)) ? 1 : (t == BRBD_JOINT_FREEFLYER ? 7 : 4);
}
inline int joint_nv(int t) { return (t <= BRBD_JOINT_PZ || joint_unaligned(t) || joint_unbounded(t)) ? 1 : (t == BRBD_JOINT_FREEFLYER ? 6 : 3); }

template<class S> struct Model
{
  int njoints = 0, nq = 0, nv = 0;
  std::vector<int> parents, type, idx_q, idx_v, nvs;
  std::vector<SE3<S>> jointPlacements;
  std::vector<Inertia<S>> inertias;
  std::vector<S> armature;
  std::vector<V3<S>> axis; // unit axis of the unaligned joints (joint-revolute-unaligned.hpp, joint-prismatic-unaligned.hpp)
  Motion<S> gravity; // model.gravity, linear = (0,0,-9.81) by default (model.hxx:40)

  explicit Model(const brbd_flat_model & f)
  {
    njoints = f.njoints; nq = f.nq; nv = f.nv;
    parents.assign(f.parents, f.parents + njoints);
    type.assign(f.joint_type, f.joint_type + njoints);
    idx_q.assign(f.idx_q, f.idx_q + njoints);
    idx_v.assign(f.idx_v, f.idx_v + njoints);
    nvs.resize(njoints);
    jointPlacements.resize(njoints);
    inertias.resize(njoints);
    axis.resize(njoints);
    for (int i = 0; i < njoints; ++i)
    {
      nvs[i] = (i == 0) ? 0 : joint_nv(type[i]);
      if (f.axis) axis[i] = V3<S>(S(f.axis[3 * i]), S(f.axis[3 * i + 1]), S(f.axis[3 * i + 2]));
      const double * P = f.placement + 12 * i;
      for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) jointPlacements[i].R(r, c) = S(P[3 * r + c]);
      for (int k = 0; k < 3; ++k) jointPlacements[i].p[k] = S(P[9 + k]);
      const double * Y = f.inertia + 10 * i;
      inertias[i].mass = S(Y[0]);
      for (int k = 0; k < 3; ++k) inertias[i].c[k] = S(Y[1 + k]);
      for (int k = 0; k < 6; ++k) inertias[i].I.d[k] = S(Y[4 + k]);
    }
    armature.resize(nv);
    for (int k = 0; k < nv; ++k) armature[k] = S(f.armature[k]);
    gravity.lin = V3<S>(S(f.gravity[0]), S(f.gravity[1]), S(f.gravity[2]));
  }
};

// JointData: result of jmodel.calc() restated with a dense S.
template<class S> struct JointData
{
  int nvj = 0;
  SE3<S> M;       // joint transform (used for FF / spherical / planar; revolute & prismatic use shortcuts)
  Motion<S> v;    // joint velocity S*qdot
  S Sm[6][6];     // motion subspace columns, joint frame: Sm[row][col]
  S U[6][6], Dinv[6][6], UDinv[6][6], StU[6][6];
};

template<class S> struct Data
{
  std::vector<SE3<S>> liMi, oMi;
  std::vector<Motion<S>> v, a, a_gf, ov, oa, oa_gf;
  std::vector<Force<S>> f, h, of, oh;
  std::vector<Inertia<S>> Ycrb, oYcrb, oinertias;
  std::vector<Mat6<S>> oYaba, doYcrb;
  std::vector<JointData<S>> joints;
  std::vector<S> tau, u, ddq;
  std::vector<S> M;               // nv x nv col-major (data.M)
  std::vector<S> Minv;            // nv x nv, (r,c) -> r*nv + c
  std::vector<S> dtau_dq, dtau_dv;// nv x nv, (r,c) -> r*nv + c (row-major in the reference, data.hpp:401-416)
  Mat6x<S> J, dJ, dVdq, dAdq, dAdv, dFdq, dFdv, dFda, SDinv, Ag;
  std::vector<Mat6x<S>> Fcrb;
  std::vector<int> lastChild, nvSubtree, parents_fromRow;
  int nv;

  explicit Data(const Model<S> & model)
  {
    const int n = model.njoints;
    nv = model.nv;
    liMi.resize(n); oMi.resize(n);
    v.resize(n); a.resize(n); a_gf.resize(n); ov.resize(n); oa.resize(n); oa_gf.resize(n);
    f.resize(n); h.resize(n); of.resize(n); oh.resize(n);
    Ycrb.resize(n); oYcrb.resize(n); oinertias.resize(n);
    oYaba.resize(n); doYcrb.resize(n);
    joints.resize(n);
    tau.assign(nv, S(0)); u.assign(nv, S(0)); ddq.assign(nv, S(0));
    M.assign((size_t)nv * nv, S(0));
    Minv.assign((size_t)nv * nv, S(0));
    dtau_dq.assign((size_t)nv * nv, S(0));
    dtau_dv.assign((size_t)nv * nv, S(0));
    J.resize(nv); dJ.resize(nv); dVdq.resize(nv); dAdq.resize(nv); dAdv.resize(nv);
    dFdq.resize(nv); dFdv.resize(nv); dFda.resize(nv); SDinv.resize(nv); Ag.resize(nv);
    Fcrb.resize(n);
    for (int i = 0; i < n; ++i) Fcrb[i].resize(nv);
    // data.hxx:197-242 (no mimic joints)
    lastChild.assign(n, -1);
    nvSubtree.assign(n, 0);
    for (int i = n - 1; i >= 0; --i)
    {
      if (lastChild[i] == -1) lastChild[i] = i;
      const int parent = model.parents[i];
      lastChild[parent] = std::max(lastChild[i], lastChild[parent]);
      const int lc = lastChild[i];
      nvSubtree[i] = (lc == 0) ? 0 : model.idx_v[lc] + model.nvs[lc] - (i == 0 ? 0 : model.idx_v[i]);
    }
    // data.hxx:246-315
    parents_fromRow.assign(nv, -1);
    for (int j = 1; j < n; ++j)
    {
      const int parent = model.parents[j];
      const int idx_vj = model.idx_v[j];
      if (parent > 0) parents_fromRow[idx_vj] = model.idx_v[parent] + model.nvs[parent] - 1;
      else parents_fromRow[idx_vj] = -1;
      for (int row = 1; row < model.nvs[j]; ++row) parents_fromRow[idx_vj + row] = idx_vj + row - 1;
    }
  }
};

// Eigen Quaternion::matrix() == toRotationMatrix (Eigen 3.4 Geometry/Quaternion.h);
// called from joint-free-flyer.hpp:339,351 and joint-spherical.hpp:529. Not re-normalised.
template<class S> inline M3<S> quatToMatrix(S x, S y, S z, S w)
{
  const S tx = S(2) * x, ty = S(2) * y, tz = S(2) * z;
  const S twx = tx * w, twy = ty * w, twz = tz * w;
  const S txx = tx * x, txy = ty * x, txz = tz * x;
  const S tyy = ty * y, tyz = tz * y, tzz = tz * z;
  M3<S> R;
  R(0, 0) = S(1) - (tyy + tzz); R(0, 1) = txy - twz;          R(0, 2) = txz + twy;
  R(1, 0) = txy + twz;          R(1, 1) = S(1) - (txx + tzz); R(1, 2) = tyz - twx;
  R(2, 0) = txz - twy;          R(2, 1) = tyz + twx;          R(2, 2) = S(1) - (txx + tyy);
  return R;
}

// jmodel.calc(jdata, q, v) followed by liMi = jointPlacements[i] * jdata.M()
//   revolute : joint-revolute.hpp:791-820 ; prismatic: joint-prismatic.hpp:698-725
//   freeflyer: joint-free-flyer.hpp:342-374 ; spherical: joint-spherical.hpp:524-570 ; planar: joint-planar.hpp:600-647
template<class S>
inline void jointCalc(const Model<S> & model, int i, const S * q, const S * vq, JointData<S> & jd, SE3<S> & liMi)
{
  const int t = model.type[i];
  const SE3<S> & P = model.jointPlacements[i];
  jd.nvj = model.nvs[i];
  for (int r = 0; r < 6; ++r) for (int c = 0; c < 6; ++c) jd.Sm[r][c] = structural_zero<S>();
  jd.v = Motion<S>();
  const S * qj = q + model.idx_q[i];
  const S * vj = vq ? vq + model.idx_v[i] : nullptr;
  if (t <= BRBD_JOINT_RZ || (t >= BRBD_JOINT_RUBX && t <= BRBD_JOINT_RUBZ))
  {
    // data.M.setValues(sa, ca); jointPlacements[i] * jdata.M() converts the TransformRevolute to a
    // plain SE3 (operator PlainType, joint-revolute.hpp:112-122, _setRotation 199-232) and uses the
    // generic SE3 product (se3-base.hpp:138-141, se3-tpl.hpp:314-317). The se3action shortcut at
    // joint-revolute.hpp:123-156 has no caller.
    // JointModelRevoluteUnbounded::calc (joint-revolute-unbounded.hpp:154-162): ca = q[0], sa = q[1], no sincos
    const bool unbounded = t >= BRBD_JOINT_RUBX;
    const int axis = unbounded ? t - BRBD_JOINT_RUBX : t - BRBD_JOINT_RX;
    S sa, ca;
    if (unbounded) { ca = qj[0]; sa = qj[1]; }
    else sincos_s(qj[0], sa, ca);
    jd.M = SE3<S>::Identity();
    if (axis == 0) { jd.M.R(1, 1) = ca; jd.M.R(1, 2) = -sa; jd.M.R(2, 1) = sa; jd.M.R(2, 2) = ca; }
    else if (axis == 1) { jd.M.R(0, 0) = ca; jd.M.R(0, 2) = sa; jd.M.R(2, 0) = -sa; jd.M.R(2, 2) = ca; }
    else { jd.M.R(0, 0) = ca; jd.M.R(0, 1) = -sa; jd.M.R(1, 0) = sa; jd.M.R(1, 1) = ca; }
    jd.Sm[ANGULAR + axis][0] = structural_one<S>();
    if (vj) jd.v.ang[axis] = vj[0];
  }
  else if (t <= BRBD_JOINT_PZ)
  {
    // TransformPrismatic::plain(), joint-prismatic.hpp:247-254, then the generic SE3 product.
    const int axis = t - BRBD_JOINT_PX;
    jd.M = SE3<S>::Identity();
    jd.M.p[axis] = qj[0];
    jd.Sm[LINEAR + axis][0] = structural_one<S>();
    if (vj) jd.v.lin[axis] = vj[0];
  }
  else if (t == BRBD_JOINT_REVOLUTE_UNALIGNED || t == BRBD_JOINT_REVOLUTE_UNBOUNDED_UNALIGNED)
  {
    // JointModelRevoluteUnalignedTpl::calc, joint-revolute-unaligned.hpp:668-672: toRotationMatrix(axis, cos, sin) of
    // math/rotation.hpp:26-55 (Eigen's AngleAxis formula); S = (0, axis), v = axis qdot (:84-103)
    // (unbounded: ca = q[0], sa = q[1], joint-revolute-unbounded-unaligned.hpp:192-200)
    const V3<S> & ax = model.axis[i];
    S sa, ca;
    if (t == BRBD_JOINT_REVOLUTE_UNBOUNDED_UNALIGNED) { ca = qj[0]; sa = qj[1]; }
    else sincos_s(qj[0], sa, ca);
    const V3<S> sin_axis = sa * ax, cos1_axis = (S(1) - ca) * ax;
    jd.M = SE3<S>::Identity();
    S tmp = cos1_axis[0] * ax[1];
    jd.M.R(0, 1) = tmp - sin_axis[2]; jd.M.R(1, 0) = tmp + sin_axis[2];
    tmp = cos1_axis[0] * ax[2];
    jd.M.R(0, 2) = tmp + sin_axis[1]; jd.M.R(2, 0) = tmp - sin_axis[1];
    tmp = cos1_axis[1] * ax[2];
    jd.M.R(1, 2) = tmp - sin_axis[0]; jd.M.R(2, 1) = tmp + sin_axis[0];
    for (int k = 0; k < 3; ++k) jd.M.R(k, k) = cos1_axis[k] * ax[k] + ca;
    for (int k = 0; k < 3; ++k) jd.Sm[ANGULAR + k][0] = ax[k];
    if (vj) jd.v.ang = vj[0] * ax;
  }
  else if (t == BRBD_JOINT_PRISMATIC_UNALIGNED)
  {
    // JointModelPrismaticUnalignedTpl::calc, joint-prismatic-unaligned.hpp: translation = axis q; S = (axis, 0)
    const V3<S> & ax = model.axis[i];
    jd.M = SE3<S>::Identity();
    jd.M.p = qj[0] * ax;
    for (int k = 0; k < 3; ++k) jd.Sm[LINEAR + k][0] = ax[k];
    if (vj) jd.v.lin = vj[0] * ax;
  }
  else
  if (t == BRBD_JOINT_FREEFLYER)
  {
    jd.M.p = V3<S>(qj[0], qj[1], qj[2]);
    jd.M.R = quatToMatrix(qj[3], qj[4], qj[5], qj[6]);
    for (int k = 0; k < 6; ++k) jd.Sm[k][k] = structural_one<S>();
    if (vj) { jd.v.lin = V3<S>(vj[0], vj[1], vj[2]); jd.v.ang = V3<S>(vj[3], vj[4], vj[5]); }
  }
  else if (t == BRBD_JOINT_SPHERICAL)
  {
    jd.M.p = V3<S>();
    jd.M.R = quatToMatrix(qj[0], qj[1], qj[2], qj[3]);
    for (int k = 0; k < 3; ++k) jd.Sm[ANGULAR + k][k] = structural_one<S>();
    if (vj) jd.v.ang = V3<S>(vj[0], vj[1], vj[2]);
  }
  else // planar
  {
    const S c = qj[2], s = qj[3];
    jd.M.R = M3<S>::Identity();
    jd.M.R(0, 0) = c; jd.M.R(0, 1) = -s; jd.M.R(1, 0) = s; jd.M.R(1, 1) = c;
    jd.M.p = V3<S>(qj[0], qj[1], S(0));
    jd.Sm[0][0] = structural_one<S>(); jd.Sm[1][1] = structural_one<S>(); jd.Sm[5][2] = structural_one<S>();
    if (vj) { jd.v.lin = V3<S>(vj[0], vj[1], S(0)); jd.v.ang = V3<S>(S(0), S(0), vj[2]); }
  }
  liMi = P * jd.M;
}

template<class S> inline Motion<S> Scol(const JointData<S> & jd, int k)
{
  return Motion<S>{V3<S>(jd.Sm[0][k], jd.Sm[1][k], jd.Sm[2][k]), V3<S>(jd.Sm[3][k], jd.Sm[4][k], jd.Sm[5][k])};
}
// jdata.S() * x
template<class S> inline Motion<S> Stimes(const JointData<S> & jd, const S * x)
{
  Motion<S> r;
  for (int k = 0; k < jd.nvj; ++k)
  {
    const Motion<S> s = Scol(jd, k);
    r.lin += x[k] * s.lin;
    r.ang += x[k] * s.ang;
  }
  return r;
}
template<class S> inline S dot6(const Motion<S> & m, const Force<S> & f) { return dot(m.lin, f.lin) + dot(m.ang, f.ang); }
template<class S> inline S dot6(const Force<S> & f, const Motion<S> & m) { return dot(m.lin, f.lin) + dot(m.ang, f.ang); }
template<class S> inline S dot6ff(const Force<S> & a, const Force<S> & b) { return dot(a.lin, b.lin) + dot(a.ang, b.ang); }

// ---------------------------------------------------------------------------------------------
// RNEA — algorithm/rnea.hxx:45-79 (forward), 92-107 (backward), 117-161 (driver)
// ---------------------------------------------------------------------------------------------
template<class S>
void rnea(const Model<S> & model, Data<S> & data, const S * q, const S * v, const S * a)
{
  for (int k = 0; k < model.nv; ++k) data.tau[k] = S(0);
  data.v[0] = Motion<S>();
  data.a_gf[0] = Motion<S>{-model.gravity.lin, -model.gravity.ang};
  for (int i = 1; i < model.njoints; ++i)
  {
    JointData<S> & jd = data.joints[i];
    const int parent = model.parents[i];
    jointCalc(model, i, q, v, jd, data.liMi[i]);
    data.v[i] = jd.v;
    if (parent > 0) data.v[i] += data.liMi[i].actInv(data.v[parent]);
    data.a_gf[i] = mcross(data.v[i], jd.v); // jdata.c() == 0 for all supported joints
    data.a_gf[i] += Stimes(jd, a + model.idx_v[i]);
    data.a_gf[i] += data.liMi[i].actInv(data.a_gf[parent]);
    data.h[i] = model.inertias[i] * data.v[i];
    data.f[i] = model.inertias[i] * data.a_gf[i];
    data.f[i] += fcross(data.v[i], data.h[i]);
  }
  for (int i = model.njoints - 1; i > 0; --i)
  {
    const JointData<S> & jd = data.joints[i];
    const int parent = model.parents[i];
    for (int k = 0; k < jd.nvj; ++k) data.tau[model.idx_v[i] + k] += dot6(Scol(jd, k), data.f[i]);
    if (parent > 0) data.f[parent] += data.liMi[i].act(data.f[i]);
  }
  for (int k = 0; k < model.nv; ++k) data.tau[k] += model.armature[k] * a[k];
}

// PerformStYSInversion — multibody/joint/joint-common-operations.hpp:23-33:
// Dinv = I; StYS.llt().solveInPlace(Dinv)
template<class S> inline void lltInverse(int n, const S A[6][6], S Ainv[6][6])
{
  S L[6][6];
  for (int i = 0; i < n; ++i)
    for (int j = 0; j <= i; ++j)
    {
      S s = A[i][j];
      for (int k = 0; k < j; ++k) s -= L[i][k] * L[j][k];
      if (i == j) L[i][i] = sqrt_s(s);
      else L[i][j] = s / L[j][j];
    }
  for (int c = 0; c < n; ++c)
  {
    S y[6];
    for (int i = 0; i < n; ++i)
    {
      S s = (i == c) ? S(1) : S(0);
      for (int k = 0; k < i; ++k) s -= L[i][k] * y[k];
      y[i] = s / L[i][i];
    }
    for (int i = n - 1; i >= 0; --i)
    {
      S s = y[i];
      for (int k = i + 1; k < n; ++k) s -= L[k][i] * Ainv[k][c];
      Ainv[i][c] = s / L[i][i];
    }
  }
}

// Shared by aba (aba.hxx:101-138) and computeABADerivatives pass 1 (aba-derivatives.hxx:38-79)
template<class S>
inline void abaWorldForward1(const Model<S> & model, Data<S> & data, int i, const S * q, const S * v)
{
  JointData<S> & jd = data.joints[i];
  const int parent = model.parents[i];
  jointCalc(model, i, q, v, jd, data.liMi[i]);
  if (parent > 0) data.oMi[i] = data.oMi[parent] * data.liMi[i];
  else data.oMi[i] = data.liMi[i];
  for (int k = 0; k < jd.nvj; ++k) data.J.set(model.idx_v[i] + k, data.oMi[i].act(Scol(jd, k)));
  Motion<S> & ov = data.ov[i];
  ov = data.oMi[i].act(jd.v);
  if (parent > 0) ov += data.ov[parent];
  data.oa_gf[i] = Motion<S>(); // oMi.act(jdata.c()) with c == 0
  if (parent > 0) data.oa_gf[i] += mcross(data.ov[parent], ov);
  data.oinertias[i] = data.oYcrb[i] = act(data.oMi[i], model.inertias[i]);
  data.oYaba[i] = data.oYcrb[i].matrix();
  data.oh[i] = data.oYcrb[i] * ov;
  data.of[i] = fcross(ov, data.oh[i]);
}

// Core of AbaWorldConventionBackwardStep (aba.hxx:152-192), shared with
// ComputeABADerivativesBackwardStep1 (aba-derivatives.hxx:124-132, 161-169).
template<class S> inline void abaWorldBackwardCore(const Model<S> & model, Data<S> & data, int i, bool articulate)
{
  JointData<S> & jd = data.joints[i];
  const int parent = model.parents[i];
  const int iv = model.idx_v[i], nvj = jd.nvj;
  Mat6<S> & Ia = data.oYaba[i];
  Force<S> & fi = data.of[i];
  S Jc[6][6]; // Jcols[row][k]
  for (int k = 0; k < nvj; ++k) for (int r = 0; r < 6; ++r) Jc[r][k] = data.J(r, iv + k);
  if (!articulate)
  {
    for (int k = 0; k < nvj; ++k) data.u[iv + k] -= dot6(data.J.motion(iv + k), fi);
    for (int r = 0; r < 6; ++r)
      for (int k = 0; k < nvj; ++k)
      {
        S acc = Ia(r, 0) * Jc[0][k];
        for (int c = 1; c < 6; ++c) acc += Ia(r, c) * Jc[c][k];
        jd.U[r][k] = acc;
      }
    for (int a = 0; a < nvj; ++a)
      for (int b = 0; b < nvj; ++b)
      {
        S acc = Jc[0][a] * jd.U[0][b];
        for (int r = 1; r < 6; ++r) acc += Jc[r][a] * jd.U[r][b];
        jd.StU[a][b] = acc;
      }
    for (int k = 0; k < nvj; ++k) jd.StU[k][k] += model.armature[iv + k];
    lltInverse(nvj, jd.StU, jd.Dinv);
    for (int r = 0; r < 6; ++r)
      for (int k = 0; k < nvj; ++k)
      {
        S acc = jd.U[r][0] * jd.Dinv[0][k];
        for (int c = 1; c < nvj; ++c) acc += jd.U[r][c] * jd.Dinv[c][k];
        jd.UDinv[r][k] = acc;
      }
    return;
  }
  if (parent > 0)
  {
    for (int r = 0; r < 6; ++r)
      for (int c = 0; c < 6; ++c)
      {
        S acc = jd.UDinv[r][0] * jd.U[c][0];
        for (int k = 1; k < nvj; ++k) acc += jd.UDinv[r][k] * jd.U[c][k];
        Ia(r, c) -= acc;
      }
    const Force<S> Iaa = mul6(Ia, data.oa_gf[i]);
    S fa[6], ud[6];
    toVec(Iaa, fa);
    for (int r = 0; r < 6; ++r)
    {
      S acc = jd.UDinv[r][0] * data.u[iv];
      for (int k = 1; k < nvj; ++k) acc += jd.UDinv[r][k] * data.u[iv + k];
      ud[r] = acc;
    }
    for (int k = 0; k < 3; ++k) { fi.lin[k] += fa[k] + ud[k]; fi.ang[k] += fa[3 + k] + ud[3 + k]; }
    data.oYaba[parent] += Ia;
    data.of[parent] += fi;
  }
}

// AbaWorldConventionForwardStep2 core (aba.hxx:222-226) shared with aba-derivatives.hxx:209-214
template<class S> inline void abaWorldForward2Core(const Model<S> & model, Data<S> & data, int i)
{
  JointData<S> & jd = data.joints[i];
  const int parent = model.parents[i];
  const int iv = model.idx_v[i], nvj = jd.nvj;
  data.oa_gf[i] += data.oa_gf[parent];
  S ag[6];
  toVec(data.oa_gf[i], ag);
  for (int k = 0; k < nvj; ++k)
  {
    S t1 = jd.Dinv[k][0] * data.u[iv];
    for (int c = 1; c < nvj; ++c) t1 += jd.Dinv[k][c] * data.u[iv + c];
    S t2 = jd.UDinv[0][k] * ag[0];
    for (int r = 1; r < 6; ++r) t2 += jd.UDinv[r][k] * ag[r];
    data.ddq[iv + k] = t1 - t2;
  }
  for (int k = 0; k < nvj; ++k)
  {
    const Motion<S> Jk = data.J.motion(iv + k);
    data.oa_gf[i].lin += data.ddq[iv + k] * Jk.lin;
    data.oa_gf[i].ang += data.ddq[iv + k] * Jk.ang;
  }
}

// ABA, WORLD convention — algorithm/aba.hxx:242-293 (what abaInParallel calls, parallel/aba.hpp:82)
template<class S>
void abaWorld(const Model<S> & model, Data<S> & data, const S * q, const S * v, const S * tau)
{
  data.oa_gf[0] = Motion<S>{-model.gravity.lin, -model.gravity.ang};
  data.of[0] = Force<S>();
  for (int k = 0; k < model.nv; ++k) data.u[k] = tau[k];
  for (int i = 1; i < model.njoints; ++i) abaWorldForward1(model, data, i, q, v);
  for (int i = model.njoints - 1; i > 0; --i)
  {
    abaWorldBackwardCore(model, data, i, false);
    abaWorldBackwardCore(model, data, i, true);
  }
  for (int i = 1; i < model.njoints; ++i)
  {
    abaWorldForward2Core(model, data, i);
    // "consistent output" (aba.hxx:228-230)
    data.oa[i] = data.oa_gf[i] + model.gravity;
    data.of[i] = data.oinertias[i] * data.oa_gf[i] + fcross(data.ov[i], data.oh[i]);
  }
  for (int i = model.njoints - 1; i > 0; --i) data.of[model.parents[i]] += data.of[i];
}

// ---------------------------------------------------------------------------------------------
// CRBA — algorithm/crba.hxx
// ---------------------------------------------------------------------------------------------
// WORLD: 498-548 (driver), 35-58 (forward), 80-99 (backward)
template<class S> void crbaWorld(const Model<S> & model, Data<S> & data, const S * q)
{
  const int nv = model.nv;
  data.oYcrb[0] = Inertia<S>();
  for (int i = 1; i < model.njoints; ++i)
  {
    JointData<S> & jd = data.joints[i];
    const int parent = model.parents[i];
    jointCalc(model, i, q, (const S *)nullptr, jd, data.liMi[i]);
    if (parent > 0) data.oMi[i] = data.oMi[parent] * data.liMi[i];
    else data.oMi[i] = data.liMi[i];
    for (int k = 0; k < jd.nvj; ++k) data.J.set(model.idx_v[i] + k, data.oMi[i].act(Scol(jd, k)));
    data.oYcrb[i] = act(data.oMi[i], model.inertias[i]);
  }
  for (int i = model.njoints - 1; i > 0; --i)
  {
    const int iv = model.idx_v[i], nvj = model.nvs[i];
    for (int k = 0; k < nvj; ++k) data.Ag.set(iv + k, data.oYcrb[i] * data.J.motion(iv + k));
    for (int k = 0; k < nvj; ++k)
      for (int c = 0; c < data.nvSubtree[i]; ++c)
        data.M[(size_t)(iv + c) * nv + iv + k] = dot6(data.J.motion(iv + k), data.Ag.force(iv + c));
    data.oYcrb[model.parents[i]] += data.oYcrb[i];
  }
  for (int k = 0; k < nv; ++k) data.M[(size_t)k * nv + k] += model.armature[k];
  // centroidal fix-up of Ag (crba.hxx:539-545) is not part of M; kept for completeness
  const V3<S> com = data.oYcrb[0].c;
  for (int k = 0; k < nv; ++k)
  {
    Force<S> fk = data.Ag.force(k);
    fk.ang += cross(fk.lin, com);
    data.Ag.set(k, fk);
  }
}
// LOCAL (the default of crba(), crba.hpp:51): 454-491 (driver), 234-248 (forward), 273-309 (backward)
template<class S> void crbaLocal(const Model<S> & model, Data<S> & data, const S * q)
{
  const int nv = model.nv;
  for (int i = 1; i < model.njoints; ++i)
  {
    jointCalc(model, i, q, (const S *)nullptr, data.joints[i], data.liMi[i]);
    data.Ycrb[i] = model.inertias[i];
  }
  for (int i = model.njoints - 1; i > 0; --i)
  {
    const JointData<S> & jd = data.joints[i];
    const int iv = model.idx_v[i], nvj = jd.nvj, ns = data.nvSubtree[i];
    for (int k = 0; k < nvj; ++k) data.Fcrb[i].set(iv + k, data.Ycrb[i] * Scol(jd, k));
    for (int k = 0; k < nvj; ++k)
      for (int c = 0; c < ns; ++c)
        data.M[(size_t)(iv + c) * nv + iv + k] = dot6(Scol(jd, k), data.Fcrb[i].force(iv + c));
    const int parent = model.parents[i];
    if (parent > 0)
    {
      data.Ycrb[parent] += act(data.liMi[i], data.Ycrb[i]);
      for (int c = 0; c < ns; ++c) data.Fcrb[parent].set(iv + c, data.liMi[i].act(data.Fcrb[i].force(iv + c)));
    }
  }
  for (int k = 0; k < nv; ++k) data.M[(size_t)k * nv + k] += model.armature[k];
}

// ---------------------------------------------------------------------------------------------
// Shared derivative pieces
// ---------------------------------------------------------------------------------------------
// addForceCrossMatrix — rnea-derivatives.hxx:340-351
template<class S> inline void addForceCrossMatrix(const Force<S> & f, Mat6<S> & mout)
{
  M3<S> LA = mout.block(LINEAR, ANGULAR), AL = mout.block(ANGULAR, LINEAR), AA = mout.block(ANGULAR, ANGULAR);
  addSkew(-f.lin, LA);
  addSkew(-f.lin, AL);
  addSkew(-f.ang, AA);
  mout.setBlock(LINEAR, ANGULAR, LA);
  mout.setBlock(ANGULAR, LINEAR, AL);
  mout.setBlock(ANGULAR, ANGULAR, AA);
}
// dJ, dVdq, dAdq, dAdv columns — rnea-derivatives.hxx:319-332 == aba-derivatives.hxx:241-251
template<class S> inline void derivColumns(const Model<S> & model, Data<S> & data, int i)
{
  const int parent = model.parents[i];
  const int iv = model.idx_v[i], nvj = model.nvs[i];
  for (int k = 0; k < nvj; ++k)
  {
    const int c = iv + k;
    const Motion<S> Jk = data.J.motion(c);
    const Motion<S> dJk = mcross(data.ov[i], Jk);
    data.dJ.set(c, dJk);
    Motion<S> dAdq = mcross(data.oa_gf[parent], Jk);
    Motion<S> dAdv = dJk;
    if (parent > 0)
    {
      const Motion<S> dVdq = mcross(data.ov[parent], Jk);
      data.dVdq.set(c, dVdq);
      dAdq += mcross(data.ov[parent], dVdq);
      dAdv += dVdq;
    }
    else
      data.dVdq.set(c, Motion<S>());
    data.dAdq.set(c, dAdq);
    data.dAdv.set(c, dAdv);
  }
}
template<class S> inline Force<S> mul6m(const Mat6<S> & A, const Motion<S> & v) { return mul6(A, v); }

// ---------------------------------------------------------------------------------------------
// computeRNEADerivatives — algorithm/rnea-derivatives.hxx:263-338 (fwd), 378-459 (bwd), 472-541
// Outputs are nv x nv COLUMN-major and must be pre-zeroed by the caller (rnea-derivatives.hpp:104-106).
// ---------------------------------------------------------------------------------------------
template<class S>
void rneaDerivatives(const Model<S> & model, Data<S> & data, const S * q, const S * v, const S * a,
                     S * dq, S * dv, S * da)
{
  const int nv = model.nv;
#define AT(Mx, r, c) Mx[(size_t)(c) * nv + (r)]
  data.oa_gf[0] = Motion<S>{-model.gravity.lin, -model.gravity.ang};
  for (int i = 1; i < model.njoints; ++i)
  {
    JointData<S> & jd = data.joints[i];
    const int parent = model.parents[i];
    jointCalc(model, i, q, v, jd, data.liMi[i]);
    data.v[i] = jd.v;
    if (parent > 0)
    {
      data.oMi[i] = data.oMi[parent] * data.liMi[i];
      data.v[i] += data.liMi[i].actInv(data.v[parent]);
    }
    else
      data.oMi[i] = data.liMi[i];
    data.a[i] = Stimes(jd, a + model.idx_v[i]) + mcross(data.v[i], jd.v);
    if (parent > 0) data.a[i] += data.liMi[i].actInv(data.a[parent]);
    data.oYcrb[i] = data.oinertias[i] = act(data.oMi[i], model.inertias[i]);
    data.ov[i] = data.oMi[i].act(data.v[i]);
    data.oa[i] = data.oMi[i].act(data.a[i]);
    data.oa_gf[i] = data.oa[i] - model.gravity;
    data.oh[i] = data.oYcrb[i] * data.ov[i];
    data.of[i] = data.oYcrb[i] * data.oa_gf[i] + fcross(data.ov[i], data.oh[i]);
    for (int k = 0; k < jd.nvj; ++k) data.J.set(model.idx_v[i] + k, data.oMi[i].act(Scol(jd, k)));
    derivColumns(model, data, i);
    data.doYcrb[i] = data.oYcrb[i].variation(data.ov[i]);
    addForceCrossMatrix(data.oh[i], data.doYcrb[i]);
  }
  Mat6x<S> & dYtJ = data.Fcrb[0];
  for (int i = model.njoints - 1; i > 0; --i)
  {
    const int parent = model.parents[i];
    const int iv = model.idx_v[i], nvj = model.nvs[i];
    const int ns = data.nvSubtree[i];
    const int ivp = iv + nvj, nsp = ns - nvj;
    for (int k = 0; k < nvj; ++k) data.tau[iv + k] = dot6(data.J.motion(iv + k), data.of[i]);
    // dtau/da
    for (int k = 0; k < nvj; ++k) data.dFda.set(iv + k, data.oYcrb[i] * data.J.motion(iv + k));
    for (int k = 0; k < nvj; ++k)
      for (int c = 0; c < ns; ++c) AT(da, iv + k, iv + c) = dot6(data.J.motion(iv + k), data.dFda.force(iv + c));
    // dtau/dq
    for (int k = 0; k < nvj; ++k)
    {
      if (parent > 0)
      {
        Force<S> t = mul6m(data.doYcrb[i], data.dVdq.motion(iv + k));
        t += data.oYcrb[i] * data.dAdq.motion(iv + k);
        data.dFdq.set(iv + k, t);
      }
      else
        data.dFdq.set(iv + k, data.oYcrb[i] * data.dAdq.motion(iv + k));
    }
    // dYtJ_cols^T = J_cols^T * doYcrb  => dYtJ(:,k) = doYcrb^T J_k
    for (int k = 0; k < nvj; ++k)
      for (int r = 0; r < 6; ++r)
      {
        S acc = data.J(0, iv + k) * data.doYcrb[i](0, r);
        for (int c = 1; c < 6; ++c) acc += data.J(c, iv + k) * data.doYcrb[i](c, r);
        dYtJ(r, iv + k) = acc;
      }
    for (int r = 0; r < nsp; ++r)
      for (int k = 0; k < nvj; ++k)
        AT(dq, ivp + r, iv + k) = dot6(data.dFda.force(ivp + r), data.dAdq.motion(iv + k))
                                  + dot6(dYtJ.force(ivp + r), data.dVdq.motion(iv + k));
    for (int k = 0; k < nvj; ++k)
      for (int c = 0; c < ns; ++c) AT(dq, iv + k, iv + c) = dot6(data.J.motion(iv + k), data.dFdq.force(iv + c));
    for (int k = 0; k < nvj; ++k) data.dFdq.add(iv + k, fcross(data.J.motion(iv + k), data.of[i]));
    // dtau/dv
    for (int k = 0; k < nvj; ++k)
    {
      Force<S> t = mul6m(data.doYcrb[i], data.J.motion(iv + k));
      t += data.oYcrb[i] * data.dAdv.motion(iv + k);
      data.dFdv.set(iv + k, t);
    }
    for (int r = 0; r < nsp; ++r)
      for (int k = 0; k < nvj; ++k)
        AT(dv, ivp + r, iv + k) = dot6(data.dFda.force(ivp + r), data.dAdv.motion(iv + k))
                                  + dot6(dYtJ.force(ivp + r), data.J.motion(iv + k));
    for (int k = 0; k < nvj; ++k)
      for (int c = 0; c < ns; ++c) AT(dv, iv + k, iv + c) = dot6(data.J.motion(iv + k), data.dFdv.force(iv + c));
    if (parent > 0)
    {
      data.oYcrb[parent] += data.oYcrb[i];
      data.doYcrb[parent] += data.doYcrb[i];
      data.of[parent] += data.of[i];
    }
  }
  // restore dAdq (530-535) — internal only
  for (int k = 0; k < nv; ++k)
  {
    Motion<S> m = data.dAdq.motion(k);
    m.lin += cross(model.gravity.lin, data.J.motion(k).ang);
    data.dAdq.set(k, m);
  }
  for (int k = 0; k < nv; ++k) data.tau[k] += model.armature[k] * a[k];
  for (int k = 0; k < nv; ++k) AT(da, k, k) += model.armature[k];
#undef AT
}

// ---------------------------------------------------------------------------------------------
// computeABADerivatives — algorithm/aba-derivatives.hxx:38-79, 97-170, 188-256, 283-367, 380-453
// Outputs nv x nv COLUMN-major. dtau <- Minv (full symmetric), dq, dv full dense, data.ddq.
// ---------------------------------------------------------------------------------------------
template<class S>
void abaDerivatives(const Model<S> & model, Data<S> & data, const S * q, const S * v, const S * tau,
                    S * dq, S * dv, S * dtau)
{
  const int nv = model.nv;
#define MINV(r, c) data.Minv[(size_t)(r) * nv + (c)]
#define DTQ(r, c) data.dtau_dq[(size_t)(r) * nv + (c)]
#define DTV(r, c) data.dtau_dv[(size_t)(r) * nv + (c)]
  data.oa_gf[0] = Motion<S>{-model.gravity.lin, -model.gravity.ang};
  for (int k = 0; k < nv; ++k) data.u[k] = tau[k];
  // Minv_.triangularView<Upper>().setZero()
  for (int r = 0; r < nv; ++r) for (int c = r; c < nv; ++c) MINV(r, c) = S(0);
  // Data is created with zeros in dtau_dq / dtau_dv and only a fixed pattern is ever written
  // (data.hxx:97-99); the oracle re-zeroes so that one Data can serve several models' tests.
  std::fill(data.dtau_dq.begin(), data.dtau_dq.end(), S(0));
  std::fill(data.dtau_dv.begin(), data.dtau_dv.end(), S(0));
  data.of[0] = Force<S>();

  for (int i = 1; i < model.njoints; ++i) abaWorldForward1(model, data, i, q, v);

  Mat6x<S> & Fcrb0 = data.Fcrb[0];
  Fcrb0.setZero();
  for (int i = model.njoints - 1; i > 0; --i)
  {
    JointData<S> & jd = data.joints[i];
    const int parent = model.parents[i];
    const int iv = model.idx_v[i], nvj = jd.nvj, ns = data.nvSubtree[i];
    abaWorldBackwardCore(model, data, i, false);
    for (int a = 0; a < nvj; ++a) for (int b = 0; b < nvj; ++b) MINV(iv + a, iv + b) = jd.Dinv[a][b];
    const int nv_children = ns - nvj;
    if (nv_children > 0)
    {
      // SDinv_cols = J_cols * Dinv
      for (int r = 0; r < 6; ++r)
        for (int k = 0; k < nvj; ++k)
        {
          S acc = data.J(r, iv) * jd.Dinv[0][k];
          for (int c = 1; c < nvj; ++c) acc += data.J(r, iv + c) * jd.Dinv[c][k];
          data.SDinv(r, iv + k) = acc;
        }
      for (int k = 0; k < nvj; ++k)
        for (int c = 0; c < nv_children; ++c)
          MINV(iv + k, iv + nvj + c) = -dot6ff(data.SDinv.force(iv + k), Fcrb0.force(iv + nvj + c));
      if (parent > 0)
      {
        for (int c = 0; c < ns; ++c)
          for (int r = 0; r < 6; ++r)
          {
            S acc = jd.U[r][0] * MINV(iv, iv + c);
            for (int k = 1; k < nvj; ++k) acc += jd.U[r][k] * MINV(iv + k, iv + c);
            Fcrb0(r, iv + c) += acc;
          }
      }
    }
    else
    {
      for (int c = 0; c < ns; ++c)
        for (int r = 0; r < 6; ++r)
        {
          S acc = jd.U[r][0] * MINV(iv, iv + c);
          for (int k = 1; k < nvj; ++k) acc += jd.U[r][k] * MINV(iv + k, iv + c);
          Fcrb0(r, iv + c) = acc;
        }
    }
    abaWorldBackwardCore(model, data, i, true);
  }

  for (int i = 1; i < model.njoints; ++i)
  {
    JointData<S> & jd = data.joints[i];
    const int parent = model.parents[i];
    const int iv = model.idx_v[i], nvj = jd.nvj;
    abaWorldForward2Core(model, data, i);
    data.oa[i] = data.oa_gf[i] + model.gravity;
    data.of[i] = data.oYcrb[i] * data.oa_gf[i] + fcross(data.ov[i], data.oh[i]);
    const int nr = nv - iv; // rightCols(model.nv - idx_v)
    if (parent > 0)
    {
      for (int k = 0; k < nvj; ++k)
        for (int c = 0; c < nr; ++c)
        {
          S acc = jd.UDinv[0][k] * data.Fcrb[parent](0, iv + c);
          for (int r = 1; r < 6; ++r) acc += jd.UDinv[r][k] * data.Fcrb[parent](r, iv + c);
          MINV(iv + k, iv + c) -= acc;
        }
    }
    for (int c = 0; c < nr; ++c)
      for (int r = 0; r < 6; ++r)
      {
        S acc = data.J(r, iv) * MINV(iv, iv + c);
        for (int k = 1; k < nvj; ++k) acc += data.J(r, iv + k) * MINV(iv + k, iv + c);
        data.Fcrb[i](r, iv + c) = acc;
      }
    if (parent > 0)
      for (int c = 0; c < nr; ++c)
        for (int r = 0; r < 6; ++r) data.Fcrb[i](r, iv + c) += data.Fcrb[parent](r, iv + c);
    derivColumns(model, data, i);
    data.doYcrb[i] = data.oYcrb[i].variation(data.ov[i]);
    addForceCrossMatrix(data.oh[i], data.doYcrb[i]);
  }

  for (int i = model.njoints - 1; i > 0; --i)
  {
    const int parent = model.parents[i];
    const int iv = model.idx_v[i], nvj = model.nvs[i], ns = data.nvSubtree[i];
    // dtau/dv
    for (int k = 0; k < nvj; ++k)
    {
      Force<S> t = data.oYcrb[i] * data.dAdv.motion(iv + k);
      t += mul6m(data.doYcrb[i], data.J.motion(iv + k));
      data.dFdv.set(iv + k, t);
    }
    for (int k = 0; k < nvj; ++k)
      for (int c = 0; c < ns; ++c) DTV(iv + k, iv + c) = dot6(data.J.motion(iv + k), data.dFdv.force(iv + c));
    // dtau/dq
    for (int k = 0; k < nvj; ++k)
    {
      Force<S> t = data.oYcrb[i] * data.dAdq.motion(iv + k);
      if (parent > 0) t += mul6m(data.doYcrb[i], data.dVdq.motion(iv + k));
      data.dFdq.set(iv + k, t);
    }
    for (int k = 0; k < nvj; ++k)
      for (int c = 0; c < ns; ++c) DTQ(iv + k, iv + c) = dot6(data.J.motion(iv + k), data.dFdq.force(iv + c));
    for (int k = 0; k < nvj; ++k) data.dFdq.add(iv + k, fcross(data.J.motion(iv + k), data.of[i]));
    for (int k = 0; k < nvj; ++k) data.dFda.set(iv + k, data.oYcrb[i] * data.J.motion(iv + k));
    if (parent > 0)
    {
      for (int j = data.parents_fromRow[iv]; j >= 0; j = data.parents_fromRow[j])
        for (int k = 0; k < nvj; ++k) DTQ(iv + k, j) = dot6(data.dFda.force(iv + k), data.dAdq.motion(j));
      for (int j = data.parents_fromRow[iv]; j >= 0; j = data.parents_fromRow[j])
        for (int k = 0; k < nvj; ++k) DTV(iv + k, j) = dot6(data.dFda.force(iv + k), data.dAdv.motion(j));
      // StdY = J_cols^T * doYcrb (nvj x 6)
      S StdY[6][6];
      for (int k = 0; k < nvj; ++k)
        for (int r = 0; r < 6; ++r)
        {
          S acc = data.J(0, iv + k) * data.doYcrb[i](0, r);
          for (int c = 1; c < 6; ++c) acc += data.J(c, iv + k) * data.doYcrb[i](c, r);
          StdY[k][r] = acc;
        }
      for (int j = data.parents_fromRow[iv]; j >= 0; j = data.parents_fromRow[j])
        for (int k = 0; k < nvj; ++k)
        {
          S acc = StdY[k][0] * data.dVdq(0, j);
          for (int r = 1; r < 6; ++r) acc += StdY[k][r] * data.dVdq(r, j);
          DTQ(iv + k, j) += acc;
        }
      for (int j = data.parents_fromRow[iv]; j >= 0; j = data.parents_fromRow[j])
        for (int k = 0; k < nvj; ++k)
        {
          S acc = StdY[k][0] * data.J(0, j);
          for (int r = 1; r < 6; ++r) acc += StdY[k][r] * data.J(r, j);
          DTV(iv + k, j) += acc;
        }
      data.oYcrb[parent] += data.oYcrb[i];
      data.doYcrb[parent] += data.doYcrb[i];
      data.of[parent] += data.of[i];
    }
    for (int k = 0; k < nvj; ++k)
    {
      Motion<S> m = data.dAdq.motion(iv + k);
      m.lin += cross(model.gravity.lin, data.J.motion(iv + k).ang);
      data.dAdq.set(iv + k, m);
    }
  }
  // symmetrise Minv (448-449)
  for (int r = 0; r < nv; ++r) for (int c = 0; c < r; ++c) MINV(r, c) = MINV(c, r);
  // aba_partial_dq = -Minv * dtau_dq ; aba_partial_dv = -Minv * dtau_dv  (451-452)
  for (int c = 0; c < nv; ++c)
    for (int r = 0; r < nv; ++r)
    {
      S sq = S(0), sv = S(0);
      for (int k = 0; k < nv; ++k)
      {
        sq += (-MINV(r, k)) * DTQ(k, c);
        sv += (-MINV(r, k)) * DTV(k, c);
      }
      dq[(size_t)c * nv + r] = sq;
      dv[(size_t)c * nv + r] = sv;
      dtau[(size_t)c * nv + r] = MINV(r, c);
    }
#undef MINV
#undef DTQ
#undef DTV
}

// ---------------------------------------------------------------------------------------------
// integrate(model, q, v) — algorithm/joint-configuration.hpp:49-74 -> IntegrateStep -> per-joint Lie group
// integrate_impl (multibody/liegroup/liegroup-algo.hxx).  Quaternions are stored (x, y, z, w) as in the
// configuration vector; the Eigen operations of the reference (quaternion product, quaternion * vector,
// AngleAxis -> quaternion) are restated from their Eigen 3.4 definitions.
// ---------------------------------------------------------------------------------------------
template<class S> inline S taylor_precision3() { return S(0.0001220703125); } // pow(epsilon, 1/4), math/taylor-expansion.hpp:30-36 (2^-13 for double)
template<> inline float taylor_precision3<float>() { return std::pow(std::numeric_limits<float>::epsilon(), 0.25f); }
#ifdef ORACLE_WITH_QUAD
template<> inline __float128 taylor_precision3<__float128>() { return powq(FLT128_EPSILON, 0.25Q); }
#endif
template<> inline long double taylor_precision3<long double>() { return powl(std::numeric_limits<long double>::epsilon(), 0.25L); }

// Eigen::Quaternion operator* (Eigen/src/Geometry/Quaternion.h, quat_product)
template<class S> inline void quatMul(const S * a, const S * b, S * r)
{
  const S ax = a[0], ay = a[1], az = a[2], aw = a[3], bx = b[0], by = b[1], bz = b[2], bw = b[3];
  r[3] = aw * bw - ax * bx - ay * by - az * bz;
  r[0] = aw * bx + ax * bw + ay * bz - az * by;
  r[1] = aw * by + ay * bw + az * bx - ax * bz;
  r[2] = aw * bz + az * bw + ax * by - ay * bx;
}
// Eigen::QuaternionBase::_transformVector: v + w * (2 q x v) + q x (2 q x v)
template<class S> inline V3<S> quatRotate(const S * q, const V3<S> & v)
{
  const V3<S> qv(q[0], q[1], q[2]);
  V3<S> uv = cross(qv, v);
  uv += uv;
  return v + q[3] * uv + cross(qv, uv);
}
// quaternion::exp3 — spatial/explog-quaternion.hpp:25-64
template<class S> inline void quatExp3(const V3<S> & w, S * quat)
{
  const S eps = eps_s<S>();
  const S t2 = dot(w, w);
  const S t = sqrt_s(t2 + eps * eps);
  const S ts_prec = taylor_precision3<S>();
  if (t2 > ts_prec)
  {
    // Eigen::AngleAxis(t, w / t) -> Quaternion: w = cos(t/2), vec = sin(t/2) * axis
    S sh, ch;
    sincos_s(S(0.5) * t, sh, ch);
    for (int k = 0; k < 3; ++k) quat[k] = sh * (w[k] / t);
    quat[3] = ch;
  }
  else
  {
    const S t2_2 = t2 / S(4);
    const S a = S(0.5) * (S(1) - t2_2 / S(6) + t2_2 * t2_2 / S(120));
    for (int k = 0; k < 3; ++k) quat[k] = a * w[k];
    quat[3] = S(1) - t2_2 / S(2) + t2_2 * t2_2 / S(24);
  }
}
// quaternion::exp6 — spatial/explog-quaternion.hpp:92-136; out = (translation, quaternion)
template<class S> inline void quatExp6(const V3<S> & v, const V3<S> & w, V3<S> & trans, S * quat)
{
  const S eps = eps_s<S>();
  const S t2 = dot(w, w) + eps * eps;
  const S t = sqrt_s(t2);
  S st, ct;
  sincos_s(t, st, ct);
  const S inv_t2 = S(1) / t2;
  const S ts_prec = taylor_precision3<S>();
  const S alpha_wxv = (t < ts_prec) ? S(0.5) - t2 / S(24) : (S(1) - ct) * inv_t2;
  const S alpha_w2 = (t < ts_prec) ? S(1) / S(6) - t2 / S(120) : (t - st) * inv_t2 / t;
  const V3<S> wxv = cross(w, v);
  trans = v + alpha_wxv * wxv + alpha_w2 * cross(w, wxv);
  quatExp3(w, quat);
}
// quaternion::firstOrderNormalize — math/quaternion.hpp:90-111
template<class S> inline void quatFirstOrderNormalize(S * q)
{
  const S N2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
  const S alpha = (S(3) - N2) / S(2);
  for (int k = 0; k < 4; ++k) q[k] *= alpha;
}

template<class S> void integrate(const Model<S> & model, const S * q, const S * v, S * qout)
{
  for (int i = 1; i < model.njoints; ++i)
  {
    const S * qj = q + model.idx_q[i];
    const S * vj = v + model.idx_v[i];
    S * o = qout + model.idx_q[i];
    const int t = model.type[i];
    if (joint_unbounded(t))
    { // SpecialOrthogonalOperationTpl<2>::integrate_impl, liegroup/special-orthogonal.hpp:164-185
      const S ca = qj[0], sa = qj[1];
      S so, co;
      sincos_s(vj[0], so, co);
      o[0] = co * ca - so * sa;
      o[1] = so * ca + co * sa;
      const S norm2 = o[0] * o[0] + o[1] * o[1];
      const S k = (S(3) - norm2) / S(2);
      o[0] = o[0] * k;
      o[1] = o[1] * k;
    }
    else if (t <= BRBD_JOINT_PZ || joint_unaligned(t))
      o[0] = qj[0] + vj[0]; // VectorSpaceOperation::integrate_impl, liegroup/vector-space.hpp:142-150
    else if (t == BRBD_JOINT_FREEFLYER)
    {
      // SpecialEuclideanOperationTpl<3>::integrate_impl, liegroup/special-euclidean.hpp:660-698
      V3<S> trans;
      S quat1[4], res[4];
      quatExp6(V3<S>(vj[0], vj[1], vj[2]), V3<S>(vj[3], vj[4], vj[5]), trans, quat1);
      const V3<S> p = quatRotate(qj + 3, trans);
      for (int k = 0; k < 3; ++k) o[k] = p[k] + qj[k];
      quatMul(qj + 3, quat1, res);
      const S dp = res[0] * qj[3] + res[1] * qj[4] + res[2] * qj[5] + res[3] * qj[6];
      for (int k = 0; k < 4; ++k) res[k] = (dp < S(0)) ? -res[k] : res[k];
      quatFirstOrderNormalize(res);
      for (int k = 0; k < 4; ++k) o[3 + k] = res[k];
    }
    else if (t == BRBD_JOINT_SPHERICAL)
    {
      // SpecialOrthogonalOperationTpl<3>::integrate_impl, liegroup/special-orthogonal.hpp:467-481
      S pOmega[4], res[4];
      quatExp3(V3<S>(vj[0], vj[1], vj[2]), pOmega);
      quatMul(qj, pOmega, res);
      quatFirstOrderNormalize(res);
      for (int k = 0; k < 4; ++k) o[k] = res[k];
    }
    else
    {
      // planar: SpecialEuclideanOperationTpl<2>::integrate_impl (special-euclidean.hpp:289-306) with exp (:61-90);
      // q = (x, y, cos, sin), v = (vx, vy, omega)
      const S c0 = qj[2], s0 = qj[3];
      const S omega = vj[2];
      S sv, cv;
      sincos_s(omega, sv, cv);
      // vcross = (-v1, v0) - (-v1 R.col(0) + v0 R.col(1)), R = [[cv, -sv], [sv, cv]]
      S vc0 = -vj[1] - (-vj[1] * cv + vj[0] * (-sv));
      S vc1 = vj[0] - (-vj[1] * sv + vj[0] * cv);
      vc0 = vc0 / omega;
      vc1 = vc1 / omega;
      const S omega_abs = (omega < S(0)) ? -omega : omega;
      const S t0 = (omega_abs > S(1e-14)) ? vc0 : vj[0];
      const S t1 = (omega_abs > S(1e-14)) ? vc1 : vj[1];
      o[0] = (c0 * t0 - s0 * t1) + qj[0]; // R0 * t + t0
      o[1] = (s0 * t0 + c0 * t1) + qj[1];
      o[2] = c0 * cv - s0 * sv; // R0 * R.col(0)
      o[3] = s0 * cv + c0 * sv;
    }
  }
}

} // namespace rbdo
