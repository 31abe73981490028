// integrate.cuh — batched integrate(model, q, v) and the semi-implicit Euler step that follows ABA, so that a batched
// simulation never leaves the device.
//
// Restates integrate (reference: include/pinocchio/algorithm/joint-configuration.hpp:49-74) with the per-joint Lie
// group operations it dispatches to:
//   revolute / prismatic   VectorSpaceOperation::integrate_impl           multibody/liegroup/vector-space.hpp:142-150
//   free-flyer             SpecialEuclideanOperationTpl<3>::integrate_impl multibody/liegroup/special-euclidean.hpp:660-698
//                          quaternion::exp6 / exp3 / firstOrderNormalize   spatial/explog-quaternion.hpp:25-64, 92-136; math/quaternion.hpp:90-111
//   spherical              SpecialOrthogonalOperationTpl<3>::integrate_impl multibody/liegroup/special-orthogonal.hpp:467-481
//   planar                 SpecialEuclideanOperationTpl<2>::integrate_impl  multibody/liegroup/special-euclidean.hpp:289-306, exp :61-90
// Quaternions are (x, y, z, w) as in the configuration vector; Eigen's quaternion product, quaternion * vector and
// AngleAxis -> quaternion are restated from their definitions.
//
// One configuration per thread; a warp stages its 32 configurations through shared memory so that the caller's
// (column = configuration) blocks are read and written with coalesced accesses.  The kernels are HBM-bound:
// 8 (2 nq + nv) bytes per configuration for integrate, 8 (2 nq + 4 nv) for the Euler step.
#pragma once

#include "crba.cuh"
#include "engine.cuh"

namespace brbd
{

template<class T> BRBD_DI T taylor_precision3();
template<> BRBD_DI double taylor_precision3<double>() { return 0.0001220703125; } // pow(epsilon, 1/4) = 2^-13, math/taylor-expansion.hpp:30-36
template<> BRBD_DI float taylor_precision3<float>() { return 0.018581361f; }
template<class T> BRBD_DI T epsilon_t();
template<> BRBD_DI double epsilon_t<double>() { return 2.220446049250313e-16; }
template<> BRBD_DI float epsilon_t<float>() { return 1.1920929e-07f; }

template<class T> BRBD_DI void quat_mul(const T * a, const T * b, T * r)
{
  const T ax = a[0], ay = a[1], az = a[2], aw = a[3], bx = b[0], by = b[1], bz = b[2], bw = b[3];
  r[3] = aw * bw - ax * bx - ay * by - az * bz;
  r[0] = aw * bx + ax * bw + ay * bz - az * by;
  r[1] = aw * by + ay * bw + az * bx - ax * bz;
  r[2] = aw * bz + az * bw + ax * by - ay * bx;
}
template<class T> BRBD_DI Vec3<T> quat_rotate(const T * q, const Vec3<T> & v)
{
  const Vec3<T> qv(q[0], q[1], q[2]);
  Vec3<T> uv = cross(qv, v);
  uv += uv;
  return v + q[3] * uv + cross(qv, uv);
}
template<class T> BRBD_DI void quat_exp3(const Vec3<T> & w, T * quat)
{
  const T eps = epsilon_t<T>();
  const T t2 = dot(w, w);
  const T t = sqrt_t(t2 + eps * eps);
  if (t2 > taylor_precision3<T>())
  {
    T sh, ch;
    sincos_t(T(0.5) * t, &sh, &ch);
    quat[0] = sh * (w.x / t); quat[1] = sh * (w.y / t); quat[2] = sh * (w.z / t);
    quat[3] = ch;
  }
  else
  {
    const T t2_2 = t2 / T(4);
    const T a = T(0.5) * (T(1) - t2_2 / T(6) + t2_2 * t2_2 / T(120));
    quat[0] = a * w.x; quat[1] = a * w.y; quat[2] = a * w.z;
    quat[3] = T(1) - t2_2 / T(2) + t2_2 * t2_2 / T(24);
  }
}
template<class T> BRBD_DI void quat_exp6(const Vec3<T> & v, const Vec3<T> & w, Vec3<T> & trans, T * quat)
{
  const T eps = epsilon_t<T>();
  const T t2 = dot(w, w) + eps * eps;
  const T t = sqrt_t(t2);
  T st, ct;
  sincos_t(t, &st, &ct);
  const T inv_t2 = T(1) / t2;
  const bool small = t < taylor_precision3<T>();
  const T alpha_wxv = small ? T(0.5) - t2 / T(24) : (T(1) - ct) * inv_t2;
  const T alpha_w2 = small ? T(1) / T(6) - t2 / T(120) : (t - st) * inv_t2 / t;
  const Vec3<T> wxv = cross(w, v);
  trans = v + alpha_wxv * wxv + alpha_w2 * cross(w, wxv);
  quat_exp3(w, quat);
}
template<class T> BRBD_DI void quat_first_order_normalize(T * q)
{
  const T N2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
  const T alpha = (T(3) - N2) / T(2);
  q[0] *= alpha; q[1] *= alpha; q[2] *= alpha; q[3] *= alpha;
}

// one joint: qout_j = q_j (+) s * v_j   (s = 1 for integrate, dt for the Euler step)
template<class T> BRBD_DI void integrate_joint(int type, const T * qj, const T * vj, T s, T * o)
{
  if (type <= J_PZ) o[0] = qj[0] + s * vj[0];
  else if (type == J_FF)
  {
    Vec3<T> trans;
    T quat1[4], res[4], q0[4] = {qj[3], qj[4], qj[5], qj[6]};
    quat_exp6(Vec3<T>(s * vj[0], s * vj[1], s * vj[2]), Vec3<T>(s * vj[3], s * vj[4], s * vj[5]), trans, quat1);
    const Vec3<T> p = quat_rotate(q0, trans);
    o[0] = p.x + qj[0]; o[1] = p.y + qj[1]; o[2] = p.z + qj[2];
    quat_mul(q0, quat1, res);
    const T dp = res[0] * q0[0] + res[1] * q0[1] + res[2] * q0[2] + res[3] * q0[3];
    if (dp < T(0)) { res[0] = -res[0]; res[1] = -res[1]; res[2] = -res[2]; res[3] = -res[3]; }
    quat_first_order_normalize(res);
    o[3] = res[0]; o[4] = res[1]; o[5] = res[2]; o[6] = res[3];
  }
  else if (type == J_SPH)
  {
    T pOmega[4], res[4], q0[4] = {qj[0], qj[1], qj[2], qj[3]};
    quat_exp3(Vec3<T>(s * vj[0], s * vj[1], s * vj[2]), pOmega);
    quat_mul(q0, pOmega, res);
    quat_first_order_normalize(res);
    o[0] = res[0]; o[1] = res[1]; o[2] = res[2]; o[3] = res[3];
  }
  else
  {
    const T c0 = qj[2], s0 = qj[3];
    const T v0 = s * vj[0], v1 = s * vj[1], omega = s * vj[2];
    T sv, cv;
    sincos_t(omega, &sv, &cv);
    T vc0 = -v1 - (-v1 * cv + v0 * (-sv));
    T vc1 = v0 - (-v1 * sv + v0 * cv);
    vc0 = vc0 / omega;
    vc1 = vc1 / omega;
    const T omega_abs = omega < T(0) ? -omega : omega;
    const T t0 = omega_abs > T(1e-14) ? vc0 : v0;
    const T t1 = omega_abs > T(1e-14) ? vc1 : v1;
    o[0] = (c0 * t0 - s0 * t1) + qj[0];
    o[1] = (s0 * t0 + c0 * t1) + qj[1];
    o[2] = c0 * cv - s0 * sv;
    o[3] = s0 * cv + c0 * sv;
  }
}

// EULER = false: qout = integrate(q, v).
// EULER = true : v_next = v + dt a; q_next = integrate(q, dt v_next)  (semi-implicit Euler; `a` = the ABA result).
template<class T, bool EULER>
__global__ void __launch_bounds__(256)
integrate_kernel(const ModelPOD<T> * __restrict__ gmod, const T * __restrict__ q, int64_t ldq, const T * __restrict__ v, int64_t ldv,
                 const T * __restrict__ a, int64_t lda, T dt, T * __restrict__ qout, int64_t ldqo, T * __restrict__ vout, int64_t ldvo,
                 int64_t B)
{
  __shared__ ModelPOD<T> m;
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  copy_model_to_smem(&m, gmod);
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
    if (EULER) tile_load(sa, vpad, a + c0 * lda, lda, m.nv, nc, lane);
    BRBD_SYNCWARP();
    if (lane < nc)
    {
      T * ql = sq + lane * qpad, * vl = sv + lane * vpad;
      if (EULER)
      {
        const T * al = sa + lane * vpad;
        for (int k = 0; k < m.nv; ++k) vl[k] += dt * al[k];
      }
      for (int i = 1; i < m.njoints; ++i)
      {
        T o[7];
        const int type = m.type[i];
        T * qj = ql + m.idx_q[i];
        if (m.unb[i])
        { // SpecialOrthogonalOperationTpl<2>::integrate_impl (special-orthogonal.hpp:164-185): q = (cos, sin) turned by v,
          // then the first-order normalisation of the unit complex
          const T om = (EULER ? dt : T(1)) * vl[m.idx_v[i]];
          T so, co;
          sincos_t(om, &so, &co);
          const T c1 = co * qj[0] - so * qj[1], s1 = so * qj[0] + co * qj[1];
          const T n = (T(3) - (c1 * c1 + s1 * s1)) / T(2);
          o[0] = c1 * n; o[1] = s1 * n;
        }
        else
          integrate_joint(type, qj, vl + m.idx_v[i], EULER ? dt : T(1), o);
        const int nqj = m.unb[i] ? 2 : (type <= J_PZ ? 1 : (type == J_FF ? 7 : 4));
#pragma unroll
        for (int k = 0; k < 7; ++k)
          if (k < nqj) qj[k] = o[k];
      }
    }
    BRBD_SYNCWARP();
    tile_store(qout + c0 * ldqo, ldqo, sq, qpad, m.nq, nc, lane);
    if (EULER) tile_store(vout + c0 * ldvo, ldvo, sv, vpad, m.nv, nc, lane);
    BRBD_SYNCWARP();
  }
}

} // namespace brbd
