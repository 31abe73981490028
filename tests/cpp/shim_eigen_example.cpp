// The Eigen-typed overloads of include/pinocchio_b200_shim.hpp (the reference's own parameter lists:
// rneaInParallel(num_threads, pool, q, v, a, tau), parallel/rnea.hpp:31-45; abaInParallel, parallel/aba.hpp:32-46) compiled and
// run.  Eigen is not in this image, so the test builds this file against tests/cpp/eigen_stub (a stand-in for the few members
// of Eigen::MatrixBase the shim touches); with the real Eigen on the include path the same file compiles unchanged.
// Prints "OK" on a GPU box, "NO_GPU: ..." (exit 0) without one.
#include <Eigen/Core>

#include <cmath>
#include <cstdio>

#include "pinocchio_b200_shim.hpp"

#ifndef PINOCCHIO_B200_WITH_EIGEN
#error "the shim did not see <Eigen/Core>"
#endif

namespace pb = pinocchio_b200;

static const char * kUrdf =
  "<robot name='arm'>"
  "<link name='l0'/>"
  "<link name='l1'><inertial><origin xyz='0 0 0.1'/><mass value='1.5'/><inertia ixx='0.02' ixy='0' ixz='0' iyy='0.03' iyz='0' izz='0.01'/></inertial></link>"
  "<link name='l2'><inertial><origin xyz='0.1 0 0'/><mass value='0.7'/><inertia ixx='0.01' ixy='0' ixz='0' iyy='0.01' iyz='0' izz='0.02'/></inertial></link>"
  "<link name='l3'><inertial><origin xyz='0 0.05 0'/><mass value='0.4'/><inertia ixx='0.004' ixy='0' ixz='0' iyy='0.003' iyz='0' izz='0.005'/></inertial></link>"
  "<joint name='j1' type='revolute'><parent link='l0'/><child link='l1'/><origin xyz='0 0 0.2'/><axis xyz='0 0 1'/><limit lower='-3' upper='3'/></joint>"
  "<joint name='j2' type='continuous'><parent link='l1'/><child link='l2'/><origin xyz='0 0 0.3' rpy='0.1 0 0'/><axis xyz='0 1 0'/></joint>"
  "<joint name='j3' type='prismatic'><parent link='l2'/><child link='l3'/><origin xyz='0.2 0 0'/><axis xyz='1 0 0'/><limit lower='-1' upper='1'/></joint>"
  "</robot>";

int main()
{
  try
  {
    std::unique_ptr<pb::DeviceModelPool> pool(pb::DeviceModelPool::fromUrdf(kUrdf));
    const int nq = pool->nq(), nv = pool->nv(), B = 41; // nq = 1 + 2 + 1 (the continuous joint is (cos, sin)), nv = 3
    if (nq != 4 || nv != 3) { std::printf("FAIL nq %d nv %d\n", nq, nv); return 1; }
    if (pool->getModels().size() != pool->size() || pool->getModel(0).njoints != 4 || pool->getData(0).device != 0) { std::printf("FAIL pool surface\n"); return 1; }
    Eigen::MatrixXd q(nq, B), v(nv, B), a(nv, B), tau(nv, B), a2(nv, B), M(nv * nv, B), Minv(nv * nv, B), nle(nv, B), g(nv, B), qn(nq, B), vn(nv, B);
    Eigen::MatrixXd dq(nv * nv, B), dv(nv * nv, B), da(nv * nv, B), adq(nv * nv, B), adv(nv * nv, B), adt(nv * nv, B);
    unsigned s = 99u;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (double)(s >> 8) / (double)(1u << 24) * 2.0 - 1.0; };
    for (int c = 0; c < B; ++c)
    {
      const double ang = 3.0 * rnd();
      q(0, c) = rnd(); q(1, c) = std::cos(ang); q(2, c) = std::sin(ang); q(3, c) = rnd();
      for (int r = 0; r < nv; ++r) { v(r, c) = rnd(); a(r, c) = rnd(); }
    }
    pb::rneaInParallel(1, *pool, q, v, a, tau);
    pb::abaInParallel(1, *pool, q, v, tau, a2);
    double err = 0;
    for (int c = 0; c < B; ++c) for (int r = 0; r < nv; ++r) err = std::fmax(err, std::fabs(a2(r, c) - a(r, c)));
    if (!(err < 1e-10)) { std::printf("FAIL aba(rnea(a)) != a, %.3e\n", err); return 1; }
    pb::crbaInParallel(1, *pool, q, M);
    pb::computeMinverseInParallel(1, *pool, q, Minv);
    pb::nonLinearEffectsInParallel(1, *pool, q, v, nle);
    pb::computeGeneralizedGravityInParallel(1, *pool, q, g);
    pb::computeRNEADerivativesInParallel(1, *pool, q, v, a, dq, dv, da);
    pb::computeABADerivativesInParallel(1, *pool, q, v, tau, adq, adv, adt);
    pb::abaEulerStepInParallel(1, *pool, q, v, tau, 1e-3, qn, vn);
    // tau == M a + nle (unittest/rnea.cpp:195-206); dtau_da == M upper (rnea-derivatives.cpp:294-299); ddq_dtau == Minv on the upper part
    for (int c = 0; c < B; ++c)
      for (int r = 0; r < nv; ++r)
      {
        double acc = nle(r, c);
        for (int k = 0; k < nv; ++k) acc += (r <= k ? M(k * nv + r, c) : M(r * nv + k, c)) * a(k, c);
        err = std::fmax(err, std::fabs(acc - tau(r, c)));
        for (int k = r; k < nv; ++k)
        {
          err = std::fmax(err, std::fabs(da(k * nv + r, c) - M(k * nv + r, c)));
          err = std::fmax(err, std::fabs(adt(k * nv + r, c) - Minv(k * nv + r, c)));
        }
      }
    if (!(err < 1e-9)) { std::printf("FAIL identities, %.3e\n", err); return 1; }
    { // the packed output format: P(k, c) == M(rows[k] + nv * cols[k], c), and M has nothing outside the pattern
      std::vector<int32_t> rows, cols;
      pool->crbaPattern(rows, cols);
      Eigen::MatrixXd P((long)rows.size(), B);
      pb::crbaPackedInParallel(1, *pool, q, P);
      std::vector<char> in_pattern((size_t)nv * nv, 0);
      for (size_t k = 0; k < rows.size(); ++k)
      {
        in_pattern[(size_t)cols[k] * nv + rows[k]] = 1;
        for (int c = 0; c < B; ++c) err = std::fmax(err, std::fabs(P((long)k, c) - M(cols[k] * nv + rows[k], c)));
      }
      for (int e = 0; e < nv * nv; ++e)
        if (!in_pattern[e])
          for (int c = 0; c < B; ++c) err = std::fmax(err, std::fabs(M(e, c)));
      if (!(err < 1e-12)) { std::printf("FAIL packed crba, %.3e\n", err); return 1; }
      // and back to the dense layout on the host: exactly M's pattern entries, zeros elsewhere
      Eigen::MatrixXd M2(nv * nv, B);
      pb::expandPackedCrba(2, *pool, pb::ConstMatrixView{P.data(), (int64_t)P.rows(), (int64_t)P.cols(), (int64_t)P.outerStride()},
                           pb::MatrixView{M2.data(), (int64_t)M2.rows(), (int64_t)M2.cols(), (int64_t)M2.outerStride()});
      for (int e = 0; e < nv * nv; ++e)
        for (int c = 0; c < B; ++c)
          if (in_pattern[e] ? std::fabs(M2(e, c) - M(e, c)) > 1e-12 : M2(e, c) != 0.0) { std::printf("FAIL expandPackedCrba\n"); return 1; }
    }
    // a block of columns with the parent's outer stride, as q.middleCols(5, 7) would be
    Eigen::MatrixXd tau2(nv, B);
    Eigen::ColsBlock qb(q, 5, 7), vb(v, 5, 7), ab(a, 5, 7), tb(tau2, 5, 7);
    pb::rneaInParallel(1, *pool, qb, vb, ab, tb);
    for (int c = 5; c < 12; ++c) for (int r = 0; r < nv; ++r) if (tau2(r, c) != tau(r, c)) { std::printf("FAIL block view\n"); return 1; }
    bool threw = false;
    Eigen::MatrixXd bad(nv + 1, B);
    try { pb::rneaInParallel(1, *pool, q, v, a, bad); } catch (const std::invalid_argument &) { threw = true; }
    if (!threw) { std::printf("FAIL wrong-size argument did not throw\n"); return 1; }
    std::printf("OK eigen overloads, max err %.2e\n", err);
    return 0;
  }
  catch (const std::runtime_error & e)
  {
    std::printf("NO_GPU: %s\n", e.what());
    return 0;
  }
}
