// Exercises include/pinocchio_b200_shim.hpp the way a caller of the reference's batched API would
// (benchmark/timings-parallel.cpp:141-158, unittest/parallel-rnea.cpp:21-55): build a pool, call
// rneaInParallel / abaInParallel on column-per-configuration matrices, check the reference's identity
// aba(q, v, rnea(q, v, a)) == a (unittest/aba.cpp:143-154) and the argument-size error behaviour.
// Prints "OK" on a GPU box, "NO_GPU: <message>" (exit 0) when no CUDA device exists — there is no CPU fallback.
#include <cmath>
#include <cstdio>
#include <vector>

#include "pinocchio_b200_shim.hpp"

namespace pb = pinocchio_b200;

int main()
{
  // a 6-revolute chain shaped like buildModels::manipulator (RX,RY,RZ,RY,RX,RY), made-up constants
  const int nj = 7, nq = 6, nv = 6;
  std::vector<int32_t> parents = {0, 0, 1, 2, 3, 4, 5}, types = {-1, 0, 1, 2, 1, 0, 1}, idx = {0, 0, 1, 2, 3, 4, 5};
  std::vector<double> placement(12 * nj, 0.0), inertia(10 * nj, 0.0), armature(nv, 0.01);
  for (int i = 0; i < nj; ++i)
  {
    double * P = &placement[12 * i];
    P[0] = P[4] = P[8] = 1.0;          // identity rotation (row-major)
    P[11] = (i > 1) ? 0.3 : 0.0;       // 0.3 m along z between joints
    double * Y = &inertia[10 * i];
    Y[0] = 1.0 + 0.1 * i;              // mass
    Y[1] = 0.01 * i; Y[2] = -0.02; Y[3] = 0.15; // lever
    Y[4] = 0.02; Y[6] = 0.03; Y[9] = 0.01;      // xx, yy, zz
  }
  brbd_flat_model flat;
  flat.njoints = nj; flat.nq = nq; flat.nv = nv;
  flat.parents = parents.data(); flat.joint_type = types.data(); flat.idx_q = idx.data(); flat.idx_v = idx.data();
  flat.placement = placement.data(); flat.inertia = inertia.data(); flat.armature = armature.data();
  flat.gravity[0] = 0; flat.gravity[1] = 0; flat.gravity[2] = -9.81;
  flat.axis = nullptr;               // no joint about an arbitrary axis in this model

  try
  {
    pb::DeviceModelPool pool(flat);
    const int B = 37;
    std::vector<double> q(nq * B), v(nv * B), a(nv * B), tau(nv * B), a2(nv * B);
    uint32_t s = 12345u;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (double)(s >> 8) / (double)(1u << 24) * 2.0 - 1.0; };
    for (auto * x : {&q, &v, &a}) for (double & e : *x) e = rnd();
    pb::rneaInParallel(1, pool, {q.data(), nq, B, nq}, {v.data(), nv, B, nv}, {a.data(), nv, B, nv}, {tau.data(), nv, B, nv});
    pb::abaInParallel(1, pool, {q.data(), nq, B, nq}, {v.data(), nv, B, nv}, {tau.data(), nv, B, nv}, {a2.data(), nv, B, nv});
    double err = 0;
    for (int k = 0; k < nv * B; ++k) err = std::fmax(err, std::fabs(a2[k] - a[k]));
    if (!(err < 1e-10)) { std::printf("FAIL aba(rnea(a)) != a, max err %.3e\n", err); return 1; }
    bool threw = false;
    try { pb::rneaInParallel(1, pool, {q.data(), nq - 1, B, nq}, {v.data(), nv, B, nv}, {a.data(), nv, B, nv}, {tau.data(), nv, B, nv}); }
    catch (const std::invalid_argument &) { threw = true; }
    if (!threw) { std::printf("FAIL wrong-size argument did not throw std::invalid_argument\n"); return 1; }
    // the rest of the shim: nle + M a == tau (unittest/rnea.cpp:195-206), Minv M == 1 on the diagonal, integrate(q, 0) == q
    std::vector<double> nle(nv * B), M(nv * nv * B), Minv(nv * nv * B), zero(nv * B, 0.0), q2(nq * B), qn(nq * B), vn(nv * B);
    pb::nonLinearEffectsInParallel(1, pool, {q.data(), nq, B, nq}, {v.data(), nv, B, nv}, {nle.data(), nv, B, nv});
    pb::crbaInParallel(1, pool, {q.data(), nq, B, nq}, {M.data(), nv * nv, B, nv * nv});
    pb::computeMinverseInParallel(1, pool, {q.data(), nq, B, nq}, {Minv.data(), nv * nv, B, nv * nv});
    pb::integrateInParallel(1, pool, {q.data(), nq, B, nq}, {zero.data(), nv, B, nv}, {q2.data(), nq, B, nq});
    pb::abaEulerStepInParallel(1, pool, {q.data(), nq, B, nq}, {v.data(), nv, B, nv}, {tau.data(), nv, B, nv}, 1e-3, {qn.data(), nq, B, nq},
                               {vn.data(), nv, B, nv});
    double err2 = 0, err3 = 0, err4 = 0;
    for (int b = 0; b < B; ++b)
    {
      const double * Mb = M.data() + (size_t)b * nv * nv, * Ib = Minv.data() + (size_t)b * nv * nv;
      auto sym = [&](const double * A, int r, int c) { return r <= c ? A[c * nv + r] : A[r * nv + c]; }; // upper triangles
      for (int r = 0; r < nv; ++r)
      {
        double t = nle[b * nv + r], d = 0;
        for (int c = 0; c < nv; ++c) { t += sym(Mb, r, c) * a[b * nv + c]; d += sym(Ib, r, c) * sym(Mb, c, r); }
        err2 = std::fmax(err2, std::fabs(t - tau[b * nv + r]));
        err3 = std::fmax(err3, std::fabs(d - 1.0));
      }
      for (int k = 0; k < nv; ++k) err4 = std::fmax(err4, std::fabs(vn[b * nv + k] - (v[b * nv + k] + 1e-3 * a[b * nv + k])));
    }
    for (int k = 0; k < nq * B; ++k) err4 = std::fmax(err4, std::fabs(q2[k] - q[k]));
    if (!(err2 < 1e-9 && err3 < 1e-9 && err4 < 1e-9)) { std::printf("FAIL nle/crba/Minv/integrate/euler identities: %.3e %.3e %.3e\n", err2, err3, err4); return 1; }
    std::printf("OK max|aba(rnea(a)) - a| = %.3e over %d configurations; M a + nle - tau %.1e, diag(Minv M) - 1 %.1e, integrate/euler %.1e\n", err,
                B, err2, err3, err4);
    return 0;
  }
  catch (const std::runtime_error & e)
  {
    std::printf("NO_GPU: %s\n", e.what());
    return 0;
  }
}
