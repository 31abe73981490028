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
    std::printf("OK max|aba(rnea(a)) - a| = %.3e over %d configurations\n", err, B);
    return 0;
  }
  catch (const std::runtime_error & e)
  {
    std::printf("NO_GPU: %s\n", e.what());
    return 0;
  }
}
