// pinocchio_b200_shim.hpp — C++ host shim above the C ABI (include/pinocchio_b200.h) that keeps the
// reference's batched signatures, so that a caller of
//   pinocchio::rneaInParallel(num_threads, pool, q, v, a, tau)   (include/pinocchio/algorithm/parallel/rnea.hpp:31-83)
//   pinocchio::abaInParallel (num_threads, pool, q, v, tau, a)   (include/pinocchio/algorithm/parallel/aba.hpp:32-84)
// only swaps the pool type.  Batched crba / computeRNEADerivatives / computeABADerivatives do not exist in the
// reference; they are defined by analogy (single-configuration semantics of algorithm/crba.hpp:47-51,
// rnea-derivatives.hpp:110-128, aba-derivatives.hpp:52-66; every matrix output is one col-major nv x nv
// matrix per column of an (nv*nv) x B block).
//
// Two overload families:
//   * MatrixView / ConstMatrixView {data, rows, cols, outer stride} — always available, no dependency;
//   * Eigen::MatrixBase<...> templates with the reference's exact parameter list — compiled only when
//     <Eigen/Core> is on the include path (it is not in this image).
// Errors: the C ABI returns a status; the shim turns BRBD_EINVAL into std::invalid_argument with the
// reference's messages (macros.hpp:185-223) and everything else into std::runtime_error.
// num_threads is accepted for signature compatibility and ignored: the GPU pool has no per-thread
// replicas and, unlike set_default_omp_options (parallel/omp.hpp:12-16), no process-global state is touched.
#ifndef PINOCCHIO_B200_SHIM_HPP
#define PINOCCHIO_B200_SHIM_HPP

#include <cstddef>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "pinocchio_b200.h"

#if defined(__has_include)
#if __has_include(<Eigen/Core>)
#include <Eigen/Core>
#define PINOCCHIO_B200_WITH_EIGEN 1
#endif
#endif

namespace pinocchio_b200
{

struct ConstMatrixView
{
  const double * data;
  int64_t rows, cols, ld; // ld = elements between consecutive columns (Eigen outerStride())
};
struct MatrixView
{
  double * data;
  int64_t rows, cols, ld;
  operator ConstMatrixView() const { return ConstMatrixView{data, rows, cols, ld}; }
};

inline void check_status(brbd_status st)
{
  if (st == BRBD_OK) return;
  const std::string msg = brbd_last_error_string();
  if (st == BRBD_EINVAL) throw std::invalid_argument(msg);
  throw std::runtime_error(msg);
}

// Device analogue of ModelPoolTpl (multibody/pool/model.hpp:19-165): one staged model replica, one
// arena and three streams per listed device.  `flags` of the calls: host pointers, FP64.
class DeviceModelPool
{
public:
  // flat: the fields of ModelTpl the algorithms read (see brbd_flat_model); devices: CUDA ordinals,
  // empty = device 0.  (With Pinocchio available, flatten a pinocchio::Model with
  // INTEGRATION.md's `flatten(const pinocchio::Model &)`.)
  // from a URDF file / text (pinocchio::urdf::buildModel + ModelPool in one step); root_joint_type: BRBD_JOINT_UNIVERSE = fixed base
  static DeviceModelPool * fromUrdf(const std::string & path_or_xml, int root_joint_type = BRBD_JOINT_UNIVERSE, const std::vector<int> & devices = {})
  {
    brbd_model * m = nullptr;
    check_status(brbd_model_from_urdf(path_or_xml.c_str(), root_joint_type, &m));
    brbd_flat_model f;
    check_status(brbd_model_get_flat(m, &f));
    DeviceModelPool * p = nullptr;
    try { p = new DeviceModelPool(f, devices); } catch (...) { brbd_model_destroy(m); throw; }
    brbd_model_destroy(m);
    return p;
  }
  explicit DeviceModelPool(const brbd_flat_model & flat, const std::vector<int> & devices = {})
  {
    check_status(brbd_model_create(&flat, &model_));
    const brbd_status st = brbd_pool_create(model_, devices.empty() ? nullptr : devices.data(), (int)devices.size(), &pool_);
    if (st != BRBD_OK)
    {
      brbd_model_destroy(model_);
      model_ = nullptr;
      check_status(st);
    }
  }
  DeviceModelPool(const DeviceModelPool &) = delete;
  DeviceModelPool & operator=(const DeviceModelPool &) = delete;
  ~DeviceModelPool()
  {
    if (pool_) brbd_pool_destroy(pool_);
    if (model_) brbd_model_destroy(model_);
  }
  // ModelPoolTpl::size (pool/model.hpp:76): number of replicas = number of devices
  size_t size() const { return (size_t)brbd_pool_size(pool_); }
  // ModelPoolTpl::update (pool/model.hpp:100-108)
  void update(const brbd_flat_model & flat)
  {
    brbd_model * m = nullptr;
    check_status(brbd_model_create(&flat, &m));
    const brbd_status st = brbd_pool_update(pool_, m);
    if (st == BRBD_OK)
    {
      brbd_model_destroy(model_);
      model_ = m;
    }
    else
      brbd_model_destroy(m);
    check_status(st);
  }
  // ModelPoolTpl::getModel / getModels (pool/model.hpp:60-89): every replica holds the same model; what comes back is the
  // flattened view of it (pointers owned by the pool, valid until the next update() / destruction)
  brbd_flat_model getModel(size_t index = 0) const
  {
    // wording of pool/model.hpp:63-65
    if (index >= size()) throw std::invalid_argument("Index greater than the size of the model vector.");
    brbd_flat_model f;
    check_status(brbd_model_get_flat(model_, &f));
    return f;
  }
  std::vector<brbd_flat_model> getModels() const { return std::vector<brbd_flat_model>(size(), getModel(0)); }
  // ModelPoolTpl::getData / getDatas (pool/model.hpp:91-...): a device replica has no per-thread Data; it owns a CUDA device and
  // its grow-only arenas
  struct DeviceData { int device; uint64_t workspace_bytes; };
  DeviceData getData(size_t index) const
  {
    if (index >= size()) throw std::invalid_argument("Index greater than the size of the data vector.");
    return DeviceData{brbd_pool_device_id(pool_, (int)index), brbd_pool_workspace_bytes(pool_, (int)index)};
  }
  std::vector<DeviceData> getDatas() const
  {
    std::vector<DeviceData> v;
    for (size_t i = 0; i < size(); ++i) v.push_back(getData(i));
    return v;
  }
  // ModelPoolTpl::resize (pool/model.hpp:110-131): the replicas of a device pool are devices
  void resize(const std::vector<int> & devices) { check_status(brbd_pool_resize(pool_, devices.empty() ? nullptr : devices.data(), (int)devices.size())); }
  // kernels generated for this model (the batched analogue of pinocchio's code generation, codegen/code-generator-algo.hpp)
  void specialize(bool rnea = true, bool aba = true, bool crba = true)
  {
    check_status(brbd_pool_specialize(pool_, (rnea ? 1 << BRBD_GEN_RNEA : 0) | (aba ? 1 << BRBD_GEN_ABA : 0) | (crba ? 1 << BRBD_GEN_CRBA : 0), 0));
  }
  int nq() const { return brbd_model_nq(model_); }
  int nv() const { return brbd_model_nv(model_); }
  // the structural pattern of crba's result (crba.hxx:94-95: M.block(idx_v, idx_v, nv, nvSubtree) is all crba writes), column-major:
  // entry k of a packed column (crbaPackedInParallel) is M(rows[k], cols[k])
  void crbaPattern(std::vector<int32_t> & rows, std::vector<int32_t> & cols) const
  {
    int64_t nnz = 0;
    check_status(brbd_model_crba_pattern(model_, nullptr, nullptr, 0, &nnz));
    rows.resize((size_t)nnz); cols.resize((size_t)nnz);
    check_status(brbd_model_crba_pattern(model_, rows.data(), cols.data(), nnz, &nnz));
  }
  int64_t crbaPatternSize() const
  {
    int64_t nnz = 0;
    check_status(brbd_model_crba_pattern(model_, nullptr, nullptr, 0, &nnz));
    return nnz;
  }
  brbd_pool * handle() { return pool_; }
  const brbd_model * model_handle() const { return model_; }
  // Unlike the reference (parallel/rnea.hpp:53 "The pool is too small"), num_threads is NOT checked against size(): the GPU
  // path has no per-thread replicas, any num_threads is accepted.

private:
  brbd_model * model_ = nullptr;
  brbd_pool * pool_ = nullptr;
};

namespace detail
{
inline void check_rows(const char * name, int64_t got, int64_t expected)
{
  // wording of PINOCCHIO_CHECK_ARGUMENT_SIZE (macros.hpp:201-223)
  if (got != expected)
    throw std::invalid_argument(std::string("wrong argument size: expected ") + std::to_string(expected) + ", got "
                                + std::to_string(got) + "\nhint: " + name + " has the wrong number of rows");
}
inline void check_cols(const char * name, int64_t got, int64_t expected)
{
  if (got != expected)
    throw std::invalid_argument(std::string("wrong argument size: expected ") + std::to_string(expected) + ", got "
                                + std::to_string(got) + "\nhint: " + name + " has the wrong number of columns");
}
inline void check_pool(size_t /*num_threads*/, const DeviceModelPool & pool)
{
  // parallel/rnea.hpp:52: "The pool should have at least one element"
  if (pool.size() == 0) throw std::invalid_argument("The pool should have at least one element");
}
} // namespace detail

// tau.col(i) = rnea(q.col(i), v.col(i), a.col(i)) — parallel/rnea.hpp:38-83
inline void rneaInParallel(size_t num_threads, DeviceModelPool & pool, ConstMatrixView q, ConstMatrixView v, ConstMatrixView a,
                           MatrixView tau)
{
  detail::check_pool(num_threads, pool);
  detail::check_rows("q", q.rows, pool.nq());   // rnea.hpp:56
  detail::check_rows("v", v.rows, pool.nv());   // :57
  detail::check_rows("a", a.rows, pool.nv());   // :58
  detail::check_rows("tau", tau.rows, pool.nv()); // :61
  detail::check_cols("v", v.cols, q.cols);      // :63-66
  detail::check_cols("a", a.cols, q.cols);
  detail::check_cols("tau", tau.cols, q.cols);
  check_status(brbd_rnea_batch(pool.handle(), q.data, q.ld, v.data, v.ld, a.data, a.ld, tau.data, tau.ld, q.cols,
                               BRBD_PTR_HOST | BRBD_FP64));
}

// a.col(i) = aba(q.col(i), v.col(i), tau.col(i), Convention::WORLD) — parallel/aba.hpp:40-84.
// (The reference checks a.rows() twice and never tau.rows(), parallel/aba.hpp:63-65; all four are checked here.)
inline void abaInParallel(size_t num_threads, DeviceModelPool & pool, ConstMatrixView q, ConstMatrixView v, ConstMatrixView tau,
                          MatrixView a)
{
  detail::check_pool(num_threads, pool);
  detail::check_rows("q", q.rows, pool.nq());
  detail::check_rows("v", v.rows, pool.nv());
  detail::check_rows("tau", tau.rows, pool.nv());
  detail::check_rows("a", a.rows, pool.nv());
  detail::check_cols("v", v.cols, q.cols);
  detail::check_cols("tau", tau.cols, q.cols);
  detail::check_cols("a", a.cols, q.cols);
  check_status(brbd_aba_batch(pool.handle(), q.data, q.ld, v.data, v.ld, tau.data, tau.ld, a.data, a.ld, q.cols,
                              BRBD_PTR_HOST | BRBD_FP64));
}

// M.col(i) = vec(crba(q.col(i))): upper triangle + armature, zeros elsewhere — crba.hpp:47-51
inline void crbaInParallel(size_t num_threads, DeviceModelPool & pool, ConstMatrixView q, MatrixView M)
{
  detail::check_pool(num_threads, pool);
  detail::check_rows("q", q.rows, pool.nq());
  detail::check_rows("M", M.rows, (int64_t)pool.nv() * pool.nv());
  detail::check_cols("M", M.cols, q.cols);
  // num_threads host threads rebuild the dense matrices from the packed transfer (brbd_pool_set_host_threads); 0 / 1: plain copy
  check_status(brbd_pool_set_host_threads(pool.handle(), (int)num_threads));
  check_status(brbd_crba_batch(pool.handle(), q.data, q.ld, M.data, M.ld, q.cols, BRBD_PTR_HOST | BRBD_FP64));
}
// M.col(i) = the dense matrix of the packed column P.col(i), rebuilt by num_threads host threads (brbd_crba_expand_packed)
inline void expandPackedCrba(size_t num_threads, DeviceModelPool & pool, ConstMatrixView P, MatrixView M)
{
  detail::check_rows("P", P.rows, pool.crbaPatternSize());
  detail::check_rows("M", M.rows, (int64_t)pool.nv() * pool.nv());
  detail::check_cols("M", M.cols, P.cols);
  check_status(brbd_crba_expand_packed(pool.model_handle(), P.data, P.ld, M.data, M.ld, P.cols, (int)num_threads, BRBD_PTR_HOST | BRBD_FP64));
}
// P.col(i) = the entries of crba(q.col(i)) inside the structural pattern (pool.crbaPattern), column-major — opt-in output format
inline void crbaPackedInParallel(size_t num_threads, DeviceModelPool & pool, ConstMatrixView q, MatrixView P)
{
  detail::check_pool(num_threads, pool);
  detail::check_rows("q", q.rows, pool.nq());
  detail::check_rows("P", P.rows, pool.crbaPatternSize());
  detail::check_cols("P", P.cols, q.cols);
  check_status(brbd_crba_packed_batch(pool.handle(), q.data, q.ld, P.data, P.ld, q.cols, BRBD_PTR_HOST | BRBD_FP64));
}

// computeRNEADerivatives per column — rnea-derivatives.hpp:110-128 (outputs need not be pre-zeroed here)
inline void computeRNEADerivativesInParallel(size_t num_threads, DeviceModelPool & pool, ConstMatrixView q, ConstMatrixView v,
                                             ConstMatrixView a, MatrixView rnea_partial_dq, MatrixView rnea_partial_dv,
                                             MatrixView rnea_partial_da)
{
  detail::check_pool(num_threads, pool);
  const int64_t nn = (int64_t)pool.nv() * pool.nv();
  detail::check_rows("q", q.rows, pool.nq());
  detail::check_rows("v", v.rows, pool.nv());
  detail::check_rows("a", a.rows, pool.nv());
  detail::check_rows("rnea_partial_dq", rnea_partial_dq.rows, nn);
  detail::check_rows("rnea_partial_dv", rnea_partial_dv.rows, nn);
  detail::check_rows("rnea_partial_da", rnea_partial_da.rows, nn);
  detail::check_cols("v", v.cols, q.cols);
  detail::check_cols("a", a.cols, q.cols);
  check_status(brbd_rnea_derivatives_batch(pool.handle(), q.data, q.ld, v.data, v.ld, a.data, a.ld, rnea_partial_dq.data,
                                           rnea_partial_dq.ld, rnea_partial_dv.data, rnea_partial_dv.ld, rnea_partial_da.data,
                                           rnea_partial_da.ld, nullptr, 0, q.cols, BRBD_PTR_HOST | BRBD_FP64));
}

// computeABADerivatives per column — aba-derivatives.hpp:52-66
inline void computeABADerivativesInParallel(size_t num_threads, DeviceModelPool & pool, ConstMatrixView q, ConstMatrixView v,
                                            ConstMatrixView tau, MatrixView aba_partial_dq, MatrixView aba_partial_dv,
                                            MatrixView aba_partial_dtau)
{
  detail::check_pool(num_threads, pool);
  const int64_t nn = (int64_t)pool.nv() * pool.nv();
  detail::check_rows("q", q.rows, pool.nq());
  detail::check_rows("v", v.rows, pool.nv());
  detail::check_rows("tau", tau.rows, pool.nv());
  detail::check_rows("aba_partial_dq", aba_partial_dq.rows, nn);
  detail::check_rows("aba_partial_dv", aba_partial_dv.rows, nn);
  detail::check_rows("aba_partial_dtau", aba_partial_dtau.rows, nn);
  detail::check_cols("v", v.cols, q.cols);
  detail::check_cols("tau", tau.cols, q.cols);
  check_status(brbd_aba_derivatives_batch(pool.handle(), q.data, q.ld, v.data, v.ld, tau.data, tau.ld, aba_partial_dq.data,
                                          aba_partial_dq.ld, aba_partial_dv.data, aba_partial_dv.ld, aba_partial_dtau.data,
                                          aba_partial_dtau.ld, nullptr, 0, q.cols, BRBD_PTR_HOST | BRBD_FP64));
}

// ---- the callers' other needs on the same sweeps (no batched version upstream; single-configuration semantics) ----
// nle.col(i) = nonLinearEffects(q.col(i), v.col(i)) — algorithm/rnea.hpp:105
inline void nonLinearEffectsInParallel(size_t num_threads, DeviceModelPool & pool, ConstMatrixView q, ConstMatrixView v, MatrixView nle)
{
  detail::check_pool(num_threads, pool);
  detail::check_rows("q", q.rows, pool.nq());
  detail::check_rows("v", v.rows, pool.nv());
  detail::check_rows("nle", nle.rows, pool.nv());
  detail::check_cols("v", v.cols, q.cols);
  detail::check_cols("nle", nle.cols, q.cols);
  check_status(brbd_nle_batch(pool.handle(), q.data, q.ld, v.data, v.ld, nle.data, nle.ld, q.cols, BRBD_PTR_HOST | BRBD_FP64));
}
// g.col(i) = computeGeneralizedGravity(q.col(i)) — algorithm/rnea.hpp:133
inline void computeGeneralizedGravityInParallel(size_t num_threads, DeviceModelPool & pool, ConstMatrixView q, MatrixView g)
{
  detail::check_pool(num_threads, pool);
  detail::check_rows("q", q.rows, pool.nq());
  detail::check_rows("g", g.rows, pool.nv());
  detail::check_cols("g", g.cols, q.cols);
  check_status(brbd_gravity_batch(pool.handle(), q.data, q.ld, g.data, g.ld, q.cols, BRBD_PTR_HOST | BRBD_FP64));
}
// Minv.col(i) = vec(computeMinverse(q.col(i))): upper triangle, zeros below — algorithm/aba.hpp:106
inline void computeMinverseInParallel(size_t num_threads, DeviceModelPool & pool, ConstMatrixView q, MatrixView Minv)
{
  detail::check_pool(num_threads, pool);
  detail::check_rows("q", q.rows, pool.nq());
  detail::check_rows("Minv", Minv.rows, (int64_t)pool.nv() * pool.nv());
  detail::check_cols("Minv", Minv.cols, q.cols);
  check_status(brbd_minverse_batch(pool.handle(), q.data, q.ld, Minv.data, Minv.ld, q.cols, BRBD_PTR_HOST | BRBD_FP64));
}
// qout.col(i) = integrate(model, q.col(i), v.col(i)) — algorithm/joint-configuration.hpp:49-74
inline void integrateInParallel(size_t num_threads, DeviceModelPool & pool, ConstMatrixView q, ConstMatrixView v, MatrixView qout)
{
  detail::check_pool(num_threads, pool);
  detail::check_rows("q", q.rows, pool.nq());
  detail::check_rows("v", v.rows, pool.nv());
  detail::check_rows("qout", qout.rows, pool.nq());
  detail::check_cols("v", v.cols, q.cols);
  detail::check_cols("qout", qout.cols, q.cols);
  check_status(brbd_integrate_batch(pool.handle(), q.data, q.ld, v.data, v.ld, qout.data, qout.ld, q.cols, BRBD_PTR_HOST | BRBD_FP64));
}
// a = aba(q, v, tau); v_next = v + dt a; q_next = integrate(q, dt v_next) per column — examples/simulation-pendulum.py:153-157
inline void abaEulerStepInParallel(size_t num_threads, DeviceModelPool & pool, ConstMatrixView q, ConstMatrixView v, ConstMatrixView tau,
                                   double dt, MatrixView q_next, MatrixView v_next)
{
  detail::check_pool(num_threads, pool);
  detail::check_rows("q", q.rows, pool.nq());
  detail::check_rows("v", v.rows, pool.nv());
  detail::check_rows("tau", tau.rows, pool.nv());
  detail::check_rows("q_next", q_next.rows, pool.nq());
  detail::check_rows("v_next", v_next.rows, pool.nv());
  detail::check_cols("v", v.cols, q.cols);
  detail::check_cols("tau", tau.cols, q.cols);
  detail::check_cols("q_next", q_next.cols, q.cols);
  detail::check_cols("v_next", v_next.cols, q.cols);
  check_status(brbd_aba_euler_step_batch(pool.handle(), q.data, q.ld, v.data, v.ld, tau.data, tau.ld, dt, q_next.data, q_next.ld,
                                         v_next.data, v_next.ld, q.cols, BRBD_PTR_HOST | BRBD_FP64));
}

#ifdef PINOCCHIO_B200_WITH_EIGEN
// The reference's parameter lists (parallel/rnea.hpp:31-45, parallel/aba.hpp:32-46): outputs are passed as
// const references and written through const_cast, exactly as the reference does (rnea.hpp:59).
namespace detail
{
template<class D> ConstMatrixView cview(const Eigen::MatrixBase<D> & m)
{
  static_assert(!D::IsRowMajor, "columns must be configurations (column-major)");
  return ConstMatrixView{m.derived().data(), (int64_t)m.rows(), (int64_t)m.cols(), (int64_t)m.derived().outerStride()};
}
template<class D> MatrixView mview(const Eigen::MatrixBase<D> & m)
{
  static_assert(!D::IsRowMajor, "columns must be configurations (column-major)");
  D & w = const_cast<D &>(m.derived());
  return MatrixView{w.data(), (int64_t)w.rows(), (int64_t)w.cols(), (int64_t)w.outerStride()};
}
} // namespace detail
template<class Q, class V1, class V2, class V3>
void rneaInParallel(const size_t num_threads, DeviceModelPool & pool, const Eigen::MatrixBase<Q> & q, const Eigen::MatrixBase<V1> & v,
                    const Eigen::MatrixBase<V2> & a, const Eigen::MatrixBase<V3> & tau)
{
  rneaInParallel(num_threads, pool, detail::cview(q), detail::cview(v), detail::cview(a), detail::mview(tau));
}
template<class Q, class V1, class V2, class V3>
void abaInParallel(const size_t num_threads, DeviceModelPool & pool, const Eigen::MatrixBase<Q> & q, const Eigen::MatrixBase<V1> & v,
                   const Eigen::MatrixBase<V2> & tau, const Eigen::MatrixBase<V3> & a)
{
  abaInParallel(num_threads, pool, detail::cview(q), detail::cview(v), detail::cview(tau), detail::mview(a));
}
template<class Q, class M1>
void crbaInParallel(const size_t num_threads, DeviceModelPool & pool, const Eigen::MatrixBase<Q> & q, const Eigen::MatrixBase<M1> & M)
{
  crbaInParallel(num_threads, pool, detail::cview(q), detail::mview(M));
}
template<class Q, class M1>
void crbaPackedInParallel(const size_t num_threads, DeviceModelPool & pool, const Eigen::MatrixBase<Q> & q, const Eigen::MatrixBase<M1> & P)
{
  crbaPackedInParallel(num_threads, pool, detail::cview(q), detail::mview(P));
}
// argument order of computeRNEADerivatives(model, data, q, v, a, rnea_partial_dq, rnea_partial_dv, rnea_partial_da), rnea-derivatives.hpp:110-128
template<class Q, class V1, class V2, class M1, class M2, class M3>
void computeRNEADerivativesInParallel(const size_t num_threads, DeviceModelPool & pool, const Eigen::MatrixBase<Q> & q,
                                      const Eigen::MatrixBase<V1> & v, const Eigen::MatrixBase<V2> & a, const Eigen::MatrixBase<M1> & rnea_partial_dq,
                                      const Eigen::MatrixBase<M2> & rnea_partial_dv, const Eigen::MatrixBase<M3> & rnea_partial_da)
{
  computeRNEADerivativesInParallel(num_threads, pool, detail::cview(q), detail::cview(v), detail::cview(a), detail::mview(rnea_partial_dq),
                                   detail::mview(rnea_partial_dv), detail::mview(rnea_partial_da));
}
// computeABADerivatives(model, data, q, v, tau, aba_partial_dq, aba_partial_dv, aba_partial_dtau), aba-derivatives.hpp:52-66
template<class Q, class V1, class V2, class M1, class M2, class M3>
void computeABADerivativesInParallel(const size_t num_threads, DeviceModelPool & pool, const Eigen::MatrixBase<Q> & q,
                                     const Eigen::MatrixBase<V1> & v, const Eigen::MatrixBase<V2> & tau, const Eigen::MatrixBase<M1> & aba_partial_dq,
                                     const Eigen::MatrixBase<M2> & aba_partial_dv, const Eigen::MatrixBase<M3> & aba_partial_dtau)
{
  computeABADerivativesInParallel(num_threads, pool, detail::cview(q), detail::cview(v), detail::cview(tau), detail::mview(aba_partial_dq),
                                  detail::mview(aba_partial_dv), detail::mview(aba_partial_dtau));
}
template<class Q, class V1, class V2>
void nonLinearEffectsInParallel(const size_t num_threads, DeviceModelPool & pool, const Eigen::MatrixBase<Q> & q, const Eigen::MatrixBase<V1> & v,
                                const Eigen::MatrixBase<V2> & nle)
{
  nonLinearEffectsInParallel(num_threads, pool, detail::cview(q), detail::cview(v), detail::mview(nle));
}
template<class Q, class V1>
void computeGeneralizedGravityInParallel(const size_t num_threads, DeviceModelPool & pool, const Eigen::MatrixBase<Q> & q, const Eigen::MatrixBase<V1> & g)
{
  computeGeneralizedGravityInParallel(num_threads, pool, detail::cview(q), detail::mview(g));
}
template<class Q, class M1>
void computeMinverseInParallel(const size_t num_threads, DeviceModelPool & pool, const Eigen::MatrixBase<Q> & q, const Eigen::MatrixBase<M1> & Minv)
{
  computeMinverseInParallel(num_threads, pool, detail::cview(q), detail::mview(Minv));
}
template<class Q, class V1, class Q2>
void integrateInParallel(const size_t num_threads, DeviceModelPool & pool, const Eigen::MatrixBase<Q> & q, const Eigen::MatrixBase<V1> & v,
                         const Eigen::MatrixBase<Q2> & qout)
{
  integrateInParallel(num_threads, pool, detail::cview(q), detail::cview(v), detail::mview(qout));
}
template<class Q, class V1, class V2, class Q2, class V3>
void abaEulerStepInParallel(const size_t num_threads, DeviceModelPool & pool, const Eigen::MatrixBase<Q> & q, const Eigen::MatrixBase<V1> & v,
                            const Eigen::MatrixBase<V2> & tau, double dt, const Eigen::MatrixBase<Q2> & q_next, const Eigen::MatrixBase<V3> & v_next)
{
  abaEulerStepInParallel(num_threads, pool, detail::cview(q), detail::cview(v), detail::cview(tau), dt, detail::mview(q_next), detail::mview(v_next));
}
#endif // PINOCCHIO_B200_WITH_EIGEN

} // namespace pinocchio_b200

#endif // PINOCCHIO_B200_SHIM_HPP
