/*
 * pinocchio_b200.h — C ABI of the B200-native batched rigid-body-dynamics engine.
 *
 * This is the drop-in boundary for Pinocchio's many-configuration path:
 *   rneaInParallel   (reference: include/pinocchio/algorithm/parallel/rnea.hpp:31-83)
 *   abaInParallel    (reference: include/pinocchio/algorithm/parallel/aba.hpp:32-84)
 *   ModelPoolTpl     (reference: include/pinocchio/multibody/pool/model.hpp:19-165)
 * extended by analogy to batched crba / computeRNEADerivatives / computeABADerivatives
 *   (reference single-configuration entry points: include/pinocchio/algorithm/crba.hpp:47-51,
 *    rnea-derivatives.hpp:110-128, aba-derivatives.hpp:52-66).
 *
 * Conventions (identical to the reference):
 *   - every batched argument is a column-major block whose COLUMNS are configurations;
 *     `ld*` is the leading dimension in elements (Eigen's outerStride()).
 *   - free-flyer q = [x y z qx qy qz qw]; tangent vectors are body-frame (linear, angular).
 *   - matrix outputs (M and the derivative matrices) are one col-major nv x nv matrix per configuration,
 *     stored contiguously in one column of an (nv*nv) x B block.
 *   - plain pointers and sizes only; no exceptions cross this boundary; every function
 *     returns a brbd_status and brbd_last_error_string() explains the last failure of the
 *     calling thread.
 *   - there is NO CPU fallback: if no CUDA device is usable the calls fail with BRBD_ECUDA.
 */
#ifndef PINOCCHIO_B200_H
#define PINOCCHIO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum brbd_status {
  BRBD_OK = 0,
  BRBD_EINVAL = 1,            /* bad argument / size mismatch (reference: std::invalid_argument) */
  BRBD_EUNSUPPORTED_JOINT = 2,/* joint type outside {RX,RY,RZ,PX,PY,PZ,FreeFlyer,Spherical,Planar} */
  BRBD_ETOPOLOGY = 3,         /* parents[i] >= i, or the tree is not depth-first "compact"
                                 (reference: CRBAChecker, crba.hxx:573-595) */
  BRBD_ECUDA = 4,             /* CUDA runtime error or no device */
  BRBD_ENOMEM = 5
} brbd_status;

/* Joint type tags (reference variant alternatives: multibody/joint/joint-collection.hpp:85-111). */
typedef enum brbd_joint_type {
  BRBD_JOINT_RX = 0, BRBD_JOINT_RY = 1, BRBD_JOINT_RZ = 2,   /* joint-revolute.hpp */
  BRBD_JOINT_PX = 3, BRBD_JOINT_PY = 4, BRBD_JOINT_PZ = 5,   /* joint-prismatic.hpp */
  BRBD_JOINT_FREEFLYER = 6,                                  /* joint-free-flyer.hpp (nq 7, nv 6) */
  BRBD_JOINT_SPHERICAL = 7,                                  /* joint-spherical.hpp  (nq 4, nv 3) */
  BRBD_JOINT_PLANAR = 8,                                     /* joint-planar.hpp     (nq 4, nv 3) */
  BRBD_JOINT_REVOLUTE_UNALIGNED = 9,                         /* joint-revolute-unaligned.hpp: axis in brbd_flat_model::axis */
  BRBD_JOINT_PRISMATIC_UNALIGNED = 10,                       /* joint-prismatic-unaligned.hpp */
  /* joint-revolute-unbounded.hpp:121-235 (URDF "continuous" joints, parsers/urdf/model.hxx:269-273): revolute joints whose
     configuration is (cos q, sin q) — nq 2, nv 1 */
  BRBD_JOINT_RUBX = 11, BRBD_JOINT_RUBY = 12, BRBD_JOINT_RUBZ = 13,
  BRBD_JOINT_REVOLUTE_UNBOUNDED_UNALIGNED = 14,              /* joint-revolute-unbounded-unaligned.hpp: axis in brbd_flat_model::axis */
  BRBD_JOINT_UNIVERSE = -1                                   /* slot 0 only */
} brbd_joint_type;

/*
 * Flattened Model: the fields of ModelTpl the five algorithms read
 * (reference: multibody/model.hpp:97-205). Joint 0 is the universe. All arrays are
 * borrowed for the duration of brbd_model_create only.
 */
typedef struct brbd_flat_model {
  int32_t njoints;            /* including the universe                                   */
  int32_t nq, nv;
  const int32_t * parents;    /* [njoints]  parents[0] = 0                                */
  const int32_t * joint_type; /* [njoints]  brbd_joint_type                               */
  const int32_t * idx_q;      /* [njoints]                                                */
  const int32_t * idx_v;      /* [njoints]                                                */
  const double * placement;   /* [njoints*12] jointPlacements: R row-major (9) then p (3) */
  const double * inertia;     /* [njoints*10] mass, lever(3), Symmetric3 packed
                                 (xx,xy,yy,xz,yz,zz) — spatial/symmetric3.hpp:46-51      */
  const double * armature;    /* [nv]                                                     */
  double gravity[3];          /* linear part of model.gravity (default 0,0,-9.81)         */
  const double * axis;        /* [njoints*3] unit axis of the *_UNALIGNED joints, joint frame (other
                                 entries ignored); may be NULL when the model has none   */
} brbd_flat_model;

typedef struct brbd_model brbd_model; /* validated host copy + derived topology tables */
typedef struct brbd_pool brbd_pool;   /* device analogue of ModelPoolTpl                 */

/* flags for the *_batch calls */
enum {
  BRBD_PTR_HOST = 0,     /* pointers are host memory: the call stages H2D / D2H itself    */
  BRBD_PTR_DEVICE = 1,   /* pointers are device memory on the pool's (single) device      */
  BRBD_FP64 = 0,         /* element type double (default, parity mode)                    */
  BRBD_FP32 = 2,         /* element type float  (tolerance documented in DESIGN.md)       */
  BRBD_ASYNC = 4         /* device pointers only: do not synchronise before returning     */
};

const char * brbd_last_error_string(void);
const char * brbd_version(void);
int brbd_device_count(void);

brbd_status brbd_model_create(const brbd_flat_model * flat, brbd_model ** out);
void brbd_model_destroy(brbd_model * m);
/* URDF -> model on the host C++ side (src/parsers/urdf/model.cpp:67-294, include/pinocchio/parsers/urdf/model.hxx:205-372:
 * name-sorted depth-first visit, fixed joints merged into the parent body, axis-aligned / unaligned / continuous joints).
 * `path_or_xml`: a file name, or the XML text itself when it starts with '<'.  `root_joint_type`: BRBD_JOINT_UNIVERSE for a
 * fixed base, else BRBD_JOINT_FREEFLYER / PLANAR / SPHERICAL (pinocchio::urdf::buildModel(file, root_joint, model)).
 * Mimic joints are refused (BRBD_EUNSUPPORTED_JOINT). */
brbd_status brbd_model_from_urdf(const char * path_or_xml, int root_joint_type, brbd_model ** out);
int brbd_model_nq(const brbd_model * m);
int brbd_model_nv(const brbd_model * m);
int brbd_model_njoints(const brbd_model * m);
/* The model as it was given to brbd_model_create (ModelPoolTpl::getModel, pool/model.hpp:60-74): the pointers of
 * `out` refer to storage owned by `m` and stay valid until brbd_model_destroy(m). */
brbd_status brbd_model_get_flat(const brbd_model * m, brbd_flat_model * out);

/* One staged model replica + one scratch arena + one stream per listed device
 * (device analogue of ModelPoolTpl(model, pool_size), pool/model.hpp:41-46). */
brbd_status brbd_pool_create(const brbd_model * m, const int * device_ids, int n_devices,
                             brbd_pool ** out);
void brbd_pool_destroy(brbd_pool * p);
int brbd_pool_size(const brbd_pool * p);               /* number of devices */
/* ModelPoolTpl::resize (pool/model.hpp:110-131): replace the set of devices (replicas); NULL / 0 = device 0.
 * Staging and workspaces of the dropped replicas are released, the model is re-staged on the new ones. */
brbd_status brbd_pool_resize(brbd_pool * p, const int * device_ids, int n_devices);
/* ModelPoolTpl::getModel(index): the model every replica holds (owned by the pool). */
const brbd_model * brbd_pool_model(const brbd_pool * p);
/* ModelPoolTpl::getData(index) has no per-thread Data here; what a replica owns is a device and its grow-only arenas:
 * CUDA ordinal of replica `index` (-1 if out of range) and the bytes it currently holds on that device. */
int brbd_pool_device_id(const brbd_pool * p, int index);
uint64_t brbd_pool_workspace_bytes(const brbd_pool * p, int index);
/* Replace the model held by every replica (ModelPoolTpl::update, pool/model.hpp:100-108). */
brbd_status brbd_pool_update(brbd_pool * p, const brbd_model * m);
/* Use an external CUDA stream (cudaStream_t as void*) for device-pointer calls on a
 * single-device pool; NULL restores the pool's own (non-blocking) stream.  The legacy default
 * stream is selected by its CUDA handle cudaStreamLegacy ((cudaStream_t)0x1). */
brbd_status brbd_pool_set_stream(brbd_pool * p, void * cuda_stream);
brbd_status brbd_pool_synchronize(brbd_pool * p);
/* Number of kernels this pool has launched since creation (bench.py's gpu_launches). */
int64_t brbd_pool_launch_count(const brbd_pool * p);
/* CUDA-event time (ms) of the kernels of the last device-pointer call on device 0. */
double brbd_pool_last_kernel_ms(const brbd_pool * p);

/* tau[:,i] = rnea(q[:,i], v[:,i], a[:,i])     — replaces rneaInParallel (parallel/rnea.hpp:38) */
brbd_status brbd_rnea_batch(brbd_pool * p, const void * q, int64_t ldq, const void * v, int64_t ldv,
                            const void * a, int64_t lda, void * tau, int64_t ldtau, int64_t batch,
                            int flags);
/* a[:,i] = aba(q[:,i], v[:,i], tau[:,i], WORLD) — replaces abaInParallel (parallel/aba.hpp:40) */
brbd_status brbd_aba_batch(brbd_pool * p, const void * q, int64_t ldq, const void * v, int64_t ldv,
                           const void * tau, int64_t ldtau, void * a, int64_t lda, int64_t batch,
                           int flags);
/* M[:,i] = vec(crba(q[:,i])): upper triangle + armature on the diagonal, strictly-lower part
 * written as zeros (crba.hpp:15-22; a fresh Data holds zeros there, data.hxx:43). ldM >= nv*nv. */
brbd_status brbd_crba_batch(brbd_pool * p, const void * q, int64_t ldq, void * M, int64_t ldM,
                            int64_t batch, int flags);
/* Packed CRBA (opt-in; not a reference signature): P[:,i] = the nnz entries of crba(q[:,i]) that lie inside the structural
 * pattern of the model — for every column the dofs of its own joint and of the joint's ancestors; everything else in
 * the reference's data.M is a structural zero (crba.hxx:94-95 writes M.block(idx_v, idx_v, nv, nvSubtree) only) — in
 * column-major order: P[k,i] = M_i[rows[k], cols[k]] with (rows, cols) from brbd_model_crba_pattern.  For a humanoid
 * that is a third of nv*nv, i.e. a third of the device-to-host traffic that bounds a host-pointer CRBA call.
 * Runs the kernel generated for the pool's model (specialises CRBA at the first call: needs NVRTC).  ldP >= nnz. */
brbd_status brbd_crba_packed_batch(brbd_pool * p, const void * q, int64_t ldq, void * P, int64_t ldP,
                                   int64_t batch, int flags);
/* Host utility for callers of brbd_crba_packed_batch that need data.M after all: M[:,i] = the dense column-major nv x nv matrix of
 * P[:,i] (zeros outside the pattern), rebuilt by `threads` host threads with non-temporal stores.  Host blocks only; a copy,
 * no arithmetic. */
brbd_status brbd_crba_expand_packed(const brbd_model * m, const void * P, int64_t ldP, void * M, int64_t ldM,
                                    int64_t batch, int threads, int flags);
/* The pattern: rows / cols may be NULL (size query); otherwise `capacity` entries each. */
brbd_status brbd_model_crba_pattern(const brbd_model * m, int32_t * rows, int32_t * cols, int64_t capacity,
                                    int64_t * nnz);
/* computeRNEADerivatives per column (rnea-derivatives.hpp:110-128). dtau_dq, dtau_dv: full
 * tree-sparse matrices; dtau_da: upper triangle of M (others zero). tau may be NULL. */
brbd_status brbd_rnea_derivatives_batch(brbd_pool * p, const void * q, int64_t ldq, const void * v,
                                        int64_t ldv, const void * a, int64_t lda, void * dtau_dq,
                                        int64_t ld_dq, void * dtau_dv, int64_t ld_dv,
                                        void * dtau_da, int64_t ld_da, void * tau, int64_t ldtau,
                                        int64_t batch, int flags);
/* computeABADerivatives per column (aba-derivatives.hpp:52-66). ddq_dtau = Minv (full
 * symmetric). ddq may be NULL. */
brbd_status brbd_aba_derivatives_batch(brbd_pool * p, const void * q, int64_t ldq, const void * v,
                                       int64_t ldv, const void * tau, int64_t ldtau, void * ddq_dq,
                                       int64_t ld_dq, void * ddq_dv, int64_t ld_dv,
                                       void * ddq_dtau, int64_t ld_dtau, void * ddq, int64_t ldddq,
                                       int64_t batch, int flags);

/* ---- the callers' other needs, on the same sweeps ---------------------------------------- */
/* nle[:,i] = nonLinearEffects(q[:,i], v[:,i])  (algorithm/rnea.hpp:105, rnea.hxx:227-343) = rnea(q, v, 0) */
brbd_status brbd_nle_batch(brbd_pool * p, const void * q, int64_t ldq, const void * v, int64_t ldv,
                           void * nle, int64_t ldn, int64_t batch, int flags);
/* g[:,i] = computeGeneralizedGravity(q[:,i])   (algorithm/rnea.hpp:133, rnea.hxx:346-452) = rnea(q, 0, 0) */
brbd_status brbd_gravity_batch(brbd_pool * p, const void * q, int64_t ldq, void * g, int64_t ldg,
                               int64_t batch, int flags);
/* Minv[:,i] = vec(computeMinverse(q[:,i]))      (algorithm/aba.hpp:106, aba.hxx:613-902): upper triangle of
 * M^-1, strictly-lower part written as zeros (a fresh data.Minv). ldM >= nv*nv. */
brbd_status brbd_minverse_batch(brbd_pool * p, const void * q, int64_t ldq, void * Minv, int64_t ldM,
                                int64_t batch, int flags);
/* qout[:,i] = integrate(model, q[:,i], v[:,i])  (algorithm/joint-configuration.hpp:49-74) */
brbd_status brbd_integrate_batch(brbd_pool * p, const void * q, int64_t ldq, const void * v, int64_t ldv,
                                 void * qout, int64_t ldqo, int64_t batch, int flags);
/* One semi-implicit Euler step of forward dynamics, entirely on the device:
 *   a = aba(q, v, tau);  v_next = v + dt a;  q_next = integrate(q, dt v_next)
 * (the loop of examples/simulation-pendulum.py:153-157: a = aba(...); v += a dt; q = integrate(model, q, v dt)).
 * q_next / v_next must not alias q / v. */
brbd_status brbd_aba_euler_step_batch(brbd_pool * p, const void * q, int64_t ldq, const void * v, int64_t ldv,
                                      const void * tau, int64_t ldtau, double dt, void * q_next,
                                      int64_t ldqn, void * v_next, int64_t ldvn, int64_t batch, int flags);

/* ---- per-model code generation (the GPU analogue of include/pinocchio/codegen/code-generator-algo.hpp:22-570) --------
 * brbd_codegen_source runs the algorithm once on a recording scalar and returns the CUDA source of a kernel specialised
 * for THIS model: the tree unrolled, joint types resolved, placements / inertias folded in as constants.  *source is
 * malloc'ed; release it with brbd_codegen_free.  brbd_pool_specialize compiles it (NVRTC) and makes the pool's
 * brbd_rnea_batch / brbd_aba_batch use it for large batches. */
enum { BRBD_GEN_RNEA = 0, BRBD_GEN_ABA = 1, BRBD_GEN_CRBA = 2,
       BRBD_GEN_RNEA_DERIVATIVES = 3, BRBD_GEN_ABA_DERIVATIVES = 4 /* small models only: every result stays alive to the end */ };
enum {
  BRBD_GEN_EXPLICIT_SLOTS = 1, /* long-lived values in explicit on-chip slots instead of compiler-managed local memory */
  BRBD_GEN_HOST = 2,           /* emit the host-callable variant (tests of the generator only)                          */
  BRBD_GEN_FP32 = 4,           /* element type float                                                                     */
  BRBD_GEN_DIRECT_IO = 8,      /* no shared-memory input / output tiles: every lane reads its own column from global memory */
  BRBD_GEN_CRBA_COMPACT = 16,  /* CRBA: only the entries of the structural pattern are staged on chip; the kernel writes the  *
                                * dense matrices (zeros put back by the flush) or the packed entries (brbd_crba_packed_batch)  */
  BRBD_GEN_CRBA_BULK = 32      /* CRBA: every lane hands the staged columns of its configuration to the copy engine           *
                                * (cp.async.bulk, asynchronous); bits 29..30: staging rows per lane - 1                        */
  /* bits 8..19: threads per block (0 = default), bits 20..23: min blocks per SM of __launch_bounds__,                    *
   * bits 24..28: CRBA: adjacent columns per flush (0 = 1); with BRBD_GEN_CRBA_COMPACT: pattern entries per flush / 8   */
};
typedef struct brbd_codegen_info {
  int32_t record_slots, park_slots, nodes, live_nodes, adds, muls, recips, sqrts, sincos, loads, stores, threads_per_block;
  int32_t smem_slots, tmem_slots; /* explicit slots: values per configuration in shared memory / tensor memory */
  int32_t dynamic_smem_bytes;     /* per CTA: the warps' input / output tiles + the shared-memory park slots     */
  int32_t copies;                 /* identical kernels brbd_gen_<algo>_<c> at different code addresses            */
} brbd_codegen_info;
brbd_status brbd_codegen_source(const brbd_model * m, int algo, int flags, char ** source, brbd_codegen_info * info);
void brbd_codegen_free(char * source);
/* Generate, compile (NVRTC, sm_100a) and load kernels specialised for the pool's model.  algo_mask: bit (1 << BRBD_GEN_*);
 * flags: BRBD_GEN_FP32 to specialise the float instantiation, BRBD_GEN_EXPLICIT_SLOTS.  Afterwards brbd_rnea_batch /
 * brbd_aba_batch / brbd_crba_batch (and the calls built on them) run the specialised kernel from the batch size where it beats
 * the small-batch paths (measured per algorithm: CRBA always, RNEA from 4096, ABA from 8192 configurations per device, models of
 * at most 8 dofs always); brbd_pool_set_specialized_min_batch (>= 0) overrides that for every algorithm, -1 restores it;
 * brbd_pool_update drops the kernels. */
brbd_status brbd_pool_specialize(brbd_pool * p, int algo_mask, int flags);
int brbd_pool_specialized(const brbd_pool * p); /* mask of algorithms that have a specialised kernel */
brbd_status brbd_pool_set_specialized_min_batch(brbd_pool * p, int64_t min_batch);

/* Page-lock caller-owned host memory (cudaHostRegister / cudaHostUnregister).  Host-pointer calls work on
 * pageable memory too, but only pinned memory reaches the full link bandwidth (measured on this pool's
 * B200 hosts: 57 GB/s pinned vs 11-22 GB/s pageable) and lets the upload / compute / download pipeline of a
 * call overlap.  An Eigen::MatrixXd that is reused across calls should be registered once. */
brbd_status brbd_host_register(void * ptr, uint64_t bytes);
/* Host threads a host-pointer call may use — the `num_threads` of the reference's parallel API (parallel/rnea.hpp:38), which the
 * reference spends on the algorithm itself.  Here they serve one purpose: with n >= 2 (and a single-device pool, NVRTC available),
 * a host-pointer brbd_crba_batch moves only the entries inside the structural pattern over PCIe (brbd_crba_packed_batch's
 * format, a third of nv * nv for a humanoid) and n threads rebuild the caller's dense matrices while the next chunk is in
 * flight.  Results are bit-identical to the packed kernel's; 0 or 1 (default): the dense block is copied as it is.  Applies to
 * batches of at least 4096 configurations; the first such call generates and compiles the packed kernel (about a second). */
brbd_status brbd_pool_set_host_threads(brbd_pool * p, int n);
brbd_status brbd_host_unregister(void * ptr);

/* Register-resident DFMA loop; returns achieved FP64 FLOP/s on the pool's device 0
 * (SURVEY.md §7 hard part 4: the FP64 roofline denominator is measured, not assumed). */
brbd_status brbd_measure_fp64_peak(brbd_pool * p, double * flops_per_s, double * elapsed_ms);

#ifdef __cplusplus
}
#endif
#endif /* PINOCCHIO_B200_H */
