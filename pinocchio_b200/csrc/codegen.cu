// codegen.cu — per-model code generation behind the C ABI (host code only; compiled by nvcc because it instantiates the
// engine's device headers with the recording scalar, BRBD_DI = __host__ __device__).
//   brbd_codegen_source : model + algorithm -> CUDA source of a kernel specialised for that model (codegen/trace.hpp,
//                         codegen/emit.hpp); also the host-callable variant used by the CPU tests of the generator.
// The GPU analogue of the reference's code generation (include/pinocchio/codegen/code-generator-algo.hpp:22-570).
#define BRBD_DI __host__ __device__ __forceinline__
#define BRBD_SYNCWARP() ((void)0) // no kernel of this translation unit is ever launched
#include "codegen/sym.hpp"
#include "host_ctx.hpp"
#include "codegen/trace.hpp"
#include "codegen/trace_derivs.hpp"
#include "codegen/emit.hpp"

#include <sstream>

using namespace brbd;

namespace
{
const char * algo_name(int algo)
{
  const char * names[] = {"rnea", "aba", "crba", "rnea_derivatives", "aba_derivatives"};
  return names[algo];
}

const char * math_macros(bool fp32)
{
  return fp32 ? "typedef float real;\n#define BRBD_C(x) ((real)(x))\n#define BRBD_SINCOS(x, s, c) sincosf((x), (s), (c))\n#define BRBD_SIN(x) sinf(x)\n"
                "#define BRBD_COS(x) cosf(x)\n#define BRBD_SQRT(x) sqrtf(x)\n#define BRBD_MAX(a, b) fmaxf((a), (b))\n"
              : "typedef double real;\n#define BRBD_C(x) (x)\n#define BRBD_SINCOS(x, s, c) sincos((x), (s), (c))\n#define BRBD_SIN(x) sin(x)\n"
                "#define BRBD_COS(x) cos(x)\n#define BRBD_SQRT(x) sqrt(x)\n#define BRBD_MAX(a, b) fmax((a), (b))\n";
}

// park_st / park_ld helpers of the generated kernel for one group shape.  Tensor memory (space 0): the group is a run of
// 32-bit columns of the warp's slice, written / read with tcgen05.st / tcgen05.ld 32x32b in power-of-two chunks (SASS STTM /
// LDTM; thread t of a warp owns lane 32 * (warp % 4) + t, see tmem.cuh).  Shared memory (space 1): slot-major, [slot][thread].
void emit_park_helpers(std::ostringstream & os, int n, int space, bool fp32, int nt)
{
  const int wpv = fp32 ? 1 : 2, words = n * wpv;
  auto word = [&](int w) -> std::string {
    const std::string v = "v" + std::to_string(w / wpv);
    if (fp32) return "__float_as_uint(" + v + ")";
    return (w % 2 == 0) ? "(unsigned)__double2loint(" + v + ")" : "(unsigned)__double2hiint(" + v + ")";
  };
  const char sp = space == 0 ? 'T' : 'S';
  // store
  os << "__device__ __forceinline__ void park_st" << sp << n << "(" << (space == 0 ? "unsigned a" : "real * a");
  for (int k = 0; k < n; ++k) os << ", real v" << k;
  os << ")\n{\n";
  if (space == 1)
    for (int k = 0; k < n; ++k) os << "  a[" << k * nt << "] = v" << k << ";\n";
  else
    for (int w0 = 0; w0 < words;)
    {
      int c = 16;
      while (c > words - w0) c >>= 1;
      os << "  asm volatile(\"tcgen05.st.sync.aligned.32x32b.x" << c << ".b32 [%0], {";
      for (int k = 0; k < c; ++k) os << (k ? ", " : "") << "%" << k + 1;
      os << "};\" ::\"r\"(a + " << w0 << ")";
      for (int k = 0; k < c; ++k) os << ", \"r\"(" << word(w0 + k) << ")";
      os << " : \"memory\");\n";
      w0 += c;
    }
  os << "}\n";
  // load
  os << "__device__ __forceinline__ void park_ld" << sp << n << "(" << (space == 0 ? "unsigned a" : "const real * a");
  for (int k = 0; k < n; ++k) os << ", real & v" << k;
  os << ")\n{\n";
  if (space == 1)
    for (int k = 0; k < n; ++k) os << "  v" << k << " = a[" << k * nt << "];\n";
  else
  {
    os << "  unsigned w[" << words << "];\n  asm volatile(\"tcgen05.wait::st.sync.aligned;\" ::: \"memory\");\n";
    for (int w0 = 0; w0 < words;)
    {
      int c = 16;
      while (c > words - w0) c >>= 1;
      os << "  asm volatile(\"tcgen05.ld.sync.aligned.32x32b.x" << c << ".b32 {";
      for (int k = 0; k < c; ++k) os << (k ? ", " : "") << "%" << k;
      os << "}, [%" << c << "];\" : ";
      for (int k = 0; k < c; ++k) os << (k ? ", " : "") << "\"=r\"(w[" << w0 + k << "])";
      os << " : \"r\"(a + " << w0 << ") : \"memory\");\n";
      w0 += c;
    }
    os << "  asm volatile(\"tcgen05.wait::ld.sync.aligned;\" ::: \"memory\");\n";
    for (int k = 0; k < n; ++k)
    {
      if (fp32) os << "  v" << k << " = __uint_as_float(w[" << k << "]);\n";
      else os << "  v" << k << " = __hiloint2double((int)w[" << 2 * k + 1 << "], (int)w[" << 2 * k << "]);\n";
    }
  }
  os << "}\n";
}

// Device wrapper.  A warp owns 32 consecutive configurations per round of the persistent grid.  Its q / v / x columns (the
// caller's layout: one configuration = `rows` contiguous elements) are copied with COALESCED loads into a per-warp tile in
// shared memory, row per configuration with an odd pitch, so that each lane then walks its own row conflict-free; the result
// overwrites the x row and leaves through the same tile with coalesced stores.  (First version: one strided LDG per element and
// lane, 32 sectors per instruction — the generated RNEA was bound by exactly that.)
// sincos of the generated FP64 device code.  CUDA's sincos costs ~100 instructions per call site in straight-line code (a
// third of them UMOV pairs materialising its polynomial coefficients, plus the out-of-line slow path of the Payne-Hanek
// reduction), i.e. 2.9 k of the 7.7 k instructions of the humanoid RNEA — and the generated kernels run at the rate their
// instructions can be FETCHED.  Same method, leaner: Cody-Waite reduction by pi/2 in three FMA steps, the fdlibm kernels
// (__kernel_sin / __kernel_cos minimax polynomials on [-pi/4, pi/4], error < 1 ulp), coefficients as constant-bank operands;
// No Payne-Hanek path: with FMA the three-step reduction keeps its absolute error near 1e-16 while n = rint(2 x / pi) is exact,
// i.e. far beyond any joint angle; the quadrant is taken from a 64-bit n.  (A shared out-of-line slow path was tried: the call
// sites' ABI spills cost more than the whole saving.)
const char * device_sincos(bool fp32)
{
  if (fp32) return "#define BRBD_SINCOS(x, s, c) sincosf((x), (s), (c))\n";
  return "__constant__ double BRBD_SC[16] = {6.36619772367581382433e-01, 1.57079632679489655800e+00, 6.12323399573676603587e-17, 8.47842766036889956997e-32,\n"
         "  -1.66666666666666324348e-01, 8.33333333332248946124e-03, -1.98412698298579493134e-04, 2.75573137070700676789e-06, -2.50507602534068634195e-08, 1.58969099521155010221e-10,\n"
         "  4.16666666666666019037e-02, -1.38888888888741095749e-03, 2.48015872894767294178e-05, -2.75573143513906633035e-07, 2.08757232129817482790e-09, -1.13596475577881948265e-11};\n"
         "__device__ __forceinline__ void brbd_sincos(double x, double * s, double * c)\n{\n"
         "  const double n = rint(x * BRBD_SC[0]);\n  const int i = (int)(__double2ll_rn(n) & 3ll);\n"
         "  double r = fma(-n, BRBD_SC[1], x);\n  r = fma(-n, BRBD_SC[2], r);\n  r = fma(-n, BRBD_SC[3], r);\n"
         "  const double z = r * r;\n"
         "  double ps = fma(z, BRBD_SC[9], BRBD_SC[8]);\n  ps = fma(z, ps, BRBD_SC[7]);\n  ps = fma(z, ps, BRBD_SC[6]);\n  ps = fma(z, ps, BRBD_SC[5]);\n  ps = fma(z, ps, BRBD_SC[4]);\n"
         "  const double sr = fma(z * r, ps, r);\n"
         "  double pc = fma(z, BRBD_SC[15], BRBD_SC[14]);\n  pc = fma(z, pc, BRBD_SC[13]);\n  pc = fma(z, pc, BRBD_SC[12]);\n  pc = fma(z, pc, BRBD_SC[11]);\n  pc = fma(z, pc, BRBD_SC[10]);\n"
         "  const double cr = fma(z * z, pc, fma(z, -0.5, 1.0));\n"
         "  const double s0 = (i & 1) ? cr : sr, c0 = (i & 1) ? sr : cr;\n"
         "  *s = (i & 2) ? -s0 : s0;\n  *c = ((i + 1) & 2) ? -c0 : c0;\n}\n"
         "#define BRBD_SINCOS(x, s, c) brbd_sincos((x), (s), (c))\n";
}

std::string wrap_device(const std::string & body, const char * name, bool fp32, int nrec, const cg::EmitStats & st, int nt, int minb,
                        int tmem_cols, int nq, int nv, int copies, bool direct_io, const std::string & ktable)
{
  std::ostringstream os;
  // direct_io: no shared-memory tiles, every lane reads / writes its own column in global memory (strided, through L1); leaves
  // shared memory free, so more warps fit per SM
  const int qp = nq | 1, vp = nv | 1, tile = direct_io ? 0 : 32 * (qp + 2 * vp), warps = nt / 32;
  os << "// generated by pinocchio_b200 codegen: " << name << (fp32 ? " (FP32)" : " (FP64)") << ", " << nrec << " record slots, "
     << st.tmem_slots << " tensor-memory + " << st.smem_slots << " shared-memory park slots per configuration\n";
  os << math_macros(fp32) << "#undef BRBD_SINCOS\n" << device_sincos(fp32) << ktable;
  // record store: written in pass 2, read once in pass 3.  With L2 cache policies (experiment, BRBD_GEN_REC_POLICY=1) the stores
  // ask L2 to keep the lines (evict_last) and the loads release them (evict_first), so that the 139 MB of records of a 65 536 x
  // humanoid batch do not leave the 126 MB L2 between the passes.
  // Measured (profiles/r2_aba_rec_policy.txt): ABA 0.155 -> 0.146 ms at 65 536 x simple_humanoid, 2.46 -> 2.34 ms at 2^20: the default.
  // BRBD_GEN_REC_POLICY: 0 = off, 1 = records only, 2 = the input columns are also loaded evict_first (they are read once).
  const int rec_policy_mode = std::getenv("BRBD_GEN_REC_POLICY") ? std::atoi(std::getenv("BRBD_GEN_REC_POLICY")) : 1;
  const bool rec_policy = rec_policy_mode != 0;
  const char * keep_fraction = std::getenv("BRBD_GEN_REC_KEEP") ? std::getenv("BRBD_GEN_REC_KEEP") : "1.0";
  if (rec_policy)
    os << (fp32 ? "__device__ __forceinline__ real ld_rec(const real * p, unsigned long long pol) { real v; asm volatile(\"ld.global.cg.L2::cache_hint.f32 %0, [%1], %2;\" : \"=f\"(v) : \"l\"(p), \"l\"(pol) : \"memory\"); return v; }\n"
                  "__device__ __forceinline__ void st_rec(real * p, real v, unsigned long long pol) { asm volatile(\"st.global.cg.L2::cache_hint.f32 [%0], %1, %2;\" ::\"l\"(p), \"f\"(v), \"l\"(pol) : \"memory\"); }\n"
                : "__device__ __forceinline__ real ld_rec(const real * p, unsigned long long pol) { real v; asm volatile(\"ld.global.cg.L2::cache_hint.f64 %0, [%1], %2;\" : \"=d\"(v) : \"l\"(p), \"l\"(pol) : \"memory\"); return v; }\n"
                  "__device__ __forceinline__ void st_rec(real * p, real v, unsigned long long pol) { asm volatile(\"st.global.cg.L2::cache_hint.f64 [%0], %1, %2;\" ::\"l\"(p), \"d\"(v), \"l\"(pol) : \"memory\"); }\n");
  else
  os << (fp32 ? "__device__ __forceinline__ real ld_rec(const real * p) { real v; asm volatile(\"ld.global.cg.f32 %0, [%1];\" : \"=f\"(v) : \"l\"(p) : \"memory\"); return v; }\n"
              : "__device__ __forceinline__ real ld_rec(const real * p) { real v; asm volatile(\"ld.global.cg.f64 %0, [%1];\" : \"=d\"(v) : \"l\"(p) : \"memory\"); return v; }\n");
  // warp-cooperative tile copies: `rows` elements per configuration, configurations `ld` apart in global memory, `pitch` apart
  // in the tile; lanes beyond the batch end re-read the last valid configuration (their results are never stored)
  // global -> shared with cp.async (LDGSTS): every lane fires all its element copies without waiting for any of them; one
  // wait per tile.  (A plain `t[..] = g[..]` loop serialises one L2 round trip per element: 106 of them per tile.)
  os << "__device__ __forceinline__ void cp_elem(real * dst, const real * src)\n{\n"
        "  asm volatile(\"cp.async.ca.shared.global [%0], [%1], " << (fp32 ? 4 : 8) << ";\" ::\"r\"((unsigned)__cvta_generic_to_shared(dst)), \"l\"(src) : \"memory\");\n}\n";
  os << "__device__ __forceinline__ void tile_in(real * t, int pitch, const real * __restrict__ g, long long ld, int rows, int nvalid, int lane)\n{\n"
        "  if (ld == rows && nvalid == 32)\n  {\n    int c = 0, r = lane;\n    while (r >= rows) { r -= rows; ++c; }\n"
        "    for (int k = lane; k < 32 * rows; k += 32)\n    {\n      cp_elem(t + c * pitch + r, g + k);\n      r += 32;\n"
        "      while (r >= rows) { r -= rows; ++c; }\n    }\n  }\n  else\n"
        "    for (int c = 0; c < 32; ++c)\n    {\n      const long long cs = c < nvalid ? c : nvalid - 1;\n"
        "      for (int r = lane; r < rows; r += 32) cp_elem(t + c * pitch + r, g + cs * ld + r);\n    }\n}\n";
  os << "__device__ __forceinline__ void tile_out(real * __restrict__ g, long long ld, const real * t, int pitch, int rows, int nvalid, int lane)\n{\n"
        "  if (ld == rows && nvalid == 32)\n  {\n    int c = 0, r = lane;\n    while (r >= rows) { r -= rows; ++c; }\n"
        "    for (int k = lane; k < 32 * rows; k += 32)\n    {\n      g[k] = t[c * pitch + r];\n      r += 32;\n"
        "      while (r >= rows) { r -= rows; ++c; }\n    }\n  }\n  else\n"
        "    for (int c = 0; c < nvalid; ++c)\n      for (int r = lane; r < rows; r += 32) g[c * ld + r] = t[c * pitch + r];\n}\n";
  if (direct_io && rec_policy_mode == 2)
  {
    os << (fp32 ? "__device__ __forceinline__ real ld_in(const real * p, unsigned long long pol) { real v; asm volatile(\"ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;\" : \"=f\"(v) : \"l\"(p), \"l\"(pol)); return v; }\n"
                : "__device__ __forceinline__ real ld_in(const real * p, unsigned long long pol) { real v; asm volatile(\"ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;\" : \"=d\"(v) : \"l\"(p), \"l\"(pol)); return v; }\n");
    os << "#define BRBD_IN0(k) ld_in(tq + (k), pol_drop)\n#define BRBD_IN1(k) ld_in(tv + (k), pol_drop)\n#define BRBD_IN2(k) ld_in(tx + (k), pol_drop)\n";
  }
  else if (direct_io)
  {
    os << (fp32 ? "__device__ __forceinline__ real ld_in(const real * p) { real v; asm volatile(\"ld.global.nc.f32 %0, [%1];\" : \"=f\"(v) : \"l\"(p)); return v; }\n"
                : "__device__ __forceinline__ real ld_in(const real * p) { real v; asm volatile(\"ld.global.nc.f64 %0, [%1];\" : \"=d\"(v) : \"l\"(p)); return v; }\n");
    os << "#define BRBD_IN0(k) ld_in(tq + (k))\n#define BRBD_IN1(k) ld_in(tv + (k))\n#define BRBD_IN2(k) ld_in(tx + (k))\n";
  }
  else
    os << "#define BRBD_IN0(k) tq[(k)]\n#define BRBD_IN1(k) tv[(k)]\n#define BRBD_IN2(k) tx[(k)]\n";
  // record store [CTA][slot][thread of the CTA]: a slot of a warp is one coalesced row, and the slot offset is an immediate
  if (rec_policy)
  {
    os << "#define BRBD_REC_ST(k, val) st_rec(rec + (k) * " << nt << ", (val), pol_keep)\n";
    os << "#define BRBD_REC_LD(k) ld_rec(rec + (k) * " << nt << ", pol_drop)\n";
  }
  else
  {
  os << "#define BRBD_REC_ST(k, val) __stcg(rec + (k) * " << nt << ", (val))\n";
  os << "#define BRBD_REC_LD(k) ld_rec(rec + (k) * " << nt << ")\n";
  }
  if (direct_io) os << "#define BRBD_OUT0(row, val) do { if (live) to[(row)] = (val); } while (0)\n#define BRBD_SYNC() __syncthreads()\n";
  else os << "#define BRBD_OUT0(row, val) tx[(row)] = (val)\n#define BRBD_SYNC() __syncthreads()\n";
  const int wpv = fp32 ? 1 : 2;
  for (const auto & sh : st.park_shapes)
  {
    emit_park_helpers(os, sh.first, sh.second, fp32, nt);
    const char sp = sh.second == 0 ? 'T' : 'S';
    const std::string base = sh.second == 0 ? "tm + (s) * " + std::to_string(wpv) : "park + (s) * " + std::to_string(nt);
    os << "#define BRBD_PARK_ST" << sp << sh.first << "(s, ...) park_st" << sp << sh.first << "(" << base << ", __VA_ARGS__)\n";
    os << "#define BRBD_PARK_LD" << sp << sh.first << "(s, ...) park_ld" << sp << sh.first << "(" << base << ", __VA_ARGS__)\n";
  }
  // COPIES identical kernels (brbd_gen_<name>_<c>): the same code at DIFFERENT addresses, launched side by side on a slice of
  // the batch each (launch_gen.cu).  With one copy all 148 SMs stream the same instruction lines from L2 at the same time and a
  // pass of the 35-dof humanoid ABA takes 124 us instead of the 63 us it takes on <= 64 SMs (scripts/gen_fetch_probe.py).
  for (int copy = 0; copy < copies; ++copy)
  {
  os << "extern \"C\" __global__ void __launch_bounds__(" << nt << ", " << minb << ")\nbrbd_gen_" << name << "_" << copy
     << "(const real * __restrict__ q, long long ldq, const real * __restrict__ v, long long ldv, const real * __restrict__ x, long long ldx,\n"
        "     real * __restrict__ out, long long ldo, real * __restrict__ recbase, long long B)\n{\n";
  os << "  extern __shared__ __align__(16) unsigned char smem_raw[];\n  real * smem = reinterpret_cast<real *>(smem_raw);\n";
  os << "  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;\n";
  if (rec_policy)
    os << "  unsigned long long pol_keep, pol_drop;\n"
          "  asm volatile(\"createpolicy.fractional.L2::evict_last.b64 %0, " << keep_fraction << ";\" : \"=l\"(pol_keep));\n"
          "  asm volatile(\"createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\" : \"=l\"(pol_drop));\n";
  if (!direct_io)
  {
    os << "  real * tile_q = smem + warp * " << tile << ";\n  real * tile_v = tile_q + " << 32 * qp << ";\n  real * tile_x = tile_v + " << 32 * vp << ";\n";
    os << "  real * tq = tile_q + lane * " << qp << ";\n  real * tv = tile_v + lane * " << vp << ";\n  real * tx = tile_x + lane * " << vp << ";\n";
  }
  os << "  real * park = smem + " << warps * tile << " + threadIdx.x;\n";
  os << "  unsigned tm = 0;\n";
  if (st.tmem_slots > 0)
  { // the CTA's tensor-memory columns: warp w works in lane quadrant w % 4, warps 4.. in a second column range
    os << "  __shared__ unsigned tmem_base_slot;\n";
    os << "  if (warp == 0)\n  {\n";
    os << "    asm volatile(\"tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], " << tmem_cols
       << ";\" ::\"r\"((unsigned)__cvta_generic_to_shared(&tmem_base_slot)) : \"memory\");\n";
    os << "    asm volatile(\"tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\" ::: \"memory\");\n  }\n";
    os << "  asm volatile(\"tcgen05.fence::before_thread_sync;\" ::: \"memory\");\n  __syncthreads();\n"
          "  asm volatile(\"tcgen05.fence::after_thread_sync;\" ::: \"memory\");\n";
    os << "  tm = tmem_base_slot + ((((unsigned)warp) & 3u) * 32u << 16) + ((unsigned)warp >> 2) * " << st.tmem_slots * wpv << "u;\n";
  }
  if (!direct_io) os << "  for (int k = threadIdx.x; k < " << warps * tile << "; k += " << nt << ") smem[k] = BRBD_C(0.0);\n  __syncthreads();\n";
  os << "  const long long nthreads = (long long)gridDim.x * " << nt << ";\n";
  os << "  real * __restrict__ rec = recbase + (long long)blockIdx.x * " << (long long)nrec * nt << "ll + threadIdx.x;\n";
  os << "  const long long rounds = (B + nthreads - 1) / nthreads;\n";
  os << "  for (long long rd = 0; rd < rounds; ++rd)\n  {\n";
  os << "    const long long cfg0 = rd * nthreads + (long long)blockIdx.x * " << nt << " + warp * 32; // first configuration of the warp's tile\n";
  // a warp whose tile lies beyond the batch still walks the body (the CTA barriers in it count every warp) on whatever its
  // tile holds, and stores nothing
  if (direct_io)
  {
    os << "    const long long cfg_raw = cfg0 + lane;\n    const bool live = cfg_raw < B;\n    const long long cfg = live ? cfg_raw : B - 1;\n";
    os << "    const real * __restrict__ tq = q + cfg * ldq;\n    const real * __restrict__ tv = v + cfg * ldv;\n"
          "    const real * __restrict__ tx = x + cfg * ldx;\n    real * __restrict__ to = out + cfg * ldo;\n";
    if (const char * e = std::getenv("BRBD_GEN_PREFETCH"))
    { // experiment: this round's input columns into L2 (1) / L1 (2) before the straight-line body touches them one by one
      const char * lvl = std::atoi(e) == 2 ? "L1" : "L2";
      const int es = fp32 ? 4 : 8, step = 128 / es;
      for (int k = 0; k < nq; k += step) os << "    asm volatile(\"prefetch.global." << lvl << " [%0];\" ::\"l\"(tq + " << k << "));\n";
      for (int k = 0; k < nv; k += step)
        os << "    asm volatile(\"prefetch.global." << lvl << " [%0];\" ::\"l\"(tv + " << k << "));\n    asm volatile(\"prefetch.global." << lvl << " [%0];\" ::\"l\"(tx + " << k << "));\n";
    }
  }
  else
  {
    os << "    const int nvalid = cfg0 >= B ? 0 : (int)(B - cfg0 < 32 ? B - cfg0 : 32);\n";
    os << "    if (nvalid > 0)\n    {\n";
    os << "      tile_in(tile_q, " << qp << ", q + cfg0 * ldq, ldq, " << nq << ", nvalid, lane);\n";
    os << "      tile_in(tile_v, " << vp << ", v + cfg0 * ldv, ldv, " << nv << ", nvalid, lane);\n";
    os << "      tile_in(tile_x, " << vp << ", x + cfg0 * ldx, ldx, " << nv << ", nvalid, lane);\n    }\n";
    os << "    asm volatile(\"cp.async.wait_all;\" ::: \"memory\");\n    __syncwarp();\n";
  }
  os << "    {\n" << body << "    }\n";
  if (!direct_io)
  {
    os << "    __syncwarp();\n";
    os << "    if (nvalid > 0) tile_out(out + cfg0 * ldo, ldo, tile_x, " << vp << ", " << nv << ", nvalid, lane);\n";
    os << "    __syncwarp();\n";
  }
  os << "  }\n";
  if (st.tmem_slots > 0)
  {
    os << "  asm volatile(\"tcgen05.wait::st.sync.aligned;\" ::: \"memory\");\n";
    os << "  asm volatile(\"tcgen05.fence::before_thread_sync;\" ::: \"memory\");\n  __syncthreads();\n";
    os << "  if (warp == 0)\n    asm volatile(\"tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, " << tmem_cols
       << ";\" ::\"r\"(tmem_base_slot) : \"memory\");\n";
  }
  os << "}\n";
  } // copies
  return os.str();
}

// Device wrappers of the generated CRBA.  Every lane computes the entries of one column of ITS configuration's matrix into
// its row of the warp's staging tile; the tile then leaves as column `col` of 32 consecutive configurations:
//   brbd_gen_crba_0   : the warp writes the 32 segments of nv elements with coalesced stores (any layout); the flush is a small
//                       rolled loop that stays in the instruction cache, only the arithmetic is straight-line;
//   brbd_gen_crba_tma : ONE cp.async.bulk.tensor.2d store per column and warp (two with an odd leading dimension), issued
//                       by one lane and asynchronous — the scheme of crba_tma_kernel (crba_dfs.cuh), including its shifted boxes
//                       for odd nv: same tensor maps (crba_tma_setup), same CrbaTmaGeom.
// q is read directly (nq strided loads per lane).
//   compact staging (pat != nullptr; brbd_gen_crba_0 only): the staging row holds the entries of the group's structural pattern
//                       only (a third of the group for a humanoid: three times the warps fit per SM), the flush puts the zeros
//                       back from a position table in shared memory.  `ldx` carries the output mode: 0 = dense (the caller's
//                       nv x nv matrices), 1 = packed (the rows leave as they are: nnz entries per configuration).
std::string wrap_device_crba(const std::string & body, bool fp32, const cg::EmitStats & st, int nt, int nq, int nv, const std::string & ktable, int nbuf,
                             int group, const cg::CrbaPattern * pat)
{
  std::ostringstream os;
  const int pitch = pat ? (pat->maxrow | 1) : ((nv * group) | 1); // LSU variant: odd pitch, conflict-free rows
  const int epad = nv + (nv & 1); // TMA variant: the box's inner extent (even: rows stay 16-byte aligned)
  os << "// generated by pinocchio_b200 codegen: crba" << (fp32 ? " (FP32)" : " (FP64)") << "\n";
  os << math_macros(fp32) << "#undef BRBD_SINCOS\n" << device_sincos(fp32) << ktable;
  os << (fp32 ? "__device__ __forceinline__ real ld_in(const real * p) { real v; asm volatile(\"ld.global.nc.f32 %0, [%1];\" : \"=f\"(v) : \"l\"(p)); return v; }\n"
              : "__device__ __forceinline__ real ld_in(const real * p) { real v; asm volatile(\"ld.global.nc.f64 %0, [%1];\" : \"=d\"(v) : \"l\"(p)); return v; }\n");
  os << "#define BRBD_IN0(k) ld_in(tq + (k))\n#define BRBD_SYNC()\n";
  os << "struct __align__(64) TensorMap { unsigned long long opaque[16]; };\n";
  // column `col` of the warp's configurations: element k = c * nv + r of the staging buffer goes to gM[c * ldM + r]
  os << "__device__ __noinline__ void flush_col(real * colbuf, real * __restrict__ g, long long ldM, int nvalid, int lane, int len)\n{\n"
        "  __syncwarp();\n  int c = 0, r = lane;\n  while (r >= len) { r -= len; ++c; }\n"
        "  const int total = nvalid * len;\n"
        "#pragma unroll 4\n  for (int k = lane; k < total; k += 32)\n  {\n    g[c * ldM + r] = colbuf[c * " << pitch << " + r];\n    r += 32;\n"
        "    while (r >= len) { r -= len; ++c; }\n  }\n  __syncwarp();\n}\n";
  os << "__device__ __forceinline__ void tma_store_2d(const void * tmap, const void * ssrc, int x, int y)\n{\n"
        "  asm volatile(\"cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];\" ::\"l\"(tmap),\n"
        "               \"r\"((unsigned)__cvta_generic_to_shared(ssrc)), \"r\"(x), \"r\"(y) : \"memory\");\n}\n";
  // ---- LSU variant ----
  if (pat)
  {
    os << "__constant__ short BRBD_CPOS[" << nv * nv << "] = {";
    for (int k = 0; k < nv * nv; ++k) os << (k ? "," : "") << pat->pos[k];
    os << "};\n__constant__ int BRBD_GBASE[" << nv << "] = {";
    for (int k = 0; k < nv; ++k) os << (k ? "," : "") << pat->gbase[k];
    os << "};\n__constant__ int BRBD_GNNZ[" << nv << "] = {";
    for (int k = 0; k < nv; ++k) os << (k ? "," : "") << pat->gnnz[k];
    os << "};\n";
    // dense: element r of the group's run (len = columns * nv contiguous elements of configuration c) is the staged entry
    // tab[r] of that configuration, or a structural zero
    os << "__device__ __noinline__ void flush_dense(const real * colbuf, real * __restrict__ g, long long ldM, int nvalid, int lane, int len, const short * tab)\n{\n"
          "  __syncwarp();\n  int c = 0, r = lane;\n  while (r >= len) { r -= len; ++c; }\n"
          "  const int total = nvalid * len;\n"
          "#pragma unroll 4\n  for (int k = lane; k < total; k += 32)\n  {\n    const int o = tab[r];\n"
          "    g[c * ldM + r] = o >= 0 ? colbuf[c * " << pitch << " + o] : BRBD_C(0.0);\n    r += 32;\n"
          "    while (r >= len) { r -= len; ++c; }\n  }\n  __syncwarp();\n}\n";
    os << "#define BRBD_OUT0(row, val) cb[(row)] = (val)\n#define BRBD_COLBEGIN(col)\n#define BRBD_CLEAR(row)\n";
    os << "#define BRBD_FLUSH(cc) do { if (packed) flush_col(colbuf, gM + BRBD_GBASE[(cc) & 0xffff], ldM, nvalid, lane, BRBD_GNNZ[(cc) & 0xffff]); \\\n"
          "    else flush_dense(colbuf, gM + ((cc) & 0xffff) * " << nv << ", ldM, nvalid, lane, ((cc) >> 16) * " << nv << ", cpos + ((cc) & 0xffff) * " << nv << "); } while (0)\n";
  }
  else
  {
  os << "#define BRBD_OUT0(row, val) cb[(row)] = (val)\n#define BRBD_COLBEGIN(col)\n#define BRBD_CLEAR(row) cb[(row)] = BRBD_C(0.0)\n";
  os << "#define BRBD_FLUSH(cc) do { flush_col(colbuf + buf * " << 32 * pitch << ", gM + ((cc) & 0xffff) * " << nv << ", ldM, nvalid, lane, ((cc) >> 16) * " << nv << "); buf = buf + 1 == " << nbuf
     << " ? 0 : buf + 1; cb = colbuf + buf * " << 32 * pitch << " + lane * " << pitch << "; } while (0)\n";
  }
  os << "extern \"C\" __global__ void __launch_bounds__(" << nt << ", 1)\nbrbd_gen_crba_0"
     << "(const real * __restrict__ q, long long ldq, const real * __restrict__ v, long long ldv, const real * __restrict__ x, long long ldx,\n"
        "     real * __restrict__ out, long long ldM, real * __restrict__ recbase, long long B)\n{\n";
  os << "  extern __shared__ __align__(128) unsigned char smem_raw[];\n  real * smem = reinterpret_cast<real *>(smem_raw);\n";
  os << "  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;\n";
  os << "  real * colbuf = smem + warp * " << nbuf * 32 * pitch << ";\n  real * cb = colbuf + lane * " << pitch << ";\n  int buf = 0;\n";
  os << "  for (int k = lane; k < " << nbuf * 32 * pitch << "; k += 32) colbuf[k] = BRBD_C(0.0);\n  __syncwarp();\n";
  if (pat)
    os << "  __shared__ short cpos[" << nv * nv << "];\n  for (int k = threadIdx.x; k < " << nv * nv << "; k += " << nt << ") cpos[k] = BRBD_CPOS[k];\n"
          "  __syncthreads();\n  const bool packed = ldx != 0;\n  (void)buf;\n";
  os << "  const long long nthreads = (long long)gridDim.x * " << nt << ";\n";
  os << "  const long long rounds = (B + nthreads - 1) / nthreads;\n";
  if (const char * e = std::getenv("BRBD_GEN_CRBA_STAGGER")) // experiment: warps start that many cycles apart (out of phase)
    os << "  if (rounds > 2) { const long long t0_ = clock64(); while (clock64() - t0_ < (long long)((warp * 7 + blockIdx.x) % " << nt / 32 << ") * " << std::atoi(e) << "ll) { } }\n";
  os << "  for (long long rd = 0; rd < rounds; ++rd)\n  {\n";
  os << "    const long long cfg0 = rd * nthreads + (long long)blockIdx.x * " << nt << " + warp * 32;\n";
  os << "    if (cfg0 >= B) continue; // warp-uniform\n";
  os << "    const int nvalid = (int)(B - cfg0 < 32 ? B - cfg0 : 32);\n";
  os << "    const long long cfg = cfg0 + (lane < nvalid ? lane : nvalid - 1);\n";
  os << "    const real * __restrict__ tq = q + cfg * ldq;\n    real * __restrict__ gM = out + cfg0 * ldM;\n";
  os << "    {\n" << body << "    }\n  }\n}\n";
  if (group > 1 || pat) return os.str(); // the tensor-store variant takes one whole column per store
  // ---- TMA variant ----
  os << "#undef BRBD_OUT0\n#undef BRBD_COLBEGIN\n#undef BRBD_CLEAR\n#undef BRBD_FLUSH\n";
  os << "#define BRBD_OUT0(row, val) myrow[cs + (row)] = (val)\n#define BRBD_CLEAR(row) myrow[ps[buf] + (row)] = BRBD_C(0.0)\n";
  // the tile is free again once the engine has read the previous column block out of it; then this column's shift: one
  // element early where the segment starts at an odd element, two where it starts at an even one, none for the plain columns
  os << "#define BRBD_COLBEGIN(col) do { if (lane == 0) asm volatile(\"cp.async.bulk.wait_group.read " << nbuf - 1 << ";\" ::: \"memory\"); __syncwarp(); \\\n"
        "    if (odd) { const int a_ = (half + ((col) & 0xffff)) & 1; plain = ((col) & 0xffff) == 0 || ((col) == " << nv - 1 << " && a_ == 0); cs = plain ? 0 : (a_ ? 1 : 2); } } while (0)\n";
  os << "#define BRBD_FLUSH(cc) BRBD_FLUSH1(((cc) & 0xffff))\n";
  os << "#define BRBD_FLUSH1(col) do { asm volatile(\"fence.proxy.async.shared::cta;\" ::: \"memory\"); __syncwarp(); \\\n"
        "    if (odd) { \\\n"
        "      const int a0_ = (col) & 1, a1_ = ((col) + 1) & 1; \\\n"
        "      const bool plain0_ = (col) == 0 || ((col) == " << nv - 1 << " && a0_ == 0), plain1_ = (col) == 0 || ((col) == " << nv - 1 << " && a1_ == 0); \\\n"
        "      if (lane == 0) { \\\n"
        "        if (pairs) { const int y_ = (int)(cfg0 >> 1); \\\n"
        "          if (!plain0_) tma_store_2d(&map0, emb, (col) * " << nv << " - (a0_ ? 1 : 2), y_); \\\n"
        "          if (!plain1_ && cfg0 + 1 < B) tma_store_2d(&map1, emb + " << 16 * epad << ", (col) * " << nv << " - (a1_ ? 1 : 2) + 1, y_); } \\\n"
        "        else if (!plain0_) tma_store_2d(&map0, emb, (col) * " << nv << " - (a0_ ? 1 : 2), (int)cfg0); \\\n"
        "        asm volatile(\"cp.async.bulk.commit_group;\" ::: \"memory\"); } \\\n"
        "      if (plain && live) { real * __restrict__ g_ = out + cfg * ldM + (long long)(col) * " << nv << "; for (int e_ = 0; e_ < " << nv << "; ++e_) g_[e_] = myrow[e_]; } \\\n"
        "    } else { \\\n"
        "      if (lane == 0) { tma_store_2d(&map0, emb, (col) * " << nv << ", (int)cfg0); asm volatile(\"cp.async.bulk.commit_group;\" ::: \"memory\"); } } \\\n"
        "    ps[buf] = cs; buf = buf + 1 == " << nbuf << " ? 0 : buf + 1; emb = em + buf * " << 32 * epad << "; myrow = emb + rowoff; } while (0)\n";
  os << "extern \"C\" __global__ void __launch_bounds__(" << nt << ", 1)\nbrbd_gen_crba_tma"
     << "(const real * __restrict__ q, long long ldq, real * __restrict__ out, long long ldM, long long B, int odd, int pairs,\n"
        "     const __grid_constant__ TensorMap map0, const __grid_constant__ TensorMap map1)\n{\n";
  os << "  extern __shared__ __align__(128) unsigned char smem_raw[];\n  real * smem = reinterpret_cast<real *>(smem_raw);\n";
  os << "  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;\n";
  os << "  real * em = smem + warp * " << nbuf * 32 * epad << ";\n  real * emb = em;\n";
  os << "  const int half = (odd && pairs) ? (lane & 1) : 0;\n";
  os << "  const int rowoff = ((odd && pairs) ? (lane >> 1) + 16 * half : lane) * " << epad << ";\n  real * myrow = em + rowoff;\n";
  os << "  for (int b = 0; b < " << nbuf << "; ++b) for (int k = 0; k < " << epad << "; ++k) em[b * " << 32 * epad << " + rowoff + k] = BRBD_C(0.0);\n  __syncwarp();\n";
  os << "  int ps[" << nbuf << "] = {0}, cs = 0, buf = 0;\n  bool plain = false;\n";
  os << "  const long long nthreads = (long long)gridDim.x * " << nt << ";\n";
  os << "  const long long rounds = (B + nthreads - 1) / nthreads;\n";
  os << "  for (long long rd = 0; rd < rounds; ++rd)\n  {\n";
  os << "    const long long cfg0 = rd * nthreads + (long long)blockIdx.x * " << nt << " + warp * 32;\n";
  os << "    if (cfg0 >= B) continue; // warp-uniform\n";
  os << "    const bool live = cfg0 + lane < B;\n    const long long cfg = live ? cfg0 + lane : B - 1;\n";
  os << "    const real * __restrict__ tq = q + cfg * ldq;\n";
  os << "    {\n" << body << "    }\n  }\n";
  os << "  if (lane == 0) asm volatile(\"cp.async.bulk.wait_group 0;\" ::: \"memory\");\n  __syncwarp();\n}\n";
  (void)st; (void)nq;
  return os.str();
}

// Device wrapper of the generated CRBA, bulk-copy variant (brbd_gen_crba_0, same arguments as the LSU variant).  Every lane
// stages `group` adjacent columns of ITS configuration — group * nv contiguous elements of the caller's matrix — in its own row
// of shared memory and hands the row to the copy engine with ONE cp.async.bulk.global.shared::cta: the store is asynchronous
// (the lane goes on with the next columns in another of its `nbuf` rows), lane-local (no warp synchronisation: the lane that
// wrote the row issues the copy) and one run of >= 600 bytes per copy.  Measured (scripts/micro/bulk_store_bw.cu): such copies
// from 4-8 warps per SM sustain 5.2-5.7 TB/s into this very layout, coalesced STG from as few warps 2.5-4.6 TB/s, and both
// fall with the number of resident warps (more configurations, i.e. more DRAM pages, written at the same time).
// A bulk copy needs 16-byte aligned source, destination and size: the row is filled `sh` elements in, sh = the destination's
// misalignment in elements (odd nv: every other configuration / column group), so that source and destination are congruent;
// the few elements before / after the aligned interior leave through plain stores.
std::string wrap_device_crba_bulk(const std::string & body, bool fp32, int nt, int nq, int nv, const std::string & ktable, int nbuf, int group)
{
  std::ostringstream os;
  const int A = fp32 ? 4 : 2; // elements per 16 bytes
  const int pitch = crba_bulk_pitch(nv, group, fp32); // rows 2 * odd (4 * odd) elements apart: at most 2-way (4-way) bank conflicts on the row writes
  os << "// generated by pinocchio_b200 codegen: crba" << (fp32 ? " (FP32)" : " (FP64)") << ", bulk-copy variant\n";
  os << math_macros(fp32) << "#undef BRBD_SINCOS\n" << device_sincos(fp32) << ktable;
  os << (fp32 ? "__device__ __forceinline__ real ld_in(const real * p) { real v; asm volatile(\"ld.global.nc.f32 %0, [%1];\" : \"=f\"(v) : \"l\"(p)); return v; }\n"
              : "__device__ __forceinline__ real ld_in(const real * p) { real v; asm volatile(\"ld.global.nc.f64 %0, [%1];\" : \"=d\"(v) : \"l\"(p)); return v; }\n");
  os << "#define BRBD_IN0(k) ld_in(tq + (k))\n#define BRBD_SYNC()\n";
  // row[sh + e] -> g[e], e in [0, L): plain stores for the unaligned head / tail, one bulk copy for the rest
  // the copies carry an L2 evict_first policy — M is written once and never read again here (BRBD_GEN_CRBA_POLICY=0 turns it off;
  // measured, profiles/r2_crba_out_policy.txt: 65 536 x simple_humanoid 0.161 -> 0.156 ms, talos 0.186 -> 0.167 ms, 2^20 unchanged)
  const bool out_policy = !(std::getenv("BRBD_GEN_CRBA_POLICY") && std::atoi(std::getenv("BRBD_GEN_CRBA_POLICY")) == 0);
  os << "__device__ __forceinline__ void bulk_flush(const real * row, int sh, real * __restrict__ g, int L, bool live" << (out_policy ? ", unsigned long long pol" : "") << ")\n{\n"
        "  const int e0 = (" << A << " - sh) & " << A - 1 << ", n = (L - e0) & ~" << A - 1 << ";\n"
        "  if (live)\n  {\n"
        "    for (int e = 0; e < e0; ++e) g[e] = row[sh + e];\n"
        "    for (int e = e0 + n; e < L; ++e) g[e] = row[sh + e];\n"
        "    asm volatile(\"fence.proxy.async.shared::cta;\" ::: \"memory\");\n"
        "    if (n > 0)\n"
     << (out_policy ? "      asm volatile(\"cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;\" ::\"l\"(g + e0), \"r\"((unsigned)__cvta_generic_to_shared(row + sh + e0)),\n"
                      : "      asm volatile(\"cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\" ::\"l\"(g + e0), \"r\"((unsigned)__cvta_generic_to_shared(row + sh + e0)),\n")
     << "                   \"r\"(n * " << (fp32 ? 4 : 8) << ")" << (out_policy ? ", \"l\"(pol)" : "") << " : \"memory\");\n"
        "  }\n"
        "  asm volatile(\"cp.async.bulk.commit_group;\" ::: \"memory\");\n}\n";
  os << "#define BRBD_OUT0(row, val) myrow[sh + (row)] = (val)\n#define BRBD_CLEAR(row) myrow[ps[buf] + (row)] = BRBD_C(0.0)\n";
  // the row is free again once the copy engine has read the group of nbuf flushes ago out of it
  os << "#define BRBD_COLBEGIN(col) do { asm volatile(\"cp.async.bulk.wait_group.read " << nbuf - 1 << ";\" ::: \"memory\"); sh = (par + ((col) & 0xffff) * " << nv << ") & " << A - 1 << "; } while (0)\n";
  os << "#define BRBD_FLUSH(cc) do { bulk_flush(myrow, sh, gcfg + ((cc) & 0xffff) * " << nv << ", ((cc) >> 16) * " << nv << ", live" << (out_policy ? ", pol_out" : "") << "); ps[buf] = sh; \\\n"
        "    buf = buf + 1 == " << nbuf << " ? 0 : buf + 1; myrow = em + buf * " << 32 * pitch << " + lane * " << pitch << "; } while (0)\n";
  os << "extern \"C\" __global__ void __launch_bounds__(" << nt << ", 1)\nbrbd_gen_crba_0"
     << "(const real * __restrict__ q, long long ldq, const real * __restrict__ v, long long ldv, const real * __restrict__ x, long long ldx,\n"
        "     real * __restrict__ out, long long ldM, real * __restrict__ recbase, long long B)\n{\n";
  os << "  extern __shared__ __align__(128) unsigned char smem_raw[];\n  real * smem = reinterpret_cast<real *>(smem_raw);\n";
  os << "  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;\n";
  os << "  real * em = smem + warp * " << nbuf * 32 * pitch << ";\n  real * myrow = em + lane * " << pitch << ";\n";
  os << "  for (int k = lane; k < " << nbuf * 32 * pitch << "; k += 32) em[k] = BRBD_C(0.0);\n  __syncwarp();\n";
  os << "  int ps[" << nbuf << "] = {0}, sh = 0, buf = 0;\n";
  if (out_policy) os << "  unsigned long long pol_out;\n  asm volatile(\"createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\" : \"=l\"(pol_out));\n";
  os << "  const long long nthreads = (long long)gridDim.x * " << nt << ";\n";
  os << "  const long long rounds = (B + nthreads - 1) / nthreads;\n";
  os << "  for (long long rd = 0; rd < rounds; ++rd)\n  {\n";
  os << "    const long long cfg0 = rd * nthreads + (long long)blockIdx.x * " << nt << " + warp * 32;\n";
  os << "    if (cfg0 >= B) continue; // warp-uniform\n";
  os << "    const bool live = cfg0 + lane < B;\n    const long long cfg = live ? cfg0 + lane : B - 1;\n";
  os << "    const real * __restrict__ tq = q + cfg * ldq;\n    real * __restrict__ gcfg = out + cfg * ldM;\n";
  if (const char * e = std::getenv("BRBD_GEN_PREFETCH")) // experiment: the lane's q column of the next round into L2 (1) / of this round into L1 (2)
  {
    const int mode = std::atoi(e);
    const int es = fp32 ? 4 : 8, step = 128 / es;
    if (mode == 1)
    {
      os << "    if (cfg + nthreads < B)\n    {\n      const real * nq_ = q + (cfg + nthreads) * ldq;\n";
      for (int k = 0; k < nq; k += step) os << "      asm volatile(\"prefetch.global.L2 [%0];\" ::\"l\"(nq_ + " << k << "));\n";
      os << "      asm volatile(\"prefetch.global.L2 [%0];\" ::\"l\"(nq_ + " << nq - 1 << "));\n    }\n";
    }
    else if (mode == 2)
    {
      for (int k = 0; k < nq; k += step) os << "    asm volatile(\"prefetch.global.L1 [%0];\" ::\"l\"(tq + " << k << "));\n";
      os << "    asm volatile(\"prefetch.global.L1 [%0];\" ::\"l\"(tq + " << nq - 1 << "));\n";
    }
  }
  os << "    const int par = (int)((reinterpret_cast<unsigned long long>(gcfg) / " << (fp32 ? 4 : 8) << "ull) & " << A - 1 << "ull);\n";
  os << "    {\n" << body << "    }\n  }\n";
  os << "  asm volatile(\"cp.async.bulk.wait_group 0;\" ::: \"memory\");\n}\n";
  return os.str();
}

// Device wrapper of the generated derivative kernels (small models): every lane reads its own q / v / x column.  Results:
// `staged` — each lane fills its rows of three per-warp tiles (32 x nv^2) in shared memory and hands every row to the copy
// engine with one cp.async.bulk when the configuration is finished: asynchronous, lane-local, one run of nv^2 elements per copy
// (the scheme of wrap_device_crba_bulk, including the alignment shift; first version: the warp wrote the tiles with coalesced
// stores — 906 MB per algorithm at 2^20 x manipulator at the 2.2 TB/s that pattern reaches from 8 warps,
// profiles/r2_bulk_store_bw.txt); otherwise every lane stores its own results directly.
std::string wrap_device_derivs(const std::string & body, const char * name, bool fp32, int nt, const std::string & ktable, int nv, bool staged)
{
  std::ostringstream os;
  const int nn = nv * nv, pitch = crba_bulk_pitch(nn, 1, fp32), vp = crba_bulk_pitch(nv, 1, fp32), tile = 32 * (3 * pitch + vp);
  const int A = fp32 ? 4 : 2, es = fp32 ? 4 : 8;
  os << "// generated by pinocchio_b200 codegen: " << name << (fp32 ? " (FP32)" : " (FP64)") << "\n";
  os << math_macros(fp32) << "#undef BRBD_SINCOS\n" << device_sincos(fp32) << ktable;
  os << (fp32 ? "__device__ __forceinline__ real ld_in(const real * p) { real v; asm volatile(\"ld.global.nc.f32 %0, [%1];\" : \"=f\"(v) : \"l\"(p)); return v; }\n"
              : "__device__ __forceinline__ real ld_in(const real * p) { real v; asm volatile(\"ld.global.nc.f64 %0, [%1];\" : \"=d\"(v) : \"l\"(p)); return v; }\n");
  os << "#define BRBD_IN0(k) ld_in(tq + (k))\n#define BRBD_IN1(k) ld_in(tv + (k))\n#define BRBD_IN2(k) ld_in(tx + (k))\n#define BRBD_SYNC()\n";
  if (staged)
  {
    // row[sh + e] -> g[e], e in [0, L): plain stores for the unaligned head / tail, one bulk copy for the rest
    os << "__device__ __forceinline__ int misalign(const real * g) { return (int)((reinterpret_cast<unsigned long long>(g) / " << es << "ull) & " << A - 1 << "ull); }\n";
    // the copies carry an L2 evict_first policy: the results are written once, and what should stay in L2 are the lanes' local-memory
    // spill slots, which the result stream otherwise pushes out to DRAM (ncu of C3: 3.4 GB written for 0.96 GB of results)
    os << "__device__ __forceinline__ void bulk_flush(const real * row, int sh, real * __restrict__ g, int L, bool live, unsigned long long pol)\n{\n"
          "  const int e0 = (" << A << " - sh) & " << A - 1 << ", n = (L - e0) & ~" << A - 1 << ";\n"
          "  if (live)\n  {\n"
          "    for (int e = 0; e < e0; ++e) g[e] = row[sh + e];\n"
          "    for (int e = e0 + n; e < L; ++e) g[e] = row[sh + e];\n"
          "    asm volatile(\"fence.proxy.async.shared::cta;\" ::: \"memory\");\n"
          "    if (n > 0)\n"
          "      asm volatile(\"cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;\" ::\"l\"(g + e0), \"r\"((unsigned)__cvta_generic_to_shared(row + sh + e0)),\n"
          "                   \"r\"(n * " << es << "), \"l\"(pol) : \"memory\");\n"
          "  }\n"
          "  asm volatile(\"cp.async.bulk.commit_group;\" ::: \"memory\");\n}\n";
    for (int k = 0; k < 4; ++k) os << "#define BRBD_OUT" << k << "(i, val) s" << k << "[sh" << k << " + (i)] = (val)\n";
  }
  else
    for (int k = 0; k < 4; ++k) os << "#define BRBD_OUT" << k << "(i, val) do { if (live" << (k == 3 ? " && o3" : "") << ") p" << k << "[(i)] = (val); } while (0)\n";
  os << "extern \"C\" __global__ void __launch_bounds__(" << nt << ", 1)\nbrbd_gen_" << name << "_0"
     << "(const real * __restrict__ q, long long ldq, const real * __restrict__ v, long long ldv, const real * __restrict__ x, long long ldx,\n"
        "     real * __restrict__ o0, long long ld0, real * __restrict__ o1, long long ld1, real * __restrict__ o2, long long ld2, real * __restrict__ o3, long long ld3, long long B)\n{\n";
  os << "  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;\n";
  if (staged)
  {
    os << "  extern __shared__ __align__(128) unsigned char smem_raw[];\n  real * tile = reinterpret_cast<real *>(smem_raw) + warp * " << tile << ";\n";
    os << "  real * s0 = tile + lane * " << pitch << ";\n  real * s1 = s0 + " << 32 * pitch << ";\n  real * s2 = s1 + " << 32 * pitch << ";\n"
          "  real * s3 = tile + " << 3 * 32 * pitch << " + lane * " << vp << ";\n";
    os << "  unsigned long long pol_out;\n  asm volatile(\"createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\" : \"=l\"(pol_out));\n";
  }
  os << "  const long long nthreads = (long long)gridDim.x * " << nt << ";\n";
  os << "  for (long long cfg0 = (long long)blockIdx.x * " << nt << " + warp * 32; cfg0 < B; cfg0 += nthreads)\n  {\n";
  os << "    const long long cfg_raw = cfg0 + lane;\n    const bool live = cfg_raw < B;\n    const long long cfg = live ? cfg_raw : B - 1;\n";
  os << "    const int nvalid = (int)(B - cfg0 < 32 ? B - cfg0 : 32);\n";
  os << "    const real * __restrict__ tq = q + cfg * ldq;\n    const real * __restrict__ tv = v + cfg * ldv;\n    const real * __restrict__ tx = x + cfg * ldx;\n";
  os << "    real * __restrict__ p0 = o0 + cfg * ld0;\n    real * __restrict__ p1 = o1 + cfg * ld1;\n    real * __restrict__ p2 = o2 + cfg * ld2;\n"
        "    real * __restrict__ p3 = o3 ? o3 + cfg * ld3 : o3;\n";
  if (staged) // the rows are free once the copy engine has read the previous configuration out of them
    os << "    asm volatile(\"cp.async.bulk.wait_group.read 0;\" ::: \"memory\");\n"
          "    const int sh0 = misalign(p0), sh1 = misalign(p1), sh2 = misalign(p2), sh3 = misalign(p3);\n";
  os << "    {\n" << body << "    }\n";
  if (staged)
  {
    os << "    bulk_flush(s0, sh0, p0, " << nn << ", live, pol_out);\n    bulk_flush(s1, sh1, p1, " << nn << ", live, pol_out);\n    bulk_flush(s2, sh2, p2, " << nn << ", live, pol_out);\n";
    os << "    if (o3) bulk_flush(s3, sh3, p3, " << nv << ", live, pol_out);\n";
  }
  os << "    (void)live; (void)nvalid; (void)p0; (void)p1; (void)p2; (void)p3;\n  }\n";
  if (staged) os << "  asm volatile(\"cp.async.bulk.wait_group 0;\" ::: \"memory\");\n";
  os << "}\n";
  return os.str();
}

std::string wrap_host(const std::string & body, const char * name, bool fp32, const cg::EmitStats & st, const std::string & ktable, int nv,
                      const cg::CrbaPattern * pat = nullptr)
{
  std::ostringstream os;
  os << "#define BRBD_NV " << nv << "\n";
  os << "// generated by pinocchio_b200 codegen (host variant, C++, tests only): " << name << "\n#include <math.h>\n";
  os << math_macros(fp32) << ktable;
  os << "#define BRBD_IN0(k) qc[(k)]\n#define BRBD_IN1(k) vc[(k)]\n#define BRBD_IN2(k) xc[(k)]\n";
  os << "#define BRBD_REC_ST(k, val) rec[(k)] = (val)\n#define BRBD_REC_LD(k) rec[(k)]\n#define BRBD_OUT0(row, val) oc[(row)] = (val)\n";
  for (const auto & sh : st.park_shapes)
  { // plain C: one pair of functions per group shape; tensor-memory groups sit after the shared-memory ones in `park`
    const char sp = sh.second == 0 ? 'T' : 'S';
    const int off = sh.second == 0 ? st.smem_slots : 0;
    os << "static void park_st" << sp << sh.first << "(real * a";
    for (int k = 0; k < sh.first; ++k) os << ", real v" << k;
    os << ") {";
    for (int k = 0; k < sh.first; ++k) os << " a[" << k << "] = v" << k << ";";
    os << " }\nstatic void park_ld" << sp << sh.first << "(const real * a";
    for (int k = 0; k < sh.first; ++k) os << ", real & v" << k;
    os << ") {";
    for (int k = 0; k < sh.first; ++k) os << " v" << k << " = a[" << k << "];";
    os << " }\n";
    os << "#define BRBD_PARK_ST" << sp << sh.first << "(s, ...) park_st" << sp << sh.first << "(park + " << off << " + (s), __VA_ARGS__)\n";
    os << "#define BRBD_PARK_LD" << sp << sh.first << "(s, ...) park_ld" << sp << sh.first << "(park + " << off << " + (s), __VA_ARGS__)\n";
  }
  if (std::string(name).find("derivatives") != std::string::npos)
  { // oc = [block 0 | block 1 | block 2 | vector]
    os << "#undef BRBD_OUT0\n#define BRBD_OUT0(i, val) oc[(i)] = (val)\n#define BRBD_OUT1(i, val) oc[BRBD_NV * BRBD_NV + (i)] = (val)\n"
          "#define BRBD_OUT2(i, val) oc[2 * BRBD_NV * BRBD_NV + (i)] = (val)\n#define BRBD_OUT3(i, val) oc[3 * BRBD_NV * BRBD_NV + (i)] = (val)\n";
    os << "extern \"C\" void brbd_gen_" << name << "_host(const real * qc, const real * vc, const real * xc, real * oc, real * rec, real * park)\n{\n";
    os << body << "}\n";
    return os.str();
  }
  if (std::string(name) == "crba" && pat)
  { // compact staging: oc = [dense nv x nv matrix | packed entries]; the flush expands the group's row through the position table
    // exactly as the device wrapper's flush_dense / flush_col do
    os << "static const short BRBD_CPOS[" << nv * nv << "] = {";
    for (int k = 0; k < nv * nv; ++k) os << (k ? "," : "") << pat->pos[k];
    os << "};\nstatic const int BRBD_GBASE[" << nv << "] = {";
    for (int k = 0; k < nv; ++k) os << (k ? "," : "") << pat->gbase[k];
    os << "};\nstatic const int BRBD_GNNZ[" << nv << "] = {";
    for (int k = 0; k < nv; ++k) os << (k ? "," : "") << pat->gnnz[k];
    os << "};\n";
    os << "#undef BRBD_OUT0\n#define BRBD_OUT0(row, val) colbuf[(row)] = (val)\n#define BRBD_COLBEGIN(col) for (int r_ = 0; r_ < " << (pat->maxrow | 1) << "; ++r_) colbuf[r_] = nan(\"\")\n#define BRBD_CLEAR(row)\n"
          "#define BRBD_FLUSH(cc) do { const int lo_ = (cc) & 0xffff, len_ = ((cc) >> 16) * BRBD_NV; \\\n"
          "    for (int r_ = 0; r_ < len_; ++r_) { const int o_ = BRBD_CPOS[lo_ * BRBD_NV + r_]; oc[lo_ * BRBD_NV + r_] = o_ >= 0 ? colbuf[o_] : BRBD_C(0.0); } \\\n"
          "    for (int o_ = 0; o_ < BRBD_GNNZ[lo_]; ++o_) oc[BRBD_NV * BRBD_NV + BRBD_GBASE[lo_] + o_] = colbuf[o_]; } while (0)\n";
    os << "extern \"C\" void brbd_gen_crba_host(const real * qc, const real * vc, const real * xc, real * oc, real * rec, real * park)\n{\n"
          "  real colbuf[" << (pat->maxrow | 1) << "];\n";
    os << body << "}\n";
    return os.str();
  }
  if (std::string(name) == "crba")
  { // matrix output: BRBD_OUT0 fills the staging row of the current column, BRBD_FLUSH copies it into column `col` of oc
    os << "#undef BRBD_OUT0\n#define BRBD_OUT0(row, val) colbuf[(row)] = (val)\n#define BRBD_COLBEGIN(col)\n#define BRBD_CLEAR(row) colbuf[(row)] = BRBD_C(0.0)\n"
          "#define BRBD_FLUSH(cc) for (int r_ = 0; r_ < BRBD_NV; ++r_) oc[((cc) & 0xffff) * BRBD_NV + r_] = colbuf[r_]\n";
    os << "extern \"C\" void brbd_gen_crba_host(const real * qc, const real * vc, const real * xc, real * oc, real * rec, real * park)\n{\n"
          "  static real colbuf[BRBD_NV]; // keeps the last column of the previous call, as a lane's staging row does from round to round\n";
    os << body << "}\n";
    return os.str();
  }
  os << "extern \"C\" void brbd_gen_" << name << "_host(const real * qc, const real * vc, const real * xc, real * oc, real * rec, real * park)\n{\n";
  os << body << "}\n";
  return os.str();
}

// Host emulation of ONE LANE of wrap_device_crba_bulk (tests of the generator only): the staging rows persist from call to call as
// a lane's rows do from round to round of the persistent grid, every call may find its matrix at another misalignment
// (brbd_gen_crba_host_par, set by the test: the `par` of the device wrapper), a group is staged `sh` elements in and leaves as
// row[sh .. sh + L), and what a row still holds from `nbuf` groups ago — written at THAT group's shift — is cleared first.
std::string wrap_host_crba_bulk(const std::string & body, bool fp32, const std::string & ktable, int nv, int nbuf, int group)
{
  std::ostringstream os;
  const int A = fp32 ? 4 : 2, pitch = crba_bulk_pitch(nv, group, fp32);
  os << "#define BRBD_NV " << nv << "\n// generated by pinocchio_b200 codegen (host emulation of the bulk-copy CRBA wrapper, tests only)\n#include <math.h>\n";
  os << math_macros(fp32) << ktable;
  os << "#define BRBD_IN0(k) qc[(k)]\n#define BRBD_IN1(k) vc[(k)]\n#define BRBD_IN2(k) xc[(k)]\n";
  os << "extern \"C\" { int brbd_gen_crba_host_par = 0; }\n";
  os << "static real em[" << nbuf << "][" << pitch << "];\nstatic int ps[" << nbuf << "], buf = 0;\n";
  os << "#define BRBD_OUT0(row, val) myrow[sh + (row)] = (val)\n#define BRBD_CLEAR(row) myrow[ps[buf] + (row)] = BRBD_C(0.0)\n";
  os << "#define BRBD_COLBEGIN(col) sh = (brbd_gen_crba_host_par + ((col) & 0xffff) * BRBD_NV) & " << A - 1 << "\n";
  os << "#define BRBD_FLUSH(cc) do { const int lo_ = (cc) & 0xffff, len_ = ((cc) >> 16) * BRBD_NV; \\\n"
        "    for (int e_ = 0; e_ < len_; ++e_) oc[lo_ * BRBD_NV + e_] = myrow[sh + e_]; \\\n"
        "    ps[buf] = sh; buf = buf + 1 == " << nbuf << " ? 0 : buf + 1; myrow = em[buf]; } while (0)\n";
  os << "extern \"C\" void brbd_gen_crba_host(const real * qc, const real * vc, const real * xc, real * oc, real * rec, real * park)\n{\n"
        "  real * myrow = em[buf];\n  int sh = 0;\n";
  os << body << "  (void)sh; (void)myrow;\n}\n";
  return os.str();
}
} // namespace

namespace brbd
{
int crba_pattern_nnz(const brbd_model & m) { return cg::crba_pattern(m.pd, 1).nnz; }
void crba_pattern_index(const brbd_model & m, std::vector<int32_t> & idx)
{
  const cg::CrbaPattern pat = cg::crba_pattern(m.pd, 1);
  idx.resize((size_t)pat.nnz);
  for (int k = 0; k < pat.nnz; ++k) idx[k] = pat.cols[k] * m.pd.nv + pat.rows[k];
}
} // namespace brbd

extern "C" {

brbd_status brbd_codegen_source(const brbd_model * m, int algo, int flags, char ** source, brbd_codegen_info * info)
{
  if (!m || !source) return fail(BRBD_EINVAL, "null argument");
  *source = nullptr;
  if (algo < BRBD_GEN_RNEA || algo > BRBD_GEN_ABA_DERIVATIVES) return fail(BRBD_EINVAL, "code generation: unknown algorithm");
  if (algo >= BRBD_GEN_RNEA_DERIVATIVES && m->pd.nv > 16)
    return fail(BRBD_EINVAL, "code generation: the derivative programs keep all 3 nv^2 results alive and are limited to nv <= 16");
  // CRBA: staging tiles per warp.  Rotating over 2 or 3 (so that a tensor store drains while the next column is assembled)
  // was measured and does not pay — 65 536 x simple_humanoid: 0.233 ms with one tile and 16 warps, 0.287 with two and 12
  // (profiles/r2_gen_crba_experiments.txt): the store path, not the wait for the tile, bounds the kernel
  int crba_nbuf = 1;
  if (const char * e = std::getenv("BRBD_GEN_CRBA_NBUF")) crba_nbuf = std::max(1, std::min(4, std::atoi(e)));
  if (flags & BRBD_GEN_HOST) crba_nbuf = 1;
  // CRBA: adjacent columns per flush (see trace_crba); the host variant and the TMA variant take one column at a time
  int crba_group = (flags >> 24) & 0x1f ? (flags >> 24) & 0x1f : 1;
  if ((flags & BRBD_GEN_HOST) && !(flags & (BRBD_GEN_CRBA_COMPACT | BRBD_GEN_CRBA_BULK))) crba_group = 1;
  // CRBA, compact staging: only the entries of the structural pattern are staged (see CrbaPattern); one staging tile per warp
  const bool crba_compact = algo == BRBD_GEN_CRBA && (flags & BRBD_GEN_CRBA_COMPACT) != 0;
  // (the flags' group field then counts ENTRIES per group, in units of 8, instead of columns)
  int crba_budget = 0;
  if (crba_compact) { crba_nbuf = 1; crba_budget = 8 * ((flags >> 24) & 0x1f ? (flags >> 24) & 0x1f : 1); crba_group = 31; }
  cg::CrbaPattern pat;
  if (crba_compact) pat = cg::crba_pattern(m->pd, crba_group, crba_budget);
  // CRBA, bulk-copy variant: `group` adjacent columns per lane and copy, rotating over bits 29..30 (+1) staging rows
  const bool crba_bulk = algo == BRBD_GEN_CRBA && !crba_compact && (flags & BRBD_GEN_CRBA_BULK) != 0 && !(flags & BRBD_GEN_HOST);
  if (crba_bulk) crba_nbuf = ((flags >> 29) & 3) + 1;
  // host emulation of the bulk-copy wrapper's staging (tests of the generator): same groups, same rotation of staging rows
  const bool crba_bulk_host = algo == BRBD_GEN_CRBA && !crba_compact && (flags & BRBD_GEN_CRBA_BULK) != 0 && (flags & BRBD_GEN_HOST) != 0;
  if (crba_bulk_host) crba_nbuf = ((flags >> 29) & 3) + 1;
  cg::Tracer T(m->pd, (flags & BRBD_GEN_EXPLICIT_SLOTS) != 0);
  if (algo == BRBD_GEN_ABA) cg::trace_aba(T);
  else if (algo == BRBD_GEN_CRBA) cg::trace_crba(T, crba_nbuf, crba_group, crba_compact, crba_budget);
  else if (algo == BRBD_GEN_RNEA_DERIVATIVES) cg::trace_rnea_derivatives(T);
  else if (algo == BRBD_GEN_ABA_DERIVATIVES) cg::trace_aba_derivatives(T);
  else cg::trace_rnea(T);
  cg::EmitStats st;
  // derivative programs: results through per-warp tiles when a warp's three nv^2 blocks fit beside those of its CTA's other warps
  const size_t derivs_tile_bytes = (size_t)32 * (3 * crba_bulk_pitch(m->pd.nv * m->pd.nv, 1, (flags & BRBD_GEN_FP32) != 0) + crba_bulk_pitch(m->pd.nv, 1, (flags & BRBD_GEN_FP32) != 0)) *
                                   ((flags & BRBD_GEN_FP32) ? 4 : 8);
  const bool derivs_staged = derivs_tile_bytes * ((((flags >> 8) & 0xfff) ? ((flags >> 8) & 0xfff) : 128) / 32) <= 220 * 1024;
  const int nt = (flags >> 8) & 0xfff ? (flags >> 8) & 0xfff : 128;
  const int minb = (flags >> 20) & 0xf ? (flags >> 20) & 0xf : 1;
  const bool fp32 = (flags & BRBD_GEN_FP32) != 0;
  const bool direct_io = (flags & BRBD_GEN_DIRECT_IO) != 0;
  // tensor memory: 512 columns of 32 bit per SM, shared by the resident CTAs; a CTA's warps 0-3 / 4-7 / ... use their lane
  // quadrant, so the columns split over ceil(warps / 4) ranges
  const int wpv = fp32 ? 1 : 2, ranges = (nt / 32 + 3) / 4;
  const int cols_per_cta = 512 / (minb > 0 ? minb : 1);
  int tmem_cols = 32;
  while (tmem_cols * 2 <= cols_per_cta) tmem_cols *= 2;
  const int tmem_capacity = (flags & BRBD_GEN_EXPLICIT_SLOTS) ? tmem_cols / ranges / wpv : 0;
  // copies of the body at different code addresses (see wrap_device): 1 below ~6000 statements, 3 above
  int copies = 1;
  if (const char * e = std::getenv("BRBD_GEN_COPIES")) copies = std::max(1, std::min(8, std::atoi(e)));
  int sync_every = 0;
  if (const char * e = std::getenv("BRBD_GEN_SYNC")) sync_every = std::atoi(e);
  if (flags & BRBD_GEN_HOST) sync_every = 0;
  cg::ConstTable K;
  const std::string body = cg::emit_body(T.g, st, K, tmem_capacity, sync_every);
  if (st.tmem_slots > 0)
  { // allocate no more columns than the kernel uses
    int need = 32;
    while (need < st.tmem_slots * wpv * ranges) need *= 2;
    tmem_cols = need;
  }
  const std::string src = (algo >= BRBD_GEN_RNEA_DERIVATIVES && !(flags & BRBD_GEN_HOST))
                            ? wrap_device_derivs(body, algo_name(algo), fp32, nt, K.definition("__constant__"), m->pd.nv, derivs_staged)
                            : crba_bulk ? wrap_device_crba_bulk(body, fp32, nt, m->pd.nq, m->pd.nv, K.definition("__constant__"), crba_nbuf, crba_group)
                            : (algo == BRBD_GEN_CRBA && !(flags & BRBD_GEN_HOST))
                            ? wrap_device_crba(body, fp32, st, nt, m->pd.nq, m->pd.nv, K.definition("__constant__"), crba_nbuf, crba_group, crba_compact ? &pat : nullptr)
                            : crba_bulk_host ? wrap_host_crba_bulk(body, fp32, K.definition("static const"), m->pd.nv, crba_nbuf, crba_group)
                            : (flags & BRBD_GEN_HOST) ? wrap_host(body, algo_name(algo), fp32, st, K.definition("static const"), m->pd.nv, crba_compact ? &pat : nullptr)
                                                  : wrap_device(body, algo_name(algo), fp32, T.nrec, st, nt, minb, tmem_cols, m->pd.nq, m->pd.nv, copies,
                                                                direct_io, K.definition("__constant__"));
  char * buf = (char *)std::malloc(src.size() + 1);
  if (!buf) return fail(BRBD_ENOMEM, "out of memory");
  std::memcpy(buf, src.c_str(), src.size() + 1);
  *source = buf;
  if (info)
  {
    info->record_slots = T.nrec; info->park_slots = st.slots; info->smem_slots = st.smem_slots; info->tmem_slots = st.tmem_slots; info->nodes = st.nodes; info->live_nodes = st.live;
    info->adds = st.add; info->muls = st.mul; info->recips = st.recip; info->sqrts = st.sqrt_; info->sincos = st.sincos;
    info->loads = st.inputs + st.rec_ld + st.park_ld; info->stores = st.rec_st + st.park_st + st.outputs;
    info->threads_per_block = nt;
    info->copies = copies;
    if (algo >= BRBD_GEN_RNEA_DERIVATIVES) info->dynamic_smem_bytes = derivs_staged ? (int32_t)(derivs_tile_bytes * (nt / 32)) : 0;
    else if (crba_bulk)
      info->dynamic_smem_bytes = (int32_t)((size_t)crba_nbuf * (nt / 32) * 32 * crba_bulk_pitch(m->pd.nv, crba_group, fp32) * (fp32 ? 4 : 8));
    else if (algo == BRBD_GEN_CRBA && crba_compact) info->dynamic_smem_bytes = (int32_t)((size_t)(nt / 32) * 32 * (pat.maxrow | 1) * (fp32 ? 4 : 8));
    else if (algo == BRBD_GEN_CRBA) info->dynamic_smem_bytes = (int32_t)((size_t)crba_nbuf * (nt / 32) * 32 * (m->pd.nv * crba_group + 2) * (fp32 ? 4 : 8));
    else info->dynamic_smem_bytes = (int32_t)(((direct_io ? 0 : (size_t)(nt / 32) * 32 * ((m->pd.nq | 1) + 2 * (m->pd.nv | 1))) + (size_t)st.smem_slots * nt) * (fp32 ? 4 : 8));
  }
  return BRBD_OK;
}
void brbd_codegen_free(char * source) { std::free(source); }

brbd_status brbd_model_crba_pattern(const brbd_model * m, int32_t * rows, int32_t * cols, int64_t capacity, int64_t * nnz)
{
  if (!m || !nnz) return fail(BRBD_EINVAL, "null argument");
  const cg::CrbaPattern pat = cg::crba_pattern(m->pd, 1);
  *nnz = pat.nnz;
  if (!rows && !cols) return BRBD_OK; // size query
  if (capacity < pat.nnz) return fail(BRBD_EINVAL, "crba pattern: capacity smaller than the number of structural non-zeros");
  for (int k = 0; k < pat.nnz; ++k)
  {
    if (rows) rows[k] = pat.rows[k];
    if (cols) cols[k] = pat.cols[k];
  }
  return BRBD_OK;
}

} // extern "C"
