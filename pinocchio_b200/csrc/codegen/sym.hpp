// codegen/sym.hpp — tracing scalar of the per-model code generator.
//
// The generic kernels walk the kinematic tree at run time: a type-tag switch per joint, the joint record and the constants
// fetched from the constant bank, loop-carried register moves.  The reference removes the same overhead on the CPU with code
// generation (include/pinocchio/codegen/code-generator-algo.hpp:22-570, CppAD + CppADCodeGen: the algorithm is run once on a
// recording scalar, the tape is emitted as straight-line C).  This is the GPU analogue: the device algorithm, templated on
// its scalar, is run once on the host with `Sym`; every arithmetic operation appends a node to an expression graph with
//   * constant folding (the model's placements / inertias are constants: an identity rotation costs nothing),
//   * algebraic simplification (x*0, x*1, x+0, x-0, 0-x, -(-x), x*(-1)),
//   * common-subexpression elimination (hash-consing; commutative operands ordered),
// and the graph is emitted as straight-line CUDA for ONE model (emit.hpp), compiled by NVRTC at pool specialisation.
// Host only: nothing in this directory runs an algorithm on numbers — it builds programs.
#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#ifndef BRBD_HD
#ifdef __CUDACC__
#define BRBD_HD __host__ __device__
#else
#define BRBD_HD
#endif
#endif

namespace brbd
{
namespace cg
{

enum Op : uint8_t
{
  OP_CONST, OP_INPUT, OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_NEG, OP_SQRT, OP_SIN, OP_COS, OP_MAX,
  OP_FETCH,  // value read back from a parked slot: a = slot handle, b = element
  OP_RECIP
};

struct Node
{
  Op op;
  int a, b;      // operand node ids (or input array / index, or fetch handle / element)
  double val;    // OP_CONST
};

// side effects, in program order, interleaved with the pure nodes by `at` (= number of nodes when issued)
struct Effect
{
  enum Kind : uint8_t { OUTPUT, PARK, RELEASE, RECORD_ST, COMMENT, FLUSH } kind;
  int at;                 // issued after node id at-1 was created
  int handle;             // PARK / RELEASE / RECORD_ST: slot-group handle; OUTPUT: output array
  int index;              // OUTPUT: row; RECORD_ST: first record slot
  std::vector<int> vals;  // node ids
  std::string text;
};

struct Graph
{
  std::vector<Node> nodes;
  std::vector<Effect> effects;
  std::unordered_map<uint64_t, std::vector<int>> cse;
  int n_handles = 0;
  std::vector<int> handle_size; // elements per park handle

  int add_raw(Op op, int a, int b, double v)
  {
    nodes.push_back(Node{op, a, b, v});
    return (int)nodes.size() - 1;
  }
  static uint64_t key(Op op, int a, int b, double v)
  {
    uint64_t bits;
    std::memcpy(&bits, &v, 8);
    uint64_t h = (uint64_t)op * 0x9E3779B97F4A7C15ull;
    h ^= ((uint64_t)(uint32_t)a + 0x7F4A7C15ull) * 0xBF58476D1CE4E5B9ull;
    h ^= ((uint64_t)(uint32_t)b + 0x1CE4E5B9ull) * 0x94D049BB133111EBull;
    h ^= bits * 0xD6E8FEB86659FD93ull;
    return h;
  }
  int intern(Op op, int a, int b, double v)
  {
    const uint64_t k = key(op, a, b, v);
    auto & bucket = cse[k];
    for (int id : bucket)
    {
      const Node & n = nodes[id];
      if (n.op == op && n.a == a && n.b == b && std::memcmp(&n.val, &v, 8) == 0) return id;
    }
    const int id = add_raw(op, a, b, v);
    bucket.push_back(id);
    return id;
  }
  int find(Op op, int a, int b, double v) const // id of an interned node, -1 if there is none
  {
    auto it = cse.find(key(op, a, b, v));
    if (it == cse.end()) return -1;
    for (int id : it->second)
    {
      const Node & n = nodes[id];
      if (n.op == op && n.a == a && n.b == b && std::memcmp(&n.val, &v, 8) == 0) return id;
    }
    return -1;
  }
  int constant(double v) { return intern(OP_CONST, -1, -1, v == 0.0 ? 0.0 : v); } // -0.0 folds into +0.0
  bool is_const(int id) const { return nodes[id].op == OP_CONST; }
  double cval(int id) const { return nodes[id].val; }
  bool is_c(int id, double v) const { return is_const(id) && cval(id) == v; }

  int neg(int a)
  {
    if (is_const(a)) return constant(-cval(a));
    if (nodes[a].op == OP_NEG) return nodes[a].a;
    if (nodes[a].op == OP_SUB) return intern(OP_SUB, nodes[a].b, nodes[a].a, 0.0); // -(x - y) = y - x
    return intern(OP_NEG, a, -1, 0.0);
  }
  int add(int a, int b)
  {
    if (is_const(a) && is_const(b)) return constant(cval(a) + cval(b));
    if (is_c(a, 0.0)) return b;
    if (is_c(b, 0.0)) return a;
    if (nodes[b].op == OP_NEG) return sub(a, nodes[b].a);
    if (nodes[a].op == OP_NEG) return sub(b, nodes[a].a);
    if (a > b) std::swap(a, b);
    return intern(OP_ADD, a, b, 0.0);
  }
  int sub(int a, int b)
  {
    if (is_const(a) && is_const(b)) return constant(cval(a) - cval(b));
    if (is_c(b, 0.0)) return a;
    if (is_c(a, 0.0)) return neg(b);
    if (a == b) return constant(0.0);
    if (nodes[b].op == OP_NEG) return add(a, nodes[b].a);
    return intern(OP_SUB, a, b, 0.0);
  }
  int mul(int a, int b)
  {
    if (is_const(a) && is_const(b)) return constant(cval(a) * cval(b));
    if (is_c(a, 0.0) || is_c(b, 0.0)) return constant(0.0);
    if (is_c(a, 1.0)) return b;
    if (is_c(b, 1.0)) return a;
    if (is_c(a, -1.0)) return neg(b);
    if (is_c(b, -1.0)) return neg(a);
    // signs travel outward so that (-x) * y and x * y share one product
    if (nodes[a].op == OP_NEG && nodes[b].op == OP_NEG) return mul(nodes[a].a, nodes[b].a);
    if (nodes[a].op == OP_NEG) return neg(mul(nodes[a].a, b));
    if (nodes[b].op == OP_NEG) return neg(mul(a, nodes[b].a));
    if (a > b) std::swap(a, b);
    return intern(OP_MUL, a, b, 0.0);
  }
  int div(int a, int b)
  {
    if (is_const(a) && is_const(b)) return constant(cval(a) / cval(b));
    if (is_c(a, 0.0)) return constant(0.0);
    if (is_c(b, 1.0)) return a;
    if (is_const(b)) return mul(a, constant(1.0 / cval(b)));
    if (is_c(a, 1.0)) return intern(OP_RECIP, b, -1, 0.0);
    // one reciprocal shared by every quotient with this denominator (the reference divides by Dinv once, aba.hxx:172)
    return mul(a, intern(OP_RECIP, b, -1, 0.0));
  }
  int sqrt_(int a)
  {
    if (is_const(a)) return constant(std::sqrt(cval(a)));
    return intern(OP_SQRT, a, -1, 0.0);
  }
  int sin_(int a)
  {
    if (is_const(a)) return constant(std::sin(cval(a)));
    return intern(OP_SIN, a, -1, 0.0);
  }
  int cos_(int a)
  {
    if (is_const(a)) return constant(std::cos(cval(a)));
    return intern(OP_COS, a, -1, 0.0);
  }
  int max_(int a, int b)
  {
    if (is_const(a) && is_const(b)) return constant(cval(a) > cval(b) ? cval(a) : cval(b));
    if (a == b) return a;
    return intern(OP_MAX, a, b, 0.0);
  }
  int input(int array, int index) { return intern(OP_INPUT, array, index, 0.0); }
};

inline Graph *& current_graph()
{
  static thread_local Graph * g = nullptr;
  return g;
}

// The recording scalar.  Marked __host__ __device__ only so that the engine's device headers (templated on their scalar,
// BRBD_DI = __host__ __device__ in the generator's translation unit) accept it; it is never instantiated in device code.
struct Sym
{
  int id;
  BRBD_HD Sym() : id(0)
  {
#ifndef __CUDA_ARCH__
    id = current_graph()->constant(0.0);
#endif
  }
  BRBD_HD Sym(double v) : id(0)
  {
#ifndef __CUDA_ARCH__
    id = current_graph()->constant(v);
#endif
  }
  BRBD_HD Sym(int v) : id(0)
  {
#ifndef __CUDA_ARCH__
    id = current_graph()->constant((double)v);
#endif
  }
  struct Raw {};
  BRBD_HD Sym(Raw, int node) : id(node) {}
  static Sym node(int n) { return Sym(Raw{}, n); }
};

#ifndef __CUDA_ARCH__
#define BRBD_CG_BIN(OPNAME, FN)                                                                                          \
  BRBD_HD inline Sym OPNAME(const Sym & a, const Sym & b) { return Sym::node(current_graph()->FN(a.id, b.id)); }
#else
#define BRBD_CG_BIN(OPNAME, FN) BRBD_HD inline Sym OPNAME(const Sym & a, const Sym &) { return a; }
#endif
BRBD_CG_BIN(operator+, add)
BRBD_CG_BIN(operator-, sub)
BRBD_CG_BIN(operator*, mul)
BRBD_CG_BIN(operator/, div)
#undef BRBD_CG_BIN
BRBD_HD inline Sym operator-(const Sym & a)
{
#ifndef __CUDA_ARCH__
  return Sym::node(current_graph()->neg(a.id));
#else
  return a;
#endif
}
BRBD_HD inline Sym & operator+=(Sym & a, const Sym & b) { a = a + b; return a; }
BRBD_HD inline Sym & operator-=(Sym & a, const Sym & b) { a = a - b; return a; }
BRBD_HD inline Sym & operator*=(Sym & a, const Sym & b) { a = a * b; return a; }

} // namespace cg

// hooks of spatial.cuh / engine.cuh for the recording scalar
BRBD_HD inline void sincos_t(cg::Sym x, cg::Sym * s, cg::Sym * c)
{
#ifndef __CUDA_ARCH__
  *s = cg::Sym::node(cg::current_graph()->sin_(x.id));
  *c = cg::Sym::node(cg::current_graph()->cos_(x.id));
#endif
}
BRBD_HD inline cg::Sym sqrt_t(cg::Sym x)
{
#ifndef __CUDA_ARCH__
  return cg::Sym::node(cg::current_graph()->sqrt_(x.id));
#else
  return x;
#endif
}
BRBD_HD inline cg::Sym max_t(cg::Sym a, cg::Sym b)
{
#ifndef __CUDA_ARCH__
  return cg::Sym::node(cg::current_graph()->max_(a.id, b.id));
#else
  return a;
#endif
}

} // namespace brbd
