// codegen/emit.hpp — expression graph (sym.hpp) -> straight-line source for ONE model and ONE algorithm.
//
// The body is emitted against a small macro vocabulary (BRBD_IN0/1/2(k) inputs, BRBD_REC_ST/LD record store, BRBD_OUT
// outputs, BRBD_PARK_ST/LD explicit slots) so that the same text compiles as the device kernel (NVRTC / nvcc, wrapper from
// kernel_source()) and as a plain C++ function on the host — the latter only for the CPU tests of the generator
// (tests/test_codegen.py), never as a product path.
#pragma once

#include <cstdio>
#include <map>
#include <set>
#include <sstream>
#include <string>
#include <vector>

#include "sym.hpp"

namespace brbd
{
namespace cg
{

struct EmitStats
{
  int nodes = 0, live = 0, add = 0, mul = 0, recip = 0, sqrt_ = 0, sincos = 0, max_ = 0, inputs = 0, rec_st = 0, rec_ld = 0,
      park_st = 0, park_ld = 0, outputs = 0, slots = 0;
  int tmem_slots = 0, smem_slots = 0;        // explicit parking: values per configuration in tensor memory / shared memory
  std::set<std::pair<int, int>> park_shapes; // (number of values, space: 0 tensor memory, 1 shared memory) of the park groups
};

inline std::string lit_text(double v)
{
  char buf[64];
  std::snprintf(buf, sizeof buf, "%.17g", v);
  std::string s(buf);
  if (s.find_first_of(".eEn") == std::string::npos) s += ".0";
  return s;
}
// Constants.  A double whose low 32 bits are zero (0.5, 2, -1, ...) is an immediate operand of DFMA / DMUL / DADD; any other
// one costs two UMOV to materialise when written as a literal (970 of the 16 k instructions of the humanoid ABA).  Those go
// to a __constant__ table instead and are read as c[bank][offset] operands: no instruction at all.
struct ConstTable
{
  std::vector<double> vals;
  std::map<uint64_t, int> index;
  std::string ref(double v)
  {
    uint64_t bits;
    std::memcpy(&bits, &v, 8);
    if ((bits & 0xffffffffull) == 0) return "BRBD_C(" + lit_text(v) + ")";
    auto it = index.find(bits);
    if (it == index.end())
    {
      it = index.emplace(bits, (int)vals.size()).first;
      vals.push_back(v);
    }
    return "BRBD_K[" + std::to_string(it->second) + "]";
  }
  std::string definition(const char * qualifier) const
  {
    std::ostringstream os;
    os << qualifier << " real BRBD_K[" << (vals.empty() ? 1 : vals.size()) << "] = {";
    for (size_t k = 0; k < vals.size(); ++k) os << (k ? ", " : "") << "BRBD_C(" << lit_text(vals[k]) << ")";
    if (vals.empty()) os << "BRBD_C(0.0)";
    os << "};\n";
    return os.str();
  }
};

// Emits the body.  `slots_out`: number of explicit park slots the body needs (0 when the tracer did not park explicitly).
// `tmem_capacity`: values per configuration that fit in the warp's tensor-memory slice; park groups beyond it go to shared memory
// `sync_every` > 0: a BRBD_SYNC() (CTA barrier in the device wrapper) after every that many emitted statements.  The body is
// hundreds of KB of straight-line code, far beyond the instruction caches (L0 ~6 KB per scheduler, L1.5 32 KB per SM): warps
// that drift apart each stream it from L2 on their own and starve (ncu: 66 % of the stall samples `no_instruction`); kept within
// one cache window of each other they share ONE sequential, prefetchable stream.
inline std::string emit_body(const Graph & g, EmitStats & st, ConstTable & K, int tmem_capacity = 0, int sync_every = 0)
{
  const int N = (int)g.nodes.size();
  // ---- liveness: outputs and record stores are roots; a live FETCH keeps its parked value alive -------------------
  std::vector<char> live(N, 0);
  std::map<int, const Effect *> park_of; // handle -> PARK effect
  for (const Effect & e : g.effects)
  {
    if (e.kind == Effect::OUTPUT || e.kind == Effect::RECORD_ST)
      for (int v : e.vals) live[v] = 1;
    if (e.kind == Effect::PARK) park_of[e.handle] = &e;
  }
  // record loads keep nothing alive beyond the stores (already roots); record stores whose slot is never loaded are dropped below
  std::vector<char> rec_loaded;
  for (int id = N - 1; id >= 0; --id)
  {
    if (!live[id]) continue;
    const Node & n = g.nodes[id];
    switch (n.op)
    {
    case OP_CONST: case OP_INPUT: break;
    case OP_FETCH: live[park_of.at(n.a)->vals[n.b]] = 1; break;
    case OP_NEG: case OP_SQRT: case OP_SIN: case OP_COS: case OP_RECIP: live[n.a] = 1; break;
    default: live[n.a] = 1; live[n.b] = 1; break;
    }
  }
  // which record slots are read at all
  int nrec = 0;
  for (const Effect & e : g.effects)
    if (e.kind == Effect::RECORD_ST) nrec = std::max(nrec, e.index + (int)e.vals.size());
  rec_loaded.assign(nrec, 0);
  for (int id = 0; id < N; ++id)
    if (live[id] && g.nodes[id].op == OP_INPUT && g.nodes[id].a == 100) rec_loaded[g.nodes[id].b] = 1;
  // which parked elements are fetched
  std::map<std::pair<int, int>, char> fetched;
  for (int id = 0; id < N; ++id)
    if (live[id] && g.nodes[id].op == OP_FETCH) fetched[{g.nodes[id].a, g.nodes[id].b}] = 1;

  // ---- explicit slots: each park group (handle) gets a contiguous range for the elements that are fetched later; first fit
  // over the groups' lifetimes (PARK .. RELEASE in effect order), tensor memory first, shared memory when it is full ------------
  struct Group { int space = 0, first = 0; std::vector<int> elems; }; // elems: fetched element indices, in order
  std::map<int, Group> group_of;
  {
    std::vector<char> used[2];
    auto alloc = [&](int space, int n, int cap) -> int {
      std::vector<char> & u = used[space];
      for (int s0 = 0; cap < 0 || s0 + n <= cap; ++s0)
      {
        if ((int)u.size() < s0 + n) u.resize(s0 + n, 0);
        bool ok = true;
        for (int k = 0; k < n && ok; ++k) ok = !u[s0 + k];
        if (ok) { for (int k = 0; k < n; ++k) u[s0 + k] = 1; return s0; }
      }
      return -1;
    };
    for (const Effect & e : g.effects)
    {
      if (e.kind == Effect::PARK)
      {
        Group gr;
        for (int k = 0; k < (int)e.vals.size(); ++k)
          if (fetched.count({e.handle, k})) gr.elems.push_back(k);
        if (gr.elems.empty()) continue;
        const int n = (int)gr.elems.size();
        int s0 = tmem_capacity > 0 ? alloc(0, n, tmem_capacity) : -1;
        gr.space = s0 >= 0 ? 0 : 1;
        if (s0 < 0) s0 = alloc(1, n, -1);
        gr.first = s0;
        group_of[e.handle] = gr;
      }
      else if (e.kind == Effect::RELEASE)
      {
        auto it = group_of.find(e.handle);
        if (it == group_of.end()) continue;
        for (int k = 0; k < (int)it->second.elems.size(); ++k) used[it->second.space][it->second.first + k] = 0;
      }
    }
    // a slot counts up to its high-water mark
    for (int sp = 0; sp < 2; ++sp)
    {
      int hw = 0;
      for (auto & kv : group_of)
        if (kv.second.space == sp) hw = std::max(hw, kv.second.first + (int)kv.second.elems.size());
      (sp == 0 ? st.tmem_slots : st.smem_slots) = hw;
    }
    st.slots = st.tmem_slots + st.smem_slots;
  }

  // ---- emission in creation order, effects interleaved by `at` ------------------------------------------------------
  std::ostringstream os;
  std::vector<char> emitted(N, 0);
  auto name = [&](int id) -> std::string {
    const Node & n = g.nodes[id];
    if (n.op == OP_CONST) return K.ref(n.val);
    return "t" + std::to_string(id);
  };
  // sin / cos of the same argument leave as one sincos
  std::map<int, std::pair<int, int>> sc_of_arg; // arg -> (sin node, cos node)
  for (int id = 0; id < N; ++id)
  {
    if (!live[id]) continue;
    if (g.nodes[id].op == OP_SIN) sc_of_arg[g.nodes[id].a].first = id + 1;
    if (g.nodes[id].op == OP_COS) sc_of_arg[g.nodes[id].a].second = id + 1;
  }
  auto emit_node = [&](int id) {
    if (emitted[id]) return;
    emitted[id] = 1;
    const Node & n = g.nodes[id];
    const std::string lhs = "  const real t" + std::to_string(id) + " = ";
    switch (n.op)
    {
    case OP_CONST: break;
    case OP_INPUT:
      if (n.a == 100) { os << lhs << "BRBD_REC_LD(" << n.b << ");\n"; st.rec_ld++; }
      else { os << lhs << "BRBD_IN" << n.a << "(" << n.b << ");\n"; st.inputs++; }
      break;
    case OP_FETCH: {
      // the whole group comes back with one vector load; elements nobody reads any more land in dummies
      const Group & gr = group_of.at(n.a);
      const int nn = (int)gr.elems.size();
      std::vector<std::string> names;
      for (int k = 0; k < nn; ++k)
      {
        const int fid = g.find(OP_FETCH, n.a, gr.elems[k], n.val); // same fetch generation
        if (fid >= 0 && live[fid])
        {
          names.push_back("t" + std::to_string(fid));
          emitted[fid] = 1;
        }
        else
          names.push_back("u" + std::to_string(id) + "_" + std::to_string(k)); // nobody reads this element of this fetch
      }
      os << "  real";
      for (int k = 0; k < nn; ++k) os << (k ? ", " : " ") << names[k];
      os << ";\n  BRBD_PARK_LD" << (gr.space == 0 ? "T" : "S") << nn << "(" << gr.first;
      for (int k = 0; k < nn; ++k) os << ", " << names[k];
      os << ");\n";
      st.park_ld += nn;
      st.park_shapes.insert({nn, gr.space});
      break;
    }
    case OP_ADD: os << lhs << name(n.a) << " + " << name(n.b) << ";\n"; st.add++; break;
    case OP_SUB: os << lhs << name(n.a) << " - " << name(n.b) << ";\n"; st.add++; break;
    case OP_MUL: os << lhs << name(n.a) << " * " << name(n.b) << ";\n"; st.mul++; break;
    case OP_DIV: os << lhs << name(n.a) << " / " << name(n.b) << ";\n"; st.recip++; break;
    case OP_RECIP: os << lhs << "BRBD_C(1.0) / " << name(n.a) << ";\n"; st.recip++; break;
    case OP_NEG: os << lhs << "-" << name(n.a) << ";\n"; break;
    case OP_SQRT: os << lhs << "BRBD_SQRT(" << name(n.a) << ");\n"; st.sqrt_++; break;
    case OP_MAX: os << lhs << "BRBD_MAX(" << name(n.a) << ", " << name(n.b) << ");\n"; st.max_++; break;
    case OP_SIN: case OP_COS: {
      const auto sc = sc_of_arg[n.a];
      if (sc.first && sc.second)
      {
        os << "  real t" << sc.first - 1 << ", t" << sc.second - 1 << ";\n  BRBD_SINCOS(" << name(n.a) << ", &t" << sc.first - 1 << ", &t"
           << sc.second - 1 << ");\n";
        emitted[sc.first - 1] = emitted[sc.second - 1] = 1;
        st.sincos++;
      }
      else
      {
        os << lhs << (n.op == OP_SIN ? "BRBD_SIN(" : "BRBD_COS(") << name(n.a) << ");\n";
        st.sincos++;
      }
      break;
    }
    }
  };
  size_t ei = 0;
  int since_sync = 0;
  for (int id = 0; id <= N; ++id)
  {
    if (sync_every > 0 && since_sync >= sync_every)
    {
      os << "  BRBD_SYNC();\n";
      since_sync = 0;
    }
    while (ei < g.effects.size() && g.effects[ei].at <= id)
    {
      const Effect & e = g.effects[ei++];
      switch (e.kind)
      {
      case Effect::OUTPUT: os << "  BRBD_OUT" << e.handle << "(" << e.index << ", " << name(e.vals[0]) << ");\n"; st.outputs++; break;
      case Effect::RECORD_ST:
        for (int k = 0; k < (int)e.vals.size(); ++k)
          if (rec_loaded[e.index + k]) { os << "  BRBD_REC_ST(" << e.index + k << ", " << name(e.vals[k]) << ");\n"; st.rec_st++; }
        break;
      case Effect::PARK: {
        auto it = group_of.find(e.handle);
        if (it == group_of.end()) break;
        const Group & gr = it->second;
        const int nn = (int)gr.elems.size();
        os << "  BRBD_PARK_ST" << (gr.space == 0 ? "T" : "S") << nn << "(" << gr.first;
        for (int k = 0; k < nn; ++k) os << ", " << name(e.vals[gr.elems[k]]);
        os << "); // " << e.text << "\n";
        st.park_st += nn;
        st.park_shapes.insert({nn, gr.space});
        break;
      }
      case Effect::RELEASE: break;
      case Effect::FLUSH:
        os << (e.handle == 0 ? "  BRBD_COLBEGIN(" : (e.handle == 1 ? "  BRBD_FLUSH(" : "  BRBD_CLEAR(")) << e.index << ");\n";
        break;
      case Effect::COMMENT: os << "  // " << e.text << "\n"; break;
      }
    }
    if (id < N && live[id] && g.nodes[id].op != OP_CONST) { emit_node(id); ++since_sync; }
  }
  st.nodes = N;
  for (int id = 0; id < N; ++id) st.live += live[id] && g.nodes[id].op != OP_CONST;
  return os.str();
}

} // namespace cg
} // namespace brbd
