// urdf_loader.cpp — URDF -> flattened model on the host C++ side (brbd_model_from_urdf), so that a C++ caller does not
// have to fill brbd_flat_model by hand.  Host code only, no CUDA.
//
// Restates the rules of the reference's URDF path for the joints the engine supports:
//   src/parsers/urdf/model.cpp:30-42   convertFromUrdf(Inertial): Y = (mass, origin.p, R I R^T)
//   src/parsers/urdf/model.cpp:67-294  parseTree: children visited depth-first, in urdfdom's order (the joints are kept in a
//                                      std::map<std::string, ...>, i.e. sorted by joint name, pixi.lock:180 urdfdom 4.0.1)
//   include/pinocchio/parsers/urdf/model.hxx:205-211,601-616  optional root joint
//   :269-273  REVOLUTE -> RX/RY/RZ/RevoluteUnaligned, CONTINUOUS -> RUBX/RUBY/RUBZ/RevoluteUnboundedUnaligned
//   :347-361  fixed joints: the child's inertia is added to the parent joint's body, I_parent += X . I
//   :437-480,564-574  axis classification (Eigen isApprox against the unit axes at 1e-12, else *Unaligned(axis.normalized()))
//   :298-319  mimic joints -> rejected here with the reference's message for unsupported mimics
//   include/pinocchio/multibody/model.hxx:61-170,406-413  addJoint / appendBodyToJoint
// The XML reader below handles what URDF files use: elements, attributes, comments, the XML declaration; no entities beyond
// the five predefined ones, no CDATA, no namespaces.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/pinocchio_b200.h"

namespace brbd
{
brbd_status fail(brbd_status s, const std::string & msg); // capi.cu
}

namespace
{
// ---- minimal XML -----------------------------------------------------------------------------------------------------
struct XmlNode
{
  std::string tag;
  std::map<std::string, std::string> attr;
  std::vector<std::unique_ptr<XmlNode>> children;
  const XmlNode * child(const char * name) const
  {
    for (const auto & c : children)
      if (c->tag == name) return c.get();
    return nullptr;
  }
  std::string get(const char * key, const char * dflt = "") const
  {
    auto it = attr.find(key);
    return it == attr.end() ? std::string(dflt) : it->second;
  }
};

struct XmlParser
{
  const std::string & s;
  size_t p = 0;
  std::string err;
  explicit XmlParser(const std::string & text) : s(text) {}
  void skip_ws() { while (p < s.size() && std::isspace((unsigned char)s[p])) ++p; }
  bool starts(const char * t) const { return s.compare(p, std::strlen(t), t) == 0; }
  bool skip_misc()
  { // whitespace, comments, processing instructions, DOCTYPE, text
    for (;;)
    {
      skip_ws();
      if (starts("<!--"))
      {
        const size_t e = s.find("-->", p);
        if (e == std::string::npos) { err = "unterminated comment"; return false; }
        p = e + 3;
      }
      else if (starts("<?"))
      {
        const size_t e = s.find("?>", p);
        if (e == std::string::npos) { err = "unterminated processing instruction"; return false; }
        p = e + 2;
      }
      else if (starts("<!"))
      {
        const size_t e = s.find('>', p);
        if (e == std::string::npos) { err = "unterminated declaration"; return false; }
        p = e + 1;
      }
      else if (p < s.size() && s[p] != '<')
      { // character data: URDF keeps nothing there
        const size_t e = s.find('<', p);
        p = e == std::string::npos ? s.size() : e;
      }
      else
        return true;
    }
  }
  static std::string unescape(const std::string & v)
  {
    std::string o;
    for (size_t i = 0; i < v.size(); ++i)
    {
      if (v[i] != '&') { o += v[i]; continue; }
      const char * ents[] = {"&amp;", "&lt;", "&gt;", "&quot;", "&apos;"};
      const char repl[] = {'&', '<', '>', '"', '\''};
      bool hit = false;
      for (int k = 0; k < 5; ++k)
        if (v.compare(i, std::strlen(ents[k]), ents[k]) == 0) { o += repl[k]; i += std::strlen(ents[k]) - 1; hit = true; break; }
      if (!hit) o += v[i];
    }
    return o;
  }
  std::unique_ptr<XmlNode> element()
  {
    if (!skip_misc()) return nullptr;
    if (p >= s.size() || s[p] != '<') { err = "expected an element"; return nullptr; }
    ++p;
    std::unique_ptr<XmlNode> n(new XmlNode());
    while (p < s.size() && !std::isspace((unsigned char)s[p]) && s[p] != '>' && s[p] != '/') n->tag += s[p++];
    for (;;)
    {
      skip_ws();
      if (p >= s.size()) { err = "unterminated element <" + n->tag + ">"; return nullptr; }
      if (s[p] == '/')
      {
        if (p + 1 >= s.size() || s[p + 1] != '>') { err = "malformed empty element <" + n->tag + ">"; return nullptr; }
        p += 2;
        return n;
      }
      if (s[p] == '>') { ++p; break; }
      std::string key;
      while (p < s.size() && s[p] != '=' && !std::isspace((unsigned char)s[p])) key += s[p++];
      skip_ws();
      if (p >= s.size() || s[p] != '=') { err = "attribute without a value in <" + n->tag + ">"; return nullptr; }
      ++p;
      skip_ws();
      if (p >= s.size() || (s[p] != '"' && s[p] != '\'')) { err = "unquoted attribute value in <" + n->tag + ">"; return nullptr; }
      const char quote = s[p++];
      const size_t e = s.find(quote, p);
      if (e == std::string::npos) { err = "unterminated attribute value in <" + n->tag + ">"; return nullptr; }
      n->attr[key] = unescape(s.substr(p, e - p));
      p = e + 1;
    }
    for (;;)
    {
      if (!skip_misc()) return nullptr;
      if (p >= s.size()) { err = "missing </" + n->tag + ">"; return nullptr; }
      if (starts("</"))
      {
        const size_t e = s.find('>', p);
        if (e == std::string::npos) { err = "unterminated closing tag"; return nullptr; }
        p = e + 1;
        return n;
      }
      std::unique_ptr<XmlNode> c = element();
      if (!c) return nullptr;
      n->children.push_back(std::move(c));
    }
  }
};

// ---- small fixed-size algebra (host, double) ----------------------------------------------------------------------------
struct M3 { double a[9]; }; // row-major
struct Se3 { M3 R; double p[3]; };
struct Inr { double m; double c[3]; double I[9]; }; // mass, lever, inertia about the CoM (full symmetric 3x3)

M3 eye() { return M3{{1, 0, 0, 0, 1, 0, 0, 0, 1}}; }
M3 mul(const M3 & A, const M3 & B)
{
  M3 C;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) C.a[3 * r + c] = A.a[3 * r] * B.a[c] + A.a[3 * r + 1] * B.a[3 + c] + A.a[3 * r + 2] * B.a[6 + c];
  return C;
}
M3 transpose(const M3 & A)
{
  M3 T;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) T.a[3 * r + c] = A.a[3 * c + r];
  return T;
}
void mulv(const M3 & A, const double * v, double * o)
{
  for (int r = 0; r < 3; ++r) o[r] = A.a[3 * r] * v[0] + A.a[3 * r + 1] * v[1] + A.a[3 * r + 2] * v[2];
}
Se3 compose(const Se3 & a, const Se3 & b) // se3-tpl.hpp:314-317
{
  Se3 r;
  r.R = mul(a.R, b.R);
  double t[3];
  mulv(a.R, b.p, t);
  for (int k = 0; k < 3; ++k) r.p[k] = a.p[k] + t[k];
  return r;
}
Se3 identity() { return Se3{eye(), {0, 0, 0}}; }
// aI = aXb.act(bI), inertia.hpp:872-880
Inr act(const Se3 & X, const Inr & Y)
{
  Inr r;
  r.m = Y.m;
  double t[3];
  mulv(X.R, Y.c, t);
  for (int k = 0; k < 3; ++k) r.c[k] = X.p[k] + t[k];
  const M3 I{{Y.I[0], Y.I[1], Y.I[2], Y.I[3], Y.I[4], Y.I[5], Y.I[6], Y.I[7], Y.I[8]}};
  const M3 RI = mul(mul(X.R, I), transpose(X.R));
  std::memcpy(r.I, RI.a, sizeof(r.I));
  return r;
}
// Ya += Yb, inertia.hpp:659-673
void add(Inr & A, const Inr & B)
{
  const double eps = 2.220446049250313e-16;
  const double mab = A.m + B.m, inv = 1.0 / std::max(mab, eps);
  const double AB[3] = {A.c[0] - B.c[0], A.c[1] - B.c[1], A.c[2] - B.c[2]};
  for (int k = 0; k < 3; ++k) A.c[k] = A.c[k] * (A.m * inv) + (B.m * inv) * B.c[k];
  const double k = A.m * B.m * inv, x = AB[0], y = AB[1], z = AB[2];
  const double S[9] = {k * (y * y + z * z), -k * x * y, -k * x * z, -k * x * y, k * (x * x + z * z), -k * y * z, -k * x * z, -k * y * z, k * (x * x + y * y)};
  for (int e = 0; e < 9; ++e) A.I[e] += B.I[e] + S[e];
  A.m = mab;
}
bool is_zero(const Inr & Y)
{
  if (Y.m != 0) return false;
  for (double v : Y.c) if (v != 0) return false;
  for (double v : Y.I) if (v != 0) return false;
  return true;
}

bool parse_doubles(const std::string & text, int n, double * out)
{
  std::istringstream is(text);
  for (int k = 0; k < n; ++k)
    if (!(is >> out[k])) return false;
  return true;
}
// urdf::Rotation::setFromRPY -> quaternion -> matrix (src/parsers/urdf/utils.cpp:10-15)
M3 rpy_to_matrix(double r, double pch, double yw)
{
  const double phi = r / 2, the = pch / 2, psi = yw / 2;
  double x = std::sin(phi) * std::cos(the) * std::cos(psi) - std::cos(phi) * std::sin(the) * std::sin(psi);
  double y = std::cos(phi) * std::sin(the) * std::cos(psi) + std::sin(phi) * std::cos(the) * std::sin(psi);
  double z = std::cos(phi) * std::cos(the) * std::sin(psi) - std::sin(phi) * std::sin(the) * std::cos(psi);
  double w = std::cos(phi) * std::cos(the) * std::cos(psi) + std::sin(phi) * std::sin(the) * std::sin(psi);
  const double s = std::sqrt(x * x + y * y + z * z + w * w);
  if (s == 0.0) { x = y = z = 0; w = 1; }
  else { x /= s; y /= s; z /= s; w /= s; }
  // Eigen Quaternion::toRotationMatrix
  const double tx = 2 * x, ty = 2 * y, tz = 2 * z, twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y,
               tyz = tz * y, tzz = tz * z;
  return M3{{1 - (tyy + tzz), txy - twz, txz + twy, txy + twz, 1 - (txx + tzz), tyz - twx, txz - twy, tyz + twx, 1 - (txx + tyy)}};
}
bool parse_origin(const XmlNode * o, Se3 & X, std::string & err)
{
  X = identity();
  if (!o) return true;
  double xyz[3] = {0, 0, 0}, rpy[3] = {0, 0, 0};
  if (o->attr.count("xyz") && !parse_doubles(o->get("xyz"), 3, xyz)) { err = "malformed origin xyz"; return false; }
  if (o->attr.count("rpy") && !parse_doubles(o->get("rpy"), 3, rpy)) { err = "malformed origin rpy"; return false; }
  X.R = rpy_to_matrix(rpy[0], rpy[1], rpy[2]);
  std::memcpy(X.p, xyz, sizeof(xyz));
  return true;
}
bool parse_inertial(const XmlNode * link, Inr & Y, std::string & err)
{
  std::memset(&Y, 0, sizeof(Y));
  const XmlNode * in = link->child("inertial");
  if (!in) return true;
  Se3 X;
  if (!parse_origin(in->child("origin"), X, err)) return false;
  const XmlNode * m = in->child("mass");
  const XmlNode * I = in->child("inertia");
  if (!m || !I) { err = "inertial of link " + link->get("name") + " lacks mass or inertia"; return false; }
  Y.m = std::atof(m->get("value", "0").c_str());
  std::memcpy(Y.c, X.p, sizeof(Y.c));
  auto g = [&](const char * k) { return std::atof(I->get(k, "0").c_str()); };
  const M3 Im{{g("ixx"), g("ixy"), g("ixz"), g("ixy"), g("iyy"), g("iyz"), g("ixz"), g("iyz"), g("izz")}};
  const M3 RI = mul(mul(X.R, Im), transpose(X.R));
  std::memcpy(Y.I, RI.a, sizeof(Y.I));
  return true;
}
// extractCartesianAxis, parsers/urdf/model.hxx:564-574 (Eigen isApprox at 1e-12)
int axis_tag(const double * a)
{
  const double n = std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
  for (int k = 0; k < 3; ++k)
  {
    double d[3] = {a[0], a[1], a[2]};
    d[k] -= 1.0;
    if (std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]) <= 1e-12 * std::min(n, 1.0)) return k;
  }
  return -1;
}

struct Builder
{
  std::vector<int32_t> parents{0}, types{BRBD_JOINT_UNIVERSE}, idx_q{0}, idx_v{0};
  std::vector<Se3> placement{identity()};
  std::vector<Inr> inertia{Inr{}};
  std::vector<double> axis{0, 0, 0};
  int nq = 0, nv = 0;
  static int nq_of(int t)
  {
    if (t >= BRBD_JOINT_RUBX) return 2;
    return (t <= BRBD_JOINT_PZ || t == BRBD_JOINT_REVOLUTE_UNALIGNED || t == BRBD_JOINT_PRISMATIC_UNALIGNED) ? 1 : (t == BRBD_JOINT_FREEFLYER ? 7 : 4);
  }
  static int nv_of(int t)
  {
    if (t >= BRBD_JOINT_RUBX) return 1;
    return (t <= BRBD_JOINT_PZ || t == BRBD_JOINT_REVOLUTE_UNALIGNED || t == BRBD_JOINT_PRISMATIC_UNALIGNED) ? 1 : (t == BRBD_JOINT_FREEFLYER ? 6 : 3);
  }
  int add_joint(int parent, int type, const Se3 & X, const double * ax)
  { // ModelTpl::addJoint, model.hxx:61-170
    parents.push_back(parent); types.push_back(type); idx_q.push_back(nq); idx_v.push_back(nv);
    placement.push_back(X);
    inertia.push_back(Inr{});
    double a[3] = {0, 0, 0};
    if (ax)
    {
      const double n = std::sqrt(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]);
      for (int k = 0; k < 3; ++k) a[k] = ax[k] / n;
    }
    axis.insert(axis.end(), a, a + 3);
    nq += nq_of(type);
    nv += nv_of(type);
    return (int)parents.size() - 1;
  }
  void append_body(int joint, const Inr & Y, const Se3 & X)
  { // appendBodyToJoint, model.hxx:406-413: inertias[joint] += X.act(Y)
    if (is_zero(Y)) return;
    add(inertia[joint], act(X, Y));
  }
};

struct BodyFrame { int joint; Se3 placement; }; // where a link sits: parent joint and placement in that joint's frame

} // namespace

extern "C" brbd_status brbd_model_from_urdf(const char * path_or_xml, int root_joint_type, brbd_model ** out)
{
  using brbd::fail;
  if (!path_or_xml || !out) return fail(BRBD_EINVAL, "null argument");
  *out = nullptr;
  std::string text(path_or_xml);
  {
    size_t k = 0;
    while (k < text.size() && std::isspace((unsigned char)text[k])) ++k;
    if (k >= text.size() || text[k] != '<')
    {
      std::ifstream fh(path_or_xml);
      if (!fh) return fail(BRBD_EINVAL, std::string("The file ") + path_or_xml + " does not contain a valid URDF model."); // urdf/model.cpp
      std::stringstream ss;
      ss << fh.rdbuf();
      text = ss.str();
    }
  }
  XmlParser xp(text);
  std::unique_ptr<XmlNode> robot = xp.element();
  if (!robot || robot->tag != "robot") return fail(BRBD_EINVAL, "URDF: " + (xp.err.empty() ? std::string("the root element is not <robot>") : xp.err));
  if (root_joint_type != BRBD_JOINT_UNIVERSE && root_joint_type != BRBD_JOINT_FREEFLYER && root_joint_type != BRBD_JOINT_PLANAR
      && root_joint_type != BRBD_JOINT_SPHERICAL)
    return fail(BRBD_EINVAL, "URDF: the root joint must be BRBD_JOINT_UNIVERSE (none), FREEFLYER, PLANAR or SPHERICAL");

  std::map<std::string, const XmlNode *> links, joints; // std::map: name-sorted, the order urdfdom hands the children out in
  for (const auto & c : robot->children)
  {
    if (c->tag == "link") links[c->get("name")] = c.get();
    else if (c->tag == "joint") joints[c->get("name")] = c.get();
  }
  std::map<std::string, std::vector<std::pair<std::string, std::string>>> children;
  std::map<std::string, bool> has_parent;
  for (const auto & kv : joints)
  {
    const XmlNode * p = kv.second->child("parent");
    const XmlNode * c = kv.second->child("child");
    if (!p || !c) return fail(BRBD_EINVAL, "URDF: joint " + kv.first + " lacks a parent or a child link");
    if (!links.count(p->get("link")) || !links.count(c->get("link")))
      return fail(BRBD_EINVAL, "URDF: joint " + kv.first + " refers to an unknown link");
    children[p->get("link")].push_back({kv.first, c->get("link")});
    has_parent[c->get("link")] = true;
  }
  std::string root;
  int nroots = 0;
  for (const auto & kv : links)
    if (!has_parent.count(kv.first)) { root = kv.first; ++nroots; }
  if (nroots != 1) return fail(BRBD_EINVAL, "URDF must have exactly one root link");

  Builder B;
  std::map<std::string, BodyFrame> body;
  std::string err;
  Inr Yroot;
  if (!parse_inertial(links[root], Yroot, err)) return fail(BRBD_EINVAL, "URDF: " + err);
  if (root_joint_type == BRBD_JOINT_UNIVERSE) body[root] = BodyFrame{0, identity()}; // the root link is welded to the universe
  else
  { // addRootJoint, parsers/urdf/model.hxx:205-211
    const int j = B.add_joint(0, root_joint_type, identity(), nullptr);
    B.append_body(j, Yroot, identity());
    body[root] = BodyFrame{j, identity()};
  }
  // depth-first, children in joint-name order (src/parsers/urdf/model.cpp:290-293)
  std::vector<std::pair<std::string, size_t>> stack{{root, 0}};
  while (!stack.empty())
  {
    const std::string link = stack.back().first;
    const size_t k = stack.back().second;
    const auto & ch = children[link];
    if (k >= ch.size()) { stack.pop_back(); continue; }
    stack.back().second = k + 1;
    const std::string & jname = ch[k].first;
    const std::string & child = ch[k].second;
    const XmlNode * j = joints[jname];
    const std::string jtype = j->get("type");
    const BodyFrame frame = body[link];
    Se3 X;
    Inr Y;
    if (!parse_origin(j->child("origin"), X, err) || !parse_inertial(links[child], Y, err)) return fail(BRBD_EINVAL, "URDF: " + err);
    double ax[3] = {1, 0, 0};
    if (const XmlNode * a = j->child("axis"))
      if (a->attr.count("xyz") && !parse_doubles(a->get("xyz"), 3, ax)) return fail(BRBD_EINVAL, "URDF: malformed axis of joint " + jname);
    const Se3 M = compose(frame.placement, X);
    // (talos_reduced.urdf carries <mimic> on FIXED joints: urdfdom keeps those fixed, and so does the reference)
    if (j->child("mimic") && jtype != "fixed")
      return fail(BRBD_EUNSUPPORTED_JOINT, "Cannot mimic this type. Only revolute, prismatic and helicoidal can be mimicked (joint " + jname
                                             + "; mimic joints are not supported by the batched engine)");
    if (jtype == "fixed")
    { // addFixedJointAndBody, parsers/urdf/model.hxx:347-361: the body goes to the parent joint
      B.append_body(frame.joint, Y, M);
      body[child] = BodyFrame{frame.joint, M};
    }
    else
    {
      int tag;
      const int k3 = axis_tag(ax);
      if (jtype == "revolute") tag = k3 < 0 ? BRBD_JOINT_REVOLUTE_UNALIGNED : BRBD_JOINT_RX + k3;
      else if (jtype == "continuous") tag = k3 < 0 ? BRBD_JOINT_REVOLUTE_UNBOUNDED_UNALIGNED : BRBD_JOINT_RUBX + k3;
      else if (jtype == "prismatic") tag = k3 < 0 ? BRBD_JOINT_PRISMATIC_UNALIGNED : BRBD_JOINT_PX + k3;
      else if (jtype == "floating") tag = BRBD_JOINT_FREEFLYER;
      else if (jtype == "planar") tag = BRBD_JOINT_PLANAR;
      else return fail(BRBD_EUNSUPPORTED_JOINT, "The type of joint " + jname + " (" + jtype + ") is not supported."); // model.cpp
      const bool needs_axis = tag == BRBD_JOINT_REVOLUTE_UNALIGNED || tag == BRBD_JOINT_PRISMATIC_UNALIGNED || tag == BRBD_JOINT_REVOLUTE_UNBOUNDED_UNALIGNED;
      if (needs_axis && !(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2] > 0)) return fail(BRBD_EINVAL, "URDF: zero axis on joint " + jname);
      const int jid = B.add_joint(frame.joint, tag, M, needs_axis ? ax : nullptr);
      B.append_body(jid, Y, identity());
      body[child] = BodyFrame{jid, identity()};
    }
    stack.push_back({child, 0});
  }
  // flatten and hand over to brbd_model_create (validation, topology tables, re-framing of the unaligned joints)
  const int n = (int)B.parents.size();
  std::vector<double> plc(12 * (size_t)n), inr(10 * (size_t)n), arm((size_t)std::max(1, B.nv), 0.0);
  for (int i = 0; i < n; ++i)
  {
    std::memcpy(&plc[12 * i], B.placement[i].R.a, 9 * sizeof(double));
    std::memcpy(&plc[12 * i + 9], B.placement[i].p, 3 * sizeof(double));
    const Inr & Y = B.inertia[i];
    const double v[10] = {Y.m, Y.c[0], Y.c[1], Y.c[2], Y.I[0], Y.I[3], Y.I[4], Y.I[6], Y.I[7], Y.I[8]}; // (xx, xy, yy, xz, yz, zz)
    std::memcpy(&inr[10 * i], v, sizeof(v));
  }
  brbd_flat_model f;
  f.njoints = n; f.nq = B.nq; f.nv = B.nv;
  f.parents = B.parents.data(); f.joint_type = B.types.data(); f.idx_q = B.idx_q.data(); f.idx_v = B.idx_v.data();
  f.placement = plc.data(); f.inertia = inr.data(); f.armature = arm.data(); f.axis = B.axis.data();
  f.gravity[0] = 0; f.gravity[1] = 0; f.gravity[2] = -9.81; // model.hxx:40
  return brbd_model_create(&f, out);
}
