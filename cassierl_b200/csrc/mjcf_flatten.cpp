// Host-side MJCF reader + planar flattener (see mjcf_flatten.h).
#include "mjcf_flatten.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <vector>

#include "planar_engine.cuh"

namespace cassie {
namespace {

// ------------------------------------------------------------------ tolerant XML subset
// The reference model files are not well-formed (comments closed by '--->',
// cassie2d_stiff.xml:69,73), so this is a small hand parser: elements, attributes in single
// or double quotes, comments, self-closing tags.  No entities, no CDATA, no text nodes.
struct XmlNode {
  std::string tag;
  std::map<std::string, std::string> attr;
  std::vector<std::unique_ptr<XmlNode>> kids;
  const XmlNode* child(const char* t) const {
    for (auto& k : kids) if (k->tag == t) return k.get();
    return nullptr;
  }
  bool has(const char* a) const { return attr.count(a) > 0; }
  std::string get(const char* a, const std::string& d = "") const {
    auto it = attr.find(a);
    return it == attr.end() ? d : it->second;
  }
};

struct XmlParser {
  const std::string& s;
  size_t i = 0;
  std::string err;
  explicit XmlParser(const std::string& str) : s(str) {}
  void skip_ws() { while (i < s.size() && isspace((unsigned char)s[i])) i++; }
  bool skip_misc() {  // whitespace, comments, <?...?>, <!DOCTYPE>
    for (;;) {
      skip_ws();
      if (s.compare(i, 4, "<!--") == 0) {
        size_t e = s.find("-->", i + 4);
        if (e == std::string::npos) { err = "unterminated comment"; return false; }
        i = e + 3;
      } else if (s.compare(i, 2, "<?") == 0) {
        size_t e = s.find("?>", i);
        if (e == std::string::npos) { err = "unterminated <?"; return false; }
        i = e + 2;
      } else if (s.compare(i, 2, "<!") == 0) {
        size_t e = s.find('>', i);
        if (e == std::string::npos) { err = "unterminated <!"; return false; }
        i = e + 1;
      } else
        return true;
    }
  }
  std::unique_ptr<XmlNode> element() {
    if (!skip_misc()) return nullptr;
    if (i >= s.size() || s[i] != '<') { err = "expected '<' at offset " + std::to_string(i); return nullptr; }
    i++;
    auto n = std::make_unique<XmlNode>();
    while (i < s.size() && (isalnum((unsigned char)s[i]) || s[i] == '_' || s[i] == '-' || s[i] == ':')) n->tag += s[i++];
    if (n->tag.empty()) { err = "empty tag at offset " + std::to_string(i); return nullptr; }
    for (;;) {
      skip_ws();
      if (i >= s.size()) { err = "eof in tag <" + n->tag; return nullptr; }
      if (s[i] == '/') {
        if (i + 1 < s.size() && s[i + 1] == '>') { i += 2; return n; }
        err = "stray '/' in <" + n->tag; return nullptr;
      }
      if (s[i] == '>') { i++; break; }
      std::string key;
      while (i < s.size() && s[i] != '=' && !isspace((unsigned char)s[i]) && s[i] != '>' && s[i] != '/') key += s[i++];
      skip_ws();
      if (i >= s.size() || s[i] != '=') { err = "attribute '" + key + "' without value in <" + n->tag; return nullptr; }
      i++;
      skip_ws();
      if (i >= s.size() || (s[i] != '\'' && s[i] != '"')) { err = "unquoted attribute '" + key + "'"; return nullptr; }
      const char qc = s[i++];
      size_t e = s.find(qc, i);
      if (e == std::string::npos) { err = "unterminated attribute '" + key + "'"; return nullptr; }
      n->attr[key] = s.substr(i, e - i);
      i = e + 1;
    }
    for (;;) {  // children until the closing tag
      if (!skip_misc()) return nullptr;
      if (i >= s.size()) { err = "eof inside <" + n->tag + ">"; return nullptr; }
      if (s.compare(i, 2, "</") == 0) {
        size_t e = s.find('>', i);
        if (e == std::string::npos) { err = "unterminated closing tag"; return nullptr; }
        i = e + 1;
        return n;
      }
      if (s[i] != '<') {  // text content: ignore
        while (i < s.size() && s[i] != '<') i++;
        continue;
      }
      auto k = element();
      if (!k) return nullptr;
      n->kids.push_back(std::move(k));
    }
  }
};

bool parse_doubles(const std::string& str, int n, double* out) {
  std::istringstream is(str);
  for (int i = 0; i < n; i++)
    if (!(is >> out[i])) return false;
  return true;
}

// ------------------------------------------------------------------ 3-D intermediate model
struct V3 { double x = 0, y = 0, z = 0; };
struct M3 { double m[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}; };
V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
V3 operator*(double s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
V3 normalized(V3 a) { double n = std::sqrt(dot(a, a)); return (1.0 / n) * a; }
V3 mul(const M3& A, V3 v) {
  return {A.m[0] * v.x + A.m[1] * v.y + A.m[2] * v.z, A.m[3] * v.x + A.m[4] * v.y + A.m[5] * v.z,
          A.m[6] * v.x + A.m[7] * v.y + A.m[8] * v.z};
}
V3 tmul(const M3& A, V3 v) {
  return {A.m[0] * v.x + A.m[3] * v.y + A.m[6] * v.z, A.m[1] * v.x + A.m[4] * v.y + A.m[7] * v.z,
          A.m[2] * v.x + A.m[5] * v.y + A.m[8] * v.z};
}
M3 mul(const M3& A, const M3& B) {
  M3 R;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      R.m[3 * i + j] = A.m[3 * i] * B.m[j] + A.m[3 * i + 1] * B.m[3 + j] + A.m[3 * i + 2] * B.m[6 + j];
  return R;
}
M3 transpose(const M3& A) {
  M3 R;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) R.m[3 * i + j] = A.m[3 * j + i];
  return R;
}
M3 from_columns(V3 x, V3 y, V3 z) {
  M3 R;
  R.m[0] = x.x; R.m[3] = x.y; R.m[6] = x.z;
  R.m[1] = y.x; R.m[4] = y.y; R.m[7] = y.z;
  R.m[2] = z.x; R.m[5] = z.y; R.m[8] = z.z;
  return R;
}
M3 axis_angle(V3 a, double ang) {
  const double c = std::cos(ang), s = std::sin(ang), t = 1 - c;
  M3 R;
  R.m[0] = t * a.x * a.x + c;       R.m[1] = t * a.x * a.y - s * a.z; R.m[2] = t * a.x * a.z + s * a.y;
  R.m[3] = t * a.x * a.y + s * a.z; R.m[4] = t * a.y * a.y + c;       R.m[5] = t * a.y * a.z - s * a.x;
  R.m[6] = t * a.x * a.z - s * a.y; R.m[7] = t * a.y * a.z + s * a.x; R.m[8] = t * a.z * a.z + c;
  return R;
}

struct Joint3 {
  std::string name;
  bool hinge = true, limited = false;
  V3 axis{0, 0, 1}, pos;
  double ref = 0, lo = 0, hi = 0, damping = 0, armature = 0;
  double solref[2] = {0.02, 1}, solimp[5] = {0.9, 0.95, 0.001, 0.5, 2};
};
struct Geom3 {
  int type = 0;  // 0 plane 2 sphere 3 capsule
  V3 pos, from, to;
  double radius = 0;
  int contype = 1, conaffinity = 1, condim = 3;
  double friction[3] = {1, 0.005, 0.0001}, solref[2] = {0.02, 1}, solimp[5] = {0.9, 0.95, 0.001, 0.5, 2};
};
struct Site3 { std::string name; V3 pos; };
struct Body3 {
  std::string name;
  int parent = 0;
  V3 pos, ipos;
  M3 mat;
  double mass = 0, I[6] = {0, 0, 0, 0, 0, 0};
  std::vector<Joint3> joints;
  std::vector<Geom3> geoms;
  std::vector<Site3> sites;
};
struct Connect3 { int b1 = 0, b2 = 0; V3 anchor; double solref[2] = {0.02, 1}, solimp[5] = {0.9, 0.95, 0.001, 0.5, 2}; };
struct Motor3 { std::string joint; double gear = 1, lo = 0, hi = 0; bool limited = false; };
struct Model3 {
  std::vector<Body3> bodies;  // [0] = world
  std::vector<Connect3> connects;
  std::vector<Motor3> motors;
  double timestep = 0.002, tolerance = 1e-8, impratio = 1, gravity[3] = {0, 0, -9.81};
  int iterations = 100;
  std::string cone = "pyramidal", solver = "Newton";
};

struct Defaults { const XmlNode* joint = nullptr; const XmlNode* geom = nullptr; const XmlNode* motor = nullptr; };

std::string dattr(const XmlNode& e, const XmlNode* d, const char* name, const std::string& fallback) {
  if (e.has(name)) return e.get(name);
  if (d && d->has(name)) return d->get(name);
  return fallback;
}

void read_solimp(const std::string& s, double out[5]) {
  double v[5] = {0.9, 0.95, 0.001, 0.5, 2};  // 3-parameter (MuJoCo 1.50) input keeps midpoint .5, power 2
  std::istringstream is(s);
  for (int i = 0; i < 5; i++) { double x; if (is >> x) v[i] = x; else break; }
  for (int i = 0; i < 5; i++) out[i] = v[i];
}

bool read_geom(const XmlNode& g, const Defaults& df, Geom3* out, bool* skip, std::string* err) {
  const std::string type = dattr(g, df.geom, "type", "sphere");
  *skip = false;
  if (type == "mesh") { *skip = true; return true; }
  out->contype = atoi(dattr(g, df.geom, "contype", "1").c_str());
  out->conaffinity = atoi(dattr(g, df.geom, "conaffinity", "1").c_str());
  out->condim = atoi(dattr(g, df.geom, "condim", "3").c_str());
  parse_doubles(dattr(g, df.geom, "friction", "1 0.005 0.0001"), 3, out->friction);
  parse_doubles(dattr(g, df.geom, "solref", "0.02 1"), 2, out->solref);
  read_solimp(dattr(g, df.geom, "solimp", "0.9 0.95 0.001"), out->solimp);
  double sz[3] = {0, 0, 0};
  parse_doubles(dattr(g, df.geom, "size", "0"), 1, sz);
  double p[3] = {0, 0, 0};
  parse_doubles(dattr(g, df.geom, "pos", "0 0 0"), 3, p);
  out->pos = {p[0], p[1], p[2]};
  if (type == "plane") out->type = 0;
  else if (type == "sphere") { out->type = 2; out->radius = sz[0]; }
  else if (type == "capsule") {
    out->type = 3; out->radius = sz[0];
    double ft[6];
    if (!g.has("fromto") || !parse_doubles(g.get("fromto"), 6, ft)) { *err = "capsule without fromto"; return false; }
    out->from = {ft[0], ft[1], ft[2]};
    out->to = {ft[3], ft[4], ft[5]};
  } else { *err = "unsupported geom type '" + type + "'"; return false; }
  return true;
}

bool read_body(const XmlNode& e, int parent, const Defaults& df, double ang, Model3* M, std::string* err) {
  Body3 b;
  b.name = e.get("name");
  b.parent = parent;
  double p[3] = {0, 0, 0};
  if (e.has("pos") && !parse_doubles(e.get("pos"), 3, p)) { *err = "bad body pos"; return false; }
  b.pos = {p[0], p[1], p[2]};
  if (e.has("xyaxes")) {
    double a[6];
    if (!parse_doubles(e.get("xyaxes"), 6, a)) { *err = "bad xyaxes"; return false; }
    V3 x = normalized({a[0], a[1], a[2]});
    V3 y{a[3], a[4], a[5]};
    y = normalized(y - dot(x, y) * x);  // MuJoCo: Gram-Schmidt [EXT]
    b.mat = from_columns(x, y, cross(x, y));
  }
  if (const XmlNode* in = e.child("inertial")) {
    double ip[3], fi[6];
    if (!parse_doubles(in->get("pos"), 3, ip) || !in->has("mass") || !parse_doubles(in->get("fullinertia"), 6, fi)) {
      *err = "inertial needs pos, mass, fullinertia (body " + b.name + ")";
      return false;
    }
    b.ipos = {ip[0], ip[1], ip[2]};
    b.mass = atof(in->get("mass").c_str());
    for (int i = 0; i < 6; i++) b.I[i] = fi[i];
  }
  for (auto& k : e.kids) {
    if (k->tag == "joint") {
      Joint3 j;
      j.name = k->get("name");
      const std::string type = dattr(*k, df.joint, "type", "hinge");
      if (type != "hinge" && type != "slide") { *err = "unsupported joint type '" + type + "'"; return false; }
      j.hinge = type == "hinge";
      double a[3] = {0, 0, 1}, jp[3] = {0, 0, 0}, rg[2] = {0, 0};
      parse_doubles(dattr(*k, df.joint, "axis", "0 0 1"), 3, a);
      parse_doubles(dattr(*k, df.joint, "pos", "0 0 0"), 3, jp);
      parse_doubles(dattr(*k, df.joint, "range", "0 0"), 2, rg);
      j.axis = normalized({a[0], a[1], a[2]});
      j.pos = {jp[0], jp[1], jp[2]};
      const double sc = j.hinge ? ang : 1.0;
      j.ref = atof(dattr(*k, df.joint, "ref", "0").c_str()) * sc;
      j.lo = rg[0] * sc; j.hi = rg[1] * sc;
      j.limited = dattr(*k, df.joint, "limited", "false") == "true";
      j.damping = atof(dattr(*k, df.joint, "damping", "0").c_str());
      j.armature = atof(dattr(*k, df.joint, "armature", "0").c_str());
      parse_doubles(dattr(*k, df.joint, "solreflimit", "0.02 1"), 2, j.solref);
      read_solimp(dattr(*k, df.joint, "solimplimit", "0.9 0.95 0.001"), j.solimp);
      b.joints.push_back(j);
    } else if (k->tag == "geom") {
      Geom3 g; bool skip;
      if (!read_geom(*k, df, &g, &skip, err)) return false;
      if (!skip) b.geoms.push_back(g);
    } else if (k->tag == "site") {
      Site3 s;
      s.name = k->get("name");
      double sp[3] = {0, 0, 0};
      parse_doubles(k->get("pos", "0 0 0"), 3, sp);
      s.pos = {sp[0], sp[1], sp[2]};
      b.sites.push_back(s);
    }
  }
  const int id = (int)M->bodies.size();
  M->bodies.push_back(b);
  for (auto& k : e.kids)
    if (k->tag == "body" && !read_body(*k, id, df, ang, M, err)) return false;
  return true;
}

bool read_model3(const std::string& xml, Model3* M, std::string* err) {
  XmlParser P(xml);
  auto root = P.element();
  if (!root) { *err = "XML: " + P.err; return false; }
  if (root->tag != "mujoco") { *err = "root element is not <mujoco>"; return false; }
  double ang = M_PI / 180.0;
  if (const XmlNode* c = root->child("compiler"))
    if (c->get("angle", "degree") == "radian") ang = 1.0;
  if (const XmlNode* o = root->child("option")) {
    if (o->has("timestep")) M->timestep = atof(o->get("timestep").c_str());
    if (o->has("iterations")) M->iterations = atoi(o->get("iterations").c_str());
    if (o->has("tolerance")) M->tolerance = atof(o->get("tolerance").c_str());
    if (o->has("impratio")) M->impratio = atof(o->get("impratio").c_str());
    if (o->has("gravity")) parse_doubles(o->get("gravity"), 3, M->gravity);
    M->cone = o->get("cone", M->cone);
    M->solver = o->get("solver", M->solver);
    if (o->has("integrator") && o->get("integrator") != "Euler") { *err = "only the Euler integrator is supported"; return false; }
  }
  Defaults df;
  if (const XmlNode* d = root->child("default")) { df.joint = d->child("joint"); df.geom = d->child("geom"); df.motor = d->child("motor"); }
  const XmlNode* wb = root->child("worldbody");
  if (!wb) { *err = "no <worldbody>"; return false; }
  Body3 world;
  world.name = "world";
  for (auto& k : wb->kids)
    if (k->tag == "geom") {
      Geom3 g; bool skip;
      if (!read_geom(*k, df, &g, &skip, err)) return false;
      if (!skip) world.geoms.push_back(g);
    }
  M->bodies.push_back(world);
  for (auto& k : wb->kids)
    if (k->tag == "body" && !read_body(*k, 0, df, ang, M, err)) return false;
  auto body_id = [&](const std::string& n) { for (size_t i = 0; i < M->bodies.size(); i++) if (M->bodies[i].name == n) return (int)i; return -1; };
  if (const XmlNode* eq = root->child("equality"))
    for (auto& k : eq->kids) {
      if (k->tag != "connect") { *err = "unsupported equality '" + k->tag + "'"; return false; }
      Connect3 c;
      c.b1 = body_id(k->get("body1")); c.b2 = body_id(k->get("body2"));
      double a[3];
      if (c.b1 < 0 || c.b2 < 0 || !parse_doubles(k->get("anchor"), 3, a)) { *err = "bad <connect>"; return false; }
      c.anchor = {a[0], a[1], a[2]};
      parse_doubles(k->get("solref", "0.02 1"), 2, c.solref);
      read_solimp(k->get("solimp", "0.9 0.95 0.001"), c.solimp);
      M->connects.push_back(c);
    }
  if (const XmlNode* ac = root->child("actuator"))
    for (auto& k : ac->kids) {
      if (k->tag != "motor") { *err = "unsupported actuator '" + k->tag + "'"; return false; }
      Motor3 mo;
      mo.joint = k->get("joint");
      mo.gear = atof(k->get("gear", "1").c_str());
      mo.limited = dattr(*k, df.motor, "ctrllimited", "false") == "true";
      double rg[2] = {0, 0};
      parse_doubles(k->get("ctrlrange", "0 0"), 2, rg);
      mo.lo = rg[0]; mo.hi = rg[1];
      M->motors.push_back(mo);
    }
  return true;
}

// ------------------------------------------------------------------ 3-D forward kinematics
struct Pose3 { std::vector<V3> xpos; std::vector<M3> xmat; };
struct JointRef { int body, idx; };

void fk3(const Model3& M, const std::vector<JointRef>& jr, const std::vector<double>& q, Pose3* P) {
  const size_t nb = M.bodies.size();
  P->xpos.assign(nb, V3{});
  P->xmat.assign(nb, M3{});
  for (size_t b = 1; b < nb; b++) {
    const Body3& B = M.bodies[b];
    V3 pos = P->xpos[B.parent] + mul(P->xmat[B.parent], B.pos);
    M3 mat = mul(P->xmat[B.parent], B.mat);
    for (size_t d = 0; d < jr.size(); d++) {
      if (jr[d].body != (int)b) continue;
      const Joint3& J = B.joints[jr[d].idx];
      const V3 ax = mul(mat, J.axis);
      if (J.hinge) {
        const V3 an = pos + mul(mat, J.pos);
        const M3 R = axis_angle(ax, q[d] - J.ref);
        pos = an + mul(R, pos - an);
        mat = mul(R, mat);
      } else
        pos = pos + (q[d] - J.ref) * ax;
    }
    P->xpos[b] = pos;
    P->xmat[b] = mat;
  }
}

struct V2 { double x, z; };
V2 xz(V3 v) { return {v.x, v.z}; }

// ------------------------------------------------------------------ flatten one 3-D model
// `anchor_q` = configuration at which the connect's second anchor is made to coincide with the
// first (qpos0 of the ORIGINAL file for both variants, DynamicModel.cpp:141-167).
bool flatten(const Model3& M, const std::vector<double>& anchor_q, PlanarModel<double>* out, std::string* err) {
  PlanarModel<double>& m = *out;
  std::memset(&m, 0, sizeof(m));
  const int nb = (int)M.bodies.size();
  // dof list in file order
  std::vector<JointRef> jr;
  for (int b = 1; b < nb; b++)
    for (size_t j = 0; j < M.bodies[b].joints.size(); j++) jr.push_back({b, (int)j});
  if ((int)jr.size() != kNV) { *err = "expected 13 joints, found " + std::to_string(jr.size()); return false; }
  auto J = [&](int d) -> const Joint3& { return M.bodies[jr[d].body].joints[jr[d].idx]; };
  // link of each body: bodies without joints are welded to the parent's link
  std::vector<int> link_root(nb, 0);  // body that owns the link
  for (int b = 1; b < nb; b++) link_root[b] = M.bodies[b].joints.empty() ? link_root[M.bodies[b].parent] : b;
  const int pelvis = jr[0].body;
  if (jr[1].body != pelvis || jr[2].body != pelvis || M.bodies[pelvis].joints.size() != 3 || M.bodies[pelvis].parent != 0) {
    *err = "expected a root body with exactly three joints (slide x, slide z, hinge y)"; return false;
  }
  // leg link bodies, dof d = 3 + 5L + a
  int lb[2][kLegLinks];
  for (int L = 0; L < 2; L++)
    for (int a = 0; a < kLegLinks; a++) {
      const int d = 3 + 5 * L + a;
      lb[L][a] = jr[d].body;
      if (M.bodies[lb[L][a]].joints.size() != 1 || !J(d).hinge) { *err = "leg joint " + J(d).name + ": expected one hinge per body"; return false; }
    }
  for (int L = 0; L < 2; L++)
    for (int a = 0; a < kLegLinks; a++) {
      const int want = a == kThigh ? pelvis : lb[L][link_parent(a)];
      if (link_root[M.bodies[lb[L][a]].parent] != want) { *err = "unexpected kinematic tree at joint " + J(3 + 5 * L + a).name; return false; }
    }
  // qpos0 pose
  std::vector<double> q0(kNV);
  for (int d = 0; d < kNV; d++) q0[d] = J(d).ref;
  Pose3 P0;
  fk3(M, jr, q0, &P0);
  auto world_axis = [&](int d) { return mul(P0.xmat[jr[d].body], J(d).axis); };
  const double tol = 1e-9;
  {
    const V3 a0 = world_axis(0), a1 = world_axis(1), a2 = world_axis(2);
    if (J(0).hinge || J(1).hinge || !J(2).hinge || std::fabs(a0.x - 1) > tol || std::fabs(a1.z - 1) > tol || std::fabs(a2.y - 1) > tol) {
      *err = "root joints must be slide +x, slide +z, hinge +y"; return false;
    }
  }
  auto pivot = [&](int d) { return P0.xpos[jr[d].body] + mul(P0.xmat[jr[d].body], J(d).pos); };
  const V3 piv0 = pivot(2);
  m.pel_org[0] = piv0.x; m.pel_org[1] = piv0.z;
  for (int i = 0; i < 3; i++) m.pel_ref[i] = J(i).ref;
  // composite inertia of a link = all bodies whose link_root is the link's body
  auto composite = [&](int root, V3 piv, double* com2, double* mass, double* inertia) {
    double mt = 0; V3 c{};
    for (int b = 1; b < nb; b++) if (link_root[b] == root) {
      const V3 cb = P0.xpos[b] + mul(P0.xmat[b], M.bodies[b].ipos);
      mt += M.bodies[b].mass; c = c + M.bodies[b].mass * cb;
    }
    c = (1.0 / mt) * c;
    double I = 0;
    for (int b = 1; b < nb; b++) if (link_root[b] == root) {
      const Body3& B = M.bodies[b];
      M3 Ib; Ib.m[0] = B.I[0]; Ib.m[4] = B.I[1]; Ib.m[8] = B.I[2];
      Ib.m[1] = Ib.m[3] = B.I[3]; Ib.m[2] = Ib.m[6] = B.I[4]; Ib.m[5] = Ib.m[7] = B.I[5];
      const M3 Iw = mul(mul(P0.xmat[b], Ib), transpose(P0.xmat[b]));
      const V3 cb = P0.xpos[b] + mul(P0.xmat[b], B.ipos);
      const double dx = cb.x - c.x, dz = cb.z - c.z;
      I += Iw.m[4] + B.mass * (dx * dx + dz * dz);
    }
    com2[0] = c.x - piv.x; com2[1] = c.z - piv.z; *mass = mt; *inertia = I;
  };
  composite(pelvis, piv0, m.pel_com, &m.pel_mass, &m.pel_inertia);
  m.total_mass = m.pel_mass;
  for (int L = 0; L < 2; L++)
    for (int a = 0; a < kLegLinks; a++) {
      const int d = 3 + 5 * L + a;
      const V3 ax = world_axis(d);
      if (std::fabs(std::fabs(ax.y) - 1) > tol) { *err = "hinge " + J(d).name + " is not about the world y axis"; return false; }
      m.sgn[L][a] = ax.y > 0 ? 1.0 : -1.0;
      m.ang0[L][a] = -m.sgn[L][a] * J(d).ref;
      const V3 pv = pivot(d);
      const V3 pp = a == kThigh ? piv0 : pivot(3 + 5 * L + link_parent(a));
      m.off[L][a][0] = pv.x - pp.x; m.off[L][a][1] = pv.z - pp.z;
      composite(lb[L][a], pv, m.com[L][a], &m.mass[L][a], &m.inertia[L][a]);
      m.total_mass += m.mass[L][a];
    }
  for (int d = 0; d < kNV; d++) {
    m.damping[d] = J(d).damping; m.armature[d] = J(d).armature;
    m.has_limit[d] = J(d).limited ? 1 : 0;
    m.lim_lo[d] = J(d).lo; m.lim_hi[d] = J(d).hi;
    if (J(d).limited && d < 3) { *err = "limited root joints are not supported"; return false; }
    if (J(d).limited) {
      for (int i = 0; i < 2; i++) m.lim_solref[i] = J(d).solref[i];
      for (int i = 0; i < 5; i++) m.lim_solimp[i] = J(d).solimp[i];
    }
  }
  if (m.lim_solref[0] == 0) { m.lim_solref[0] = 0.02; m.lim_solref[1] = 1; double si[5] = {0.9, 0.95, 0.001, 0.5, 2}; for (int i = 0; i < 5; i++) m.lim_solimp[i] = si[i]; }
  // actuators
  if ((int)M.motors.size() != kNU) { *err = "expected 6 motors"; return false; }
  for (int a = 0; a < kNU; a++) {
    int d = -1;
    for (int i = 0; i < kNV; i++) if (J(i).name == M.motors[a].joint) d = i;
    if (d < 3) { *err = "motor on unknown or root joint '" + M.motors[a].joint + "'"; return false; }
    m.act_dof[a] = d; m.act_gear[a] = M.motors[a].gear;
    m.act_lo[a] = M.motors[a].limited ? M.motors[a].lo : -1e30;
    m.act_hi[a] = M.motors[a].limited ? M.motors[a].hi : 1e30;
  }
  // sites in file order (xml_parser.h:151-163)
  {
    int ns = 0;
    for (int b = 1; b < nb; b++)
      for (auto& s : M.bodies[b].sites) {
        if (ns >= kNumSites) { *err = "more than 6 sites"; return false; }
        const int root = link_root[b];
        int link = -2; V3 pv = piv0;
        if (root == pelvis) link = -1;
        for (int L = 0; L < 2; L++) for (int a = 0; a < kLegLinks; a++) if (lb[L][a] == root) { link = 5 * L + a; pv = pivot(3 + link); }
        if (link == -2) { *err = "site on unknown link"; return false; }
        const V3 w = P0.xpos[b] + mul(P0.xmat[b], s.pos);
        m.site_link[ns] = link; m.site_off[ns][0] = w.x - pv.x; m.site_off[ns][1] = w.z - pv.z;
        ns++;
      }
    if (ns != kNumSites) { *err = "expected 6 sites, found " + std::to_string(ns); return false; }
  }
  // geoms: floor plane in the world body, pelvis sphere, four capsules per leg
  const Geom3* floor = nullptr;
  for (auto& g : M.bodies[0].geoms) if (g.type == 0) floor = &g;
  if (!floor) { *err = "no floor plane"; return false; }
  auto collides = [&](const Geom3& g) { return (floor->contype & g.conaffinity) || (g.contype & floor->conaffinity); };
  bool have_sphere = false;
  int ncap[2] = {0, 0};
  const Geom3* any = nullptr;
  std::vector<int> cap_body(kNumCaps, -1);
  int sph_body = -1;
  for (int b = 1; b < nb; b++)
    for (auto& g : M.bodies[b].geoms) {
      if (!collides(g)) continue;
      any = &g;
      const int root = link_root[b];
      if (g.type == 2) {
        if (root != pelvis || have_sphere) { *err = "only one colliding sphere on the pelvis is supported"; return false; }
        const V3 w = P0.xpos[b] + mul(P0.xmat[b], g.pos);
        m.sph_c[0] = w.x - piv0.x; m.sph_c[1] = w.z - piv0.z; m.sph_r = g.radius; have_sphere = true; sph_body = b;
      } else if (g.type == 3) {
        int L = -1, a = -1;
        for (int l = 0; l < 2; l++) for (int x = 0; x < kLegLinks; x++) if (lb[l][x] == root) { L = l; a = x; }
        if (L < 0 || a != ncap[L] || a > kToe) { *err = "colliding capsules must be, per leg, on thigh, knee(shin), tarsus, toe in this order"; return false; }
        const int c = 4 * L + a;
        const V3 pv = pivot(3 + 5 * L + a);
        const V3 wt = P0.xpos[b] + mul(P0.xmat[b], g.to), wf = P0.xpos[b] + mul(P0.xmat[b], g.from);
        m.cap_link[c] = a;
        m.cap_to[c][0] = wt.x - pv.x; m.cap_to[c][1] = wt.z - pv.z;
        m.cap_from[c][0] = wf.x - pv.x; m.cap_from[c][1] = wf.z - pv.z;
        m.cap_r[c] = g.radius; cap_body[c] = b;
        ncap[L]++;
      } else { *err = "unsupported colliding geom"; return false; }
    }
  if (!have_sphere || ncap[0] != 4 || ncap[1] != 4) { *err = "expected pelvis sphere + 4 capsules per leg"; return false; }
  // contact parameter mixing floor x geom [EXT mj_contactParam]: max condim / friction, equal solref/solimp
  {
    const int condim = floor->condim > any->condim ? floor->condim : any->condim;
    if (condim != 3) { *err = "floor contacts must be condim 3"; return false; }
    if (M.cone != "elliptic" || M.solver != "PGS") { *err = "only solver=PGS cone=elliptic is implemented"; return false; }
    m.con_mu = floor->friction[0] > any->friction[0] ? floor->friction[0] : any->friction[0];
    for (int i = 0; i < 2; i++) m.con_solref[i] = 0.5 * (floor->solref[i] + any->solref[i]);
    for (int i = 0; i < 5; i++) m.con_solimp[i] = 0.5 * (floor->solimp[i] + any->solimp[i]);
  }
  // connects: rod anchor <-> tarsus link, one per leg, in leg order
  if (M.connects.size() != 2) { *err = "expected two <connect> constraints"; return false; }
  Pose3 PA;
  fk3(M, jr, anchor_q, &PA);
  int eq_b1[2], eq_b2[2];
  for (int L = 0; L < 2; L++) {
    const Connect3& c = M.connects[L];
    if (link_root[c.b1] != lb[L][kRod] || link_root[c.b2] != lb[L][kTarsus]) { *err = "connect must join the rod to the tarsus link of the same leg"; return false; }
    eq_b1[L] = c.b1; eq_b2[L] = c.b2;
    // anchor 1 in the rod zero frame
    const V3 w1 = P0.xpos[c.b1] + mul(P0.xmat[c.b1], c.anchor);
    const V3 pr = pivot(3 + 5 * L + kRod), pt = pivot(3 + 5 * L + kTarsus);
    m.eq_a1[L][0] = w1.x - pr.x; m.eq_a1[L][1] = w1.z - pr.z;
    // anchor 2: body-2 local coordinates of anchor 1 at `anchor_q`, then into the zero frame
    const V3 wa = PA.xpos[c.b1] + mul(PA.xmat[c.b1], c.anchor);
    const V3 a2 = tmul(PA.xmat[c.b2], wa - PA.xpos[c.b2]);
    const V3 w2 = P0.xpos[c.b2] + mul(P0.xmat[c.b2], a2);
    m.eq_a2[L][0] = w2.x - pt.x; m.eq_a2[L][1] = w2.z - pt.z;
    for (int i = 0; i < 2; i++) m.eq_solref[i] = c.solref[i];
    for (int i = 0; i < 5; i++) m.eq_solimp[i] = c.solimp[i];
  }
  m.timestep = M.timestep; m.gravity_z = M.gravity[2]; m.tolerance = M.tolerance; m.impratio = M.impratio;
  m.iterations = M.iterations;
  if (M.gravity[0] != 0 || M.gravity[1] != 0) { *err = "gravity must be along z"; return false; }

  // ---- invweight0 / meaninertia at qpos0 (mj_setConst [EXT]); planar: tran = (Axx + Azz)/3
  {
    double qd0[kNV] = {0};
    Kin<double> k;
    forward_kinematics(m, q0.data(), qd0, k);
    double Mm[kNV][kNV], Dinv[kNV];
    std::memset(Mm, 0, sizeof(Mm));
    mass_matrix(m, k, Mm);
    double tr = 0;
    for (int i = 0; i < kNV; i++) tr += Mm[i][i];
    m.meaninertia = tr / kNV;
    factor(Mm, Dinv);
    for (int d = 0; d < kNV; d++) {
      double e[kNV] = {0};
      e[d] = 1;
      solve(Mm, Dinv, e);
      m.lim_diag[d] = e[d];
    }
    auto body_tran = [&](int b) {
      const int root = link_root[b];
      int L = 0, a = -1;
      for (int l = 0; l < 2; l++) for (int x = 0; x < kLegLinks; x++) if (lb[l][x] == root) { L = l; a = x; }
      const V3 cb = P0.xpos[b] + mul(P0.xmat[b], M.bodies[b].ipos);
      double Jx[8], Jz[8], x[kNV];
      const double lpx = a >= 0 ? k.px[L][a] : 0.0, lpz = a >= 0 ? k.pz[L][a] : 0.0;
      point_jac(m, k, L, a, cb.x - piv0.x - lpx, cb.z - piv0.z - lpz, Jx, Jz);
      double s = 0;
      expand_row(Jx, L, x); solve(Mm, Dinv, x); s += dot8_dense(Jx, L, x);
      expand_row(Jz, L, x); solve(Mm, Dinv, x); s += dot8_dense(Jz, L, x);
      s /= 3.0;
      return s > kMinVal ? s : kMinVal;
    };
    m.sph_diag = body_tran(sph_body);
    for (int c = 0; c < kNumCaps; c++) m.cap_diag[c] = body_tran(cap_body[c]);
    for (int L = 0; L < 2; L++) m.eq_diag[L] = body_tran(eq_b1[L]) + body_tran(eq_b2[L]);
  }
  return true;
}

}  // namespace

bool flatten_mjcf_text(const std::string& xml, FlatModels* out, std::string* err) {
  Model3 M;
  if (!read_model3(xml, &M, err)) return false;
  std::vector<double> q0;
  for (size_t b = 1; b < M.bodies.size(); b++)
    for (auto& j : M.bodies[b].joints) q0.push_back(j.ref);
  if (!flatten(M, q0, &out->phys, err)) return false;
  // RBDL-loader variant (DynamicModel.cpp:84-103): a body whose LAST joint has |ref| >= 1e-3
  // (degrees in the file) loses its xyaxes rotation and its joints lose their ref.
  // Not modelled: RBDL gets xyaxes normalised but not orthogonalised (known quirks).
  Model3 R = M;
  for (size_t b = 1; b < R.bodies.size(); b++) {
    Body3& B = R.bodies[b];
    if (B.joints.empty()) continue;
    const Joint3& last = B.joints.back();
    if (last.hinge && std::fabs(last.ref * 180.0 / M_PI) >= 1e-3) {
      B.mat = M3{};
      for (auto& j : B.joints) j.ref = 0.0;
    }
  }
  return flatten(R, q0, &out->ctrl, err);
}

bool flatten_mjcf_file(const std::string& path, FlatModels* out, std::string* err) {
  std::ifstream f(path);
  if (!f) { *err = "cannot open model file '" + path + "'"; return false; }
  std::stringstream ss;
  ss << f.rdbuf();
  return flatten_mjcf_text(ss.str(), out, err);
}

namespace {
template <class A, class B> void cp(A& d, const B& s) { d = (A)s; }
template <class A, class B, size_t N> void cp(A (&d)[N], const B (&s)[N]) { for (size_t i = 0; i < N; i++) cp(d[i], s[i]); }
}  // namespace

template <typename T>
PlanarModel<T> cast_model(const PlanarModel<double>& s) {
  PlanarModel<T> o;
#define CP(f) cp(o.f, s.f)
  CP(pel_org); CP(pel_ref); CP(pel_com); CP(pel_mass); CP(pel_inertia);
  CP(off); CP(sgn); CP(ang0); CP(com); CP(mass); CP(inertia);
  CP(damping); CP(armature); CP(lim_lo); CP(lim_hi); CP(lim_diag); CP(has_limit); CP(lim_solref); CP(lim_solimp);
  CP(act_dof); CP(act_gear); CP(act_lo); CP(act_hi);
  CP(eq_a1); CP(eq_a2); CP(eq_solref); CP(eq_solimp); CP(eq_diag);
  CP(sph_c); CP(sph_r); CP(sph_diag); CP(cap_link); CP(cap_to); CP(cap_from); CP(cap_r); CP(cap_diag);
  CP(con_solref); CP(con_solimp); CP(con_mu);
  CP(site_link); CP(site_off);
  CP(timestep); CP(gravity_z); CP(tolerance); CP(meaninertia); CP(impratio); CP(iterations); CP(total_mass);
#undef CP
  return o;
}
template PlanarModel<float> cast_model<float>(const PlanarModel<double>&);
template PlanarModel<double> cast_model<double>(const PlanarModel<double>&);

}  // namespace cassie
