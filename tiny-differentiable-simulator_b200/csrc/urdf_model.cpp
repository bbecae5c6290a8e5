// Host-side model compiler: URDF text -> flat model (include/tds_b200_model.h).
//
// Behavioural mirror of the reference's setup path (run once per model, never per step):
//   UrdfParser::load_urdf_from_string      src/urdf/urdf_parser.hpp:729-925   (link ordering: pre-order DFS
//                                           from the root, children in joint document order, :677-706)
//   UrdfToMultiBody::convert_to_multi_body src/urdf/urdf_to_multi_body.hpp:41-277
//   MultiBody::initialize                  src/multi_body.hpp:324-378         (q / qd indices)
// including its quirks (SURVEY.md section 8 "parity traps" 11): only ixx/iyy/izz are read, the inertial
// rpy rotates both the inertia and the com by R^T, axes equal to exactly +1 become *_X/Y/Z joints,
// a missing <axis> defaults to (0,0,1), joint damping/stiffness are not transferred, mesh / cylinder
// collision shapes are dropped, the plane constant is 0; a <plane> collision shape on a link of the robot itself is kept like any
// other shape (record: P = unit normal) - it can only meet the shapes of another multibody.
// Own minimal XML reader (no third-party parser): elements, attributes, comments, declarations.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "tds_b200_model.h"

namespace {

struct XmlNode {
  std::string name;
  std::vector<std::pair<std::string, std::string>> attrs;
  std::vector<std::unique_ptr<XmlNode>> children;
  const char* attr(const char* key) const {
    for (auto& a : attrs)
      if (a.first == key) return a.second.c_str();
    return nullptr;
  }
  const XmlNode* child(const char* nm) const {
    for (auto& c : children)
      if (c->name == nm) return c.get();
    return nullptr;
  }
  const XmlNode* first_child() const { return children.empty() ? nullptr : children[0].get(); }
};

struct XmlReader {
  const std::string& s;
  size_t p = 0;
  std::string err;
  explicit XmlReader(const std::string& text) : s(text) {}
  void skip_ws() { while (p < s.size() && isspace((unsigned char)s[p])) ++p; }
  bool starts(const char* lit) const { return s.compare(p, strlen(lit), lit) == 0; }
  bool skip_misc() {  // whitespace, comments, declarations, text
    while (p < s.size()) {
      skip_ws();
      if (starts("<!--")) {
        size_t e = s.find("-->", p + 4);
        if (e == std::string::npos) { err = "unterminated comment"; return false; }
        p = e + 3;
      } else if (starts("<?")) {
        size_t e = s.find("?>", p + 2);
        if (e == std::string::npos) { err = "unterminated declaration"; return false; }
        p = e + 2;
      } else if (starts("<!")) {
        size_t e = s.find('>', p);
        if (e == std::string::npos) { err = "unterminated doctype"; return false; }
        p = e + 1;
      } else if (p < s.size() && s[p] != '<') {
        while (p < s.size() && s[p] != '<') ++p;  // character data: ignored
      } else {
        return true;
      }
    }
    return true;
  }
  static bool name_char(char c) { return isalnum((unsigned char)c) || c == '_' || c == '-' || c == ':' || c == '.'; }
  std::unique_ptr<XmlNode> parse_element() {
    if (p >= s.size() || s[p] != '<') { err = "expected '<'"; return nullptr; }
    ++p;
    auto node = std::make_unique<XmlNode>();
    size_t b = p;
    while (p < s.size() && name_char(s[p])) ++p;
    node->name = s.substr(b, p - b);
    if (node->name.empty()) { err = "empty element name"; return nullptr; }
    while (true) {
      skip_ws();
      if (p >= s.size()) { err = "unexpected end in tag"; return nullptr; }
      if (s[p] == '/') {
        if (p + 1 < s.size() && s[p + 1] == '>') { p += 2; return node; }
        err = "malformed empty tag"; return nullptr;
      }
      if (s[p] == '>') { ++p; break; }
      b = p;
      while (p < s.size() && name_char(s[p])) ++p;
      std::string key = s.substr(b, p - b);
      skip_ws();
      if (key.empty() || p >= s.size() || s[p] != '=') { err = "malformed attribute in <" + node->name + ">"; return nullptr; }
      ++p;
      skip_ws();
      if (p >= s.size() || (s[p] != '"' && s[p] != '\'')) { err = "attribute value must be quoted"; return nullptr; }
      char qc = s[p++];
      b = p;
      while (p < s.size() && s[p] != qc) ++p;
      if (p >= s.size()) { err = "unterminated attribute value"; return nullptr; }
      node->attrs.emplace_back(key, s.substr(b, p - b));
      ++p;
    }
    while (true) {
      if (!skip_misc()) return nullptr;
      if (p >= s.size()) { err = "missing </" + node->name + ">"; return nullptr; }
      if (starts("</")) {
        p += 2;
        b = p;
        while (p < s.size() && name_char(s[p])) ++p;
        if (s.substr(b, p - b) != node->name) { err = "mismatched </" + s.substr(b, p - b) + ">"; return nullptr; }
        skip_ws();
        if (p >= s.size() || s[p] != '>') { err = "malformed end tag"; return nullptr; }
        ++p;
        return node;
      }
      auto c = parse_element();
      if (!c) return nullptr;
      node->children.push_back(std::move(c));
    }
  }
  std::unique_ptr<XmlNode> parse_document(const char* root_name) {
    while (true) {
      if (!skip_misc()) return nullptr;
      if (p >= s.size()) { err = std::string("no <") + root_name + "> element"; return nullptr; }
      auto e = parse_element();
      if (!e) return nullptr;
      if (e->name == root_name) return e;
    }
  }
};

struct V3d { double v[3] = {0, 0, 0}; };

bool parse_v3(const char* str, V3d* out) {  // urdf_parser.hpp:77-97
  std::istringstream iss(str);
  std::string piece;
  int k = 0;
  double vals[3] = {0, 0, 0};
  while (iss >> piece) {
    if (k < 3) vals[k] = atof(piece.c_str());
    ++k;
  }
  if (k < 3) return false;
  for (int i = 0; i < 3; ++i) out->v[i] = vals[i];
  return true;
}

// TinyMatrix3x3::setEulerZYX(roll, pitch, yaw) (right-associative build), tiny_matrix3x3.h:192-215
void rpy_matrix(const double* rpy, double* R) {
  double ci = cos(rpy[0]), cj = cos(rpy[1]), ch = cos(rpy[2]);
  double si = sin(rpy[0]), sj = sin(rpy[1]), sh = sin(rpy[2]);
  double cc = ci * ch, cs = ci * sh, sc = si * ch, ss = si * sh;
  R[0] = cj * ch; R[1] = sj * sc - cs; R[2] = sj * cc + ss;
  R[3] = cj * sh; R[4] = sj * ss + cc; R[5] = sj * cs - sc;
  R[6] = -sj;     R[7] = cj * si;      R[8] = cj * ci;
}

struct Inertial { double mass = 0; V3d xxyyzz, rpy, xyz; };
struct Shape { int type = -1; double p[3] = {0, 0, 0}; V3d xyz, rpy; };
struct ULink { std::string name; Inertial inertial; std::vector<Shape> collisions, visuals; };
struct UJoint { std::string name, parent, child; int type = TDSJ_FIXED; V3d xyz, rpy, axis; };

bool parse_origin(const XmlNode* o, V3d* xyz, V3d* rpy, std::string* err) {
  if (!o) return true;
  if (o->attr("xyz") && !parse_v3(o->attr("xyz"), xyz)) { *err = "malformed origin xyz"; return false; }
  if (o->attr("rpy") && !parse_v3(o->attr("rpy"), rpy)) { *err = "malformed origin rpy"; return false; }
  return true;
}

bool parse_geometry(const XmlNode* geom, Shape* sh, std::string* err) {  // urdf_parser.hpp:164-258
  if (!geom || !geom->first_child()) { *err = "geometry tag contains no child element"; return false; }
  const XmlNode* s = geom->first_child();
  if (s->name == "sphere") {
    if (!s->attr("radius")) { *err = "sphere needs radius"; return false; }
    sh->type = TDSG_SPHERE; sh->p[0] = atof(s->attr("radius"));
  } else if (s->name == "box") {
    V3d e;
    if (!s->attr("size") || !parse_v3(s->attr("size"), &e)) { *err = "box needs size"; return false; }
    sh->type = TDSG_BOX; memcpy(sh->p, e.v, sizeof e.v);
  } else if (s->name == "capsule") {
    if (!s->attr("length") || !s->attr("radius")) { *err = "capsule needs length and radius"; return false; }
    sh->type = TDSG_CAPSULE; sh->p[0] = atof(s->attr("radius")); sh->p[1] = atof(s->attr("length"));
  } else if (s->name == "cylinder") {
    if (!s->attr("length") || !s->attr("radius")) { *err = "cylinder needs length and radius"; return false; }
    sh->type = 5;  // TINY_CYLINDER_TYPE: parsed, dropped by convert_collisions
  } else if (s->name == "mesh" || s->name == "cdf") {
    if (!s->attr("filename") || !s->attr("filename")[0]) { *err = "mesh filename is empty"; return false; }
    sh->type = TDSG_MESH;
  } else if (s->name == "plane") {
    V3d nrm;
    if (!s->attr("normal") || !parse_v3(s->attr("normal"), &nrm)) { *err = "plane requires a normal"; return false; }
    sh->type = TDSG_PLANE; memcpy(sh->p, nrm.v, sizeof nrm.v);
  } else {
    *err = "unknown geometry type " + s->name; return false;
  }
  return true;
}

bool parse_link(const XmlNode* x, ULink* l, std::string* err) {  // urdf_parser.hpp:353-464
  l->name = x->attr("name") ? x->attr("name") : "";
  if (l->name.empty()) { *err = "link with no name"; return false; }
  if (const XmlNode* i = x->child("inertial")) {
    if (!parse_origin(i->child("origin"), &l->inertial.xyz, &l->inertial.rpy, err)) return false;
    const XmlNode* m = i->child("mass");
    if (!m || !m->attr("value")) { *err = "inertial needs <mass value>"; return false; }
    l->inertial.mass = atof(m->attr("value"));
    const XmlNode* in = i->child("inertia");
    if (!in || !in->attr("ixx") || !in->attr("iyy") || !in->attr("izz")) { *err = "inertia needs ixx,iyy,izz"; return false; }
    l->inertial.xxyyzz.v[0] = atof(in->attr("ixx"));
    l->inertial.xxyyzz.v[1] = atof(in->attr("iyy"));
    l->inertial.xxyyzz.v[2] = atof(in->attr("izz"));
  }
  for (auto& c : x->children) {
    if (c->name != "visual" && c->name != "collision") continue;
    Shape sh;
    if (!parse_origin(c->child("origin"), &sh.xyz, &sh.rpy, err)) return false;
    if (!parse_geometry(c->child("geometry"), &sh, err)) { *err += " (link " + l->name + ")"; return false; }
    (c->name == "visual" ? l->visuals : l->collisions).push_back(sh);
  }
  return true;
}

bool parse_joint(const XmlNode* x, UJoint* j, std::string* err) {  // urdf_parser.hpp:466-675
  j->name = x->attr("name") ? x->attr("name") : "";
  if (j->name.empty()) { *err = "unnamed joint"; return false; }
  if (!parse_origin(x->child("origin"), &j->xyz, &j->rpy, err)) return false;
  const XmlNode* p = x->child("parent");
  const XmlNode* c = x->child("child");
  if (p) { if (!p->attr("link")) { *err = "joint parent without link"; return false; } j->parent = p->attr("link"); }
  if (c) { if (!c->attr("link")) { *err = "joint child without link"; return false; } j->child = c->attr("link"); }
  const char* t = x->attr("type");
  if (!t) { *err = "joint " + j->name + " has no type"; return false; }
  std::string ts = t;
  if (ts == "revolute" || ts == "continuous") j->type = TDSJ_REVOLUTE_AXIS;
  else if (ts == "prismatic") j->type = TDSJ_PRISMATIC_AXIS;
  else if (ts == "fixed") j->type = TDSJ_FIXED;
  else if (ts == "spherical") j->type = TDSJ_SPHERICAL;
  else { *err = "joint " + j->name + " has unsupported type " + ts; return false; }
  if (j->type != TDSJ_FIXED) {
    const XmlNode* a = x->child("axis");
    if (!a) { j->axis.v[0] = 0; j->axis.v[1] = 0; j->axis.v[2] = 1; }  // urdf_parser.hpp:600-602
    else if (a->attr("xyz") && !parse_v3(a->attr("xyz"), &j->axis)) { *err = "malformed axis of joint " + j->name; return false; }
  }
  return true;
}

void assign_links(const std::string& link_name, const std::vector<UJoint>& joints,
                  std::map<std::string, int>& index) {  // urdf_parser.hpp:677-706
  for (auto& j : joints) {
    if (j.parent != link_name) continue;
    int idx = (int)index.size() - 1;
    index[j.child] = idx;
    assign_links(j.child, joints, index);
  }
}

// RigidBodyInertia after Transform(rot=R(rpy)).apply(rbi): com' = R^T com, I' = R^T diag R
// (urdf_to_multi_body.hpp:54-66,177-190; transform.hpp:385-404 with zero translation).
void pack_rbi(const Inertial& in, double* rec /* mass, com[3], inertia[9] */) {
  double R[9];
  rpy_matrix(in.rpy.v, R);
  rec[0] = in.mass;
  for (int i = 0; i < 3; ++i) rec[1 + i] = R[0 * 3 + i] * in.xyz.v[0] + R[1 * 3 + i] * in.xyz.v[1] + R[2 * 3 + i] * in.xyz.v[2];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += R[k * 3 + i] * in.xxyyzz.v[k] * R[k * 3 + j];
      rec[4 + i * 3 + j] = s;
    }
}

thread_local std::string g_error;

}  // namespace

extern "C" const char* tds_b200_last_error(void) { return g_error.c_str(); }
extern "C" void tds_b200_set_error(const char* msg) { g_error = msg ? msg : ""; }

// urdf: path to a URDF file or (if it starts with '<') URDF text.  plane_urdf: likewise or NULL/"" for
// no ground plane (the reference creates the plane as multibody 0, locomotion_contact_simulation.h:108).
// Returns the number of doubles of the flat model (written to out if cap is large enough) or <0.
extern "C" int tds_b200_urdf_to_model(const char* urdf, const char* plane_urdf, int floating, double* out, int cap) {
  g_error.clear();
  auto load_text = [](const char* src, std::string* text) -> bool {
    if (!src) return false;
    const char* p = src;
    while (*p && isspace((unsigned char)*p)) ++p;
    if (*p == '<') { *text = src; return true; }
    std::ifstream ifs(src);
    if (!ifs.is_open()) { g_error = std::string("cannot open ") + src; return false; }
    *text = std::string((std::istreambuf_iterator<char>(ifs)), std::istreambuf_iterator<char>());
    return true;
  };
  std::string text;
  if (!load_text(urdf, &text)) { if (g_error.empty()) g_error = "no urdf given"; return -1; }
  XmlReader rd(text);
  auto robot = rd.parse_document("robot");
  if (!robot) { g_error = "XML error: " + rd.err; return -2; }
  if (!robot->attr("name")) { g_error = "expected a name for robot"; return -2; }
  std::vector<UJoint> joints;
  std::map<std::string, int> joint_names;
  std::map<std::string, std::string> link_to_joint;
  for (auto& c : robot->children) {
    if (c->name != "joint") continue;
    UJoint j;
    if (!parse_joint(c.get(), &j, &g_error)) return -3;
    if (joint_names.count(j.name)) { g_error = "joint " + j.name + " is not unique"; return -3; }
    joint_names[j.name] = (int)joints.size();
    link_to_joint[j.child] = j.name;
    joints.push_back(j);
  }
  std::vector<ULink> links;
  std::vector<std::string> roots;
  for (auto& c : robot->children) {
    if (c->name != "link") continue;
    ULink l;
    if (!parse_link(c.get(), &l, &g_error)) return -3;
    if (!link_to_joint.count(l.name)) roots.push_back(l.name);
    links.push_back(l);
  }
  if (roots.size() != 1) { g_error = roots.empty() ? "no parent link" : "multiple parent links"; return -4; }
  std::map<std::string, int> index;
  index[roots[0]] = -1;
  assign_links(roots[0], joints, index);
  if (index.size() != link_to_joint.size() + 1) { g_error = "inconsistent joint/link connections"; return -4; }
  const int n_links = (int)link_to_joint.size();
  std::vector<const ULink*> ordered(n_links, nullptr);
  const ULink* base = nullptr;
  for (auto& l : links) {
    auto it = index.find(l.name);
    if (it == index.end()) { g_error = "link inconsistency: " + l.name; return -4; }
    if (it->second >= 0) ordered[it->second] = &l; else base = &l;
  }
  std::vector<const UJoint*> ojoints(n_links, nullptr);
  for (auto& j : joints) ojoints[index[j.child]] = &j;

  // plane (static multibody 0)
  double plane_n[3] = {0, 0, 1};
  int has_plane = 0;
  if (plane_urdf && plane_urdf[0]) {
    std::string ptext;
    if (!load_text(plane_urdf, &ptext)) return -1;
    XmlReader prd(ptext);
    auto probot = prd.parse_document("robot");
    if (!probot) { g_error = "plane XML error: " + prd.err; return -2; }
    for (auto& c : probot->children) {
      if (c->name != "link") continue;
      ULink l;
      if (!parse_link(c.get(), &l, &g_error)) return -3;
      for (auto& sh : l.collisions)
        if (sh.type == TDSG_PLANE) { has_plane = 1; memcpy(plane_n, sh.p, sizeof plane_n); }
    }
    if (!has_plane) { g_error = "plane URDF has no <plane> collision shape"; return -5; }
    double len = sqrt(plane_n[0] * plane_n[0] + plane_n[1] * plane_n[1] + plane_n[2] * plane_n[2]);
    for (int k = 0; k < 3; ++k) plane_n[k] = plane_n[k] * (1.0 / len);  // Plane::set_normal -> normalize, geometry.hpp:183
  }

  // geoms / visuals in the reference's enumeration order
  std::vector<double> geoms, vis;
  auto add_shapes = [&](int link_index, const ULink& l) {
    for (auto& sh : l.collisions) {  // convert_collisions, urdf_to_multi_body.hpp:222-277
      if (sh.type != TDSG_SPHERE && sh.type != TDSG_BOX && sh.type != TDSG_CAPSULE && sh.type != TDSG_PLANE) continue;
      double rec[TDSM_GEOM] = {0};
      rec[TDSM_G_LINK] = link_index;
      rec[TDSM_G_TYPE] = sh.type;
      memcpy(rec + TDSM_G_P, sh.p, sizeof sh.p);
      if (sh.type == TDSG_PLANE) {   // a plane shape on a link of the robot: Plane::set_normal normalises (geometry.hpp:183)
        const double len = sqrt(sh.p[0] * sh.p[0] + sh.p[1] * sh.p[1] + sh.p[2] * sh.p[2]);
        for (int k = 0; k < 3; ++k) rec[TDSM_G_P + k] = sh.p[k] * (1.0 / len);
      }
      rpy_matrix(sh.rpy.v, rec + TDSM_G_R);
      memcpy(rec + TDSM_G_T, sh.xyz.v, sizeof sh.xyz.v);
      geoms.insert(geoms.end(), rec, rec + TDSM_GEOM);
    }
    if (link_index >= 0)
      for (auto& sh : l.visuals) {
        double rec[TDSM_VIS] = {0};
        rec[TDSM_V_LINK] = link_index;
        rpy_matrix(sh.rpy.v, rec + TDSM_V_R);
        memcpy(rec + TDSM_V_T, sh.xyz.v, sizeof sh.xyz.v);
        vis.insert(vis.end(), rec, rec + TDSM_VIS);
      }
  };
  add_shapes(-1, *base);
  for (int i = 0; i < n_links; ++i) add_shapes(i, *ordered[i]);
  const int n_geoms = (int)(geoms.size() / TDSM_GEOM), n_vis = (int)(vis.size() / TDSM_VIS);
  const int total = TDSM_HEADER + TDSM_BASE + n_links * TDSM_LINK + n_geoms * TDSM_GEOM + n_vis * TDSM_VIS;
  for (int i = 0; i < n_links; ++i) {  // validate before sizing so that errors surface on the first call
    const UJoint& j = *ojoints[i];
    if ((j.type == TDSJ_REVOLUTE_AXIS || j.type == TDSJ_PRISMATIC_AXIS) &&
        j.axis.v[0] == 0.0 && j.axis.v[1] == 0.0 && j.axis.v[2] == 0.0) { g_error = "zero joint axis on " + j.name; return -6; }
  }
  if (!out || cap < total) return total;
  memset(out, 0, sizeof(double) * total);
  double* b = out + TDSM_HEADER;
  pack_rbi(base->inertial, b);
  double* L = b + TDSM_BASE;
  int q_index = floating ? 7 : 0, qd_index = floating ? 6 : 0;  // MultiBody::initialize, multi_body.hpp:324-349
  for (int i = 0; i < n_links; ++i) {
    const UJoint& j = *ojoints[i];
    double* r = L + (size_t)i * TDSM_LINK;
    int jt = j.type;
    if (jt == TDSJ_REVOLUTE_AXIS || jt == TDSJ_PRISMATIC_AXIS) {  // urdf_to_multi_body.hpp:115-160
      int nz = -1;
      for (int k = 0; k < 3; ++k)
        if (j.axis.v[k] == 1.0) { if (nz >= 0) break; nz = k; }
      if (nz >= 0) {
        jt = (jt == TDSJ_REVOLUTE_AXIS ? TDSJ_REVOLUTE_X : TDSJ_PRISMATIC_X) + nz;
        r[TDSM_L_AXIS + nz] = 1.0;
      } else {
        double nrm = sqrt(j.axis.v[0] * j.axis.v[0] + j.axis.v[1] * j.axis.v[1] + j.axis.v[2] * j.axis.v[2]);
        if (nrm == 0.0) { g_error = "zero joint axis on " + j.name; return -6; }
        memcpy(r + TDSM_L_AXIS, j.axis.v, sizeof j.axis.v);
      }
    }
    r[TDSM_L_PARENT] = index[j.parent];
    r[TDSM_L_JTYPE] = jt;
    if (jt == TDSJ_FIXED) { r[TDSM_L_QIDX] = -2; r[TDSM_L_QDIDX] = -2; }
    else if (jt == TDSJ_SPHERICAL) {   // quaternion xyzw: 4 coordinates, 3 velocities (multi_body.hpp:324-349)
      r[TDSM_L_QIDX] = q_index; r[TDSM_L_QDIDX] = qd_index;
      q_index += 4; qd_index += 3;
    } else { r[TDSM_L_QIDX] = q_index++; r[TDSM_L_QDIDX] = qd_index++; }
    rpy_matrix(j.rpy.v, r + TDSM_L_XT_R);
    memcpy(r + TDSM_L_XT_T, j.xyz.v, sizeof j.xyz.v);
    pack_rbi(ordered[i]->inertial, r + TDSM_L_MASS);
  }
  out[TDSM_H_MAGIC] = TDSM_MAGIC;
  out[TDSM_H_NLINKS] = n_links;
  out[TDSM_H_FLOATING] = floating ? 1 : 0;
  out[TDSM_H_NQ] = q_index;
  out[TDSM_H_NQD] = qd_index;
  out[TDSM_H_NGEOMS] = n_geoms;
  out[TDSM_H_NVIS] = n_vis;
  out[TDSM_H_HASPLANE] = has_plane;
  if (has_plane) memcpy(out + TDSM_H_PLANE_N, plane_n, sizeof plane_n);
  out[TDSM_H_PLANE_C] = 0.0;
  memcpy(L + (size_t)n_links * TDSM_LINK, geoms.data(), sizeof(double) * geoms.size());
  memcpy(L + (size_t)n_links * TDSM_LINK + geoms.size(), vis.data(), sizeof(double) * vis.size());
  return total;
}
