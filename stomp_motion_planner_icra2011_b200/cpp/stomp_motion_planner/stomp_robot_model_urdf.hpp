// stomp_robot_model_urdf.hpp — StompRobotModel::init for the B200 engine, host C++ above the C ABI: URDF text -> the tables
// stomp_engine_set_robot / set_dynamics / build_sdf_bodies take.  No ROS, urdf, kdl_parser or tinyxml exist here, so this file
// carries its own small XML reader and restates what the reference obtains from those packages:
//   * kdl_parser: every URDF joint becomes one KDL segment named after its child link,
//     pose(q) = Frame(Rot(R_pj a, q) R_pj, p_pj) with (R_pj, p_pj) the joint <origin> and a the joint <axis>; segment numbers
//     are the DFS pre-order over children in document order (TreeFkSolverJointPosAxisPartial::assignSegmentNumber,
//     src/treefksolverjointposaxis_partial.cpp:180-191)
//   * StompRobotModel::init (src/stomp_robot_model.cpp:58-226): joint limits (continuous joints have none, :160-161), the
//     planning group's joints, the inverse-dynamics chain (:181-185)
//   * generateLinkCollisionPoints / addCollisionPointsFromLinkRadius (:228-306): one sphere every radius / 2 from the link
//     origin to each child's KDL JointOrigin() (zero for a fixed child joint), ceil(distance / spacing) + 1 points, the first
//     point of every child after the first skipped; StompPlanningGroup::addCollisionPoint keeps the points a group joint moves
//   * generateAttachedObjectCollisionPoints (:377-453): the bounding sphere of every attached shape as one more collision point
//   * the links' <collision> primitives, for StompCollisionSpace::setStartState's voxelisation of the robot bodies outside
//     the planning group (src/stomp_collision_space.cpp:167-188,567-588): bodiesAtState() places them at a joint state
// Pinned by tests/test_urdf_cpp_cpu.py against stomp_motion_planner_icra2011_b200/urdf.py (which tests/test_urdf_cpu.py and
// tests/test_reference_pinning.py hold to the compiled reference's sphere generator).
#pragma once
#include <cmath>
#include <cctype>
#include <cstdlib>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <utility>
#include <vector>

#include "stomp_b200_facade.hpp"

namespace stomp_motion_planner {

// ---- a reader for the XML subset URDF uses: elements, attributes, comments, declarations; text content is skipped ------------
struct XmlElement {
  std::string name;
  std::vector<std::pair<std::string, std::string> > attributes;
  std::vector<std::unique_ptr<XmlElement> > children;
  const std::string* attr(const std::string& key) const {
    for (const auto& kv : attributes)
      if (kv.first == key) return &kv.second;
    return nullptr;
  }
  const XmlElement* child(const std::string& n) const {
    for (const auto& c : children)
      if (c->name == n) return c.get();
    return nullptr;
  }
};

class XmlReader {
 public:
  explicit XmlReader(const std::string& text) : s_(text) {}
  std::unique_ptr<XmlElement> parse(std::string& err) {
    skipMisc();
    std::unique_ptr<XmlElement> root = element(err);
    if (!root && err.empty()) err = "no root element";
    return root;
  }

 private:
  const std::string& s_;
  size_t i_ = 0;
  bool starts(const char* lit) const { return s_.compare(i_, std::char_traits<char>::length(lit), lit) == 0; }
  void skipSpace() { while (i_ < s_.size() && std::isspace(static_cast<unsigned char>(s_[i_]))) ++i_; }
  void skipMisc() {   // whitespace, text, comments, processing instructions, doctype
    for (;;) {
      while (i_ < s_.size() && s_[i_] != '<') ++i_;
      if (i_ >= s_.size()) return;
      if (starts("<!--")) { size_t e = s_.find("-->", i_ + 4); i_ = e == std::string::npos ? s_.size() : e + 3; continue; }
      if (starts("<?")) { size_t e = s_.find("?>", i_ + 2); i_ = e == std::string::npos ? s_.size() : e + 2; continue; }
      if (starts("<!")) { size_t e = s_.find('>', i_ + 2); i_ = e == std::string::npos ? s_.size() : e + 1; continue; }
      return;
    }
  }
  std::string name() {
    size_t b = i_;
    while (i_ < s_.size() && (std::isalnum(static_cast<unsigned char>(s_[i_])) || s_[i_] == '_' || s_[i_] == ':' || s_[i_] == '-' || s_[i_] == '.')) ++i_;
    return s_.substr(b, i_ - b);
  }
  std::unique_ptr<XmlElement> element(std::string& err) {
    if (i_ >= s_.size() || s_[i_] != '<') { err = "expected '<'"; return nullptr; }
    ++i_;
    std::unique_ptr<XmlElement> el(new XmlElement());
    el->name = name();
    if (el->name.empty()) { err = "element without a name"; return nullptr; }
    for (;;) {
      skipSpace();
      if (i_ >= s_.size()) { err = "unterminated element <" + el->name + ">"; return nullptr; }
      if (starts("/>")) { i_ += 2; return el; }
      if (s_[i_] == '>') { ++i_; break; }
      std::string key = name();
      skipSpace();
      if (key.empty() || i_ >= s_.size() || s_[i_] != '=') { err = "bad attribute in <" + el->name + ">"; return nullptr; }
      ++i_;
      skipSpace();
      if (i_ >= s_.size() || (s_[i_] != '"' && s_[i_] != '\'')) { err = "attribute value of <" + el->name + "> is not quoted"; return nullptr; }
      const char q = s_[i_++];
      size_t e = s_.find(q, i_);
      if (e == std::string::npos) { err = "unterminated attribute value in <" + el->name + ">"; return nullptr; }
      el->attributes.emplace_back(key, s_.substr(i_, e - i_));
      i_ = e + 1;
    }
    for (;;) {   // children until the closing tag
      skipMisc();
      if (i_ >= s_.size()) { err = "missing </" + el->name + ">"; return nullptr; }
      if (starts("</")) {
        i_ += 2;
        const std::string closing = name();
        skipSpace();
        if (closing != el->name || i_ >= s_.size() || s_[i_] != '>') { err = "mismatched </" + closing + "> for <" + el->name + ">"; return nullptr; }
        ++i_;
        return el;
      }
      std::unique_ptr<XmlElement> c = element(err);
      if (!c) return nullptr;
      el->children.push_back(std::move(c));
    }
  }
};

// ---- the model -------------------------------------------------------------------------------------------------------
struct CollisionLinkConfig {      // collision_links/<link>/{link_radius, link_clearance, link_extension} (stomp_robot_model.cpp:361-373)
  std::string link;
  double link_radius = 0.0, link_extension = 0.0;
  double link_clearance = -1.0;   // < 0: the model-wide collision_clearance
};

constexpr int kLinkBodyMesh = 3;  // LinkBody::type of a <mesh> geometry (the primitives use stomp_body_type)
struct LinkBody {                 // one <collision> geometry of a link, in the link (segment) frame
  int segment = -1;
  int type = STOMP_BODY_BOX;
  double dimensions[3] = {0, 0, 0};
  double rot[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, pos[3] = {0, 0, 0};
  // <mesh filename="..." scale="sx sy sz"/>: the file is read by loadLinkMeshes; vertices [n][3] in the link body's frame,
  // already multiplied by the URDF scale (what the reference receives from planning_models as a shapes::Mesh)
  std::string mesh_filename;
  double mesh_scale[3] = {1, 1, 1};
  std::vector<double> mesh_vertices;
};

struct StompRobotModelUrdf : StompRobotModel {
  std::vector<std::string> segment_names, segment_joint_names;   // child link / joint of every segment (root: link, "")
  std::vector<LinkBody> link_bodies;
  int segmentOfLink(const std::string& link) const {
    for (size_t i = 0; i < segment_names.size(); ++i)
      if (segment_names[i] == link) return int(i);
    return -1;
  }
};

namespace urdf_detail {
inline bool numbers(const std::string* text, int n, const double* dflt, double* out, std::string& err) {
  if (!text) { for (int i = 0; i < n; ++i) out[i] = dflt[i]; return true; }
  std::istringstream is(*text);
  int k = 0;
  double v;
  while (is >> v) { if (k < n) out[k] = v; ++k; }
  if (k != n) { err = "expected " + std::to_string(n) + " numbers, got '" + *text + "'"; return false; }
  return true;
}
inline void rpyMatrix(double r, double p, double y, double* R) {   // urdf::Rotation::setFromRPY -> KDL::Rotation (fixed axes)
  const double cr = std::cos(r), sr = std::sin(r), cp = std::cos(p), sp = std::sin(p), cy = std::cos(y), sy = std::sin(y);
  const double m[9] = {cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr, sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr,
                       -sp, cp * sr, cp * cr};
  for (int i = 0; i < 9; ++i) R[i] = m[i];
}
inline bool origin(const XmlElement* el, double* xyz, double* R, std::string& err) {
  static const double zero[3] = {0, 0, 0};
  double rpy[3];
  const XmlElement* o = el ? el->child("origin") : nullptr;
  if (!numbers(o ? o->attr("xyz") : nullptr, 3, zero, xyz, err) || !numbers(o ? o->attr("rpy") : nullptr, 3, zero, rpy, err)) return false;
  rpyMatrix(rpy[0], rpy[1], rpy[2], R);
  return true;
}
inline double number(const std::string* s, double dflt) { return s ? std::atof(s->c_str()) : dflt; }
struct Joint {
  std::string name, type, parent, child;
  double xyz[3], rot[9], axis[3], lower = 0.0, upper = 0.0;
};
inline void matmul3(const double* A, const double* B, double* C) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) C[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
}
}  // namespace urdf_detail

// group_joints: the planning group's joint names in order (planning_groups.yaml); collision_links: in the order their spheres
// are to be generated; joint_state: values of the joints outside the group (the robot start state); dynamics_chain: root and tip
// LINK of the inverse-dynamics chain ("" = none; the reference hard-codes the PR2's torso_lift_link .. r_gripper_tool_frame).
inline bool loadRobotModelFromUrdf(const std::string& urdf_xml, const std::vector<std::string>& group_joints,
                                   const std::string& reference_frame, const std::vector<CollisionLinkConfig>& collision_links,
                                   double collision_clearance, const std::map<std::string, double>& joint_state,
                                   const std::string& chain_root_link, const std::string& chain_tip_link, StompRobotModelUrdf& out,
                                   std::string& err) {
  using namespace urdf_detail;
  XmlReader reader(urdf_xml);
  std::unique_ptr<XmlElement> robot = reader.parse(err);
  if (!robot) return false;
  if (robot->name != "robot") { err = "root element is <" + robot->name + ">, not <robot>"; return false; }
  std::vector<const XmlElement*> links;
  std::vector<Joint> joints;
  for (const auto& c : robot->children) {
    if (c->name == "link") {
      if (!c->attr("name")) { err = "link without a name"; return false; }
      links.push_back(c.get());
    } else if (c->name == "joint") {
      Joint j;
      const XmlElement *par = c->child("parent"), *chi = c->child("child");
      if (!c->attr("name") || !c->attr("type") || !par || !chi || !par->attr("link") || !chi->attr("link")) { err = "incomplete joint"; return false; }
      j.name = *c->attr("name"); j.type = *c->attr("type"); j.parent = *par->attr("link"); j.child = *chi->attr("link");
      static const char* known[] = {"revolute", "continuous", "prismatic", "fixed", "floating", "planar"};
      bool ok = false;
      for (const char* k : known) ok = ok || j.type == k;
      if (!ok) { err = "unknown joint type '" + j.type + "'"; return false; }
      if (!origin(c.get(), j.xyz, j.rot, err)) return false;
      static const double x_axis[3] = {1, 0, 0};
      const XmlElement* ax = c->child("axis");
      if (!numbers(ax ? ax->attr("xyz") : nullptr, 3, x_axis, j.axis, err)) return false;
      const XmlElement* lim = c->child("limit");
      j.lower = number(lim ? lim->attr("lower") : nullptr, 0.0);
      j.upper = number(lim ? lim->attr("upper") : nullptr, 0.0);
      joints.push_back(j);
    }
  }
  std::map<std::string, std::vector<int> > children_of;
  std::map<std::string, int> joint_by_name;
  std::map<std::string, bool> is_child;
  for (size_t k = 0; k < joints.size(); ++k) {
    children_of[joints[k].parent].push_back(int(k));
    joint_by_name[joints[k].name] = int(k);
    is_child[joints[k].child] = true;
  }
  std::vector<std::string> roots;
  for (const XmlElement* l : links)
    if (!is_child.count(*l->attr("name"))) roots.push_back(*l->attr("name"));
  if (roots.size() != 1) { err = "URDF must have exactly one root link, found " + std::to_string(roots.size()); return false; }
  std::map<std::string, int> group_index;
  for (size_t g = 0; g < group_joints.size(); ++g) {
    if (!joint_by_name.count(group_joints[g])) { err = "planning group joints not in the URDF: " + group_joints[g]; return false; }
    group_index[group_joints[g]] = int(g);
  }
  out = StompRobotModelUrdf();
  // DFS pre-order, children in document order (explicit stack: deep chains must not recurse)
  struct Item { std::string link; int parent_seg; int joint; };
  std::vector<Item> stack;
  stack.push_back(Item{roots[0], -1, -1});
  while (!stack.empty()) {
    const Item it = stack.back();
    stack.pop_back();
    stomp_segment g;
    std::memset(&g, 0, sizeof(g));
    g.parent = it.parent_seg;
    g.group_index = -1;
    g.rot[0] = g.rot[4] = g.rot[8] = 1.0;
    g.axis[2] = 1.0;
    std::string joint_name;
    if (it.joint < 0) {
      g.joint_type = STOMP_JOINT_FIXED;      // KDL tree root segment
    } else {
      const Joint& j = joints[it.joint];
      joint_name = j.name;
      g.joint_type = (j.type == "revolute" || j.type == "continuous") ? STOMP_JOINT_REVOLUTE
                     : j.type == "prismatic" ? STOMP_JOINT_PRISMATIC : STOMP_JOINT_FIXED;
      for (int k = 0; k < 9; ++k) g.rot[k] = j.rot[k];
      for (int k = 0; k < 3; ++k) g.pos[k] = j.xyz[k];
      if (g.joint_type != STOMP_JOINT_FIXED) {
        // kdl_parser: the axis, normalised, expressed in the parent frame (R_pj * a); then normalised once more like
        // scenes.Robot.add_segment does, so that both builders produce the same bits
        double n = std::sqrt(j.axis[0] * j.axis[0] + j.axis[1] * j.axis[1] + j.axis[2] * j.axis[2]);
        if (n == 0.0) n = 1.0;
        const double a[3] = {j.axis[0] / n, j.axis[1] / n, j.axis[2] / n};
        double b[3];
        for (int r = 0; r < 3; ++r) b[r] = j.rot[r * 3] * a[0] + j.rot[r * 3 + 1] * a[1] + j.rot[r * 3 + 2] * a[2];
        const double n2 = std::sqrt(b[0] * b[0] + b[1] * b[1] + b[2] * b[2]);
        for (int r = 0; r < 3; ++r) g.axis[r] = b[r] / n2;
        const auto gi = group_index.find(j.name);
        g.group_index = gi == group_index.end() ? -1 : gi->second;
      }
      const auto js = joint_state.find(j.name);
      g.fixed_value = js == joint_state.end() ? 0.0 : js->second;
    }
    const int seg = int(out.segments.size());
    out.segments.push_back(g);
    out.segment_names.push_back(it.link);
    out.segment_joint_names.push_back(joint_name);
    const auto ch = children_of.find(it.link);
    if (ch != children_of.end())
      for (size_t k = ch->second.size(); k-- > 0;) stack.push_back(Item{joints[ch->second[k]].child, seg, ch->second[k]});
  }
  out.reference_segment = out.segmentOfLink(reference_frame);
  if (out.reference_segment < 0) { err = "reference frame '" + reference_frame + "' is not a link of the URDF"; return false; }
  for (const std::string& name : group_joints) {
    const Joint& j = joints[joint_by_name[name]];
    stomp_joint_limit lim;
    std::memset(&lim, 0, sizeof(lim));
    if (j.type != "continuous") { lim.has_limits = 1; lim.min = j.lower; lim.max = j.upper; }
    out.joint_limits.push_back(lim);
  }
  // collision spheres (addCollisionPointsFromLinkRadius)
  auto moved_by_group = [&](int seg) {
    for (; seg >= 0; seg = out.segments[seg].parent)
      if (out.segments[seg].group_index >= 0) return true;
    return false;
  };
  for (const CollisionLinkConfig& cfg : collision_links) {
    const int seg = out.segmentOfLink(cfg.link);
    if (seg < 0) continue;                       // links outside the model are ignored like in the reference
    const double radius = cfg.link_radius, clearance = cfg.link_clearance < 0.0 ? collision_clearance : cfg.link_clearance;
    bool first_child = true;
    for (size_t c = 0; c < out.segments.size(); ++c) {
      if (out.segments[c].parent != seg) continue;
      const bool fixed = out.segments[c].joint_type == STOMP_JOINT_FIXED;
      double o[3] = {0, 0, 0};
      if (!fixed) for (int k = 0; k < 3; ++k) o[k] = out.segments[c].pos[k];
      const double spacing = radius / 2.0;
      const double distance = std::sqrt(o[0] * o[0] + o[1] * o[1] + o[2] * o[2]) + cfg.link_extension;
      const int num_points = int(std::ceil(distance / spacing)) + 1;
      for (int i = 0; i < num_points; ++i) {
        if (!first_child && i == 0) continue;
        stomp_sphere sp;
        std::memset(&sp, 0, sizeof(sp));
        sp.segment = seg; sp.radius = radius; sp.clearance = clearance;
        const double f = num_points > 1 ? i / (num_points - 1.0) : 0.0;
        for (int k = 0; k < 3; ++k) sp.pos[k] = o[k] * f;
        if (moved_by_group(seg)) out.collision_points.push_back(sp);
      }
      first_child = false;
    }
  }
  // link inertias (<inertial>): mass, centre of mass and the inertia tensor turned into the link's axes
  out.link_inertias.assign(out.segments.size(), stomp_link_inertia());
  for (auto& li : out.link_inertias) std::memset(&li, 0, sizeof(li));
  for (const XmlElement* l : links) {
    const int seg = out.segmentOfLink(*l->attr("name"));
    if (seg < 0) continue;
    const XmlElement* in = l->child("inertial");
    if (in && in->child("mass")) {
      double xyz[3], R[9];
      if (!origin(in, xyz, R, err)) return false;
      const XmlElement* it = in->child("inertia");
      const double ixx = number(it ? it->attr("ixx") : nullptr, 0), iyy = number(it ? it->attr("iyy") : nullptr, 0),
                   izz = number(it ? it->attr("izz") : nullptr, 0), ixy = number(it ? it->attr("ixy") : nullptr, 0),
                   ixz = number(it ? it->attr("ixz") : nullptr, 0), iyz = number(it ? it->attr("iyz") : nullptr, 0);
      const double Ic[9] = {ixx, ixy, ixz, ixy, iyy, iyz, ixz, iyz, izz};
      double Rt[9], T[9], Il[9];
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) Rt[r * 3 + c] = R[c * 3 + r];
      matmul3(R, Ic, T);
      matmul3(T, Rt, Il);
      stomp_link_inertia& li = out.link_inertias[seg];
      li.mass = number(in->child("mass")->attr("value"), 0.0);
      for (int k = 0; k < 3; ++k) li.com[k] = xyz[k];
      li.inertia[0] = Il[0]; li.inertia[1] = Il[4]; li.inertia[2] = Il[8]; li.inertia[3] = Il[1]; li.inertia[4] = Il[2]; li.inertia[5] = Il[5];
    }
    // <collision> geometries (a mesh keeps its file name; loadLinkMeshes reads the vertices)
    for (const auto& c : l->children) {
      if (c->name != "collision") continue;
      const XmlElement* geo = c->child("geometry");
      if (!geo) continue;
      LinkBody b;
      b.segment = seg;
      if (!origin(c.get(), b.pos, b.rot, err)) return false;
      static const double zero3[3] = {0, 0, 0};
      if (const XmlElement* box = geo->child("box")) {
        b.type = STOMP_BODY_BOX;
        if (!numbers(box->attr("size"), 3, zero3, b.dimensions, err)) return false;
      } else if (const XmlElement* cyl = geo->child("cylinder")) {
        b.type = STOMP_BODY_CYLINDER;
        b.dimensions[0] = number(cyl->attr("radius"), 0.0);
        b.dimensions[1] = number(cyl->attr("length"), 0.0);
      } else if (const XmlElement* sph = geo->child("sphere")) {
        b.type = STOMP_BODY_SPHERE;
        b.dimensions[0] = number(sph->attr("radius"), 0.0);
      } else if (const XmlElement* mesh = geo->child("mesh")) {
        static const double one3[3] = {1, 1, 1};
        b.type = kLinkBodyMesh;
        if (const std::string* fn = mesh->attr("filename")) b.mesh_filename = *fn;
        if (!numbers(mesh->attr("scale"), 3, one3, b.mesh_scale, err)) return false;
      } else {
        continue;
      }
      out.link_bodies.push_back(b);
    }
  }
  if (!chain_root_link.empty() || !chain_tip_link.empty()) {
    out.chain_root_segment = out.segmentOfLink(chain_root_link);
    out.chain_tip_segment = out.segmentOfLink(chain_tip_link);
    if (out.chain_root_segment < 0 || out.chain_tip_segment < 0) { err = "dynamics chain link is not a link of the URDF"; return false; }
  }
  return true;
}

// StompRobotModel::generateAttachedObjectCollisionPoints (src/stomp_robot_model.cpp:377-453): every shape attached to a link
// becomes ONE collision point of that link = the bounding sphere of the padded shape (geometric_shapes'
// bodies::computeBoundingSphere, restated: sphere r; box sqrt(sum (d/2 + pad)^2); cylinder sqrt((r + pad)^2 + (l/2 + pad)^2)),
// centred at the shape's position in the link frame, with the default clearance.  Kept only if a group joint moves the link
// (StompPlanningGroup::addCollisionPoint).  Returns false for an unknown link or shape type.
inline bool addAttachedObjectCollisionPoint(StompRobotModelUrdf& m, const std::string& link, int body_type, const double dimensions[3],
                                            const double position[3], double padding, double clearance) {
  const int seg = m.segmentOfLink(link);
  if (seg < 0) return false;
  double radius;
  if (body_type == STOMP_BODY_SPHERE) {
    radius = dimensions[0] + padding;
  } else if (body_type == STOMP_BODY_BOX) {
    double sum = 0.0;
    for (int k = 0; k < 3; ++k) sum += (dimensions[k] / 2.0 + padding) * (dimensions[k] / 2.0 + padding);
    radius = std::sqrt(sum);
  } else if (body_type == STOMP_BODY_CYLINDER) {
    radius = std::sqrt((dimensions[0] + padding) * (dimensions[0] + padding) + (dimensions[1] / 2.0 + padding) * (dimensions[1] / 2.0 + padding));
  } else {
    return false;
  }
  bool moved = false;
  for (int sg = seg; sg >= 0; sg = m.segments[sg].parent) moved = moved || m.segments[sg].group_index >= 0;
  if (!moved) return true;
  stomp_sphere sp;
  std::memset(&sp, 0, sizeof(sp));
  sp.segment = seg; sp.radius = radius; sp.clearance = clearance;
  for (int k = 0; k < 3; ++k) sp.pos[k] = position[k];
  m.collision_points.push_back(sp);
  return true;
}

// World (reference-frame) poses of the links' collision bodies at a joint state: what StompCollisionSpace::updateRobotBodiesPoses
// + addAllBodiesButExcludeLinksToPoints hand to getVoxelsInBody (src/stomp_collision_space.cpp:522-588).  group_values: the
// planning group's joints (the start state); the other joints take their fixed_value.  exclude_links: the group's
// distance_exclude_links (its own links).  Quaternions are x, y, z, w.
// STL reader for <mesh> collision geometry (binary: 80-byte header, uint32 count, 50 bytes per facet; ASCII: "vertex x y z"
// lines).  Only the vertices matter: the reference turns a mesh into the convex hull of its vertices (bodies::ConvexMesh).
inline bool readStlVertices(const std::string& path, std::vector<double>& vertices, std::string& err) {
  std::FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) { err = "cannot open mesh file " + path; return false; }
  std::vector<unsigned char> buf;
  unsigned char chunk[65536];
  size_t got;
  while ((got = std::fread(chunk, 1, sizeof(chunk), f)) > 0) buf.insert(buf.end(), chunk, chunk + got);
  std::fclose(f);
  vertices.clear();
  uint32_t count = 0;
  if (buf.size() >= 84) std::memcpy(&count, &buf[80], 4);
  if (buf.size() >= 84 && buf.size() == 84 + size_t(count) * 50) {
    for (uint32_t t = 0; t < count; ++t)
      for (int v = 0; v < 3; ++v)
        for (int k = 0; k < 3; ++k) {
          float x;
          std::memcpy(&x, &buf[84 + size_t(t) * 50 + 12 + size_t(v) * 12 + size_t(k) * 4], 4);
          vertices.push_back(double(x));
        }
  } else {
    const std::string text(buf.begin(), buf.end());
    size_t at = 0;
    while ((at = text.find("vertex", at)) != std::string::npos) {
      at += 6;
      const char* c = text.c_str() + at;
      char* end = nullptr;
      double xyz[3];
      bool ok = true;
      for (int k = 0; k < 3 && ok; ++k) { xyz[k] = std::strtod(c, &end); ok = end != c; c = end; }
      if (ok) vertices.insert(vertices.end(), xyz, xyz + 3);
    }
  }
  if (vertices.size() < 12) { err = "mesh file " + path + " holds fewer than 4 vertices"; return false; }
  return true;
}

// Reads the vertices of every <mesh> collision geometry (URDF scale applied).  `resolve` maps the URDF's file name
// ("package://pr2_description/meshes/...") to a path on disk; the default takes it literally.
template <typename Resolve>
inline bool loadLinkMeshes(StompRobotModelUrdf& m, Resolve resolve, std::string& err) {
  for (LinkBody& lb : m.link_bodies) {
    if (lb.type != kLinkBodyMesh || !lb.mesh_vertices.empty()) continue;
    if (!readStlVertices(resolve(lb.mesh_filename), lb.mesh_vertices, err)) return false;
    for (size_t v = 0; v < lb.mesh_vertices.size(); ++v) lb.mesh_vertices[v] *= lb.mesh_scale[v % 3];
  }
  return true;
}
inline bool loadLinkMeshes(StompRobotModelUrdf& m, std::string& err) {
  return loadLinkMeshes(m, [](const std::string& name) { return name; }, err);
}

namespace urdf_detail {
// pose of every collision geometry of the links (not excluded) at a joint state, in the reference segment's frame:
// visit(link_body, position[3], quaternion_xyzw[4])
template <typename Visit>
inline void linkBodyPoses(const StompRobotModelUrdf& m, const std::vector<double>& group_values,
                          const std::vector<std::string>& exclude_links, Visit visit) {
  const size_t S = m.segments.size();
  std::vector<double> R(S * 9), p(S * 3);
  for (size_t s = 0; s < S; ++s) {          // DFS pre-order: parents come first
    const stomp_segment& g = m.segments[s];
    const double q = g.group_index >= 0 && size_t(g.group_index) < group_values.size() ? group_values[g.group_index] : g.fixed_value;
    double Rl[9], pl[3] = {g.pos[0], g.pos[1], g.pos[2]};
    if (g.joint_type == STOMP_JOINT_REVOLUTE) {   // Frame(Rot(axis, q) * rot, pos)
      const double c = std::cos(q), sn = std::sin(q), v = 1.0 - c, x = g.axis[0], y = g.axis[1], z = g.axis[2];
      const double Rq[9] = {c + v * x * x, v * x * y - sn * z, v * x * z + sn * y, v * x * y + sn * z, c + v * y * y, v * y * z - sn * x,
                            v * x * z - sn * y, v * y * z + sn * x, c + v * z * z};
      matmul3(Rq, g.rot, Rl);
    } else {
      for (int k = 0; k < 9; ++k) Rl[k] = g.rot[k];
      if (g.joint_type == STOMP_JOINT_PRISMATIC)
        for (int k = 0; k < 3; ++k) pl[k] += q * g.axis[k];
    }
    if (g.parent < 0) {
      for (int k = 0; k < 9; ++k) R[s * 9 + k] = Rl[k];
      for (int k = 0; k < 3; ++k) p[s * 3 + k] = pl[k];
    } else {
      const double* Rp = &R[size_t(g.parent) * 9];
      matmul3(Rp, Rl, &R[s * 9]);
      for (int k = 0; k < 3; ++k) p[s * 3 + k] = Rp[k * 3] * pl[0] + Rp[k * 3 + 1] * pl[1] + Rp[k * 3 + 2] * pl[2] + p[size_t(g.parent) * 3 + k];
    }
  }
  // express everything in the reference segment's frame
  const double* Rr = &R[size_t(m.reference_segment) * 9];
  const double* pr = &p[size_t(m.reference_segment) * 3];
  for (const LinkBody& lb : m.link_bodies) {
    bool excluded = false;
    for (const std::string& e : exclude_links) excluded = excluded || m.segment_names[lb.segment] == e;
    if (excluded) continue;
    const double* Rs = &R[size_t(lb.segment) * 9];
    const double* ps = &p[size_t(lb.segment) * 3];
    double Rw[9], pw[3], Rb[9], pb[3];
    matmul3(Rs, lb.rot, Rw);
    for (int k = 0; k < 3; ++k) pw[k] = Rs[k * 3] * lb.pos[0] + Rs[k * 3 + 1] * lb.pos[1] + Rs[k * 3 + 2] * lb.pos[2] + ps[k];
    double Rrt[9];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) Rrt[r * 3 + c] = Rr[c * 3 + r];
    matmul3(Rrt, Rw, Rb);
    for (int k = 0; k < 3; ++k) pb[k] = Rrt[k * 3] * (pw[0] - pr[0]) + Rrt[k * 3 + 1] * (pw[1] - pr[1]) + Rrt[k * 3 + 2] * (pw[2] - pr[2]);
    // rotation matrix -> quaternion (x, y, z, w)
    const double tr = Rb[0] + Rb[4] + Rb[8];
    double qx, qy, qz, qw;
    if (tr > 0.0) {
      const double s4 = std::sqrt(tr + 1.0) * 2.0;
      qw = 0.25 * s4; qx = (Rb[7] - Rb[5]) / s4; qy = (Rb[2] - Rb[6]) / s4; qz = (Rb[3] - Rb[1]) / s4;
    } else if (Rb[0] > Rb[4] && Rb[0] > Rb[8]) {
      const double s4 = std::sqrt(1.0 + Rb[0] - Rb[4] - Rb[8]) * 2.0;
      qw = (Rb[7] - Rb[5]) / s4; qx = 0.25 * s4; qy = (Rb[1] + Rb[3]) / s4; qz = (Rb[2] + Rb[6]) / s4;
    } else if (Rb[4] > Rb[8]) {
      const double s4 = std::sqrt(1.0 + Rb[4] - Rb[0] - Rb[8]) * 2.0;
      qw = (Rb[2] - Rb[6]) / s4; qx = (Rb[1] + Rb[3]) / s4; qy = 0.25 * s4; qz = (Rb[5] + Rb[7]) / s4;
    } else {
      const double s4 = std::sqrt(1.0 + Rb[8] - Rb[0] - Rb[4]) * 2.0;
      qw = (Rb[3] - Rb[1]) / s4; qx = (Rb[2] + Rb[6]) / s4; qy = (Rb[5] + Rb[7]) / s4; qz = 0.25 * s4;
    }
    const double quat[4] = {qx, qy, qz, qw};
    visit(lb, pb, quat);
  }
}
}  // namespace urdf_detail

// StompCollisionSpace::addAllBodiesButExcludeLinksToPoints' input (src/stomp_collision_space.cpp:564-586): the primitive
// collision bodies of the robot's links at a joint state, in the reference frame; meshBodiesAtState gives the mesh ones.
inline std::vector<stomp_body> bodiesAtState(const StompRobotModelUrdf& m, const std::vector<double>& group_values,
                                             const std::vector<std::string>& exclude_links, double scale = 1.0, double padding = 0.0) {
  std::vector<stomp_body> out;
  urdf_detail::linkBodyPoses(m, group_values, exclude_links, [&](const LinkBody& lb, const double* pb, const double* quat) {
    if (lb.type == kLinkBodyMesh) return;
    stomp_body b;
    std::memset(&b, 0, sizeof(b));
    b.type = lb.type;
    for (int k = 0; k < 3; ++k) { b.dimensions[k] = lb.dimensions[k]; b.position[k] = pb[k]; }
    for (int k = 0; k < 4; ++k) b.orientation[k] = quat[k];
    b.scale = scale;
    b.padding = padding;
    out.push_back(b);
  });
  return out;
}

// The mesh collision geometries (after loadLinkMeshes) as stomp_mesh_body records for stomp_engine_build_sdf_meshes.  The
// records point into the model's vertex arrays: keep the model alive and unchanged while they are in use.
inline std::vector<stomp_mesh_body> meshBodiesAtState(const StompRobotModelUrdf& m, const std::vector<double>& group_values,
                                                      const std::vector<std::string>& exclude_links, double scale = 1.0,
                                                      double padding = 0.0) {
  std::vector<stomp_mesh_body> out;
  urdf_detail::linkBodyPoses(m, group_values, exclude_links, [&](const LinkBody& lb, const double* pb, const double* quat) {
    if (lb.type != kLinkBodyMesh || lb.mesh_vertices.size() < 12) return;
    stomp_mesh_body b;
    std::memset(&b, 0, sizeof(b));
    b.vertices = lb.mesh_vertices.data();
    b.num_vertices = int32_t(lb.mesh_vertices.size() / 3);
    for (int k = 0; k < 3; ++k) b.position[k] = pb[k];
    for (int k = 0; k < 4; ++k) b.orientation[k] = quat[k];
    b.scale = scale;
    b.padding = padding;
    out.push_back(b);
  });
  return out;
}

}  // namespace stomp_motion_planner
