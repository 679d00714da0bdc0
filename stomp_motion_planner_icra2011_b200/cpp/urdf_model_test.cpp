// Host-only driver of stomp_robot_model_urdf.hpp (no GPU needed): tests/test_urdf_cpp_cpu.py builds it, feeds it a URDF and a
// small spec file and compares the dumped tables with stomp_motion_planner_icra2011_b200/urdf.py.
// usage: urdf_model_test robot.urdf spec.txt
//   spec lines:  group j1 j2 ... | reference link | clearance c | collision link radius extension | state joint value |
//                chain root tip | exclude link ... | start q1 q2 ... | padding scale pad |
//                attach link type d0 d1 d2 x y z padding | meshdir directory (package:// names are looked up there)
#include <algorithm>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <sstream>
#include <stomp_motion_planner/stomp_robot_model_urdf.hpp>
using namespace stomp_motion_planner;

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  std::ifstream uf(argv[1]);
  std::stringstream ub;
  ub << uf.rdbuf();
  std::vector<std::string> group, exclude;
  std::string reference, chain_root, chain_tip, meshdir;
  std::vector<CollisionLinkConfig> col;
  std::map<std::string, double> state;
  std::vector<double> start;
  double clearance = 0.07, scale = 1.0, padding = 0.0;
  struct Attach { std::string link; int type; double dims[3], pos[3], padding; };
  std::vector<Attach> attach;
  std::ifstream sf(argv[2]);
  std::string line;
  while (std::getline(sf, line)) {
    std::istringstream is(line);
    std::string key, w;
    is >> key;
    if (key == "group") while (is >> w) group.push_back(w);
    else if (key == "reference") is >> reference;
    else if (key == "clearance") is >> clearance;
    else if (key == "collision") { CollisionLinkConfig c; is >> c.link >> c.link_radius >> c.link_extension; col.push_back(c); }
    else if (key == "state") { double v; is >> w >> v; state[w] = v; }
    else if (key == "chain") is >> chain_root >> chain_tip;
    else if (key == "exclude") while (is >> w) exclude.push_back(w);
    else if (key == "start") { double v; while (is >> v) start.push_back(v); }
    else if (key == "padding") is >> scale >> padding;
    else if (key == "meshdir") is >> meshdir;
    else if (key == "attach") { Attach at; is >> at.link >> at.type >> at.dims[0] >> at.dims[1] >> at.dims[2] >> at.pos[0] >> at.pos[1] >> at.pos[2] >> at.padding; attach.push_back(at); }
  }
  StompRobotModelUrdf m;
  std::string err;
  if (!loadRobotModelFromUrdf(ub.str(), group, reference, col, clearance, state, chain_root, chain_tip, m, err)) {
    std::printf("error %s\n", err.c_str());
    return 1;
  }
  for (const Attach& at : attach)
    if (!addAttachedObjectCollisionPoint(m, at.link, at.type, at.dims, at.pos, at.padding, clearance)) { std::printf("error attach %s\n", at.link.c_str()); return 1; }
  std::printf("reference %d chain %d %d\n", m.reference_segment, m.chain_root_segment, m.chain_tip_segment);
  for (size_t s = 0; s < m.segments.size(); ++s) {
    const stomp_segment& g = m.segments[s];
    std::printf("segment %s %d %d %d", m.segment_names[s].c_str(), g.parent, g.joint_type, g.group_index);
    for (double v : g.rot) std::printf(" %.17g", v);
    for (double v : g.pos) std::printf(" %.17g", v);
    for (double v : g.axis) std::printf(" %.17g", v);
    std::printf(" %.17g\n", g.fixed_value);
  }
  for (const stomp_joint_limit& l : m.joint_limits) std::printf("limit %d %.17g %.17g\n", l.has_limits, l.min, l.max);
  for (const stomp_sphere& sp : m.collision_points)
    std::printf("sphere %d %.17g %.17g %.17g %.17g %.17g\n", sp.segment, sp.radius, sp.clearance, sp.pos[0], sp.pos[1], sp.pos[2]);
  for (size_t s = 0; s < m.link_inertias.size(); ++s) {
    const stomp_link_inertia& li = m.link_inertias[s];
    if (li.mass == 0.0) continue;
    std::printf("inertia %zu %.17g %.17g %.17g %.17g", s, li.mass, li.com[0], li.com[1], li.com[2]);
    for (double v : li.inertia) std::printf(" %.17g", v);
    std::printf("\n");
  }
  if (!meshdir.empty()) {
    auto resolve = [&](const std::string& name) {
      const std::string tag = "package://";
      return meshdir + "/" + (name.compare(0, tag.size(), tag) == 0 ? name.substr(tag.size()) : name);
    };
    if (!loadLinkMeshes(m, resolve, err)) { std::printf("error %s\n", err.c_str()); return 1; }
    for (const stomp_mesh_body& b : meshBodiesAtState(m, start, exclude, scale, padding)) {
      std::printf("mesh %d", b.num_vertices);
      for (double v : b.position) std::printf(" %.17g", v);
      for (double v : b.orientation) std::printf(" %.17g", v);
      std::printf(" %.17g %.17g", b.scale, b.padding);
      double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
      for (int v = 0; v < b.num_vertices; ++v)
        for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], b.vertices[v * 3 + k]); hi[k] = std::max(hi[k], b.vertices[v * 3 + k]); }
      std::printf(" %.17g %.17g %.17g %.17g %.17g %.17g\n", lo[0], lo[1], lo[2], hi[0], hi[1], hi[2]);
    }
  }
  for (const stomp_body& b : bodiesAtState(m, start, exclude, scale, padding)) {
    std::printf("body %d %.17g %.17g %.17g", b.type, b.dimensions[0], b.dimensions[1], b.dimensions[2]);
    for (double v : b.position) std::printf(" %.17g", v);
    for (double v : b.orientation) std::printf(" %.17g", v);
    std::printf(" %.17g %.17g\n", b.scale, b.padding);
  }
  return 0;
}
