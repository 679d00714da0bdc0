// oracle/ref_shim/kdl/tree.hpp — stand-in for KDL::Joint / Segment / Tree / TreeElement (orocos KDL 1.0 API, as built by
// kdl_parser).  TEST INFRASTRUCTURE; third-party semantics restated:
//   Joint::pose(q)   = Frame(Rot2(axis, q), origin)           (RotAxis)   | Frame(origin + q * axis)  (TransAxis)
//   Segment(f_tip)   stores joint.pose(0)^-1 * f_tip;  Segment::pose(q) = joint.pose(q) * f_tip_stored
//   Tree::addSegment numbers joints (q_nr) in insertion order; a fixed joint gets q_nr 0; children keep insertion order.
#ifndef STOMP_REF_SHIM_KDL_TREE
#define STOMP_REF_SHIM_KDL_TREE
#include <map>
#include <string>
#include <vector>
#include <kdl/frames.hpp>
#include <kdl/jntarray.hpp>
#include <kdl/chain.hpp>

namespace KDL {

class Tree;

class TreeElement;
typedef std::map<std::string, TreeElement> SegmentMap;

class TreeElement {
 public:
  Segment segment;
  unsigned int q_nr;
  SegmentMap::const_iterator parent;
  std::vector<SegmentMap::const_iterator> children;
  TreeElement() : q_nr(0) {}
  TreeElement(const Segment& s, const SegmentMap::const_iterator& p, unsigned int q) : segment(s), q_nr(q), parent(p) {}
  static TreeElement Root(const std::string& name) { TreeElement e; e.segment = Segment(name, Joint(name + "_joint", Joint::None)); return e; }
};

class Tree {
 public:
  explicit Tree(const std::string& root_name = "root") : nrOfJoints(0), nrOfSegments(0), root_name_(root_name) {
    root_ = segments.insert(std::make_pair(root_name, TreeElement::Root(root_name))).first;
  }
  Tree(const Tree& o) { *this = o; }
  Tree& operator=(const Tree& o) {  // iterators must point into the new map: rebuild by re-adding in DFS order
    segments.clear();
    nrOfJoints = nrOfSegments = 0;
    root_name_ = o.root_name_;
    root_ = segments.insert(std::make_pair(root_name_, TreeElement::Root(root_name_))).first;
    copyChildren(o, o.getRootSegment());
    return *this;
  }
  bool addSegment(const Segment& segment, const std::string& hook_name) {
    SegmentMap::iterator parent = segments.find(hook_name);
    if (parent == segments.end()) return false;
    unsigned int q_nr = segment.getJoint().getType() != Joint::None ? nrOfJoints : 0;
    std::pair<SegmentMap::iterator, bool> retval =
        segments.insert(std::make_pair(segment.getName(), TreeElement(segment, parent, q_nr)));
    if (!retval.second) return false;
    parent->second.children.push_back(retval.first);
    nrOfSegments++;
    if (segment.getJoint().getType() != Joint::None) nrOfJoints++;
    return true;
  }
  unsigned int getNrOfJoints() const { return nrOfJoints; }
  unsigned int getNrOfSegments() const { return nrOfSegments; }
  SegmentMap::const_iterator getSegment(const std::string& name) const { return segments.find(name); }
  SegmentMap::const_iterator getRootSegment() const { return root_; }
  const SegmentMap& getSegments() const { return segments; }
  bool getChain(const std::string& root, const std::string& tip, Chain& chain) const {
    chain = Chain();
    std::vector<SegmentMap::const_iterator> path;
    SegmentMap::const_iterator it = segments.find(tip);
    if (it == segments.end() || segments.find(root) == segments.end()) return false;
    while (it->first != root) {
      if (it == root_) return false;
      path.push_back(it);
      it = it->second.parent;
    }
    for (size_t i = path.size(); i-- > 0;) chain.addSegment(path[i]->second.segment);
    return true;
  }

 private:
  void copyChildren(const Tree& o, SegmentMap::const_iterator node) {
    for (size_t i = 0; i < node->second.children.size(); ++i) {
      SegmentMap::const_iterator c = node->second.children[i];
      addSegment(c->second.segment, node->first);
      copyChildren(o, c);
    }
  }
  SegmentMap segments;
  unsigned int nrOfJoints, nrOfSegments;
  std::string root_name_;
  SegmentMap::const_iterator root_;
};

}  // namespace KDL
#endif
