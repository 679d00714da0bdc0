// oracle/ref_shim/tf/transform_listener.h — tf::quaternionMsgToTF and the Bullet types tf re-exports.  TEST INFRASTRUCTURE.
#ifndef STOMP_REF_SHIM_TF
#define STOMP_REF_SHIM_TF
#include <LinearMath/bullet_shim.h>
#include <ros_msgs_shim.h>
namespace tf {
// identity "transform": the harness expresses attached-object poses directly in the owner link's frame
class TransformListener {
 public:
  void transformPose(const std::string&, const geometry_msgs::PoseStamped& in, geometry_msgs::PoseStamped& out) const { out = in; }
};
template <typename M> class MessageFilter {};
inline void poseTFToMsg(const btTransform& t, geometry_msgs::Pose& p) {
  p.position.x = t.getOrigin().x(); p.position.y = t.getOrigin().y(); p.position.z = t.getOrigin().z();
  p.orientation.x = t.getRotation().x(); p.orientation.y = t.getRotation().y();
  p.orientation.z = t.getRotation().z(); p.orientation.w = t.getRotation().w();
}
inline void poseMsgToTF(const geometry_msgs::Pose& p, btTransform& t) {
  t = btTransform(btQuaternion(p.orientation.x, p.orientation.y, p.orientation.z, p.orientation.w),
                  btVector3(p.position.x, p.position.y, p.position.z));
}
inline void quaternionMsgToTF(const geometry_msgs::Quaternion& m, btQuaternion& q) { q = btQuaternion(m.x, m.y, m.z, m.w); }
}
#endif
