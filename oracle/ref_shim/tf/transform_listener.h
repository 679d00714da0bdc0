// oracle/ref_shim/tf/transform_listener.h — tf::quaternionMsgToTF and the Bullet types tf re-exports.  TEST INFRASTRUCTURE.
#ifndef STOMP_REF_SHIM_TF
#define STOMP_REF_SHIM_TF
#include <LinearMath/bullet_shim.h>
#include <ros_msgs_shim.h>
namespace tf {
class TransformListener {};
template <typename M> class MessageFilter {};
inline void quaternionMsgToTF(const geometry_msgs::Quaternion& m, btQuaternion& q) { q = btQuaternion(m.x, m.y, m.z, m.w); }
}
#endif
