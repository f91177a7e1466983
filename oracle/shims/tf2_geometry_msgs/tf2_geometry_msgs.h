// shim: geometry_msgs::Pose and tf2::fromMsg as used by Frame(const geometry_msgs::Pose&) (include/bio_ik/frame.h:69-73)
#pragma once
#include <tf2/LinearMath/Quaternion.h>
namespace geometry_msgs
{
struct Point { double x = 0, y = 0, z = 0; };
struct QuaternionMsg { double x = 0, y = 0, z = 0, w = 1; };
struct Pose
{
    Point position;
    QuaternionMsg orientation;
};
}
namespace tf2
{
inline void fromMsg(const geometry_msgs::QuaternionMsg& in, Quaternion& out) { out = Quaternion(in.x, in.y, in.z, in.w); }
}
