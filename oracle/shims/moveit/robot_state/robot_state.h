// shim: only RobotFK_MoveIt (src/forward_kinematics.h:1468-1503, never used by bio2) touches RobotState
#pragma once
#include <moveit/robot_model/robot_model.h>
namespace moveit
{
namespace core
{
class RobotState
{
public:
    explicit RobotState(const RobotModelConstPtr&) {}
    void setToDefaultValues() {}
    void setVariablePositions(const std::vector<double>&) {}
    void setVariablePositions(const double*) {}
    void update() {}
    void updateLinkTransforms() {}
    const Eigen::Isometry3d& getGlobalLinkTransform(const std::string&) const { return t_; }
    const Eigen::Isometry3d& getGlobalLinkTransform(const LinkModel*) const { return t_; }
    void setJointPositions(const JointModel*, const double*) {}
    void setJointPositions(const std::string&, const double*) {}
    void updateLinkTransform(const LinkModel*) {}
    Eigen::Isometry3d t_;
};
}
}
namespace robot_state
{
typedef moveit::core::RobotState RobotState;
}
