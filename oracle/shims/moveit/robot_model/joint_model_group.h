#pragma once
#include <moveit/robot_model/robot_model.h>
namespace moveit
{
namespace core
{
class JointModelGroup
{
public:
    const RobotModel* parent_ = nullptr;
    std::string name_;
    std::vector<std::string> variable_names_;
    std::vector<const JointModel*> active_joints_;
    std::vector<std::string> tips_;
    const std::vector<std::string>& getVariableNames() const { return variable_names_; }
    const std::vector<const JointModel*>& getActiveJointModels() const { return active_joints_; }
    const RobotModel& getParentModel() const { return *parent_; }
    bool getEndEffectorTips(std::vector<std::string>& tips) const
    {
        tips = tips_;
        return true;
    }
};
}
}
