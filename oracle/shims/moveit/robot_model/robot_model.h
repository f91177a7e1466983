// shim: moveit::core::RobotModel / LinkModel / JointModel / JointModelGroup accessors the bio2 path uses
// (SURVEY.md Appendix B), over a flattened robot table.  Built by oracle/ref_harness.cpp from a BioikRobot.
#pragma once
#include <cmath>
#include <Eigen/Dense>
#include <memory>
#include <string>
#include <vector>
// urdf::ModelInterface::getLink(name)->inertial: what BalanceGoal::describe reads (src/goal_types.cpp:236-250)
namespace urdf
{
struct Vector3
{
    double x = 0, y = 0, z = 0;
};
struct Pose
{
    Vector3 position;
};
struct Inertial
{
    Pose origin;
    double mass = 0;
};
struct Link
{
    std::shared_ptr<Inertial> inertial;
};
struct ModelInterface
{
    std::vector<std::pair<std::string, std::shared_ptr<Link>>> links_;
    std::shared_ptr<const Link> getLink(const std::string& name) const
    {
        for(auto& l : links_)
            if(l.first == name) return l.second;
        return nullptr;
    }
};
}
namespace moveit
{
namespace core
{
class LinkModel;
class JointModel;
struct VariableBounds
{
    double min_position_ = 0, max_position_ = 0;
    bool position_bounded_ = false;
    double max_velocity_ = 0;
    bool velocity_bounded_ = true;
};
class JointModel
{
public:
    enum JointType { UNKNOWN, REVOLUTE, PRISMATIC, PLANAR, FLOATING, FIXED };
    std::string name_;
    JointType type_ = FIXED;
    int joint_index_ = 0, first_variable_index_ = 0;
    size_t variable_count_ = 0;
    std::vector<std::string> variable_names_;
    const JointModel* mimic_ = nullptr;
    double mimic_factor_ = 1, mimic_offset_ = 0;
    const LinkModel *child_link_ = nullptr, *parent_link_ = nullptr;
    Eigen::Vector3d axis_;
    virtual ~JointModel() {}
    JointType getType() const { return type_; }
    const std::string& getName() const { return name_; }
    size_t getJointIndex() const { return joint_index_; }
    size_t getFirstVariableIndex() const { return first_variable_index_; }
    size_t getVariableCount() const { return variable_count_; }
    const std::vector<std::string>& getVariableNames() const { return variable_names_; }
    const JointModel* getMimic() const { return mimic_; }
    double getMimicFactor() const { return mimic_factor_; }
    double getMimicOffset() const { return mimic_offset_; }
    const LinkModel* getChildLinkModel() const { return child_link_; }
    const LinkModel* getParentLinkModel() const { return parent_link_; }
    // Reached by the reference for PLANAR joints only (forward_kinematics.h:128-135).  MoveIt's PlanarJointModel::computeTransform is
    //   transf = Eigen::Isometry3d(Eigen::Translation3d(v[0], v[1], 0.0) * Eigen::AngleAxisd(v[2], Eigen::Vector3d::UnitZ()));
    // restated here with Eigen's AngleAxis::toRotationMatrix arithmetic for the axis (0, 0, 1): third-party behaviour, see README.md.
    // sincos hook: libm by default, the arithmetic contract's det_sincos when ref_set_contract_math(1) is active.
    static void (*&planarSinCosHook())(double, double*, double*)
    {
        static void (*hook)(double, double*, double*) = nullptr;
        return hook;
    }
    void computeTransform(const double* v, Eigen::Isometry3d& t) const
    {
        t = Eigen::Isometry3d();
        if(type_ != PLANAR) return;
        double s, c;
        if(auto hook = planarSinCosHook())
            hook(v[2], &s, &c);
        else
            s = std::sin(v[2]), c = std::cos(v[2]);
        const double one_minus_c = 1.0 - c; // cos1_axis = (1 - c) * axis
        t.R(0, 0) = 0.0 * 0.0 + c, t.R(0, 1) = 0.0 - s, t.R(0, 2) = 0.0 + 0.0;
        t.R(1, 0) = 0.0 + s, t.R(1, 1) = 0.0 * 0.0 + c, t.R(1, 2) = 0.0 - 0.0;
        t.R(2, 0) = 0.0 - 0.0, t.R(2, 1) = 0.0 + 0.0, t.R(2, 2) = one_minus_c * 1.0 + c;
        t.t = Eigen::Vector3d(v[0], v[1], 0.0);
    }
};
class RevoluteJointModel : public JointModel
{
public:
    const Eigen::Vector3d& getAxis() const { return axis_; }
};
class PrismaticJointModel : public JointModel
{
public:
    const Eigen::Vector3d& getAxis() const { return axis_; }
};
class FixedJointModel : public JointModel
{
};
class LinkModel
{
public:
    std::string name_;
    int link_index_ = 0;
    const JointModel* parent_joint_ = nullptr;
    const LinkModel* parent_link_ = nullptr;
    Eigen::Isometry3d joint_origin_transform_;
    const std::string& getName() const { return name_; }
    size_t getLinkIndex() const { return link_index_; }
    const JointModel* getParentJointModel() const { return parent_joint_; }
    const LinkModel* getParentLinkModel() const { return parent_link_; }
    const Eigen::Isometry3d& getJointOriginTransform() const { return joint_origin_transform_; }
};
class JointModelGroup;
class RobotModel
{
public:
    std::vector<std::unique_ptr<LinkModel>> links_;
    std::vector<std::unique_ptr<JointModel>> joints_;
    std::vector<const LinkModel*> link_ptrs_;
    std::vector<const JointModel*> joint_ptrs_, mimic_joints_;
    std::vector<std::string> link_names_, variable_names_;
    std::vector<VariableBounds> bounds_;
    std::vector<const JointModel*> joint_of_variable_;
    std::vector<std::unique_ptr<JointModelGroup>> groups_;
    std::shared_ptr<urdf::ModelInterface> urdf_ = std::make_shared<urdf::ModelInterface>();
    const std::shared_ptr<urdf::ModelInterface>& getURDF() const { return urdf_; }

    const std::vector<const LinkModel*>& getLinkModels() const { return link_ptrs_; }
    size_t getLinkModelCount() const { return link_ptrs_.size(); }
    const std::vector<std::string>& getLinkModelNames() const { return link_names_; }
    const LinkModel* getLinkModel(const std::string& name) const
    {
        for(auto* l : link_ptrs_)
            if(l->getName() == name) return l;
        return nullptr;
    }
    const LinkModel* getLinkModel(size_t i) const { return link_ptrs_[i]; }
    size_t getJointModelCount() const { return joint_ptrs_.size(); }
    const JointModel* getJointModel(size_t i) const { return joint_ptrs_[i]; }
    const JointModel* getJointModel(const std::string& name) const
    {
        for(auto* j : joint_ptrs_)
            if(j->getName() == name) return j;
        return nullptr;
    }
    const std::vector<const JointModel*>& getMimicJointModels() const { return mimic_joints_; }
    size_t getVariableCount() const { return variable_names_.size(); }
    const std::vector<std::string>& getVariableNames() const { return variable_names_; }
    size_t getVariableIndex(const std::string& name) const
    {
        for(size_t i = 0; i < variable_names_.size(); i++)
            if(variable_names_[i] == name) return i;
        return (size_t)-1;
    }
    const VariableBounds& getVariableBounds(const std::string& name) const { return bounds_[getVariableIndex(name)]; }
    const JointModel* getJointOfVariable(size_t i) const { return joint_of_variable_[i]; }
    const JointModel* getJointOfVariable(const std::string& name) const { return joint_of_variable_[getVariableIndex(name)]; }
    void interpolate(const double*, const double*, double, double*) const {}
};
typedef std::shared_ptr<const RobotModel> RobotModelConstPtr;
typedef std::shared_ptr<RobotModel> RobotModelPtr;
}
}
#include <moveit/robot_model/joint_model_group.h>
