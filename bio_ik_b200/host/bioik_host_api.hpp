// bioik_host_api.hpp — header-only C++ host side above the C ABI (include/bioik_b200.h).
//
// Mirrors the reference's public goal / solver interface with the same class names and argument
// meaning (include/bio_ik/goal.h, include/bio_ik/goal_types.h, src/ik_base.h:128-210) without MoveIt:
// the kinematic tree is a flattened RobotModel.  Goals only DESCRIBE themselves
// into the flattened BioikGoal record; evaluation runs in the CUDA kernels behind the ABI.  Errors are
// reported the way the reference does (ERROR(...) -> std::runtime_error, src/utils.h:122-129).
// For the adapter that plugs this into the real bio_ik tree see INTEGRATION.md.
#pragma once

#include "../../include/bioik_b200.h"

#include <cfloat>
#include <cmath>
#include <cstdint>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace bio_ik
{

// tf2::Vector3 / tf2::Quaternion look-alikes (only what goal construction needs)
struct Vector3
{
    double x_, y_, z_;
    Vector3(double x = 0, double y = 0, double z = 0) : x_(x), y_(y), z_(z) {}
    double x() const { return x_; }
    double y() const { return y_; }
    double z() const { return z_; }
    Vector3 normalized() const
    {
        double s = 1.0 / std::sqrt(x_ * x_ + y_ * y_ + z_ * z_); // tf2: v * (1.0 / length())
        return Vector3(x_ * s, y_ * s, z_ * s);
    }
};
struct Quaternion
{
    double x_, y_, z_, w_;
    Quaternion(double x = 0, double y = 0, double z = 0, double w = 1) : x_(x), y_(y), z_(z), w_(w) {}
    Quaternion normalized() const
    {
        double s = 1.0 / std::sqrt(x_ * x_ + y_ * y_ + z_ * z_ + w_ * w_);
        return Quaternion(x_ * s, y_ * s, z_ * s, w_ * s);
    }
};

// ---- flattened moveit::core::RobotModel --------------------------------------------------------
struct RobotModel
{
    struct Link
    {
        std::string name, parent, joint_name;
        int joint_type = BIOIK_JOINT_FIXED;
        double origin[7] = {0, 0, 0, 0, 0, 0, 1}; // px py pz qx qy qz qw
        double axis[3] = {0, 0, 1};
        double lower = 0, upper = 0, velocity = 1;
        bool bounded = true;
        std::string mimic; // joint name
        double mimic_factor = 1, mimic_offset = 0;
        double mass = 0;           // URDF inertial (BalanceGoal, src/goal_types.cpp:236-250)
        double com[3] = {0, 0, 0}; // inertial origin in the link frame
    };
    std::vector<Link> links;
    std::vector<std::string> variable_names;
    std::vector<int32_t> link_parent, joint_type, joint_first_var, joint_mimic, var_bounded;
    std::vector<double> link_origin, joint_axis, joint_mimic_factor, joint_mimic_offset, var_min, var_max, var_max_velocity, link_mass, link_com;
    std::map<std::string, int> link_index, joint_index, variable_index;
    static int variableCount(int joint_type) { return joint_type == BIOIK_JOINT_REVOLUTE || joint_type == BIOIK_JOINT_PRISMATIC ? 1 : (joint_type == BIOIK_JOINT_FLOATING ? 7 : (joint_type == BIOIK_JOINT_PLANAR ? 3 : 0)); }

    void addLink(const Link& l) { links.push_back(l); }
    // call once after all links were added (parents before children)
    void finalize()
    {
        link_index.clear(), joint_index.clear(), variable_index.clear(), variable_names.clear();
        for(size_t i = 0; i < links.size(); i++) link_index[links[i].name] = (int)i, joint_index[links[i].joint_name] = (int)i;
        link_parent.clear(), joint_type.clear(), joint_first_var.clear(), joint_mimic.clear(), link_origin.clear(), joint_axis.clear();
        joint_mimic_factor.clear(), joint_mimic_offset.clear(), var_min.clear(), var_max.clear(), var_bounded.clear(), var_max_velocity.clear();
        link_mass.clear(), link_com.clear();
        for(auto& l : links)
        {
            if(!l.parent.empty() && !link_index.count(l.parent)) throw std::runtime_error("link not found " + l.parent);
            link_parent.push_back(l.parent.empty() ? -1 : link_index[l.parent]);
            joint_type.push_back(l.joint_type);
            const int cnt = variableCount(l.joint_type);
            joint_first_var.push_back(cnt ? (int)variable_names.size() : -1);
            auto addVariable = [&](const std::string& name, double lo, double hi, bool bounded) {
                variable_index[name] = (int)variable_names.size();
                variable_names.push_back(name);
                var_min.push_back(lo), var_max.push_back(hi), var_bounded.push_back(bounded), var_max_velocity.push_back(l.velocity);
            };
            if(cnt == 1) addVariable(l.joint_name, l.lower, l.upper, l.bounded);
            if(cnt == 7) // MoveIt FloatingJointModel: variable names and default bounds
            {
                const char* names[7] = {"trans_x", "trans_y", "trans_z", "rot_x", "rot_y", "rot_z", "rot_w"};
                for(int k = 0; k < 7; k++) addVariable(l.joint_name + "/" + names[k], k < 3 ? -1e308 : -1.0, k < 3 ? 1e308 : 1.0, k >= 3);
            }
            if(cnt == 3) // PlanarJointModel
            {
                const char* names[3] = {"x", "y", "theta"};
                for(int k = 0; k < 3; k++) addVariable(l.joint_name + "/" + names[k], k < 2 ? -1e308 : -3.14159265358979323846, k < 2 ? 1e308 : 3.14159265358979323846, false);
            }
            for(double o : l.origin) link_origin.push_back(o);
            for(double a : l.axis) joint_axis.push_back(a);
            joint_mimic.push_back(l.mimic.empty() ? -1 : joint_index.at(l.mimic));
            joint_mimic_factor.push_back(l.mimic_factor), joint_mimic_offset.push_back(l.mimic_offset);
            link_mass.push_back(l.mass);
            for(double c : l.com) link_com.push_back(c);
        }
    }
    size_t getVariableCount() const { return variable_names.size(); }
    BioikRobot toABI() const
    {
        BioikRobot r{};
        r.n_links = (int32_t)links.size(), r.n_vars = (int32_t)variable_names.size();
        r.link_mass = link_mass.data(), r.link_com = link_com.data();
        r.link_parent = link_parent.data(), r.joint_type = joint_type.data(), r.joint_first_var = joint_first_var.data();
        r.link_origin = link_origin.data(), r.joint_axis = joint_axis.data(), r.joint_mimic = joint_mimic.data();
        r.joint_mimic_factor = joint_mimic_factor.data(), r.joint_mimic_offset = joint_mimic_offset.data();
        r.var_min = var_min.data(), r.var_max = var_max.data(), r.var_bounded = var_bounded.data(), r.var_max_velocity = var_max_velocity.data();
        return r;
    }
};

struct JointModelGroup
{
    std::string name;
    std::vector<std::string> joint_names; // active joints of the group, in group order
    std::vector<std::string> tip_links;   // end-effector tips
};

// ---- goals (include/bio_ik/goal.h:97-119, goal_types.h) ---------------------------------------------
class Goal
{
protected:
    bool secondary_ = false;
    double weight_ = 1;

public:
    virtual ~Goal() {}
    bool isSecondary() const { return secondary_; }
    double getWeight() const { return weight_; }
    void setWeight(double w) { weight_ = w; }
    // flattened description: goal type, link name / variable name it refers to, parameter block
    virtual int type() const = 0;
    virtual std::string linkName() const { return ""; }
    // every link the goal refers to (GoalContext::addLink in describe()); the first one is its tip.  Default: linkName()
    virtual std::vector<std::string> linkNames(const struct RobotModel&) const
    {
        std::vector<std::string> r;
        if(!linkName().empty()) r.push_back(linkName());
        return r;
    }
    virtual std::string variableName() const { return ""; }
    virtual void params(double* p) const { (void)p; }
};

class LinkGoalBase : public Goal
{
    std::string link_name_;

public:
    LinkGoalBase() {}
    LinkGoalBase(const std::string& link_name, double weight) : link_name_(link_name) { weight_ = weight; }
    void setLinkName(const std::string& n) { link_name_ = n; }
    const std::string& getLinkName() const { return link_name_; }
    std::string linkName() const override { return link_name_; }
};

#define BIOIK_V3(p, o, v) (p)[(o)] = (v).x(), (p)[(o) + 1] = (v).y(), (p)[(o) + 2] = (v).z()

class PositionGoal : public LinkGoalBase
{
    Vector3 position_;

public:
    PositionGoal() {}
    PositionGoal(const std::string& link_name, const Vector3& position, double weight = 1.0) : LinkGoalBase(link_name, weight), position_(position) {}
    void setPosition(const Vector3& p) { position_ = p; }
    int type() const override { return BIOIK_GOAL_POSITION; }
    void params(double* p) const override { BIOIK_V3(p, 0, position_); }
};

class OrientationGoal : public LinkGoalBase
{
    Quaternion orientation_;

public:
    OrientationGoal() {}
    OrientationGoal(const std::string& link_name, const Quaternion& orientation, double weight = 1.0) : LinkGoalBase(link_name, weight), orientation_(orientation.normalized()) {}
    void setOrientation(const Quaternion& q) { orientation_ = q.normalized(); }
    int type() const override { return BIOIK_GOAL_ORIENTATION; }
    void params(double* p) const override { p[3] = orientation_.x_, p[4] = orientation_.y_, p[5] = orientation_.z_, p[6] = orientation_.w_; }
};

class PoseGoal : public LinkGoalBase
{
    Vector3 position_;
    Quaternion orientation_;
    double rotation_scale_ = 0.5;

public:
    PoseGoal() {}
    PoseGoal(const std::string& link_name, const Vector3& position, const Quaternion& orientation, double weight = 1.0)
        : LinkGoalBase(link_name, weight), position_(position), orientation_(orientation.normalized())
    {
    }
    void setPosition(const Vector3& p) { position_ = p; }
    void setOrientation(const Quaternion& q) { orientation_ = q.normalized(); }
    double getRotationScale() const { return rotation_scale_; }
    void setRotationScale(double s) { rotation_scale_ = s; }
    int type() const override { return BIOIK_GOAL_POSE; }
    void params(double* p) const override
    {
        BIOIK_V3(p, 0, position_);
        p[3] = orientation_.x_, p[4] = orientation_.y_, p[5] = orientation_.z_, p[6] = orientation_.w_, p[7] = rotation_scale_;
    }
};

class LookAtGoal : public LinkGoalBase
{
    Vector3 axis_{1, 0, 0}, target_;

public:
    LookAtGoal() {}
    LookAtGoal(const std::string& link_name, const Vector3& axis, const Vector3& target, double weight = 1.0) : LinkGoalBase(link_name, weight), axis_(axis), target_(target) {}
    void setAxis(const Vector3& a) { axis_ = a.normalized(); }
    void setTarget(const Vector3& t) { target_ = t; }
    int type() const override { return BIOIK_GOAL_LOOK_AT; }
    void params(double* p) const override { BIOIK_V3(p, 0, axis_), BIOIK_V3(p, 3, target_); }
};

class MaxDistanceGoal : public LinkGoalBase
{
protected:
    Vector3 target;
    double distance = 1;

public:
    MaxDistanceGoal() {}
    MaxDistanceGoal(const std::string& link_name, const Vector3& target, double distance, double weight = 1.0) : LinkGoalBase(link_name, weight), target(target), distance(distance) {}
    void setTarget(const Vector3& t) { target = t; }
    void setDistance(double d) { distance = d; }
    int type() const override { return BIOIK_GOAL_MAX_DISTANCE; }
    void params(double* p) const override { BIOIK_V3(p, 0, target), p[3] = distance; }
};
class MinDistanceGoal : public MaxDistanceGoal
{
public:
    using MaxDistanceGoal::MaxDistanceGoal;
    int type() const override { return BIOIK_GOAL_MIN_DISTANCE; }
};

class LineGoal : public LinkGoalBase
{
    Vector3 position, direction;

public:
    LineGoal() {}
    LineGoal(const std::string& link_name, const Vector3& position, const Vector3& direction, double weight = 1.0) : LinkGoalBase(link_name, weight), position(position), direction(direction.normalized()) {}
    void setPosition(const Vector3& p) { position = p; }
    void setDirection(const Vector3& d) { direction = d.normalized(); }
    int type() const override { return BIOIK_GOAL_LINE; }
    void params(double* p) const override { BIOIK_V3(p, 0, position), BIOIK_V3(p, 3, direction); }
};

class PlaneGoal : public LinkGoalBase
{
    Vector3 position, normal{0, 0, 1};

public:
    PlaneGoal() {}
    PlaneGoal(const std::string& link_name, const Vector3& position, const Vector3& normal, double weight = 1.0) : LinkGoalBase(link_name, weight), position(position), normal(normal.normalized()) {}
    void setPosition(const Vector3& p) { position = p; }
    void setNormal(const Vector3& n) { normal = n.normalized(); }
    int type() const override { return BIOIK_GOAL_PLANE; }
    void params(double* p) const override { BIOIK_V3(p, 0, position), BIOIK_V3(p, 3, normal); }
};

class SideGoal : public LinkGoalBase
{
protected:
    Vector3 axis{0, 0, 1}, direction{0, 0, 1};

public:
    SideGoal() {}
    SideGoal(const std::string& link_name, const Vector3& axis, const Vector3& direction, double weight = 1.0) : LinkGoalBase(link_name, weight), axis(axis), direction(direction) {}
    void setAxis(const Vector3& a) { axis = a.normalized(); }
    void setDirection(const Vector3& d) { direction = d.normalized(); }
    int type() const override { return BIOIK_GOAL_SIDE; }
    void params(double* p) const override { BIOIK_V3(p, 0, axis), BIOIK_V3(p, 3, direction); }
};
class DirectionGoal : public SideGoal
{
public:
    using SideGoal::SideGoal;
    int type() const override { return BIOIK_GOAL_DIRECTION; }
};

class ConeGoal : public LinkGoalBase // goal_types.h:646-712 (the three reference constructors)
{
    Vector3 position, axis{0, 0, 1}, direction{0, 0, 1};
    double position_weight = 0, angle = 0;

public:
    ConeGoal() {}
    ConeGoal(const std::string& link_name, const Vector3& axis, const Vector3& direction, double angle, double weight = 1.0) : LinkGoalBase(link_name, weight), axis(axis), direction(direction), angle(angle) {}
    ConeGoal(const std::string& link_name, const Vector3& position, const Vector3& axis, const Vector3& direction, double angle, double weight = 1.0)
        : LinkGoalBase(link_name, weight), position(position), axis(axis), direction(direction), position_weight(1), angle(angle)
    {
    }
    ConeGoal(const std::string& link_name, const Vector3& position, double position_weight, const Vector3& axis, const Vector3& direction, double angle, double weight = 1.0)
        : LinkGoalBase(link_name, weight), position(position), axis(axis), direction(direction), position_weight(position_weight), angle(angle)
    {
    }
    void setPosition(const Vector3& p) { position = p; }
    void setPositionWeight(double w) { position_weight = w; }
    void setAxis(const Vector3& a) { axis = a.normalized(); }
    void setDirection(const Vector3& d) { direction = d.normalized(); }
    void setAngle(double a) { angle = a; }
    int type() const override { return BIOIK_GOAL_CONE; }
    void params(double* p) const override { BIOIK_V3(p, 0, position), p[3] = position_weight, BIOIK_V3(p, 4, axis), BIOIK_V3(p, 7, direction), p[10] = angle; }
};

#define BIOIK_JOINT_SPACE_GOAL(NAME, TYPE, SECONDARY_DEFAULT)                                    \
    class NAME : public Goal                                                                      \
    {                                                                                             \
    public:                                                                                       \
        NAME(double weight = 1.0, bool secondary = SECONDARY_DEFAULT) { weight_ = weight, secondary_ = secondary; } \
        int type() const override { return TYPE; }                                                \
    };
BIOIK_JOINT_SPACE_GOAL(AvoidJointLimitsGoal, BIOIK_GOAL_AVOID_JOINT_LIMITS, true)
BIOIK_JOINT_SPACE_GOAL(CenterJointsGoal, BIOIK_GOAL_CENTER_JOINTS, true)
BIOIK_JOINT_SPACE_GOAL(MinimalDisplacementGoal, BIOIK_GOAL_MINIMAL_DISPLACEMENT, true)
BIOIK_JOINT_SPACE_GOAL(RegularizationGoal, BIOIK_GOAL_REGULARIZATION, false)

class JointVariableGoal : public Goal
{
    std::string variable_name;
    double variable_position = 0;

public:
    JointVariableGoal() {}
    JointVariableGoal(const std::string& variable_name, double variable_position, double weight = 1.0, bool secondary = false) : variable_name(variable_name), variable_position(variable_position)
    {
        weight_ = weight, secondary_ = secondary;
    }
    void setVariablePosition(double p) { variable_position = p; }
    int type() const override { return BIOIK_GOAL_JOINT_VARIABLE; }
    std::string variableName() const override { return variable_name; }
    void params(double* p) const override { p[0] = variable_position; }
};

// goal_types.h:540-568, src/goal_types.cpp:231-272: the centre of mass over every link with a positive inertial mass is pulled onto
// the line through `target` along `axis`; each of those links becomes a tip link of the problem (link order)
class BalanceGoal : public Goal
{
    Vector3 target_{0, 0, 0}, axis_{0, 0, 1};

public:
    BalanceGoal() {}
    BalanceGoal(const Vector3& target, double weight = 1.0) : target_(target) { weight_ = weight; }
    const Vector3& getTarget() const { return target_; }
    const Vector3& getAxis() const { return axis_; }
    void setTarget(const Vector3& t) { target_ = t; }
    void setAxis(const Vector3& a) { axis_ = a; }
    int type() const override { return BIOIK_GOAL_BALANCE; }
    std::vector<std::string> linkNames(const RobotModel& robot) const override
    {
        std::vector<std::string> r;
        for(auto& l : robot.links)
            if(l.mass > 0) r.push_back(l.name);
        return r;
    }
    void params(double* p) const override { BIOIK_V3(p, 0, target_), BIOIK_V3(p, 3, axis_); }
};

// ---- Problem (src/problem.cpp:72-228): tips, active variables, flattened goals ---------------------------
class Problem
{
public:
    std::vector<int32_t> tip_link_indices, active_variables;
    std::vector<BioikGoal> goals;
    double dpos = DBL_MAX, drot = DBL_MAX, dtwist = 1e-5;

    // fixed_joints: BioIKKinematicsQueryOptions::fixed_joints (joint names whose variables stay at the seed)
    void initialize(const RobotModel& robot, const JointModelGroup& group, const std::vector<const Goal*>& goal_list, const std::vector<std::string>& fixed_joints = {})
    {
        tip_link_indices.clear(), active_variables.clear(), goals.clear();
        std::vector<int> link_tip(robot.links.size(), -1);
        auto jointOfVariable = [&](int v) {
            for(size_t l = 0; l < robot.links.size(); l++)
                if(robot.joint_first_var[l] >= 0 && v >= robot.joint_first_var[l] && v < robot.joint_first_var[l] + RobotModel::variableCount(robot.links[l].joint_type)) return (int)l;
            return -1;
        };
        auto isFixed = [&](int link) {
            for(auto& f : fixed_joints)
                if(link >= 0 && f == robot.links[link].joint_name) return true;
            return false;
        };
        // src/problem.cpp:103-126: a fixed joint's variable stays out; otherwise the variable must belong to the joint group
        auto addActive = [&](const std::string& variable) {
            auto it = robot.variable_index.find(variable);
            if(it == robot.variable_index.end()) throw std::runtime_error("joint variable not found " + variable);
            const int joint = jointOfVariable(it->second);
            if(isFixed(joint)) return;
            for(int v : active_variables)
                if(v == it->second) return;
            bool in_group = false;
            for(auto& jn : group.joint_names) in_group = in_group || (joint >= 0 && jn == robot.links[joint].joint_name);
            if(!in_group) throw std::runtime_error("joint variable not found " + variable);
            active_variables.push_back(it->second);
        };
        for(const Goal* g : goal_list)
        {
            BioikGoal bg{};
            bg.type = g->type(), bg.secondary = g->isSecondary(), bg.weight = g->getWeight();
            bool first = true;
            for(auto& ln : g->linkNames(robot))
            {
                auto it = robot.link_index.find(ln);
                if(it == robot.link_index.end()) throw std::runtime_error("link not found " + ln);
                if(link_tip[it->second] < 0) link_tip[it->second] = (int)tip_link_indices.size(), tip_link_indices.push_back(it->second);
                if(first) bg.tip = link_tip[it->second];
                first = false;
            }
            std::string vn = g->variableName();
            if(!vn.empty()) addActive(vn), bg.var = robot.variable_index.at(vn);
            g->params(bg.p);
            goals.push_back(bg);
        }
        // active variables from the active subtree (:191-204): every variable of every used, non-mimic, non-fixed joint of the group
        std::vector<int> usage(robot.links.size(), 0);
        for(int tip : tip_link_indices)
            for(int l = tip; l >= 0; l = robot.link_parent[l]) usage[l] = 1;
        for(auto& jn : group.joint_names)
        {
            int l = robot.joint_index.at(jn);
            if(!usage[l] || !robot.links[l].mimic.empty() || isFixed(l)) continue;
            for(int k = 0; k < RobotModel::variableCount(robot.links[l].joint_type); k++) addActive(robot.variable_names[robot.joint_first_var[l] + k]);
        }
    }
    BioikProblem toABI() const
    {
        BioikProblem p{};
        p.n_tips = (int32_t)tip_link_indices.size(), p.tip_links = tip_link_indices.data();
        p.n_active = (int32_t)active_variables.size(), p.active_vars = active_variables.data();
        p.n_goals = (int32_t)goals.size(), p.goals = goals.data();
        p.dpos = dpos, p.drot = drot, p.dtwist = dtwist;
        return p;
    }
};

// ---- IKFactory-style solver over the ABI (src/utils.h:398-444, src/ik_base.h:128-210) --------------------
class IKSolverB200
{
    bioik_ctx* ctx_ = nullptr;
    size_t n_vars_ = 0;
    int queries_ = 0;

    void check(int rc, const char* what)
    {
        if(rc != BIOIK_OK) throw std::runtime_error(std::string(what) + ": " + bioik_last_error(ctx_));
    }

public:
    // name: "bio2", "bio2_memetic" or "bio2_memetic_l" (src/ik_evolution_2.cpp:652-654)
    IKSolverB200(const std::string& name, const RobotModel& robot, int population = 18, uint32_t random_seed = 1, int device = 0)
    {
        BioikSolverCfg cfg{};
        cfg.population = population, cfg.memetic_iters = 8, cfg.table_seed = random_seed, cfg.device = device;
        if(name == "bio2")
            cfg.memetic = 0, cfg.generations = 16;
        else if(name == "bio2_memetic")
            cfg.memetic = 'q', cfg.generations = 8;
        else if(name == "bio2_memetic_l")
            cfg.memetic = 'l', cfg.generations = 8;
        else
            throw std::runtime_error("class not found " + name);
        BioikRobot r = robot.toABI();
        n_vars_ = robot.getVariableCount();
        int rc = bioik_create(&r, &cfg, &ctx_);
        if(rc != BIOIK_OK) throw std::runtime_error(std::string("bioik_create: ") + bioik_last_error(nullptr));
    }
    ~IKSolverB200() { bioik_destroy(ctx_); }
    IKSolverB200(const IKSolverB200&) = delete;
    IKSolverB200& operator=(const IKSolverB200&) = delete;

    void initialize(const Problem& problem) // IKBase::initialize
    {
        BioikProblem p = problem.toABI();
        check(bioik_set_problem(ctx_, &p), "bioik_set_problem");
    }
    void cancel() { check(bioik_cancel(ctx_), "bioik_cancel"); } // IKBase::canceled; callable from another thread
    // bioik_set_option, e.g. BIOIK_OPT_REFERENCE_STALE_TIPS (quirk Q2 of the reference's memetic step on multi-tip problems)
    void setOption(int32_t option, int32_t value) { check(bioik_set_option(ctx_, option, value), "bioik_set_option"); }
    struct Result
    {
        std::vector<double> solutions, fitness;
        std::vector<int32_t> success, steps;
    };
    // B queries; goal_params [B][n_goals][BIOIK_GOAL_NPARAM] or empty, seeds [B][n_vars]
    Result solveBatch(const std::vector<double>& goal_params, const std::vector<double>& seeds, const std::vector<uint32_t>& rng_seeds, int steps, bool early_exit = false)
    {
        int B = (int)rng_seeds.size();
        Result r;
        r.solutions.resize((size_t)B * n_vars_), r.fitness.resize(B), r.success.resize(B), r.steps.resize(B);
        check(bioik_solve_batch(ctx_, B, goal_params.empty() ? nullptr : goal_params.data(), seeds.data(), rng_seeds.data(), steps, early_exit, r.solutions.data(), r.fitness.data(), r.success.data(), r.steps.data()),
              "bioik_solve_batch");
        return r;
    }
    struct IslandResult
    {
        std::vector<double> solutions, fitness;
        std::vector<int32_t> success, island, steps;
    };
    // Q MoveIt-style queries, `islands` differently seeded runs each, reduced like IKParallel::solve (src/ik_parallel.h:218-258)
    // and angle-wrapped like the plugin (src/kinematics_plugin.cpp:580-611).  rng_seeds: [Q * islands].
    // early_exit: 0 = every island runs `steps` steps, 1 = an island stops at its own success, 2 = the driver's `finished` flag among the islands
    IslandResult solveIslands(const std::vector<double>& goal_params, const std::vector<double>& seeds, int islands, const std::vector<uint32_t>& rng_seeds, int steps, int early_exit = 2, bool wrap = true)
    {
        int Q = (int)(rng_seeds.size() / (size_t)islands);
        IslandResult r;
        r.solutions.resize((size_t)Q * n_vars_), r.fitness.resize(Q), r.success.resize(Q), r.island.resize(Q), r.steps.resize(Q);
        check(bioik_solve_islands(ctx_, Q, islands, goal_params.empty() ? nullptr : goal_params.data(), seeds.data(), rng_seeds.data(), steps, early_exit, wrap, r.solutions.data(), r.fitness.data(), r.success.data(),
                                  r.island.data(), r.steps.data()),
              "bioik_solve_islands");
        return r;
    }
    // the reference's solver interface in its resumable form: IKBase::initialize / step / getSolution (src/ik_base.h:138-154) for
    // Q queries x `islands` runs; the state stays on the device between the calls
    void begin(const std::vector<double>& goal_params, const std::vector<double>& seeds, int islands, const std::vector<uint32_t>& rng_seeds, int max_steps = 0, int early_exit = 2)
    {
        queries_ = (int)(rng_seeds.size() / (size_t)islands);
        check(bioik_begin(ctx_, queries_, islands, goal_params.empty() ? nullptr : goal_params.data(), seeds.data(), rng_seeds.data(), max_steps, early_exit), "bioik_begin");
    }
    int step(int nsteps = 1) // returns the number of runs that would execute a further step
    {
        int32_t active = 0;
        check(bioik_step(ctx_, nsteps, &active), "bioik_step");
        return active;
    }
    IslandResult getSolution(bool wrap = false)
    {
        IslandResult r;
        r.solutions.resize((size_t)queries_ * n_vars_), r.fitness.resize(queries_), r.success.resize(queries_), r.island.resize(queries_), r.steps.resize(queries_);
        check(bioik_get_solution(ctx_, wrap, r.solutions.data(), r.fitness.data(), r.success.data(), r.island.data(), r.steps.data()), "bioik_get_solution");
        return r;
    }
    std::vector<double> forwardKinematics(const std::vector<double>& variables, int n_tips)
    {
        int B = (int)(variables.size() / n_vars_);
        std::vector<double> tips((size_t)B * n_tips * 7);
        check(bioik_fk_batch(ctx_, B, variables.data(), tips.data()), "bioik_fk_batch");
        return tips;
    }
};

} // namespace bio_ik
