// shim: base struct of BioIKKinematicsQueryOptions (include/bio_ik/goal.h:121)
#pragma once
#include <moveit/robot_model/robot_model.h>
namespace kinematics
{
struct KinematicsQueryOptions
{
    bool lock_redundant_joints = false;
    bool return_approximate_solution = false;
};
}
