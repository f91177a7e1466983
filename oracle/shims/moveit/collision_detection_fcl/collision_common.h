// shim: with FCL >= 0.6 the reference compiles TouchGoal and the collision members out (goal_types.h:330, problem.h)
#pragma once
#define FCL_VERSION_CHECK(major, minor, patch) ((major << 16) | (minor << 8) | (patch))
#define MOVEIT_FCL_VERSION FCL_VERSION_CHECK(0, 6, 0)
