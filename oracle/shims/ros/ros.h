// shim: nothing of ROS is used on the bio2 path (logging macros are the reference's own)
#pragma once
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <functional>
#include <iostream>
#include <map>
#include <random>
#include <set>
#include <sstream>
#include <vector>
#include <stdexcept>
#include <string>
namespace ros
{
struct WallTime
{
    double t = 0;
    static WallTime now() { return WallTime(); }
    double toSec() const { return t; }
};
}
