// shim: nothing of ROS is used on the bio2 path (logging macros are the reference's own)
#pragma once
#include <algorithm>
#include <chrono>
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
struct WallTime // wall-clock seconds; only IKParallel's timeout loop reads it (src/ik_parallel.h:162,168)
{
    double t = 0;
    static WallTime now()
    {
        WallTime w;
        w.t = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
        return w;
    }
    double toSec() const { return t; }
};
}
