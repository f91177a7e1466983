// shim: the KDL frame operations used by the reference (include/bio_ik/frame.h:62-68,96-102,240-259; src/problem.cpp:278-322).
// Semantics of orocos-KDL frames.hpp / frames.cpp: Rotation::Quaternion, GetQuaternion, GetRot (axis * angle), diff, Equal.
#pragma once
#include <cmath>
namespace KDL
{
class Vector
{
public:
    double data[3];
    Vector() { data[0] = data[1] = data[2] = 0; }
    Vector(double x, double y, double z) { data[0] = x, data[1] = y, data[2] = z; }
    double x() const { return data[0]; }
    double y() const { return data[1]; }
    double z() const { return data[2]; }
    void x(double v) { data[0] = v; }
    void y(double v) { data[1] = v; }
    void z(double v) { data[2] = v; }
    static Vector Zero() { return Vector(0, 0, 0); }
};
inline Vector operator-(const Vector& a, const Vector& b) { return Vector(a.data[0] - b.data[0], a.data[1] - b.data[1], a.data[2] - b.data[2]); }
inline Vector operator*(const Vector& a, double s) { return Vector(a.data[0] * s, a.data[1] * s, a.data[2] * s); }
class Rotation
{
public:
    double data[9];
    Rotation() { data[0] = data[4] = data[8] = 1, data[1] = data[2] = data[3] = data[5] = data[6] = data[7] = 0; }
    static Rotation Quaternion(double x, double y, double z, double w)
    {
        Rotation R;
        double x2 = x * x, y2 = y * y, z2 = z * z, w2 = w * w;
        R.data[0] = w2 + x2 - y2 - z2, R.data[1] = 2 * x * y - 2 * w * z, R.data[2] = 2 * x * z + 2 * w * y;
        R.data[3] = 2 * x * y + 2 * w * z, R.data[4] = w2 - x2 + y2 - z2, R.data[5] = 2 * y * z - 2 * w * x;
        R.data[6] = 2 * x * z - 2 * w * y, R.data[7] = 2 * y * z + 2 * w * x, R.data[8] = w2 - x2 - y2 + z2;
        return R;
    }
    void GetQuaternion(double& x, double& y, double& z, double& w) const
    {
        double trace = data[0] + data[4] + data[8];
        const double epsilon = 1E-12;
        if(trace > epsilon)
        {
            double s = 0.5 / std::sqrt(trace + 1.0);
            w = 0.25 / s, x = (data[7] - data[5]) * s, y = (data[2] - data[6]) * s, z = (data[3] - data[1]) * s;
        }
        else if(data[0] > data[4] && data[0] > data[8])
        {
            double s = 2.0 * std::sqrt(1.0 + data[0] - data[4] - data[8]);
            w = (data[7] - data[5]) / s, x = 0.25 * s, y = (data[1] + data[3]) / s, z = (data[2] + data[6]) / s;
        }
        else if(data[4] > data[8])
        {
            double s = 2.0 * std::sqrt(1.0 + data[4] - data[0] - data[8]);
            w = (data[2] - data[6]) / s, x = (data[1] + data[3]) / s, y = 0.25 * s, z = (data[5] + data[7]) / s;
        }
        else
        {
            double s = 2.0 * std::sqrt(1.0 + data[8] - data[0] - data[4]);
            w = (data[3] - data[1]) / s, x = (data[2] + data[6]) / s, y = (data[5] + data[7]) / s, z = 0.25 * s;
        }
    }
    Rotation Inverse() const
    {
        Rotation R;
        for(int i = 0; i < 3; i++)
            for(int j = 0; j < 3; j++) R.data[i * 3 + j] = data[j * 3 + i];
        return R;
    }
    Vector operator*(const Vector& v) const
    {
        return Vector(data[0] * v.data[0] + data[1] * v.data[1] + data[2] * v.data[2], data[3] * v.data[0] + data[4] * v.data[1] + data[5] * v.data[2],
                      data[6] * v.data[0] + data[7] * v.data[1] + data[8] * v.data[2]);
    }
    // axis * angle (KDL::Rotation::GetRot)
    Vector GetRot() const
    {
        Vector axis(data[7] - data[5], data[2] - data[6], data[3] - data[1]);
        double sa = std::sqrt(axis.data[0] * axis.data[0] + axis.data[1] * axis.data[1] + axis.data[2] * axis.data[2]) / 2.0;
        double ca = (data[0] + data[4] + data[8] - 1) / 2.0;
        double angle = std::atan2(sa, ca);
        if(sa > 1e-12) return axis * (angle / (2.0 * sa));
        if(ca > 0) return axis * 0.5;
        return Vector(angle, 0, 0);
    }
};
inline Rotation operator*(const Rotation& a, const Rotation& b)
{
    Rotation R;
    for(int i = 0; i < 3; i++)
        for(int j = 0; j < 3; j++) R.data[i * 3 + j] = a.data[i * 3] * b.data[j] + a.data[i * 3 + 1] * b.data[3 + j] + a.data[i * 3 + 2] * b.data[6 + j];
    return R;
}
class Frame
{
public:
    Vector p;
    Rotation M;
};
class Twist
{
public:
    Vector vel, rot;
    Twist() {}
    Twist(const Vector& v, const Vector& r) : vel(v), rot(r) {}
    static Twist Zero() { return Twist(); }
};
inline Vector diff(const Vector& a, const Vector& b, double dt = 1) { return (b - a) * (1.0 / dt); }
inline Vector diff(const Rotation& R_a_b1, const Rotation& R_a_b2, double dt = 1)
{
    Rotation R_b1_b2 = R_a_b1.Inverse() * R_a_b2;
    return R_a_b1 * R_b1_b2.GetRot() * (1.0 / dt);
}
inline bool Equal(double a, double b, double eps) { double d = a - b; return (d < eps) && (d > -eps); }
inline bool Equal(const Vector& a, const Vector& b, double eps = 1e-6) { return Equal(a.data[0], b.data[0], eps) && Equal(a.data[1], b.data[1], eps) && Equal(a.data[2], b.data[2], eps); }
inline bool Equal(const Twist& a, const Twist& b, double eps = 1e-6) { return Equal(a.rot, b.rot, eps) && Equal(a.vel, b.vel, eps); }
}
