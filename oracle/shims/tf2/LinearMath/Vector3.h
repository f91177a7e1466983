// shim: tf2::Vector3 (Bullet LinearMath semantics): 4 doubles, left-to-right arithmetic
#pragma once
#include <cmath>
typedef double tf2Scalar;
namespace tf2
{
inline tf2Scalar tf2Sqrt(tf2Scalar x) { return std::sqrt(x); }
inline tf2Scalar tf2Acos(tf2Scalar x)
{
    if(x < tf2Scalar(-1)) x = tf2Scalar(-1);
    if(x > tf2Scalar(1)) x = tf2Scalar(1);
    return std::acos(x);
}
// test hook (ref_harness.cpp): replaces the acos of Vector3::angle - ConeGoal's only transcendental (goal_types.h:706) -
// by the arithmetic contract's det_acos, so the reference's code can be compared with the GPU path bit for bit
inline double (*&vector3AngleAcosHook())(double)
{
    static double (*hook)(double) = nullptr;
    return hook;
}
class Vector3
{
public:
    tf2Scalar m_floats[4];
    Vector3() {}
    Vector3(const tf2Scalar& x, const tf2Scalar& y, const tf2Scalar& z) { m_floats[0] = x, m_floats[1] = y, m_floats[2] = z, m_floats[3] = tf2Scalar(0.); }
    const tf2Scalar& x() const { return m_floats[0]; }
    const tf2Scalar& y() const { return m_floats[1]; }
    const tf2Scalar& z() const { return m_floats[2]; }
    const tf2Scalar& getX() const { return m_floats[0]; }
    const tf2Scalar& getY() const { return m_floats[1]; }
    const tf2Scalar& getZ() const { return m_floats[2]; }
    void setX(tf2Scalar v) { m_floats[0] = v; }
    void setY(tf2Scalar v) { m_floats[1] = v; }
    void setZ(tf2Scalar v) { m_floats[2] = v; }
    Vector3& operator+=(const Vector3& v) { m_floats[0] += v.m_floats[0], m_floats[1] += v.m_floats[1], m_floats[2] += v.m_floats[2]; return *this; }
    Vector3& operator-=(const Vector3& v) { m_floats[0] -= v.m_floats[0], m_floats[1] -= v.m_floats[1], m_floats[2] -= v.m_floats[2]; return *this; }
    Vector3& operator*=(const tf2Scalar& s) { m_floats[0] *= s, m_floats[1] *= s, m_floats[2] *= s; return *this; }
    Vector3& operator/=(const tf2Scalar& s) { return *this *= tf2Scalar(1.0) / s; }
    tf2Scalar dot(const Vector3& v) const { return m_floats[0] * v.m_floats[0] + m_floats[1] * v.m_floats[1] + m_floats[2] * v.m_floats[2]; }
    tf2Scalar length2() const { return dot(*this); }
    tf2Scalar length() const { return tf2Sqrt(length2()); }
    tf2Scalar distance2(const Vector3& v) const;
    tf2Scalar distance(const Vector3& v) const;
    Vector3& normalize() { return *this /= length(); }
    Vector3 normalized() const;
    tf2Scalar angle(const Vector3& v) const
    {
        tf2Scalar s = tf2Sqrt(length2() * v.length2());
        if(auto hook = vector3AngleAcosHook())
        {
            tf2Scalar c = dot(v) / s;
            if(c < tf2Scalar(-1)) c = tf2Scalar(-1);
            if(c > tf2Scalar(1)) c = tf2Scalar(1);
            return hook(c);
        }
        return tf2Acos(dot(v) / s);
    }
    Vector3 cross(const Vector3& v) const
    {
        return Vector3(m_floats[1] * v.m_floats[2] - m_floats[2] * v.m_floats[1], m_floats[2] * v.m_floats[0] - m_floats[0] * v.m_floats[2], m_floats[0] * v.m_floats[1] - m_floats[1] * v.m_floats[0]);
    }
};
inline Vector3 operator+(const Vector3& a, const Vector3& b) { return Vector3(a.m_floats[0] + b.m_floats[0], a.m_floats[1] + b.m_floats[1], a.m_floats[2] + b.m_floats[2]); }
inline Vector3 operator-(const Vector3& a, const Vector3& b) { return Vector3(a.m_floats[0] - b.m_floats[0], a.m_floats[1] - b.m_floats[1], a.m_floats[2] - b.m_floats[2]); }
inline Vector3 operator-(const Vector3& v) { return Vector3(-v.m_floats[0], -v.m_floats[1], -v.m_floats[2]); }
inline Vector3 operator*(const Vector3& v, const tf2Scalar& s) { return Vector3(v.m_floats[0] * s, v.m_floats[1] * s, v.m_floats[2] * s); }
inline Vector3 operator*(const tf2Scalar& s, const Vector3& v) { return v * s; }
inline Vector3 operator/(const Vector3& v, const tf2Scalar& s) { return v * (tf2Scalar(1.0) / s); }
inline tf2Scalar Vector3::distance2(const Vector3& v) const { return (v - *this).length2(); }
inline tf2Scalar Vector3::distance(const Vector3& v) const { return (v - *this).length(); }
inline Vector3 Vector3::normalized() const { return *this / length(); }
}
