// shim: tf2::Quaternion (Bullet LinearMath semantics)
#pragma once
#include "Vector3.h"
namespace tf2
{
class Quaternion
{
public:
    tf2Scalar m_floats[4];
    Quaternion() {}
    Quaternion(const tf2Scalar& x, const tf2Scalar& y, const tf2Scalar& z, const tf2Scalar& w) { m_floats[0] = x, m_floats[1] = y, m_floats[2] = z, m_floats[3] = w; }
    Quaternion(const Vector3& axis, const tf2Scalar& angle) { setRotation(axis, angle); }
    void setRotation(const Vector3& axis, const tf2Scalar& angle)
    {
        tf2Scalar d = axis.length();
        tf2Scalar s = std::sin(angle * tf2Scalar(0.5)) / d;
        m_floats[0] = axis.x() * s, m_floats[1] = axis.y() * s, m_floats[2] = axis.z() * s, m_floats[3] = std::cos(angle * tf2Scalar(0.5));
    }
    const tf2Scalar& x() const { return m_floats[0]; }
    const tf2Scalar& y() const { return m_floats[1]; }
    const tf2Scalar& z() const { return m_floats[2]; }
    const tf2Scalar& w() const { return m_floats[3]; }
    const tf2Scalar& getX() const { return m_floats[0]; }
    const tf2Scalar& getY() const { return m_floats[1]; }
    const tf2Scalar& getZ() const { return m_floats[2]; }
    const tf2Scalar& getW() const { return m_floats[3]; }
    void setX(tf2Scalar v) { m_floats[0] = v; }
    void setY(tf2Scalar v) { m_floats[1] = v; }
    void setZ(tf2Scalar v) { m_floats[2] = v; }
    void setW(tf2Scalar v) { m_floats[3] = v; }
    Quaternion& operator+=(const Quaternion& q) { m_floats[0] += q.m_floats[0], m_floats[1] += q.m_floats[1], m_floats[2] += q.m_floats[2], m_floats[3] += q.m_floats[3]; return *this; }
    Quaternion& operator-=(const Quaternion& q) { m_floats[0] -= q.m_floats[0], m_floats[1] -= q.m_floats[1], m_floats[2] -= q.m_floats[2], m_floats[3] -= q.m_floats[3]; return *this; }
    Quaternion& operator*=(const tf2Scalar& s) { m_floats[0] *= s, m_floats[1] *= s, m_floats[2] *= s, m_floats[3] *= s; return *this; }
    Quaternion& operator/=(const tf2Scalar& s) { return *this *= tf2Scalar(1.0) / s; }
    tf2Scalar dot(const Quaternion& q) const { return m_floats[0] * q.x() + m_floats[1] * q.y() + m_floats[2] * q.z() + m_floats[3] * q.m_floats[3]; }
    tf2Scalar length2() const { return dot(*this); }
    tf2Scalar length() const { return tf2Sqrt(length2()); }
    Quaternion& normalize() { return *this /= length(); }
    Quaternion operator*(const tf2Scalar& s) const { return Quaternion(x() * s, y() * s, z() * s, m_floats[3] * s); }
    Quaternion operator/(const tf2Scalar& s) const { return *this * (tf2Scalar(1.0) / s); }
    Quaternion normalized() const { return *this / length(); }
    Quaternion operator+(const Quaternion& q2) const { return Quaternion(m_floats[0] + q2.x(), m_floats[1] + q2.y(), m_floats[2] + q2.z(), m_floats[3] + q2.m_floats[3]); }
    Quaternion operator-(const Quaternion& q2) const { return Quaternion(m_floats[0] - q2.x(), m_floats[1] - q2.y(), m_floats[2] - q2.z(), m_floats[3] - q2.m_floats[3]); }
    Quaternion operator-() const { return Quaternion(-m_floats[0], -m_floats[1], -m_floats[2], -m_floats[3]); }
    Quaternion inverse() const { return Quaternion(-m_floats[0], -m_floats[1], -m_floats[2], m_floats[3]); }
    tf2Scalar angleShortestPath(const Quaternion& q) const
    {
        tf2Scalar s = tf2Sqrt(length2() * q.length2());
        if(dot(q) < 0)
            return tf2Acos(dot(-q) / s) * tf2Scalar(2.0);
        else
            return tf2Acos(dot(q) / s) * tf2Scalar(2.0);
    }
    tf2Scalar getAngle() const
    {
        if(auto hook = vector3AngleAcosHook()) // same test hook as Vector3::angle: the arithmetic contract's acos (frameTwist of the numeric Jacobian)
        {
            tf2Scalar c = m_floats[3];
            if(c < tf2Scalar(-1)) c = tf2Scalar(-1);
            if(c > tf2Scalar(1)) c = tf2Scalar(1);
            return tf2Scalar(2.) * hook(c);
        }
        return tf2Scalar(2.) * tf2Acos(m_floats[3]);
    }
    Vector3 getAxis() const
    {
        tf2Scalar s_squared = tf2Scalar(1.) - m_floats[3] * m_floats[3];
        if(s_squared < tf2Scalar(10.) * 2.2204460492503131e-16) return Vector3(1.0, 0.0, 0.0);
        tf2Scalar s = tf2Sqrt(s_squared);
        return Vector3(m_floats[0] / s, m_floats[1] / s, m_floats[2] / s);
    }
};
inline Quaternion operator*(const Quaternion& q1, const Quaternion& q2)
{
    return Quaternion(q1.w() * q2.x() + q1.x() * q2.w() + q1.y() * q2.z() - q1.z() * q2.y(), q1.w() * q2.y() + q1.y() * q2.w() + q1.z() * q2.x() - q1.x() * q2.z(),
                      q1.w() * q2.z() + q1.z() * q2.w() + q1.x() * q2.y() - q1.y() * q2.x(), q1.w() * q2.w() - q1.x() * q2.x() - q1.y() * q2.y() - q1.z() * q2.z());
}
}
