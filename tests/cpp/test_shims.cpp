// Known-answer probe of oracle/shims (TEST INFRASTRUCTURE): the stand-in third-party headers restate Eigen / KDL / tf2 arithmetic that
// the reference's sources call.  tests/test_shims.py feeds rotations through them and compares with scipy.spatial.transform.Rotation
// and with the closed forms of SURVEY.md Appendix C, so a shim that disagreed with the real library would not go unnoticed.
//   stdin:  N, then N lines "qx qy qz qw  px py pz qx2 qy2 qz2 qw2"
//   stdout: per line: Eigen quaternion of the matrix of q (4) | KDL GetQuaternion of the same matrix (4) | KDL::diff(R(q), R(q2)) (3) |
//           tf2 quatRotate-free checks: tf2::Quaternion product q * q2 (4), angleShortestPath(q, q2) (1), Vector3(p).angle((1,2,3)) (1)
#include <Eigen/Dense>
#include <kdl/frames.hpp>
#include <tf2/LinearMath/Quaternion.h>
#include <tf2/LinearMath/Vector3.h>

#include <cstdio>

int main()
{
    int n = 0;
    if(std::scanf("%d", &n) != 1) return 1;
    for(int i = 0; i < n; i++)
    {
        double q[4], p[3], r[4];
        if(std::scanf("%lf %lf %lf %lf %lf %lf %lf %lf %lf %lf %lf", &q[0], &q[1], &q[2], &q[3], &p[0], &p[1], &p[2], &r[0], &r[1], &r[2], &r[3]) != 11) return 1;
        Eigen::Quaterniond eq(q[3], q[0], q[1], q[2]);
        Eigen::Matrix3d R = eq.toRotationMatrix();
        Eigen::Quaterniond back(R);
        KDL::Rotation K = KDL::Rotation::Quaternion(q[0], q[1], q[2], q[3]), K2 = KDL::Rotation::Quaternion(r[0], r[1], r[2], r[3]);
        double kx, ky, kz, kw;
        K.GetQuaternion(kx, ky, kz, kw);
        KDL::Vector d = KDL::diff(K, K2);
        tf2::Quaternion a(q[0], q[1], q[2], q[3]), b(r[0], r[1], r[2], r[3]);
        tf2::Quaternion ab = a * b;
        tf2::Vector3 v(p[0], p[1], p[2]);
        std::printf("%.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", back.x(), back.y(), back.z(), back.w(), kx, ky, kz, kw, d.x(), d.y(), d.z(), ab.x(), ab.y(),
                    ab.z(), ab.w(), a.angleShortestPath(b), v.angle(tf2::Vector3(1, 2, 3)));
    }
    return 0;
}
