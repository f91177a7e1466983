// bioik_oracle.hpp — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
//
// A CPU restatement of the reference's bio2 / bio2_memetic hot path
// (TAMS-Group/bio_ik @ 1de2678), each function citing the reference file:line it
// follows.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may build, load or call anything under oracle/.  The
// product library (libbioik_b200.so) never links or calls this code.
//
// PARITY PINNED against the reference's own code.  The reference's build system cannot run in this image (catkin,
// ROS, MoveIt, tf2, Eigen, KDL, FCL, Boost are absent), but the two translation units of this path
// (src/ik_evolution_2.cpp, src/problem.cpp, with every bio_ik header they include) compile where they lie against
// the stand-in third-party headers of oracle/shims/ (oracle/Makefile target `ref` -> oracle/_ref/).
// tests/test_reference_pin.py compares this restatement with that build BIT FOR BIT - lookup tables, exact FK,
// delta frames, approximated frames, every device goal class, Problem::initialize's ordering, and whole 25-step
// solves (genes, gradients, species fitness, solutions, success) in all three bio2 modes, on all five BASELINE
// configurations, at population 4..128, with mimic / prismatic joints and random trees.  Two switches are flipped
// for that comparison, both documented deviations of the contract below: libm sin/cos (Options::libm_sincos) and
// the stale-tip quirk Q2 (Options::stale_tips).  The reference's own unit tests pin only the concat / invert /
// change algebra (test/utest.cpp:63-81, restated in tests/test_oracle.py).
// What the pin does NOT cover: the shims restate third-party behaviour (tf2 LinearMath, KDL frames, Eigen's
// matrix->quaternion), so a difference between a shim and the real library would go unnoticed; ConeGoal's acos is
// compared to rounding only (det_acos vs libm).
//
// Arithmetic contract (shared with the CUDA kernels so results are bit-identical,
// far inside the 1e-5 tolerance of BASELINE.json):
//   * IEEE-754 binary64, round-to-nearest, NO implicit FMA contraction
//     (build: -ffp-contract=off); the reference itself is built -ffast-math
//     (CMakeLists.txt:85-88) so no bit-exact CPU truth exists.
//   * explicit fused multiply-add exactly where the reference's AVX+FMA
//     approximator uses _mm256_fmadd_pd (src/forward_kinematics.h:949-950,
//     1091-1092) and inside det_sincos below.
//   * sin/cos of the half joint angle (src/forward_kinematics.h:103-104) use
//     det_sincos(): a fixed sequence of IEEE operations (Cody–Waite reduction with
//     explicit FMA + fdlibm-style minimax polynomials, <= 2 ulp from libm) so CPU
//     and GPU agree bit-for-bit.  Define BIOIK_ORACLE_LIBM_SINCOS at run time via
//     Options::libm_sincos to use libm instead (tests quantify the difference).
//   * Q2 of SURVEY.md §8(a): computeApproximateMutation1 leaves tips a variable
//     does not influence untouched (stale memory, src/forward_kinematics.h:1016).
//     The oracle implements the evidently intended semantics out[t] = in[t];
//     identical for single-tip problems.
#pragma once

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>

namespace bioik_oracle
{

// ---------------------------------------------------------------------------
// tf2 LinearMath look-alikes and include/bio_ik/frame.h
// ---------------------------------------------------------------------------
struct Vec3
{
    double x = 0, y = 0, z = 0;
    Vec3() {}
    Vec3(double x, double y, double z) : x(x), y(y), z(z) {}
};
struct Quat
{
    double x = 0, y = 0, z = 0, w = 1;
    Quat() {}
    Quat(double x, double y, double z, double w) : x(x), y(y), z(z), w(w) {}
};
// include/bio_ik/frame.h:51-55 (pad omitted; it carries no information)
struct Frame
{
    Vec3 pos;
    Quat rot;
    Frame() {}
    Frame(const Vec3& p, const Quat& q) : pos(p), rot(q) {}
    static Frame identity() { return Frame(Vec3(0, 0, 0), Quat(0, 0, 0, 1)); }
};

inline Vec3 operator+(const Vec3& a, const Vec3& b) { return Vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline Vec3 operator-(const Vec3& a, const Vec3& b) { return Vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline Vec3 operator-(const Vec3& a) { return Vec3(-a.x, -a.y, -a.z); }
inline Vec3 operator*(const Vec3& a, double s) { return Vec3(a.x * s, a.y * s, a.z * s); }
inline double dot(const Vec3& a, const Vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline double length2(const Vec3& a) { return dot(a, a); }
inline double length(const Vec3& a) { return std::sqrt(length2(a)); }
// tf2::Vector3::distance2(v) = (v - *this).length2()
inline double distance2(const Vec3& self, const Vec3& v) { return length2(v - self); }
inline double distance(const Vec3& self, const Vec3& v) { return length(v - self); }
inline Vec3 cross(const Vec3& a, const Vec3& b) { return Vec3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
// tf2: operator/(v, s) = v * (1.0 / s); normalized() = *this / length()
inline Vec3 normalized(const Vec3& a) { return a * (1.0 / length(a)); }

inline Quat operator+(const Quat& a, const Quat& b) { return Quat(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
inline Quat operator-(const Quat& a, const Quat& b) { return Quat(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
inline double dot(const Quat& a, const Quat& b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
inline double length2(const Quat& a) { return dot(a, a); }
inline Quat inverse(const Quat& q) { return Quat(-q.x, -q.y, -q.z, q.w); }
inline Quat normalized(const Quat& q)
{
    double s = 1.0 / std::sqrt(length2(q));
    return Quat(q.x * s, q.y * s, q.z * s, q.w * s);
}
// tf2::operator*(Quaternion, Quaternion) — Bullet's Hamilton product, evaluated left to right
inline Quat tf2_mul(const Quat& q1, const Quat& q2)
{
    return Quat(q1.w * q2.x + q1.x * q2.w + q1.y * q2.z - q1.z * q2.y, //
                q1.w * q2.y + q1.y * q2.w + q1.z * q2.x - q1.x * q2.z, //
                q1.w * q2.z + q1.z * q2.w + q1.x * q2.y - q1.y * q2.x, //
                q1.w * q2.w - q1.x * q2.x - q1.y * q2.y - q1.z * q2.z);
}

// include/bio_ik/frame.h:108-149
inline void quat_mul_vec(const Quat& q, const Vec3& v, Vec3& r)
{
    double v_x = v.x, v_y = v.y, v_z = v.z;
    double q_x = q.x, q_y = q.y, q_z = q.z, q_w = q.w;
    if((v_x == 0 && v_y == 0 && v_z == 0) || (q_x == 0 && q_y == 0 && q_z == 0 && q_w == 1))
    {
        r = v;
        return;
    }
    double t_x = q_y * v_z - q_z * v_y;
    double t_y = q_z * v_x - q_x * v_z;
    double t_z = q_x * v_y - q_y * v_x;
    double r_x = q_w * t_x + q_y * t_z - q_z * t_y;
    double r_y = q_w * t_y + q_z * t_x - q_x * t_z;
    double r_z = q_w * t_z + q_x * t_y - q_y * t_x;
    r_x += r_x;
    r_y += r_y;
    r_z += r_z;
    r_x += v_x;
    r_y += v_y;
    r_z += v_z;
    r = Vec3(r_x, r_y, r_z);
}
// include/bio_ik/frame.h:151-172
inline void quat_mul_quat(const Quat& p, const Quat& q, Quat& r)
{
    double p_x = p.x, p_y = p.y, p_z = p.z, p_w = p.w;
    double q_x = q.x, q_y = q.y, q_z = q.z, q_w = q.w;
    double r_x = (p_w * q_x + p_x * q_w) + (p_y * q_z - p_z * q_y);
    double r_y = (p_w * q_y - p_x * q_z) + (p_y * q_w + p_z * q_x);
    double r_z = (p_w * q_z + p_x * q_y) - (p_y * q_x - p_z * q_w);
    double r_w = (p_w * q_w - p_x * q_x) - (p_y * q_y + p_z * q_z);
    r = Quat(r_x, r_y, r_z, r_w);
}
// include/bio_ik/frame.h:174-187
inline void concat(const Frame& a, const Frame& b, Frame& r)
{
    Vec3 d;
    quat_mul_vec(a.rot, b.pos, d);
    Vec3 p = a.pos + d;
    Quat q;
    quat_mul_quat(a.rot, b.rot, q);
    r.pos = p;
    r.rot = q;
}
inline void concat(const Frame& a, const Frame& b, const Frame& c, Frame& r)
{
    Frame tmp;
    concat(a, b, tmp);
    concat(tmp, c, r);
}
// include/bio_ik/frame.h:189-209
inline void invert(const Frame& a, Frame& r)
{
    Quat qi = inverse(a.rot);
    Vec3 p;
    quat_mul_vec(qi, -a.pos, p);
    r.rot = qi;
    r.pos = p;
}
inline void change(const Frame& a, const Frame& b, const Frame& c, Frame& r)
{
    Frame tmp;
    invert(b, tmp);
    concat(a, tmp, c, r);
}
// include/bio_ik/frame.h:231-238
inline void normalizeFast(Quat& q)
{
    double f = (3.0 - length2(q)) * 0.5;
    q = Quat(q.x * f, q.y * f, q.z * f, q.w * f);
}

// include/bio_ik/frame.h:240-259 (frameTwist) with tf2's Quaternion::getAngle / getAxis (Appendix C of SURVEY.md);
// acos through `acos_fn` (libm in the reference, det_acos under the arithmetic contract).  vel / rot: the KDL::Twist members.
inline void frameTwist(const Frame& a, const Frame& b, double (*acos_fn)(double), Vec3& vel, Vec3& rot)
{
    Frame ia, frame;
    invert(a, ia);
    concat(ia, b, frame); // inverse(a) * b
    vel = frame.pos;
    double w = frame.rot.w; // getAngle(): 2 * tf2Acos(w), tf2Acos clamps to [-1, 1]
    if(w < -1.0) w = -1.0;
    if(w > 1.0) w = 1.0;
    double ra = 2.0 * acos_fn(w);
    if(ra > +M_PI) ra -= 2 * M_PI;
    // getAxis()
    double s_squared = 1.0 - frame.rot.w * frame.rot.w;
    Vec3 axis(1.0, 0.0, 0.0);
    if(!(s_squared < 10.0 * 2.2204460492503131e-16))
    {
        double sq = std::sqrt(s_squared);
        axis = Vec3(frame.rot.x / sq, frame.rot.y / sq, frame.rot.z / sq);
    }
    rot = axis * ra;
}

// src/utils.h:319-333
inline double mix(double a, double b, double f) { return a * (1.0 - f) + b * f; }
inline double clamp(double v, double lo, double hi)
{
    if(v < lo) v = lo;
    if(v > hi) v = hi;
    return v;
}

// ---------------------------------------------------------------------------
// det_sincos: the arithmetic-contract replacement for cos()/sin() at
// src/forward_kinematics.h:103-104.  Constants: pi/2 split in three doubles,
// fdlibm __kernel_sin/__kernel_cos minimax coefficients.
// ---------------------------------------------------------------------------
inline void det_sincos(double x, double* s_out, double* c_out)
{
    if(!(std::fabs(x) <= 1.0e5)) x = std::fmod(x, 6.283185307179586); // exact remainder; NaN/inf -> NaN
    double fn = std::rint(x * 0.6366197723675814);
    double r = std::fma(fn, -1.5707963267948966, x);
    r = std::fma(fn, -6.123233995736766e-17, r);
    r = std::fma(fn, 1.4973849048591698e-33, r);
    double z = r * r;
    double ps = std::fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
    ps = std::fma(z, ps, 2.75573137070700676789e-06);
    ps = std::fma(z, ps, -1.98412698298579493134e-04);
    ps = std::fma(z, ps, 8.33333333332248946124e-03);
    ps = std::fma(z, ps, -1.66666666666666324348e-01);
    double sr = std::fma(r * z, ps, r);
    double pc = std::fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
    pc = std::fma(z, pc, -2.75573143513906633035e-07);
    pc = std::fma(z, pc, 2.48015872894767294178e-05);
    pc = std::fma(z, pc, -1.38888888888741095749e-03);
    pc = std::fma(z, pc, 4.16666666666666019037e-02);
    double cr = std::fma(z * z, pc, std::fma(z, -0.5, 1.0));
    long long q = (long long)fn;
    double s = (q & 1) ? cr : sr;
    double c = (q & 1) ? sr : cr;
    if(q & 2) s = -s;
    if((q + 1) & 2) c = -c;
    *s_out = s;
    *c_out = c;
}

// det_acos: the arithmetic-contract acos for ConeGoal (tf2::Vector3::angle -> tf2Acos -> acos): the fdlibm
// e_acos.c algorithm (rational approximation + sqrt), plain IEEE operations only, so CPU and GPU agree bit-for-bit.
inline double det_acos(double x)
{
    const double one = 1.0, pi = 3.14159265358979311600e+00, pio2_hi = 1.57079632679489655800e+00, pio2_lo = 6.12323399573676603587e-17;
    const double pS0 = 1.66666666666666657415e-01, pS1 = -3.25565818622400915405e-01, pS2 = 2.01212532134862925881e-01, pS3 = -4.00555345006794114027e-02, pS4 = 7.91534994289814532176e-04,
                 pS5 = 3.47933107596021167570e-05;
    const double qS1 = -2.40339491173441421878e+00, qS2 = 2.02094576023350569471e+00, qS3 = -6.88283971605453293030e-01, qS4 = 7.70381505559019352791e-02;
    const double ax = std::fabs(x);
    if(!(ax < 1.0))
    {
        if(x == 1.0) return 0.0;
        if(x == -1.0) return pi + 2.0 * pio2_lo;
        return (x - x) / (x - x); // NaN
    }
    if(ax < 0.5)
    {
        if(ax < 6.938893903907228e-18) return pio2_hi + pio2_lo; // 2^-57
        double z = x * x;
        double p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
        double q = one + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
        double r = p / q;
        return pio2_hi - (x - (pio2_lo - x * r));
    }
    if(x < 0)
    {
        double z = (one + x) * 0.5;
        double p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
        double q = one + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
        double s = std::sqrt(z);
        double r = p / q;
        double w = r * s - pio2_lo;
        return pi - 2.0 * (s + w);
    }
    double z = (one - x) * 0.5;
    double s = std::sqrt(z);
    uint64_t bits;
    std::memcpy(&bits, &s, 8);
    bits &= 0xFFFFFFFF00000000ull; // df = s with the low word cleared
    double df;
    std::memcpy(&df, &bits, 8);
    double c = (z - df * df) / (s + df);
    double p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
    double q = one + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
    double r = p / q;
    double w = r * s + c;
    return 2.0 * (df + w);
}

struct Options
{
    bool libm_sincos = false; // use libm sin/cos instead of det_sincos (deviation study only)
    bool fma_approx = true;   // FMA in the approximator like the reference's AVX path; false = scalar path (:1174-1233)
    bool fma_approx1 = true;  // same switch for computeApproximateMutation1 alone (experiments)
    int island_stride = 0;    // BIOIK_OPT_ISLAND_STREAM_STRIDE of the product (not a reference feature): island i of solveIslands starts
                              // i * island_stride solver steps into the query-independent random streams
    bool stale_tips = false;  // emulate quirk Q2 of the reference (forward_kinematics.h:940): tips a variable does not move keep
                              // whatever the output buffer held before (pinning study only; never set on the parity path)
};

// ---------------------------------------------------------------------------
// src/utils.h:369-385
// ---------------------------------------------------------------------------
struct XORShift64
{
    uint64_t v = 88172645463325252ull;
    inline uint64_t operator()()
    {
        v ^= v << 13;
        v ^= v >> 7;
        v ^= v << 17;
        return v;
    }
};

// ---------------------------------------------------------------------------
// src/ik_base.h:49-126.  The two lookup buffers are static in the reference
// (shared by every solver); here they live in a Tables object shared by all
// queries of a batch (SURVEY.md §8(c) batch RNG contract).
// ---------------------------------------------------------------------------
static const size_t random_buffer_size = 1024 * 1024 * 8; // src/ik_base.h:74

struct Tables
{
    std::vector<double> uniform, gauss;
    // Random::Random(seed): rng(seed); make_random_buffer(); make_random_gauss_buffer()  (src/ik_base.h:118-125)
    explicit Tables(uint32_t seed, size_t size = random_buffer_size)
    {
        std::minstd_rand rng(seed);
        std::normal_distribution<double> normal_distribution;
        uniform.resize(size);
        for(auto& r : uniform) r = std::uniform_real_distribution<double>(0, 1)(rng); // :57,:80-81
        gauss.resize(size);
        for(auto& r : gauss) r = normal_distribution(rng); // :62,:98-99
    }
};

struct Random
{
    std::minstd_rand rng;
    XORShift64 _xorshift;
    const double* random_buffer;
    size_t random_buffer_index;
    const double* random_gauss_buffer;
    size_t random_gauss_index;

    // "a freshly constructed solver sharing the static tables", rng = minstd_rand(seed_q)
    Random(const Tables& t, uint32_t seed) : rng(seed)
    {
        random_buffer = t.uniform.data();
        random_buffer_index = _xorshift(); // src/ik_base.h:122
        random_gauss_buffer = t.gauss.data();
        random_gauss_index = _xorshift(); // src/ik_base.h:124
    }
    inline double random() { return std::uniform_real_distribution<double>(0, 1)(rng); }                      // :57
    inline size_t random_index(size_t s) { return std::uniform_int_distribution<size_t>(0, s - 1)(rng); }       // :59
    inline double random(double min, double max) { return random() * (max - min) + min; }                      // :64
    inline size_t fast_random_index(size_t mod) { return _xorshift() % mod; }                                  // :71
    inline double fast_random()                                                                                 // :86-91
    {
        double r = random_buffer[random_buffer_index & (random_buffer_size - 1)];
        random_buffer_index++;
        return r;
    }
    inline const double* fast_random_gauss_n(size_t n) // :110-116
    {
        size_t i = random_gauss_index;
        random_gauss_index += n;
        if(random_gauss_index >= random_buffer_size) i = 0, random_gauss_index = n;
        return random_gauss_buffer + i;
    }
};

// ---------------------------------------------------------------------------
// Flattened moveit::core::RobotModel
// ---------------------------------------------------------------------------
enum JointType
{
    FIXED = 0,
    REVOLUTE = 1,
    PRISMATIC = 2,
    FLOATING = 3,
    PLANAR = 4
};

struct RobotModel
{
    struct Link
    {
        int parent = -1;
        int joint_type = FIXED;
        int first_var = -1;
        Frame origin; // LinkModel::getJointOriginTransform()
        Vec3 axis;
        int mimic = -1; // child link of the mimicked joint
        double mimic_factor = 1, mimic_offset = 0;
    };
    std::vector<Link> links;
    size_t n_vars = 0;
    std::vector<double> var_min, var_max, var_max_velocity;
    std::vector<int> var_bounded;
    std::vector<int> var_joint; // getJointOfVariable -> child link index of the joint
    std::vector<double> link_mass; // URDF inertial mass per link (empty: none), link_com [3 * links]: inertial origin (BalanceGoal)
    std::vector<double> link_com;

    static int variableCount(int joint_type)
    {
        switch(joint_type)
        {
        case REVOLUTE:
        case PRISMATIC: return 1;
        case FLOATING: return 7;
        case PLANAR: return 3;
        default: return 0;
        }
    }
    void finalize()
    {
        var_joint.assign(n_vars, -1);
        for(size_t l = 0; l < links.size(); l++)
        {
            int cnt = variableCount(links[l].joint_type);
            for(int k = 0; k < cnt; k++)
                var_joint.at(links[l].first_var + k) = (int)l;
            if(links[l].parent >= (int)l) throw std::runtime_error("links must be ordered parents first");
        }
    }
};

// include/bio_ik/robot_info.h:70-113
struct RobotInfo
{
    struct VariableInfo
    {
        double clip_min, clip_max, span, min, max, max_velocity, max_velocity_rcp;
    };
    std::vector<VariableInfo> variables;
    RobotInfo() {}
    explicit RobotInfo(const RobotModel& model)
    {
        for(size_t ivar = 0; ivar < model.n_vars; ivar++)
        {
            VariableInfo info;
            bool bounded = model.var_bounded[ivar] != 0;
            int j = model.var_joint[ivar];
            if(j >= 0 && model.links[j].joint_type == REVOLUTE)
                if(model.var_max[ivar] - model.var_min[ivar] >= 2 * M_PI * 0.9999) bounded = false;
            info.min = model.var_min[ivar];
            info.max = model.var_max[ivar];
            info.clip_min = bounded ? info.min : -DBL_MAX;
            info.clip_max = bounded ? info.max : +DBL_MAX;
            info.span = info.max - info.min;
            if(!(info.span >= 0 && info.span < FLT_MAX)) info.span = 1;
            info.max_velocity = model.var_max_velocity[ivar];
            info.max_velocity_rcp = info.max_velocity > 0.0 ? 1.0 / info.max_velocity : 0.0;
            variables.push_back(info);
        }
    }
    inline double clip(double p, size_t i) const // :109-113 (clamp2 == clamp semantically)
    {
        auto& info = variables[i];
        if(p < info.clip_min) p = info.clip_min;
        if(p > info.clip_max) p = info.clip_max;
        return p;
    }
    inline double getSpan(size_t i) const { return variables[i].span; }
    inline double getClipMin(size_t i) const { return variables[i].clip_min; }
    inline double getClipMax(size_t i) const { return variables[i].clip_max; }
    inline double getMin(size_t i) const { return variables[i].min; }
    inline double getMax(size_t i) const { return variables[i].max; }
    inline double getMaxVelocityRcp(size_t i) const { return variables[i].max_velocity_rcp; }
};

// ---------------------------------------------------------------------------
// src/forward_kinematics.h: RobotJointEvaluator (:65-214), RobotFK_Fast_Base
// (:217-360), RobotFK_Jacobian (:553-731), RobotFK_Mutator (:783-1234).
// The per-variable joint-frame cache (:145-189) is a pure memoisation and is
// omitted.  RobotFK_Fast's incremental FK (:363-550) is not used by bio2.
// ---------------------------------------------------------------------------
class RobotFK
{
public:
    const RobotModel* robot_model = nullptr;
    Options opt;
    std::vector<double> variables;
    std::vector<Frame> tip_frames;
    std::vector<int> tip_links;
    std::vector<int> link_schedule;
    std::vector<Frame> global_frames;
    std::vector<std::vector<int>> joint_dependencies;
    std::vector<int> tip_dependencies;
    // Mutator state
    std::vector<double> mutation_approx_jacobian; // [tip*6+row][icol] row-major here (storage order is irrelevant)
    size_t jac_cols = 0;
    std::vector<std::vector<Frame>> mutation_approx_frames; // [tip][ivar]
    std::vector<size_t> mutation_approx_variable_indices;
    std::vector<std::vector<int>> mutation_approx_mask;
    std::vector<std::vector<size_t>> mutation_approx_map;
    std::vector<Frame> tip_frames_aligned;

    RobotFK() {}
    explicit RobotFK(const RobotModel* model, Options o = Options()) : robot_model(model), opt(o) {}

    // :78-139
    void getJointFrame(int link, const double* vars, Frame& frame) const
    {
        const auto& L = robot_model->links[link];
        switch(L.joint_type)
        {
        case FIXED: frame = Frame::identity(); return;
        case REVOLUTE:
        {
            double v = vars[L.first_var];
            double half_angle = v * 0.5;
            double fcos, fsin;
            if(opt.libm_sincos)
            {
                fcos = std::cos(half_angle);
                fsin = std::sin(half_angle);
            }
            else
                det_sincos(half_angle, &fsin, &fcos);
            frame = Frame(Vec3(0.0, 0.0, 0.0), Quat(L.axis.x * fsin, L.axis.y * fsin, L.axis.z * fsin, fcos));
            return;
        }
        case PRISMATIC:
        {
            double v = vars[L.first_var];
            frame = Frame(L.axis * v, Quat(0.0, 0.0, 0.0, 1.0));
            return;
        }
        case FLOATING:
        {
            const double* vv = vars + L.first_var;
            frame.pos = Vec3(vv[0], vv[1], vv[2]);
            frame.rot = normalized(Quat(vv[3], vv[4], vv[5], vv[6]));
            return;
        }
        case PLANAR:
        {
            // :128-135: joint_model->computeTransform + Frame(Isometry3d).  MoveIt's PlanarJointModel::computeTransform is
            // Translation3d(x, y, 0) * AngleAxisd(theta, UnitZ()); Eigen's AngleAxis::toRotationMatrix for that axis gives
            // [[c, -s, 0], [s, c, 0], [0, 0, (1 - c) + c]], and Frame(Isometry3d) (include/bio_ik/frame.h:74-79) converts with
            // Eigen's matrix -> quaternion (Shoemake).  Third-party arithmetic restated; sin / cos as for revolute joints.
            const double* vv = vars + L.first_var;
            double s, c;
            if(opt.libm_sincos)
                s = std::sin(vv[2]), c = std::cos(vv[2]);
            else
                det_sincos(vv[2], &s, &c);
            const double m00 = 0.0 * 0.0 + c, m01 = 0.0 - s, m10 = 0.0 + s, m11 = 0.0 * 0.0 + c, m22 = (1.0 - c) * 1.0 + c;
            frame.pos = Vec3(vv[0], vv[1], 0.0);
            double t = m00 + m11 + m22;
            if(t > 0.0)
            {
                t = std::sqrt(t + 1.0);
                double w = 0.5 * t;
                t = 0.5 / t;
                frame.rot = Quat((0.0 - 0.0) * t, (0.0 - 0.0) * t, (m10 - m01) * t, w);
            }
            else
            {
                // largest diagonal element: m11 > m00 never (equal), m22 > m00 whenever the trace is not positive
                if(m22 > m00)
                {
                    t = std::sqrt(m22 - m00 - m11 + 1.0);
                    double z = 0.5 * t;
                    t = 0.5 / t;
                    frame.rot = Quat((0.0 + 0.0) * t, (0.0 + 0.0) * t, z, (m10 - m01) * t);
                }
                else
                {
                    t = std::sqrt(m00 - m11 - m22 + 1.0);
                    double x = 0.5 * t;
                    t = 0.5 / t;
                    frame.rot = Quat(x, (m10 + m01) * t, (0.0 + 0.0) * t, (0.0 - 0.0) * t);
                }
            }
            return;
        }
        default: throw std::runtime_error("oracle: unknown joint type");
        }
    }

    // :230-246
    void updateMimic(std::vector<double>& values) const
    {
        for(size_t l = 0; l < robot_model->links.size(); l++)
        {
            const auto& L = robot_model->links[l];
            if(L.mimic < 0) continue;
            int src = robot_model->links[L.mimic].first_var;
            int dest = L.first_var;
            values[dest] = values[src] * L.mimic_factor + L.mimic_offset;
        }
    }

    // RobotFK_Fast_Base::initialize :253-330 + RobotFK_Jacobian::initialize :566-599
    void initialize(const std::vector<size_t>& tip_link_indices)
    {
        tip_links.clear();
        for(auto t : tip_link_indices) tip_links.push_back((int)t);
        tip_frames.resize(tip_links.size());
        global_frames.resize(robot_model->links.size());
        link_schedule.clear();
        for(int tip_link : tip_links)
        {
            std::vector<int> chain;
            for(int link = tip_link; link >= 0; link = robot_model->links[link].parent) chain.push_back(link);
            std::reverse(chain.begin(), chain.end());
            for(int link : chain)
            {
                if(std::find(link_schedule.begin(), link_schedule.end(), link) != link_schedule.end()) continue;
                link_schedule.push_back(link);
            }
        }
        size_t tip_count = tip_links.size();
        joint_dependencies.assign(robot_model->links.size(), {});
        for(int link : link_schedule) joint_dependencies[link].push_back(link);
        for(int link : link_schedule)
        {
            int mimic = robot_model->links[link].mimic;
            if(mimic >= 0)
            {
                while(robot_model->links[mimic].mimic >= 0 && robot_model->links[mimic].mimic != link) mimic = robot_model->links[mimic].mimic;
                joint_dependencies[mimic].push_back(link);
            }
        }
        tip_dependencies.assign(robot_model->links.size() * tip_count, 0);
        for(size_t tip_index = 0; tip_index < tip_count; tip_index++)
            for(int link = tip_links[tip_index]; link >= 0; link = robot_model->links[link].parent) tip_dependencies[link * tip_count + tip_index] = 1;
    }

    // :331-354
    void applyConfiguration(const std::vector<double>& jj0)
    {
        variables = jj0;
        updateMimic(variables);
        for(int link : link_schedule)
        {
            const auto& L = robot_model->links[link];
            Frame jf;
            getJointFrame(link, variables.data(), jf);
            if(L.parent >= 0)
                concat(global_frames[L.parent], L.origin, jf, global_frames[link]);
            else
                concat(L.origin, jf, global_frames[link]);
        }
        for(size_t itip = 0; itip < tip_links.size(); itip++) tip_frames[itip] = global_frames[tip_links[itip]];
    }
    const std::vector<Frame>& getTipFrames() const { return tip_frames; }

    double& jac(size_t row, size_t col) { return mutation_approx_jacobian[row * jac_cols + col]; }

    // :600-730
    void computeJacobian(const std::vector<size_t>& variable_indices)
    {
        size_t tip_count = tip_frames.size();
        jac_cols = variable_indices.size();
        mutation_approx_jacobian.assign(tip_count * 6 * jac_cols, 0.0);
        for(size_t icol = 0; icol < variable_indices.size(); icol++)
        {
            size_t ivar = variable_indices[icol];
            int var_joint = robot_model->var_joint[ivar];
            if(robot_model->links[var_joint].mimic >= 0) continue;
            for(int joint : joint_dependencies[var_joint])
            {
                double scale = 1;
                for(int m = joint; robot_model->links[m].mimic >= 0 && robot_model->links[m].mimic != joint; m = robot_model->links[m].mimic) scale *= robot_model->links[m].mimic_factor;
                const auto& J = robot_model->links[joint];
                switch(J.joint_type)
                {
                case FIXED: continue;
                case REVOLUTE:
                {
                    const Frame& link_frame = global_frames[joint];
                    for(size_t itip = 0; itip < tip_count; itip++)
                    {
                        if(!tip_dependencies[joint * tip_count + itip]) continue;
                        const Frame& tip_frame = tip_frames[itip];
                        Quat q = tf2_mul(inverse(link_frame.rot), tip_frame.rot);
                        q = inverse(q);
                        Vec3 rot = J.axis;
                        quat_mul_vec(q, rot, rot);
                        Vec3 vel = link_frame.pos - tip_frame.pos;
                        quat_mul_vec(inverse(tip_frame.rot), vel, vel);
                        vel = cross(vel, rot);
                        jac(itip * 6 + 0, icol) += vel.x * scale;
                        jac(itip * 6 + 1, icol) += vel.y * scale;
                        jac(itip * 6 + 2, icol) += vel.z * scale;
                        jac(itip * 6 + 3, icol) += rot.x * scale;
                        jac(itip * 6 + 4, icol) += rot.y * scale;
                        jac(itip * 6 + 5, icol) += rot.z * scale;
                    }
                    continue;
                }
                case PRISMATIC:
                {
                    const Frame& link_frame = global_frames[joint];
                    for(size_t itip = 0; itip < tip_count; itip++)
                    {
                        if(!tip_dependencies[joint * tip_count + itip]) continue;
                        const Frame& tip_frame = tip_frames[itip];
                        Quat q = tf2_mul(inverse(link_frame.rot), tip_frame.rot);
                        q = inverse(q);
                        Vec3 v;
                        quat_mul_vec(q, J.axis, v);
                        jac(itip * 6 + 0, icol) += v.x * scale;
                        jac(itip * 6 + 1, icol) += v.y * scale;
                        jac(itip * 6 + 2, icol) += v.z * scale;
                    }
                    continue;
                }
                default:
                {
                    // :695-726 numeric differentiation (floating joints; planar would take the same route through computeTransform)
                    const double step_size = 0.00001, inv_step_size = 1.0 / step_size;
                    size_t ivar2 = ivar;
                    if(J.mimic >= 0) ivar2 = ivar2 - robot_model->links[var_joint].first_var + J.first_var;
                    const Frame link_frame_1 = global_frames[joint];
                    double v0 = variables[ivar2];
                    variables[ivar2] = v0 + step_size;
                    Frame joint_frame_2;
                    getJointFrame(joint, variables.data(), joint_frame_2);
                    variables[ivar2] = v0;
                    Frame link_frame_2;
                    if(J.parent >= 0)
                        concat(global_frames[J.parent], J.origin, joint_frame_2, link_frame_2);
                    else
                        concat(J.origin, joint_frame_2, link_frame_2);
                    for(size_t itip = 0; itip < tip_count; itip++)
                    {
                        if(!tip_dependencies[joint * tip_count + itip]) continue;
                        const Frame tip_frame_1 = tip_frames[itip];
                        Frame tip_frame_2;
                        change(link_frame_2, link_frame_1, tip_frame_1, tip_frame_2);
                        Vec3 tv, tr;
                        frameTwist(tip_frame_1, tip_frame_2, opt.libm_sincos ? static_cast<double (*)(double)>(std::acos) : det_acos, tv, tr);
                        jac(itip * 6 + 0, icol) += tv.x * inv_step_size * scale;
                        jac(itip * 6 + 1, icol) += tv.y * inv_step_size * scale;
                        jac(itip * 6 + 2, icol) += tv.z * inv_step_size * scale;
                        jac(itip * 6 + 3, icol) += tr.x * inv_step_size * scale;
                        jac(itip * 6 + 4, icol) += tr.y * inv_step_size * scale;
                        jac(itip * 6 + 5, icol) += tr.z * inv_step_size * scale;
                    }
                    continue;
                }
                }
            }
        }
    }

    // :802-930
    void initializeMutationApproximator(const std::vector<size_t>& variable_indices)
    {
        mutation_approx_variable_indices = variable_indices;
        size_t tip_count = tip_links.size();
        tip_frames_aligned = tip_frames;
        if(mutation_approx_frames.size() < tip_count) mutation_approx_frames.resize(tip_count);
        for(size_t itip = 0; itip < tip_count; itip++) mutation_approx_frames[itip].resize(robot_model->n_vars);
        for(size_t itip = 0; itip < tip_count; itip++)
            for(auto ivar : variable_indices) mutation_approx_frames[itip][ivar] = Frame::identity();
        computeJacobian(variable_indices);
        for(size_t icol = 0; icol < variable_indices.size(); icol++)
        {
            size_t ivar = variable_indices[icol];
            for(size_t itip = 0; itip < tip_count; itip++)
            {
                {
                    Vec3 t(jac(itip * 6 + 0, icol), jac(itip * 6 + 1, icol), jac(itip * 6 + 2, icol));
                    quat_mul_vec(tip_frames[itip].rot, t, t);
                    mutation_approx_frames[itip][ivar].pos = t;
                }
                {
                    Quat q(jac(itip * 6 + 3, icol) * 0.5, jac(itip * 6 + 4, icol) * 0.5, jac(itip * 6 + 5, icol) * 0.5, 1.0);
                    quat_mul_quat(tip_frames[itip].rot, q, q);
                    q = q - tip_frames[itip].rot;
                    mutation_approx_frames[itip][ivar].rot = q;
                }
            }
        }
        if(mutation_approx_mask.size() < tip_count) mutation_approx_mask.resize(tip_count);
        if(mutation_approx_map.size() < tip_count) mutation_approx_map.resize(tip_count);
        for(size_t itip = 0; itip < tip_count; itip++)
        {
            if(mutation_approx_mask[itip].size() < robot_model->n_vars) mutation_approx_mask[itip].resize(robot_model->n_vars);
            mutation_approx_map[itip].clear();
            for(size_t ii = 0; ii < variable_indices.size(); ii++)
            {
                auto ivar = variable_indices[ii];
                auto& frame = mutation_approx_frames[itip][ivar];
                bool b = false;
                b |= (frame.pos.x != 0.0);
                b |= (frame.pos.y != 0.0);
                b |= (frame.pos.z != 0.0);
                b |= (frame.rot.x != 0.0);
                b |= (frame.rot.y != 0.0);
                b |= (frame.rot.z != 0.0);
                mutation_approx_mask[itip][ivar] = b;
                if(b) mutation_approx_map[itip].push_back(ii);
            }
        }
    }

    inline double madd(double f, double d, double acc) const { return opt.fma_approx ? std::fma(f, d, acc) : acc + d * f; }

    // :933-1058 (AVX+FMA form :944-950 when opt.fma_approx).  Q2: unmasked tips copy the input.
    void computeApproximateMutation1(size_t variable_index, double variable_delta, const std::vector<Frame>& input, std::vector<Frame>& output) const
    {
        size_t tip_count = tip_links.size();
        output.resize(tip_count);
        for(size_t itip = 0; itip < tip_count; itip++)
        {
            if(mutation_approx_mask[itip][variable_index] == 0)
            {
                if(!opt.stale_tips) output[itip] = input[itip]; // intended semantics, see header comment (reference leaves stale data)
                continue;
            }
            const Frame& jd = mutation_approx_frames[itip][variable_index];
            const Frame& tf = input[itip];
            Frame o;
            auto madd1 = [&](double f, double d, double acc) { return opt.fma_approx1 ? madd(f, d, acc) : acc + d * f; };
            o.pos.x = madd1(variable_delta, jd.pos.x, tf.pos.x);
            o.pos.y = madd1(variable_delta, jd.pos.y, tf.pos.y);
            o.pos.z = madd1(variable_delta, jd.pos.z, tf.pos.z);
            o.rot.x = madd1(variable_delta, jd.rot.x, tf.rot.x);
            o.rot.y = madd1(variable_delta, jd.rot.y, tf.rot.y);
            o.rot.z = madd1(variable_delta, jd.rot.z, tf.rot.z);
            o.rot.w = madd1(variable_delta, jd.rot.w, tf.rot.w);
            output[itip] = o;
        }
    }

    // :1061-1233 (AVX+FMA form :1075-1107 when opt.fma_approx)
    void computeApproximateMutations(size_t mutation_count, const double* const* mutation_values, std::vector<std::vector<Frame>>& tip_frame_mutations) const
    {
        const double* p_variables = variables.data();
        size_t tip_count = tip_links.size();
        tip_frame_mutations.resize(mutation_count);
        for(auto& m : tip_frame_mutations) m.resize(tip_count);
        for(size_t itip = 0; itip < tip_count; itip++)
        {
            const auto& joint_deltas = mutation_approx_frames[itip];
            const Frame& tip_frame = tip_frames_aligned[itip];
            for(size_t imutation = 0; imutation < mutation_count; imutation++)
            {
                Frame o = tip_frame;
                for(size_t vii : mutation_approx_map[itip])
                {
                    size_t variable_index = mutation_approx_variable_indices[vii];
                    double variable_delta = mutation_values[imutation][vii] - p_variables[variable_index];
                    const Frame& jd = joint_deltas[variable_index];
                    o.pos.x = madd(variable_delta, jd.pos.x, o.pos.x);
                    o.pos.y = madd(variable_delta, jd.pos.y, o.pos.y);
                    o.pos.z = madd(variable_delta, jd.pos.z, o.pos.z);
                    o.rot.x = madd(variable_delta, jd.rot.x, o.rot.x);
                    o.rot.y = madd(variable_delta, jd.rot.y, o.rot.y);
                    o.rot.z = madd(variable_delta, jd.rot.z, o.rot.z);
                    o.rot.w = madd(variable_delta, jd.rot.w, o.rot.w);
                }
                tip_frame_mutations[imutation][itip] = o;
            }
        }
    }
};

// ---------------------------------------------------------------------------
// Goals (include/bio_ik/goal_types.h evaluate() bodies) and Problem
// (src/problem.h, src/problem.cpp), flattened the way the C ABI carries them.
// ---------------------------------------------------------------------------
enum GoalKind
{
    G_POSITION = 1,
    G_ORIENTATION = 2,
    G_POSE = 3,
    G_LOOK_AT = 4,
    G_MAX_DISTANCE = 5,
    G_MIN_DISTANCE = 6,
    G_LINE = 7,
    G_PLANE = 8,
    G_AVOID_JOINT_LIMITS = 9,
    G_CENTER_JOINTS = 10,
    G_REGULARIZATION = 11,
    G_MINIMAL_DISPLACEMENT = 12,
    G_JOINT_VARIABLE = 13,
    G_SIDE = 14,
    G_DIRECTION = 15,
    G_CONE = 16,
    G_BALANCE = 17
};
static const int GOAL_NPARAM = 12;

// BalanceGoal::balance_infos, src/goal_types.cpp:231-259
struct BalanceInfo
{
    size_t tip_index; // goal_link_indices_[i]
    Vec3 center;
    double weight;
};

struct GoalInfo
{
    int type = 0;
    int tip_index = 0; // goal_link_indices_[0]
    bool secondary = false;
    long var_index = 0; // goal_variable_indices_[0]: gene index, or -1-robot_var for a fixed joint (goal.h:70-77)
    double weight = 1, weight_sq = 1;
    double p[GOAL_NPARAM] = {0};
};

struct Problem
{
    const RobotModel* robot_model = nullptr;
    RobotInfo modelInfo;
    std::vector<size_t> tip_link_indices;
    std::vector<size_t> active_variables;
    std::vector<GoalInfo> goals, secondary_goals;
    std::vector<double> initial_guess;
    std::vector<double> minimal_displacement_factors;
    std::vector<BalanceInfo> balance_infos; // filled by describeBalance() when the problem has a BalanceGoal
    double dpos = DBL_MAX, drot = DBL_MAX, dtwist = 1e-5;

    // BalanceGoal::describe, src/goal_types.cpp:231-259: the links with a positive inertial mass in link order, weight = mass / total
    void describeBalance()
    {
        balance_infos.clear();
        double total = 0.0;
        for(size_t l = 0; l < robot_model->links.size() && !robot_model->link_mass.empty(); l++)
        {
            double mass = robot_model->link_mass[l];
            if(!(mass > 0)) continue;
            size_t tip = tip_link_indices.size();
            for(size_t t = 0; t < tip_link_indices.size(); t++)
                if(tip_link_indices[t] == l) tip = t;
            if(tip == tip_link_indices.size()) throw std::runtime_error("oracle: BalanceGoal needs every link with mass among the tip links");
            balance_infos.push_back(BalanceInfo{tip, Vec3(robot_model->link_com[3 * l], robot_model->link_com[3 * l + 1], robot_model->link_com[3 * l + 2]), mass});
            total += mass;
        }
        for(auto& b : balance_infos) b.weight /= total;
    }

    // src/problem.cpp:206-225
    void initVelocityWeights()
    {
        minimal_displacement_factors.resize(active_variables.size());
        double s = 0;
        for(auto ivar : active_variables) s += modelInfo.getMaxVelocityRcp(ivar);
        if(s > 0)
        {
            for(size_t i = 0; i < active_variables.size(); i++) minimal_displacement_factors[i] = modelInfo.getMaxVelocityRcp(active_variables[i]) / s;
        }
        else
        {
            for(size_t i = 0; i < active_variables.size(); i++) minimal_displacement_factors[i] = 1.0 / active_variables.size();
        }
    }
    // src/problem.cpp:90-95
    void sanitizeThresholds()
    {
        if(dpos < 0.0 || dpos >= FLT_MAX || !std::isfinite(dpos)) dpos = DBL_MAX;
        if(drot < 0.0 || drot >= FLT_MAX || !std::isfinite(drot)) drot = DBL_MAX;
        if(dtwist < 0.0 || dtwist >= FLT_MAX || !std::isfinite(dtwist)) dtwist = DBL_MAX;
    }

    // Goal::evaluate bodies
    double evaluate(const GoalInfo& g, const Frame* tip_frames, const double* x) const
    {
        const double* p = g.p;
        switch(g.type)
        {
        case G_POSITION: // goal_types.h:96
            return distance2(tip_frames[g.tip_index].pos, Vec3(p[0], p[1], p[2]));
        case G_ORIENTATION: // goal_types.h:115-119
        {
            Quat o(p[3], p[4], p[5], p[6]);
            const Quat& q = tip_frames[g.tip_index].rot;
            return std::fmin(length2(o - q), length2(o + q));
        }
        case G_POSE: // goal_types.h:149-180
        {
            double e = 0.0;
            e += distance2(tip_frames[g.tip_index].pos, Vec3(p[0], p[1], p[2]));
            Quat o(p[3], p[4], p[5], p[6]);
            const Quat& q = tip_frames[g.tip_index].rot;
            e += std::fmin(length2(o - q), length2(o + q)) * (p[7] * p[7]);
            return e;
        }
        case G_LOOK_AT: // goal_types.h:204-211
        {
            const Frame& fb = tip_frames[g.tip_index];
            Vec3 axis;
            quat_mul_vec(fb.rot, Vec3(p[0], p[1], p[2]), axis);
            return distance2(normalized(Vec3(p[3], p[4], p[5]) - fb.pos), normalized(axis));
        }
        case G_MAX_DISTANCE: // goal_types.h:235-240
        {
            double d = std::fmax(0.0, distance(tip_frames[g.tip_index].pos, Vec3(p[0], p[1], p[2])) - p[3]);
            return d * d;
        }
        case G_MIN_DISTANCE: // goal_types.h:264-269
        {
            double d = std::fmax(0.0, p[3] - distance(tip_frames[g.tip_index].pos, Vec3(p[0], p[1], p[2])));
            return d * d;
        }
        case G_LINE: // goal_types.h:293-297
        {
            const Frame& fb = tip_frames[g.tip_index];
            Vec3 position(p[0], p[1], p[2]), direction(p[3], p[4], p[5]);
            return distance2(position, fb.pos - direction * dot(direction, fb.pos - position));
        }
        case G_PLANE: // goal_types.h:321-327
        {
            Vec3 position(p[0], p[1], p[2]), normal(p[3], p[4], p[5]);
            double signed_dist = dot(tip_frames[g.tip_index].pos - position, normal);
            return signed_dist * signed_dist;
        }
        case G_AVOID_JOINT_LIMITS: // goal_types.h:387-401
        {
            double sum = 0.0;
            for(size_t i = 0; i < active_variables.size(); i++)
            {
                size_t ivar = active_variables[i];
                if(modelInfo.getClipMax(ivar) == DBL_MAX) continue;
                double d = x[i] - (modelInfo.getMin(ivar) + modelInfo.getMax(ivar)) * 0.5;
                d = std::fmax(0.0, std::fabs(d) * 2.0 - modelInfo.getSpan(ivar) * 0.5);
                d *= minimal_displacement_factors[i];
                sum += d * d;
            }
            return sum;
        }
        case G_CENTER_JOINTS: // goal_types.h:412-425
        {
            double sum = 0.0;
            for(size_t i = 0; i < active_variables.size(); i++)
            {
                size_t ivar = active_variables[i];
                if(modelInfo.getClipMax(ivar) == DBL_MAX) continue;
                double d = x[i] - (modelInfo.getMin(ivar) + modelInfo.getMax(ivar)) * 0.5;
                d *= minimal_displacement_factors[i];
                sum += d * d;
            }
            return sum;
        }
        case G_REGULARIZATION: // goal_types.h:435-444
        {
            double sum = 0.0;
            for(size_t i = 0; i < active_variables.size(); i++)
            {
                double d = x[i] - initial_guess[active_variables[i]];
                sum += d * d;
            }
            return sum;
        }
        case G_MINIMAL_DISPLACEMENT: // goal_types.h:455-465
        {
            double sum = 0.0;
            for(size_t i = 0; i < active_variables.size(); i++)
            {
                double d = x[i] - initial_guess[active_variables[i]];
                d *= minimal_displacement_factors[i];
                sum += d * d;
            }
            return sum;
        }
        case G_JOINT_VARIABLE: // goal_types.h:494-498, goal.h:70-77
        {
            double v = g.var_index >= 0 ? x[g.var_index] : initial_guess[-1 - g.var_index];
            double d = p[0] - v;
            return d * d;
        }
        case G_SIDE: // goal_types.h:606-613
        {
            Vec3 v;
            quat_mul_vec(tip_frames[g.tip_index].rot, Vec3(p[0], p[1], p[2]), v);
            double f = std::fmax(0.0, dot(v, Vec3(p[3], p[4], p[5])));
            return f * f;
        }
        case G_DIRECTION: // goal_types.h:637-643
        {
            Vec3 v;
            quat_mul_vec(tip_frames[g.tip_index].rot, Vec3(p[0], p[1], p[2]), v);
            return distance2(v, Vec3(p[3], p[4], p[5]));
        }
        case G_CONE: // goal_types.h:700-711; tf2::Vector3::angle(v) = tf2Acos(dot(v) / sqrt(length2() * v.length2())), tf2Acos clamps to [-1, 1]
        {
            double sum = 0.0;
            const Frame& fb = tip_frames[g.tip_index];
            Vec3 v;
            quat_mul_vec(fb.rot, Vec3(p[4], p[5], p[6]), v);
            Vec3 direction(p[7], p[8], p[9]);
            double s = std::sqrt(length2(v) * length2(direction));
            double c = dot(v, direction) / s;
            if(c < -1.0) c = -1.0;
            if(c > 1.0) c = 1.0;
            double d = std::fmax(0.0, det_acos(c) - p[10]);
            sum += d * d;
            double w = p[3];
            sum += w * w * length2(Vec3(p[0], p[1], p[2]) - fb.pos);
            return sum;
        }
        case G_BALANCE: // src/goal_types.cpp:261-272
        {
            Vec3 center(0, 0, 0);
            for(size_t i = 0; i < balance_infos.size(); i++)
            {
                auto& info = balance_infos[i];
                auto& frame = tip_frames[info.tip_index];
                Vec3 c = info.center;
                quat_mul_vec(frame.rot, c, c);
                c = c + frame.pos;
                center = center + c * info.weight;
            }
            Vec3 target(p[0], p[1], p[2]), axis(p[3], p[4], p[5]);
            center = center - target;
            center = center - axis * dot(axis, center);
            return length2(center);
        }
        default: throw std::runtime_error("oracle: unsupported goal type");
        }
    }
    // src/problem.cpp:244-257
    double computeGoalFitness(const GoalInfo& g, const Frame* tip_frames, const double* x) const { return evaluate(g, tip_frames, x) * g.weight_sq; }
    double computeGoalFitness(const std::vector<GoalInfo>& gg, const Frame* tip_frames, const double* x) const
    {
        double sum = 0.0;
        for(auto& g : gg) sum += computeGoalFitness(g, tip_frames, x);
        return sum;
    }

    // src/problem.cpp:259-341.  KDL semantics per SURVEY.md Appendix C:
    //   kdl_diff.vel = Ra^T (pb - pa);  kdl_diff.rot = Ra^T (Ra * rotvec(Ra^T Rb)) = rotvec(Ra^T Rb)
    //   KDL::Equal(a, 0, eps): every |component| < eps.
    static void quatToMatrix(const Quat& q, double R[9]) // KDL::Rotation::Quaternion
    {
        double x = q.x, y = q.y, z = q.z, w = q.w;
        double x2 = x * x, y2 = y * y, z2 = z * z, w2 = w * w;
        R[0] = w2 + x2 - y2 - z2;
        R[1] = 2 * x * y - 2 * w * z;
        R[2] = 2 * x * z + 2 * w * y;
        R[3] = 2 * x * y + 2 * w * z;
        R[4] = w2 - x2 + y2 - z2;
        R[5] = 2 * y * z - 2 * w * x;
        R[6] = 2 * x * z - 2 * w * y;
        R[7] = 2 * y * z + 2 * w * x;
        R[8] = w2 - x2 - y2 + z2;
    }
    static void kdlTwist(const Frame& fa, const Frame& fb, double vel[3], double rot[3])
    {
        double Ra[9], Rb[9];
        quatToMatrix(fa.rot, Ra);
        quatToMatrix(fb.rot, Rb);
        double d[3] = {fb.pos.x - fa.pos.x, fb.pos.y - fa.pos.y, fb.pos.z - fa.pos.z};
        for(int i = 0; i < 3; i++) vel[i] = Ra[0 + i] * d[0] + Ra[3 + i] * d[1] + Ra[6 + i] * d[2];
        double M[9]; // Ra^T Rb
        for(int i = 0; i < 3; i++)
            for(int j = 0; j < 3; j++) M[i * 3 + j] = Ra[0 + i] * Rb[0 + j] + Ra[3 + i] * Rb[3 + j] + Ra[6 + i] * Rb[6 + j];
        // KDL::Rotation::GetRot(): axis * angle
        double ax = M[7] - M[5], ay = M[2] - M[6], az = M[3] - M[1];
        double sa = std::sqrt(ax * ax + ay * ay + az * az) * 0.5;
        double ca = (M[0] + M[4] + M[8] - 1.0) * 0.5;
        double angle = std::atan2(sa, ca);
        if(sa > 1e-12)
        {
            double f = angle / (2.0 * sa);
            rot[0] = ax * f;
            rot[1] = ay * f;
            rot[2] = az * f;
        }
        else if(ca > 0)
        {
            rot[0] = ax * 0.5;
            rot[1] = ay * 0.5;
            rot[2] = az * 0.5;
        }
        else
        {
            // angle ~ pi: axis from the diagonal (never a success, any finite value will do)
            rot[0] = angle;
            rot[1] = 0;
            rot[2] = 0;
        }
    }
    static bool allBelow(const double v[3], double eps) { return std::fabs(v[0]) < eps && std::fabs(v[1]) < eps && std::fabs(v[2]) < eps; }
    static double angleShortestPath(const Quat& a, const Quat& b) // tf2::Quaternion::angleShortestPath
    {
        double s = std::sqrt(length2(a) * length2(b));
        double d = dot(a, b);
        if(d < 0)
            return std::acos(dot(a, Quat(-b.x, -b.y, -b.z, -b.w)) / s) * 2.0;
        else
            return std::acos(d / s) * 2.0;
    }
    bool checkSolutionActiveVariables(const std::vector<Frame>& tip_frames, const double* x) const
    {
        for(auto& goal : goals)
        {
            Frame fa = Frame::identity();
            const Frame& fb = tip_frames[goal.tip_index];
            switch(goal.type)
            {
            case G_POSITION:
            {
                fa.pos = Vec3(goal.p[0], goal.p[1], goal.p[2]);
                if(dpos != DBL_MAX)
                {
                    double p_dist = length(fb.pos - fa.pos);
                    if(!(p_dist <= dpos)) return false;
                }
                if(dtwist != DBL_MAX)
                {
                    double vel[3], rot[3];
                    kdlTwist(fa, fb, vel, rot);
                    if(!allBelow(vel, dtwist)) return false;
                }
                continue;
            }
            case G_ORIENTATION:
            {
                fa.rot = Quat(goal.p[3], goal.p[4], goal.p[5], goal.p[6]);
                if(drot != DBL_MAX)
                {
                    double r_dist = angleShortestPath(fb.rot, fa.rot) * 180 / M_PI;
                    if(!(r_dist <= drot)) return false;
                }
                if(dtwist != DBL_MAX)
                {
                    double vel[3], rot[3];
                    kdlTwist(fa, fb, vel, rot);
                    if(!allBelow(rot, dtwist)) return false;
                }
                continue;
            }
            case G_POSE:
            {
                fa.pos = Vec3(goal.p[0], goal.p[1], goal.p[2]);
                fa.rot = Quat(goal.p[3], goal.p[4], goal.p[5], goal.p[6]);
                if(dpos != DBL_MAX || drot != DBL_MAX)
                {
                    double p_dist = length(fb.pos - fa.pos);
                    double r_dist = angleShortestPath(fb.rot, fa.rot) * 180 / M_PI;
                    if(!(p_dist <= dpos)) return false;
                    if(!(r_dist <= drot)) return false;
                }
                if(dtwist != DBL_MAX)
                {
                    double vel[3], rot[3];
                    kdlTwist(fa, fb, vel, rot);
                    if(!allBelow(vel, dtwist) || !allBelow(rot, dtwist)) return false;
                }
                continue;
            }
            default:
            {
                double dmax = DBL_MAX;
                dmax = std::fmin(dmax, dpos);
                dmax = std::fmin(dmax, dtwist);
                double d = computeGoalFitness(goal, tip_frames.data(), x);
                if(!(d < dmax * dmax)) return false;
            }
            }
        }
        return true;
    }
};

// ---------------------------------------------------------------------------
// IKBase (src/ik_base.h:128-210) + IKEvolution2<memetic> (src/ik_evolution_2.cpp)
// ---------------------------------------------------------------------------
struct SolverConfig
{
    size_t population = 18; // children.size(): 2 parents + child_count (src/ik_evolution_2.cpp:137-138,182)
    size_t generations = 8; // :349-350
    int memetic = 'q';      // 0, 'q', 'l'  (:652-654)
    size_t memetic_iters = 8; // :453
};

struct IKEvolution2 : Random
{
    struct Individual
    {
        std::vector<double> genes, gradients;
        double fitness = 0;
    };
    struct Species
    {
        std::vector<Individual> individuals;
        double fitness = 0;
        bool improved = false;
    };

    SolverConfig cfg;
    RobotFK model;
    RobotInfo modelInfo;
    Problem problem;
    std::vector<Frame> null_tip_frames;
    std::vector<double> initial_guess, solution, temp_joint_variables;
    double solution_fitness = 0;
    std::vector<Species> species;
    std::vector<Individual> children;
    std::vector<std::vector<Frame>> phenotypes, phenotypes2, phenotypes3;
    std::vector<size_t> child_indices;
    std::vector<double*> genotypes;
    std::vector<size_t> quaternion_genes;
    std::vector<double> genes_min, genes_max, genes_span, gradient, temp;
    std::vector<double> temp_active_variable_positions;

    IKEvolution2(const RobotModel* robot, const Tables& tables, uint32_t seed, const SolverConfig& c, Options opt = Options()) : Random(tables, seed), cfg(c), model(robot, opt), modelInfo(*robot) {}

    // src/ik_base.h:163-207
    double computeSecondaryFitnessActiveVariables(const double* x) { return problem.computeGoalFitness(problem.secondary_goals, null_tip_frames.data(), x); }
    double computeFitnessActiveVariables(const std::vector<Frame>& tip_frames, const double* x) { return problem.computeGoalFitness(problem.goals, tip_frames.data(), x); }
    double computeCombinedFitnessActiveVariables(const std::vector<Frame>& tip_frames, const double* x)
    {
        double ret = 0.0;
        ret += problem.computeGoalFitness(problem.goals, tip_frames.data(), x);
        ret += problem.computeGoalFitness(problem.secondary_goals, null_tip_frames.data(), x);
        return ret;
    }
    double* extractActiveVariables(const std::vector<double>& variable_positions)
    {
        temp_active_variable_positions.resize(problem.active_variables.size());
        for(size_t i = 0; i < temp_active_variable_positions.size(); i++) temp_active_variable_positions[i] = variable_positions[problem.active_variables[i]];
        return temp_active_variable_positions.data();
    }
    double computeFitness(const std::vector<double>& variable_positions, const std::vector<Frame>& tip_frames) { return computeFitnessActiveVariables(tip_frames, extractActiveVariables(variable_positions)); }
    double computeFitness(const std::vector<double>& variable_positions)
    {
        model.applyConfiguration(variable_positions);
        return computeFitness(variable_positions, model.getTipFrames());
    }
    bool checkSolution(const std::vector<double>& variable_positions, const std::vector<Frame>& tips) { return problem.checkSolutionActiveVariables(tips, extractActiveVariables(variable_positions)); }

    // src/ik_evolution_2.cpp:101-107
    void genesToJointVariables(const Individual& individual, std::vector<double>& variables)
    {
        variables.resize(model.robot_model->n_vars);
        for(size_t i = 0; i < problem.active_variables.size(); i++) variables[problem.active_variables[i]] = individual.genes[i];
    }
    const std::vector<double>& getSolution() const { return solution; }

    // src/ik_base.h:154-161 + src/ik_evolution_2.cpp:111-230
    void initialize(const Problem& p)
    {
        problem = p;
        model.initialize(problem.tip_link_indices);
        null_tip_frames.assign(problem.tip_link_indices.size(), Frame()); // reference: uninitialised; secondary goals must not read them

        quaternion_genes.clear();
        for(size_t igene = 0; igene < problem.active_variables.size(); igene++)
        {
            size_t ivar = problem.active_variables[igene];
            int j = model.robot_model->var_joint[ivar];
            if((size_t)model.robot_model->links[j].first_var + 3 != ivar) continue;
            if(model.robot_model->links[j].joint_type != FLOATING) continue;
            quaternion_genes.push_back(igene);
        }

        initial_guess = problem.initial_guess;
        solution = initial_guess;
        solution_fitness = computeFitness(solution);
        temp_joint_variables = initial_guess;

        size_t population_size = 2;
        size_t child_count = cfg.population - population_size;

        // Q4 batch contract: a fresh solver per query => Species value-initialised (fitness 0.0, improved false)
        species.clear();
        species.resize(2);
        for(auto& s : species)
        {
            s.individuals.resize(population_size);
            auto& v = s.individuals[0];
            v.genes.resize(problem.active_variables.size());
            for(size_t i = 0; i < v.genes.size(); i++) v.genes[i] = initial_guess[problem.active_variables[i]];
            v.gradients.clear();
            v.gradients.resize(problem.active_variables.size(), 0);
            for(size_t i = 1; i < s.individuals.size(); i++)
            {
                s.individuals[i].genes = s.individuals[0].genes;
                s.individuals[i].gradients = s.individuals[0].gradients;
            }
        }
        children.resize(population_size + child_count);
        for(auto& child : children)
        {
            child.genes.resize(problem.active_variables.size());
            child.gradients.resize(problem.active_variables.size());
        }
        genes_min.resize(problem.active_variables.size());
        genes_max.resize(problem.active_variables.size());
        genes_span.resize(problem.active_variables.size());
        for(size_t i = 0; i < problem.active_variables.size(); i++)
        {
            genes_min[i] = modelInfo.getClipMin(problem.active_variables[i]);
            genes_max[i] = modelInfo.getClipMax(problem.active_variables[i]);
            genes_span[i] = modelInfo.getSpan(problem.active_variables[i]);
        }
    }

    // src/ik_evolution_2.cpp:242-326
    void reproduce(const std::vector<Individual>& population)
    {
        auto gene_count = children[0].genes.size();
        size_t s = (children.size() - population.size()) * gene_count + children.size() * 4 + 4;
        const double* rr = fast_random_gauss_n(s);
        // :257 rounds the BYTE address up to a multiple of 4: a no-op for an 8-byte aligned pointer (Q1)
        for(size_t child_index = population.size(); child_index < children.size(); child_index++)
        {
            double mutation_rate = (1 << fast_random_index(16)) * (1.0 / (1 << 23));
            auto& parent = population[0];
            auto& parent2 = population[1];
            double fmix = (child_index % 2 == 0) * 0.2;
            double gradient_factor = child_index % 3;
            auto& child = children[child_index];
            for(size_t gene_index = 0; gene_index < gene_count; gene_index++)
            {
                double r = rr[gene_index];
                double f = mutation_rate * genes_span[gene_index];
                double gene = parent.genes[gene_index];
                double parent_gene = gene;
                gene += r * f;
                double parent_gradient = mix(parent.gradients[gene_index], parent2.gradients[gene_index], fmix);
                double gradient = parent_gradient * gradient_factor;
                gene += gradient;
                gene = clamp(gene, genes_min[gene_index], genes_max[gene_index]);
                child.genes[gene_index] = gene;
                child.gradients[gene_index] = mix(parent_gradient, gene - parent_gene, 0.3);
            }
            rr += (gene_count + 3) / 4 * 4;
            for(auto quaternion_gene_index : quaternion_genes)
            {
                Quat q(child.genes[quaternion_gene_index], child.genes[quaternion_gene_index + 1], child.genes[quaternion_gene_index + 2], child.genes[quaternion_gene_index + 3]);
                normalizeFast(q);
                child.genes[quaternion_gene_index] = q.x;
                child.genes[quaternion_gene_index + 1] = q.y;
                child.genes[quaternion_gene_index + 2] = q.z;
                child.genes[quaternion_gene_index + 3] = q.w;
            }
        }
    }

    // src/ik_evolution_2.cpp:328-646
    void step()
    {
        for(size_t ispecies = 0; ispecies < species.size(); ispecies++)
        {
            auto& species = this->species[ispecies];
            auto& population = species.individuals;
            {
                // :341-346
                genesToJointVariables(species.individuals[0], temp_joint_variables);
                model.applyConfiguration(temp_joint_variables);
                model.initializeMutationApproximator(problem.active_variables);

                for(size_t generation = 0; generation < cfg.generations; generation++)
                {
                    reproduce(population); // :360
                    size_t child_count = children.size();
                    // :366-378 pre-selection by secondary objectives
                    if(problem.secondary_goals.size())
                    {
                        child_count = random_index(children.size() - population.size() - 1) + 1 + population.size();
                        for(size_t child_index = population.size(); child_index < children.size(); child_index++) children[child_index].fitness = computeSecondaryFitnessActiveVariables(children[child_index].genes.data());
                        // reference: std::sort (unstable).  Oracle: stable order — a defined tie rule the GPU can match (SURVEY.md §7 hard part 6).
                        std::stable_sort(children.begin() + population.size(), children.end(), [](const Individual& a, const Individual& b) { return a.fitness < b.fitness; });
                    }
                    // :381-388 keep parents
                    for(size_t i = 0; i < population.size(); i++)
                    {
                        children[i].genes = population[i].genes;
                        children[i].gradients = population[i].gradients;
                    }
                    // :391-398 genotype-phenotype mapping
                    genotypes.resize(child_count);
                    for(size_t i = 0; i < child_count; i++) genotypes[i] = children[i].genes.data();
                    model.computeApproximateMutations(child_count, genotypes.data(), phenotypes);
                    // :401-407 fitness
                    for(size_t child_index = 0; child_index < child_count; child_index++) children[child_index].fitness = computeFitnessActiveVariables(phenotypes[child_index], genotypes[child_index]);
                    // :410-431 selection
                    child_indices.resize(child_count);
                    for(size_t i = 0; i < child_count; i++) child_indices[i] = i;
                    for(size_t i = 0; i < population.size(); i++)
                    {
                        size_t jmin = i;
                        double fmin = children[child_indices[i]].fitness;
                        for(size_t j = i + 1; j < child_count; j++)
                        {
                            double f = children[child_indices[j]].fitness;
                            if(f < fmin) jmin = j, fmin = f;
                        }
                        std::swap(child_indices[i], child_indices[jmin]);
                    }
                    for(size_t i = 0; i < population.size(); i++)
                    {
                        std::swap(population[i].genes, children[child_indices[i]].genes);
                        std::swap(population[i].gradients, children[child_indices[i]].gradients);
                    }
                }
            }

            // :436-570 memetic optimisation
            if(cfg.memetic == 'q' || cfg.memetic == 'l')
            {
                auto& individual = population[0];
                gradient.resize(problem.active_variables.size());
                if(genotypes.empty()) genotypes.emplace_back();
                phenotypes2.resize(1);
                phenotypes3.resize(1);
                double dp = 0.0000001;
                if(fast_random() < 0.5) dp = -dp;
                for(size_t generation = 0; generation < cfg.memetic_iters; generation++)
                {
                    temp = individual.genes;
                    genotypes[0] = temp.data();
                    model.computeApproximateMutations(1, genotypes.data(), phenotypes2);
                    double f2p = computeFitnessActiveVariables(phenotypes2[0], genotypes[0]);
                    double fa = f2p + computeSecondaryFitnessActiveVariables(genotypes[0]);
                    for(size_t i = 0; i < problem.active_variables.size(); i++)
                    {
                        genotypes[0][i] = individual.genes[i] + dp;
                        model.computeApproximateMutation1(problem.active_variables[i], +dp, phenotypes2[0], phenotypes3[0]);
                        double fb = computeCombinedFitnessActiveVariables(phenotypes3[0], genotypes[0]);
                        genotypes[0][i] = individual.genes[i];
                        double d = fb - fa;
                        gradient[i] = d;
                    }
                    double sum = dp * dp;
                    for(size_t i = 0; i < problem.active_variables.size(); i++) sum += std::fabs(gradient[i]);
                    double f = 1.0 / sum * dp;
                    for(size_t i = 0; i < problem.active_variables.size(); i++) gradient[i] *= f;

                    for(size_t i = 0; i < problem.active_variables.size(); i++) genotypes[0][i] = individual.genes[i] - gradient[i];
                    model.computeApproximateMutations(1, genotypes.data(), phenotypes3);
                    double f1 = computeCombinedFitnessActiveVariables(phenotypes3[0], genotypes[0]);
                    double f2 = fa;
                    for(size_t i = 0; i < problem.active_variables.size(); i++) genotypes[0][i] = individual.genes[i] + gradient[i];
                    model.computeApproximateMutations(1, genotypes.data(), phenotypes3);
                    double f3 = computeCombinedFitnessActiveVariables(phenotypes3[0], genotypes[0]);

                    if(cfg.memetic == 'q')
                    {
                        double v1 = (f2 - f1);
                        double v2 = (f3 - f2);
                        double v = (v1 + v2) * 0.5;
                        double a = (v1 - v2);
                        double step_size = v / a;
                        for(size_t i = 0; i < problem.active_variables.size(); i++) genotypes[0][i] = modelInfo.clip(individual.genes[i] + gradient[i] * step_size * 1.0, problem.active_variables[i]);
                        model.computeApproximateMutations(1, genotypes.data(), phenotypes2);
                        double f4p = computeFitnessActiveVariables(phenotypes2[0], genotypes[0]);
                        if(f4p < f2p)
                        {
                            individual.genes = temp;
                            continue;
                        }
                        else
                            break;
                    }
                    if(cfg.memetic == 'l')
                    {
                        double cost_diff = (f3 - f1) * 0.5;
                        double step_size = f2 / cost_diff;
                        for(size_t i = 0; i < problem.active_variables.size(); i++) temp[i] = modelInfo.clip(individual.genes[i] - gradient[i] * step_size, problem.active_variables[i]);
                        model.computeApproximateMutations(1, genotypes.data(), phenotypes2);
                        double f4p = computeFitnessActiveVariables(phenotypes2[0], genotypes[0]);
                        if(f4p < f2p)
                        {
                            individual.genes = temp;
                            continue;
                        }
                        else
                            break;
                    }
                }
            }
        }

        // :604-645 species block
        for(auto& species : this->species)
        {
            genesToJointVariables(species.individuals[0], temp_joint_variables);
            double fitness = computeFitness(temp_joint_variables);
            species.improved = (fitness != species.fitness);
            species.fitness = fitness;
        }
        std::stable_sort(species.begin(), species.end(), [](const Species& a, const Species& b) { return a.fitness < b.fitness; }); // 2 elements: == std::sort
        for(size_t species_index = 1; species_index < species.size(); species_index++)
        {
            if(fast_random() < 0.1 || !species[species_index].improved)
            {
                {
                    auto& individual = species[species_index].individuals[0];
                    for(size_t i = 0; i < individual.genes.size(); i++) individual.genes[i] = random(modelInfo.getMin(problem.active_variables[i]), modelInfo.getMax(problem.active_variables[i]));
                    for(auto& v : individual.gradients) v = 0;
                }
                for(size_t i = 0; i < species[species_index].individuals.size(); i++) species[species_index].individuals[i] = species[species_index].individuals[0];
            }
        }
        if(species[0].fitness < solution_fitness)
        {
            genesToJointVariables(species[0].individuals[0], solution);
            solution_fitness = species[0].fitness;
        }
    }
};

// ---------------------------------------------------------------------------
// Batch contract around IKParallel::solverthread (src/ik_parallel.h:148-190):
// the wall-clock timeout becomes a step budget; success is tested after steps
// 1,5,9,... (the reference tests after the first step of every 4-step burst
// ... precisely: step(); up to 3 more step()s; test) and after the last step.
// ---------------------------------------------------------------------------
struct QueryResult
{
    std::vector<double> solution;
    double fitness = 0;
    int success = 0;
    int steps = 0;
};

inline QueryResult solveQuery(const RobotModel& robot, const Tables& tables, const Problem& problem, uint32_t rng_seed, const SolverConfig& cfg, int steps, bool early_exit, Options opt = Options(), IKEvolution2** keep = nullptr)
{
    IKEvolution2* solver = new IKEvolution2(&robot, tables, rng_seed, cfg, opt);
    solver->initialize(problem);
    QueryResult res;
    int done = 0;
    bool success = false;
    while(done < steps)
    {
        // src/ik_parallel.h:165-168: one step, then up to three more
        int burst = std::min(4, steps - done);
        for(int k = 0; k < burst; k++) solver->step();
        done += burst;
        // :173-181
        std::vector<double> result = solver->getSolution();
        solver->model.applyConfiguration(result);
        success = solver->checkSolution(result, solver->model.getTipFrames());
        if(success && early_exit) break;
    }
    res.solution = solver->getSolution();
    solver->model.applyConfiguration(res.solution);
    res.success = solver->checkSolution(res.solution, solver->model.getTipFrames());
    res.fitness = solver->computeFitness(res.solution, solver->model.getTipFrames());
    res.steps = done;
    if(keep)
        *keep = solver;
    else
        delete solver;
    return res;
}

// ---------------------------------------------------------------------------
// After the solver: what IKParallel and the MoveIt plugin do with per-thread results.
// ---------------------------------------------------------------------------
// src/ik_parallel.h:218-258 — choice among `count` solver results of one problem (threads in the reference, islands
// in the batch build).  solutions: full variable vectors; fitness: primary fitness of each (ik_parallel.h:181).
inline size_t selectBestResult(Problem& problem, const std::vector<std::vector<double>>& solutions, const std::vector<double>& fitness, const std::vector<int>& success, double& best_fitness)
{
    size_t count = solutions.size(), best_index = 0;
    best_fitness = DBL_MAX;
    std::vector<Frame> null_tip_frames(problem.tip_link_indices.size());
    for(size_t i = 0; i < count; i++) // :224-248
    {
        if(success[i])
        {
            double f;
            if(problem.secondary_goals.empty())
                f = fitness[i];
            else
            {
                std::vector<double> active(problem.active_variables.size());
                for(size_t k = 0; k < active.size(); k++) active[k] = solutions[i][problem.active_variables[k]]; // extractActiveVariables
                f = fitness[i] + problem.computeGoalFitness(problem.secondary_goals, null_tip_frames.data(), active.data());
            }
            if(f < best_fitness)
            {
                best_fitness = f;
                best_index = i;
            }
        }
    }
    if(best_fitness == DBL_MAX) // :252-262
        for(size_t i = 0; i < count; i++)
            if(fitness[i] < best_fitness)
            {
                best_fitness = fitness[i];
                best_index = i;
            }
    return best_index;
}

// src/ik_parallel.h:148-190 with the wall-clock timeout replaced by a step budget: `islands` solvers of one problem run in
// lock step (the reference's threads run freely; lock step is the deterministic reading of the same loop).  After every
// burst of 4 steps each solver records its solution, success flag and fitness (:173-181); query_exit = the `finished` flag
// (:160,180): once any solver has succeeded nobody starts another burst.  island_exit = a solver that succeeded stops by
// itself (the batch contract's per-run early exit).
struct IslandRuns
{
    std::vector<std::vector<double>> solutions;
    std::vector<double> fitness;
    std::vector<int> success, steps;
};
inline IslandRuns solveIslands(const RobotModel& robot, const Tables& tables, const Problem& problem, const uint32_t* rng_seeds, size_t islands, const SolverConfig& cfg, int steps, bool island_exit, bool query_exit,
                               Options opt = Options())
{
    std::vector<std::unique_ptr<IKEvolution2>> solvers;
    for(size_t i = 0; i < islands; i++)
    {
        solvers.emplace_back(new IKEvolution2(&robot, tables, rng_seeds[i], cfg, opt));
        solvers.back()->initialize(problem);
        if(opt.island_stride > 0 && i > 0)
        {
            // The cursors of the table-driven streams (XORShift64 state, fast_random index, gaussian index) advance by the same
            // amounts for every query and seed, so a throw-away clone that takes i * stride steps holds exactly the cursors
            // island i starts from; its minstd_rand stays the freshly seeded one.
            IKEvolution2 ahead(&robot, tables, rng_seeds[i], cfg, opt);
            ahead.initialize(problem);
            for(size_t k = 0; k < i * (size_t)opt.island_stride; k++) ahead.step();
            solvers.back()->_xorshift = ahead._xorshift;
            solvers.back()->random_buffer_index = ahead.random_buffer_index;
            solvers.back()->random_gauss_index = ahead.random_gauss_index;
        }
    }
    IslandRuns r;
    r.solutions.resize(islands), r.fitness.assign(islands, 0.0), r.success.assign(islands, 0), r.steps.assign(islands, 0);
    std::vector<int> stopped(islands, 0);
    int done = 0;
    bool finished = false;
    while(done < steps && !finished)
    {
        int burst = std::min(4, steps - done);
        for(size_t i = 0; i < islands; i++)
        {
            if(stopped[i]) continue;
            for(int k = 0; k < burst; k++) solvers[i]->step();
            r.steps[i] = done + burst;
            std::vector<double> result = solvers[i]->getSolution();
            solvers[i]->model.applyConfiguration(result);
            bool ok = solvers[i]->checkSolution(result, solvers[i]->model.getTipFrames());
            if(ok && (island_exit || query_exit)) stopped[i] = 1;
            if(ok && query_exit) finished = true;
        }
        done += burst;
    }
    for(size_t i = 0; i < islands; i++)
    {
        r.solutions[i] = solvers[i]->getSolution();
        solvers[i]->model.applyConfiguration(r.solutions[i]);
        r.success[i] = solvers[i]->checkSolution(r.solutions[i], solvers[i]->model.getTipFrames());
        r.fitness[i] = solvers[i]->computeFitness(r.solutions[i], solvers[i]->model.getTipFrames());
    }
    return r;
}

// src/kinematics_plugin.cpp:580-611 — angle wrap of the returned state (in place)
inline void wrapAngles(const RobotModel& robot, const Problem& problem, std::vector<double>& state)
{
    bool has_mimic = false;
    for(auto& l : robot.links) has_mimic = has_mimic || l.mimic >= 0;
    for(auto ivar : problem.active_variables)
    {
        auto v = state[ivar];
        int j = robot.var_joint[ivar];
        if(j >= 0 && robot.links[j].joint_type == REVOLUTE && !has_mimic) // :583-584
        {
            auto r = problem.initial_guess[ivar];
            auto lo = problem.modelInfo.getMin(ivar);
            auto hi = problem.modelInfo.getMax(ivar);
            if(r < v - M_PI || r > v + M_PI) // :590-598
            {
                v -= r;
                v /= (2 * M_PI);
                v += 0.5;
                v -= std::floor(v);
                v -= 0.5;
                v *= (2 * M_PI);
                v += r;
            }
            if(v > hi) v -= std::ceil(std::max(0.0, v - hi) / (2 * M_PI)) * (2 * M_PI); // :601-604
            if(v < lo) v += std::ceil(std::max(0.0, lo - v) / (2 * M_PI)) * (2 * M_PI);
            if(v < lo) v = lo; // :607-610
            if(v > hi) v = hi;
        }
        state[ivar] = v;
    }
}

} // namespace bioik_oracle
