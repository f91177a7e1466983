// bioik_host.hpp — pure C++ host logic of libbioik_b200.so (no CUDA runtime calls):
// robot table intake, the host halves of RobotFK_Fast_Base::initialize /
// RobotFK_Jacobian::initialize / RobotInfo / Problem::initialize (flattening into DProblem),
// the reference's RNG lookup tables and the query-independent RNG schedules.
// Shared by bioik_capi.cu and by the test-only host simulation (tests/hostsim).
#pragma once

#include "../../include/bioik_b200.h"
#include "bioik_dev.cuh"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <random>
#include <string>
#include <vector>

namespace bioik
{

struct HostRobot
{
    int n_links = 0, n_vars = 0;
    std::vector<int> parent, jtype, first_var, mimic;
    std::vector<double> origin, axis, mimic_factor, mimic_offset, var_min, var_max, var_vel;
    std::vector<int> var_bounded, var_joint;
    std::vector<double> link_mass, link_com; // URDF inertials (BalanceGoal); empty = none
};

inline int var_count(int t)
{
    switch(t)
    {
    case BIOIK_JOINT_REVOLUTE:
    case BIOIK_JOINT_PRISMATIC: return 1;
    case BIOIK_JOINT_FLOATING: return 7;
    case BIOIK_JOINT_PLANAR: return 3;
    default: return 0;
    }
}


inline int host_fail(std::string& err, int code, const char* msg)
{
    err = msg;
    return code;
}

// BioikRobot -> HostRobot with validation
inline int intake_robot(const BioikRobot* robot, HostRobot& R, std::string& err)
{
    if(robot->n_links < 1 || robot->n_vars < 1) return host_fail(err, BIOIK_E_INVALID, "empty robot");
    R.n_links = robot->n_links;
    R.n_vars = robot->n_vars;
    R.parent.assign(robot->link_parent, robot->link_parent + R.n_links);
    R.jtype.assign(robot->joint_type, robot->joint_type + R.n_links);
    R.first_var.assign(robot->joint_first_var, robot->joint_first_var + R.n_links);
    R.origin.assign(robot->link_origin, robot->link_origin + 7 * R.n_links);
    R.axis.assign(robot->joint_axis, robot->joint_axis + 3 * R.n_links);
    R.mimic.assign(R.n_links, -1);
    R.mimic_factor.assign(R.n_links, 1.0);
    R.mimic_offset.assign(R.n_links, 0.0);
    if(robot->joint_mimic) R.mimic.assign(robot->joint_mimic, robot->joint_mimic + R.n_links);
    if(robot->joint_mimic_factor) R.mimic_factor.assign(robot->joint_mimic_factor, robot->joint_mimic_factor + R.n_links);
    if(robot->joint_mimic_offset) R.mimic_offset.assign(robot->joint_mimic_offset, robot->joint_mimic_offset + R.n_links);
    R.var_min.assign(robot->var_min, robot->var_min + R.n_vars);
    R.var_max.assign(robot->var_max, robot->var_max + R.n_vars);
    R.var_bounded.assign(robot->var_bounded, robot->var_bounded + R.n_vars);
    R.var_vel.assign(robot->var_max_velocity, robot->var_max_velocity + R.n_vars);
    R.link_mass.clear(), R.link_com.clear();
    if(robot->link_mass)
    {
        R.link_mass.assign(robot->link_mass, robot->link_mass + R.n_links);
        R.link_com.assign(3 * (size_t)R.n_links, 0.0);
        if(robot->link_com) R.link_com.assign(robot->link_com, robot->link_com + 3 * (size_t)R.n_links);
    }
    R.var_joint.assign(R.n_vars, -1);
    for(int l = 0; l < R.n_links; l++)
    {
        if(R.parent[l] >= l) return host_fail(err, BIOIK_E_INVALID, "links must be ordered parents before children");
        if(R.mimic[l] >= R.n_links) return host_fail(err, BIOIK_E_INVALID, "mimic joint out of range");
        int cnt = var_count(R.jtype[l]);
        for(int k = 0; k < cnt; k++)
        {
            int v = R.first_var[l] + k;
            if(v < 0 || v >= R.n_vars) return host_fail(err, BIOIK_E_INVALID, "joint variable index out of range");
            R.var_joint[v] = l;
        }
    }
    return BIOIK_OK;
}

// Random::Random (src/ik_base.h:118-125): one engine fills the uniform table, then the gauss table.
inline void make_tables(uint32_t seed, std::vector<double>& uniform, std::vector<double>& gauss)
{
    const size_t size = 1024 * 1024 * 8;
    std::minstd_rand rng(seed);
    std::normal_distribution<double> normal_distribution;
    uniform.resize(size);
    for(auto& r : uniform) r = std::uniform_real_distribution<double>(0, 1)(rng);
    gauss.resize(size);
    for(auto& r : gauss) r = normal_distribution(rng);
}

struct XORShift64 // src/utils.h:369-385
{
    uint64_t v = 88172645463325252ull;
    uint64_t operator()()
    {
        v ^= v << 13;
        v ^= v >> 7;
        v ^= v << 17;
        return v;
    }
};

// The reproduce() calls of a query consume fast_random_gauss_n / fast_random_index in a fixed,
// query-independent order (src/ik_evolution_2.cpp:254-265): tabulate slab starts and rate exponents.
inline void make_schedules(int steps, int gens, int C, int n, std::vector<int32_t>& gauss_off, std::vector<uint8_t>& rate_exp)
{
    const size_t size = 1024 * 1024 * 8;
    XORShift64 x;
    x();                        // random_buffer_index  (src/ik_base.h:122)
    size_t gauss_index = x();   // random_gauss_index   (src/ik_base.h:124)
    size_t s = (size_t)(C - 2) * n + (size_t)C * 4 + 4; // :254
    int calls = steps * 2 * gens;
    gauss_off.resize(calls);
    rate_exp.resize((size_t)calls * (C - 2));
    for(int c = 0; c < calls; c++)
    {
        size_t i = gauss_index; // fast_random_gauss_n, src/ik_base.h:110-116
        gauss_index += s;
        if(gauss_index >= size) i = 0, gauss_index = s;
        gauss_off[c] = (int32_t)i;
        for(int k = 0; k < C - 2; k++) rate_exp[(size_t)c * (C - 2) + k] = (uint8_t)(x() % 16); // :265
    }
}

inline int build_problem(const HostRobot& R, const BioikProblem* p, DProblem& P, std::string& err)
{
    memset(&P, 0, sizeof(P));
    if(p->n_tips < 1 || p->n_active < 1 || p->n_goals < 1) return host_fail(err, BIOIK_E_INVALID, "problem needs at least one tip, one active variable and one goal");
    if(p->n_tips > MAX_TIPS || p->n_active > MAX_GENES || p->n_goals > MAX_GOALS || R.n_vars > MAX_VARS)
        return host_fail(err, BIOIK_E_LIMIT, "problem exceeds compiled-in capacity (tips<=24, genes<=48, goals<=24, vars<=64)");
    P.n_vars = R.n_vars;
    P.n = p->n_active;
    P.T = p->n_tips;
    P.G = p->n_goals;
    // link schedule: chains root->tip, de-duplicated (src/forward_kinematics.h:268-282)
    std::vector<int> schedule, slot_of_link(R.n_links, -1);
    for(int t = 0; t < p->n_tips; t++)
    {
        int tip = p->tip_links[t];
        if(tip < 0 || tip >= R.n_links) return host_fail(err, BIOIK_E_INVALID, "tip link out of range");
        std::vector<int> chain;
        for(int l = tip; l >= 0; l = R.parent[l]) chain.push_back(l);
        std::reverse(chain.begin(), chain.end());
        for(int l : chain)
            if(slot_of_link[l] < 0)
            {
                slot_of_link[l] = (int)schedule.size();
                schedule.push_back(l);
            }
    }
    if((int)schedule.size() > MAX_SLOTS) return host_fail(err, BIOIK_E_LIMIT, "link schedule longer than 96 links");
    P.L = (int)schedule.size();
    for(int s = 0; s < P.L; s++)
    {
        int l = schedule[s];
        DSlot& S = P.slots[s];
        S.parent = R.parent[l] >= 0 ? slot_of_link[R.parent[l]] : -1;
        S.type = R.jtype[l];
        S.var = R.first_var[l];
        S.tipmask = 0;
        for(int k = 0; k < 7; k++) S.origin[k] = R.origin[7 * l + k];
        for(int k = 0; k < 3; k++) S.axis[k] = R.axis[3 * l + k];
    }
    // tip_dependencies (src/forward_kinematics.h:588-598)
    for(int t = 0; t < p->n_tips; t++)
    {
        P.tip_slot[t] = slot_of_link[p->tip_links[t]];
        for(int l = p->tip_links[t]; l >= 0; l = R.parent[l]) P.slots[slot_of_link[l]].tipmask |= (1 << t);
    }
    // joint_dependencies (src/forward_kinematics.h:570-587), keyed by the joint's child link
    std::vector<std::vector<int>> deps(R.n_links);
    for(int l : schedule) deps[l].push_back(l);
    for(int l : schedule)
    {
        int m = R.mimic[l];
        if(m >= 0)
        {
            int hops = 0;
            while(R.mimic[m] >= 0 && R.mimic[m] != l)
            {
                m = R.mimic[m];
                if(++hops > R.n_links) return host_fail(err, BIOIK_E_INVALID, "mimic joints form a cycle");
            }
            deps[m].push_back(l);
        }
    }
    // mimic list in model order (src/forward_kinematics.h:230-246)
    P.n_mimic = 0;
    for(int l = 0; l < R.n_links; l++)
        if(R.mimic[l] >= 0)
        {
            if(P.n_mimic >= MAX_VARS) return host_fail(err, BIOIK_E_LIMIT, "more than 64 mimic joints");
            if(var_count(R.jtype[l]) < 1 || var_count(R.jtype[R.mimic[l]]) < 1) return host_fail(err, BIOIK_E_INVALID, "a mimic joint and the joint it mimics need a variable each");
            DMimic& M = P.mimics[P.n_mimic++];
            M.dest = R.first_var[l];
            M.src = R.first_var[R.mimic[l]];
            M.factor = R.mimic_factor[l];
            M.offset = R.mimic_offset[l];
        }
    // RobotInfo (include/bio_ik/robot_info.h:70-105) + velocity weights (src/problem.cpp:206-225)
    for(int v = 0; v < R.n_vars; v++) P.gene_of_var[v] = -1;
    double rcp_sum = 0;
    std::vector<double> rcp(p->n_active);
    int ndep = 0;
    for(int i = 0; i < p->n_active; i++)
    {
        int v = p->active_vars[i];
        if(v < 0 || v >= R.n_vars) return host_fail(err, BIOIK_E_INVALID, "active variable out of range");
        if(P.gene_of_var[v] >= 0) return host_fail(err, BIOIK_E_INVALID, "duplicate active variable");
        P.gene_of_var[v] = i;
        DGene& Gn = P.genes[i];
        Gn.var = v;
        Gn.var_in_joint = R.var_joint[v] >= 0 ? v - R.first_var[R.var_joint[v]] : 0;
        bool bounded = R.var_bounded[v] != 0;
        int j = R.var_joint[v];
        if(j >= 0 && R.jtype[j] == BIOIK_JOINT_REVOLUTE)
            if(R.var_max[v] - R.var_min[v] >= 2 * M_PI * 0.9999) bounded = false;
        Gn.vmin = R.var_min[v];
        Gn.vmax = R.var_max[v];
        Gn.clip_min = bounded ? Gn.vmin : -DBL_MAX;
        Gn.clip_max = bounded ? Gn.vmax : +DBL_MAX;
        Gn.span = Gn.vmax - Gn.vmin;
        if(!(Gn.span >= 0 && Gn.span < FLT_MAX)) Gn.span = 1;
        double vel = R.var_vel[v];
        rcp[i] = vel > 0.0 ? 1.0 / vel : 0.0;
        // dependencies of the variable's joint; none if that joint itself mimics another (:623)
        Gn.dep_start = ndep;
        Gn.dep_count = 0;
        Gn.tipmask = 0;
        if(j >= 0 && R.mimic[j] < 0)
            for(int jl : deps[j])
            {
                double scale = 1;
                int hops = 0;
                for(int m = jl; R.mimic[m] >= 0 && R.mimic[m] != jl && hops <= R.n_links; m = R.mimic[m], hops++) scale *= R.mimic_factor[m];
                P.dep_slot[ndep] = slot_of_link[jl];
                P.dep_scale[ndep] = scale;
                Gn.tipmask |= P.slots[slot_of_link[jl]].tipmask;
                ndep++;
                Gn.dep_count++;
            }
    }
    {
        // quaternion genes of floating joints (src/ik_evolution_2.cpp:118-126): the gene of rot_x; reproduce() normalises the four
        // doubles starting there, so the joint's rotation variables must be consecutive genes
        P.n_quat = 0;
        for(int i = 0; i < p->n_active; i++)
        {
            const int v = p->active_vars[i], j = R.var_joint[v];
            if(j < 0 || R.jtype[j] != BIOIK_JOINT_FLOATING || R.first_var[j] + 3 != v) continue;
            for(int k = 1; k < 4; k++)
                if(i + k >= p->n_active || p->active_vars[i + k] != v + k) return host_fail(err, BIOIK_E_UNSUPPORTED_JOINT, "the rotation variables of a floating joint must be consecutive active variables");
            P.quat_gene[P.n_quat++] = i;
        }
    }
    {
        // per-tip gene lists of the tip-major generation kernel
        int k = 0;
        for(int t = 0; t < P.T; t++)
        {
            P.tip_gene_start[t] = k;
            for(int i = 0; i < p->n_active; i++)
                if((P.genes[i].tipmask >> t) & 1) P.tip_gene[k++] = (int16_t)i;
        }
        for(int t = P.T; t <= MAX_TIPS; t++) P.tip_gene_start[t] = k;
    }
    {
        // kinematics_plugin.cpp:583-584: the angle wrap applies to revolute variables of robots without mimic joints
        bool any_mimic = false;
        for(int l = 0; l < R.n_links; l++) any_mimic = any_mimic || R.mimic[l] >= 0;
        for(int i = 0; i < p->n_active; i++)
        {
            int j = R.var_joint[p->active_vars[i]];
            P.wrap_gene[i] = (!any_mimic && j >= 0 && R.jtype[j] == BIOIK_JOINT_REVOLUTE) ? 1 : 0;
        }
    }
    for(int i = 0; i < p->n_active; i++) rcp_sum += rcp[i];
    for(int i = 0; i < p->n_active; i++) P.genes[i].vel_weight = rcp_sum > 0 ? rcp[i] / rcp_sum : 1.0 / p->n_active;
    // goals
    P.has_secondary = 0;
    for(int g = 0; g < p->n_goals; g++)
    {
        const BioikGoal& bg = p->goals[g];
        if(bg.type < BIOIK_GOAL_POSITION || bg.type > BIOIK_GOAL_BALANCE) return host_fail(err, BIOIK_E_UNSUPPORTED_GOAL, "goal type has no device implementation (callback / FCL goals stay on the CPU solver)");
        DGoal& D = P.goals[g];
        D.type = bg.type;
        D.tip = bg.tip;
        if(D.tip < 0 || D.tip >= p->n_tips) return host_fail(err, BIOIK_E_INVALID, "goal tip index out of range");
        D.secondary = bg.secondary ? 1 : 0;
        D.weight_sq = bg.weight * bg.weight;
        D.var_index = 0;
        if(bg.type == BIOIK_GOAL_JOINT_VARIABLE)
        {
            if(bg.var < 0 || bg.var >= R.n_vars) return host_fail(err, BIOIK_E_INVALID, "goal variable out of range");
            D.var_index = P.gene_of_var[bg.var] >= 0 ? P.gene_of_var[bg.var] : -1 - bg.var;
        }
        if(D.secondary) P.has_secondary = 1;
        if(bg.type == BIOIK_GOAL_BALANCE && P.n_balance == 0)
        {
            // BalanceGoal::describe (src/goal_types.cpp:231-259): every link whose inertial has a positive mass, in link order, with
            // weight = mass / total (total summed in the same order); each of them must be a tip link of the problem
            double total = 0.0;
            for(int l = 0; l < R.n_links && !R.link_mass.empty(); l++)
            {
                const double mass = R.link_mass[l];
                if(!(mass > 0)) continue;
                if(P.n_balance >= MAX_TIPS) return host_fail(err, BIOIK_E_LIMIT, "BalanceGoal: more than 24 links with mass");
                if(slot_of_link[l] < 0) return host_fail(err, BIOIK_E_INVALID, "BalanceGoal: every link with mass must be a tip link of the problem");
                int tip = -1;
                for(int t = 0; t < p->n_tips; t++)
                    if(p->tip_links[t] == l) tip = t;
                if(tip < 0) return host_fail(err, BIOIK_E_INVALID, "BalanceGoal: every link with mass must be a tip link of the problem");
                P.balance_tip[P.n_balance] = tip;
                for(int k = 0; k < 3; k++) P.balance_center[P.n_balance][k] = R.link_com[3 * (size_t)l + k];
                P.balance_weight[P.n_balance] = mass;
                total += mass;
                P.n_balance++;
            }
            if(P.n_balance == 0) return host_fail(err, BIOIK_E_INVALID, "BalanceGoal needs BioikRobot::link_mass (no link has mass)");
            for(int i = 0; i < P.n_balance; i++) P.balance_weight[i] /= total;
        }
        if(bg.type >= BIOIK_GOAL_AVOID_JOINT_LIMITS && bg.type <= BIOIK_GOAL_JOINT_VARIABLE) P.n_joint_goals++;
    }
    // thresholds (src/problem.cpp:90-95)
    P.dpos = p->dpos;
    P.drot = p->drot;
    P.dtwist = p->dtwist;
    if(P.dpos < 0.0 || P.dpos >= FLT_MAX || !std::isfinite(P.dpos)) P.dpos = DBL_MAX;
    if(P.drot < 0.0 || P.drot >= FLT_MAX || !std::isfinite(P.drot)) P.drot = DBL_MAX;
    if(P.dtwist < 0.0 || P.dtwist >= FLT_MAX || !std::isfinite(P.dtwist)) P.dtwist = DBL_MAX;
    return BIOIK_OK;
}


} // namespace bioik
