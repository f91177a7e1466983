// C++ test of the header-only host API (bio_ik_b200/host/bioik_host_api.hpp) over libbioik_b200.so.
// Reads like a use of the reference: construct goals by class name, initialise a Problem, create a solver by its
// IKFactory name, solve.  Without a GPU it checks the flattening and that creation FAILS LOUDLY (no CPU fallback);
// with `gpu` as argv[1] it solves a batch and verifies FK(solution) against the goals.
#include "../../bio_ik_b200/host/bioik_host_api.hpp"

#include <cstdio>
#include <cstring>
#include <random>

using namespace bio_ik;

#define CHECK(c)                                                   \
    do                                                             \
    {                                                              \
        if(!(c))                                                   \
        {                                                          \
            std::printf("CHECK failed line %d: %s\n", __LINE__, #c); \
            return 1;                                              \
        }                                                          \
    } while(0)

static RobotModel makeArm()
{
    RobotModel rm;
    auto link = [&](const char* name, const char* parent, int type, double x, double y, double z, double ax, double ay, double az, double lo, double hi, bool bounded = true) {
        RobotModel::Link l;
        l.name = name, l.parent = parent, l.joint_name = std::string(name) + "_joint", l.joint_type = type;
        l.origin[0] = x, l.origin[1] = y, l.origin[2] = z;
        l.axis[0] = ax, l.axis[1] = ay, l.axis[2] = az;
        l.lower = lo, l.upper = hi, l.bounded = bounded, l.velocity = 2.0;
        rm.addLink(l);
    };
    link("base", "", BIOIK_JOINT_FIXED, 0, 0, 0, 0, 0, 1, 0, 0);
    link("torso", "base", BIOIK_JOINT_PRISMATIC, 0, 0, 0.7, 0, 0, 1, 0, 0.3);
    link("pan", "torso", BIOIK_JOINT_REVOLUTE, 0, -0.2, 0, 0, 0, 1, -2.1, 0.6);
    link("lift", "pan", BIOIK_JOINT_REVOLUTE, 0.1, 0, 0, 0, 1, 0, -0.35, 1.3);
    link("roll", "lift", BIOIK_JOINT_REVOLUTE, 0, 0, 0, 1, 0, 0, -3.7, 0.6);
    link("elbow", "roll", BIOIK_JOINT_REVOLUTE, 0.4, 0, 0, 0, 1, 0, -2.1, -0.15);
    link("froll", "elbow", BIOIK_JOINT_REVOLUTE, 0, 0, 0, 1, 0, 0, -3.14159265358979, 3.14159265358979, false);
    link("wflex", "froll", BIOIK_JOINT_REVOLUTE, 0.32, 0, 0, 0, 1, 0, -2.0, -0.1);
    link("wroll", "wflex", BIOIK_JOINT_REVOLUTE, 0, 0, 0, 1, 0, 0, -3.14159265358979, 3.14159265358979, false);
    rm.finalize();
    return rm;
}

int main(int argc, char** argv)
{
    const bool gpu = argc > 1 && !std::strcmp(argv[1], "gpu");
    RobotModel rm = makeArm();
    CHECK(rm.getVariableCount() == 8);
    JointModelGroup arm{"arm", {"pan_joint", "lift_joint", "roll_joint", "elbow_joint", "froll_joint", "wflex_joint", "wroll_joint"}, {"wroll"}};

    PoseGoal pose("wroll", Vector3(0.5, -0.3, 0.9), Quaternion(0, 0, 0.5, 2.0));
    MinimalDisplacementGoal mind(0.5);
    JointVariableGoal jv("elbow_joint", -1.0, 0.25);
    Problem problem;
    problem.initialize(rm, arm, {&pose, &mind, &jv});
    // goal-named variables come first (src/problem.cpp:145-148), then the active subtree; the torso is not in the group
    CHECK(problem.active_variables.size() == 7);
    CHECK(rm.variable_names[problem.active_variables[0]] == "elbow_joint");
    CHECK(problem.tip_link_indices.size() == 1 && problem.goals.size() == 3);
    CHECK(problem.goals[0].type == BIOIK_GOAL_POSE && problem.goals[0].p[7] == 0.5);
    double qn = 0;
    for(int k = 3; k < 7; k++) qn += problem.goals[0].p[k] * problem.goals[0].p[k];
    CHECK(std::fabs(qn - 1.0) < 1e-15); // constructor normalises the orientation (goal_types.h:139)
    CHECK(problem.goals[1].secondary == 1 && problem.goals[2].secondary == 0 && problem.goals[2].var == rm.variable_index.at("elbow_joint"));
    bool threw = false;
    try
    {
        PositionGoal bad("no_such_link", Vector3());
        Problem p2;
        p2.initialize(rm, arm, {&bad});
    }
    catch(std::runtime_error&)
    {
        threw = true;
    }
    CHECK(threw);
    threw = false;
    try
    {
        IKSolverB200 s("bio1", rm);
    }
    catch(std::runtime_error&)
    {
        threw = true;
    }
    CHECK(threw); // IKFactory: class not found
    {
        // src/problem.cpp:103-126: a goal variable outside the joint group is an error; a fixed joint's variable stays at the seed
        threw = false;
        try
        {
            JointVariableGoal torso("torso_joint", 0.1);
            Problem p3;
            p3.initialize(rm, arm, {&pose, &torso});
        }
        catch(std::runtime_error&)
        {
            threw = true;
        }
        CHECK(threw);
        Problem p4;
        p4.initialize(rm, arm, {&pose}, {"froll_joint"});
        CHECK(p4.active_variables.size() == 6);
        for(int v : p4.active_variables) CHECK(rm.variable_names[v] != "froll_joint");
    }
    {
        // floating base: 7 variables with MoveIt's names, all of them active; links with mass become tips of a BalanceGoal
        RobotModel fb;
        RobotModel::Link w, b, l1, l2;
        w.name = "world", w.joint_name = "world_joint";
        b.name = "base", b.parent = "world", b.joint_name = "virtual_joint", b.joint_type = BIOIK_JOINT_FLOATING, b.mass = 5.0, b.com[2] = 0.1;
        l1.name = "l1", l1.parent = "base", l1.joint_name = "j1", l1.joint_type = BIOIK_JOINT_REVOLUTE, l1.origin[2] = 0.25, l1.lower = -2.5, l1.upper = 2.5, l1.mass = 1.0, l1.com[0] = 0.1;
        l2.name = "l2", l2.parent = "l1", l2.joint_name = "j2", l2.joint_type = BIOIK_JOINT_REVOLUTE, l2.origin[0] = 0.3, l2.axis[0] = 0, l2.axis[1] = 1, l2.axis[2] = 0, l2.lower = -1.8, l2.upper = 1.8;
        fb.addLink(w), fb.addLink(b), fb.addLink(l1), fb.addLink(l2);
        fb.finalize();
        CHECK(fb.getVariableCount() == 9 && fb.variable_names[3] == "virtual_joint/rot_x" && fb.var_bounded[3] == 1 && fb.var_bounded[0] == 0);
        JointModelGroup all{"all", {"virtual_joint", "j1", "j2"}, {"l2"}};
        PositionGoal tip("l2", Vector3(0.3, 0, 0.5));
        BalanceGoal bal(Vector3(0, 0, 0), 0.5);
        Problem pb;
        pb.initialize(fb, all, {&tip, &bal});
        CHECK(pb.active_variables.size() == 9);
        CHECK(pb.tip_link_indices.size() == 3 && pb.tip_link_indices[0] == 3 && pb.tip_link_indices[1] == 1 && pb.tip_link_indices[2] == 2);
        CHECK(pb.goals[1].type == BIOIK_GOAL_BALANCE && pb.goals[1].tip == 1 && pb.goals[1].p[5] == 1.0);
        BioikRobot abi = fb.toABI();
        CHECK(abi.link_mass && abi.link_mass[1] == 5.0 && abi.link_com[3 * 1 + 2] == 0.1);
    }

    if(!gpu)
    {
        // no device: the product must refuse to run, not fall back to a CPU path
        try
        {
            IKSolverB200 s("bio2_memetic", rm);
            std::printf("note: a CUDA device is present; run with `gpu` for the solve test\n");
        }
        catch(std::runtime_error& e)
        {
            CHECK(std::strstr(e.what(), "CUDA") != nullptr);
        }
        std::printf("host api ok (no-gpu leg)\n");
        return 0;
    }

    // GPU leg: FK -> IK -> FK round trip through the C++ API
    const int B = 256, n_vars = 8;
    IKSolverB200 solver("bio2_memetic", rm, 64, 1, 0);
    PoseGoal only_pose("wroll", Vector3(), Quaternion());
    Problem pr;
    pr.initialize(rm, arm, {&only_pose});
    solver.initialize(pr);
    std::mt19937 gen(7);
    std::vector<double> targets(B * n_vars, 0.0), seeds(B * n_vars, 0.0);
    for(int b = 0; b < B; b++)
        for(int v : pr.active_variables)
        {
            std::uniform_real_distribution<double> u(rm.var_min[v], rm.var_max[v]);
            targets[b * n_vars + v] = u(gen), seeds[b * n_vars + v] = u(gen);
        }
    std::vector<double> tips = solver.forwardKinematics(targets, 1);
    std::vector<double> gp(B * BIOIK_GOAL_NPARAM, 0.0);
    for(int b = 0; b < B; b++)
    {
        for(int k = 0; k < 7; k++) gp[b * BIOIK_GOAL_NPARAM + k] = tips[b * 7 + k];
        gp[b * BIOIK_GOAL_NPARAM + 7] = 0.5;
    }
    std::vector<uint32_t> rs(B);
    for(int b = 0; b < B; b++) rs[b] = 1 + b;
    auto res = solver.solveBatch(gp, seeds, rs, 25);
    int ok = 0;
    std::vector<double> reached = solver.forwardKinematics(res.solutions, 1);
    for(int b = 0; b < B; b++)
        if(res.success[b])
        {
            ok++;
            for(int k = 0; k < 3; k++) CHECK(std::fabs(reached[b * 7 + k] - tips[b * 7 + k]) < 2e-5);
        }
    std::printf("host api ok (gpu leg): %d / %d solved\n", ok, B);
    CHECK(ok > B * 8 / 10);
    // one query at a time, 16 islands each, the reference driver's early exit and the plugin's angle wrap
    {
        const int Q = 8, islands = 16;
        std::vector<double> qgp(gp.begin(), gp.begin() + Q * BIOIK_GOAL_NPARAM), qseeds(seeds.begin(), seeds.begin() + Q * n_vars);
        std::vector<uint32_t> irs(Q * islands);
        for(size_t i = 0; i < irs.size(); i++) irs[i] = 100 + (uint32_t)i;
        auto ir = solver.solveIslands(qgp, qseeds, islands, irs, 25);
        std::vector<double> at = solver.forwardKinematics(ir.solutions, 1);
        int iok = 0;
        for(int q = 0; q < Q; q++)
        {
            CHECK(ir.island[q] >= 0 && ir.island[q] < islands);
            if(!ir.success[q]) continue;
            iok++;
            for(int k = 0; k < 3; k++) CHECK(std::fabs(at[q * 7 + k] - tips[q * 7 + k]) < 2e-5); // the wrapped angles are the same pose
        }
        std::printf("host api ok (islands): %d / %d solved\n", iok, Q);
        CHECK(iok >= Q - 1);
        // the same through the resumable interface: begin, step() one at a time, getSolution - the same bits
        solver.begin(qgp, qseeds, islands, irs, 0, 0);
        for(int k = 0; k < 6; k++) solver.step(1);
        auto one = solver.getSolution(false);
        auto whole = solver.solveIslands(qgp, qseeds, islands, irs, 6, 0, false);
        CHECK(one.solutions == whole.solutions && one.island == whole.island);
    }
    return 0;
}
