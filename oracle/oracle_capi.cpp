// oracle_capi.cpp — C entry points of the CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
// Loaded with ctypes by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs only.  Takes the same flattened POD tables as the product's C ABI
// (include/bioik_b200.h) so both sides are driven by identical inputs.
#include "../include/bioik_b200.h"
#include "bioik_oracle.hpp"

#include <atomic>
#include <memory>
#include <thread>

using namespace bioik_oracle;

namespace
{
thread_local std::string g_error;

RobotModel makeRobot(const BioikRobot* r)
{
    RobotModel m;
    m.links.resize(r->n_links);
    m.n_vars = r->n_vars;
    for(int l = 0; l < r->n_links; l++)
    {
        auto& L = m.links[l];
        L.parent = r->link_parent[l];
        L.joint_type = r->joint_type[l];
        L.first_var = r->joint_first_var[l];
        const double* o = r->link_origin + 7 * l;
        L.origin = Frame(Vec3(o[0], o[1], o[2]), Quat(o[3], o[4], o[5], o[6]));
        L.axis = Vec3(r->joint_axis[3 * l], r->joint_axis[3 * l + 1], r->joint_axis[3 * l + 2]);
        L.mimic = r->joint_mimic ? r->joint_mimic[l] : -1;
        L.mimic_factor = r->joint_mimic_factor ? r->joint_mimic_factor[l] : 1.0;
        L.mimic_offset = r->joint_mimic_offset ? r->joint_mimic_offset[l] : 0.0;
    }
    m.var_min.assign(r->var_min, r->var_min + r->n_vars);
    m.var_max.assign(r->var_max, r->var_max + r->n_vars);
    m.var_bounded.assign(r->var_bounded, r->var_bounded + r->n_vars);
    m.var_max_velocity.assign(r->var_max_velocity, r->var_max_velocity + r->n_vars);
    if(r->link_mass)
    {
        m.link_mass.assign(r->link_mass, r->link_mass + r->n_links);
        m.link_com.assign(3 * (size_t)r->n_links, 0.0);
        if(r->link_com) m.link_com.assign(r->link_com, r->link_com + 3 * (size_t)r->n_links);
    }
    m.finalize();
    return m;
}

Problem makeProblem(const RobotModel& robot, const BioikProblem* p)
{
    Problem pr;
    pr.robot_model = &robot;
    pr.modelInfo = RobotInfo(robot);
    for(int i = 0; i < p->n_tips; i++) pr.tip_link_indices.push_back(p->tip_links[i]);
    for(int i = 0; i < p->n_active; i++) pr.active_variables.push_back(p->active_vars[i]);
    for(int g = 0; g < p->n_goals; g++)
    {
        const BioikGoal& bg = p->goals[g];
        GoalInfo gi;
        gi.type = bg.type;
        gi.tip_index = bg.tip;
        gi.secondary = bg.secondary != 0;
        gi.weight = bg.weight;
        gi.weight_sq = bg.weight * bg.weight;
        gi.var_index = 0;
        if(bg.type == BIOIK_GOAL_JOINT_VARIABLE)
        {
            gi.var_index = -1 - (long)bg.var;
            for(int i = 0; i < p->n_active; i++)
                if(p->active_vars[i] == bg.var) gi.var_index = i;
        }
        for(int k = 0; k < GOAL_NPARAM; k++) gi.p[k] = bg.p[k];
        if(bg.type == BIOIK_GOAL_BALANCE && pr.balance_infos.empty()) pr.describeBalance();
        if(gi.secondary)
            pr.secondary_goals.push_back(gi);
        else
            pr.goals.push_back(gi);
    }
    pr.dpos = p->dpos;
    pr.drot = p->drot;
    pr.dtwist = p->dtwist;
    pr.sanitizeThresholds();
    pr.initVelocityWeights();
    return pr;
}

// per-query parameters: goal_params [n_goals][NPARAM] in BioikProblem goal order
void applyQuery(Problem& pr, const BioikProblem* p, const double* goal_params, const double* seed, size_t n_vars)
{
    pr.initial_guess.assign(seed, seed + n_vars);
    if(!goal_params) return;
    size_t ip = 0, is = 0;
    for(int g = 0; g < p->n_goals; g++)
    {
        GoalInfo& gi = p->goals[g].secondary ? pr.secondary_goals[is++] : pr.goals[ip++];
        for(int k = 0; k < GOAL_NPARAM; k++) gi.p[k] = goal_params[g * GOAL_NPARAM + k];
    }
}

SolverConfig makeCfg(const BioikSolverCfg* c)
{
    SolverConfig s;
    s.population = c->population;
    s.generations = c->generations;
    s.memetic = c->memetic;
    s.memetic_iters = c->memetic_iters;
    return s;
}

template <class F> void parallelFor(int n, int nthreads, F f)
{
    if(nthreads <= 1 || n <= 1)
    {
        for(int i = 0; i < n; i++) f(i);
        return;
    }
    std::atomic<int> next(0);
    std::vector<std::thread> pool;
    for(int t = 0; t < nthreads; t++)
        pool.emplace_back([&]() {
            for(;;)
            {
                int i = next.fetch_add(1);
                if(i >= n) break;
                f(i);
            }
        });
    for(auto& t : pool) t.join();
}
} // namespace

extern "C" {

const char* oracle_last_error() { return g_error.c_str(); }

// ---- lookup tables (src/ik_base.h:118-125) -------------------------------------------------
void* oracle_tables_create(uint32_t seed) { return new Tables(seed); }
void oracle_tables_destroy(void* t) { delete(Tables*)t; }
const double* oracle_tables_uniform(void* t) { return ((Tables*)t)->uniform.data(); }
const double* oracle_tables_gauss(void* t) { return ((Tables*)t)->gauss.data(); }

// ---- known-answer probes ---------------------------------------------------------------------
void oracle_xorshift(int n, uint64_t* out)
{
    XORShift64 x;
    for(int i = 0; i < n; i++) out[i] = x();
}
void oracle_minstd_uniform(uint32_t seed, int n, double* out)
{
    std::minstd_rand rng(seed);
    for(int i = 0; i < n; i++) out[i] = std::uniform_real_distribution<double>(0, 1)(rng);
}
void oracle_minstd_normal(uint32_t seed, int n, double* out)
{
    std::minstd_rand rng(seed);
    std::normal_distribution<double> nd;
    for(int i = 0; i < n; i++) out[i] = nd(rng);
}
void oracle_minstd_index(uint32_t seed, uint64_t s, int n, uint64_t* out)
{
    std::minstd_rand rng(seed);
    for(int i = 0; i < n; i++) out[i] = std::uniform_int_distribution<size_t>(0, s - 1)(rng);
}
void oracle_acos(int n, const double* x, double* out)
{
    for(int i = 0; i < n; i++) out[i] = det_acos(x[i]);
}
void oracle_sincos(int n, const double* x, double* s, double* c)
{
    for(int i = 0; i < n; i++) det_sincos(x[i], s + i, c + i);
}

// ---- frame algebra (include/bio_ik/frame.h), frames as 7 doubles px py pz qx qy qz qw --------
static Frame F7(const double* f) { return Frame(Vec3(f[0], f[1], f[2]), Quat(f[3], f[4], f[5], f[6])); }
static void W7(const Frame& f, double* o)
{
    o[0] = f.pos.x, o[1] = f.pos.y, o[2] = f.pos.z, o[3] = f.rot.x, o[4] = f.rot.y, o[5] = f.rot.z, o[6] = f.rot.w;
}
void oracle_concat(const double* a, const double* b, double* r)
{
    Frame o;
    concat(F7(a), F7(b), o);
    W7(o, r);
}
void oracle_invert(const double* a, double* r)
{
    Frame o;
    invert(F7(a), o);
    W7(o, r);
}
void oracle_change(const double* a, const double* b, const double* c, double* r)
{
    Frame o;
    change(F7(a), F7(b), F7(c), o);
    W7(o, r);
}

// ---- exact FK, approximator, approximate fitness ----------------------------------------------
int oracle_fk_batch(const BioikRobot* robot, const BioikProblem* problem, int libm_sincos, int B, const double* variables, double* out_tip_frames, double* out_link_frames)
{
    try
    {
        RobotModel rm = makeRobot(robot);
        Problem pr = makeProblem(rm, problem);
        Options opt;
        opt.libm_sincos = libm_sincos != 0;
        RobotFK fk(&rm, opt);
        fk.initialize(pr.tip_link_indices);
        size_t T = pr.tip_link_indices.size();
        for(int b = 0; b < B; b++)
        {
            std::vector<double> v(variables + (size_t)b * rm.n_vars, variables + (size_t)(b + 1) * rm.n_vars);
            fk.applyConfiguration(v);
            for(size_t t = 0; t < T; t++) W7(fk.tip_frames[t], out_tip_frames + ((size_t)b * T + t) * 7);
            if(out_link_frames)
                for(size_t l = 0; l < rm.links.size(); l++) W7(fk.global_frames[l], out_link_frames + ((size_t)b * rm.links.size() + l) * 7);
        }
        return 0;
    }
    catch(std::exception& e)
    {
        g_error = e.what();
        return 1;
    }
}

// deviation-study switch for the component entry points below: bit0 = libm sincos (what the reference calls)
static int g_component_flags = 0;
void oracle_set_component_flags(int flags) { g_component_flags = flags; }
static Options componentOptions()
{
    Options opt;
    opt.libm_sincos = (g_component_flags & 1) != 0;
    opt.fma_approx = (g_component_flags & 2) == 0;
    opt.fma_approx1 = (g_component_flags & 6) == 0;
    return opt;
}

// delta frames [B][T][n_active][7]; mask [B][T][n_active] (may be NULL); jacobian [B][6T][n_active] (may be NULL)
int oracle_approx_batch(const BioikRobot* robot, const BioikProblem* problem, int B, const double* variables, double* out_delta, int32_t* out_mask, double* out_jacobian)
{
    try
    {
        RobotModel rm = makeRobot(robot);
        Problem pr = makeProblem(rm, problem);
        RobotFK fk(&rm, componentOptions());
        fk.initialize(pr.tip_link_indices);
        size_t T = pr.tip_link_indices.size(), n = pr.active_variables.size();
        for(int b = 0; b < B; b++)
        {
            std::vector<double> v(variables + (size_t)b * rm.n_vars, variables + (size_t)(b + 1) * rm.n_vars);
            fk.applyConfiguration(v);
            fk.initializeMutationApproximator(pr.active_variables);
            for(size_t t = 0; t < T; t++)
                for(size_t i = 0; i < n; i++)
                {
                    W7(fk.mutation_approx_frames[t][pr.active_variables[i]], out_delta + (((size_t)b * T + t) * n + i) * 7);
                    if(out_mask) out_mask[((size_t)b * T + t) * n + i] = fk.mutation_approx_mask[t][pr.active_variables[i]];
                }
            if(out_jacobian)
                for(size_t r = 0; r < 6 * T; r++)
                    for(size_t c = 0; c < n; c++) out_jacobian[((size_t)b * 6 * T + r) * n + c] = fk.jac(r, c);
        }
        return 0;
    }
    catch(std::exception& e)
    {
        g_error = e.what();
        return 1;
    }
}

int oracle_approx_fitness_batch(const BioikRobot* robot, const BioikProblem* problem, int B, int M, const double* goal_params, const double* seeds, const double* base_variables, const double* genotypes, double* out_primary,
                                double* out_secondary)
{
    try
    {
        RobotModel rm = makeRobot(robot);
        Problem pr0 = makeProblem(rm, problem);
        size_t n = pr0.active_variables.size();
        RobotFK fk(&rm, componentOptions());
        fk.initialize(pr0.tip_link_indices);
        std::vector<Frame> null_frames(pr0.tip_link_indices.size());
        for(int b = 0; b < B; b++)
        {
            Problem pr = pr0;
            applyQuery(pr, problem, goal_params ? goal_params + (size_t)b * problem->n_goals * GOAL_NPARAM : nullptr, seeds + (size_t)b * rm.n_vars, rm.n_vars);
            std::vector<double> v(base_variables + (size_t)b * rm.n_vars, base_variables + (size_t)(b + 1) * rm.n_vars);
            fk.applyConfiguration(v);
            fk.initializeMutationApproximator(pr.active_variables);
            std::vector<const double*> gp(M);
            for(int m = 0; m < M; m++) gp[m] = genotypes + ((size_t)b * M + m) * n;
            std::vector<std::vector<Frame>> ph;
            fk.computeApproximateMutations(M, gp.data(), ph);
            for(int m = 0; m < M; m++)
            {
                if(out_primary) out_primary[(size_t)b * M + m] = pr.computeGoalFitness(pr.goals, ph[m].data(), gp[m]);
                if(out_secondary) out_secondary[(size_t)b * M + m] = pr.computeGoalFitness(pr.secondary_goals, null_frames.data(), gp[m]);
            }
        }
        return 0;
    }
    catch(std::exception& e)
    {
        g_error = e.what();
        return 1;
    }
}

// approximated tip frames [B][M][T][7] of genotypes [B][M][n] at base points [B][n_vars] (K2 alone)
int oracle_approx_frames_batch(const BioikRobot* robot, const BioikProblem* problem, int B, int M, const double* base_variables, const double* genotypes, double* out_frames)
{
    try
    {
        RobotModel rm = makeRobot(robot);
        Problem pr = makeProblem(rm, problem);
        size_t n = pr.active_variables.size(), T = pr.tip_link_indices.size();
        RobotFK fk(&rm, componentOptions());
        fk.initialize(pr.tip_link_indices);
        for(int b = 0; b < B; b++)
        {
            std::vector<double> v(base_variables + (size_t)b * rm.n_vars, base_variables + (size_t)(b + 1) * rm.n_vars);
            fk.applyConfiguration(v);
            fk.initializeMutationApproximator(pr.active_variables);
            std::vector<const double*> gp(M);
            for(int m = 0; m < M; m++) gp[m] = genotypes + ((size_t)b * M + m) * n;
            std::vector<std::vector<Frame>> ph;
            fk.computeApproximateMutations(M, gp.data(), ph);
            for(int m = 0; m < M; m++)
                for(size_t t = 0; t < T; t++) W7(ph[m][t], out_frames + (((size_t)b * M + m) * T + t) * 7);
        }
        return 0;
    }
    catch(std::exception& e)
    {
        g_error = e.what();
        return 1;
    }
}

// ---- the batch solve (SURVEY.md §8(c) batch contract) ------------------------------------------
// flags: bit0 = libm sincos, bit1 = scalar (non-FMA) approximator, bit2 = scalar computeApproximateMutation1 only,
//        bit3 = emulate the reference's stale-tip quirk Q2 (pinning study, test_reference_pin.py)
// trace outputs (may be NULL): genes/gradients [B][2][2][n], species_fitness [B][2]
int oracle_solve_batch(const BioikRobot* robot, const BioikProblem* problem, const BioikSolverCfg* cfg, void* tables, int B, const double* goal_params, const double* seeds, const uint32_t* rng_seeds, int steps, int early_exit,
                       int flags, int nthreads, double* out_solutions, double* out_fitness, int32_t* out_success, int32_t* out_steps, double* out_genes, double* out_gradients, double* out_species_fitness)
{
    try
    {
        RobotModel rm = makeRobot(robot);
        Problem pr0 = makeProblem(rm, problem);
        SolverConfig sc = makeCfg(cfg);
        const Tables& tb = *(Tables*)tables;
        Options opt;
        opt.libm_sincos = (flags & 1) != 0;
        opt.fma_approx = (flags & 2) == 0;
        opt.fma_approx1 = (flags & 4) == 0;
        opt.stale_tips = (flags & 8) != 0;
        size_t n = pr0.active_variables.size();
        std::atomic<int> failed(0);
        std::string err;
        parallelFor(B, nthreads, [&](int b) {
            try
            {
                Problem pr = pr0;
                applyQuery(pr, problem, goal_params ? goal_params + (size_t)b * problem->n_goals * GOAL_NPARAM : nullptr, seeds + (size_t)b * rm.n_vars, rm.n_vars);
                IKEvolution2* solver = nullptr;
                QueryResult res = solveQuery(rm, tb, pr, rng_seeds[b], sc, steps, early_exit != 0, opt, &solver);
                std::unique_ptr<IKEvolution2> guard(solver);
                if(out_solutions) std::copy(res.solution.begin(), res.solution.end(), out_solutions + (size_t)b * rm.n_vars);
                if(out_fitness) out_fitness[b] = res.fitness;
                if(out_success) out_success[b] = res.success;
                if(out_steps) out_steps[b] = res.steps;
                for(int s = 0; s < 2; s++)
                {
                    for(int k = 0; k < 2; k++)
                    {
                        if(out_genes) std::copy(solver->species[s].individuals[k].genes.begin(), solver->species[s].individuals[k].genes.end(), out_genes + (((size_t)b * 2 + s) * 2 + k) * n);
                        if(out_gradients) std::copy(solver->species[s].individuals[k].gradients.begin(), solver->species[s].individuals[k].gradients.end(), out_gradients + (((size_t)b * 2 + s) * 2 + k) * n);
                    }
                    if(out_species_fitness) out_species_fitness[(size_t)b * 2 + s] = solver->species[s].fitness;
                }
            }
            catch(std::exception& e)
            {
                if(!failed.exchange(1)) err = e.what();
            }
        });
        if(failed)
        {
            g_error = err;
            return 1;
        }
        return 0;
    }
    catch(std::exception& e)
    {
        g_error = e.what();
        return 1;
    }
}

// IKParallel's result selection (src/ik_parallel.h:218-258) over the `islands` runs of each of Q queries + the plugin's
// angle wrap (src/kinematics_plugin.cpp:580-611); same contract as bioik_solve_islands' reduction step.
// goal_params [Q*islands][G][NPARAM] or NULL, seeds [Q*islands][n_vars], run inputs sol [Q*islands][n_vars], fit, succ, steps.
int oracle_select_islands(const BioikRobot* robot, const BioikProblem* problem, int Q, int islands, const double* goal_params, const double* seeds, const double* sol, const double* fit, const int32_t* succ, const int32_t* steps, int wrap,
                          double* out_solutions, double* out_fitness, int32_t* out_success, int32_t* out_island, int32_t* out_steps)
{
    try
    {
        RobotModel rm = makeRobot(robot);
        Problem pr0 = makeProblem(rm, problem);
        for(int q = 0; q < Q; q++)
        {
            size_t b0 = (size_t)q * islands;
            Problem pr = pr0;
            applyQuery(pr, problem, goal_params ? goal_params + b0 * problem->n_goals * GOAL_NPARAM : nullptr, seeds + b0 * rm.n_vars, rm.n_vars);
            std::vector<std::vector<double>> sols(islands);
            std::vector<double> f(islands);
            std::vector<int> ok(islands);
            for(int k = 0; k < islands; k++)
            {
                sols[k].assign(sol + (b0 + k) * rm.n_vars, sol + (b0 + k + 1) * rm.n_vars);
                f[k] = fit[b0 + k];
                ok[k] = succ[b0 + k];
            }
            double best;
            size_t k = selectBestResult(pr, sols, f, ok, best);
            std::vector<double> state = sols[k];
            if(wrap) wrapAngles(rm, pr, state);
            std::copy(state.begin(), state.end(), out_solutions + (size_t)q * rm.n_vars);
            if(out_fitness) out_fitness[q] = best;
            if(out_success) out_success[q] = ok[k];
            if(out_island) out_island[q] = (int32_t)k;
            if(out_steps) out_steps[q] = steps ? steps[b0 + k] : 0;
        }
        return 0;
    }
    catch(std::exception& e)
    {
        g_error = e.what();
        return 1;
    }
}

// The whole of bioik_solve_islands on the CPU: lock-step islands (solveIslands), IKParallel's selection, the plugin's wrap.
// early_exit: 0 none, 1 per island, 2 per query (the reference's `finished` flag).  goal_params [Q][G][NPARAM] or NULL, seeds [Q][n_vars],
// rng_seeds [Q * islands].  Optional per-run outputs run_* [Q * islands]...
int oracle_solve_islands(const BioikRobot* robot, const BioikProblem* problem, const BioikSolverCfg* cfg, void* tables, int Q, int islands, const double* goal_params, const double* seeds, const uint32_t* rng_seeds, int steps,
                         int early_exit, int wrap, int flags, int nthreads, double* out_solutions, double* out_fitness, int32_t* out_success, int32_t* out_island, int32_t* out_steps, double* run_solutions, double* run_fitness,
                         int32_t* run_success, int32_t* run_steps)
{
    try
    {
        RobotModel rm = makeRobot(robot);
        Problem pr0 = makeProblem(rm, problem);
        SolverConfig sc = makeCfg(cfg);
        Options opt;
        opt.libm_sincos = (flags & 1) != 0;
        opt.fma_approx = (flags & 2) == 0;
        opt.fma_approx1 = (flags & 4) == 0;
        opt.stale_tips = (flags & 8) != 0;
        opt.island_stride = (flags >> 8) & 0xff; // bits 8..15: BIOIK_OPT_ISLAND_STREAM_STRIDE
        const Tables& tb = *(const Tables*)tables;
        std::atomic<int> failed(0);
        std::string err;
        parallelFor(Q, nthreads, [&](int q) {
            try
            {
                Problem pr = pr0;
                applyQuery(pr, problem, goal_params ? goal_params + (size_t)q * problem->n_goals * GOAL_NPARAM : nullptr, seeds + (size_t)q * rm.n_vars, rm.n_vars);
                IslandRuns r = solveIslands(rm, tb, pr, rng_seeds + (size_t)q * islands, islands, sc, steps, early_exit == 1, early_exit == 2, opt);
                double best;
                size_t k = selectBestResult(pr, r.solutions, r.fitness, r.success, best);
                std::vector<double> state = r.solutions[k];
                if(wrap) wrapAngles(rm, pr, state);
                std::copy(state.begin(), state.end(), out_solutions + (size_t)q * rm.n_vars);
                if(out_fitness) out_fitness[q] = best;
                if(out_success) out_success[q] = r.success[k];
                if(out_island) out_island[q] = (int32_t)k;
                if(out_steps) out_steps[q] = r.steps[k];
                for(int i = 0; i < islands; i++)
                {
                    size_t b = (size_t)q * islands + i;
                    if(run_solutions) std::copy(r.solutions[i].begin(), r.solutions[i].end(), run_solutions + b * rm.n_vars);
                    if(run_fitness) run_fitness[b] = r.fitness[i];
                    if(run_success) run_success[b] = r.success[i];
                    if(run_steps) run_steps[b] = r.steps[i];
                }
            }
            catch(std::exception& e)
            {
                if(!failed.exchange(1)) err = e.what();
            }
        });
        if(failed)
        {
            g_error = err;
            return 1;
        }
        return 0;
    }
    catch(std::exception& e)
    {
        g_error = e.what();
        return 1;
    }
}

int oracle_hardware_threads() { return (int)std::thread::hardware_concurrency(); }
}
