// bioik_serial.cuh — the fused per-task "serial" kernel: everything of step() that is one
// dependent chain per species:
//   MEMETIC  quadratic / linear line search on individuals[0]           src/ik_evolution_2.cpp:436-570
//   SPECIES  exact fitness, species sort, wipeout, solution update      src/ik_evolution_2.cpp:604-645
//            + the driver's success test                                src/ik_parallel.h:173-181
//   PREPARE  exact FK + Jacobian + delta frames for the next step       src/ik_evolution_2.cpp:341-346
// One thread per task (query, species slot); the two species of a query sit in adjacent lanes and
// meet through warp shuffles in the species block.  All per-thread arrays are shared-memory COLUMNS
// (element e of thread t at base[e * blockDim.x + t]: conflict-free), the flattened problem sits in
// the constant bank (kernel parameter), so the dependent chains never wait on local or global memory.
// The exact FK of the species block is reused by PREPARE unless the species was wiped out.
#pragma once

#include "bioik_dev.cuh"

#ifdef BIOIK_HOSTSIM
#define BIOIK_PROBLEM_PARAM const DProblem& P
#else
#define BIOIK_PROBLEM_PARAM const __grid_constant__ DProblem P
#endif

namespace bioik
{

enum SerialPhase { PH_MEMETIC = 1, PH_SPECIES = 2, PH_PREPARE = 4 };

// host-computed launch plan: which per-thread arrays fit in shared memory at which block size
struct SerialPlan
{
    int block;       // threads per block
    int delta_smem;  // delta frames [T][n][7] in shared memory (else read from the HBM state, L1/L2 cached)
    int frames_smem; // link frames [L][7] in shared memory (else thread-local memory)
    int per_thread;  // doubles per thread
    size_t smem_bytes;
};

__host__ __device__ inline int serial_fixed_doubles(const DProblem& P) { return 4 * P.n /*ind,temp,grad,stash*/ + 3 * 7 * P.T /*ph2,ph3,tip0*/ + P.n /*base*/ + GOAL_NPARAM * P.G + P.n_vars; }

inline SerialPlan make_serial_plan(const DProblem& P, size_t smem_limit = 200 * 1024)
{
    SerialPlan pl;
    const int fixed = serial_fixed_doubles(P), dl = 7 * P.T * P.n, fr = 7 * P.L;
    const int blocks[] = {128, 64, 32};
    for(int variant = 0; variant < 3; variant++) // 0: everything on chip, 1: delta in HBM state, 2: frames local too
        for(int b : blocks)
        {
            pl.block = b;
            pl.delta_smem = variant < 1;
            pl.frames_smem = variant < 2;
            pl.per_thread = fixed + (pl.delta_smem ? dl : 0) + (pl.frames_smem ? fr : 0);
            pl.smem_bytes = (size_t)pl.per_thread * b * sizeof(double);
            // want at least 64 resident threads per SM for the on-chip variants
            if(pl.smem_bytes * (b < 64 ? 2 : 1) <= smem_limit) return pl;
        }
    return pl; // 32 threads, delta + frames off chip: always fits (fixed part <= ~1 KB/thread)
}

template <class AT> BIOIK_HD void copy_tips(const DProblem& P, CCol frames, AT tips)
{
    for(int t = 0; t < P.T; t++)
        for(int k = 0; k < 7; k++) tips[7 * t + k] = frames[7 * P.tip_slot[t] + k];
}

__global__ void __launch_bounds__(128) k_serial(BIOIK_PROBLEM_PARAM, DState S, int step, int phases, int delta_smem, int frames_smem)
{
    extern __shared__ double smem[];
    const int tid = threadIdx.x, bs = blockDim.x;
    const int task_raw = blockIdx.x * bs + tid;
    const bool valid = task_raw < 2 * S.B;
    const int task = valid ? task_raw : 2 * S.B - 1;
    const int q = task >> 1, slot = task & 1;
    const bool active = valid && !S.done[q];
    const int n = P.n, T = P.T, G = P.G;

    // ---- per-thread columns --------------------------------------------------------------------
    int off = 0;
    auto col = [&](int len) {
        Col c{smem + (size_t)off * bs + tid, bs};
        off += len;
        return c;
    };
    Col ind = col(n), temp = col(n), grad = col(n), stash = col(n), ph2 = col(7 * T), ph3 = col(7 * T), tip0 = col(7 * T), base = col(n), gp = col(GOAL_NPARAM * G), vars = col(P.n_vars);
    Col delta, frames;
    if(delta_smem)
        delta = col(7 * T * n);
    else
        delta = Col{S.delta + (size_t)task * T * n * 7, 1};
    double lf[MAX_SLOTS * 7]; // only touched when the link frames do not fit in shared memory
    if(frames_smem)
        frames = col(7 * P.L);
    else
        frames = Col{lf, 1};

    const double* seed = S.seeds + (size_t)q * P.n_vars;
    if(active)
    {
        const double* g = S.goal_params + (size_t)q * G * GOAL_NPARAM;
        for(int k = 0; k < G * GOAL_NPARAM; k++) gp[k] = g[k];
        const double* gi = S.genes + ((size_t)task * 2 + 0) * n;
        for(int i = 0; i < n; i++) ind[i] = gi[i];
    }
    CCol cgp = gp;

    // ---- MEMETIC (src/ik_evolution_2.cpp:436-570) ----------------------------------------------------
    if(active && (phases & PH_MEMETIC) && S.memetic)
    {
        {
            const double* t0 = S.tip0 + (size_t)task * T * 7;
            for(int k = 0; k < 7 * T; k++) tip0[k] = t0[k];
            const double* b0 = S.base + (size_t)task * n;
            for(int i = 0; i < n; i++) base[i] = b0[i];
            if(delta_smem)
            {
                const double* d0 = S.delta + (size_t)task * T * n * 7;
                for(int k = 0; k < 7 * T * n; k++) delta[k] = d0[k];
            }
        }
        CCol cdelta = delta, ctip0 = tip0, cbase = base;
        double dp = 0.0000001;                               // :450
        if(S.uniform[(6165936u + (uint32_t)step * 3u + (uint32_t)slot) & ((1u << 23) - 1)] < 0.5) dp = -dp; // :451 fast_random()
        for(int generation = 0; generation < S.memetic_iters; generation++)
        {
            for(int i = 0; i < n; i++) temp[i] = ind[i];       // :460
            approx_frames(T, n, ctip0, cdelta, cbase, CCol(temp), ph2); // :462
            double f2p = goal_fitness_t(P, 0, cgp, CCol(ph2), CCol(temp), seed);                       // :463
            double fa = f2p + (P.has_secondary ? goal_fitness_secondary(P, cgp, CCol(temp), seed) : 0.0); // :464
            for(int i = 0; i < n; i++)                           // :465-474
            {
                temp[i] = ind[i] + dp;
                approx_frames1(T, n, cdelta, i, dp, CCol(ph2), ph3);
                double fb = 0.0;
                fb += goal_fitness_t(P, 0, cgp, CCol(ph3), CCol(temp), seed);
                fb += P.has_secondary ? goal_fitness_secondary(P, cgp, CCol(temp), seed) : 0.0;
                temp[i] = ind[i];
                grad[i] = fb - fa;
            }
            double sum = dp * dp; // :477-482
            for(int i = 0; i < n; i++) sum += BIOIK_FABS(grad[i]);
            double f = 1.0 / sum * dp;
            for(int i = 0; i < n; i++) grad[i] *= f;

            for(int i = 0; i < n; i++) temp[i] = ind[i] - grad[i]; // :485-488
            approx_frames(T, n, ctip0, cdelta, cbase, CCol(temp), ph3);
            double f1 = 0.0;
            f1 += goal_fitness_t(P, 0, cgp, CCol(ph3), CCol(temp), seed);
            f1 += P.has_secondary ? goal_fitness_secondary(P, cgp, CCol(temp), seed) : 0.0;
            double f2 = fa;
            for(int i = 0; i < n; i++) temp[i] = ind[i] + grad[i]; // :492-495
            approx_frames(T, n, ctip0, cdelta, cbase, CCol(temp), ph3);
            double f3 = 0.0;
            f3 += goal_fitness_t(P, 0, cgp, CCol(ph3), CCol(temp), seed);
            f3 += P.has_secondary ? goal_fitness_secondary(P, cgp, CCol(temp), seed) : 0.0;

            if(S.memetic == 'q') // :498-542
            {
                double v1 = (f2 - f1);
                double v2 = (f3 - f2);
                double v = (v1 + v2) * 0.5;
                double a = (v1 - v2);
                double step_size = v / a;
                for(int i = 0; i < n; i++) temp[i] = clampd(ind[i] + grad[i] * step_size * 1.0, P.genes[i].clip_min, P.genes[i].clip_max); // :525
            }
            else // 'l', :545-568
            {
                double cost_diff = (f3 - f1) * 0.5;
                double step_size = f2 / cost_diff;
                for(int i = 0; i < n; i++) temp[i] = clampd(ind[i] - grad[i] * step_size, P.genes[i].clip_min, P.genes[i].clip_max); // :554
            }
            approx_frames(T, n, ctip0, cdelta, cbase, CCol(temp), ph2);
            double f4p = goal_fitness_t(P, 0, cgp, CCol(ph2), CCol(temp), seed);
            if(f4p < f2p) // :530-538 / :559-567
            {
                for(int i = 0; i < n; i++) ind[i] = temp[i];
                continue;
            }
            else
                break;
        }
    }

    // ---- SPECIES (src/ik_evolution_2.cpp:604-645) --------------------------------------------------------
    int my_task = task;        // where this thread's species lives after the sort
    bool frames_valid = false; // `frames` holds the exact FK of `ind`
    if(phases & PH_SPECIES)
    {
        double f = 0.0;
        if(active)
        {
            assemble_variables(P, seed, CCol(ind), vars); // genesToJointVariables :610
            exact_fk(P, CCol(vars), frames);               // computeFitness :611 -> applyConfiguration
            copy_tips(P, CCol(frames), ph2);
            f = goal_fitness_t(P, 0, cgp, CCol(ph2), CCol(ind), seed);
            frames_valid = true;
        }
        const double fo = __shfl_xor_sync(0xffffffffu, f, 1);
        int improved = 0, nslot = slot;
        if(active)
        {
            improved = (f != S.sfit[task]) ? 1 : 0; // :612
            // :617 sort ascending: the two species swap places iff species[1].fitness < species[0].fitness
            const bool swap = slot == 0 ? (fo < f) : (f < fo);
            nslot = swap ? (slot ^ 1) : slot;
            my_task = 2 * q + nslot;
            // the rest of this species (individual 1, both gradient vectors) before anybody overwrites it
            const double* g1 = S.genes + ((size_t)task * 2 + 1) * n;
            const double* r0 = S.grads + ((size_t)task * 2 + 0) * n;
            const double* r1 = S.grads + ((size_t)task * 2 + 1) * n;
            for(int i = 0; i < n; i++)
            {
                temp[i] = g1[i];
                grad[i] = r0[i];
                stash[i] = r1[i];
            }
        }
        __syncwarp(); // both species of a query have read their state
        if(active)
        {
            S.sfit[my_task] = f;
            S.impr[my_task] = improved;
            if(nslot == 1)
            {
                // :620-637 wipeout of species[1]; then the random_index draws of the next step's generations (:369)
                uint32_t rng = S.rng[q];
                const double u = S.uniform[(6165936u + (uint32_t)step * (S.memetic ? 3u : 1u) + (S.memetic ? 2u : 0u)) & ((1u << 23) - 1)];
                if(u < 0.1 || !improved)
                {
                    for(int i = 0; i < n; i++)
                    {
                        double g = minstd_random(rng, P.genes[i].vmin, P.genes[i].vmax); // :629
                        ind[i] = g;
                        temp[i] = g; // individuals[i] = individuals[0]
                        grad[i] = 0.0;
                        stash[i] = 0.0;
                    }
                    frames_valid = false;
                }
                if(P.has_secondary)
                    for(int sl = 0; sl < 2; sl++)
                        for(int g = 0; g < S.gens; g++) S.ccount[((size_t)q * 2 + sl) * S.gens + g] = (int32_t)minstd_index(rng, (uint32_t)(S.C - 3)) + 3;
                S.rng[q] = rng;
            }
            else
            {
                // :640-644 solution update, then the driver's test (src/ik_parallel.h:165-181)
                bool sol_is_ind = false;
                if(f < S.solfit[q])
                {
                    for(int i = 0; i < n; i++) S.sol[(size_t)q * n + i] = ind[i];
                    S.solfit[q] = f;
                    sol_is_ind = true;
                }
                const int steps = S.steps[q] + 1;
                S.steps[q] = steps;
                if((steps % 4) == 0 || steps == S.total_steps)
                {
                    int ok;
                    if(sol_is_ind)
                        ok = check_solution(P, cgp, CCol(ph2), CCol(ind), seed) ? 1 : 0;
                    else
                    {
                        // exact FK of the (older) solution; frames/vars/ph3 are scratch here
                        const double* sol = S.sol + (size_t)q * n;
                        assemble_variables(P, seed, sol, vars);
                        exact_fk(P, CCol(vars), frames);
                        copy_tips(P, CCol(frames), ph3);
                        ok = check_solution(P, cgp, CCol(ph3), sol, seed) ? 1 : 0;
                        frames_valid = false;
                    }
                    S.success[q] = ok;
                    if(ok && S.early_exit) S.done[q] = 1;
                }
            }
            // write the species back (to its new slot)
            double* og0 = S.genes + ((size_t)my_task * 2 + 0) * n;
            double* og1 = S.genes + ((size_t)my_task * 2 + 1) * n;
            double* or0 = S.grads + ((size_t)my_task * 2 + 0) * n;
            double* or1 = S.grads + ((size_t)my_task * 2 + 1) * n;
            for(int i = 0; i < n; i++)
            {
                og0[i] = ind[i];
                og1[i] = temp[i];
                or0[i] = grad[i];
                or1[i] = stash[i];
            }
        }
    }
    else if(active && (phases & PH_MEMETIC))
    {
        double* og0 = S.genes + ((size_t)task * 2 + 0) * n;
        for(int i = 0; i < n; i++) og0[i] = ind[i];
    }

    // ---- PREPARE (src/ik_evolution_2.cpp:341-346) ---------------------------------------------------------
    if(active && (phases & PH_PREPARE))
    {
        if(!frames_valid)
        {
            assemble_variables(P, seed, CCol(ind), vars);
            exact_fk(P, CCol(vars), frames);
        }
        double* t0 = S.tip0 + (size_t)my_task * T * 7;
        for(int t = 0; t < T; t++)
            for(int k = 0; k < 7; k++) t0[7 * t + k] = frames[7 * P.tip_slot[t] + k];
        double* b0 = S.base + (size_t)my_task * n;
        // p_variables = post-mimic variables of the base configuration (:1063): for an active,
        // non-mimic variable that is the gene itself
        assemble_variables(P, seed, CCol(ind), vars);
        for(int i = 0; i < n; i++) b0[i] = vars[P.genes[i].var];
        double* d0 = S.delta + (size_t)my_task * T * n * 7;
        for(int t = 0; t < T; t++)
            for(int i = 0; i < n; i++)
            {
                bool masked;
                store_frame(d0 + ((size_t)t * n + i) * 7, delta_frame(P, CCol(frames), i, t, masked));
            }
    }
}

} // namespace bioik
