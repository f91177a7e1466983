// bioik_memetic_group.cuh — the memetic line search (src/ik_evolution_2.cpp:436-570) with a GROUP of W lanes
// per task instead of one thread.
//
// An iteration of the reference is n+4 dependent-looking evaluations, but only four of them depend on each
// other:   f2p  ->  { n gradient probes }  ->  { f1, f3 }  ->  f4p
// and a full approximation (computeApproximateMutations, forward_kinematics.h:1061-1110) is 7T independent
// FMA chains.  The group therefore runs
//   * the 7T (or 2 x 7T) frame components of a full approximation on 7T (14T) lanes,
//   * the n one-variable probes (computeApproximateMutation1 + combined fitness) on n lanes,
//   * the element-wise gene / gradient updates on n lanes,
// while every VALUE is produced by the same operations in the same order as in the thread-per-task kernel
// (k_serial, PH_MEMETIC) — the two are interchangeable bit for bit (tests/test_hostsim_parity.py).
// The thread-per-task version left the SMs with one warp or less each (B = 10 000 queries -> 625 warps on 148 SMs)
// and paid the full latency of every FP64 operation; here the same batch is 5 000 warps at W = 8.
#pragma once

#include "bioik_dev.cuh"
#include "bioik_serial.cuh"

namespace bioik
{

// genes of a probe: individual.genes with element i replaced (ik_evolution_2.cpp:468-471)
struct ProbeGenes
{
    const double* ind;
    int i;
    double v;
    BIOIK_HD double operator[](int k) const { return k == i ? v : ind[k]; }
};

// shared-memory block of one group, in doubles
struct GroupLayout
{
    int n, T, G, W;
    __host__ __device__ int o_ind() const { return 0; }
    __host__ __device__ int o_graw() const { return n; }
    __host__ __device__ int o_grad() const { return 2 * n; }
    __host__ __device__ int o_ta() const { return 3 * n; }
    __host__ __device__ int o_tb() const { return 4 * n; }
    __host__ __device__ int o_base() const { return 5 * n; }
    __host__ __device__ int o_clip() const { return 6 * n; } // [n][2]
    __host__ __device__ int o_tip0() const { return 8 * n; }
    __host__ __device__ int o_f2() const { return o_tip0() + 7 * T; }
    __host__ __device__ int o_pl() const { return o_f2() + 7 * T; } // [W][7T] per-lane frames; rows 0 and 1 double as the f1 / f3 frames
    __host__ __device__ int o_delta() const { return o_pl() + W * 7 * T; }
    __host__ __device__ int o_gp() const { return o_delta() + 7 * T * n; }
    __host__ __device__ int o_sc() const { return o_gp() + GOAL_NPARAM * G; }
    __host__ __device__ int o_carry() const { return o_sc() + 8; }               // [7T] reference-quirk mode: the frames left in phenotypes3
    __host__ __device__ int o_prev() const { return o_carry() + 7 * T; }          // [T][n] int32: last earlier gene that moves tip t (-1: none)
    __host__ __device__ int total() const { return (o_prev() + (T * n + 1) / 2) | 1; } // odd stride: the groups of a warp start in different banks
};

inline int memetic_group_width(int n) { return n <= 8 ? 8 : (n <= 16 ? 16 : 32); }

// W lanes per task, 32 / W tasks per warp; blockDim.x = 32 * warps (any number of warps, no block-level sync)
//
// STALE = the reference-quirk mode (bioik_set_option BIOIK_OPT_REFERENCE_STALE_TIPS): the reference's
// computeApproximateMutation1 skips the tips a variable cannot move (forward_kinematics.h:940), so a probe scores those
// tips on what phenotypes3[0] held before: the frame written by the last earlier probe of the iteration that moves the tip,
// else the frames of the previous f3 evaluation (:494) - of the previous iteration, of the other species (the solver has ONE
// phenotypes3 and treats species 0, then species 1), or of the previous step.  Here a group owns a QUERY, runs its two
// species one after the other and carries those frames in `carry` (HBM between steps; identity frames at the start, which
// is what the harness of the reference build pre-fills - a fresh reference solver reads uninitialised memory there).
template <int W, bool STALE = false> __global__ void __launch_bounds__(128) k_memetic_group(BIOIK_PROBLEM_PARAM, DState S, int step)
{
    extern __shared__ double smem[];
    constexpr int GPW = 32 / W;
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp_in_block = threadIdx.x >> 5, warps_per_block = blockDim.x >> 5;
    const int gl = lane % W, gw = lane / W;
    const int unit_raw = (blockIdx.x * warps_per_block + warp_in_block) * GPW + gw; // task (query, slot), or query in STALE mode
    const int units = STALE ? S.B : 2 * S.B;
    const bool valid = unit_raw < units;
    const int unit = valid ? unit_raw : units - 1;
    const int q = STALE ? unit : (unit >> 1);
    const bool live = valid && !run_done(S, q, step) && S.memetic;
    const int n = P.n, T = P.T, G = P.G, T7 = 7 * P.T;

    const GroupLayout L{n, T, G, W};
    double* Wk = smem + (size_t)(warp_in_block * GPW + gw) * L.total();
    double *ind = Wk + L.o_ind(), *graw = Wk + L.o_graw(), *grad = Wk + L.o_grad(), *ta = Wk + L.o_ta(), *tb = Wk + L.o_tb(), *base = Wk + L.o_base(), *clip = Wk + L.o_clip();
    double *tip0 = Wk + L.o_tip0(), *f2 = Wk + L.o_f2(), *pl = Wk + L.o_pl(), *delta = Wk + L.o_delta(), *gp = Wk + L.o_gp(), *sc = Wk + L.o_sc(), *carry = Wk + L.o_carry();
    int32_t* prev = (int32_t*)(Wk + L.o_prev());
    const double* seed = S.seeds + (size_t)q * P.n_vars;
    if(STALE && live)
    {
        for(int c = gl; c < T7; c += W) carry[c] = S.carry[(size_t)q * T7 + c];
        for(int k = gl; k < T * n; k += W)
        {
            const int t = k / n, i = k - t * n;
            int j = i - 1;
            while(j >= 0 && !((P.genes[j].tipmask >> t) & 1)) j--;
            prev[k] = j;
        }
    }
    for(int slot_it = 0; slot_it < (STALE ? 2 : 1); slot_it++)
    {
    const int task = STALE ? 2 * q + slot_it : unit;
    const int slot = task & 1;
    bool alive = live;

    // ---- stage the task ------------------------------------------------------------------------
    if(alive)
    {
        const double* g = S.goal_params + (size_t)q * G * GOAL_NPARAM;
        for(int k = gl; k < G * GOAL_NPARAM; k += W) gp[k] = g[k];
        const double* gi = S.genes + ((size_t)task * 2 + 0) * n;
        const double* b0 = S.base + (size_t)task * n;
        for(int i = gl; i < n; i += W)
        {
            ind[i] = gi[i];
            base[i] = b0[i];
            clip[2 * i + 0] = P.genes[i].clip_min;
            clip[2 * i + 1] = P.genes[i].clip_max;
        }
        const double* t0 = S.tip0 + (size_t)task * T * 7;
        for(int k = gl; k < T7; k += W) tip0[k] = t0[k];
        const double* d0 = S.delta + (size_t)task * T * n * 7;
        for(int k = gl; k < T7 * n; k += W) delta[k] = d0[k];
    }
    __syncwarp();

    double dp = 0.0000001;                                                                                // :450
    if(S.uniform[(6165936u + (uint32_t)step * 3u + (uint32_t)slot) & ((1u << 23) - 1)] < 0.5) dp = -dp; // :451 fast_random()
    const bool quad = S.memetic == 'q';

    // component c (= 7 t + k) of the full approximation of genotype x: the FMA chain of approx_frames_sparse
    auto chain = [&](const double* x, int c) {
        const int t = c / 7;
        double f = tip0[c];
        const double* D = delta + (size_t)t * n * 7 + (c - 7 * t);
        for(int i = 0; i < n; i++)
        {
            if(!((P.genes[i].tipmask >> t) & 1)) continue;
            const double d = x[i] - base[i]; // :1086
            f = BIOIK_FMA(d, D[7 * i], f);
        }
        return f;
    };
    // computeCombinedFitnessActiveVariables (src/ik_base.h:179-185) / computeFitnessActiveVariables
    auto primary = [&](const double* frames, const double* x) { return goal_fitness_t(P, 0, (const double*)gp, frames, x, seed); };
    auto combined = [&](const double* frames, const double* x) {
        const double prim = primary(frames, x);
        return prim + (P.has_secondary ? goal_fitness_secondary(P, (const double*)gp, x, seed) : 0.0);
    };

    for(int generation = 0; generation < S.memetic_iters; generation++)
    {
        if(!__any_sync(FULL, alive)) break;
        // (1) genotype = individual.genes -> phenotypes2 (:460-462)
        if(alive)
            for(int c = gl; c < T7; c += W) f2[c] = chain(ind, c);
        __syncwarp();
        // (2) f2p, fa (:463-464)
        if(alive && gl == 0)
        {
            const double prim = primary(f2, ind);
            sc[0] = prim;
            sc[1] = prim + (P.has_secondary ? goal_fitness_secondary(P, (const double*)gp, (const double*)ind, seed) : 0.0);
        }
        __syncwarp();
        // (3) gradient probes (:465-474): lane i moves variable i by dp
        if(alive)
        {
            const double fa = sc[1];
            double* ph3 = pl + gl * T7;
            for(int i = gl; i < n; i += W)
            {
                for(int t = 0; t < T; t++)
                {
                    const double* D = delta + ((size_t)t * n + i) * 7;
                    if(STALE && !((P.genes[i].tipmask >> t) & 1))
                    {
                        // the tip keeps what the buffer held: the write of the last earlier probe that moves it, else the carried frame
                        const int j = prev[t * n + i];
                        const double* Dj = delta + ((size_t)t * n + (j >= 0 ? j : 0)) * 7;
                        for(int k = 0; k < 7; k++) ph3[7 * t + k] = j >= 0 ? BIOIK_FMA(dp, Dj[k], f2[7 * t + k]) : carry[7 * t + k];
                        continue;
                    }
                    for(int k = 0; k < 7; k++) ph3[7 * t + k] = BIOIK_FMA(dp, D[k], f2[7 * t + k]); // :469
                }
                const ProbeGenes x{ind, i, ind[i] + dp}; // :468
                const double prim = goal_fitness_t(P, 0, (const double*)gp, (const double*)ph3, x, seed);
                const double comb = prim + (P.has_secondary ? goal_fitness_secondary(P, (const double*)gp, x, seed) : 0.0);
                graw[i] = comb - fa; // :472-473
            }
        }
        __syncwarp();
        // (4) normalise (:477-482) and the two support points (:485-486, :492-493)
        if(alive)
        {
            double sum = dp * dp;
            for(int i = 0; i < n; i++) sum += BIOIK_FABS(graw[i]);
            const double f = 1.0 / sum * dp;
            for(int i = gl; i < n; i += W)
            {
                const double g = graw[i] * f;
                grad[i] = g;
                ta[i] = ind[i] - g;
                tb[i] = ind[i] + g;
            }
        }
        __syncwarp();
        // (5) both support points -> frames (rows 0 and 1 of the per-lane frame block)
        if(alive)
            for(int c = gl; c < 2 * T7; c += W)
            {
                const int which = c >= T7 ? 1 : 0;
                const int cc = c - which * T7;
                pl[which * T7 + cc] = chain(which ? tb : ta, cc);
            }
        __syncwarp();
        // (6) f1, f3 (:487-488, :494-495)
        if(alive && gl < 2) sc[2 + gl] = combined(pl + gl * T7, gl ? tb : ta);
        __syncwarp();
        // (7) step size and the candidate (:502-506,:525 / :549-554)
        if(alive)
        {
            if(STALE)
                for(int c = gl; c < T7; c += W) carry[c] = pl[T7 + c]; // phenotypes3 now holds the frames of the f3 evaluation (:494)
            const double f2v = sc[1], f1 = sc[2], f3 = sc[3];
            if(quad)
            {
                double v1 = (f2v - f1);
                double v2 = (f3 - f2v);
                double v = (v1 + v2) * 0.5;
                double a = (v1 - v2);
                double step_size = v / a;
                for(int i = gl; i < n; i += W) ta[i] = clampd(ind[i] + grad[i] * step_size * 1.0, clip[2 * i], clip[2 * i + 1]);
            }
            else
            {
                double cost_diff = (f3 - f1) * 0.5;
                double step_size = f2v / cost_diff;
                for(int i = gl; i < n; i += W) ta[i] = clampd(ind[i] - grad[i] * step_size, clip[2 * i], clip[2 * i + 1]);
            }
        }
        __syncwarp();
        // (8) candidate -> phenotypes2 (:526 / :555)
        if(alive)
            for(int c = gl; c < T7; c += W) f2[c] = chain(ta, c);
        __syncwarp();
        // (9) f4p
        if(alive && gl == 0) sc[4] = primary(f2, ta);
        __syncwarp();
        // (10) accept / stop (:530-538 / :559-567)
        if(alive)
        {
            if(sc[4] < sc[0])
                for(int i = gl; i < n; i += W) ind[i] = ta[i];
            else
                alive = false;
        }
        __syncwarp();
    }

    // individuals[0].genes back to the state (gradients are not touched by the memetic step)
    if(live)
    {
        double* og0 = S.genes + ((size_t)task * 2 + 0) * n;
        for(int i = gl; i < n; i += W) og0[i] = ind[i];
    }
    __syncwarp(); // the next species of the query re-uses the block
    } // species of the query (STALE), or the single task
    if(STALE && live)
        for(int c = gl; c < T7; c += W) S.carry[(size_t)q * T7 + c] = carry[c];
}

#ifdef BIOIK_HOSTSIM
typedef void (*MemeticGroupKernel)(const DProblem&, DState, int);
#else
typedef void (*MemeticGroupKernel)(const DProblem, DState, int);
#endif

inline MemeticGroupKernel select_memetic_group(int W, bool stale = false)
{
#ifdef BIOIK_SLIM
    return (MemeticGroupKernel)k_memetic_group<8>;
#else
    if(stale) return W == 8 ? (MemeticGroupKernel)k_memetic_group<8, true> : (W == 16 ? (MemeticGroupKernel)k_memetic_group<16, true> : (MemeticGroupKernel)k_memetic_group<32, true>);
    return W == 8 ? (MemeticGroupKernel)k_memetic_group<8> : (W == 16 ? (MemeticGroupKernel)k_memetic_group<16> : (MemeticGroupKernel)k_memetic_group<32>);
#endif
}

// the reference-quirk mode changes something only if some gene cannot move some tip
__host__ __device__ inline bool stale_tips_matter(const DProblem& P)
{
    for(int i = 0; i < P.n; i++)
        if((P.genes[i].tipmask & ((1 << P.T) - 1)) != ((1 << P.T) - 1)) return true;
    return false;
}

} // namespace bioik
