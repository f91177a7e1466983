// hostsim.cpp — TEST-ONLY host simulation of the CUDA kernels (there is no GPU in the build
// container).  Compiles bio_ik_b200/csrc/bioik_kernels.cuh unchanged with g++ by shimming the few
// CUDA constructs it uses: thread-per-item kernels run as plain loops, the warp-cooperative
// k_evolve runs on 32 OS threads with barrier-based __syncwarp/__shfl emulation.  This lets the
// CPU test-suite check kernel logic + host flattening bit-for-bit against the oracle before any
// GPU time is spent.  It is NOT part of libbioik_b200.so and is never used as a fallback.
#define BIOIK_HOSTSIM 1

#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <thread>
#include <vector>

// ---- CUDA shims ---------------------------------------------------------------------------
struct sim_uint3
{
    unsigned x = 0, y = 0, z = 0;
};
static thread_local sim_uint3 threadIdx, blockIdx;
static sim_uint3 blockDim, gridDim;
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __restrict__
#define __shared__
namespace bioik { double smem[1 << 18]; } // the `extern __shared__ double smem[]` of k_evolve

static std::barrier<>* g_warp_barrier = nullptr;
// barriers of the aligned lane groups (16- and 8-lane masks): kernels that put several tasks in a warp synchronise per group
static std::barrier<>* g_group_barrier[2][4] = {{nullptr, nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr, nullptr}};
static uint64_t g_shfl_slots[32];
static inline std::barrier<>& sim_barrier(unsigned mask)
{
    if(mask == 0xffffffffu) return *g_warp_barrier;
    const int width = __builtin_popcount(mask), first = __builtin_ctz(mask);
    return *g_group_barrier[width == 16 ? 0 : 1][first / width];
}
static inline void __syncwarp(unsigned mask = 0xffffffffu) { sim_barrier(mask).arrive_and_wait(); }
template <class T> static inline T sim_shfl(T v, int src, unsigned mask = 0xffffffffu)
{
    static_assert(sizeof(T) <= 8, "");
    uint64_t raw = 0;
    memcpy(&raw, &v, sizeof(T));
    g_shfl_slots[threadIdx.x & 31] = raw;
    sim_barrier(mask).arrive_and_wait();
    uint64_t got = g_shfl_slots[src & 31];
    sim_barrier(mask).arrive_and_wait();
    T r;
    memcpy(&r, &got, sizeof(T));
    return r;
}
template <class T> static inline T __shfl_xor_sync(unsigned m, T v, int o) { return sim_shfl(v, (int)(threadIdx.x & 31) ^ o, m); }
template <class T> static inline T __shfl_sync(unsigned m, T v, int src) { return sim_shfl(v, src, m); }
static inline unsigned __reduce_min_sync(unsigned mask, unsigned v)
{
    unsigned m = v;
    for(int o = __builtin_popcount(mask) / 2; o > 0; o >>= 1) // butterfly inside the aligned group
    {
        unsigned other = sim_shfl(m, (int)(threadIdx.x & 31) ^ o, mask);
        m = other < m ? other : m;
    }
    return m;
}
static inline unsigned __reduce_or_sync(unsigned mask, unsigned v)
{
    unsigned acc = v;
    for(int o = __builtin_popcount(mask) / 2; o > 0; o >>= 1) acc |= sim_shfl(acc, (int)(threadIdx.x & 31) ^ o, mask);
    return acc;
}
static inline unsigned __ballot_sync(unsigned mask, bool pred)
{
    unsigned bit = pred ? (1u << (threadIdx.x & 31)) : 0u, acc = bit;
    for(int o = __builtin_popcount(mask) / 2; o > 0; o >>= 1) acc |= sim_shfl(acc, (int)(threadIdx.x & 31) ^ o, mask);
    return acc;
}
static inline bool __any_sync(unsigned m, bool pred) { return __ballot_sync(m, pred) != 0; }
static inline int atomicMin(int* a, int v)
{
    int old = *a;
    if(v < old) *a = v;
    return old;
}
// the queue operations of the persistent kernel: in the simulation only one lane of the single resident warp executes them at a time
static inline int atomicAdd(int* a, int v)
{
    int old = *a;
    *a = old + v;
    return old;
}
static inline int atomicCAS(int* a, int cmp, int v)
{
    int old = *a;
    if(old == cmp) *a = v;
    return old;
}
static inline int atomicExch(int* a, int v)
{
    int old = *a;
    *a = v;
    return old;
}
static inline void __threadfence() {}
using std::max;
using std::min;
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline double __longlong_as_double(long long v)
{
    double r;
    memcpy(&r, &v, 8);
    return r;
}
static inline long long __double_as_longlong(double d)
{
    long long r;
    memcpy(&r, &d, 8);
    return r;
}
using std::fabs;

#include "../../bio_ik_b200/csrc/bioik_host.hpp"
#include "../../bio_ik_b200/csrc/bioik_kernels.cuh"

using namespace bioik;

namespace
{
template <class F> void launch_serial(int blocks, int tpb, F f)
{
    blockDim.x = tpb;
    gridDim.x = blocks;
    for(int b = 0; b < blocks; b++)
        for(int t = 0; t < tpb; t++)
        {
            blockIdx.x = b;
            threadIdx.x = t;
            f();
        }
}
template <class F> void launch_warp(int blocks, F f)
{
    blockDim.x = 32;
    gridDim.x = blocks;
    std::barrier<> bar(32);
    g_warp_barrier = &bar;
    std::barrier<> b16[2] = {std::barrier<>(16), std::barrier<>(16)};
    std::barrier<> b8[4] = {std::barrier<>(8), std::barrier<>(8), std::barrier<>(8), std::barrier<>(8)};
    for(int i = 0; i < 2; i++) g_group_barrier[0][i] = &b16[i];
    for(int i = 0; i < 4; i++) g_group_barrier[1][i] = &b8[i];
    std::vector<std::thread> lanes;
    for(int l = 0; l < 32; l++)
        lanes.emplace_back([&, l]() {
            for(int b = 0; b < blocks; b++)
            {
                blockIdx.x = b;
                threadIdx.x = l;
                f();
                bar.arrive_and_wait();
            }
        });
    for(auto& t : lanes) t.join();
    g_warp_barrier = nullptr;
}

struct SimTables
{
    std::vector<double> uniform, gauss;
};
std::map<uint32_t, std::unique_ptr<SimTables>> g_tables;
std::string g_err;
} // namespace

extern "C" {
const char* hostsim_last_error() { return g_err.c_str(); }

const double* hostsim_tables(uint32_t seed, int which)
{
    auto& t = g_tables[seed];
    if(!t)
    {
        t.reset(new SimTables());
        make_tables(seed, t->uniform, t->gauss);
    }
    return which ? t->gauss.data() : t->uniform.data();
}

// device-side RNG helpers, for direct comparison with libstdc++
void hostsim_minstd_uniform(uint32_t seed, int n, double* out)
{
    uint32_t s = seed % 2147483647u;
    if(s == 0) s = 1;
    for(int i = 0; i < n; i++) out[i] = minstd_uniform01(s);
}
void hostsim_minstd_index(uint32_t seed, uint32_t m, int n, uint64_t* out)
{
    uint32_t s = seed % 2147483647u;
    if(s == 0) s = 1;
    for(int i = 0; i < n; i++) out[i] = minstd_index(s, m);
}
void hostsim_sincos(int n, const double* x, double* s, double* c)
{
    for(int i = 0; i < n; i++) d_sincos(x[i], s[i], c[i]);
}

// lanes per task wanted for the single-pose generation kernel (8, 16, 32)
static int g_lpt_want = 16;
void hostsim_set_evolve_lanes(int lpt) { g_lpt_want = lpt; }

// reference-quirk mode of the memetic step (bioik_set_option BIOIK_OPT_REFERENCE_STALE_TIPS) for the next hostsim_solve calls
static int g_stale = 0;
void hostsim_set_stale_tips(int on) { g_stale = on; }

// islands per query of the next hostsim_solve calls (0: plain batch), for early_exit == 2
static int g_islands = 0, g_island_stride = 0;
void hostsim_set_islands(int islands) { g_islands = islands; }
void hostsim_set_island_stride(int stride) { g_island_stride = stride; } // BIOIK_OPT_ISLAND_STREAM_STRIDE

// the launch sequence of enqueue_solve() in bioik_capi.cu, on host memory
int hostsim_solve(const BioikRobot* robot, const BioikProblem* problem, const BioikSolverCfg* cfg, int B, const double* goal_params, const double* seeds, const uint32_t* rng_seeds, int steps, int early_exit, double* out_solutions,
                  double* out_fitness, int32_t* out_success, int32_t* out_steps, double* out_genes, double* out_gradients, double* out_species_fitness, int use_fast)
{
    HostRobot R;
    int rc = intake_robot(robot, R, g_err);
    if(rc) return rc;
    static DProblem P; // large
    rc = build_problem(R, problem, P, g_err);
    if(rc) return rc;
    std::vector<int32_t> go;
    std::vector<uint8_t> re;
    make_schedules(std::max(steps, 1) + (g_islands > 1 ? (g_islands - 1) * g_island_stride : 0), cfg->generations, cfg->population, P.n, go, re);
    size_t n = P.n, T = P.T, gens = cfg->generations;
    std::vector<double> genes(B * 4 * n), grads(B * 4 * n), sfit(B * 2), sol(B * n), solfit(B), base(B * 2 * n), tip0(B * 2 * T * 7), delta(B * 2 * T * n * 7);
    std::vector<int32_t> impr(B * 2), done(B), stp(B), succ(B), cc(B * 2 * gens), qstep(B);
    std::vector<double> carry(B * T * 7);
    std::vector<uint32_t> rng(B);
    std::vector<double> gp_default;
    if(!goal_params)
    {
        gp_default.resize((size_t)B * P.G * GOAL_NPARAM);
        for(int b = 0; b < B; b++)
            for(int g = 0; g < P.G; g++)
                for(int k = 0; k < GOAL_NPARAM; k++) gp_default[((size_t)b * P.G + g) * GOAL_NPARAM + k] = problem->goals[g].p[k];
        goal_params = gp_default.data();
    }
    DState S;
    memset(&S, 0, sizeof(S));
    S.B = B, S.C = cfg->population, S.gens = cfg->generations, S.memetic = cfg->memetic, S.memetic_iters = cfg->memetic_iters, S.total_steps = steps, S.early_exit = early_exit, S.islands = g_islands, S.island_stride = g_islands > 1 ? g_island_stride : 0;
    S.goal_params = goal_params, S.seeds = seeds, S.rng_seeds = rng_seeds;
    S.genes = genes.data(), S.grads = grads.data(), S.sfit = sfit.data(), S.impr = impr.data(), S.sol = sol.data(), S.solfit = solfit.data(), S.rng = rng.data(), S.done = done.data(), S.steps = stp.data(),
    S.success = succ.data(), S.ccount = cc.data(), S.qstep = qstep.data(), S.carry = carry.data(), S.cancel = nullptr, S.base = base.data(), S.tip0 = tip0.data(), S.delta = delta.data();
    S.uniform = hostsim_tables(cfg->table_seed, 0), S.gauss = hostsim_tables(cfg->table_seed, 1), S.gauss_off = go.data(), S.rate_exp = re.data();
    S.gauss_absmax = 0;
    for(size_t i = 0; i < (size_t)1 << 23; i++) S.gauss_absmax = std::max(S.gauss_absmax, std::fabs(S.gauss[i]));
    const int TPB = 128;
    int qblocks = (B + TPB - 1) / TPB, tblocks = (2 * B + TPB - 1) / TPB;
    int evolve_lpt = 32;
    EvolveFastKernel fast = use_fast ? select_evolve_fast(P, S.C, 8, &evolve_lpt, g_lpt_want) : nullptr;
    // no fast instantiation (e.g. quaternion genes): like the library, the production sequence then runs the generic k_evolve
    std::vector<double> mtab;
    if(fast)
    {
        int calls = (int)go.size();
        long long total = (long long)calls * P.n * mtab_row(S.C);
        mtab.resize(total);
        launch_serial((int)((total + 255) / 256), 256, [&]() { k_mutation_table(&P, calls, S.C, S.gauss, S.gauss_off, S.rate_exp, mtab.data()); });
    }
    launch_serial(qblocks, TPB, [&]() { k_init(&P, S); });
    if(use_fast)
    {
        // production launch sequence of enqueue_solve(): k_evolve_fast + the fused k_serial (32-thread blocks here)
        SerialPlan pl = make_serial_plan(P);
        const bool stale = g_stale && S.memetic && stale_tips_matter(P);
        const bool group_memetic = use_fast >= 6 || stale; // 6 = k_memetic_group + k_serial(SPECIES|PREPARE), the library's default sequence
        if(use_fast >= 2 && use_fast <= 5)
        {
            // forced placement variant of the serial kernel: 2 = all on chip, 3 = frames local, 4 = delta in HBM, 5 = both off chip
            pl.delta_smem = (use_fast == 2 || use_fast == 3);
            pl.frames_smem = (use_fast == 2 || use_fast == 4);
        }
        pl.block = 32;
        SerialKernel ks = select_serial(pl);
        const int sgrid = (2 * B + 31) / 32;
        bool pds = true, pfs = false;
        PersistKernel pk = use_fast == 10 ? select_persist(P, S.C, &pds, &pfs) : nullptr;
        if(use_fast == 10 && (!pk || !fast || evolve_lpt != 16))
        {
            g_err = "no persistent kernel for this problem shape";
            return 1;
        }
        if(pk)
        {
            // the persistent kernel (bioik_persist.cuh) with ONE resident warp: it drains both queues by itself, which exercises
            // the whole dependency logic (PREPARE items, group counters, the hand-over between evolve and serial items); two
            // launches with a cut in the middle exercise the resume path of bioik_step
            const int groups = (B + PERSIST_GROUP_QUERIES - 1) / PERSIST_GROUP_QUERIES;
            std::vector<unsigned long long> slots((size_t)(steps + 1) * (B + groups) + 8);
            std::vector<int32_t> ctr(PQ_INTS), gcount(groups);
            bool prepared = false;
            const int cut = steps / 2;
            for(int part = 0; part < 2; part++)
            {
                const int a = part ? cut : 0, b = part ? steps : cut;
                if(b <= a) continue;
                PersistArgs A;
                A.s0 = a, A.s1 = b, A.last = b == steps ? 1 : 0, A.prepared = prepared ? 1 : 0, A.groups = groups, A.sm_count = 1;
                A.cap_e = (b - a) * B, A.cap_s = (b - a + 1) * groups;
                A.slots_e = slots.data(), A.slots_s = slots.data() + A.cap_e, A.ctr = ctr.data(), A.gcount = gcount.data();
                const int fill = std::max(std::max(A.cap_e, A.cap_s), groups);
                launch_serial((fill + 255) / 256, 256, [&]() { k_persist_init(A, B); });
                launch_warp(1, [&]() { pk(P, &P, S, A, mtab.data()); });
                prepared = b != steps;
            }
        }
        else if(steps > 0)
            launch_warp(sgrid, [&]() { ks(P, S, 0, PH_PREPARE); });
        for(int step = 0; step < steps && !pk; step++)
        {
            if(fast)
                launch_warp((2 * B + 32 / evolve_lpt - 1) / (32 / evolve_lpt), [&]() { fast(&P, S, step, mtab.data()); });
            else
                launch_warp(2 * B, [&]() { k_evolve(&P, S, step); });
            int phases = (S.memetic ? PH_MEMETIC : 0) | PH_SPECIES | (step + 1 < steps ? PH_PREPARE : 0);
            if(group_memetic && S.memetic)
            {
                const int MW = use_fast == 7 ? 8 : (use_fast == 8 ? 16 : (use_fast == 9 ? 32 : memetic_group_width(P.n))); // 7..9 force a width
                MemeticGroupKernel mgk = select_memetic_group(MW, stale);
                launch_warp(((stale ? B : 2 * B) + 32 / MW - 1) / (32 / MW), [&]() { mgk(P, S, step); });
                phases &= ~PH_MEMETIC;
            }
            launch_warp(sgrid, [&]() { ks(P, S, step, phases); });
        }
    }
    else
        for(int step = 0; step < steps; step++)
        {
            // the generic kernels (BIOIK_FORCE_GENERIC=1 path of the library)
            launch_serial(tblocks, TPB, [&]() { k_prepare(&P, S, step); });
            launch_warp(2 * B, [&]() { k_evolve(&P, S, step); });
            if(S.memetic) launch_serial(tblocks, TPB, [&]() { k_memetic(&P, S, step); });
            launch_serial(qblocks, TPB, [&]() { k_species(&P, S, step); });
        }
    launch_serial(qblocks, TPB, [&]() { k_finalize(&P, S, out_solutions, out_fitness, out_success, out_steps); });
    if(out_genes) memcpy(out_genes, genes.data(), genes.size() * 8);
    if(out_gradients) memcpy(out_gradients, grads.data(), grads.size() * 8);
    if(out_species_fitness) memcpy(out_species_fitness, sfit.data(), sfit.size() * 8);
    return 0;
}

// k_select_islands on host memory (the reduction step of bioik_solve_islands)
int hostsim_select_islands(const BioikRobot* robot, const BioikProblem* problem, int Q, int islands, const double* goal_params, const double* seeds, const double* sol, const double* fit, const int32_t* succ, const int32_t* steps, int wrap,
                           double* out_solutions, double* out_fitness, int32_t* out_success, int32_t* out_island, int32_t* out_steps)
{
    HostRobot R;
    int rc = intake_robot(robot, R, g_err);
    if(rc) return rc;
    static DProblem P;
    rc = build_problem(R, problem, P, g_err);
    if(rc) return rc;
    std::vector<double> gp_default;
    if(!goal_params)
    {
        gp_default.resize((size_t)Q * islands * P.G * GOAL_NPARAM);
        for(int b = 0; b < Q * islands; b++)
            for(int g = 0; g < P.G; g++)
                for(int k = 0; k < GOAL_NPARAM; k++) gp_default[((size_t)b * P.G + g) * GOAL_NPARAM + k] = problem->goals[g].p[k];
        goal_params = gp_default.data();
    }
    launch_serial((Q + 127) / 128, 128, [&]() { k_select_islands(&P, Q, islands, goal_params, seeds, sol, fit, succ, steps, wrap, out_solutions, out_fitness, out_success, out_island, out_steps); });
    return 0;
}

int hostsim_fk(const BioikRobot* robot, const BioikProblem* problem, int B, const double* variables, double* out_tips, double* out_delta)
{
    HostRobot R;
    int rc = intake_robot(robot, R, g_err);
    if(rc) return rc;
    static DProblem P;
    rc = build_problem(R, problem, P, g_err);
    if(rc) return rc;
    if(out_tips) launch_serial((B + 127) / 128, 128, [&]() { k_fk_batch(&P, B, variables, out_tips); });
    if(out_delta) launch_serial((B + 127) / 128, 128, [&]() { k_approx_batch(&P, B, variables, out_delta); });
    return 0;
}
}
