// bioik_evolve_fast.cuh — register-blocked generation kernel (the dominant kernel).
//
// Same arithmetic as k_evolve / the reference (src/ik_evolution_2.cpp:351-432), different
// schedule:
//   * one warp per task, lane l owns child slots l, l+32, ... (CPL of them) and evaluates
//     them TOGETHER: per gene the warp-uniform operands (parent gene, limits, base, delta
//     frame) are read once from shared memory and applied to CPL children, tip-frame
//     accumulators live in registers, genes are never stored;
//   * the mutation term r * (mutation_rate * span) of every (reproduce call, gene, child)
//     is query-independent — gaussians and rate exponents are consumed in a fixed order
//     (src/ik_evolution_2.cpp:254-265,288-293) — so it is tabulated once per problem
//     (k_mutation_table) in a [call][gene][child] layout: lanes read consecutive doubles;
//   * joint-space goals (sums over genes) are accumulated inside the gene loop in gene
//     order, link goals are evaluated from the register tip frames; goal order of the
//     final weighted sum is kept (src/problem.cpp:251-257);
//   * the two winners are re-derived from the table instead of storing 2*n doubles per child.
#pragma once

#include "bioik_dev.cuh"

#ifdef BIOIK_HOSTSIM
#define BIOIK_LDG(p) (*(p))
#else
#define BIOIK_LDG(p) __ldg(p)
#endif

namespace bioik
{

constexpr int FAST_MAX_JOINT_GOALS = 4;

// mutation table: mtab[(call * n + gene) * C + child] = r * (mutation_rate * span)   (:288-293)
__global__ void k_mutation_table(const DProblem* __restrict__ Pp, int calls, int C, const double* __restrict__ gauss, const int32_t* __restrict__ gauss_off, const uint8_t* __restrict__ rate_exp, double* __restrict__ mtab)
{
    const DProblem& P = *Pp;
    const int n = P.n;
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)calls * n * C;
    if(idx >= total) return;
    int c = (int)(idx % C);
    int i = (int)((idx / C) % n);
    int call = (int)(idx / ((long long)C * n));
    double m = 0.0;
    if(c >= 2)
    {
        const int stride4 = (n + 3) / 4 * 4;
        double r = gauss[gauss_off[call] + (size_t)(c - 2) * stride4 + i];
        double mutation_rate = (double)(1 << rate_exp[(size_t)call * (C - 2) + (c - 2)]) * (1.0 / (double)(1 << 23)); // :265
        double f = mutation_rate * P.genes[i].span;                                                                   // :290
        m = r * f;                                                                                                    // :293
    }
    mtab[idx] = m;
}

// per-warp shared-memory carve-up (doubles); every block is 16-byte aligned
struct FastSmem
{
    int n, T, G, nj;
    __host__ __device__ int off_rec() const { return 0; }                        // [n][4]  g0, base, clip_min, clip_max
    __host__ __device__ int off_term() const { return off_rec() + 4 * n; }       // [n][6]  pg[parity] * gradient_factor
    __host__ __device__ int off_delta() const { return off_term() + 6 * n; }     // [T][n][8]
    __host__ __device__ int off_par() const { return off_delta() + 8 * T * n; }  // [2 buffers][g0,g1,gr0,gr1][n]
    __host__ __device__ int off_pg() const { return off_par() + 8 * n; }         // [2][n] mix(gr0, gr1, fmix)
    __host__ __device__ int off_tip0() const { return off_pg() + 2 * n; }        // [T][8]
    __host__ __device__ int off_gp() const { return off_tip0() + 8 * T; }        // [G][12]
    __host__ __device__ int off_jrec() const { return off_gp() + 12 * G; }       // [n][4]  mid, halfspan, vel_weight, seed  (nj > 0)
    __host__ __device__ int off_fit() const { return off_jrec() + (nj ? 4 * n : 0); } // [256] primary fitness per child slot
    __host__ __device__ int off_sf() const { return off_fit() + 256; }                // [256] secondary fitness per child slot
    __host__ __device__ int total() const { return ((off_sf() + 256) + 1) & ~1; }
};

template <int T> BIOIK_HD void select_frame(const double (&F)[T][7], int tip, double* f)
{
#pragma unroll
    for(int k = 0; k < 7; k++) f[k] = F[0][k];
#pragma unroll
    for(int t = 1; t < T; t++)
        if(tip == t)
        {
#pragma unroll
            for(int k = 0; k < 7; k++) f[k] = F[t][k];
        }
}

// Goal::evaluate for the link goals, on a frame held in registers (same bodies as goal_value)
BIOIK_HD double link_goal_value(int type, const double* p, const double* f)
{
    switch(type)
    {
    case G_POSITION: return len2(p[0] - f[0], p[1] - f[1], p[2] - f[2]);
    case G_ORIENTATION: return BIOIK_FMIN(qlen2(p[3] - f[3], p[4] - f[4], p[5] - f[5], p[6] - f[6]), qlen2(p[3] + f[3], p[4] + f[4], p[5] + f[5], p[6] + f[6]));
    case G_POSE:
    {
        double e = 0.0;
        e += len2(p[0] - f[0], p[1] - f[1], p[2] - f[2]);
        e += BIOIK_FMIN(qlen2(p[3] - f[3], p[4] - f[4], p[5] - f[5], p[6] - f[6]), qlen2(p[3] + f[3], p[4] + f[4], p[5] + f[5], p[6] + f[6])) * (p[7] * p[7]);
        return e;
    }
    case G_LOOK_AT:
    {
        V3 axis = quat_mul_vec(Q4{f[3], f[4], f[5], f[6]}, V3{p[0], p[1], p[2]});
        double ax = p[3] - f[0], ay = p[4] - f[1], az = p[5] - f[2];
        double sa = 1.0 / BIOIK_SQRT(len2(ax, ay, az));
        ax = ax * sa; ay = ay * sa; az = az * sa;
        double sb = 1.0 / BIOIK_SQRT(len2(axis.x, axis.y, axis.z));
        double bx = axis.x * sb, by = axis.y * sb, bz = axis.z * sb;
        return len2(bx - ax, by - ay, bz - az);
    }
    case G_MAX_DISTANCE:
    {
        double d = BIOIK_FMAX(0.0, BIOIK_SQRT(len2(p[0] - f[0], p[1] - f[1], p[2] - f[2])) - p[3]);
        return d * d;
    }
    case G_MIN_DISTANCE:
    {
        double d = BIOIK_FMAX(0.0, p[3] - BIOIK_SQRT(len2(p[0] - f[0], p[1] - f[1], p[2] - f[2])));
        return d * d;
    }
    case G_LINE:
    {
        double rx = f[0] - p[0], ry = f[1] - p[1], rz = f[2] - p[2];
        double k = p[3] * rx + p[4] * ry + p[5] * rz;
        double qx = f[0] - p[3] * k, qy = f[1] - p[4] * k, qz = f[2] - p[5] * k;
        return len2(qx - p[0], qy - p[1], qz - p[2]);
    }
    case G_PLANE:
    {
        double sd = (f[0] - p[0]) * p[3] + (f[1] - p[1]) * p[4] + (f[2] - p[2]) * p[5];
        return sd * sd;
    }
    case G_SIDE:
    {
        V3 v = quat_mul_vec(Q4{f[3], f[4], f[5], f[6]}, V3{p[0], p[1], p[2]});
        double s = BIOIK_FMAX(0.0, v.x * p[3] + v.y * p[4] + v.z * p[5]);
        return s * s;
    }
    case G_DIRECTION:
    {
        V3 v = quat_mul_vec(Q4{f[3], f[4], f[5], f[6]}, V3{p[0], p[1], p[2]});
        return len2(p[3] - v.x, p[4] - v.y, p[5] - v.z);
    }
    default: return 0.0;
    }
}

BIOIK_HD bool is_joint_goal(int type) { return type >= G_AVOID_JOINT_LIMITS && type <= G_JOINT_VARIABLE; }

// one gene's contribution to a joint-space goal (goal_types.h:387-465,494-498), gene order = loop order
BIOIK_HD void joint_goal_accumulate(int type, int var_index, int i, double x, double clip_max, double mid, double halfspan, double vw, double seedv, double p0, double& acc)
{
    switch(type)
    {
    case G_AVOID_JOINT_LIMITS:
        if(clip_max != DBLMAX)
        {
            double d = x - mid;
            d = BIOIK_FMAX(0.0, BIOIK_FABS(d) * 2.0 - halfspan);
            d *= vw;
            acc += d * d;
        }
        break;
    case G_CENTER_JOINTS:
        if(clip_max != DBLMAX)
        {
            double d = x - mid;
            d *= vw;
            acc += d * d;
        }
        break;
    case G_REGULARIZATION:
    {
        double d = x - seedv;
        acc += d * d;
        break;
    }
    case G_MINIMAL_DISPLACEMENT:
    {
        double d = x - seedv;
        d *= vw;
        acc += d * d;
        break;
    }
    case G_JOINT_VARIABLE:
        if(var_index == i)
        {
            double d = p0 - x;
            acc = d * d;
        }
        break;
    default: break;
    }
}

#ifndef BIOIK_FAST_HELPERS_DEFINED
#define BIOIK_FAST_HELPERS_DEFINED
__device__ __forceinline__ uint64_t fast_fitness_key(double f) { return (f != f) ? 0xFFFFFFFFFFFFFFFFull : (uint64_t)__double_as_longlong(f); }
__device__ __forceinline__ void fast_warp_argmin(uint64_t& key, int& pos, int& child)
{
#pragma unroll
    for(int o = 16; o > 0; o >>= 1)
    {
        uint64_t k2 = __shfl_xor_sync(0xffffffffu, key, o);
        int p2 = __shfl_xor_sync(0xffffffffu, pos, o);
        int c2 = __shfl_xor_sync(0xffffffffu, child, o);
        if(k2 < key || (k2 == key && p2 < pos))
        {
            key = k2;
            pos = p2;
            child = c2;
        }
    }
}
#endif

constexpr int FAST_MAX_CPL = 8; // population <= 256

// T = tips, CH = children evaluated together per lane (register block), JOINT = joint-space goals present
template <int T, int CH, bool JOINT> __global__ void __launch_bounds__(128) k_evolve_fast(const DProblem* __restrict__ Pp, DState S, int step, const double* __restrict__ mtab)
{
    extern __shared__ double smem[];
    const DProblem& P = *Pp;
    const int lane = threadIdx.x & 31;
    const int warp_in_block = threadIdx.x >> 5;
    const int task = blockIdx.x * (blockDim.x >> 5) + warp_in_block;
    if(task >= S.B * 2) return;
    const int q = task >> 1, slot = task & 1;
    if(S.done[q]) return;
    const int n = P.n, C = S.C, G = P.G;
    const int nchunks = (C + 32 * CH - 1) / (32 * CH);

    // joint-space goals handled in the gene loop (host guarantees <= FAST_MAX_JOINT_GOALS)
    int nj = 0, jg_goal[FAST_MAX_JOINT_GOALS], jg_type[FAST_MAX_JOINT_GOALS], jg_var[FAST_MAX_JOINT_GOALS];
    if(JOINT)
        for(int g = 0; g < G; g++)
            if(is_joint_goal(P.goals[g].type) && nj < FAST_MAX_JOINT_GOALS)
            {
                jg_goal[nj] = g;
                jg_type[nj] = P.goals[g].type;
                jg_var[nj] = P.goals[g].var_index;
                nj++;
            }

    FastSmem L{n, T, G, JOINT ? 1 : 0};
    double* W = smem + (size_t)warp_in_block * L.total();
    double *s_rec = W + L.off_rec(), *s_term = W + L.off_term(), *s_delta = W + L.off_delta(), *s_par = W + L.off_par(), *s_pg = W + L.off_pg();
    double *s_tip0 = W + L.off_tip0(), *s_gp = W + L.off_gp(), *s_jrec = W + L.off_jrec(), *s_fit = W + L.off_fit(), *s_sf = W + L.off_sf();
    const double* seed = S.seeds + (size_t)q * P.n_vars;

    // ---- stage the task -------------------------------------------------------------------
    for(int k = lane; k < T * n * 7; k += 32)
    {
        int ti = k / 7, c7 = k - ti * 7;
        s_delta[ti * 8 + c7] = S.delta[(size_t)task * T * n * 7 + k];
    }
    for(int k = lane; k < T * 7; k += 32) s_tip0[(k / 7) * 8 + (k % 7)] = S.tip0[(size_t)task * T * 7 + k];
    for(int i = lane; i < n; i += 32)
    {
        const DGene& Gn = P.genes[i];
        s_rec[4 * i + 1] = S.base[(size_t)task * n + i];
        s_rec[4 * i + 2] = Gn.clip_min;
        s_rec[4 * i + 3] = Gn.clip_max;
        s_par[0 * n + i] = S.genes[((size_t)task * 2 + 0) * n + i];
        s_par[1 * n + i] = S.genes[((size_t)task * 2 + 1) * n + i];
        s_par[2 * n + i] = S.grads[((size_t)task * 2 + 0) * n + i];
        s_par[3 * n + i] = S.grads[((size_t)task * 2 + 1) * n + i];
        if(JOINT)
        {
            s_jrec[4 * i + 0] = (Gn.vmin + Gn.vmax) * 0.5;
            s_jrec[4 * i + 1] = Gn.span * 0.5;
            s_jrec[4 * i + 2] = Gn.vel_weight;
            s_jrec[4 * i + 3] = seed[Gn.var];
        }
    }
    for(int k = lane; k < G * GOAL_NPARAM; k += 32) s_gp[k] = S.goal_params[(size_t)q * G * GOAL_NPARAM + k];
    __syncwarp();

    int cur = 0; // parent buffer in use
    const int parity = lane & 1; // child slot c = lane + 32k is even <=> lane is even

    for(int gen = 0; gen < S.gens; gen++)
    {
        const int call = (step * 2 + slot) * S.gens + gen;
        const double* mt = mtab + (size_t)call * n * C;
        const int child_count = P.has_secondary ? S.ccount[((size_t)q * 2 + slot) * S.gens + gen] : C;
        double* par = s_par + cur * 4 * n;
        const double *p_g0 = par, *p_g1 = par + n, *p_gr0 = par + 2 * n, *p_gr1 = par + 3 * n;

        // per-generation warp-uniform tables: pg = mix(gr0, gr1, fmix), term = pg * gradient_factor  (:268-269,:294-295)
        for(int i = lane; i < n; i += 32)
        {
            double a = p_gr0[i], b = p_gr1[i];
            double pge = mix(a, b, 0.2); // child_index even: fmix = 1 * 0.2
            double pgo = mix(a, b, 0.0); // child_index odd:  fmix = 0 * 0.2
            s_pg[i] = pge;
            s_pg[n + i] = pgo;
            s_term[6 * i + 0] = pge * 0.0;
            s_term[6 * i + 1] = pge * 1.0;
            s_term[6 * i + 2] = pge * 2.0;
            s_term[6 * i + 3] = pgo * 0.0;
            s_term[6 * i + 4] = pgo * 1.0;
            s_term[6 * i + 5] = pgo * 2.0;
            s_rec[4 * i + 0] = p_g0[i];
        }
        __syncwarp();

        for(int chunk = 0; chunk < nchunks; chunk++)
        {
            const int cbase = lane + 32 * CH * chunk;
            double F[CH][T][7];
            double acc[JOINT ? CH : 1][FAST_MAX_JOINT_GOALS];
            int tsel[CH];
#pragma unroll
            for(int k = 0; k < CH; k++)
            {
                int c = cbase + 32 * k;
                tsel[k] = (parity ? 3 : 0) + (c % 3); // term column: fmix class x gradient_factor (c % 3)
#pragma unroll
                for(int t = 0; t < T; t++)
#pragma unroll
                    for(int j = 0; j < 7; j++) F[k][t][j] = s_tip0[8 * t + j];
                if(JOINT)
#pragma unroll
                    for(int j = 0; j < FAST_MAX_JOINT_GOALS; j++) acc[k][j] = 0.0;
            }

#pragma unroll 1
            for(int i = 0; i < n; i++)
            {
                const double g0 = s_rec[4 * i + 0], base = s_rec[4 * i + 1], lo = s_rec[4 * i + 2], hi = s_rec[4 * i + 3];
                double d[CH], x[CH];
#pragma unroll
                for(int k = 0; k < CH; k++)
                {
                    int c = cbase + 32 * k;
                    double m = (c < C) ? BIOIK_LDG(mt + (size_t)i * C + c) : 0.0;
                    double gene = g0;
                    gene += m;                       // gene += r * f      (:293)
                    gene += s_term[6 * i + tsel[k]]; // gene += gradient   (:296)
                    gene = clampd(gene, lo, hi);     // :297
                    if(c == 0) gene = g0;            // slots 0 and 1 carry the parents unchanged (:381-388)
                    if(c == 1) gene = p_g1[i];
                    x[k] = gene;
                    d[k] = gene - base; // :1086
                }
                const int tmask = P.genes[i].tipmask; // tips this gene can move (structural); others have an all-zero delta frame
#pragma unroll
                for(int t = 0; t < T; t++)
                {
                    if(!((tmask >> t) & 1)) continue; // fma(d, 0, F) == F
                    const double* D = s_delta + ((size_t)t * n + i) * 8;
                    const double D0 = D[0], D1 = D[1], D2 = D[2], D3 = D[3], D4 = D[4], D5 = D[5], D6 = D[6];
#pragma unroll
                    for(int k = 0; k < CH; k++)
                    {
                        F[k][t][0] = BIOIK_FMA(d[k], D0, F[k][t][0]);
                        F[k][t][1] = BIOIK_FMA(d[k], D1, F[k][t][1]);
                        F[k][t][2] = BIOIK_FMA(d[k], D2, F[k][t][2]);
                        F[k][t][3] = BIOIK_FMA(d[k], D3, F[k][t][3]);
                        F[k][t][4] = BIOIK_FMA(d[k], D4, F[k][t][4]);
                        F[k][t][5] = BIOIK_FMA(d[k], D5, F[k][t][5]);
                        F[k][t][6] = BIOIK_FMA(d[k], D6, F[k][t][6]);
                    }
                }
                if(JOINT)
                {
                    const double mid = s_jrec[4 * i + 0], halfspan = s_jrec[4 * i + 1], vw = s_jrec[4 * i + 2], seedv = s_jrec[4 * i + 3];
#pragma unroll
                    for(int j = 0; j < FAST_MAX_JOINT_GOALS; j++)
                        if(j < nj)
                        {
                            const double p0 = s_gp[jg_goal[j] * GOAL_NPARAM];
#pragma unroll
                            for(int k = 0; k < CH; k++) joint_goal_accumulate(jg_type[j], jg_var[j], i, x[k], hi, mid, halfspan, vw, seedv, p0, acc[k][j]);
                        }
                }
            }

            // fitness: weighted sum in goal order (src/problem.cpp:251-257)
#pragma unroll
            for(int k = 0; k < CH; k++)
            {
                const int c = cbase + 32 * k;
                double prim = 0.0, sec = 0.0;
                int jn = 0;
                for(int g = 0; g < G; g++)
                {
                    const DGoal& gl = P.goals[g];
                    double v;
                    if(JOINT && is_joint_goal(gl.type))
                    {
                        v = 0.0;
#pragma unroll
                        for(int j = 0; j < FAST_MAX_JOINT_GOALS; j++)
                            if(j == jn) v = acc[k][j];
                        if(gl.type == G_JOINT_VARIABLE && gl.var_index < 0)
                        {
                            double dd = s_gp[g * GOAL_NPARAM] - seed[-1 - gl.var_index];
                            v = dd * dd;
                        }
                        jn++;
                    }
                    else
                    {
                        double f[7];
                        select_frame<T>(F[k], gl.secondary ? 0 : gl.tip, f);
                        if(gl.secondary)
                        {
                            // secondary goals see null_tip_frames (identity), src/ik_base.h:163
                            f[0] = f[1] = f[2] = f[3] = f[4] = f[5] = 0.0;
                            f[6] = 1.0;
                        }
                        v = link_goal_value(gl.type, s_gp + g * GOAL_NPARAM, f);
                    }
                    if(gl.secondary)
                        sec += v * gl.weight_sq;
                    else
                        prim += v * gl.weight_sq;
                }
                if(c < C)
                {
                    s_fit[c] = prim;
                    s_sf[c] = sec;
                }
            }
        }
        __syncwarp();

        // ---- pre-selection positions (:366-378) and this lane's candidates ---------------------------
        double fit[FAST_MAX_CPL];
        int posn[FAST_MAX_CPL];
#pragma unroll
        for(int k = 0; k < FAST_MAX_CPL; k++)
        {
            int c = lane + 32 * k;
            fit[k] = 0.0;
            posn[k] = 0x7fffffff;
            if(c >= C) continue;
            fit[k] = s_fit[c];
            posn[k] = c;
            if(P.has_secondary && c >= 2)
            {
                double mine = s_sf[c];
                int rank = 0;
                for(int o = 2; o < C; o++)
                {
                    double other = s_sf[o];
                    rank += (other < mine || (other == mine && o < c)) ? 1 : 0;
                }
                posn[k] = (2 + rank < child_count) ? 2 + rank : 0x7fffffff;
            }
        }

        // ---- selection (:410-431) -------------------------------------------------------------------
        uint64_t key = 0xFFFFFFFFFFFFFFFFull;
        int bpos = 0x7fffffff, bchild = -1;
#pragma unroll
        for(int k = 0; k < FAST_MAX_CPL; k++)
        {
            uint64_t kk = fast_fitness_key(fit[k]);
            if(posn[k] != 0x7fffffff && (kk < key || (kk == key && posn[k] < bpos)))
            {
                key = kk;
                bpos = posn[k];
                bchild = lane + 32 * k;
            }
        }
        fast_warp_argmin(key, bpos, bchild);
        double f_pos0 = __shfl_sync(0xffffffffu, fit[0], 0);
        double f_pos1 = __shfl_sync(0xffffffffu, fit[0], 1);
        int w1_pos = bpos, w1_child = bchild;
        if(f_pos0 != f_pos0)
        {
            w1_pos = 0;
            w1_child = 0;
        }
        key = 0xFFFFFFFFFFFFFFFFull;
        bpos = 0x7fffffff;
        bchild = -1;
#pragma unroll
        for(int k = 0; k < FAST_MAX_CPL; k++)
        {
            int c = lane + 32 * k;
            if(posn[k] == 0x7fffffff || c == w1_child) continue;
            int p = (posn[k] == 0) ? w1_pos : posn[k];
            uint64_t kk = fast_fitness_key(fit[k]);
            if(kk < key || (kk == key && p < bpos))
            {
                key = kk;
                bpos = p;
                bchild = c;
            }
        }
        fast_warp_argmin(key, bpos, bchild);
        int w2_child = bchild;
        {
            int occ1_child = (w1_pos == 1) ? 0 : 1;
            double f_occ1 = (occ1_child == 0) ? f_pos0 : f_pos1;
            if(f_occ1 != f_occ1) w2_child = occ1_child;
        }

        // ---- new parents into the other buffer (lanes 0/1 re-derive the winners) -------------------
        double* nxt = s_par + (cur ^ 1) * 4 * n;
        if(lane < 2)
        {
            const int wc = (lane == 0) ? w1_child : w2_child;
            double* og = nxt + lane * n;        // genes of individuals[lane]
            double* ogr = nxt + (2 + lane) * n; // gradients
            if(wc < 2)
            {
                const double* sg = wc == 0 ? p_g0 : p_g1;
                const double* sgr = wc == 0 ? p_gr0 : p_gr1;
                for(int i = 0; i < n; i++)
                {
                    og[i] = sg[i];
                    ogr[i] = sgr[i];
                }
            }
            else
            {
                const int wpar = wc & 1; // 1 = odd child
                const int col = (wpar ? 3 : 0) + (wc % 3);
                for(int i = 0; i < n; i++)
                {
                    double g0 = p_g0[i];
                    double gene = g0;
                    gene += BIOIK_LDG(mt + (size_t)i * C + wc);
                    gene += s_term[6 * i + col];
                    gene = clampd(gene, s_rec[4 * i + 2], s_rec[4 * i + 3]);
                    og[i] = gene;
                    ogr[i] = mix(s_pg[wpar * n + i], gene - g0, 0.3); // :299
                }
            }
        }
        __syncwarp();
        cur ^= 1;
    }

    double* par = s_par + cur * 4 * n;
    for(int i = lane; i < n; i += 32)
    {
        S.genes[((size_t)task * 2 + 0) * n + i] = par[i];
        S.genes[((size_t)task * 2 + 1) * n + i] = par[n + i];
        S.grads[((size_t)task * 2 + 0) * n + i] = par[2 * n + i];
        S.grads[((size_t)task * 2 + 1) * n + i] = par[3 * n + i];
    }
}

typedef void (*EvolveFastKernel)(const DProblem*, DState, int, const double*);

// picks the instantiation for (tips, population, joint goals); returns nullptr if the generic kernel must be used
inline EvolveFastKernel select_evolve_fast(int T, int C, int n_joint_goals)
{
    if(T < 1 || T > 8 || n_joint_goals > FAST_MAX_JOINT_GOALS || C > 32 * FAST_MAX_CPL) return nullptr;
    const bool J = n_joint_goals > 0;
    const int cpl = (C + 31) / 32;
#define BIOIK_PICK(TT, CC) (J ? (EvolveFastKernel)k_evolve_fast<TT, CC, true> : (EvolveFastKernel)k_evolve_fast<TT, CC, false>)
    switch(T)
    {
    case 1: return cpl >= 3 ? BIOIK_PICK(1, 4) : (cpl == 2 ? BIOIK_PICK(1, 2) : BIOIK_PICK(1, 1));
    case 2: return cpl >= 2 ? BIOIK_PICK(2, 2) : BIOIK_PICK(2, 1);
    case 3: return cpl >= 2 ? BIOIK_PICK(3, 2) : BIOIK_PICK(3, 1);
    case 4: return BIOIK_PICK(4, 1);
    case 5: return BIOIK_PICK(5, 1);
    case 6: return BIOIK_PICK(6, 1);
    case 7: return BIOIK_PICK(7, 1);
    default: return BIOIK_PICK(8, 1);
    }
#undef BIOIK_PICK
}

} // namespace bioik
