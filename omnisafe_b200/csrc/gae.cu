// Dual (reward + cost) GAE as a segmented reverse inclusive scan -- sm_100a.
//
// Replaces the per-path Python recursions of the reference:
//   omnisafe/common/buffer/onpolicy_buffer.py:L148-203 (finish_path)
//   omnisafe/common/buffer/onpolicy_buffer.py:L299-303 (gae branch)
//   omnisafe/utils/math.py:L59-82                      (discount_cumsum, fp64 carry)
//   omnisafe/common/buffer/vector_onpolicy_buffer.py:L125-136 + utils/distributed.py:L382-388
//       (advantage statistics; produced here as fp64 partial sums in the epilogue)
//
// Layout: time-major SoA slabs [T][N] (env index contiguous) so that the rollout kernel appends a
// step with coalesced stores and this kernel reads rows with coalesced 128 B requests.
//
// Parallelisation: block = 16 envs (x) x 32 time-chunks (y) x 4 steps per thread (512 threads).  Every thread
// folds its 4 steps into an affine map  A_in -> b + a*A_in  (fp64), the 32 chunk maps of one env are
// combined with a warp-shuffle suffix scan (after a shared-memory transpose so that lanes run along
// time), and a second local pass replays the reference's sequential arithmetic
// (separately rounded fp64 mul/add) from the exact carry-in.  Tiles of 128 steps are walked from
// the end of the horizon to its start with a per-env carry.
#include "common.cuh"
#include <stdlib.h>

namespace osb {

constexpr int GE = 16;         // envs per block
constexpr int GC = 32;         // chunks per tile
constexpr int GL = 4;          // steps per chunk
constexpr int GT = GC * GL;    // steps per tile
constexpr int GPAD = GE + 1;
constexpr int GTHREADS = GE * GC;   // 512

struct GaeArgs {
    const float* rew;
    const float* cost;
    const float* val_r;
    const float* val_c;
    const uint8_t* flags;
    const float* boot_r;
    const float* boot_c;
    float* adv_r;
    float* adv_c;
    float* tv_r;
    float* tv_c;
    float* disc_ret;  // may be null
    double* partials; // [gridDim.x][4]
    double* sums;     // [4] written by the last block
    unsigned int* ticket;
    int T, N;
    float gamma_f;    // (float)gamma : fp32 delta arithmetic (onpolicy_buffer.py:L301)
    float pen;        // penalty_coefficient (onpolicy_buffer.py:L185)
    double g;         // gamma            (discounted return)
    double gl_r;      // gamma * lam      (python double product)
    double gl_c;      // gamma * lam_c
};

struct GaeTile {      // one thread's 4 steps (+ the value after them) of one env
    float r[GL], c[GL], vr[GL + 1], vc[GL + 1];
    unsigned f[GL];
};

__device__ __forceinline__ void gae_load_tile(const GaeArgs& p, int env, bool env_ok, int t0, GaeTile& d) {
    const int N = p.N, T = p.T;
#pragma unroll
    for (int i = 0; i < GL; ++i) {
        const int t = t0 + i;
        const bool ok = env_ok && t >= 0;
        const size_t idx = (size_t)(ok ? t : 0) * N + (env_ok ? env : 0);
        d.r[i] = ok ? __ldg(p.rew + idx) : 0.f;
        d.c[i] = ok ? __ldg(p.cost + idx) : 0.f;
        d.vr[i] = ok ? __ldg(p.val_r + idx) : 0.f;
        d.vc[i] = ok ? __ldg(p.val_c + idx) : 0.f;
        d.f[i] = ok ? (unsigned)__ldg(p.flags + idx) : 0u;
    }
    const int t = t0 + GL;
    const bool ok = env_ok && t >= 0 && t < T;
    const size_t idx = (size_t)(ok ? t : 0) * N + (env_ok ? env : 0);
    d.vr[GL] = ok ? __ldg(p.val_r + idx) : 0.f;
    d.vc[GL] = ok ? __ldg(p.val_c + idx) : 0.f;
}

// MULTI = more than one 128-step tile: prefetches the next tile (more registers, 1 CTA per SM);
// the single-tile instantiation fits 64 registers so that two CTAs share an SM.
// EST = advantage estimator (onpolicy_buffer.py:L299-331):
//   0 'gae'      adv = scan(delta, gamma*lam),  target = adv + V                    (L299-303)
//   1 'gae-rtg'  adv = scan(delta, gamma*lam),  target = discount_cumsum(rewards)   (L305-310)
//   2 'plain'    adv = delta,                   target = discount_cumsum(rewards)   (L328-331)
//   3 'vtrace'   on-policy V-trace (behaviour == target policy, so rho = c = 1, L312-326, L338-405):
//                v_t = V_t + delta_t + gamma (v_{t+1} - V_{t+1}),  adv_t = r_t + gamma v_{t+1} - V_t,
//                target = v.  The scan carries e_t = v_t - V_t (the lambda = 1 GAE recurrence) in fp64;
//                the replay pass redoes the reference's fp32 operations from that carry.
// where `rewards` is the penalised reward path INCLUDING its bootstrap slot (finish_path subtracts
// penalty * costs in place, L185) and the cost targets use the same gamma (L192-196).
// RET = produce discounted_ret (third scan); training keeps no discounted_ret, so the hot instantiation drops it.
template <bool MULTI, int EST, bool RET>
__global__ void __launch_bounds__(GTHREADS, (MULTI || EST == 1 || EST == 3) ? 1 : 2) gae_dual_kernel(GaeArgs p) {
    constexpr int NQ = (EST == 0 || EST == 3) ? 3 : 4;    // scanned quantities: adv_r, adv_c, reward-to-go, cost-to-go
    __shared__ double sa[NQ * GC * GPAD];
    __shared__ double sb[NQ * GC * GPAD];
    __shared__ double carry[NQ * GE];
    __shared__ double red[3 * (GTHREADS / 32)];
    __shared__ int s_last;

    const int x = threadIdx.x, y = threadIdx.y;     // env lane (16), time chunk (32)
    const int lin = y * GE + x;
    const int tw = lin >> 5, tl = lin & 31;          // transposed role: warp tw <-> env tw, lane tl <-> chunk
    const int env = blockIdx.x * GE + x;
    const bool env_ok = env < p.N;
    const int N = p.N, T = p.T;
    const int ntiles = (T + GT - 1) / GT;

    if (lin < NQ * GE) carry[lin] = 0.0;
    constexpr bool want_g = (EST == 1 || EST == 2) || RET;

    double st_r = 0.0, st_r2 = 0.0, st_c = 0.0;
    GaeTile cur;
    gae_load_tile(p, env, env_ok, T - GT + y * GL, cur);
    __syncthreads();

    for (int k = 0; k < ntiles; ++k) {
        const int t0 = T - (k + 1) * GT + y * GL;  // first step of my chunk (may be < 0)
        if (!MULTI && k > 0) gae_load_tile(p, env, env_ok, t0, cur);   // no prefetch in the 2-CTA/SM variant
        // fp32 deltas with the reference's three separately rounded ops; bootstrap at path ends.
        float dr[GL], dc[GL], bootr[GL], bootc[GL], r[GL], c[GL], vr[GL], vc[GL];
        bool end[GL], valid[GL];
#pragma unroll
        for (int i = 0; i < GL; ++i) {
            const int t = t0 + i;
            valid[i] = env_ok && t >= 0;
            end[i] = valid[i] && (cur.f[i] != 0u || t == T - 1);
            float nr = cur.vr[i + 1], nc = cur.vc[i + 1];
            if (end[i]) {
                const bool term = (cur.f[i] & OSB_FLAG_TERMINATED) != 0u;
                const size_t idx = (size_t)t * N + env;
                nr = term ? 0.f : __ldg(p.boot_r + idx);
                nc = term ? 0.f : __ldg(p.boot_c + idx);
            }
            const float rp = __fadd_rn(cur.r[i], -__fmul_rn(p.pen, cur.c[i]));
            // reward-to-go estimators run on the penalised path, bootstrap slot included
            bootr[i] = (EST == 0 || EST == 3) ? nr : __fadd_rn(nr, -__fmul_rn(p.pen, nc));
            bootc[i] = nc;
            dr[i] = __fadd_rn(__fadd_rn(rp, __fmul_rn(p.gamma_f, nr)), -cur.vr[i]);
            dc[i] = __fadd_rn(__fadd_rn(cur.c[i], __fmul_rn(p.gamma_f, nc)), -cur.vc[i]);
            r[i] = (EST == 0 || EST == 3) ? cur.r[i] : rp; c[i] = cur.c[i]; vr[i] = cur.vr[i]; vc[i] = cur.vc[i];
        }
        const float cur_vr_last = cur.vr[GL], cur_vc_last = cur.vc[GL];
        float nvr[GL], nvc[GL];   // V_{t+1} inside the path (V-trace replay)
#pragma unroll
        for (int i = 0; i < GL; ++i) { nvr[i] = cur.vr[i + 1]; nvc[i] = cur.vc[i + 1]; }
        // prefetch the next (earlier) tile while this one is scanned
        if (MULTI && k + 1 < ntiles) gae_load_tile(p, env, env_ok, t0 - GT, cur);
        // pass 1: fold the chunk into affine maps (a, b) per quantity.
        const double glr = (EST == 3) ? p.g : p.gl_r, glc = (EST == 3) ? p.g : p.gl_c;
        double ar = 1.0, br = 0.0, ac = 1.0, bc = 0.0, ag = 1.0, bg = 0.0, bh = 0.0;   // (ag, bh): cost-to-go
#pragma unroll
        for (int i = GL - 1; i >= 0; --i) {
            if (valid[i]) {
                if (end[i]) {
                    ar = 0.0; br = (double)dr[i];
                    ac = 0.0; bc = (double)dc[i];
                    if (want_g) { ag = 0.0; bg = (double)r[i] + p.g * (double)bootr[i]; }
                    if (EST == 1 || EST == 2) bh = (double)c[i] + p.g * (double)bootc[i];
                } else {
                    br = (double)dr[i] + glr * br; ar = glr * ar;
                    bc = (double)dc[i] + glc * bc; ac = glc * ac;
                    if (want_g) { bg = (double)r[i] + p.g * bg; ag = p.g * ag; }
                    if (EST == 1 || EST == 2) bh = (double)c[i] + p.g * bh;
                }
            }
        }
        sa[(0 * GC + y) * GPAD + x] = ar; sb[(0 * GC + y) * GPAD + x] = br;
        sa[(1 * GC + y) * GPAD + x] = ac; sb[(1 * GC + y) * GPAD + x] = bc;
        if (want_g) { sa[(2 * GC + y) * GPAD + x] = ag; sb[(2 * GC + y) * GPAD + x] = bg; }
        if (EST == 1 || EST == 2) { sa[(3 * GC + y) * GPAD + x] = ag; sb[(3 * GC + y) * GPAD + x] = bh; }
        __syncthreads();
        // transposed role: warp tw owns env tw of the block, lane tl = chunk index.
#pragma unroll
        for (int q = (EST == 2 ? 2 : 0); q < NQ; ++q) {
            if (q == 2 && !want_g) continue;
            double a = sa[(q * GC + tl) * GPAD + tw];
            double b = sb[(q * GC + tl) * GPAD + tw];
#pragma unroll
            for (int off = 1; off < GC; off <<= 1) {
                const double a2 = __shfl_down_sync(0xffffffffu, a, off);
                const double b2 = __shfl_down_sync(0xffffffffu, b, off);
                if (tl + off < GC) { b = b + a * b2; a = a * a2; }
            }
            const double cin = carry[q * GE + tw];
            const double full = b + a * cin;              // value at the first step of chunk tl
            double ain = __shfl_down_sync(0xffffffffu, full, 1);
            if (tl == GC - 1) ain = cin;                  // last chunk takes the tile carry
            __syncwarp();
            sb[(q * GC + tl) * GPAD + tw] = ain;
            if (tl == 0) carry[q * GE + tw] = full;
        }
        __syncthreads();
        // pass 2: replay sequentially from the exact carry-in with the reference's roundings.
        double Ar = sb[(0 * GC + y) * GPAD + x];
        double Ac = sb[(1 * GC + y) * GPAD + x];
        double Ag = want_g ? sb[(2 * GC + y) * GPAD + x] : 0.0;
        double Ah = (EST == 1 || EST == 2) ? sb[(3 * GC + y) * GPAD + x] : 0.0;
        // V-trace replay state: v_{t+1} of the step after this chunk (fp32, = V + e from the scan)
        float lv_r = 0.f, lv_c = 0.f;
        if (EST == 3) {
            lv_r = (float)((double)cur_vr_last + Ar);
            lv_c = (float)((double)cur_vc_last + Ac);
        }
#pragma unroll
        for (int i = GL - 1; i >= 0; --i) {
            if (EST == 3) {
                if (valid[i]) {
                    // values[index + 1] and last_v_s: the bootstrap slot at a path end, else V_{t+1} / v_{t+1}
                    const float vnr = end[i] ? bootr[i] : nvr[i], vnc = end[i] ? bootc[i] : nvc[i];
                    const float lr = end[i] ? bootr[i] : lv_r, lc = end[i] ? bootc[i] : lv_c;
                    const size_t idx = (size_t)(t0 + i) * N + env;
                    // policy_advantage = clip_rho * (r + gamma * v_{t+1} - V_t)        (L402-403)
                    const float rp = __fadd_rn(r[i], -__fmul_rn(p.pen, c[i]));
                    const float o_ar = __fadd_rn(__fadd_rn(rp, __fmul_rn(p.gamma_f, lr)), -vr[i]);
                    const float o_ac = __fadd_rn(__fadd_rn(c[i], __fmul_rn(p.gamma_f, lc)), -vc[i]);
                    // v_s[t] += delta + gamma * c * (last_v_s - V_{t+1})                 (L394-397)
                    lv_r = __fadd_rn(vr[i], __fadd_rn(dr[i], __fmul_rn(p.gamma_f, __fadd_rn(lr, -vnr))));
                    lv_c = __fadd_rn(vc[i], __fadd_rn(dc[i], __fmul_rn(p.gamma_f, __fadd_rn(lc, -vnc))));
                    Ag = end[i] ? __dadd_rn((double)r[i], __dmul_rn(p.g, (double)bootr[i]))
                                : __dadd_rn((double)r[i], __dmul_rn(p.g, Ag));
                    p.adv_r[idx] = o_ar; p.adv_c[idx] = o_ac;
                    p.tv_r[idx] = lv_r;  p.tv_c[idx] = lv_c;
                    if (RET && p.disc_ret) p.disc_ret[idx] = (float)Ag;
                    st_r += (double)o_ar; st_r2 += (double)o_ar * (double)o_ar; st_c += (double)o_ac;
                }
            } else if (valid[i]) {
                if (end[i]) {
                    Ar = (double)dr[i];
                    Ac = (double)dc[i];
                    if (want_g) Ag = __dadd_rn((double)r[i], __dmul_rn(p.g, (double)bootr[i]));
                    if (EST == 1 || EST == 2) Ah = __dadd_rn((double)c[i], __dmul_rn(p.g, (double)bootc[i]));
                } else {
                    Ar = __dadd_rn((double)dr[i], __dmul_rn(p.gl_r, Ar));
                    Ac = __dadd_rn((double)dc[i], __dmul_rn(p.gl_c, Ac));
                    if (want_g) Ag = __dadd_rn((double)r[i], __dmul_rn(p.g, Ag));
                    if (EST == 1 || EST == 2) Ah = __dadd_rn((double)c[i], __dmul_rn(p.g, Ah));
                }
                const size_t idx = (size_t)(t0 + i) * N + env;
                const float o_ar = (EST == 2) ? dr[i] : (float)Ar, o_ac = (EST == 2) ? dc[i] : (float)Ac;
                p.adv_r[idx] = o_ar;
                p.adv_c[idx] = o_ac;
                p.tv_r[idx] = (EST == 0) ? (float)(Ar + (double)vr[i]) : (float)Ag;
                p.tv_c[idx] = (EST == 0) ? (float)(Ac + (double)vc[i]) : (float)Ah;
                if (RET && p.disc_ret) p.disc_ret[idx] = (float)Ag;
                st_r += (double)o_ar;
                st_r2 += (double)o_ar * (double)o_ar;
                st_c += (double)o_ac;
            }
        }
        __syncthreads();
    }
    // epilogue: block partial sums for the advantage statistics (fixed order -> deterministic);
    // the last block to finish folds all partials into sums[4].
    st_r = warp_sum(st_r); st_r2 = warp_sum(st_r2); st_c = warp_sum(st_c);
    if (tl == 0) { red[0 * 16 + tw] = st_r; red[1 * 16 + tw] = st_r2; red[2 * 16 + tw] = st_c; }
    __syncthreads();
    if (lin == 0) {
        double a = 0, b = 0, c = 0;
        for (int w = 0; w < GTHREADS / 32; ++w) { a += red[w]; b += red[16 + w]; c += red[32 + w]; }
        double* o = p.partials + (size_t)blockIdx.x * 4;
        o[0] = a; o[1] = b; o[2] = c;
        const int nenv = min(GE, N - (int)blockIdx.x * GE);
        o[3] = (double)nenv * (double)T;
        __threadfence();
        s_last = (atomicAdd(p.ticket, 1u) == gridDim.x - 1) ? 1 : 0;
    }
    __syncthreads();
    if (s_last && lin < 32) {
        __threadfence();
        double acc[4] = {0, 0, 0, 0};
        for (int bb = lin; bb < (int)gridDim.x; bb += 32)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[q] += __ldcg(p.partials + (size_t)bb * 4 + q);
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[q] = warp_sum(acc[q]);
        if (lin == 0) {
            for (int q = 0; q < 4; ++q) p.sums[q] = acc[q];
            *p.ticket = 0u;
        }
    }
}

// sums[4] = {sum adv_r, sum adv_r^2, sum adv_c, count}; one warp, fixed order.
__global__ void gae_stats_reduce_kernel(const double* __restrict__ partials, int nblocks,
                                        double* __restrict__ sums) {
    double acc[4] = {0, 0, 0, 0};
    for (int b = threadIdx.x; b < nblocks; b += 32)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[q] += partials[(size_t)b * 4 + q];
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[q] = warp_sum(acc[q]);
    if (threadIdx.x == 0)
        for (int q = 0; q < 4; ++q) sums[q] = acc[q];
}

// moments[4] = {mean_r, std_r + 1e-8, mean_c, 1}  (vector_onpolicy_buffer.py:L131-136)
__global__ void adv_moments_kernel(const double* __restrict__ sums, int standardize_r,
                                   int standardize_c, float* __restrict__ moments) {
    if (threadIdx.x != 0) return;
    const double n = sums[3];
    const double mean_r = sums[0] / n;
    double var = sums[1] / n - mean_r * mean_r;
    if (var < 0.0) var = 0.0;
    const float mean_rf = (float)mean_r;
    const float std_rf = (float)sqrt(var);
    moments[0] = standardize_r ? mean_rf : 0.f;
    moments[1] = standardize_r ? __fadd_rn(std_rf, 1e-8f) : 1.f;
    moments[2] = standardize_c ? (float)(sums[2] / n) : 0.f;
    moments[3] = 1.f;
}

// materialise the standardised advantages (what VectorOnPolicyBuffer.get() returns)
__global__ void adv_standardize_kernel(const float* __restrict__ adv_r,
                                       const float* __restrict__ adv_c,
                                       const float* __restrict__ moments, size_t n,
                                       float* __restrict__ out_r, float* __restrict__ out_c) {
    const float m_r = moments[0], s_r = moments[1], m_c = moments[2];
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        out_r[i] = __fdiv_rn(__fadd_rn(adv_r[i], -m_r), s_r);
        out_c[i] = __fadd_rn(adv_c[i], -m_c);
    }
}

// Batched discount_cumsum (utils/math.py:L59-82): one warp per row, lanes along time,
// fp64 affine warp-shuffle suffix scan.  x: [rows][len] fp32 or fp64, out: [rows][len] fp64.
template <typename TIn>
__global__ void discount_cumsum_kernel(const TIn* __restrict__ x, int rows, int len,
                                       double discount, double* __restrict__ out) {
    const int row = blockIdx.x * (blockDim.x / 32) + (threadIdx.x / 32);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const int per = (len + 31) / 32;
    const int lo = lane * per, hi = min(len, lo + per);
    const TIn* xr = x + (size_t)row * len;
    double* orow = out + (size_t)row * len;
    double a = 1.0, b = 0.0;
    for (int i = hi - 1; i >= lo; --i) {
        if (i == len - 1) { b = (double)xr[i]; a = 0.0; }
        else { b = (double)xr[i] + discount * b; a = discount * a; }
    }
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const double a2 = __shfl_down_sync(0xffffffffu, a, off);
        const double b2 = __shfl_down_sync(0xffffffffu, b, off);
        if (lane + off < 32) { b = b + a * b2; a = a * a2; }
    }
    double cum = __shfl_down_sync(0xffffffffu, b, 1);
    if (lane == 31) cum = 0.0;
    for (int i = hi - 1; i >= lo; --i) {
        if (i == len - 1) cum = (double)xr[i];
        else cum = __dadd_rn((double)xr[i], __dmul_rn(discount, cum));
        orow[i] = cum;
    }
}

}  // namespace osb

using namespace osb;

extern "C" {

// partials [blocks][4] + one 8-byte ticket slot (zero-initialised by the caller, self-resetting)
int osb_gae_workspace_doubles(int n_envs) { return ((n_envs + GE - 1) / GE) * 4 + 8; }

int osb_adv_estimate(const float* rew, const float* cost, const float* val_r, const float* val_c,
                     const uint8_t* flags, const float* boot_r, const float* boot_c, int T, int N,
                     double gamma, double lam, double lam_c, double penalty_coef, int estimator,
                     float* adv_r, float* adv_c, float* tv_r, float* tv_c, float* disc_ret,
                     double* workspace, double* sums, void* stream) {
    OSB_CHECK_ARG(T > 0 && N > 0, "T, N must be positive");
    OSB_CHECK_ARG(estimator >= 0 && estimator <= 3, "estimator: 0 gae, 1 gae-rtg, 2 plain, 3 vtrace");
    OSB_CHECK_ARG(estimator == 0 || estimator == 3 || disc_ret == nullptr || penalty_coef == 0.0,
                  "reward-to-go estimators share the scan with discounted_ret: needs penalty_coef == 0 or disc_ret == NULL");
    OSB_CHECK_ARG(rew && cost && val_r && val_c && flags && boot_r && boot_c, "null input slab");
    OSB_CHECK_ARG(adv_r && adv_c && tv_r && tv_c && workspace && sums, "null output");
    GaeArgs a;
    a.rew = rew; a.cost = cost; a.val_r = val_r; a.val_c = val_c; a.flags = flags;
    a.boot_r = boot_r; a.boot_c = boot_c;
    a.adv_r = adv_r; a.adv_c = adv_c; a.tv_r = tv_r; a.tv_c = tv_c; a.disc_ret = disc_ret;
    a.partials = workspace;
    a.T = T; a.N = N;
    a.gamma_f = (float)gamma; a.pen = (float)penalty_coef;
    a.g = gamma; a.gl_r = gamma * lam; a.gl_c = gamma * lam_c;
    const int nblocks = (N + GE - 1) / GE;
    a.sums = sums;
    a.ticket = reinterpret_cast<unsigned int*>(workspace + (size_t)nblocks * 4 + 1);
    cudaStream_t s = (cudaStream_t)stream;
    // the 64-register instantiation (two CTAs per SM) wins at every T measured on B200: occupancy beats
    // the register-hungry prefetching variant, which is kept for experiments (OSB_GAE_PREFETCH=1)
    static const bool prefetch = getenv("OSB_GAE_PREFETCH") != nullptr;
    const bool ret = disc_ret != nullptr;
    const dim3 blk(GE, GC);
    if (estimator == 1) {
        if (ret) gae_dual_kernel<false, 1, true><<<nblocks, blk, 0, s>>>(a);
        else gae_dual_kernel<false, 1, false><<<nblocks, blk, 0, s>>>(a);
    } else if (estimator == 2) {
        if (ret) gae_dual_kernel<false, 2, true><<<nblocks, blk, 0, s>>>(a);
        else gae_dual_kernel<false, 2, false><<<nblocks, blk, 0, s>>>(a);
    } else if (estimator == 3) {
        if (ret) gae_dual_kernel<false, 3, true><<<nblocks, blk, 0, s>>>(a);
        else gae_dual_kernel<false, 3, false><<<nblocks, blk, 0, s>>>(a);
    } else if (prefetch && T > GT) {
        gae_dual_kernel<true, 0, true><<<nblocks, blk, 0, s>>>(a);
    } else if (ret) {
        gae_dual_kernel<false, 0, true><<<nblocks, blk, 0, s>>>(a);
    } else {
        // training path (no discounted_ret slab): two scans instead of three.  A segment-sequential variant
        // (32-env warps, 16-step segments composed through shared memory) was measured slower at every
        // horizon (12.7 vs 11.5 us at T=128, 1.4 vs 2.8 TB/s at T=2048): too few warps, too long chains.
        gae_dual_kernel<false, 0, false><<<nblocks, blk, 0, s>>>(a);
    }
    OSB_LAUNCH_CHECK();
    return OSB_OK;
}

int osb_gae_dual(const float* rew, const float* cost, const float* val_r, const float* val_c,
                 const uint8_t* flags, const float* boot_r, const float* boot_c, int T, int N,
                 double gamma, double lam, double lam_c, double penalty_coef, float* adv_r,
                 float* adv_c, float* tv_r, float* tv_c, float* disc_ret, double* workspace,
                 double* sums, void* stream) {
    return osb_adv_estimate(rew, cost, val_r, val_c, flags, boot_r, boot_c, T, N, gamma, lam, lam_c,
                            penalty_coef, 0, adv_r, adv_c, tv_r, tv_c, disc_ret, workspace, sums, stream);
}

int osb_adv_moments(const double* sums, int standardize_r, int standardize_c, float* moments,
                    void* stream) {
    OSB_CHECK_ARG(sums && moments, "null pointer");
    adv_moments_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(sums, standardize_r, standardize_c,
                                                          moments);
    OSB_LAUNCH_CHECK();
    return OSB_OK;
}

int osb_adv_standardize(const float* adv_r, const float* adv_c, const float* moments, long long n,
                        float* out_r, float* out_c, void* stream) {
    OSB_CHECK_ARG(adv_r && adv_c && moments && out_r && out_c && n >= 0, "bad argument");
    if (n == 0) return OSB_OK;
    int blocks = (int)((n + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    adv_standardize_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(adv_r, adv_c, moments,
                                                                    (size_t)n, out_r, out_c);
    OSB_LAUNCH_CHECK();
    return OSB_OK;
}

int osb_discount_cumsum(const void* x, int x_is_f64, int rows, int len, double discount,
                        double* out, void* stream) {
    OSB_CHECK_ARG(rows >= 0 && len >= 0 && out, "bad argument");
    if (rows == 0 || len == 0) return OSB_OK;
    OSB_CHECK_ARG(x != nullptr, "null input");
    const int wpb = 4;
    const int blocks = (rows + wpb - 1) / wpb;
    cudaStream_t s = (cudaStream_t)stream;
    if (x_is_f64)
        discount_cumsum_kernel<double><<<blocks, wpb * 32, 0, s>>>((const double*)x, rows, len,
                                                                   discount, out);
    else
        discount_cumsum_kernel<float><<<blocks, wpb * 32, 0, s>>>((const float*)x, rows, len,
                                                                  discount, out);
    OSB_LAUNCH_CHECK();
    return OSB_OK;
}

}  // extern "C"
