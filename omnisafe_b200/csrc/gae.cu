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
// Parallelisation: block = 32 envs (x) x 32 time-chunks (y) x 4 steps per thread.  Every thread
// folds its 4 steps into an affine map  A_in -> b + a*A_in  (fp64), the 32 chunk maps of one env are
// combined with a warp-shuffle suffix scan (after a shared-memory transpose so that lanes run along
// time), and a second local pass replays the reference's sequential arithmetic
// (separately rounded fp64 mul/add) from the exact carry-in.  Tiles of 128 steps are walked from
// the end of the horizon to its start with a per-env carry.
#include "common.cuh"

namespace osb {

constexpr int GE = 32;         // envs per block
constexpr int GC = 32;         // chunks per tile
constexpr int GL = 4;          // steps per chunk
constexpr int GT = GC * GL;    // steps per tile
constexpr int GPAD = GE + 1;

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
    int T, N;
    float gamma_f;    // (float)gamma : fp32 delta arithmetic (onpolicy_buffer.py:L301)
    float pen;        // penalty_coefficient (onpolicy_buffer.py:L185)
    double g;         // gamma            (discounted return)
    double gl_r;      // gamma * lam      (python double product)
    double gl_c;      // gamma * lam_c
};

__global__ void __launch_bounds__(GE* GC, 1) gae_dual_kernel(GaeArgs p) {
    extern __shared__ double sm[];
    double* sa = sm;                         // [3][GC][GPAD]
    double* sb = sm + 3 * GC * GPAD;         // [3][GC][GPAD]
    double* carry = sm + 6 * GC * GPAD;      // [3][GE]
    double* red = carry + 3 * GE;            // [3][32]

    const int x = threadIdx.x, y = threadIdx.y;
    const int env = blockIdx.x * GE + x;
    const bool env_ok = env < p.N;
    const int N = p.N, T = p.T;
    const int ntiles = (T + GT - 1) / GT;

    if (y < 3) carry[y * GE + x] = 0.0;
    __syncthreads();

    double st_r = 0.0, st_r2 = 0.0, st_c = 0.0;

    for (int k = 0; k < ntiles; ++k) {
        const int t0 = T - (k + 1) * GT + y * GL;  // first step of my chunk (may be < 0)
        float r[GL], c[GL], vr[GL + 1], vc[GL + 1];
        unsigned f[GL];
#pragma unroll
        for (int i = 0; i < GL; ++i) {
            const int t = t0 + i;
            const bool ok = env_ok && t >= 0;
            const size_t idx = (size_t)(ok ? t : 0) * N + (env_ok ? env : 0);
            r[i] = ok ? __ldg(p.rew + idx) : 0.f;
            c[i] = ok ? __ldg(p.cost + idx) : 0.f;
            vr[i] = ok ? __ldg(p.val_r + idx) : 0.f;
            vc[i] = ok ? __ldg(p.val_c + idx) : 0.f;
            f[i] = ok ? (unsigned)__ldg(p.flags + idx) : 0u;
        }
        {
            const int t = t0 + GL;
            const bool ok = env_ok && t >= 0 && t < T;
            const size_t idx = (size_t)(ok ? t : 0) * N + (env_ok ? env : 0);
            vr[GL] = ok ? __ldg(p.val_r + idx) : 0.f;
            vc[GL] = ok ? __ldg(p.val_c + idx) : 0.f;
        }
        // fp32 deltas with the reference's three separately rounded ops; bootstrap at path ends.
        float dr[GL], dc[GL], bootr[GL];
        bool end[GL], valid[GL];
#pragma unroll
        for (int i = 0; i < GL; ++i) {
            const int t = t0 + i;
            valid[i] = env_ok && t >= 0;
            end[i] = valid[i] && (f[i] != 0u || t == T - 1);
            float nr = vr[i + 1], nc = vc[i + 1];
            if (end[i]) {
                const bool term = (f[i] & OSB_FLAG_TERMINATED) != 0u;
                const size_t idx = (size_t)t * N + env;
                nr = term ? 0.f : __ldg(p.boot_r + idx);
                nc = term ? 0.f : __ldg(p.boot_c + idx);
            }
            bootr[i] = nr;
            const float rp = __fadd_rn(r[i], -__fmul_rn(p.pen, c[i]));
            dr[i] = __fadd_rn(__fadd_rn(rp, __fmul_rn(p.gamma_f, nr)), -vr[i]);
            dc[i] = __fadd_rn(__fadd_rn(c[i], __fmul_rn(p.gamma_f, nc)), -vc[i]);
        }
        // pass 1: fold the chunk into affine maps (a, b) per quantity.
        double ar = 1.0, br = 0.0, ac = 1.0, bc = 0.0, ag = 1.0, bg = 0.0;
#pragma unroll
        for (int i = GL - 1; i >= 0; --i) {
            if (valid[i]) {
                if (end[i]) {
                    ar = 0.0; br = (double)dr[i];
                    ac = 0.0; bc = (double)dc[i];
                    ag = 0.0; bg = (double)r[i] + p.g * (double)bootr[i];
                } else {
                    br = (double)dr[i] + p.gl_r * br; ar = p.gl_r * ar;
                    bc = (double)dc[i] + p.gl_c * bc; ac = p.gl_c * ac;
                    bg = (double)r[i] + p.g * bg;     ag = p.g * ag;
                }
            }
        }
        sa[(0 * GC + y) * GPAD + x] = ar; sb[(0 * GC + y) * GPAD + x] = br;
        sa[(1 * GC + y) * GPAD + x] = ac; sb[(1 * GC + y) * GPAD + x] = bc;
        sa[(2 * GC + y) * GPAD + x] = ag; sb[(2 * GC + y) * GPAD + x] = bg;
        __syncthreads();
        // transposed role: this warp owns env y of the block, lane x = chunk index.
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            double a = sa[(q * GC + x) * GPAD + y];
            double b = sb[(q * GC + x) * GPAD + y];
#pragma unroll
            for (int off = 1; off < GC; off <<= 1) {
                const double a2 = __shfl_down_sync(0xffffffffu, a, off);
                const double b2 = __shfl_down_sync(0xffffffffu, b, off);
                if (x + off < GC) { b = b + a * b2; a = a * a2; }
            }
            const double cin = carry[q * GE + y];
            const double full = b + a * cin;              // value at the first step of chunk x
            double ain = __shfl_down_sync(0xffffffffu, full, 1);
            if (x == GC - 1) ain = cin;                   // last chunk takes the tile carry
            __syncwarp();
            sb[(q * GC + x) * GPAD + y] = ain;
            if (x == 0) carry[q * GE + y] = full;
        }
        __syncthreads();
        // pass 2: replay sequentially from the exact carry-in with the reference's roundings.
        double Ar = sb[(0 * GC + y) * GPAD + x];
        double Ac = sb[(1 * GC + y) * GPAD + x];
        double Ag = sb[(2 * GC + y) * GPAD + x];
#pragma unroll
        for (int i = GL - 1; i >= 0; --i) {
            if (valid[i]) {
                if (end[i]) {
                    Ar = (double)dr[i];
                    Ac = (double)dc[i];
                    Ag = __dadd_rn((double)r[i], __dmul_rn(p.g, (double)bootr[i]));
                } else {
                    Ar = __dadd_rn((double)dr[i], __dmul_rn(p.gl_r, Ar));
                    Ac = __dadd_rn((double)dc[i], __dmul_rn(p.gl_c, Ac));
                    Ag = __dadd_rn((double)r[i], __dmul_rn(p.g, Ag));
                }
                const size_t idx = (size_t)(t0 + i) * N + env;
                const float o_ar = (float)Ar, o_ac = (float)Ac;
                p.adv_r[idx] = o_ar;
                p.adv_c[idx] = o_ac;
                p.tv_r[idx] = (float)(Ar + (double)vr[i]);
                p.tv_c[idx] = (float)(Ac + (double)vc[i]);
                if (p.disc_ret) p.disc_ret[idx] = (float)Ag;
                st_r += (double)o_ar;
                st_r2 += (double)o_ar * (double)o_ar;
                st_c += (double)o_ac;
            }
        }
        __syncthreads();
    }
    // epilogue: block partial sums for the advantage statistics (fixed order -> deterministic).
    st_r = warp_sum(st_r); st_r2 = warp_sum(st_r2); st_c = warp_sum(st_c);
    if (x == 0) { red[0 * 32 + y] = st_r; red[1 * 32 + y] = st_r2; red[2 * 32 + y] = st_c; }
    __syncthreads();
    if (y == 0) {
        double a = warp_sum(red[0 * 32 + x]);
        double b = warp_sum(red[1 * 32 + x]);
        double c = warp_sum(red[2 * 32 + x]);
        if (x == 0) {
            double* o = p.partials + (size_t)blockIdx.x * 4;
            o[0] = a; o[1] = b; o[2] = c;
            const int nenv = min(GE, N - blockIdx.x * GE);
            o[3] = (double)nenv * (double)T;
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

int osb_gae_workspace_doubles(int n_envs) { return ((n_envs + GE - 1) / GE) * 4 + 4; }

int osb_gae_dual(const float* rew, const float* cost, const float* val_r, const float* val_c,
                 const uint8_t* flags, const float* boot_r, const float* boot_c, int T, int N,
                 double gamma, double lam, double lam_c, double penalty_coef, float* adv_r,
                 float* adv_c, float* tv_r, float* tv_c, float* disc_ret, double* workspace,
                 double* sums, void* stream) {
    OSB_CHECK_ARG(T > 0 && N > 0, "T, N must be positive");
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
    const size_t smem = (size_t)(6 * GC * GPAD + 3 * GE + 3 * 32) * sizeof(double);
    static bool attr_set = false;
    if (!attr_set) {
        OSB_CUDA(cudaFuncSetAttribute(gae_dual_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
        attr_set = true;
    }
    cudaStream_t s = (cudaStream_t)stream;
    gae_dual_kernel<<<nblocks, dim3(GE, GC), smem, s>>>(a);
    OSB_LAUNCH_CHECK();
    gae_stats_reduce_kernel<<<1, 32, 0, s>>>(workspace, nblocks, sums);
    OSB_LAUNCH_CHECK();
    return OSB_OK;
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
