// Tensor-core (tcgen05, kind::tf32, TMEM accumulators) variant of the fused minibatch
// forward + loss + backward kernel -- the "fast" arithmetic mode of csrc/update.cu (which stays as
// the exact-fp32 parity path).  Same interface, same per-CTA partial-gradient outputs.
//
// All GEMM operands are fp32 tiles in shared memory in the K-major 128B-swizzled canonical layout
// (csrc/umma.cuh).  TF32 has no usable MN-major view of such a tile, so every activation that is
// needed with the sample index as the contraction dimension (weight gradients) is produced a second
// time in transposed form by a role-swapped MMA (D^T = W * X^T) instead of a transposed copy:
//
//   per 128-sample tile and network (O <= 64):
//     MMA1  Z1    [s][n] = X  * W1^T      -> H1    = tanh(.+b1)            (A operand of layer 2)
//     MMA2  Z1^T  [n][s] = W1 * X^T       -> H1^T                          (B operand of dW2)
//     MMA3  Z2    [s][n] = H1 * W2^T      -> H2
//     MMA4  OUT   [s][o] = H2 * W3^T      -> per-sample loss, dOUT
//     MMA5  dZ2   [s][k] = dOUT * W3      -> * (1 - H2^2)                  (B operand of dZ1^T)
//     MMA6  dZ2^T [k][s] = W3^T * dOUT^T  -> * (1 - H2^2)                  (A operand of dW2)
//     MMA7  dW2   [j][k] += dZ2^T * H1    (TMEM accumulator kept across tiles)
//     MMA8  dZ1^T [k][s] = W2^T * dZ2^T   -> * (1 - H1^2)                  (A operand of dW1)
//     MMA9  dW1   [j][o] += dZ1^T * X     (TMEM accumulator kept across tiles)
//   dW3, the bias gradients and d log_std stay on the CUDA cores (tiny).
#include "common.cuh"
#include "mlp.cuh"
#include "umma.cuh"

namespace osb {

using namespace umma;

constexpr int TT = 128;                 // samples per tile
constexpr uint32_t BUF = TT * 64 * 4;   // 32 KB activation buffer ([128][64] or [64][128] fp32)

enum TcLoss { TC_PPO_CLIP = 0, TC_RATIO = 1, TC_COST = 3 };

struct TcBatch {
    const float* obs; const float* act; const float* logp; const float* adv_r; const float* adv_c;
    const float* tv_r; const float* tv_c; const float* moments; const int* perm;
    long long total; unsigned perm_seed; long long mb_start; int mb_count;
};
struct TcArgs {
    TcBatch b;
    int kind; float clip, entropy_coef;
    const float* lagrange;
    const float* theta;
    float* gpart;
    float* stats_part;
    const int* stop_flag;
    int O, A, P, net_mask;
};

__device__ __forceinline__ unsigned long long tc_feistel(unsigned long long k, unsigned long long n, unsigned seed) {
    int bits = 2;
    while ((1ull << bits) < n) bits += 2;
    const int half = bits >> 1;
    const unsigned mask = (1u << half) - 1u;
    unsigned long long x = k;
    do {
        unsigned l = (unsigned)(x >> half) & mask, r = (unsigned)x & mask;
#pragma unroll
        for (int round = 0; round < 4; ++round) {
            const unsigned f = mix32(r ^ (seed + 0x9E3779B9u * (unsigned)(round + 1))) & mask;
            const unsigned nl = r;
            r = l ^ f;
            l = nl;
        }
        x = ((unsigned long long)l << half) | r;
    } while (x >= n);
    return x;
}

// TMEM column map
constexpr uint32_t C_Z = 0, C_ZT = 64, C_OUT = 192, C_DW2 = 224, C_DW1 = 288, TMEM_COLS = 512;

__global__ void __launch_bounds__(NTHREADS, 1) minibatch_grad_tc_kernel(TcArgs p) {
    if (p.stop_flag && *p.stop_flag) return;
    const int net = blockIdx.y;
    if (!((p.net_mask >> net) & 1)) return;

    extern __shared__ __align__(16) uint8_t smem_raw[];
    const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;   // tiles need 1024 B alignment
    const uint32_t B0 = smem_u32(smem_raw) + pad;   // X -> H2 -> dZ1^T
    const uint32_t B1 = B0 + BUF;                   // X^T
    const uint32_t B2 = B1 + BUF;                   // H1 -> dZ2
    const uint32_t B3 = B2 + BUF;                   // H1^T
    const uint32_t B4 = B3 + BUF;                   // dOUT (first 16 KB) -> dZ2^T
    const uint32_t sW1 = B4 + BUF;                  // [64][64]
    const uint32_t sW2 = sW1 + 16384;               // [64][64]
    const uint32_t sW2T = sW2 + 16384;              // [64][64]
    const uint32_t sW3 = sW2T + 16384;              // [16][64]   (2 atoms x 16 rows)
    const uint32_t sW3T = sW3 + 4096;               // [64][32]   (cols >= out zero)
    float* sB1 = reinterpret_cast<float*>(smem_raw + pad + 5 * BUF + 3 * 16384 + 4096 + 8192);
    float* sB2 = sB1 + 64;
    float* sB3 = sB2 + 64;            // [16]
    float* sLs = sB3 + 16;            // logstd[16], sigma[16], dlogstd acc[16]
    float* sStat = sLs + 48;          // [8]
    float* sRed = sStat + 8;          // [16 + 4 * 16]
    long long* sRow = reinterpret_cast<long long*>(sRed + 80);   // [128]
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q = warp & 3, h = warp >> 2;
    const int O = p.O, A = p.A;
    const NetLayout L = net_layout(net, O, A);
    const int noff = net_offset(net, O, A);
    const float* theta = p.theta + noff;
    float* gout = p.gpart + (size_t)blockIdx.x * p.P + noff;
    const int ntiles = (p.b.mb_count + TT - 1) / TT;
    const float inv_b = 1.0f / (float)p.b.mb_count;

    // ---- weights -> swizzled K-major tiles (loads batched so they are all in flight together) ----
    {
        float w1v[16], w2v[16], w3v[4];
        const int k = tid & 63;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int n = (tid >> 6) + 4 * j;
            w1v[j] = (k < O) ? __ldg(theta + L.off_w1 + n * O + k) : 0.f;
            w2v[j] = __ldg(theta + L.off_w2 + n * 64 + k);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int o = (tid >> 6) + 4 * j;
            w3v[j] = (o < L.out) ? __ldg(theta + L.off_w3 + o * 64 + k) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int n = (tid >> 6) + 4 * j;
            sts(tile_addr(sW1, n, k, 64), tf32r(w1v[j]));
            const float w2 = tf32r(w2v[j]);
            sts(tile_addr(sW2, n, k, 64), w2);
            sts(tile_addr(sW2T, k, n, 64), w2);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int o = (tid >> 6) + 4 * j;
            const float w = tf32r(w3v[j]);
            sts(tile_addr(sW3, o, k, 16), w);
            sts(tile_addr(sW3T, k, o, 64), w);
        }
    }
    for (int i = tid; i < 64 * 16; i += NTHREADS) sts(tile_addr(sW3T, i >> 4, 16 + (i & 15), 64), 0.f);
    if (tid < 64) { sB1[tid] = __ldg(theta + L.off_b1 + tid); sB2[tid] = __ldg(theta + L.off_b2 + tid); }
    if (tid < 16) {
        sB3[tid] = (tid < L.out) ? __ldg(theta + L.off_b3 + tid) : 0.f;
        const float ls = (net == 0 && tid < A) ? __ldg(theta + L.off_logstd + tid) : 0.f;
        sLs[tid] = ls; sLs[16 + tid] = expf(ls); sLs[32 + tid] = 0.f;
    }
    if (tid < 8) sStat[tid] = 0.f;
    if (tid == 0) { mbar_init(&bar, 1); mbar_init_fence(); }
    if (warp == 0) tmem_alloc(&tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    uint32_t phase = 0;

    const float lam = (p.lagrange != nullptr) ? __ldg(p.lagrange) : 0.f;
    const float m_r = __ldg(p.b.moments + 0), s_r = __ldg(p.b.moments + 1), m_c = __ldg(p.b.moments + 2);
    float aw3[4] = {0.f, 0.f, 0.f, 0.f};
    float ab1 = 0.f, ab2 = 0.f, ab3 = 0.f;
    bool first_tile = true;

    const uint32_t aB0 = B0, aB1 = B1, aB2 = B2, aB3 = B3, aB4 = B4;
    const uint32_t aW1 = sW1, aW2 = sW2, aW2T = sW2T, aW3 = sW3, aW3T = sW3T;

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        // ---- P0: gather X and X^T --------------------------------------------------------------
        if (tid < TT) {
            const int local = tile * TT + tid;
            long long row = -1;
            if (local < p.b.mb_count) {
                const long long k = p.b.mb_start + local;
                row = p.b.perm ? (long long)p.b.perm[k] : (long long)tc_feistel((unsigned long long)k, (unsigned long long)p.b.total, p.b.perm_seed);
            }
            sRow[tid] = row;
        }
        __syncthreads();
        {
            const int k = tid & 63;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                float xv[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int m = (tid >> 6) + 4 * (16 * half + j);
                    const long long row = sRow[m];
                    xv[j] = (row >= 0 && k < O) ? __ldg(p.b.obs + row * O + k) : 0.f;
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int m = (tid >> 6) + 4 * (16 * half + j);
                    const float v = tf32r(xv[j]);
                    sts(tile_addr(B0, m, k, TT), v);
                    sts(tile_addr(B1, k, m, 64), v);
                }
            }
        }
        fence_async_smem();
        __syncthreads();
        // ---- P1: Z1 and Z1^T ---------------------------------------------------------------------
        if (tid == 0) {
            tc_fence_after();
            tc_gemm(tmem + C_Z, aB0, TT, aW1, 64, 128, 64, 64, false);
            tc_gemm(tmem + C_ZT, aW1, 64, aB0, TT, 64, 128, 64, false);
            mma_commit(&bar);
        }
        mbar_wait(&bar, phase); phase ^= 1;
        tc_fence_after();
        {
            float v[32];
            tmem_ld32(tmem + lane_base + C_Z + 32 * h, v);        // row s = 32q+lane, cols 32h..
            const int s = 32 * q + lane;
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = tanh_fast(v[i] + sB1[32 * h + i]);
            store_row32(B2, s, 32 * h, TT, v);
            // transposed: row n = 16q + lane (lane < 16), cols s in [64h, 64h+64)
            const int n = 16 * q + lane;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                tmem_ld32(tmem + lane_base + C_ZT + 64 * h + 32 * half, v);
                if (lane < 16) {
                    const float bb = sB1[n];
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = tanh_fast(v[i] + bb);
                    store_row32(B3, n, 64 * h + 32 * half, 64, v);
                }
            }
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        // ---- P2: Z2 -> H2 (into B0; X is dead) ---------------------------------------------------
        if (tid == 0) {
            tc_fence_after();
            tc_gemm(tmem + C_Z, aB2, TT, aW2, 64, 128, 64, 64, false);
            mma_commit(&bar);
        }
        mbar_wait(&bar, phase); phase ^= 1;
        tc_fence_after();
        {
            float v[32];
            tmem_ld32(tmem + lane_base + C_Z + 32 * h, v);
            const int s = 32 * q + lane;
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = tanh_fast(v[i] + sB2[32 * h + i]);
            store_row32(B0, s, 32 * h, TT, v);
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        // ---- P3: OUT -> loss -> dOUT (B4, cols 0..31) --------------------------------------------
        if (tid == 0) {
            tc_fence_after();
            tc_gemm(tmem + C_OUT, aB0, TT, aW3, 16, 128, 16, 64, false);
            mma_commit(&bar);
        }
        // per-sample scalars: issue the global loads before blocking on the MMA
        float pf_act[16], pf_logp = 0.f, pf_advr = 0.f, pf_advc = 0.f, pf_tv = 0.f;
        {
            const long long prow = (h == 0) ? sRow[32 * q + lane] : -1;
#pragma unroll
            for (int a = 0; a < 16; ++a) pf_act[a] = 0.f;
            if (prow >= 0) {
                if (net == 0) {
#pragma unroll
                    for (int a = 0; a < 16; ++a)
                        if (a < A) pf_act[a] = __ldg(p.b.act + prow * A + a);
                    pf_logp = __ldg(p.b.logp + prow);
                    pf_advr = __ldg(p.b.adv_r + prow);
                    pf_advc = __ldg(p.b.adv_c + prow);
                } else {
                    pf_tv = __ldg((net == 1 ? p.b.tv_r : p.b.tv_c) + prow);
                }
            }
        }
        mbar_wait(&bar, phase); phase ^= 1;
        tc_fence_after();
        {
            float st[4] = {0.f, 0.f, 0.f, 0.f};
            float dls[16];
#pragma unroll
            for (int a = 0; a < 16; ++a) dls[a] = 0.f;
            if (h == 0) {
                float o16[16];
                tmem_ld16(tmem + lane_base + C_OUT, o16);
                const int s = 32 * q + lane;
                const long long row = sRow[s];
                float d32[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) d32[i] = 0.f;
                if (row >= 0) {
                    if (net != 0) {
                        const float v = o16[0] + sB3[0];
                        const float d = v - pf_tv;
                        st[0] = d * d; st[3] = 1.f;
                        d32[0] = 2.f * d * inv_b;
                    } else {
                        float logp_new = 0.f, diff[16];
#pragma unroll
                        for (int a = 0; a < 16; ++a) {
                            diff[a] = 0.f;
                            if (a < A) {
                                const float mu = o16[a] + sB3[a];
                                const float sd = sLs[16 + a];
                                const float d = pf_act[a] - mu;
                                diff[a] = d;
                                logp_new += -(d * d) / (2.f * sd * sd) - sLs[a] - 0.9189385332046727f;
                            }
                        }
                        const float ratio = expf(logp_new - pf_logp);
                        const float adv_r = (pf_advr - m_r) / s_r;
                        const float adv_c = pf_advc - m_c;
                        const float adv = (adv_r - lam * adv_c) / (1.f + lam);
                        float dlogp, loss;
                        if (p.kind == TC_PPO_CLIP) {
                            const float rc = fminf(fmaxf(ratio, 1.f - p.clip), 1.f + p.clip);
                            const float s1 = ratio * adv, s2 = rc * adv;
                            loss = -fminf(s1, s2);
                            dlogp = (s1 <= s2) ? -adv * ratio * inv_b : 0.f;
                        } else if (p.kind == TC_RATIO) {
                            loss = -ratio * adv; dlogp = -adv * ratio * inv_b;
                        } else {
                            loss = ratio * adv_c; dlogp = adv_c * ratio * inv_b;
                        }
                        st[0] = loss; st[1] = ratio; st[3] = 1.f;
#pragma unroll
                        for (int a = 0; a < 16; ++a)
                            if (a < A) {
                                const float sd = sLs[16 + a];
                                const float iv = 1.f / (sd * sd);
                                d32[a] = dlogp * diff[a] * iv;
                                dls[a] = dlogp * (diff[a] * diff[a] * iv - 1.f);
                            }
                    }
                }
                store_row32(B4, s, 0, TT, d32);
            }
            // deterministic reductions over the 128 sample threads (warps 0..3)
#pragma unroll
            for (int i = 0; i < 4; ++i) st[i] = warp_sum(st[i]);
#pragma unroll
            for (int a = 0; a < 16; ++a) dls[a] = warp_sum(dls[a]);
            if (h == 0 && lane == 0) {
#pragma unroll
                for (int i = 0; i < 4; ++i) sRed[q * 4 + i] = st[i];
#pragma unroll
                for (int a = 0; a < 16; ++a) sRed[16 + q * 16 + a] = dls[a];
            }
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (tid < 4) sStat[tid] += sRed[tid] + sRed[4 + tid] + sRed[8 + tid] + sRed[12 + tid];
        if (net == 0 && tid >= 32 && tid < 48) {
            const int a = tid - 32;
            sLs[32 + a] += sRed[16 + a] + sRed[32 + a] + sRed[48 + a] + sRed[64 + a];
        }
        // ---- P4: dZ2 and dZ2^T -------------------------------------------------------------------
        if (tid == 0) {
            tc_fence_after();
            tc_gemm(tmem + C_Z, aB4, TT, aW3T, 64, 128, 64, 16, false);
            tc_gemm(tmem + C_ZT, aW3T, 64, aB4, TT, 64, 128, 16, false);
            mma_commit(&bar);
        }
        {   // CUDA cores meanwhile: dW3[o][k] += sum_s dOUT[s][o] H2[s][k]; db3
            const int k = tid & 63, og = tid >> 6;
            const uint32_t hb = B0 + (uint32_t)((k >> 5) * TT * 128 + ((k & 3) << 2));
            const uint32_t hc = (uint32_t)((k >> 2) & 7);
            for (int s0 = 0; s0 < TT; s0 += 8) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint32_t rowoff = (uint32_t)((s0 + j) * 128);
                    const float hv = lds(hb + rowoff + ((hc ^ (uint32_t)j) << 4));
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int o = og + 4 * i;
                        if (o < L.out)
                            aw3[i] = fmaf(lds(B4 + rowoff + ((((uint32_t)o >> 2) ^ (uint32_t)j) << 4) + (((uint32_t)o & 3) << 2)), hv, aw3[i]);
                    }
                }
            }
            if (tid < L.out) {
                float c = 0.f;
                for (int s0 = 0; s0 < TT; s0 += 8)
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        c += lds(B4 + (uint32_t)((s0 + j) * 128) + ((((uint32_t)tid >> 2) ^ (uint32_t)j) << 4) + (((uint32_t)tid & 3) << 2));
                ab3 += c;
            }
        }
        __syncthreads();           // all CUDA-core reads of dOUT (B4) done before the epilogue overwrites it
        mbar_wait(&bar, phase); phase ^= 1;
        tc_fence_after();
        {
            float v[32], hh[32];
            tmem_ld32(tmem + lane_base + C_Z + 32 * h, v);
            const int s = 32 * q + lane;
            load_row32(B0, s, 32 * h, TT, hh);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] *= (1.f - hh[i] * hh[i]);
            store_row32(B2, s, 32 * h, TT, v);                   // dZ2 [s][k]
            const int k = 16 * q + lane;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                tmem_ld32(tmem + lane_base + C_ZT + 64 * h + 32 * half, v);
                if (lane < 16) {
                    const int s0 = 64 * h + 32 * half;
                    const uint32_t hb = B0 + (uint32_t)((k >> 5) * TT * 128 + ((k & 3) << 2) + s0 * 128);
                    const uint32_t hc = (uint32_t)((k >> 2) & 7);
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const float hv = lds(hb + (uint32_t)(i * 128) + ((hc ^ (uint32_t)(i & 7)) << 4));
                        v[i] *= (1.f - hv * hv);
                    }
                    store_row32(B4, k, s0, 64, v);               // dZ2^T [k][s]
                }
            }
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        // ---- P5: dW2 += dZ2^T H1 ; dZ1^T = W2^T dZ2^T --------------------------------------------
        if (tid == 0) {
            tc_fence_after();
            tc_gemm(tmem + C_DW2, aB4, 64, aB3, 64, 64, 64, 128, !first_tile);
            tc_gemm(tmem + C_ZT, aW2T, 64, aB2, TT, 64, 128, 64, false);
            mma_commit(&bar);
        }
        if (tid < 64) {   // db2[j] = sum_s dZ2^T[j][s]
            float c = 0.f, v[32];
#pragma unroll
            for (int a4 = 0; a4 < 4; ++a4) {
                load_row32(B4, tid, 32 * a4, 64, v);
#pragma unroll
                for (int i = 0; i < 32; ++i) c += v[i];
            }
            ab2 += c;
        }
        mbar_wait(&bar, phase); phase ^= 1;
        tc_fence_after();
        {
            float v[32], hh[32];
            const int k = 16 * q + lane;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                tmem_ld32(tmem + lane_base + C_ZT + 64 * h + 32 * half, v);
                if (lane < 16) {
                    const int s0 = 64 * h + 32 * half;
                    load_row32(B3, k, s0, 64, hh);
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] *= (1.f - hh[i] * hh[i]);
                    store_row32(B0, k, s0, 64, v);               // dZ1^T [k][s] (H2 is dead)
                }
            }
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        // ---- P6: dW1 += dZ1^T X --------------------------------------------------------------------
        if (tid == 0) {
            tc_fence_after();
            tc_gemm(tmem + C_DW1, aB0, 64, aB1, 64, 64, 64, 128, !first_tile);
            mma_commit(&bar);
        }
        if (tid < 64) {   // db1[j] = sum_s dZ1^T[j][s]
            float c = 0.f, v[32];
#pragma unroll
            for (int a4 = 0; a4 < 4; ++a4) {
                load_row32(B0, tid, 32 * a4, 64, v);
#pragma unroll
                for (int i = 0; i < 32; ++i) c += v[i];
            }
            ab1 += c;
        }
        mbar_wait(&bar, phase); phase ^= 1;      // B0 / B1 are rewritten by the next tile's gather
        tc_fence_after();
        first_tile = false;
        __syncthreads();
    }

    // ---- write this CTA's partial gradient segment ----------------------------------------------
    {
        float v[32];
        const int j = 16 * q + lane;
        tmem_ld32(tmem + lane_base + C_DW2 + 32 * h, v);
        if (lane < 16)
#pragma unroll
            for (int i = 0; i < 32; ++i) gout[L.off_w2 + j * 64 + 32 * h + i] = v[i];
        tmem_ld32(tmem + lane_base + C_DW1 + 32 * h, v);
        if (lane < 16)
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const int o = 32 * h + i;
                if (o < O) gout[L.off_w1 + j * O + o] = v[i];
            }
        const int k = tid & 63, og = tid >> 6;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int o = og + 4 * i;
            if (o < L.out) gout[L.off_w3 + o * 64 + k] = aw3[i];
        }
        if (tid < 64) { gout[L.off_b1 + tid] = ab1; gout[L.off_b2 + tid] = ab2; }
        if (tid < L.out) gout[L.off_b3 + tid] = ab3;
        __syncthreads();
        if (net == 0 && tid < A) {
            float g = sLs[32 + tid];
            if (blockIdx.x == 0 && p.kind == TC_PPO_CLIP) g -= p.entropy_coef / (float)A;
            gout[L.off_logstd + tid] = g;
        }
        if (tid < 8) p.stats_part[((size_t)blockIdx.x * 3 + net) * 8 + tid] = sStat[tid];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, TMEM_COLS);
}

}  // namespace osb

using namespace osb;

static size_t tc_smem_bytes() {
    return 1024 + 5 * (size_t)BUF + 3 * 16384 + 4096 + 8192 + (64 + 64 + 16 + 48 + 8 + 80) * 4 + 128 * 8 + 64;
}

extern "C" {

int osb_update_grid_blocks(int mb_count);

// Tensor-core (TF32 tcgen05) variant of osb_minibatch_grad; O <= 64, loss kinds 0 / 1 / 3.
int osb_minibatch_grad_tc(const float* theta, int O, int A, const float* obs, const float* act,
                          const float* logp, const float* adv_r, const float* adv_c,
                          const float* tv_r, const float* tv_c, const float* moments,
                          const int* perm, long long total, unsigned perm_seed, long long mb_start,
                          int mb_count, int loss_kind, float clip, float entropy_coef,
                          const float* lagrange, int net_mask, float* gpart, float* stats_part,
                          const int* stop_flag, void* stream) {
    OSB_CHECK_ARG(theta && obs && act && logp && adv_r && adv_c && tv_r && tv_c && moments, "null input");
    OSB_CHECK_ARG(O > 0 && O <= 64 && A > 0 && A <= 16 && mb_count > 0 && total > 0, "tensor-core path needs O <= 64, A <= 16");
    OSB_CHECK_ARG(mb_start >= 0 && mb_start + mb_count <= total, "minibatch window out of range");
    OSB_CHECK_ARG(loss_kind == 0 || loss_kind == 1 || loss_kind == 3, "tensor-core path: loss kind 0, 1 or 3");
    TcArgs p;
    p.b = TcBatch{obs, act, logp, adv_r, adv_c, tv_r, tv_c, moments, perm, total, perm_seed, mb_start, mb_count};
    p.kind = loss_kind; p.clip = clip; p.entropy_coef = entropy_coef; p.lagrange = lagrange;
    p.theta = theta; p.gpart = gpart; p.stats_part = stats_part; p.stop_flag = stop_flag;
    p.O = O; p.A = A; p.P = actor_layout(O, A).size + 2 * critic_layout(O, A).size; p.net_mask = net_mask;
    const size_t smem = tc_smem_bytes();
    static bool attr = false;
    if (!attr) {
        OSB_CUDA(cudaFuncSetAttribute(minibatch_grad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = true;
    }
    dim3 grid(osb_update_grid_blocks(mb_count), 3);
    minibatch_grad_tc_kernel<<<grid, NTHREADS, smem, (cudaStream_t)stream>>>(p);
    OSB_LAUNCH_CHECK();
    return OSB_OK;
}

}  // extern "C"
