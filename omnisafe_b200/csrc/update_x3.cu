// Parity-grade tensor-core variant of the fused minibatch forward + loss + backward kernel
// (PolicyGradient._update, algorithms/on_policy/base/policy_gradient.py:L345-524; PPO._loss_pi base/ppo.py:L35-87;
// PPOLag._compute_adv_surrogate naive_lagrange/ppo_lag.py:L82-102): split-bf16 arithmetic (csrc/x3.cuh), i.e.
// every GEMM is six tcgen05 kind::f16 MMAs over the three bf16 pieces of its fp32 operands with fp32
// accumulation in TMEM -- fp32-level results on the tensor cores (the reference computes fp32 Linear layers,
// omnisafe/utils/model.py:L105-111).
//
// One CTA = one network x a strided set of 128-sample tiles (same grid / per-CTA partial-gradient contract as
// minibatch_grad_tc_kernel).  Warp-specialised: 16 epilogue warps + 1 MMA-issue warp, linked by mbarriers only:
//
//   per tile (activations stored ONCE as [sample][feature] bf16x3 tiles; the weight-gradient GEMMs read the
//   same tiles MN-major, so there are no transposed copies):
//     Z1   = X  W1^T            -> H1 = tanh(. + b1)
//     Z2   = H1 W2^T            -> H2 = tanh(. + b2)
//     OUT  = H2 W3^T            -> per-sample loss, dOUT
//     dZ2' = dOUT W3            -> dZ2 = dZ2' (1 - H2^2)   (stored over H2)     | dW3^T += H2^T dOUT
//     dZ1' = dZ2 W2             -> dZ1 = dZ1' (1 - H1^2)   (stored over H1)     | dW2 += dZ2^T H1, db2 += dZ2^T 1
//                                                                               | dW1 += dZ1^T X,  db1 += dZ1^T 1
//   Every epilogue writes its activation in two column halves, each announced by its own mbarrier, so the
//   next layer's MMAs start on k-steps 0-1 while the epilogue still produces k-steps 2-3; weight / bias
//   gradient MMAs run behind the dependent chain while the epilogue warps work.
#include "common.cuh"
#include "mlp.cuh"
#include "x3.cuh"

namespace osb {

using namespace x3;

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 512;\n" ::: "memory"); }      // epilogue warps only
__device__ __forceinline__ void loss_bar_sync() { asm volatile("bar.sync 2, 128;\n" ::: "memory"); }     // loss warps (h == 0)

enum X3Loss { X3_PPO_CLIP = 0, X3_RATIO = 1, X3_COST = 3 };

struct X3Batch {
    const float* obs; const float* act; const float* logp; const float* adv_r; const float* adv_c;
    const float* tv_r; const float* tv_c; const float* moments; const int* perm;
    long long total; unsigned perm_seed; long long mb_start; int mb_count;
    int identity_stride;     // > 0: row = (mb_start + local) * identity_stride (full-batch passes)
};
struct X3Args {
    X3Batch b;
    int kind; float clip, entropy_coef;
    const float* lagrange;
    const float* theta;
    float* gpart;            // [gridDim.x][P]
    float* stats_part;       // [gridDim.x][3][8]
    const int* stop_flag;
    int O, A, P, net_mask;
    // ---- fused optimiser (persistent mode: one launch = one pass over [mb_start, mb_start + mb_count) in
    //      minibatches of batch_size, each followed by reduce + clip + (all-reduce) + Adam inside the kernel) ----
    int batch_size;
    float* theta_rw; float* grad; float* adam_m; float* adam_v; int* adam_step;
    float critic_norm_coef, max_grad_norm, lr[3];
    float* sumsq_part;           // [3][2][gridDim.x]
    float* train_stats;          // [3][8]
    unsigned int* bar_ctr;       // [3] per-network grid barrier counters, zero at launch
    float* const* peer_buf;      // [world] receive buffers, each [2][world][P] (pushed over NVLink), or null
    unsigned int* const* peer_flag;   // [world] flag arrays, each [2 * world + 2 * world * 160]
    int world, rank;
    unsigned int step_base;      // exchange step id of this launch's first minibatch (identical on all ranks)
    int* error_flag;
};

constexpr int XT = 128;                      // samples per tile
constexpr int NEPI = 512;                    // 16 epilogue warps: lane quarter q = warp % 4, column group h = warp / 4
constexpr int NTX3 = NEPI + 32;              // + the MMA-issue warp
constexpr uint32_t ACT_SUB = XT * 128, ACT_X3 = 3 * ACT_SUB;        // [128][64] bf16 sub-tile, x3 tile
constexpr uint32_t D_SUB = XT * 32, D_X3 = 3 * D_SUB;               // [128][16] bf16 (SW32)
constexpr uint32_t W_SUB = 64 * 128, W_X3 = 3 * W_SUB;              // [64][64]
constexpr uint32_t W3_SUB = 16 * 128, W3_X3 = 3 * W3_SUB;           // [16][64]
constexpr uint32_t OFF_X = 0, OFF_H1 = OFF_X + ACT_X3, OFF_H2 = OFF_H1 + ACT_X3, OFF_D = OFF_H2 + ACT_X3,
                   OFF_W1 = OFF_D + D_X3, OFF_W2 = OFF_W1 + W_X3, OFF_W3 = OFF_W2 + W_X3, OFF_ONES = OFF_W3 + W3_X3,
                   OFF_MISC = OFF_ONES + 512;
// misc region (floats unless noted)
constexpr int MF_B1 = 0, MF_B2 = 64, MF_B3 = 128, MF_LS = 144 /* logstd[16] sigma[16] dlogstd acc[16] */, MF_STAT = 192,
              MF_RED = 200 /* [4*8 + 4*16 + 4*16] */, MF_B3ACC = 360, MF_PART = 376 /* [2][256] */, MF_SCAL = 888 /* [8] */, MF_END = 896;
constexpr uint32_t OFF_ROWS = OFF_MISC + MF_END * 4;                 // long long [2][128]
constexpr uint32_t OFF_BARS = OFF_ROWS + 2 * XT * 8;                 // uint64 [NBAR]
enum Bar { RDY_X0 = 0, RDY_X1, RDY_H1_0, RDY_H1_1, RDY_H2_0, RDY_H2_1, RDY_D, RDY_DZ2_0, RDY_DZ2_1, RDY_DZ1,
           DONE_C1, DONE_C2, DONE_C3, DONE_C4A, DONE_C4B, DONE_C5A, DONE_C5B, DONE_C6, NBAR };
constexpr uint32_t OFF_TMEMSLOT = OFF_BARS + NBAR * 8;
constexpr uint32_t X3_SMEM = OFF_TMEMSLOT + 16;
// TMEM columns
constexpr uint32_t T_ZA = 0, T_ZB = 64, T_OUT = 128, T_DW1 = 144, T_DW2 = 208, T_DW3 = 272, T_DB1 = 288, T_DB2 = 304, T_COLS = 512;

__device__ __forceinline__ unsigned long long x3_feistel(unsigned long long k, unsigned long long n, unsigned seed) {
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

// fp32 parameters of one network -> bf16x3 weight tiles + fp32 biases in shared memory (all NEPI epilogue threads)
__device__ __forceinline__ void stage_weights_x3(uint32_t sbase, float* misc, const float* __restrict__ theta,
                                                 const NetLayout& L, int net, int O, int A, int tid) {
    for (int i = tid; i < 64 * 32; i += NEPI) {          // W1 / W2: [64 n][64 k], two columns per thread
        const int n = i >> 5, k = (i & 31) << 1;
        const float a1 = (k < O) ? __ldcg(theta + L.off_w1 + n * O + k) : 0.f;
        const float b1 = (k + 1 < O) ? __ldcg(theta + L.off_w1 + n * O + k + 1) : 0.f;
        const float a2 = __ldcg(theta + L.off_w2 + n * 64 + k), b2 = __ldcg(theta + L.off_w2 + n * 64 + k + 1);
        uint32_t w0, w1, w2;
        const uint32_t off = off128(n, k);
        split2(a1, b1, w0, w1, w2);
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(sbase + OFF_W1 + off), "r"(w0) : "memory");
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(sbase + OFF_W1 + W_SUB + off), "r"(w1) : "memory");
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(sbase + OFF_W1 + 2 * W_SUB + off), "r"(w2) : "memory");
        split2(a2, b2, w0, w1, w2);
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(sbase + OFF_W2 + off), "r"(w0) : "memory");
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(sbase + OFF_W2 + W_SUB + off), "r"(w1) : "memory");
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(sbase + OFF_W2 + 2 * W_SUB + off), "r"(w2) : "memory");
    }
    {                                                      // W3: [16 o][64 k], rows >= out zero
        const int o = tid >> 5, k = (tid & 31) << 1;
        const float a = (o < L.out) ? __ldcg(theta + L.off_w3 + o * 64 + k) : 0.f;
        const float b = (o < L.out) ? __ldcg(theta + L.off_w3 + o * 64 + k + 1) : 0.f;
        uint32_t w0, w1, w2;
        split2(a, b, w0, w1, w2);
        const uint32_t off = off128(o, k);
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(sbase + OFF_W3 + off), "r"(w0) : "memory");
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(sbase + OFF_W3 + W3_SUB + off), "r"(w1) : "memory");
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(sbase + OFF_W3 + 2 * W3_SUB + off), "r"(w2) : "memory");
    }
    if (tid < 64) { misc[MF_B1 + tid] = __ldcg(theta + L.off_b1 + tid); misc[MF_B2 + tid] = __ldcg(theta + L.off_b2 + tid); }
    if (tid < 16) {
        misc[MF_B3 + tid] = (tid < L.out) ? __ldcg(theta + L.off_b3 + tid) : 0.f;
        const float ls = (net == 0 && tid < A) ? __ldcg(theta + L.off_logstd + tid) : 0.f;
        misc[MF_LS + tid] = ls; misc[MF_LS + 16 + tid] = expf(ls);
    }
}

__device__ __forceinline__ unsigned int ld_acquire_gpu_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned int ld_acquire_sys_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_u32(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// FUSED = false: one minibatch, per-CTA partial gradients out (the contract of minibatch_grad_tc_kernel).
// FUSED = true : persistent cooperative kernel -- the CTA loops over all minibatches of one update iteration
//   (policy_gradient.py:L369-381); after each minibatch the CTAs of a network meet at a software grid barrier,
//   reduce the partial gradients slice-wise in a fixed order, clip per network (clip_grad_norm_), exchange the
//   clipped slice with the peer ranks over NVLink (clip -> average -> step, policy_gradient.py:L437-443,
//   distributed.py:L193-198), apply torch-Adam and re-stage the new weights: no relaunch, no separate
//   optimiser kernel.
template <bool FUSED>
__global__ void __launch_bounds__(NTX3, 1) minibatch_grad_x3_kernel(X3Args p) {
    if (p.stop_flag && *p.stop_flag) return;
    const int net = (gridDim.y == 1) ? (__ffs(p.net_mask) - 1) : (int)blockIdx.y;
    if (!((p.net_mask >> net) & 1)) return;

    extern __shared__ __align__(16) uint8_t smem_raw[];
    const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
    const uint32_t sbase = smem_u32(smem_raw) + pad;
    uint8_t* gbase = smem_raw + pad;
    float* misc = reinterpret_cast<float*>(gbase + OFF_MISC);
    long long* sRowBuf = reinterpret_cast<long long*>(gbase + OFF_ROWS);
    const uint32_t bars = sbase + OFF_BARS;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gbase + OFF_TMEMSLOT);
    auto bar = [&](int i) { return bars + (uint32_t)i * 8u; };

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int O = p.O, A = p.A;
    const int G = (int)gridDim.x;
    const NetLayout L = net_layout(net, O, A);
    const int noff = net_offset(net, O, A);
    const float* theta = (FUSED ? p.theta_rw : p.theta) + noff;
    float* gout = p.gpart + (size_t)blockIdx.x * p.P + noff;
    const bool is_mma_warp = warp == NEPI / 32;
    const int batch = FUSED ? p.batch_size : p.b.mb_count;
    const int n_mb = (p.b.mb_count + batch - 1) / batch;

    // ---- one-time setup -------------------------------------------------------------------------------
    if (!is_mma_warp) {
        stage_weights_x3(sbase, misc, theta, L, net, O, A, tid);
        if (tid < 128) reinterpret_cast<uint32_t*>(gbase + OFF_ONES)[tid] = 0x3F803F80u;       // bf16 1.0 x 256
        if (tid < 16) { misc[MF_LS + 32 + tid] = 0.f; misc[MF_B3ACC + tid] = 0.f; }
        if (tid < 8) misc[MF_STAT + tid] = 0.f;
    } else {
        if (lane == 0) {
            for (int i = 0; i < NBAR; ++i) {
                const uint32_t cnt = (i >= DONE_C1) ? 1u : (i == RDY_D ? 4u : 16u);
                asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar(i)), "r"(cnt) : "memory");
            }
            mbar_init_fence();
        }
        __syncwarp();
        tmem_alloc(tmem_slot, T_COLS);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (is_mma_warp) {
        // ======================= MMA-issue warp: uniform control flow, one elected lane issues ==============
        const bool leader = elect_one_sync();
        const uint64_t dX = desc128(sbase + OFF_X), dH1 = desc128(sbase + OFF_H1), dH2 = desc128(sbase + OFF_H2);
        const uint64_t dD = desc32(sbase + OFF_D), dW1 = desc128(sbase + OFF_W1), dW2 = desc128(sbase + OFF_W2);
        const uint64_t dW3 = desc128(sbase + OFF_W3), dOnes = desc32(sbase + OFF_ONES);
        const uint32_t id_fwd = idesc_bf16(128, 64, 0, 0), id_out = idesc_bf16(128, 16, 0, 0), id_bwd = idesc_bf16(128, 64, 0, 1);
        const uint32_t id_dw = idesc_bf16(64, 64, 1, 1), id_dw16 = idesc_bf16(64, 16, 1, 1);
        int it = 0;
#pragma unroll 1
        for (int mb = 0; mb < n_mb; ++mb) {
            const int count = min(batch, p.b.mb_count - mb * batch);
            const int ntiles = (count + XT - 1) / XT;
#pragma unroll 1
            for (int tile = blockIdx.x; tile < ntiles; tile += G, ++it) {
                const uint32_t par = (uint32_t)(it & 1);
                const bool acc_dw = tile != (int)blockIdx.x;        // the first tile of a minibatch overwrites the accumulators
                // Z1 = X W1^T  (k-steps 0-1 after the first column half of X, 2-3 after the second)
#pragma unroll 1
                for (int ph = 0; ph < 2; ++ph) {
                    mbar_wait_a(bar(RDY_X0 + ph), par);
                    tc_fence_after();
                    gemm_x3_warp(leader, tmem + T_ZA, desc_add(dX, 64u * ph), ACT_SUB, 32u, desc_add(dW1, 64u * ph), W_SUB, 32u, id_fwd, 2, ph > 0);
                }
                if (leader) mma_commit_a(bar(DONE_C1));
                __syncwarp();
                // Z2 = H1 W2^T
#pragma unroll 1
                for (int ph = 0; ph < 2; ++ph) {
                    mbar_wait_a(bar(RDY_H1_0 + ph), par);
                    tc_fence_after();
                    gemm_x3_warp(leader, tmem + T_ZB, desc_add(dH1, 64u * ph), ACT_SUB, 32u, desc_add(dW2, 64u * ph), W_SUB, 32u, id_fwd, 2, ph > 0);
                }
                if (leader) mma_commit_a(bar(DONE_C2));
                __syncwarp();
                // OUT = H2 W3^T
#pragma unroll 1
                for (int ph = 0; ph < 2; ++ph) {
                    mbar_wait_a(bar(RDY_H2_0 + ph), par);
                    tc_fence_after();
                    gemm_x3_warp(leader, tmem + T_OUT, desc_add(dH2, 64u * ph), ACT_SUB, 32u, desc_add(dW3, 64u * ph), W3_SUB, 32u, id_out, 2, ph > 0);
                }
                if (leader) mma_commit_a(bar(DONE_C3));
                __syncwarp();
                // dZ2' = dOUT W3 ; dW3^T += H2^T dOUT
                mbar_wait_a(bar(RDY_D), par);
                tc_fence_after();
                gemm_x3_warp(leader, tmem + T_ZA, dD, D_SUB, 32u, dW3, W3_SUB, 2048u, id_bwd, 1, false);
                if (leader) mma_commit_a(bar(DONE_C4A));
                __syncwarp();
                gemm_x3_warp(leader, tmem + T_DW3, dH2, ACT_SUB, 2048u, dD, D_SUB, 512u, id_dw16, 8, acc_dw);
                if (leader) mma_commit_a(bar(DONE_C4B));
                __syncwarp();
                // dZ1' = dZ2 W2 ; dW2 += dZ2^T H1 ; db2 += dZ2^T 1
#pragma unroll 1
                for (int ph = 0; ph < 2; ++ph) {
                    mbar_wait_a(bar(RDY_DZ2_0 + ph), par);
                    tc_fence_after();
                    gemm_x3_warp(leader, tmem + T_ZB, desc_add(dH2, 64u * ph), ACT_SUB, 32u, desc_add(dW2, 4096u * ph), W_SUB, 2048u, id_bwd, 2, ph > 0);
                }
                if (leader) mma_commit_a(bar(DONE_C5A));
                __syncwarp();
                gemm_x3_warp(leader, tmem + T_DW2, dH2, ACT_SUB, 2048u, dH1, ACT_SUB, 2048u, id_dw, 8, acc_dw);
                gemm_x3_warp(leader, tmem + T_DB2, dH2, ACT_SUB, 2048u, dOnes, 0u, 0u, id_dw16, 8, acc_dw);
                if (leader) mma_commit_a(bar(DONE_C5B));
                __syncwarp();
                // dW1 += dZ1^T X ; db1 += dZ1^T 1
                mbar_wait_a(bar(RDY_DZ1), par);
                tc_fence_after();
                gemm_x3_warp(leader, tmem + T_DW1, dH1, ACT_SUB, 2048u, dX, ACT_SUB, 2048u, id_dw, 8, acc_dw);
                gemm_x3_warp(leader, tmem + T_DB1, dH1, ACT_SUB, 2048u, dOnes, 0u, 0u, id_dw16, 8, acc_dw);
                if (leader) mma_commit_a(bar(DONE_C6));
                __syncwarp();
            }
        }
    } else {
        // ======================= epilogue warps ===============================================================
        const int q = warp & 3, h = warp >> 2;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        const int s_row = 32 * q + lane;                 // sample row of this thread in the [s][.] accumulators
        const float lam = (p.lagrange != nullptr) ? __ldg(p.lagrange) : 0.f;
        float m_r = 0.f, s_r = 1.f, m_c = 0.f;
        if (p.b.moments) { m_r = __ldg(p.b.moments + 0); s_r = __ldg(p.b.moments + 1); m_c = __ldg(p.b.moments + 2); }
        float* sB1 = misc + MF_B1; float* sB2 = misc + MF_B2; float* sB3 = misc + MF_B3; float* sLs = misc + MF_LS;
        float* sStat = misc + MF_STAT; float* sRed = misc + MF_RED; float* sB3acc = misc + MF_B3ACC;
        float* sPart = misc + MF_PART; float* sScal = misc + MF_SCAL;

        // X gather: thread -> row xm = tid / 4, columns 32 ph + 8 (tid % 4) .. + 7 in column half ph
        const int xm = tid >> 2, xc = (tid & 3) << 3;
        const bool vec = (O & 3) == 0;
        float xpre[16];
        auto prefetch_x = [&](const long long* rows) {
            const long long row = rows[xm];
#pragma unroll
            for (int ph = 0; ph < 2; ++ph) {
                const int c0 = 32 * ph + xc;
                if (vec) {
#pragma unroll
                    for (int v4 = 0; v4 < 2; ++v4) {
                        const int c = c0 + 4 * v4;
                        const float4 v = (row >= 0 && c < O) ? __ldg(reinterpret_cast<const float4*>(p.b.obs + row * O + c))
                                                             : make_float4(0.f, 0.f, 0.f, 0.f);
                        xpre[8 * ph + 4 * v4] = v.x; xpre[8 * ph + 4 * v4 + 1] = v.y; xpre[8 * ph + 4 * v4 + 2] = v.z; xpre[8 * ph + 4 * v4 + 3] = v.w;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) xpre[8 * ph + i] = (row >= 0 && c0 + i < O) ? __ldg(p.b.obs + row * O + c0 + i) : 0.f;
                }
            }
        };
        auto announce = [&](int b) {        // this warp's stores of one column half are visible to the tensor core
            fence_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(b));
        };
        unsigned int nbar = 0;              // software grid barriers passed so far (FUSED)
        auto net_barrier = [&]() {          // all CTAs of this network: writes before it are visible after it (via L2)
            epi_bar_sync();
            if (tid == 0) {
                __threadfence();
                atomicAdd(p.bar_ctr + net, 1u);
                const unsigned int target = (unsigned int)G * (nbar + 1u);
                while (ld_acquire_gpu_u32(p.bar_ctr + net) < target) {}
                __threadfence();
            }
            ++nbar;
            epi_bar_sync();
        };
        int step_t0 = 0;
        if (FUSED) step_t0 = p.adam_step[net];

        int rpar = 0, it = 0;
#pragma unroll 1
        for (int mb = 0; mb < n_mb; ++mb) {
            const long long mb_start = p.b.mb_start + (long long)mb * batch;
            const int count = min(batch, p.b.mb_count - mb * batch);
            const int ntiles = (count + XT - 1) / XT;
            const float inv_b = 1.0f / (float)count;
            auto tile_rows = [&](int tile, long long* dst) {
                if (tid < XT) {
                    const int local = tile * XT + tid;
                    long long row = -1;
                    if (local < count) {
                        const long long k = mb_start + local;
                        if (p.b.identity_stride > 0) row = k * p.b.identity_stride;
                        else row = p.b.perm ? (long long)p.b.perm[k] : (long long)x3_feistel((unsigned long long)k, (unsigned long long)p.b.total, p.b.perm_seed);
                    }
                    dst[tid] = row;
                }
            };
            const bool have_tiles = (int)blockIdx.x < ntiles;
            if (have_tiles) {
                tile_rows(blockIdx.x, sRowBuf + rpar * XT);
                epi_bar_sync();
                prefetch_x(sRowBuf + rpar * XT);
            }
#pragma unroll 1
            for (int tile = blockIdx.x; tile < ntiles; tile += G, ++it) {
                const uint32_t par = (uint32_t)(it & 1);
                long long* sRow = sRowBuf + rpar * XT;
                long long* sRowNext = sRowBuf + (rpar ^ 1) * XT;
                const bool has_next = tile + G < ntiles;
                // ---- E0: X tile (prefetched registers -> bf16x3) ------------------------------------------------
                if (has_next) tile_rows(tile + G, sRowNext);
                if (it > 0) mbar_wait_a(bar(DONE_C6), par ^ 1u);          // previous tile's dW1 / db1 read X and dZ1
                tc_fence_after();
#pragma unroll
                for (int ph = 0; ph < 2; ++ph) {
                    float v[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = xpre[8 * ph + i];
                    store8_x3(sbase + OFF_X, ACT_SUB, xm, 32 * ph + xc, v);
                    announce(RDY_X0 + ph);
                }
                epi_bar_sync();                                            // next tile's row list is complete
                // ---- E1: H1 = tanh(Z1 + b1) -------------------------------------------------------------------
                mbar_wait_a(bar(DONE_C1), par);
                tc_fence_after();
#pragma unroll
                for (int ph = 0; ph < 2; ++ph) {
                    const int c0 = 32 * ph + 8 * h;
                    float v[8];
                    tmem_ld8(tmem + lane_base + T_ZA + (uint32_t)c0, v);
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = tanh_acc(v[i] + sB1[c0 + i]);
                    store8_x3(sbase + OFF_H1, ACT_SUB, s_row, c0, v);
                    announce(RDY_H1_0 + ph);
                }
                // ---- E2: H2 = tanh(Z2 + b2) -------------------------------------------------------------------
                mbar_wait_a(bar(DONE_C2), par);
                tc_fence_after();
#pragma unroll
                for (int ph = 0; ph < 2; ++ph) {
                    const int c0 = 32 * ph + 8 * h;
                    float v[8];
                    tmem_ld8(tmem + lane_base + T_ZB + (uint32_t)c0, v);
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = tanh_acc(v[i] + sB2[c0 + i]);
                    store8_x3(sbase + OFF_H2, ACT_SUB, s_row, c0, v);
                    announce(RDY_H2_0 + ph);
                }
                // ---- E3: OUT -> loss -> dOUT (warps with h == 0: one thread per sample) -------------------------
                if (h == 0) {
                    float pf_act[16], pf_logp = 0.f, pf_advr = 0.f, pf_advc = 0.f, pf_tv = 0.f;
                    const long long prow = sRow[s_row];
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
                    mbar_wait_a(bar(DONE_C3), par);
                    tc_fence_after();
                    float st[4] = {0.f, 0.f, 0.f, 0.f};   // loss, ratio, kl (unused here), count
                    float dls[16], o16[16], d16[16];
#pragma unroll
                    for (int a = 0; a < 16; ++a) { dls[a] = 0.f; d16[a] = 0.f; }
                    tmem_ld16(tmem + lane_base + T_OUT, o16);
                    if (prow >= 0) {
                        if (net != 0) {
                            const float d = o16[0] + sB3[0] - pf_tv;
                            st[0] = d * d; st[3] = 1.f;
                            d16[0] = 2.f * d * inv_b;
                        } else {
                            float logp_new = 0.f, diff[16];
#pragma unroll
                            for (int a = 0; a < 16; ++a) {
                                diff[a] = 0.f;
                                if (a < A) {
                                    const float sd = sLs[16 + a];
                                    const float d = pf_act[a] - (o16[a] + sB3[a]);
                                    diff[a] = d;
                                    logp_new += -(d * d) / (2.f * sd * sd) - sLs[a] - 0.9189385332046727f;
                                }
                            }
                            const float ratio = expf(logp_new - pf_logp);
                            const float adv_r = (pf_advr - m_r) / s_r;
                            const float adv_c = pf_advc - m_c;
                            const float adv = (adv_r - lam * adv_c) / (1.f + lam);
                            float dlogp, loss;
                            if (p.kind == X3_PPO_CLIP) {
                                const float rc = fminf(fmaxf(ratio, 1.f - p.clip), 1.f + p.clip);
                                const float s1 = ratio * adv, s2 = rc * adv;
                                loss = -fminf(s1, s2);
                                dlogp = (s1 <= s2) ? -adv * ratio * inv_b : 0.f;
                            } else if (p.kind == X3_RATIO) {
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
                                    d16[a] = dlogp * diff[a] * iv;
                                    dls[a] = dlogp * (diff[a] * diff[a] * iv - 1.f);
                                }
                        }
                    }
                    store16_x3_sw32(sbase + OFF_D, D_SUB, s_row, d16);
                    announce(RDY_D);
                    // deterministic reductions over the 128 sample threads
#pragma unroll
                    for (int i = 0; i < 4; ++i) st[i] = warp_sum(st[i]);
                    if (net == 0) {
#pragma unroll
                        for (int a = 0; a < 16; ++a) dls[a] = warp_sum(dls[a]);
                    }
                    float db[16];   // db3[o] = sum_s dOUT[s][o]
#pragma unroll
                    for (int a = 0; a < 16; ++a) db[a] = (a < L.out) ? warp_sum(d16[a]) : 0.f;
                    if (lane == 0) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) sRed[q * 8 + i] = st[i];
#pragma unroll
                        for (int a = 0; a < 16; ++a) { sRed[32 + q * 16 + a] = dls[a]; sRed[96 + q * 16 + a] = db[a]; }
                    }
                    loss_bar_sync();
                    if (tid < 4) sStat[tid] += sRed[tid] + sRed[8 + tid] + sRed[16 + tid] + sRed[24 + tid];
                    if (net == 0 && tid >= 32 && tid < 48) {
                        const int a = tid - 32;
                        sLs[32 + a] += sRed[32 + a] + sRed[48 + a] + sRed[64 + a] + sRed[80 + a];
                    }
                    if (tid >= 64 && tid < 64 + L.out) {
                        const int a = tid - 64;
                        sB3acc[a] += sRed[96 + a] + sRed[112 + a] + sRed[128 + a] + sRed[144 + a];
                    }
                    loss_bar_sync();                          // sRed is rewritten by the next tile
                }
                // ---- E4: dZ2 = (dOUT W3) (1 - H2^2), stored over H2 once dW3 has read it --------------------------
                if (has_next) prefetch_x(sRowNext);                        // next tile's rows fly during the backward half
                mbar_wait_a(bar(DONE_C4A), par);
                tc_fence_after();
                float dz[16];
#pragma unroll
                for (int ph = 0; ph < 2; ++ph) {
                    const int c0 = 32 * ph + 8 * h;
                    float v[8], hh[8];
                    tmem_ld8(tmem + lane_base + T_ZA + (uint32_t)c0, v);
                    load8_x3(sbase + OFF_H2, ACT_SUB, s_row, c0, hh);
#pragma unroll
                    for (int i = 0; i < 8; ++i) dz[8 * ph + i] = v[i] * (1.f - hh[i] * hh[i]);
                }
                mbar_wait_a(bar(DONE_C4B), par);
#pragma unroll
                for (int ph = 0; ph < 2; ++ph) {
                    float v[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = dz[8 * ph + i];
                    store8_x3(sbase + OFF_H2, ACT_SUB, s_row, 32 * ph + 8 * h, v);
                    announce(RDY_DZ2_0 + ph);
                }
                // ---- E5: dZ1 = (dZ2 W2) (1 - H1^2), stored over H1 once dW2 has read it --------------------------
                mbar_wait_a(bar(DONE_C5A), par);
                tc_fence_after();
#pragma unroll
                for (int ph = 0; ph < 2; ++ph) {
                    const int c0 = 32 * ph + 8 * h;
                    float v[8], hh[8];
                    tmem_ld8(tmem + lane_base + T_ZB + (uint32_t)c0, v);
                    load8_x3(sbase + OFF_H1, ACT_SUB, s_row, c0, hh);
#pragma unroll
                    for (int i = 0; i < 8; ++i) dz[8 * ph + i] = v[i] * (1.f - hh[i] * hh[i]);
                }
                mbar_wait_a(bar(DONE_C5B), par);
#pragma unroll
                for (int ph = 0; ph < 2; ++ph) {
                    float v[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = dz[8 * ph + i];
                    store8_x3(sbase + OFF_H1, ACT_SUB, s_row, 32 * ph + 8 * h, v);
                }
                announce(RDY_DZ1);
                rpar ^= 1;
            }
            // ---- this CTA's partial gradient of the minibatch: TMEM accumulators -> global ------------------------
            if (have_tiles) {
                mbar_wait_a(bar(DONE_C6), (uint32_t)((it - 1) & 1));
                tc_fence_after();
                const int t_row = 16 * q + lane;       // row (lane < 16) of the M = 64 accumulators
                const int c16 = 16 * h;
                float v[16];
                tmem_ld16(tmem + lane_base + T_DW2 + (uint32_t)c16, v);
                if (lane < 16) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) __stcg(gout + L.off_w2 + t_row * 64 + c16 + i, v[i]);       // rows of P floats: 4 B aligned only
                }
                tmem_ld16(tmem + lane_base + T_DW1 + (uint32_t)c16, v);
                if (lane < 16) {
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (c16 + i < O) __stcg(gout + L.off_w1 + t_row * O + c16 + i, v[i]);
                }
                if (h == 0) {      // dW3^T [k][o]
                    tmem_ld16(tmem + lane_base + T_DW3, v);
                    if (lane < 16)
#pragma unroll
                        for (int o = 0; o < 16; ++o)
                            if (o < L.out) __stcg(gout + L.off_w3 + o * 64 + t_row, v[o]);
                } else if (h == 1) {
                    tmem_ld16(tmem + lane_base + T_DB1, v);
                    if (lane < 16) __stcg(gout + L.off_b1 + t_row, v[0]);
                } else if (h == 2) {
                    tmem_ld16(tmem + lane_base + T_DB2, v);
                    if (lane < 16) __stcg(gout + L.off_b2 + t_row, v[0]);
                }
                tc_fence_before();
                if (tid < L.out) __stcg(gout + L.off_b3 + tid, sB3acc[tid]);
                if (net == 0 && tid < A) {
                    float g = sLs[32 + tid];
                    if (blockIdx.x == 0 && p.kind == X3_PPO_CLIP) g -= p.entropy_coef / (float)A;
                    __stcg(gout + L.off_logstd + tid, g);
                }
                if (tid < 8) __stcg(p.stats_part + ((size_t)blockIdx.x * 3 + net) * 8 + tid, (tid < 4) ? sStat[tid] : 0.f);
            } else {
                for (int i = tid; i < L.size; i += NEPI) __stcg(gout + i, 0.f);      // no tile of this (short) minibatch
                if (tid < 8) __stcg(p.stats_part + ((size_t)blockIdx.x * 3 + net) * 8 + tid, 0.f);
            }
            if (!FUSED) break;

            // ================= in-kernel optimiser step =========================================================
            net_barrier();                                             // every partial gradient of this network is in L2
            const int S = (L.size + G - 1) / G;                        // parameters owned by this CTA: [p0, p0 + S)
            const int p0 = (int)blockIdx.x * S;
            const int Gh = (G + 1) >> 1;
            float ssq = 0.f, st2 = 0.f;
            for (int base = 0; base < S; base += 256) {
                const int pi = base + (tid & 255), part = tid >> 8;
                const bool valid = pi < S && p0 + pi < L.size;
                float sacc = 0.f;
                if (valid) {
                    const float* src = p.gpart + noff + p0 + pi;
                    const int b1 = min(G, (part + 1) * Gh);
                    for (int b = part * Gh; b < b1; b += 16) {
                        float t[16];
#pragma unroll
                        for (int u = 0; u < 16; ++u) t[u] = (b + u < b1) ? __ldcg(src + (size_t)(b + u) * p.P) : 0.f;
                        sacc += (((t[0] + t[1]) + (t[2] + t[3])) + ((t[4] + t[5]) + (t[6] + t[7]))) +
                                (((t[8] + t[9]) + (t[10] + t[11])) + ((t[12] + t[13]) + (t[14] + t[15])));
                    }
                }
                sPart[tid] = sacc;
                epi_bar_sync();
                if (part == 0 && valid) {
                    const int qg = noff + p0 + pi;
                    float g = sPart[tid] + sPart[256 + tid];
                    const float th = __ldcg(p.theta_rw + qg);
                    if (net != 0 && p.critic_norm_coef > 0.f) { g += 2.f * p.critic_norm_coef * th; st2 += th * th; }
                    __stcg(p.grad + qg, g);
                    ssq += g * g;
                }
                epi_bar_sync();
            }
            ssq = warp_sum(ssq); st2 = warp_sum(st2);
            if (lane == 0) { sPart[warp] = ssq; sPart[16 + warp] = st2; }
            epi_bar_sync();
            if (tid == 0) {
                float a = 0.f, b = 0.f;
                for (int w = 0; w < 16; ++w) { a += sPart[w]; b += sPart[16 + w]; }
                __stcg(p.sumsq_part + (net * 2 + 0) * G + blockIdx.x, a);
                __stcg(p.sumsq_part + (net * 2 + 1) * G + blockIdx.x, b);
            }
            net_barrier();                                             // every slice norm of this network is in L2
            const int step_t = step_t0 + mb + 1;
            if (warp == 0) {
                float tot = 0.f, t2 = 0.f;
                for (int b = lane; b < G; b += 32) { tot += __ldcg(p.sumsq_part + (net * 2 + 0) * G + b); t2 += __ldcg(p.sumsq_part + (net * 2 + 1) * G + b); }
                tot = warp_sum(tot); t2 = warp_sum(t2);
                if (lane == 0) {
                    sScal[0] = (p.max_grad_norm > 0.f) ? fminf(p.max_grad_norm / (sqrtf(tot) + 1e-6f), 1.0f) : 1.0f;
                    const double bc1 = 1.0 - pow(0.9, (double)step_t), bc2 = 1.0 - pow(0.999, (double)step_t);
                    sScal[1] = (float)((double)p.lr[net] / bc1);
                    sScal[2] = (float)sqrt(bc2);
                    sScal[3] = t2;
                }
            } else if (warp == 1 && blockIdx.x == 0) {               // loss statistics of this minibatch (logger means)
                float acc[4] = {0.f, 0.f, 0.f, 0.f};
                for (int b = lane; b < G; b += 32)
#pragma unroll
                    for (int i = 0; i < 4; ++i) acc[i] += __ldcg(p.stats_part + ((size_t)b * 3 + net) * 8 + i);
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[i] = warp_sum(acc[i]);
                if (lane == 0) { sScal[4] = acc[0]; sScal[5] = acc[1]; sScal[6] = acc[2]; sScal[7] = acc[3]; }
            }
            epi_bar_sync();
            if (blockIdx.x == 0 && tid == 0) {
                const float inv = sScal[7] > 0.f ? 1.f / sScal[7] : 0.f;
                float* ts = p.train_stats + net * 8;
                ts[0] += sScal[4] * inv + ((net != 0) ? p.critic_norm_coef * sScal[3] : 0.f);
                ts[1] += sScal[5] * inv;
                ts[2] += sScal[6] * inv;
                ts[3] += 1.f;
            }
            const float clipc = sScal[0], step_size = sScal[1], bc2_sqrt = sScal[2];
            const unsigned int xstep = p.step_base + (unsigned int)mb;
            const int xpar = (int)(xstep & 1u);
            const int cta_g = (gridDim.y == 1 ? 0 : net) * G + (int)blockIdx.x;
            bool xfail = false;
            if (p.world > 1) {
                // push the clipped slice into every rank's receive buffer [parity][source rank][P], then raise this
                // CTA's flag on every rank; the slices of different CTAs travel independently (no grid barrier)
                for (int base = 0; base < S; base += NEPI) {
                    const int pi = base + tid;
                    if (pi < S && p0 + pi < L.size) {
                        const int qg = noff + p0 + pi;
                        const float gc = __ldcg(p.grad + qg) * clipc;
                        for (int r = 0; r < p.world; ++r)
                            asm volatile("st.relaxed.sys.global.f32 [%0], %1;" ::"l"(p.peer_buf[r] + ((size_t)(xpar * p.world + p.rank)) * p.P + qg), "f"(gc) : "memory");
                    }
                }
                __threadfence_system();
                epi_bar_sync();
                if (tid < p.world) {
                    const int fbase = 2 * p.world + (xpar * p.world) * 160;
                    st_release_sys_u32(p.peer_flag[tid] + fbase + p.rank * 160 + cta_g, xstep);
                    const unsigned int* f = p.peer_flag[p.rank] + fbase + tid * 160 + cta_g;
                    const long long t0 = clock64();
                    while (ld_acquire_sys_u32(f) != xstep) {
                        if (clock64() - t0 > 20000000000LL) { *p.error_flag = 1; sScal[0] = -1.f; break; }   // ~10 s: fail loudly, never hang the GPU
                    }
                }
                epi_bar_sync();
                xfail = sScal[0] < 0.f;
            }
            if (!xfail) {
                for (int base = 0; base < S; base += NEPI) {
                    const int pi = base + tid;
                    if (pi < S && p0 + pi < L.size) {
                        const int qg = noff + p0 + pi;
                        float g;
                        if (p.world > 1) {
                            float sum = 0.f;
                            for (int r = 0; r < p.world; ++r) {
                                float v;
                                asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p.peer_buf[p.rank] + ((size_t)(xpar * p.world + r)) * p.P + qg) : "memory");
                                sum += v;
                            }
                            g = sum / (float)p.world;
                        } else {
                            g = __ldcg(p.grad + qg) * clipc;
                        }
                        __stcg(p.grad + qg, g);
                        const float th = __ldcg(p.theta_rw + qg);
                        float m = __ldcg(p.adam_m + qg), v = __ldcg(p.adam_v + qg);
                        m = __fadd_rn(m, __fmul_rn(0.1f, __fadd_rn(g, -m)));                       // exp_avg.lerp_(grad, 1 - beta1)
                        v = __fadd_rn(__fmul_rn(v, 0.999f), __fmul_rn(__fmul_rn(0.001f, g), g));   // mul_(beta2).addcmul_(g, g, 1 - beta2)
                        const float denom = __fadd_rn(__fdiv_rn(sqrtf(v), bc2_sqrt), 1e-8f);
                        __stcg(p.theta_rw + qg, __fadd_rn(th, __fmul_rn(-step_size, __fdiv_rn(m, denom))));
                        __stcg(p.adam_m + qg, m); __stcg(p.adam_v + qg, v);
                    }
                }
            }
            net_barrier();                                             // the new parameters of this network are in L2
            stage_weights_x3(sbase, misc, theta, L, net, O, A, tid);   // visible to the tensor core with the next X announce
            if (tid < 16) { misc[MF_LS + 32 + tid] = 0.f; misc[MF_B3ACC + tid] = 0.f; }
            if (tid < 8) misc[MF_STAT + tid] = 0.f;
            epi_bar_sync();
        }
        if (FUSED && blockIdx.x == 0 && tid == 0) p.adam_step[net] = step_t0 + n_mb;     // every CTA read it before the first barrier
    }
    tc_fence_before();
    __syncthreads();
    if (is_mma_warp) tmem_dealloc(tmem, T_COLS);
}

}  // namespace osb

using namespace osb;

extern "C" {

int osb_tc_grid_blocks(long long rows, int net_mask);

static int x3_set_attr() {
    static bool attr = false;
    if (!attr) {
        const size_t smem = 1024 + X3_SMEM;
        OSB_CUDA(cudaFuncSetAttribute(minibatch_grad_x3_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        OSB_CUDA(cudaFuncSetAttribute(minibatch_grad_x3_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = true;
    }
    return OSB_OK;
}

// Split-bf16 (parity-grade tensor-core) variant of osb_minibatch_grad: same arguments, O <= 64, A <= 16,
// loss kinds PPO-clip / ratio / cost surrogate.  gpart holds osb_tc_grid_blocks(mb_count, net_mask) rows of P floats.
int osb_minibatch_grad_x3(const float* theta, int O, int A, const float* obs, const float* act,
                          const float* logp, const float* adv_r, const float* adv_c,
                          const float* tv_r, const float* tv_c, const float* mu_old,
                          const float* moments, const int* perm, long long total, unsigned perm_seed,
                          long long mb_start, int mb_count, int loss_kind, float clip,
                          float entropy_coef, float focops_lam, float focops_eta,
                          const float* lagrange, const float* logstd_old, int net_mask, float* gpart,
                          float* stats_part, const int* stop_flag, void* stream) {
    (void)mu_old; (void)focops_lam; (void)focops_eta; (void)logstd_old;
    OSB_CHECK_ARG(theta && obs && act && logp && adv_r && adv_c && tv_r && tv_c && moments, "null input");
    OSB_CHECK_ARG(O > 0 && O <= 64 && A > 0 && A <= 16 && mb_count > 0 && total > 0, "bf16x3 path needs O <= 64, A <= 16");
    OSB_CHECK_ARG(mb_start >= 0 && mb_start + mb_count <= total, "minibatch window out of range");
    OSB_CHECK_ARG(loss_kind == X3_PPO_CLIP || loss_kind == X3_RATIO || loss_kind == X3_COST, "loss kind not on the bf16x3 path");
    OSB_CHECK_ARG(net_mask > 0 && net_mask < 8, "net_mask");
    X3Args p = {};
    p.b = X3Batch{obs, act, logp, adv_r, adv_c, tv_r, tv_c, moments, perm, total, perm_seed, mb_start, mb_count, 0};
    p.kind = loss_kind; p.clip = clip; p.entropy_coef = entropy_coef; p.lagrange = lagrange;
    p.theta = theta; p.gpart = gpart; p.stats_part = stats_part; p.stop_flag = stop_flag;
    p.O = O; p.A = A; p.P = actor_layout(O, A).size + 2 * critic_layout(O, A).size; p.net_mask = net_mask;
    p.batch_size = mb_count; p.world = 1;
    const int nb = osb_tc_grid_blocks(mb_count, net_mask);
    int rc = x3_set_attr();
    if (rc) return rc;
    const bool single = (net_mask & (net_mask - 1)) == 0;
    minibatch_grad_x3_kernel<false><<<dim3(nb, single ? 1 : 3), NTX3, 1024 + X3_SMEM, (cudaStream_t)stream>>>(p);
    OSB_LAUNCH_CHECK();
    return OSB_OK;
}

// One update iteration of PolicyGradient._update (policy_gradient.py:L369-381) as ONE persistent cooperative
// kernel on bf16x3 tiles: all minibatches of [0, total) in steps of batch_size, each = fused forward + loss +
// backward, fixed-order partial reduction, per-network clip_grad_norm_, (world > 1: clipped-gradient exchange
// over NVLink peer memory, policy_gradient.py:L437-443 / distributed.py:L193-198) and torch-Adam, with the
// parameters re-staged in shared memory between minibatches.  perm: [total] slab rows of this iteration or NULL
// (Feistel order keyed by perm_seed).  gpart: osb_tc_grid_blocks(batch_size, net_mask) rows of P floats.
// peer_buf / peer_flag: device arrays of `world` pointers ([2][world][P] floats, [2*world + 2*world*160] flags)
// or NULL for one rank.
int osb_ppo_update_iter_x3(float* theta, float* grad, float* adam_m, float* adam_v, int* adam_step, int O, int A,
                           const float* obs, const float* act, const float* logp, const float* adv_r,
                           const float* adv_c, const float* tv_r, const float* tv_c, const float* moments,
                           const int* perm, long long total, unsigned perm_seed, int batch_size, int loss_kind,
                           float clip, float entropy_coef, const float* lagrange, int net_mask,
                           float critic_norm_coef, float max_grad_norm, float lr_actor, float lr_critic_r,
                           float lr_critic_c, float* gpart, float* stats_part, float* train_stats,
                           const int* stop_flag, void* peer_buf, void* peer_flag, int world, int rank,
                           int* p2p_error, void* stream) {
    OSB_CHECK_ARG(theta && grad && adam_m && adam_v && adam_step && obs && act && logp && adv_r && adv_c && tv_r && tv_c && moments, "null input");
    OSB_CHECK_ARG(O > 0 && O <= 64 && A > 0 && A <= 16 && batch_size > 0 && total > 0 && total < (1ll << 31), "bf16x3 path needs O <= 64, A <= 16");
    OSB_CHECK_ARG(loss_kind == X3_PPO_CLIP || loss_kind == X3_RATIO || loss_kind == X3_COST, "loss kind not on the bf16x3 path");
    OSB_CHECK_ARG(net_mask > 0 && net_mask < 8 && gpart && stats_part && train_stats, "bad argument");
    OSB_CHECK_ARG(world >= 1 && (world == 1 || (peer_buf && peer_flag && p2p_error && rank >= 0 && rank < world && world <= 64)), "bad p2p argument");
    static float* d_ws = nullptr;            // [0, 4): barrier counters (u32); [64, 64 + 6 * 148): slice norms
    static unsigned int step_base = 0;       // identical on every rank: same call sequence
    cudaStream_t s = (cudaStream_t)stream;
    if (!d_ws) OSB_CUDA(cudaMalloc(&d_ws, (64 + 6 * 148) * sizeof(float)));
    OSB_CUDA(cudaMemsetAsync(d_ws, 0, 4 * sizeof(unsigned int), s));
    X3Args p = {};
    p.b = X3Batch{obs, act, logp, adv_r, adv_c, tv_r, tv_c, moments, perm, total, perm_seed, 0, (int)total, 0};
    p.kind = loss_kind; p.clip = clip; p.entropy_coef = entropy_coef; p.lagrange = lagrange;
    p.theta = theta; p.gpart = gpart; p.stats_part = stats_part; p.stop_flag = stop_flag;
    p.O = O; p.A = A; p.P = actor_layout(O, A).size + 2 * critic_layout(O, A).size; p.net_mask = net_mask;
    p.batch_size = batch_size; p.theta_rw = theta; p.grad = grad; p.adam_m = adam_m; p.adam_v = adam_v; p.adam_step = adam_step;
    p.critic_norm_coef = critic_norm_coef; p.max_grad_norm = max_grad_norm;
    p.lr[0] = lr_actor; p.lr[1] = lr_critic_r; p.lr[2] = lr_critic_c;
    p.sumsq_part = d_ws + 64; p.train_stats = train_stats; p.bar_ctr = reinterpret_cast<unsigned int*>(d_ws);
    p.peer_buf = (float* const*)peer_buf; p.peer_flag = (unsigned int* const*)peer_flag;
    p.world = world; p.rank = rank; p.error_flag = p2p_error;
    const int n_mb = (int)((total + batch_size - 1) / batch_size);
    p.step_base = step_base + 1u;
    step_base += (unsigned int)n_mb;
    const int first = (int)(total < batch_size ? total : batch_size);
    const int nb = osb_tc_grid_blocks(first, net_mask);
    int rc = x3_set_attr();
    if (rc) return rc;
    const bool single = (net_mask & (net_mask - 1)) == 0;
    void* args[] = {&p};
    OSB_CUDA(cudaLaunchCooperativeKernel((void*)minibatch_grad_x3_kernel<true>, dim3(nb, single ? 1 : 3), dim3(NTX3), args,
                                         1024 + X3_SMEM, s));
    return OSB_OK;
}

}  // extern "C"
