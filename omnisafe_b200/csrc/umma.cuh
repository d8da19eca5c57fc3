// tcgen05 / TMEM / mbarrier helpers (inline PTX, sm_100a) for the tensor-core MLP tiles.
//
// Operand tiles live in shared memory as fp32 in the canonical 128-byte-swizzled layout
// (rows of 128 B = 32 floats, 8-row groups of 1024 B, 16-byte chunks XOR-ed with row % 8).  The SAME
// physical tile [R rows][32 floats] can be consumed
//   * K-major  (rows = M/N index, the 32 floats of a row run along K), or
//   * MN-major (rows = K index, the 32 floats of a row run along M/N),
// which is what lets one stored activation tile feed both the forward GEMM and the weight-gradient
// GEMM without a transposed copy.  kind::tf32 reads the fp32 words directly (10-bit mantissa used).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace osb {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// byte offset of element (r, c) inside a tile of R rows x C floats (C % 32 == 0) stored as C/32 atoms
// of [R][32] floats, each atom 128B-swizzled.  Tile base must be 1024-byte aligned.
__device__ __forceinline__ uint32_t sw128_offset(int r, int c, int R) {
    const int atom = c >> 5, cc = c & 31;
    return (uint32_t)(atom * R * 128 + r * 128 + ((((cc >> 2) ^ (r & 7)) << 4) | ((cc & 3) << 2)));
}

// Shared-memory matrix descriptor, 128B swizzle (SM100 UMMA SmemDescriptor: start address [0,14),
// leading byte offset [16,30), stride byte offset [32,46), version=1 [46,48), layout type [61,64)).
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46) | (2ull << 61);
}
// K-major operand: 8-row groups are 1024 B apart; LBO is unused for swizzled K-major (set to 16 B).
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t saddr) { return desc_sw128(saddr, 16, 1024); }
// MN-major operand: LBO = distance between 32-element blocks along M/N (= one atom = rows*128 B),
// SBO = distance between 8-row K groups (1024 B).
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t saddr, uint32_t atom_bytes) {
    return desc_sw128(saddr, atom_bytes, 1024);
}

// Instruction descriptor for kind::tf32, fp32 accumulate (SM100 InstrDescriptor bit layout).
__device__ __forceinline__ uint32_t idesc_tf32(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread.
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// make the mbarrier track completion of all previously issued MMAs of this thread
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(
                     smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() {
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

// generic-proxy smem writes -> visible to the async proxy (tensor core operand reads)
__device__ __forceinline__ void fence_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
}

// TMEM allocation: one full warp; ncols power of two >= 32; result written to *slot (smem).
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(slot)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(ncols) : "memory");
}

// 32 lanes x 32 consecutive columns: thread i of the warp receives lane (warp%4)*32 + i.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// issue D[tmem] (+)= A * B^T, both K-major SW128 tiles with RA / RB rows; one thread.
__device__ __forceinline__ void tc_gemm(uint32_t tmem_d, uint32_t a_base, int RA, uint32_t b_base, int RB,
                                        int M, int N, int K, bool accumulate) {
    const uint32_t idesc = idesc_tf32(M, N, 0, 0);
    for (int ks = 0; ks < K / 8; ++ks) {
        const uint32_t offA = (uint32_t)((ks >> 2) * RA * 128 + (ks & 3) * 32);
        const uint32_t offB = (uint32_t)((ks >> 2) * RB * 128 + (ks & 3) * 32);
        mma_tf32(tmem_d, desc_kmajor(a_base + offA), desc_kmajor(b_base + offB), idesc,
                 (accumulate || ks > 0) ? 1u : 0u);
    }
}

// round-to-nearest TF32 (the MMA would otherwise truncate the low 13 mantissa bits: biased)
__device__ __forceinline__ float tf32r(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}
__device__ __forceinline__ float tanh_fast(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// explicit shared-window (32-bit address) accessors: keeps every tile access an LDS/STS
__device__ __forceinline__ float lds(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts(uint32_t a, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory");
}
__device__ __forceinline__ uint32_t tile_addr(uint32_t base, int r, int c, int R) {
    return base + sw128_offset(r, c, R);
}
// store / load 32 consecutive columns [c0, c0+32) (c0 % 32 == 0) of row r
__device__ __forceinline__ void store_row32(uint32_t base, int r, int c0, int R, const float (&v)[32]) {
    const uint32_t row = base + (uint32_t)((c0 >> 5) * R * 128 + r * 128);
#pragma unroll
    for (int i = 0; i < 8; ++i)
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(row + (uint32_t)((i ^ (r & 7)) << 4)),
                     "f"(tf32r(v[4 * i])), "f"(tf32r(v[4 * i + 1])), "f"(tf32r(v[4 * i + 2])), "f"(tf32r(v[4 * i + 3]))
                     : "memory");
}
__device__ __forceinline__ void load_row32(uint32_t base, int r, int c0, int R, float (&v)[32]) {
    const uint32_t row = base + (uint32_t)((c0 >> 5) * R * 128 + r * 128);
#pragma unroll
    for (int i = 0; i < 8; ++i)
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(v[4 * i]), "=f"(v[4 * i + 1]), "=f"(v[4 * i + 2]), "=f"(v[4 * i + 3])
                     : "r"(row + (uint32_t)((i ^ (r & 7)) << 4)));
}


}  // namespace umma
}  // namespace osb
