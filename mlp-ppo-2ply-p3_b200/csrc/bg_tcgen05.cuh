// bg_tcgen05.cuh -- thin inline-PTX wrappers (tcgen05 MMA / TMEM / mbarrier / cp.async) and the register-level
// feature-row builder shared by the tensor-core kernels K4 (mlp.cu) and the policy head (policy.cu).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "bg_device.cuh"
#include "bg_features.cuh"
#include "bg_b200.h"

namespace bg {

constexpr int kTileM = 128;            // positions per tile = UMMA M
constexpr int kHidden = BG_HIDDEN;     // UMMA N of the first layer
constexpr int kKPad = BG_FEAT_LD_BF16; // 208 = 13 x UMMA K
constexpr int kChunks = kKPad / 8;     // 26 sixteen-byte chunks per feature row

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
    // SWIZZLE_NONE, K-major: start >> 4 | LBO(2048) >> 4 << 16 | SBO(128) >> 4 << 32 | version 1 << 46
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(2048u >> 4) << 16) | ((uint64_t)(128u >> 4) << 32) |
           (1ull << 46);
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, N=128, M=128
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kHidden >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);

// D[tmem_d] (+)= A[tmem_a] * B[smem desc]
__device__ __forceinline__ void mma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
        :: "r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
// 16 bytes (8 bf16 of one row) -> 4 consecutive TMEM columns of this thread's lane
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint4& v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};\n"
                 :: "r"(taddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" :: "r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async4_s(uint32_t smem_addr, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" :: "r"(smem_addr), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16_s(uint32_t smem_addr, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" :: "r"(smem_addr), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}\n" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void umma_commit(unsigned long long* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" :: "r"(smem_u32(bar)) : "memory");
}

// general K-major SWIZZLE_NONE descriptor: core matrices of 8 rows x 16 B; lbo = byte distance between 8-column
// (16-byte) K chunks, sbo = byte distance between 8-row groups
__device__ __forceinline__ uint64_t make_smem_desc_kmajor(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major
constexpr uint32_t make_idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// board52 byte `idx` (a compile-time constant after unrolling) of a row held as 13 words in registers
__device__ __forceinline__ int reg_byte(const uint32_t (&w)[kBoardWords], int idx) { return (int)((w[idx >> 2] >> (8 * (idx & 3))) & 15u); }
// chunk k of the feature row, k constant after unrolling (same content as feature_chunk_lut)
__device__ __forceinline__ uint4 feature_chunk_regs(const uint32_t (&w)[kBoardWords], int flag, int k, const uint2* lut) {
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    if (k < 12) {
        uint2 a = lut[reg_byte(w, 2 * k)], c = lut[reg_byte(w, 2 * k + 1)];
        o = make_uint4(a.x, a.y, c.x, c.y);
    } else if (k == 12) {
        uint2 a = lut[reg_byte(w, 24)], c = lut[reg_byte(w, 25)];
        o = make_uint4(bar_off_pair_bf16(reg_byte(w, 48), reg_byte(w, 50)), a.x, a.y, c.x);
    } else if (k < 24) {
        const int q = 2 * (k - 12) - 1;
        uint2 a = lut[reg_byte(w, 24 + q)], c = lut[reg_byte(w, 25 + q)], e = lut[reg_byte(w, 26 + q)];
        o = make_uint4(a.y, c.x, c.y, e.x);
    } else if (k == 24) {
        o.x = lut[reg_byte(w, 47)].y;
        o.y = bar_off_pair_bf16(reg_byte(w, 49), reg_byte(w, 51));
        o.z = flag == 0 ? 0x00003F80u : 0x3F800000u;
        o.w = 0x3F803F80u;                           // columns 198, 199 = 1.0: multiply the folded bias (hi, lo) of W1
    }
    return o;
}
// (rows beyond the batch in the last tile are built from whatever the staging buffer holds: their accumulators are
// finite and never stored)
template <int HALF>
__device__ __forceinline__ void build_half_row(const uint32_t (&w)[kBoardWords], int flag, const uint2* lut, uint32_t tmem_row) {
#pragma unroll
    for (int i = 0; i < 13; ++i) {
        const int kc = HALF * 13 + i;
        const uint4 v = feature_chunk_regs(w, flag, kc, lut);
        tmem_st4(tmem_row + (uint32_t)(kc * 4), v);              // chunk kc = bf16 columns 8kc..8kc+7 = TMEM columns 4kc..4kc+3
    }
}

}  // namespace bg
