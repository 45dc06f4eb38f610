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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
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
// one try (the hardware suspends the thread for a bounded time while the phase is pending): true = the phase with `parity` is complete
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return done != 0;
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

// ---------------------------------------------------------------------------------------------------------------
// The A tile of the first layer.  The tensor-core kernels are free to order the K dimension as they like as long as
// the A tile and the packed W1 agree, so internally the 198 features are permuted so that every 16-byte chunk (8
// bf16) of a row is exactly TWO WHOLE POINTS:
//     internal column  0.. 95   PLAYER1 points 0..23, 4 units each      (reference features   0.. 95)
//                     96..191   PLAYER2 points 0..23                    (reference features  98..193)
//                    192..197   bar1/2, off1/15, bar2/2, off2/15, turn flags  (features 96, 97, 194, 195, 196, 197)
//                    198, 199   1.0, 1.0 (they multiply the folded bias hi / lo of W1);   200..207 zero
// (bg_w1_source_column below is the same map for bg_pack_w1.)  Chunk k < 24 is then two 8-byte loads from the 16-entry
// units table (128 bytes = one entry per bank pair: conflict-free for any index pattern), indexed by the two adjacent
// count bytes of board word k/2.  (A 256-entry table of point PAIRS, one 16-byte load per chunk, was measured slower:
// random 16-byte entries conflict in the banks -- 80 us vs 77 us per 1.2 M rows.)
__host__ __device__ __forceinline__ int bg_w1_source_column(int j) {      // internal column -> reference feature (-1: none)
    if (j < 96) return j;
    if (j < 192) return j + 2;
    switch (j) { case 192: return 96; case 193: return 97; case 194: return 194; case 195: return 195; case 196: return 196; case 197: return 197; }
    return -1;
}
// shared-memory copy of the bar/2 and off/15 tables, placed right after the 16 units entries: lut[16 + k/2] holds
// entries k of both tables for two k (a per-thread index into __constant__ memory is a serialised, long-latency load:
// it was a third of the producers' stall samples)
struct FeatureLut {
    uint2 units[16];
    uint16_t half[16];
    uint16_t off15[16];
};
__device__ __forceinline__ void load_feature_lut(FeatureLut* t) {
    if (threadIdx.x < 16) { t->units[threadIdx.x] = kUnitsBf16[threadIdx.x]; t->half[threadIdx.x] = kHalfBf16[threadIdx.x]; t->off15[threadIdx.x] = kOff15Bf16[threadIdx.x]; }
}
__device__ __forceinline__ uint32_t bar_off_pair_s(uint32_t bar, uint32_t off, const FeatureLut* t) {
    return (uint32_t)t->half[bar & 15u] | ((uint32_t)t->off15[off & 15u] << 16);
}
// chunk kc (compile-time constant after unrolling) of the internal feature row of a position held as 13 words
__device__ __forceinline__ uint4 feature_chunk_regs(const uint32_t (&w)[kBoardWords], int flag, int kc, const FeatureLut* ft) {
    const uint2* lut = ft->units;
    if (kc < 24) {
        const uint32_t x = w[kc >> 1] >> (16 * (kc & 1));                    // counts of points 2kc, 2kc+1 in the low two bytes
        const uint2 a = lut[x & 15u], c = lut[(x >> 8) & 15u];
        return make_uint4(a.x, a.y, c.x, c.y);
    }
    if (kc == 24) {
        const uint32_t m = w[12];                                            // bar1, bar2, off1, off2
        return make_uint4(bar_off_pair_s(m, m >> 16, ft), bar_off_pair_s(m >> 8, m >> 24, ft),
                          flag == 0 ? 0x00003F80u : 0x3F800000u, 0x3F803F80u);
    }
    return make_uint4(0u, 0u, 0u, 0u);
}
// (rows beyond the batch in the last tile are built from whatever the staging buffer holds: their accumulators are
// finite and never stored)
template <int HALF>
__device__ __forceinline__ void build_half_row(const uint32_t (&w)[kBoardWords], int flag, const FeatureLut* lut, uint32_t tmem_row) {
#pragma unroll
    for (int i = 0; i < 13; ++i) {
        const int kc = HALF * 13 + i;
        const uint4 v = feature_chunk_regs(w, flag, kc, lut);
        tmem_st4(tmem_row + (uint32_t)(kc * 4), v);              // chunk kc = bf16 columns 8kc..8kc+7 = TMEM columns 4kc..4kc+3
    }
}

// chunks [K0, K1) of the row (compile-time bounds): the share of one of several producer threads of a position; only the
// board words those chunks read are loaded from the staged row
template <int K0, int K1>
__device__ __forceinline__ void build_row_chunks(const uint32_t* srow, int flag, const FeatureLut* lut, uint32_t tmem_row) {
    uint32_t w[kBoardWords];
#pragma unroll
    for (int i = 0; i < kBoardWords; ++i) {
        const bool used = (i < 12 && 2 * i + 1 >= K0 && 2 * i < K1) || (i == 12 && K1 > 24);
        w[i] = used ? srow[i] : 0u;
    }
#pragma unroll
    for (int kc = K0; kc < K1; ++kc) {
        const uint4 v = feature_chunk_regs(w, flag, kc, lut);
        tmem_st4(tmem_row + (uint32_t)(kc * 4), v);
    }
}

}  // namespace bg
