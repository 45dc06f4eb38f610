// mlp.cu -- K4: the MLP leaf evaluator on the tcgen05 tensor cores, fused with the feature encoding.
//
// Replaces the value path of BackgammonPolicyNetwork.forward (src/agent/policy_network.py:58-75)
//     v = value_head( relu( fc1(x) ) ),   x = the 198 features of a position (K3's encoding)
// for B positions given as board52 + turn flag.  The 198-wide bf16 feature rows never touch HBM:
// a CTA stages 128 boards in shared memory, expands them into the tcgen05 K-major operand layout,
// multiplies by the resident W1 tile (128 hidden x 208) with 13 tcgen05.mma (M128 N128 K16,
// bf16 x bf16 -> f32 in TMEM) and reduces the hidden layer in the epilogue straight out of TMEM
// (tcgen05.ld): value = b_v + sum_h w_v[h] * relu(acc[h] + b1[h]).  HBM traffic: 53 B in, 4 B out per position.
//
// Shared-memory operand layout (SWIZZLE_NONE "interleave", K-major): element (row r, column k) of a
// 128 x 208 bf16 tile lives at  (k/8)*2048 + r*16 + (k%8)*2  bytes: core matrices of 8 rows x 16 bytes
// are contiguous (128 B), 8-row groups advance by SBO = 128 B, 8-column chunks by LBO = 2048 B.
#include <cuda_bf16.h>
#include "bg_device.cuh"
#include "bg_features.cuh"
#include "bg_internal.h"

namespace bg {

constexpr int kTileM = 128;            // positions per tile = UMMA M
constexpr int kHidden = BG_HIDDEN;     // UMMA N
constexpr int kKPad = BG_FEAT_LD_BF16; // 208 = 13 x UMMA K
constexpr int kChunks = kKPad / 8;     // 26 sixteen-byte chunks per row
constexpr int kOperandBytes = kChunks * kTileM * 16;   // 53,248
constexpr int kMlpThreads = 128;

struct MlpSmem {
    uint8_t A[kOperandBytes];          // feature tile, rebuilt per 128 positions
    uint8_t W[kOperandBytes];          // W1 tile, resident
    uint32_t boards[kTileM * kBoardWords];
    float b1[kHidden];
    float wv[kHidden];
    int8_t flag[kTileM];
    uint2 units[16];
    unsigned long long mbar;
    uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
    // SWIZZLE_NONE, K-major: start >> 4 | LBO(2048) >> 4 << 16 | SBO(128) >> 4 << 32 | version 1 << 46
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(2048u >> 4) << 16) | ((uint64_t)(128u >> 4) << 32) |
           (1ull << 46);
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, N=128, M=128
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kHidden >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);

__device__ __forceinline__ void mma_bf16_ss(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// win reward of a finished game seen from `p`, who has just borne off 15 (environment/backgammon_env.py:156-171,365-405)
__device__ __forceinline__ float win_reward(const int8_t* b, int p) {
    const int o = p ^ 1;
    if (b[50 + o] != 0) return 1.0f;
    bool bgm = b[48 + o] > 0;
    const int8_t* orow = b + 24 * o;
    const int h0 = p == 0 ? 18 : 0;
#pragma unroll
    for (int i = 0; i < 6; ++i) bgm = bgm || orow[h0 + i] > 0;
    return bgm ? 2.0f : 1.5f;
}

// terminal_aware: a row whose flag player has 15 men off gets the win reward instead of the network value
// (leaf rule of the 2-ply search).
__global__ void __launch_bounds__(kMlpThreads, 1) mlp_value_kernel(
    const int8_t* __restrict__ boards, const int8_t* __restrict__ flags, int flag_all, int flip_flags, long long B,
    const unsigned long long* __restrict__ n_rows_dev, const uint16_t* __restrict__ w1 /*[128][208] bf16*/,
    const float* __restrict__ b1, const float* __restrict__ wv, float bv, int terminal_aware,
    float* __restrict__ values) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    MlpSmem& S = *reinterpret_cast<MlpSmem*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5;
    if (n_rows_dev) B = min(B, (long long)*n_rows_dev);

    // ---- one-time setup: W1 into the operand layout, biases, mbarrier, TMEM
    for (int c = tid; c < kChunks * kHidden; c += kMlpThreads) {
        int kc = c / kHidden, n = c - kc * kHidden;             // consecutive threads -> consecutive n: conflict-free stores
        uint4 v = *reinterpret_cast<const uint4*>(w1 + (size_t)n * kKPad + kc * 8);
        *reinterpret_cast<uint4*>(S.W + kc * 2048 + n * 16) = v;
    }
    load_units_lut(S.units);
    S.b1[tid] = b1[tid];
    S.wv[tid] = wv[tid];
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" :: "r"(smem_u32(&S.mbar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n"
                     :: "r"(smem_u32(&S.tmem_base)), "r"((uint32_t)kHidden) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");     // W tile visible to the tensor-core proxy
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = S.tmem_base;
    const uint32_t a_addr = smem_u32(S.A), w_addr = smem_u32(S.W);
    uint32_t phase = 0;

    const long long n_tiles = (B + kTileM - 1) / kTileM;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long row0 = tile * kTileM;
        const int rows = (int)min((long long)kTileM, B - row0);
        // ---- stage boards (coalesced) and flags
        const uint32_t* src = reinterpret_cast<const uint32_t*>(boards + row0 * kBoardBytes);
        for (int i = tid; i < rows * kBoardWords; i += kMlpThreads) S.boards[i] = __ldg(src + i);
        if (tid < rows) S.flag[tid] = (int8_t)(((flags ? flags[row0 + tid] : flag_all) ^ flip_flags) & 1);
        __syncthreads();
        // ---- build the A tile: thread = row, 26 chunks of 16 B (rows past the end are zero)
        {
            const int8_t* b = reinterpret_cast<const int8_t*>(S.boards) + tid * kBoardBytes;
            const int fl = S.flag[tid];
#pragma unroll 2
            for (int kc = 0; kc < kChunks; ++kc) {
                uint4 v = tid < rows ? feature_chunk_lut(b, fl, kc, S.units) : make_uint4(0u, 0u, 0u, 0u);
                *reinterpret_cast<uint4*>(S.A + kc * 2048 + tid * 16) = v;
            }
        }
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        __syncthreads();
        // ---- 13 MMAs issued by one thread, completion signalled on the mbarrier
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll
            for (int ks = 0; ks < kKPad / 16; ++ks) {
                uint64_t da = make_smem_desc(a_addr + ks * 2 * 2048);
                uint64_t db = make_smem_desc(w_addr + ks * 2 * 2048);
                mma_bf16_ss(tmem, da, db, kIdesc, ks > 0 ? 1u : 0u);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
                         :: "r"(smem_u32(&S.mbar)) : "memory");
        }
        // ---- wait for the accumulator
        {
            uint32_t done = 0;
            const uint32_t bar = smem_u32(&S.mbar);
            while (!done) {
                asm volatile(
                    "{\n\t.reg .pred p;\n\t"
                    "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                    "selp.u32 %0, 1, 0, p;\n\t}\n"
                    : "=r"(done) : "r"(bar), "r"(phase) : "memory");
            }
            phase ^= 1u;
        }
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        // ---- epilogue: thread t of warp w owns TMEM lane 32w+t = row 32w+t
        float v = bv;
#pragma unroll 1
        for (int c0 = 0; c0 < kHidden; c0 += 32) {
            float acc[32];
            tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, acc);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                float h = acc[j] + S.b1[c0 + j];
                v = fmaf(S.wv[c0 + j], fmaxf(h, 0.0f), v);
            }
        }
        if (tid < rows) {
            if (terminal_aware) {
                const int8_t* b = reinterpret_cast<const int8_t*>(S.boards) + tid * kBoardBytes;
                const int fl = S.flag[tid];
                if (b[50 + fl] == 15) v = win_reward(b, fl);
            }
            values[row0 + tid] = v;
        }
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        __syncthreads();                                  // TMEM and the staging buffers are free again
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    }
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tmem), "r"((uint32_t)kHidden) : "memory");
}

__global__ void pack_w1_kernel(const float* __restrict__ w, uint16_t* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= kHidden * kKPad) return;
    int h = i / kKPad, k = i - h * kKPad;
    out[i] = k < BG_FEATURES ? __bfloat16_as_ushort(__float2bfloat16_rn(w[h * BG_FEATURES + k])) : (uint16_t)0;
}

}  // namespace bg

using namespace bg;

extern "C" int bg_pack_w1(const float* fc1_weight, uint16_t* w1_bf16, void* stream) {
    if (!fc1_weight || !w1_bf16) return bg_set_error_msg(BG_ERR_INVALID, "bg_pack_w1: null pointer");
    pack_w1_kernel<<<(kHidden * kKPad + 255) / 256, 256, 0, (cudaStream_t)stream>>>(fc1_weight, w1_bf16);
    return bg_set_error(cudaGetLastError(), "bg_pack_w1: launch");
}

extern "C" int bg_mlp_value(const int8_t* boards52, const int8_t* flags, int flag_all, int flip_flags, long long B,
                            const unsigned long long* n_rows_dev, const uint16_t* w1_bf16, const float* b1,
                            const float* wv, float bv, int terminal_aware, float* values, void* stream) {
    if (B < 0) return bg_set_error_msg(BG_ERR_INVALID, "bg_mlp_value: negative batch");
    if (B == 0) return BG_OK;
    if (!boards52 || !w1_bf16 || !b1 || !wv || !values) return bg_set_error_msg(BG_ERR_INVALID, "bg_mlp_value: null pointer");
    size_t smem = sizeof(MlpSmem) + 1024;
    cudaError_t e = cudaFuncSetAttribute(mlp_value_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return bg_set_error(e, "bg_mlp_value: cudaFuncSetAttribute");
    long long tiles = (B + kTileM - 1) / kTileM;
    long long grid = bg_sm_count();
    if (grid > tiles) grid = tiles;
    mlp_value_kernel<<<(unsigned)grid, kMlpThreads, smem, (cudaStream_t)stream>>>(
        boards52, flags, flag_all & 1, flip_flags & 1, B, n_rows_dev, w1_bf16, b1, wv, bv, terminal_aware, values);
    return bg_set_error(cudaGetLastError(), "bg_mlp_value: launch");
}
