// mlp.cu -- K4: the MLP leaf evaluator on the tcgen05 tensor cores, fused with the feature encoding.
//
// Replaces the value path of BackgammonPolicyNetwork.forward (src/agent/policy_network.py:58-75)
//     v = value_head( relu( fc1(x) ) ),   x = the 198 features of a position (K3's encoding)
// for B positions given as board52 + turn flag.  The 198-wide bf16 feature rows never touch HBM (nor shared
// memory): a CTA stages 128 boards, each producer thread expands its position's row in registers and writes it
// into TENSOR MEMORY (tcgen05.st, lane = position, two bf16 per 32-bit column), the MMA thread multiplies it by
// the W1 tile resident in shared memory with 13 tcgen05.mma (A from TMEM, B from smem; M128 N128 K16,
// bf16 x bf16 -> f32 in TMEM), and the epilogue reduces the hidden layer straight out of TMEM (tcgen05.ld):
// value = b_v + sum_h w_v[h] * relu(acc[h] + b1[h]).  HBM traffic: 53 B in, 4 B out per position.
// (With both operands in shared memory one K16 step reads 8 KB per 64 cycles -- the SM's whole shared-memory
// bandwidth -- which capped the first version at ~0.6 PFLOP/s; with A in TMEM shared memory only feeds W1.)
//
// W1 shared-memory operand layout (SWIZZLE_NONE "interleave", K-major): element (row r, column k) of the
// 128 x 208 bf16 tile lives at  (k/8)*2048 + r*16 + (k%8)*2  bytes: core matrices of 8 rows x 16 bytes
// are contiguous (128 B), 8-row groups advance by SBO = 128 B, 8-column chunks by LBO = 2048 B.
// TMEM columns: [0,256) two f32 accumulators of 128 columns; [256, 464) two A tiles of 104 columns.
#include <cuda_bf16.h>
#include "bg_device.cuh"
#include "bg_features.cuh"
#include "bg_tcgen05.cuh"
#include "bg_internal.h"

namespace bg {

constexpr int kOperandBytes = kChunks * kTileM * 16;   // 53,248
constexpr int kStages = 2;             // A tiles / TMEM accumulators in flight
constexpr int kAColsPerTile = kKPad / 2;                 // 104 TMEM columns per A tile (2 bf16 per column)
constexpr int kTmemACol0 = kStages * kHidden;            // first A column (after the accumulators)
constexpr int kTmemCols = 512;
// warp roles: 0-7 epilogue (warp w: TMEM lanes 32(w%4).., tiles of parity w/4), 8-23 A-tile producers (4 threads per
// position), 24 MMA issuer
constexpr int kEpiThreads = 256, kProdThreads = 512;
constexpr int kEpiWarps = kEpiThreads / 32, kProdWarps = kProdThreads / 32;
constexpr int kMlpThreads = kEpiThreads + kProdThreads + 32;

struct MlpSmem {
    uint8_t W[kOperandBytes];          // W1 tile, resident
    uint32_t boards[kStages][kTileM * kBoardWords];
    float b1[kHidden];
    float wv[kHidden];
    FeatureLut flut;
    unsigned long long a_full[kStages], a_empty[kStages], acc_full[kStages], acc_empty[kStages];
    uint32_t tmem_base;
};

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

// Persistent, warp-specialised, double-buffered:
//   producers (8 warps): boards of tile k+1 prefetched with cp.async; wait a_empty[s] -> expand the row in registers
//                        -> tcgen05.st into the TMEM A tile s -> arrive a_full[s]
//   MMA (1 thread)     : wait a_full[s], acc_empty[s] -> 13 x tcgen05.mma (A TMEM, B smem) into TMEM columns [128 s, 128 s + 128)
//                        -> tcgen05.commit to a_empty[s] and to acc_full[s]
//   epilogue (8 warps) : wait acc_full[s] -> tcgen05.ld 64 of the row's accumulators -> bias, ReLU, value head
//                        -> the two column halves are combined through shared memory -> store
//                        -> arrive acc_empty[s]
// terminal_aware: a row whose flag player has 15 men off gets the win reward instead of the network value
// (leaf rule of the 2-ply search).
template <bool BIAS>
__global__ void __launch_bounds__(kMlpThreads, 1) mlp_value_kernel(
    const int8_t* __restrict__ boards, const int8_t* __restrict__ flags, int flag_all, int flip_flags, long long B,
    const unsigned long long* __restrict__ row_begin_dev, const unsigned long long* __restrict__ n_rows_dev,
    const uint16_t* __restrict__ w1 /*[128][208] bf16*/,
    const float* __restrict__ b1, const float* __restrict__ wv, float bv, int terminal_aware,
    float* __restrict__ values) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    MlpSmem& S = *reinterpret_cast<MlpSmem*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (n_rows_dev) B = min(B, (long long)*n_rows_dev);
    const long long begin = row_begin_dev ? min(B, (long long)*row_begin_dev) : 0;      // rows [begin, B)

    // ---- one-time setup: W1 into the operand layout, biases, mbarriers, TMEM (2 x 128 columns)
    {   // w1 is stored in global memory in the operand layout (bg_pack_w1): a straight, coalesced, asynchronous copy
        const uint32_t w_s = smem_u32(S.W);
        for (int c = tid; c < kOperandBytes / 16; c += kMlpThreads) cp_async16_s(w_s + 16u * c, reinterpret_cast<const unsigned char*>(w1) + 16 * c);
        cp_async_commit();
        cp_async_wait_all();
    }
    load_feature_lut(&S.flut);
    if (tid < kHidden) { S.b1[tid] = BIAS ? b1[tid] : 0.0f; S.wv[tid] = wv[tid]; }
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&S.a_full[s], kProdThreads); mbar_init(&S.a_empty[s], 1);
            mbar_init(&S.acc_full[s], 1);          mbar_init(&S.acc_empty[s], kEpiThreads / 2);
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n"
                     :: "r"(smem_u32(&S.tmem_base)), "r"((uint32_t)kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");     // W tile visible to the tensor-core proxy
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = S.tmem_base;
    const long long n_tiles = (B - begin + kTileM - 1) / kTileM;

    if (warp >= kEpiWarps && warp < kEpiWarps + kProdWarps) {
        // ================= producers =================
        const int ptid = tid - kEpiThreads;                    // 0..255
        const int row = ptid & (kTileM - 1), quarter = ptid >> 7;  // four threads per position: chunks [7 q, min(7 q + 7, 26))
        // boards (and flags) of tile k+1 are prefetched with cp.async while tile k is being expanded
        const uint32_t boards_s[2] = {smem_u32(&S.boards[0][0]), smem_u32(&S.boards[1][0])};
        const unsigned char* bsrc = reinterpret_cast<const unsigned char*>(boards + begin * kBoardBytes);
        const bool base16 = (reinterpret_cast<uintptr_t>(bsrc) & 15u) == 0;   // tiles are 6,656 B apart: one alignment for all
        const long long nrows = B - begin;
        auto prefetch = [&](long long tile, int s) {
            const unsigned char* src = bsrc + tile * (kTileM * kBoardBytes);
            const long long left = nrows - tile * kTileM;
            if (base16 && left >= kTileM) {
                // a full, 16-byte aligned tile (every tile but the last when the rows start at a multiple of 4): 416 copies of 16 bytes
                if (ptid < kTileM * kBoardBytes / 16) cp_async16_s(boards_s[s] + 16u * ptid, src + 16 * ptid);
            } else {
                const int bytes = (int)min((long long)kTileM, left) * kBoardBytes;
                const int n16 = base16 ? bytes >> 4 : 0;
                for (int i = ptid; i < n16; i += kProdThreads) cp_async16_s(boards_s[s] + 16u * i, src + 16 * i);
                for (int i = 4 * n16 + ptid; i < (bytes >> 2); i += kProdThreads) cp_async4_s(boards_s[s] + 4u * i, src + 4 * i);
            }
            cp_async_commit();
        };
        if ((long long)blockIdx.x < n_tiles) prefetch(blockIdx.x, 0);
        int k = 0;
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++k) {
            const int s = k & 1;
            const uint32_t it = (uint32_t)(k >> 1);
            // the turn flag is part of chunk 24: only the last quarter's thread needs it (issued early, consumed late)
            int fl = flag_all;
            if (quarter == 3 && flags) { const long long r = begin + tile * kTileM + row; fl = r < B ? (int)__ldg(flags + r) : 0; }
            fl = (fl ^ flip_flags) & 1;
            cp_async_wait_all();                                // this thread's share of tile k's boards has landed
            asm volatile("bar.sync 1, %0;\n" :: "n"(kProdThreads) : "memory");   // ... and everybody else's; tile k-1 is fully built
            if (tile + gridDim.x < n_tiles) prefetch(tile + gridDim.x, s ^ 1);
            const uint32_t* srow = &S.boards[s][row * kBoardWords];              // stride 13 words: conflict-free
            mbar_wait(&S.a_empty[s], (it & 1u) ^ 1u);           // MMAs that read A[s] two tiles ago are done
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(kTmemACol0 + s * kAColsPerTile);
            if (quarter == 0)      build_row_chunks<0, 7>(srow, fl, &S.flut, trow);
            else if (quarter == 1) build_row_chunks<7, 14>(srow, fl, &S.flut, trow);
            else if (quarter == 2) build_row_chunks<14, 21>(srow, fl, &S.flut, trow);
            else                   build_row_chunks<21, 26>(srow, fl, &S.flut, trow);
            asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
            mbar_arrive(&S.a_full[s]);
        }
    } else if (warp == kEpiWarps + kProdWarps) {
        // ================= MMA issuer =================
        const uint32_t w_addr = smem_u32(S.W);
        int k = 0;
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++k) {
            const int s = k & 1;
            const uint32_t it = (uint32_t)(k >> 1);
            mbar_wait(&S.a_full[s], it & 1u);
            mbar_wait(&S.acc_empty[s], (it & 1u) ^ 1u);
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            if (lane == 0) {
#pragma unroll
                for (int ks = 0; ks < kKPad / 16; ++ks)
                    mma_bf16_ts(tmem + (uint32_t)(s * kHidden), tmem + (uint32_t)(kTmemACol0 + s * kAColsPerTile + ks * 8),
                                make_smem_desc(w_addr + ks * 2 * 2048), kIdesc, ks > 0 ? 1u : 0u);
                umma_commit(&S.a_empty[s]);                      // A[s] may be rebuilt
                umma_commit(&S.acc_full[s]);                     // accumulator s is complete
            }
            __syncwarp();
        }
    } else if (warp < kEpiWarps) {
        // ================= epilogue =================
        // warps 0-3 take the even tiles of this CTA (accumulator 0), warps 4-7 the odd ones (accumulator 1); a warp reads all 128
        // hidden units of its 32 rows in four passes of 32 columns: no partial sums to combine across warps, no CTA-level barrier,
        // and two tile times to finish a tile
        const int row = (warp & 3) * 32 + lane;                 // TMEM lane = position within the tile
        const int s = warp >> 2;                                // accumulator / parity of the tiles this warp handles
        int k = s;
        for (long long tile = blockIdx.x + (long long)s * gridDim.x; tile < n_tiles; tile += 2 * (long long)gridDim.x, k += 2) {
            const uint32_t it = (uint32_t)(k >> 1);
            const long long r = begin + tile * kTileM + row;
            mbar_wait(&S.acc_full[s], it & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            float v0 = 0.0f, v1 = 0.0f, v2 = 0.0f, v3 = 0.0f;
            uint32_t acc[32];
            const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(s * kHidden);
#pragma unroll
            for (int pass = 0; pass < 4; ++pass) {
                tmem_ld32(taddr + 32 * pass, acc);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 ww = *reinterpret_cast<const float4*>(&S.wv[32 * pass + j]);
                    if (BIAS) {
                        const float4 bb = *reinterpret_cast<const float4*>(&S.b1[32 * pass + j]);
                        v0 = fmaf(ww.x, fmaxf(__uint_as_float(acc[j + 0]) + bb.x, 0.0f), v0);
                        v1 = fmaf(ww.y, fmaxf(__uint_as_float(acc[j + 1]) + bb.y, 0.0f), v1);
                        v2 = fmaf(ww.z, fmaxf(__uint_as_float(acc[j + 2]) + bb.z, 0.0f), v2);
                        v3 = fmaf(ww.w, fmaxf(__uint_as_float(acc[j + 3]) + bb.w, 0.0f), v3);
                    } else {                                             // the bias came out of the GEMM (columns 198, 199)
                        v0 = fmaf(ww.x, fmaxf(__uint_as_float(acc[j + 0]), 0.0f), v0);
                        v1 = fmaf(ww.y, fmaxf(__uint_as_float(acc[j + 1]), 0.0f), v1);
                        v2 = fmaf(ww.z, fmaxf(__uint_as_float(acc[j + 2]), 0.0f), v2);
                        v3 = fmaf(ww.w, fmaxf(__uint_as_float(acc[j + 3]), 0.0f), v3);
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
            mbar_arrive(&S.acc_empty[s]);                        // accumulator s may be overwritten
            if (r < B) {
                float v = bv + ((v0 + v1) + (v2 + v3));
                if (terminal_aware) {
                    const int8_t* b = boards + r * kBoardBytes;
                    const int fl = ((flags ? flags[r] : flag_all) ^ flip_flags) & 1;
                    if (b[50 + fl] == 15) v = win_reward(b, fl);
                }
                values[r] = v;
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tmem), "r"((uint32_t)kTmemCols) : "memory");
}

// W1 in the kernels' internal K order (bg_tcgen05.cuh): points first, bar/off/flags in columns 192..197, and the bias
// as a bf16 hi / lo pair in columns 198 / 199 (the A tile has 1.0 there): b ~ hi + lo to 2^-17 relative
__global__ void pack_w1_kernel(const float* __restrict__ w, const float* __restrict__ b, uint16_t* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= kHidden * kKPad) return;
    int h = i / kKPad, k = i - h * kKPad;                    // hidden unit (operand row), internal K column
    const int src = bg_w1_source_column(k);
    uint16_t v = 0;
    if (src >= 0) v = __bfloat16_as_ushort(__float2bfloat16_rn(w[h * BG_FEATURES + src]));
    else if (b && k == BG_FEATURES) v = __bfloat16_as_ushort(__float2bfloat16_rn(b[h]));
    else if (b && k == BG_FEATURES + 1) v = __bfloat16_as_ushort(__float2bfloat16_rn(b[h] - __bfloat162float(__float2bfloat16_rn(b[h]))));
    out[(k >> 3) * (kHidden * 8) + h * 8 + (k & 7)] = v;      // tcgen05 K-major no-swizzle operand layout: (k/8)*2048 B + row*16 B + (k%8)*2 B
}

}  // namespace bg

using namespace bg;

extern "C" int bg_pack_w1(const float* fc1_weight, const float* fc1_bias, uint16_t* w1_bf16, void* stream) {
    if (!fc1_weight || !w1_bf16) return bg_set_error_msg(BG_ERR_INVALID, "bg_pack_w1: null pointer");
    pack_w1_kernel<<<(kHidden * kKPad + 255) / 256, 256, 0, (cudaStream_t)stream>>>(fc1_weight, fc1_bias, w1_bf16);
    return bg_set_error(cudaGetLastError(), "bg_pack_w1: launch");
}

int bg::mlp_value_launch(const int8_t* boards52, const int8_t* flags, int flag_all, int flip_flags, long long B,
                         const unsigned long long* row_begin_dev, const unsigned long long* n_rows_dev,
                         const uint16_t* w1_bf16, const float* b1, const float* wv, float bv, int terminal_aware,
                         float* values, cudaStream_t stream) {
    if (B < 0) return bg_set_error_msg(BG_ERR_INVALID, "bg_mlp_value: negative batch");
    if (B == 0) return BG_OK;
    if (!boards52 || !w1_bf16 || !wv || !values) return bg_set_error_msg(BG_ERR_INVALID, "bg_mlp_value: null pointer");
    size_t smem = sizeof(MlpSmem) + 1024;
    auto kern = b1 ? mlp_value_kernel<true> : mlp_value_kernel<false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return bg_set_error(e, "bg_mlp_value: cudaFuncSetAttribute");
    long long tiles = (B + kTileM - 1) / kTileM;
    long long grid = bg_sm_count();
    if (grid > tiles) grid = tiles;
    kern<<<(unsigned)grid, kMlpThreads, smem, stream>>>(
        boards52, flags, flag_all & 1, flip_flags & 1, B, row_begin_dev, n_rows_dev, w1_bf16, b1, wv, bv, terminal_aware, values);
    return bg_set_error(cudaGetLastError(), "bg_mlp_value: launch");
}

extern "C" int bg_mlp_value(const int8_t* boards52, const int8_t* flags, int flag_all, int flip_flags, long long B,
                            const unsigned long long* n_rows_dev, const uint16_t* w1_bf16, const float* b1,
                            const float* wv, float bv, int terminal_aware, float* values, void* stream) {
    return bg::mlp_value_launch(boards52, flags, flag_all, flip_flags, B, nullptr, n_rows_dev, w1_bf16, b1, wv, bv,
                                terminal_aware, values, (cudaStream_t)stream);
}
