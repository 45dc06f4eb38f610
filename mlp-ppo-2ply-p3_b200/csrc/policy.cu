// policy.cu -- the policy/value forward of the PPO rollout on the tcgen05 tensor cores, fused with the feature
// encoding, the action mask, the softmax and the sampling.
//
// Replaces, for B positions at once, BackgammonPPOAgent.select_action (src/agent/ppo_agent.py:138-191):
//     logits, value = BackgammonPolicyNetwork.forward(obs)        (src/agent/policy_network.py:58-75)
//     masked = logits + log(mask + 1e-45);  probs = softmax(masked)
//     action ~ Categorical(probs)  (training)   |   argmax(probs)  (inference);   log_prob(action)
// with obs = the 198 features of (board52, turn flag) and mask = the env's prefix mask (slot k legal iff
// k < legal_counts[b], src/environment/backgammon_env.py:228-231).
//
// Per tile of 128 positions (one CTA, 16 warps, everything on chip):
//   A  boards -> shared memory
//   B  8 warps expand the feature rows in registers and tcgen05.st them into TENSOR MEMORY (A1, as in K4)
//   C  13 x tcgen05.mma  acc1[128x128] = A1 (TMEM) x W1^T (smem)
//   D  16 warps read acc1 (tcgen05.ld), add b1, ReLU; value head partial sums; hidden activations rounded to bf16
//      and written back to TMEM as the A operand of the second GEMM (A2, 64 columns, aliasing A1)
//   E  4 chunks of 128 actions: 8 x tcgen05.mma  acc2[c&1] = A2 x Wa[128c..]^T, double buffered in TMEM; while
//      chunk c+1 multiplies, the 16 warps reduce chunk c straight out of TMEM: + bias, mask offset, online
//      max / sum-exp, and Gumbel-max sampling (argmax_i logit_i + g_i, g_i = -log(-log u_i), u_i from Philox4x32-10
//      keyed by (seed; global row, step, slot)) -- a single pass, the 500 logits of a row never leave the SM
//   F  the four column-quarter warps of a row are combined through shared memory -> action, log-prob, value
// Rows are processed in two classes so that the work follows the mask instead of the 500 slots: class A = rows with
// 1..128 legal slots (99 % of a self-play batch; mean 18.5): ONE chunk of the policy GEMM, and a warp whose 32 slots
// are illegal for all its rows skips its part of the epilogue altogether; illegal slots are left out of the softmax
// (the reference adds log(1e-45) = -103.3 to them: a relative change below 1e-38 of the normaliser).  Class B = rows
// with no legal slot (a pass: the reference samples among all 500 slots) or more than 128: all four chunks, mask
// offset applied literally.  policy_partition_kernel builds the two row lists (one array, A from the front, B from
// the back); tiles gather their rows through it and scatter their results.
// HBM traffic per position: 53 B board + 4 B count in, 12 B out.
// TMEM columns: [0,128) acc1; [128,232) A1 / [128,192) A2; [256,384) and [384,512) acc2 ping/pong.
#include <type_traits>
#include <cuda_bf16.h>
#include "bg_device.cuh"
#include "bg_features.cuh"
#include "bg_tcgen05.cuh"
#include "bg_internal.h"

namespace bg {

constexpr int kActions = BG_ACTIONS;       // 500
constexpr int kActPad = 512;
constexpr int kPolThreads = 512;
constexpr int kWaBytes = (kHidden / 8) * kActPad * 16;      // 131,072: (k/8)*8192 + n*16 + (k%8)*2
constexpr int kW1Bytes = kChunks * kTileM * 16;             // 53,248
constexpr float kMaskLog = -103.27893f;                     // log(1e-45f) in f32 (1e-45 is the smallest denormal)
constexpr uint32_t kTagGumbel = 0x47554D42u;                // "GUMB"
constexpr int kColAcc1 = 0, kColA = 128, kColAcc2 = 256;

struct PolSmem {
    uint8_t W1[kW1Bytes];
    uint8_t Wa[kWaBytes];
    uint32_t boards[2][kTileM * kBoardWords];   // double buffered: the next tile's rows are gathered during this tile's GEMMs
    float b1[kHidden], wv[kHidden], ba[kActPad];
    float part[4][kTileM];
    float red_m[4][kTileM], red_s[4][kTileM], red_g[4][kTileM], red_l[4][kTileM];
    int red_i[4][kTileM];
    FeatureLut flut;
    int rowidx[2][kTileM];                  // global row of each tile row (-1: none)
    unsigned long long bar1, bar2[2];
    uint32_t tmem_base;
    unsigned int tile_slot;                 // next tile index (dynamic schedule), broadcast by thread 0
};

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
    return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(a)) | ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(b)) << 16);
}

__global__ void __launch_bounds__(kPolThreads, 1) policy_kernel(
    const int8_t* __restrict__ boards, const int8_t* __restrict__ flags, int flag_all, long long B,
    const int32_t* __restrict__ legal_counts, const uint16_t* __restrict__ w1, const float* __restrict__ b1,
    const uint16_t* __restrict__ wa, const float* __restrict__ ba, const float* __restrict__ wv, float bv,
    unsigned long long seed, unsigned long long stream_base, uint32_t step, int greedy,
    const int32_t* __restrict__ row_list, const int32_t* __restrict__ row_list2, const unsigned int* __restrict__ class_ctr,
    unsigned int* __restrict__ tile_ctr,
    int32_t* __restrict__ actions, float* __restrict__ logp, float* __restrict__ values, float* __restrict__ logits_out) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    PolSmem& S = *reinterpret_cast<PolSmem*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // ---- one-time setup: weights into the tcgen05 operand layouts, biases, barriers, TMEM
    {   // both weight tiles are stored in global memory in their operand layouts (bg_pack_w1 / bg_pack_wa): straight,
        // coalesced, asynchronous copies (181 KB per CTA; element-wise gathers here were a fifth of the kernel's time)
        const uint32_t w1_s = smem_u32(S.W1), wa_s = smem_u32(S.Wa);
        for (int c = tid; c < kW1Bytes / 16; c += kPolThreads) cp_async16_s(w1_s + 16u * c, reinterpret_cast<const unsigned char*>(w1) + 16 * c);
        for (int c = tid; c < kWaBytes / 16; c += kPolThreads) cp_async16_s(wa_s + 16u * c, reinterpret_cast<const unsigned char*>(wa) + 16 * c);
        cp_async_commit();
        cp_async_wait_all();
    }
    load_feature_lut(&S.flut);
    if (tid < kHidden) { S.b1[tid] = b1 ? b1[tid] : 0.0f; S.wv[tid] = wv[tid]; }   // b1 == NULL: folded into W1 (bg_pack_w1)
    S.ba[tid] = tid < kActions ? ba[tid] : 0.0f;                 // kPolThreads == kActPad
    if (tid == 0) {
        mbar_init(&S.bar1, 1); mbar_init(&S.bar2[0], 1); mbar_init(&S.bar2[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n"
                     :: "r"(smem_u32(&S.tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = S.tmem_base;
    const uint32_t w1_addr = smem_u32(S.W1), wa_addr = smem_u32(S.Wa);
    // class A1 rows (1..32 legal slots) = row_list[0 .. nA1), class B rows = row_list[B-1 .. ] (from the back), class A2 rows
    // (33..128 legal slots) = row_list2[0 .. nA2); without lists every row is class B
    const long long nA1 = row_list ? (long long)class_ctr[0] : 0, nA2 = row_list ? (long long)class_ctr[3] : 0;
    const long long nB = B - nA1 - nA2;
    const long long tilesA1 = (nA1 + kTileM - 1) / kTileM, tilesA2 = (nA2 + kTileM - 1) / kTileM, tilesB = (nB + kTileM - 1) / kTileM;
    const long long n_tiles = tilesB + tilesA2 + tilesA1;
    const int q = warp & 3, cq = warp >> 2;                      // TMEM lane quadrant, column quarter
    const int row = q * 32 + lane;
    const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
    uint32_t ph1 = 0, ph2[2] = {0, 0};

    // rows of a tile: class, position in its class, number of rows, global row of tile row r
    // tile order: the (few, 5-6 x more expensive) class B tiles first, then class A; with a tile counter the CTAs pull
    // tiles dynamically so that a CTA holding a heavy tile takes fewer light ones
    auto tile_row = [&](long long tile, int r) -> int {
        if (tile < tilesB) {
            const long long i = tile * kTileM + r;
            if (i >= nB) return -1;
            return row_list ? row_list[B - 1 - i] : (int)i;
        }
        if (tile < tilesB + tilesA2) {
            const long long i = (tile - tilesB) * kTileM + r;
            return i < nA2 ? row_list2[i] : -1;
        }
        const long long i = (tile - tilesB - tilesA2) * kTileM + r;
        return i < nA1 ? row_list[i] : -1;
    };
    const uint32_t boards_s[2] = {smem_u32(&S.boards[0][0]), smem_u32(&S.boards[1][0])};
    // gather the boards of the rows listed in S.rowidx[b] into S.boards[b] (asynchronously)
    auto gather = [&](int b) {
        for (int i = tid; i < kTileM * kBoardWords; i += kPolThreads) {
            const int r = i / kBoardWords, wd = i - r * kBoardWords;
            const int g = S.rowidx[b][r];
            if (g >= 0) cp_async4_s(boards_s[b] + 4u * i, reinterpret_cast<const uint32_t*>(boards + (long long)g * kBoardBytes) + wd);
            else S.boards[b][i] = 0u;
        }
        cp_async_commit();
    };
    long long tile = blockIdx.x;
    if (tile_ctr) {
        if (tid == 0) S.tile_slot = atomicAdd(tile_ctr, 1u);
        __syncthreads();
        tile = S.tile_slot;
    }
    if (tile < n_tiles) {                                           // prologue: the first tile
        if (tid < kTileM) S.rowidx[0][tid] = tile_row(tile, tid);
        __syncthreads();
        gather(0);
    }
    for (int it = 0; tile < n_tiles; ++it) {
        const int bsel = it & 1;
        const bool class_a = tile >= tilesB;                      // one chunk of the policy GEMM
        const bool narrow = tile >= tilesB + tilesA2;             // ... and at most 32 legal slots: 8 slots per column warp
        const int n_chunks = class_a ? 1 : 4;
        // ---- A: this tile's boards have been gathered during the previous tile; the next tile is chosen and its row list read now
        if (tile_ctr && tid == 0) S.tile_slot = atomicAdd(tile_ctr, 1u);
        cp_async_wait_all();
        __syncthreads();
        const long long next_tile = tile_ctr ? (long long)S.tile_slot : tile + gridDim.x;
        int next_g = -1;
        if (tid < kTileM && next_tile < n_tiles) next_g = tile_row(next_tile, tid);
        // ---- B: feature rows -> TMEM (four threads per position: column-warp cq builds chunks 7 cq .. 7 cq + 6 of its rows)
        {
            const uint32_t* srow = &S.boards[bsel][row * kBoardWords];
            int fl = 0;
            if (cq == 3) { const int pg = S.rowidx[bsel][row]; fl = pg >= 0 ? (int)((flags ? flags[pg] : flag_all) & 1) : 0; }   // chunk 24 holds the turn flags
            const uint32_t trow = lane_base + (uint32_t)kColA;
            if (cq == 0)      build_row_chunks<0, 7>(srow, fl, &S.flut, trow);
            else if (cq == 1) build_row_chunks<7, 14>(srow, fl, &S.flut, trow);
            else if (cq == 2) build_row_chunks<14, 21>(srow, fl, &S.flut, trow);
            else              build_row_chunks<21, 26>(srow, fl, &S.flut, trow);
            asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
        }
        if (tid < kTileM) S.rowidx[bsel ^ 1][tid] = next_g;
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        __syncthreads();
        if (next_tile < n_tiles) gather(bsel ^ 1);                 // lands while the GEMMs and the epilogues of this tile run
        // ---- C: hidden layer GEMM
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll
            for (int ks = 0; ks < kKPad / 16; ++ks)
                mma_bf16_ts(tmem + kColAcc1, tmem + (uint32_t)(kColA + ks * 8), make_smem_desc(w1_addr + ks * 2 * 2048), kIdesc,
                            ks > 0 ? 1u : 0u);
            umma_commit(&S.bar1);
        }
        mbar_wait(&S.bar1, ph1); ph1 ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        // ---- D: bias + ReLU, value-head partials, hidden -> bf16 -> TMEM (A operand of the policy GEMM)
        {
            uint32_t acc[32];
            tmem_ld32(lane_base + (uint32_t)(kColAcc1 + 32 * cq), acc);
            tmem_ld_wait();
            float v = 0.0f;
            uint32_t hw[16];
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
                const float h0 = fmaxf(__uint_as_float(acc[j]) + S.b1[32 * cq + j], 0.0f);
                const float h1 = fmaxf(__uint_as_float(acc[j + 1]) + S.b1[32 * cq + j + 1], 0.0f);
                v = fmaf(S.wv[32 * cq + j], h0, v);
                v = fmaf(S.wv[32 * cq + j + 1], h1, v);
                hw[j >> 1] = pack_bf16x2(h0, h1);
            }
            S.part[cq][row] = v;
#pragma unroll
            for (int i = 0; i < 4; ++i)
                tmem_st4(lane_base + (uint32_t)(kColA + 16 * cq + 4 * i), make_uint4(hw[4 * i], hw[4 * i + 1], hw[4 * i + 2], hw[4 * i + 3]));
            asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        __syncthreads();
        // ---- E: policy GEMM in 4 chunks of 128 actions + fused masked softmax / Gumbel-max
        auto issue_chunk = [&](int c) {
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll
            for (int ks = 0; ks < kHidden / 16; ++ks)
                mma_bf16_ts(tmem + (uint32_t)(kColAcc2 + 128 * (c & 1)), tmem + (uint32_t)(kColA + ks * 8),
                            make_smem_desc_kmajor(wa_addr + c * (128 * 16) + ks * 2 * (kActPad * 16), kActPad * 16, 128), kIdesc,
                            ks > 0 ? 1u : 0u);
            umma_commit(&S.bar2[c & 1]);
        };
        if (tid == 0) { issue_chunk(0); if (n_chunks > 1) issue_chunk(1); }
        const long long gid = S.rowidx[bsel][row];                      // row index within this call (-1: none)
        const int n_legal = (gid >= 0 && legal_counts) ? legal_counts[gid] : (legal_counts ? 1 : kActions);
        const unsigned long long sid = stream_base + (unsigned long long)(gid >= 0 ? gid : 0);   // global stream id (game id)
        // class A: this warp's slots of chunk 0 (32 of them, or 8 in a narrow tile) matter only if some row of the warp has that many
        const bool warp_active = !class_a || __any_sync(kFull, n_legal > (narrow ? 8 : 32) * cq);
        float m = -INFINITY, s = 0.0f, gbest = -INFINITY, lbest = 0.0f;
        int ibest = 0;
        // W accumulator columns acc[0..W) = action slots base .. base + W - 1 of this thread's row
        auto reduce_block = [&](const uint32_t* acc, int base, auto wtag) {
            constexpr int W = decltype(wtag)::value;
            float x[W];
            float cm = -INFINITY;
#pragma unroll
            for (int j = 0; j < W; ++j) {
                const int i = base + j;
                float l = __uint_as_float(acc[j]) + S.ba[i];
                if (logits_out && gid >= 0 && i < kActions) logits_out[gid * kActions + i] = l;
                if (i >= n_legal) l = class_a ? -INFINITY : l + kMaskLog;   // logits + log(mask + 1e-45), ppo_agent.py:165
                if (i >= kActions) l = -INFINITY;                  // padding slots do not exist
                x[j] = l;
                cm = fmaxf(cm, l);
            }
            if (cm > -INFINITY) {
                const float mn = fmaxf(m, cm);
                float add = 0.0f;
#pragma unroll
                for (int j = 0; j < W; ++j) add += __expf(x[j] - mn);
                s = s * __expf(m - mn) + add;
                m = mn;
            }
            // sampling: argmax_i x_i + Gumbel_i (lowest slot on ties); a masked slot (offset -103) can only win when
            // every slot is masked (a pass: the reference samples from all 500 then), so its noise is skipped otherwise
            if (greedy) {
#pragma unroll
                for (int j = 0; j < W; ++j)
                    if (x[j] > gbest) { gbest = x[j]; ibest = base + j; lbest = x[j]; }
            } else {
                const int lim = n_legal > 0 ? min(n_legal, kActions) : kActions;
#pragma unroll
                for (int g4 = 0; g4 < W / 4; ++g4) {
                    if (base + 4 * g4 < lim) {
                        uint32_t r[4];
                        philox4x32_10((uint32_t)sid, ((uint32_t)(sid >> 32) << 8) | (uint32_t)((base >> 2) + g4), step, kTagGumbel,
                                      (uint32_t)seed, (uint32_t)(seed >> 32), r);
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int j = 4 * g4 + e;
                            const float u = ((float)(r[e] >> 8) + 0.5f) * (1.0f / 16777216.0f);      // (0,1), 24 bits
                            const float key = x[j] - __logf(-__logf(u));
                            if (base + j < lim && key > gbest) { gbest = key; ibest = base + j; lbest = x[j]; }
                        }
                    }
                }
            }
        };
#pragma unroll 1
        for (int c = 0; c < n_chunks; ++c) {
            mbar_wait(&S.bar2[c & 1], ph2[c & 1]); ph2[c & 1] ^= 1u;
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            if (warp_active) {
                if (narrow) {                                      // slots 8 cq .. 8 cq + 7: the 32 slots of a narrow tile spread over all four column warps
                    uint32_t acc[8];
                    tmem_ld8(lane_base + (uint32_t)(kColAcc2 + 8 * cq), acc);
                    tmem_ld_wait();
                    reduce_block(acc, 8 * cq, std::integral_constant<int, 8>{});
                } else {
                    uint32_t acc[32];
                    tmem_ld32(lane_base + (uint32_t)(kColAcc2 + 128 * (c & 1) + 32 * cq), acc);
                    tmem_ld_wait();
                    reduce_block(acc, 128 * c + 32 * cq, std::integral_constant<int, 32>{});
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
            __syncthreads();                                       // every warp has read acc2[c & 1]
            if (tid == 0 && c + 2 < n_chunks) issue_chunk(c + 2);
        }
        // ---- F: combine the four column quarters of each row
        S.red_m[cq][row] = m; S.red_s[cq][row] = s; S.red_g[cq][row] = gbest; S.red_l[cq][row] = lbest; S.red_i[cq][row] = ibest;
        __syncthreads();
        if (tid < kTileM && S.rowidx[bsel][tid] >= 0) {
            const int r = tid;
            float M = S.red_m[0][r];
#pragma unroll
            for (int k = 1; k < 4; ++k) M = fmaxf(M, S.red_m[k][r]);
            float sum = 0.0f, gb = -INFINITY, lb = 0.0f;
            int ib = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (S.red_m[k][r] > -INFINITY) sum += S.red_s[k][r] * __expf(S.red_m[k][r] - M);
                const float g = S.red_g[k][r];
                if (g > gb || (g == gb && g > -INFINITY && S.red_i[k][r] < ib)) { gb = g; lb = S.red_l[k][r]; ib = S.red_i[k][r]; }
            }
            const long long g = S.rowidx[bsel][r];
            actions[g] = ib;
            if (logp) logp[g] = lb - (M + __logf(sum));
            if (values) values[g] = bv + ((S.part[0][r] + S.part[1][r]) + (S.part[2][r] + S.part[3][r]));
        }
        __syncthreads();
        tile = next_tile;
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tmem), "r"(512u) : "memory");
}

// Row lists of the three classes: rows with 1..32 legal slots from the front of row_list, rows with none or more than 128
// from its back, rows with 33..128 from the front of row_list2.  Order inside a class is arbitrary (atomics); results do not
// depend on it (the random stream is keyed by the row).  ctr: [0] class A1 rows, [1] class B rows, [3] class A2 rows.
__global__ void __launch_bounds__(256) policy_partition_kernel(const int32_t* __restrict__ counts, long long B,
                                                               int32_t* __restrict__ row_list, int32_t* __restrict__ row_list2,
                                                               unsigned int* __restrict__ ctr) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const int n = g < B ? counts[g] : -1;
    const bool a1 = g < B && n >= 1 && n <= 32, a2 = g < B && n > 32 && n <= 128, b = g < B && !a1 && !a2;
    const unsigned m1 = __ballot_sync(kFull, a1), m2 = __ballot_sync(kFull, a2), mb = __ballot_sync(kFull, b);
    unsigned int base1 = 0, base2 = 0, base_b = 0;
    if (lane == 0) {
        if (m1) base1 = atomicAdd(&ctr[0], (unsigned)__popc(m1));
        if (m2) base2 = atomicAdd(&ctr[3], (unsigned)__popc(m2));
        if (mb) base_b = atomicAdd(&ctr[1], (unsigned)__popc(mb));
    }
    base1 = __shfl_sync(kFull, base1, 0); base2 = __shfl_sync(kFull, base2, 0); base_b = __shfl_sync(kFull, base_b, 0);
    const unsigned below = (1u << lane) - 1u;
    if (a1) row_list[base1 + __popc(m1 & below)] = (int32_t)g;
    if (a2) row_list2[base2 + __popc(m2 & below)] = (int32_t)g;
    if (b) row_list[B - 1 - (long long)(base_b + __popc(mb & below))] = (int32_t)g;
}

__global__ void pack_wa_kernel(const float* __restrict__ w, uint16_t* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= kActPad * kHidden) return;
    const int n = i / kHidden, k = i - n * kHidden;            // action slot (operand row), hidden unit (K)
    // tcgen05 K-major no-swizzle operand layout: (k/8)*8192 B + row*16 B + (k%8)*2 B
    out[(k >> 3) * (kActPad * 8) + n * 8 + (k & 7)] = n < kActions ? __bfloat16_as_ushort(__float2bfloat16_rn(w[i])) : (uint16_t)0;
}

}  // namespace bg

using namespace bg;

extern "C" int bg_pack_wa(const float* action_head_weight, uint16_t* wa_bf16, void* stream) {
    if (!action_head_weight || !wa_bf16) return bg_set_error_msg(BG_ERR_INVALID, "bg_pack_wa: null pointer");
    pack_wa_kernel<<<(kActPad * kHidden + 255) / 256, 256, 0, (cudaStream_t)stream>>>(action_head_weight, wa_bf16);
    return bg_set_error(cudaGetLastError(), "bg_pack_wa: launch");
}

extern "C" size_t bg_policy_workspace_bytes(long long B) { return 16 + 2 * sizeof(int32_t) * (size_t)(B > 0 ? B : 1); }

extern "C" int bg_policy_sample(const int8_t* boards52, const int8_t* flags, int flag_all, long long B,
                                const int32_t* legal_counts, const uint16_t* w1_bf16, const float* b1,
                                const uint16_t* wa_bf16, const float* ba, const float* wv, float bv,
                                unsigned long long seed, unsigned long long stream_base, uint32_t step, int greedy,
                                int32_t* actions, float* log_probs, float* values, float* logits_out, void* workspace,
                                size_t workspace_bytes, void* stream) {
    if (B < 0) return bg_set_error_msg(BG_ERR_INVALID, "bg_policy_sample: negative batch");
    if (B == 0) return BG_OK;
    if (B > 0x7FFFFFF0LL) return bg_set_error_msg(BG_ERR_INVALID, "bg_policy_sample: batch too large");
    if (!boards52 || !w1_bf16 || !wa_bf16 || !ba || !wv || !actions)
        return bg_set_error_msg(BG_ERR_INVALID, "bg_policy_sample: null pointer");
    const size_t smem = sizeof(PolSmem) + 1024;
    cudaError_t e = cudaFuncSetAttribute(policy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return bg_set_error(e, "bg_policy_sample: cudaFuncSetAttribute");
    // with a workspace: dynamic tile schedule, and -- when there is a mask and the logits are not wanted -- the two row
    // classes (1..128 legal slots / the rest)
    const int32_t* row_list = nullptr;
    const int32_t* row_list2 = nullptr;
    const unsigned int* class_ctr = nullptr;
    unsigned int* tile_ctr = nullptr;
    if (workspace) {
        if (workspace_bytes < bg_policy_workspace_bytes(B)) return bg_set_error_msg(BG_ERR_INVALID, "bg_policy_sample: workspace too small");
        unsigned int* ctr = static_cast<unsigned int*>(workspace);      // [0] class A1 rows, [1] class B rows, [2] next tile, [3] class A2 rows
        e = cudaMemsetAsync(ctr, 0, 16, (cudaStream_t)stream);
        if (e != cudaSuccess) return bg_set_error(e, "bg_policy_sample: memset");
        tile_ctr = ctr + 2;
        if (legal_counts && !logits_out) {
            int32_t* list = reinterpret_cast<int32_t*>(static_cast<unsigned char*>(workspace) + 16);
            policy_partition_kernel<<<(unsigned)((B + 255) / 256), 256, 0, (cudaStream_t)stream>>>(legal_counts, B, list, list + B, ctr);
            row_list = list; row_list2 = list + B; class_ctr = ctr;
        }
    }
    long long tiles = (B + kTileM - 1) / kTileM + 2;
    long long grid = bg_sm_count();
    if (grid > tiles) grid = tiles;
    policy_kernel<<<(unsigned)grid, kPolThreads, smem, (cudaStream_t)stream>>>(
        boards52, flags, flag_all & 1, B, legal_counts, w1_bf16, b1, wa_bf16, ba, wv, bv, seed, stream_base, step, greedy,
        row_list, row_list2, class_ctr, tile_ctr, actions, log_probs, values, logits_out);
    return bg_set_error(cudaGetLastError(), "bg_policy_sample: launch");
}
