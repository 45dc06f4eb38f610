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
// Per tile of 128 positions (one CTA of 8 warps, TWO CTAs per SM, everything on chip):
//   A  boards -> shared memory
//   B  8 warps expand the feature rows in registers and tcgen05.st them into TENSOR MEMORY (A1, as in K4)
//   C  13 x tcgen05.mma  acc[128x128] = A1 (TMEM) x W1^T (smem)
//   D  8 warps read acc (tcgen05.ld), add b1, ReLU; value head partial sums; hidden activations rounded to bf16
//      and written back to TMEM as the A operand of the second GEMM (A2, 64 columns, aliasing A1)
//   E  per chunk of 128 actions: 8 x tcgen05.mma  acc = A2 x Wa[128c..]^T (the accumulator of C is dead by now and is
//      reused), then the 8 warps reduce the chunk straight out of TMEM: + bias, mask offset, online max / sum-exp, and
//      categorical sampling (inverse CDF inside a block of 8 / 16 slots, blocks merged reservoir-style: two uniforms per block from
//      Philox4x32-10 keyed by (seed; global row, step, block) instead of one Gumbel variate per slot) -- the 500 logits of a row
//      never leave the SM
//   F  the two column halves of a row are combined through shared memory -> action, log-prob, value
// Rows are processed in two classes so that the work follows the mask instead of the 500 slots: class A = rows with
// 1..128 legal slots (99 % of a self-play batch; mean 18.5): ONE chunk of the policy GEMM, and a warp skips the blocks of 16
// (8 in a narrow tile) slots that are illegal for all its rows; illegal slots are left out of the softmax
// (the reference adds log(1e-45) = -103.3 to them: a relative change below 1e-38 of the normaliser).  Class B = rows
// with no legal slot (a pass: the reference samples among all 500 slots) or more than 128: all four chunks, mask
// offset applied literally.  policy_partition_kernel builds the two row lists (one array, A from the front, B from
// the back); tiles gather their rows through it and scatter their results.
// Two CTAs per SM instead of a hand-rolled pipeline: the phases of a tile are serial (build -> GEMM -> activation -> GEMM ->
// softmax / sampling), and a second, independent tile on the same SM fills the issue slots the first leaves idle while it waits
// for its MMAs, its gather or a barrier (round 1: one CTA of 16 warps per SM, IPC 1.18, 40 % of the issue slots).  What
// makes two CTAs fit: only ONE 32 KB chunk of the action head (slots 128 c .. 128 c + 127) is resident in shared memory --
// chunk 0 for the class A tiles; the few class B tiles stream chunks 1..3 in and put chunk 0 back -- so a CTA needs 110 KB
// and 256 TMEM columns ([0,128) accumulator of both GEMMs; [128,232) A1 / [128,192) A2).
// HBM traffic per position: 53 B board + 4 B count in, 12 B out.
#include <type_traits>
#include <cuda_bf16.h>
#include "bg_device.cuh"
#include "bg_features.cuh"
#include "bg_tcgen05.cuh"
#include "bg_internal.h"

namespace bg {

constexpr int kActions = BG_ACTIONS;       // 500
constexpr int kActPad = 512;
constexpr int kPolThreads = 256;
constexpr int kWaBytes = (kHidden / 8) * kActPad * 16;      // 131,072 in global memory: (k/8)*8192 + n*16 + (k%8)*2
constexpr int kWaChunkBytes = (kHidden / 8) * 128 * 16;     // 32,768 in shared memory: (k/8)*2048 + (n%128)*16 + (k%8)*2
constexpr int kW1Bytes = kChunks * kTileM * 16;             // 53,248
constexpr float kMaskLog = -103.27893f;                     // log(1e-45f) in f32 (1e-45 is the smallest denormal)
constexpr uint32_t kTagGumbel = 0x47554D42u;                // "GUMB"
constexpr int kColAcc = 0, kColA = 128;
constexpr uint32_t kPolTmemCols = 256;
constexpr int kMaxSplitB = 64;                              // class B tiles whose four chunks are dealt to four CTAs (more: serial chunks)
struct BPartial { float m, s, g, l; int i; };               // one chunk's share of a row: max, sum exp, best key, its logit, its slot

struct PolSmem {
    uint8_t W1[kW1Bytes];
    uint8_t Wa[kWaChunkBytes];
    uint32_t boards[2][kTileM * kBoardWords];   // double buffered: the next tile's rows are gathered during this tile's GEMMs
    float b1[kHidden], wv[kHidden], ba[kActPad];
    float part[2][kTileM];
    float red_m[2][kTileM], red_s[2][kTileM], red_g[2][kTileM], red_l[2][kTileM];
    int red_i[2][kTileM];
    FeatureLut flut;
    int rowidx[2][kTileM];                  // global row of each tile row (-1: none)
    unsigned long long bar1, bar2;
    uint32_t tmem_base;
    unsigned int tile_slot;                 // next tile index (dynamic schedule), broadcast by thread 0
    unsigned int tile_slot2;                // arrival number of this chunk of a class B tile
};
static_assert(2 * (sizeof(PolSmem) + 1024) <= 227 * 1024, "two CTAs per SM");

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
    return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(a)) | ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(b)) << 16);
}

__global__ void __launch_bounds__(kPolThreads, 2) policy_kernel(
    const int8_t* __restrict__ boards, const int8_t* __restrict__ flags, int flag_all, long long B,
    const int32_t* __restrict__ legal_counts, const uint16_t* __restrict__ w1, const float* __restrict__ b1,
    const uint16_t* __restrict__ wa, const float* __restrict__ ba, const float* __restrict__ wv, float bv,
    unsigned long long seed, unsigned long long stream_base, uint32_t step, int greedy,
    const int32_t* __restrict__ row_list, const int32_t* __restrict__ row_list2, const unsigned int* __restrict__ class_ctr,
    unsigned int* __restrict__ tile_ctr, unsigned int* __restrict__ b_done, BPartial* __restrict__ b_part,
    int32_t* __restrict__ actions, float* __restrict__ logp, float* __restrict__ values, float* __restrict__ logits_out) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    PolSmem& S = *reinterpret_cast<PolSmem*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t w1_addr = smem_u32(S.W1), wa_addr = smem_u32(S.Wa);

    // chunk c of the action head (slots 128 c .. 128 c + 127): 16 pieces of 2 KB (one per group of 8 hidden units), asynchronously
    auto stage_wa = [&](int c) {
        const unsigned char* src = reinterpret_cast<const unsigned char*>(wa) + c * 2048;
        for (int i = tid; i < kWaChunkBytes / 16; i += kPolThreads)
            cp_async16_s(wa_addr + 16u * i, src + (i >> 7) * (kActPad * 16) + (i & 127) * 16);
        cp_async_commit();
    };
    // ---- one-time setup: weights into the tcgen05 operand layouts, biases, barriers, TMEM
    {   // both weight tiles are stored in global memory in their operand layouts (bg_pack_w1 / bg_pack_wa): straight,
        // coalesced, asynchronous copies (85 KB per CTA)
        for (int c = tid; c < kW1Bytes / 16; c += kPolThreads) cp_async16_s(w1_addr + 16u * c, reinterpret_cast<const unsigned char*>(w1) + 16 * c);
        stage_wa(0);
        cp_async_wait_all();
    }
    int wa_chunk = 0;                                            // the chunk of the action head that is in shared memory (uniform over the CTA)
    load_feature_lut(&S.flut);
    if (tid < kHidden) { S.b1[tid] = b1 ? b1[tid] : 0.0f; S.wv[tid] = wv[tid]; }   // b1 == NULL: folded into W1 (bg_pack_w1)
    for (int i = tid; i < kActPad; i += kPolThreads) S.ba[i] = i < kActions ? ba[i] : 0.0f;
    if (tid == 0) {
        mbar_init(&S.bar1, 1); mbar_init(&S.bar2, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n"
                     :: "r"(smem_u32(&S.tmem_base)), "r"(kPolTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = S.tmem_base;
    // class A1 rows (1..32 legal slots) = row_list[0 .. nA1), class B rows = row_list[B-1 .. ] (from the back), class A2 rows
    // (33..128 legal slots) = row_list2[0 .. nA2); without lists every row is class B
    const long long nA1 = row_list ? (long long)class_ctr[0] : 0, nA2 = row_list ? (long long)class_ctr[3] : 0;
    const long long nB = B - nA1 - nA2;
    const long long tilesA1 = (nA1 + kTileM - 1) / kTileM, tilesA2 = (nA2 + kTileM - 1) / kTileM, tilesB = (nB + kTileM - 1) / kTileM;
    // Work items.  A class B tile costs four chunks of the action head, one after the other, with the chunks streamed through the
    // one buffer -- 25 us on the critical path of a 30 us kernel.  So (when there are row lists and few such tiles) its chunks are
    // four ITEMS, taken by four CTAs: each recomputes the hidden layer (cheap), reduces its 128 slots and leaves a partial per row
    // in b_part; the CTA that finishes last (b_done) combines the four and writes the row's results.
    const bool split_b = row_list && b_part && tilesB <= kMaxSplitB;
    const long long itemsB = split_b ? 4 * tilesB : tilesB;
    const long long n_tiles = itemsB + tilesA2 + tilesA1;         // (items)
    auto tile_of = [&](long long item) -> long long { return item < itemsB ? (split_b ? item >> 2 : item) : item - itemsB + tilesB; };
    const int q = warp & 3, cq = warp >> 2;                      // TMEM lane quadrant, column half
    const int row = q * 32 + lane;
    const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
    uint32_t ph1 = 0, ph2 = 0;

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
        if (tid < kTileM) S.rowidx[0][tid] = tile_row(tile_of(tile), tid);
        __syncthreads();
        gather(0);
    }
    for (int it = 0; tile < n_tiles; ++it) {
        const int bsel = it & 1;
        const bool class_a = tile >= itemsB;                      // one chunk of the policy GEMM
        const bool narrow = tile >= itemsB + tilesA2;             // ... and at most 32 legal slots: 16 slots per column warp
        const bool b_item = !class_a && split_b;                  // one chunk of a class B tile
        const int c_first = b_item ? (int)(tile & 3) : 0, n_chunks = (class_a || b_item) ? 1 : 4;
        // ---- A: this tile's boards have been gathered during the previous tile; the next tile is chosen and its row list read now
        // (a class B item is 2-5 x longer than a class A tile: its CTA chooses its next tile only when it is done, so that it does
        // not sit on a tile other CTAs could have finished meanwhile)
        const bool claim_early = class_a || !tile_ctr;
        if (claim_early && tile_ctr && tid == 0) S.tile_slot = atomicAdd(tile_ctr, 1u);
        cp_async_wait_all();
        __syncthreads();
        long long next_tile = claim_early ? (tile_ctr ? (long long)S.tile_slot : tile + gridDim.x) : n_tiles;
        int next_g = -1;
        if (tid < kTileM && next_tile < n_tiles) next_g = tile_row(tile_of(next_tile), tid);
        // ---- B: feature rows -> TMEM (two threads per position: column-warp cq builds chunks 13 cq .. 13 cq + 12 of its rows)
        {
            const uint32_t* srow = &S.boards[bsel][row * kBoardWords];
            int fl = 0;
            if (cq == 1) { const int pg = S.rowidx[bsel][row]; fl = pg >= 0 ? (int)((flags ? flags[pg] : flag_all) & 1) : 0; }   // chunk 24 holds the turn flags
            const uint32_t trow = lane_base + (uint32_t)kColA;
            if (cq == 0) build_row_chunks<0, 13>(srow, fl, &S.flut, trow);
            else         build_row_chunks<13, 26>(srow, fl, &S.flut, trow);
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
                mma_bf16_ts(tmem + kColAcc, tmem + (uint32_t)(kColA + ks * 8), make_smem_desc(w1_addr + ks * 2 * 2048), kIdesc,
                            ks > 0 ? 1u : 0u);
            umma_commit(&S.bar1);
        }
        mbar_wait(&S.bar1, ph1); ph1 ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        // ---- D: bias + ReLU, value-head partials, hidden -> bf16 -> TMEM (A operand of the policy GEMM); 64 hidden units per thread
        {
            float v = 0.0f;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int c0 = 64 * cq + 32 * h;
                uint32_t acc[32];
                tmem_ld32(lane_base + (uint32_t)(kColAcc + c0), acc);
                tmem_ld_wait();
                uint32_t hw[16];
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                    const float h0 = fmaxf(__uint_as_float(acc[j]) + S.b1[c0 + j], 0.0f);
                    const float h1 = fmaxf(__uint_as_float(acc[j + 1]) + S.b1[c0 + j + 1], 0.0f);
                    v = fmaf(S.wv[c0 + j], h0, v);
                    v = fmaf(S.wv[c0 + j + 1], h1, v);
                    hw[j >> 1] = pack_bf16x2(h0, h1);
                }
                // (the A1 columns these stores overwrite were read by GEMM 1, which has completed: bar1)
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    tmem_st4(lane_base + (uint32_t)(kColA + (c0 >> 1) + 4 * i), make_uint4(hw[4 * i], hw[4 * i + 1], hw[4 * i + 2], hw[4 * i + 3]));
            }
            S.part[cq][row] = v;
            asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        __syncthreads();                                           // every warp has read the accumulator of GEMM 1 and written its part of A2
        // ---- E: policy GEMM per chunk of 128 actions + fused masked softmax / sampling
        auto issue_chunk = [&]() {
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll
            for (int ks = 0; ks < kHidden / 16; ++ks)
                mma_bf16_ts(tmem + (uint32_t)kColAcc, tmem + (uint32_t)(kColA + ks * 8),
                            make_smem_desc_kmajor(wa_addr + ks * 2 * 2048, 2048, 128), kIdesc, ks > 0 ? 1u : 0u);
            umma_commit(&S.bar2);
        };
        const long long gid = S.rowidx[bsel][row];                      // row index within this call (-1: none)
        const int n_legal = (gid >= 0 && legal_counts) ? legal_counts[gid] : (legal_counts ? 1 : kActions);
        const unsigned long long sid = stream_base + (unsigned long long)(gid >= 0 ? gid : 0);   // global stream id (game id)
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
                const float before = s * __expf(m - mn);            // the slots seen so far, on the new scale
                s = before + add;
                m = mn;
                if (!greedy) {
                    // Categorical sampling, hierarchically and exactly, with TWO uniforms per block instead of one per slot: a slot of
                    // this block by inverse CDF over its exp(x - m) (u0), then the block's choice replaces the running choice with
                    // probability (block mass) / (mass so far) (u1) -- a reservoir over blocks.  The blocks of a row (its column
                    // halves, the chunks of a class B row) are merged the same way in stage F.  Masked slots have mass 0 (class A)
                    // or 1e-45 of it (class B, as in the reference).  Uniforms: Philox4x32-10 keyed by (seed; row, step, block).
                    uint32_t r[4];
                    philox4x32_10((uint32_t)sid, ((uint32_t)(sid >> 32) << 8) | (uint32_t)(base >> 3), step, kTagGumbel,
                                  (uint32_t)seed, (uint32_t)(seed >> 32), r);
                    const float u0 = ((float)(r[0] >> 8) + 0.5f) * (1.0f / 16777216.0f);          // (0,1), 24 bits
                    const float u1 = ((float)(r[1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
                    const float target = u0 * add;
                    float cum = 0.0f, lsel = 0.0f;
                    int jsel = -1;
                    bool found = false;
#pragma unroll
                    for (int j = 0; j < W; ++j) {
                        const float e = __expf(x[j] - mn);
                        cum += e;
                        if (!found && e > 0.0f) {                  // the latest slot with mass (the fallback if rounding keeps cum below the target) ...
                            jsel = j; lsel = x[j];
                            found = cum >= target;                 // ... frozen once the cumulative mass reaches the target
                        }
                    }
                    if (jsel >= 0 && u1 * s >= before) { ibest = base + jsel; lbest = lsel; }
                }
            }
            // greedy: running argmax (lowest slot on ties)
            if (greedy) {
#pragma unroll
                for (int j = 0; j < W; ++j)
                    if (x[j] > gbest) { gbest = x[j]; ibest = base + j; lbest = x[j]; }
            }
        };
#pragma unroll 1
        for (int c = c_first; c < c_first + n_chunks; ++c) {
            if (wa_chunk != c) {                                   // class B tiles (and the first class A tile after them): another chunk of the action head
                stage_wa(c);                                       // (the MMAs that read the previous chunk have completed: bar2)
                cp_async_wait_all();
                asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
                __syncthreads();
                wa_chunk = c;
            }
            if (tid == 0) issue_chunk();
            mbar_wait(&S.bar2, ph2); ph2 ^= 1u;
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            // a block of 16 (narrow: 8) slots matters to a warp only if some row of the warp has that many legal slots
            if (narrow) {                                          // slots 16 cq .. 16 cq + 15: the 32 slots of a narrow tile spread over both column warps
#pragma unroll 1
                for (int h = 0; h < 2; ++h) {
                    const int base = 16 * cq + 8 * h;
                    if (!__any_sync(kFull, n_legal > base)) break;
                    uint32_t acc[8];
                    tmem_ld8(lane_base + (uint32_t)(kColAcc + base), acc);
                    tmem_ld_wait();
                    reduce_block(acc, base, std::integral_constant<int, 8>{});
                }
            } else {
#pragma unroll 1
                for (int h = 0; h < 4; ++h) {                      // blocks of 16 slots (32 at once do not fit the registers of two CTAs per SM)
                    const int col = 64 * cq + 16 * h;
                    if (class_a && !__any_sync(kFull, n_legal > col)) break;
                    uint32_t acc[16];
                    tmem_ld16(lane_base + (uint32_t)(kColAcc + col), acc);
                    tmem_ld_wait();
                    reduce_block(acc, 128 * c + col, std::integral_constant<int, 16>{});
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
            __syncthreads();                                       // every warp has read the accumulator: the next chunk / tile may overwrite it
        }
        // ---- F: combine the two column halves of each row (and, for a chunk of a class B tile, the four chunks: last CTA in)
        S.red_m[cq][row] = m; S.red_s[cq][row] = s; S.red_g[cq][row] = gbest; S.red_l[cq][row] = lbest; S.red_i[cq][row] = ibest;
        __syncthreads();
        float M = -INFINITY, sum = 0.0f, gb = -INFINITY, lb = 0.0f;
        int ib = 0;
        // a part (m, s, choice) joins the running state: greedy -- the larger logit, the lower slot on ties; sampling -- the part's
        // choice replaces the running one with probability (its mass) / (mass so far), u uniform in (0,1)
        auto merge = [&](float pm, float ps, float pg, float pl, int pi, float u) {
            if (pm > -INFINITY) {
                const float mn = fmaxf(M, pm);
                const float before = sum * __expf(M - mn);
                sum = before + ps * __expf(pm - mn);
                M = mn;
                if (!greedy && u * sum >= before) { lb = pl; ib = pi; }
            }
            if (greedy && (pg > gb || (pg == gb && pg > -INFINITY && pi < ib))) { gb = pg; lb = pl; ib = pi; }
        };
        auto uniforms = [&](uint32_t id, float (&u)[4]) {
            uint32_t r[4];
            philox4x32_10((uint32_t)sid, ((uint32_t)(sid >> 32) << 8) | id, step, kTagGumbel, (uint32_t)seed, (uint32_t)(seed >> 32), r);
#pragma unroll
            for (int e = 0; e < 4; ++e) u[e] = ((float)(r[e] >> 8) + 0.5f) * (1.0f / 16777216.0f);
        };
        if (tid < kTileM) {                                        // (tid < 128: row == tid, and sid is this row's stream)
            float u[4] = {0.5f, 0.5f, 0.5f, 0.5f};
            if (!greedy) uniforms(0x80u + (uint32_t)c_first, u);
#pragma unroll
            for (int k = 0; k < 2; ++k) merge(S.red_m[k][tid], S.red_s[k][tid], S.red_g[k][tid], S.red_l[k][tid], S.red_i[k][tid], u[k]);
        }
        bool write_out = true;
        if (b_item) {
            const long long tb = tile >> 2;
            BPartial* mine = b_part + (tb * 4 + c_first) * kTileM;
            if (tid < kTileM) { BPartial bp; bp.m = M; bp.s = sum; bp.g = gb; bp.l = lb; bp.i = ib; mine[tid] = bp; }
            __threadfence();
            __syncthreads();
            if (tid == 0) S.tile_slot2 = atomicAdd(&b_done[tb], 1u);
            __syncthreads();
            write_out = S.tile_slot2 == 3u;                        // the other three chunks are in b_part
            if (write_out && tid < kTileM) {
                __threadfence();
                M = -INFINITY; sum = 0.0f; gb = -INFINITY; lb = 0.0f; ib = 0;
                float u[4] = {0.5f, 0.5f, 0.5f, 0.5f};
                if (!greedy) uniforms(0xC0u, u);
#pragma unroll
                for (int k = 0; k < 4; ++k) {                      // in slot order, so that ties resolve to the lowest slot as in the serial form
                    const BPartial* bp = b_part + (tb * 4 + k) * kTileM + tid;
                    merge(__ldcg(&bp->m), __ldcg(&bp->s), __ldcg(&bp->g), __ldcg(&bp->l), __ldcg(&bp->i), u[k]);
                }
            }
        }
        if (write_out && tid < kTileM && S.rowidx[bsel][tid] >= 0) {
            const long long g = S.rowidx[bsel][tid];
            actions[g] = ib;
            if (logp) logp[g] = lb - (M + __logf(sum));
            if (values) values[g] = bv + (S.part[0][tid] + S.part[1][tid]);
        }
        __syncthreads();
        if (!claim_early) {                                        // late choice of the next tile: its gather is not hidden
            if (tid == 0) S.tile_slot = atomicAdd(tile_ctr, 1u);
            __syncthreads();
            next_tile = (long long)S.tile_slot;
            if (tid < kTileM) S.rowidx[bsel ^ 1][tid] = next_tile < n_tiles ? tile_row(tile_of(next_tile), tid) : -1;
            __syncthreads();
            if (next_tile < n_tiles) gather(bsel ^ 1);
        }
        tile = next_tile;
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tmem), "r"(kPolTmemCols) : "memory");
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

// workspace: [0,16) counters, [16, 16 + 4 kMaxSplitB) arrivals per split class B tile, two int32 row lists, the chunk partials of the split tiles
static size_t pol_ws_lists() { return 16 + 4 * (size_t)kMaxSplitB; }
static size_t pol_ws_partials(long long B) { return (pol_ws_lists() + 2 * sizeof(int32_t) * (size_t)(B > 0 ? B : 1) + 15) & ~(size_t)15; }
extern "C" size_t bg_policy_workspace_bytes(long long B) { return pol_ws_partials(B) + (size_t)kMaxSplitB * 4 * kTileM * sizeof(BPartial); }

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
    cudaError_t e = cudaSuccess;
    {   // function attributes: once per device (they are not free on the launch path)
        static bool configured[64] = {};
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev < 0 || dev >= 64 || !configured[dev]) {
            e = cudaFuncSetAttribute(policy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return bg_set_error(e, "bg_policy_sample: cudaFuncSetAttribute");
            cudaFuncSetAttribute(policy_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            if (dev >= 0 && dev < 64) configured[dev] = true;
        }
    }
    // with a workspace: dynamic tile schedule, and -- when there is a mask and the logits are not wanted -- the two row
    // classes (1..128 legal slots / the rest)
    const int32_t* row_list = nullptr;
    const int32_t* row_list2 = nullptr;
    const unsigned int* class_ctr = nullptr;
    unsigned int* tile_ctr = nullptr;
    unsigned int* b_done = nullptr;
    BPartial* b_part = nullptr;
    if (workspace) {
        if (workspace_bytes < bg_policy_workspace_bytes(B)) return bg_set_error_msg(BG_ERR_INVALID, "bg_policy_sample: workspace too small");
        unsigned int* ctr = static_cast<unsigned int*>(workspace);      // [0] class A1 rows, [1] class B rows, [2] next tile, [3] class A2 rows
        e = cudaMemsetAsync(ctr, 0, pol_ws_lists(), (cudaStream_t)stream);
        if (e != cudaSuccess) return bg_set_error(e, "bg_policy_sample: memset");
        tile_ctr = ctr + 2;
        b_done = ctr + 4;
        b_part = reinterpret_cast<BPartial*>(static_cast<unsigned char*>(workspace) + pol_ws_partials(B));
        if (legal_counts && !logits_out) {
            int32_t* list = reinterpret_cast<int32_t*>(static_cast<unsigned char*>(workspace) + pol_ws_lists());
            policy_partition_kernel<<<(unsigned)((B + 255) / 256), 256, 0, (cudaStream_t)stream>>>(legal_counts, B, list, list + B, ctr);
            row_list = list; row_list2 = list + B; class_ctr = ctr;
        }
    }
    long long tiles = (B + kTileM - 1) / kTileM + 2 + 3 * kMaxSplitB;
    // two CTAs per SM (110 KB of shared memory, 256 TMEM columns, 128 registers x 256 threads each): ask for the largest carve-out
    // (cudaOccupancyMaxActiveBlocksPerMultiprocessor answers 1 for any kernel that allocates tensor memory; two such CTAs do
    // run side by side -- measured: 296 CTAs spinning 400k cycles each finish in one wave)
    const int ctas_per_sm = 2;
    long long grid = (long long)ctas_per_sm * bg_sm_count();
    if (grid > tiles) grid = tiles;
    policy_kernel<<<(unsigned)grid, kPolThreads, smem, (cudaStream_t)stream>>>(
        boards52, flags, flag_all & 1, B, legal_counts, w1_bf16, b1, wa_bf16, ba, wv, bv, seed, stream_base, step, greedy,
        row_list, row_list2, class_ctr, tile_ctr, b_done, b_part, actions, log_probs, values, logits_out);
    return bg_set_error(cudaGetLastError(), "bg_policy_sample: launch");
}
