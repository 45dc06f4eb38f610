// twoply_fused.cu -- K5: the 2-ply expansion ON CHIP.  One persistent kernel per search:
//
//   27 generator warps : an afterstate A_i per work item; its root is decoded ONCE, then for each of the 21 sorted
//                        opponent rolls (moves/get_all_dice_rolls.py:5-34) the warp runs K1's level-by-level generator
//                        (bg_movegen_warp.cuh = get_all_possible_moves, moves/get_all_moves.py:9-94) in its shared-memory
//                        scratch.  The replies stay there as 16-byte keys: 32 at a time they are expanded into the 198
//                        features (K3's encoding, ai/batching.py:78-147, turn flag = the replying player) straight into
//                        the TENSOR-MEMORY A tile (tcgen05.st; the warp fills the 32 TMEM lanes of its lane quarter).
//   1 MMA warp         : 13 x tcgen05.mma (A TMEM x W1 smem, M128 N128 K16, bf16 -> f32 in TMEM) per 128-leaf tile, as K4.
//   4 epilogue warps   : value head out of TMEM (agent/policy_network.py:58-75), terminal leaves get the win reward
//                        (environment/backgammon_env.py:156-171), then a segmented max over the leaves of the same
//                        (afterstate, roll) and ONE atomicMax per run into vmax[i*21 + r].
//
// No reply row, no feature row and no leaf value ever touches HBM: per afterstate 53 B are read and 84 B (21 maxima)
// written.  (The unfused pipeline -- bg_twoply_replies_values -- wrote and re-read ~21 x 18 rows of 57 B per afterstate
// and decoded every root 21 times.)  Leaf values are bit-identical to K4's: same operand tiles, same MMA shape, same
// epilogue arithmetic.
//
// Tiles are shared by the generator warps: a TMEM lane quarter can only be written by warps with the same (warp id % 4),
// so the warps of one class take turns (a per-class lock around "wait until the A buffer is free, write 32 rows, arrive")
// and tile t is complete when each of the four classes has contributed its quarter.  Warps that have run out of work
// keep contributing empty quarters while another class is ahead, so nobody waits for a class that has gone home.
// (i, r) items whose levels exceed this kernel's per-warp scratch (big doubles, ~2 %) go to an overflow list and through
// the ordinary path afterwards: K1's team tiers -> rows in HBM -> K4 -> overflow_max_kernel.
#include <cuda_bf16.h>
#include "bg_device.cuh"
#include "bg_movegen_common.cuh"
#include "bg_movegen_warp.cuh"
#include "bg_features.cuh"
#include "bg_tcgen05.cuh"
#include "bg_internal.h"

namespace bg {
namespace {

constexpr int kFCap = 96, kFHash = 256;          // boards per level / hash slots of a generator warp (tier 0 of K1 has 128 / 256)
constexpr int kGenWarps = 27;                    // warps 0..26; warp 27 = MMA issuer; warps 28..31 = epilogue (lane quarters 0..3)
constexpr int kMmaWarp = 27, kEpiWarp0 = 28;
constexpr int kFusedThreads = 1024;
constexpr int kWBytes = kChunks * kTileM * 16;   // 53,248: the W1 operand tile (bg_pack_w1)
constexpr int kACols = kKPad / 2;                // 104 TMEM columns per A tile
constexpr int kACol0 = 2 * kHidden;              // after the two f32 accumulators
constexpr uint32_t kInvalid = 0xFFFFFFFFu;
// Quarters of one class in flight at a time: tiles t and t+1, one per A buffer.  A third one (t+2) would wait on the same
// a_empty barrier as t for the NEXT phase, and a parity wait for "two phases ahead" passes immediately while t is still
// waiting for the MMAs of t-2 (measured: that version hung).
constexpr int kInFlight = 2;

struct FusedSmem {
    uint8_t W[kWBytes];
    WarpScratch<kFCap, kFHash> ws[kGenWarps];
    uint4 oth_chunk[kGenWarps][12];              // the 12 feature chunks of the NON-moving side of the warp's afterstate (no blot hit)
    uint4 carry_key[kGenWarps][32];              // replies waiting for a full group of 32 (always of the warp's current afterstate)
    uint8_t carry_roll[kGenWarps][32];
    uint32_t seg[4][kTileM];                     // per tile (t & 3): row -> (i*21 + r) | reward code << 29, kInvalid for an empty row
    float wv[kHidden];
    FeatureLut flut;
    unsigned long long a_full[2], a_empty[2], acc_full[2], acc_empty[2];
    uint32_t tmem_base;
    int ticket[4];                               // per class: quarters handed out so far = index of the class's next tile
    int sem[4];                                  // per class: quarters that may still be in flight (kInFlight when idle)
    int finished;                                // generator warps that have run out of work
    int final_tiles;                             // -1 until the end: total number of tiles
};

static_assert(sizeof(FusedSmem) + 1024 <= 227 * 1024, "twoply_fused_kernel: shared memory over the 227 KB per-CTA limit");

// order-preserving map f32 -> u32 (0 is below every finite value: "no reply yet")
__device__ __forceinline__ uint32_t enc_max(float v) { const uint32_t b = __float_as_uint(v); return (b & 0x80000000u) ? ~b : (b | 0x80000000u); }
__device__ __forceinline__ float dec_max(uint32_t k) { return __uint_as_float((k & 0x80000000u) ? (k ^ 0x80000000u) : ~k); }

// The internal feature row (bg_tcgen05.cuh: points of PLAYER1, points of PLAYER2, bar/off/flags, bias ones) of the leaf
// with key k, replying player p to move: p's side from the key's nibbles, the other side from the afterstate's own
// rows minus the blots hit by this reply.  25 chunks of 16 bytes -> TMEM columns [0, 100) of the row (chunk 25 is zero,
// written once at kernel start).
__device__ __forceinline__ void build_leaf_row(const uint4& k, int p, const uint32_t* rootw, const uint4* oth_chunk, uint32_t oth_bar0,
                                               uint32_t oth_off0, const FeatureLut* ft, uint32_t trow) {
    const uint2* lut = ft->units;
    const uint32_t own_col = p ? 48u : 0u, oth_col = p ? 0u : 48u;
    const uint32_t* othw = rootw + (p ? 0 : 6);
    // (a rolled loop on purpose: the generator warps' hot code has to fit the 32 KB instruction cache)
#pragma unroll 1
    for (int c = 0; c < 12; ++c) {
        const uint32_t x = (c < 4 ? k.x : (c < 8 ? k.y : k.z)) >> (8 * (c & 3));   // nibbles of points 2c, 2c+1
        const uint2 a = lut[x & 15u], b = lut[(x >> 4) & 15u];
        tmem_st4(trow + own_col + 4u * c, make_uint4(a.x, a.y, b.x, b.y));
        // the other side only differs from the afterstate's own rows where this reply hit a blot (rare)
        uint4 o = oth_chunk[c];
        const uint32_t h = (k.w >> (2 * c)) & 3u;
        if (h) {
            const uint32_t ow = othw[c >> 1] >> (16 * (c & 1));                // count bytes of the other side's points 2c, 2c+1
            const uint2 e = lut[((ow & 15u) - (h & 1u)) & 15u], f = lut[(((ow >> 8) & 15u) - (h >> 1)) & 15u];
            o = make_uint4(e.x, e.y, f.x, f.y);
        }
        tmem_st4(trow + oth_col + 4u * c, o);
    }
    const uint32_t own_pair = bar_off_pair_s((k.w >> 24) & 15u, k.w >> 28, ft);
    const uint32_t oth_pair = bar_off_pair_s(oth_bar0 + (uint32_t)__popc(k.w & 0xFFFFFFu), oth_off0, ft);
    tmem_st4(trow + 96u, make_uint4(p ? oth_pair : own_pair, p ? own_pair : oth_pair, p == 0 ? 0x00003F80u : 0x3F800000u, 0x3F803F80u));
}

// One quarter (32 rows, this warp's TMEM lane quarter q) of tile t of the class (the caller took the ticket and a slot of
// the class's semaphore, which is released here).  fill: expand `key` (of replying player p) into the A tile; else the rows
// are left as they are (an empty quarter contributed by a warp that has run out of work).  code: what the epilogue needs
// to know about the row -- (i*21 + r) | reward code << 29, or kInvalid.  ONE copy of this code in the kernel
// (__noinline__): the generator warps' hot loop must stay inside the instruction cache (with the builder inlined at every
// call site the kernel was 62 KB of SASS and stalled on instruction fetch 7.6 cycles per issue).
__device__ __noinline__ void tile_quarter(FusedSmem* smp, int q, int lane, uint32_t tmem, int t, bool fill, uint4 key, uint32_t code, int p,
                                          const uint32_t* rootw, const uint4* oth_chunk, uint32_t oth_bar0, uint32_t oth_off0) {
    FusedSmem& sm = *smp;
    const int b = t & 1;
    const uint32_t k = (uint32_t)(t >> 1);
    mbar_wait(&sm.a_empty[b], (k & 1u) ^ 1u);                                  // the MMAs of tile t-2 have read A[b]
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    if (fill) {
        const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(kACol0 + b * kACols);
        build_leaf_row(key, p, rootw, oth_chunk, oth_bar0, oth_off0, &sm.flut, trow);
        asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
    }
    sm.seg[t & 3][q * 32 + lane] = code;
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    mbar_arrive(&sm.a_full[b]);
    __syncwarp();
    if (lane == 0) atomicAdd(&sm.sem[q], 1);
}

// A slot of class q's semaphore and the class's next tile index.  While the class is saturated (both of its quarters in
// flight are waiting for older tiles to be multiplied), the pipeline is being held up by a class that has not yet taken
// its quarter of such a tile -- its warps are all busy generating.  An EMPTY quarter needs no tensor-memory write, so any
// warp may contribute it on the laggard's behalf (same semaphore / ticket protocol): the tile completes with 32 unused
// rows (the tensor pipe has plenty of slack) instead of stalling three classes.
__device__ __forceinline__ int take_ticket(FusedSmem* smp, int q, int lane, uint32_t tmem, const uint32_t* rootw, const uint4* oth_chunk) {
    FusedSmem& sm = *smp;
    for (;;) {
        int got = 0, help = -1, t = 0;
        if (lane == 0) {
            if (atomicSub(&sm.sem[q], 1) > 0) { t = atomicAdd(&sm.ticket[q], 1); got = 1; }
            else {
                atomicAdd(&sm.sem[q], 1);
                volatile int* tk = sm.ticket;
                const int mine = tk[q];
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (help < 0 && c != q && tk[c] <= mine - 3) {
                        if (atomicSub(&sm.sem[c], 1) > 0) { t = atomicAdd(&sm.ticket[c], 1); help = c; }
                        else atomicAdd(&sm.sem[c], 1);
                    }
            }
        }
        got = __shfl_sync(kFull, got, 0); help = __shfl_sync(kFull, help, 0); t = __shfl_sync(kFull, t, 0);
        if (got) return t;
        if (help >= 0) tile_quarter(smp, help, lane, tmem, t, false, make_uint4(0u, 0u, 0u, 0u), kInvalid, 0, rootw, oth_chunk, 0u, 0u);
        else __nanosleep(128);
    }
}

__global__ void __launch_bounds__(kFusedThreads, 1) twoply_fused_kernel(
    const int8_t* __restrict__ after52, const int8_t* __restrict__ movers, long long M_cap,
    const unsigned long long* __restrict__ n_rows_dev, const uint16_t* __restrict__ w1, const float* __restrict__ wv, float bv,
    uint32_t* __restrict__ vmax, int32_t* __restrict__ ovf_list, unsigned int* __restrict__ ovf_ctr,
    unsigned int* __restrict__ work_ctr, unsigned long long* __restrict__ leaves_ctr, int32_t* __restrict__ status) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    FusedSmem& sm = *reinterpret_cast<FusedSmem*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    long long M = M_cap;
    if (n_rows_dev) M = min(M, (long long)*n_rows_dev);

    {   // W1 is stored in global memory in the operand layout (bg_pack_w1): a straight asynchronous copy
        const uint32_t w_s = smem_u32(sm.W);
        for (int c = tid; c < kWBytes / 16; c += kFusedThreads) cp_async16_s(w_s + 16u * c, reinterpret_cast<const unsigned char*>(w1) + 16 * c);
        cp_async_commit();
        cp_async_wait_all();
    }
    load_feature_lut(&sm.flut);
    if (tid < kHidden) sm.wv[tid] = wv[tid];
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            mbar_init(&sm.a_full[s], kTileM);  mbar_init(&sm.a_empty[s], 1);
            mbar_init(&sm.acc_full[s], 1);     mbar_init(&sm.acc_empty[s], kTileM);
        }
        for (int c = 0; c < 4; ++c) { sm.ticket[c] = 0; sm.sem[c] = kInFlight; }
        sm.finished = 0; sm.final_tiles = -1;
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n"
                     :: "r"(smem_u32(&sm.tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = sm.tmem_base;
    if (warp < 4) {                                                            // chunk 25 (columns 200..207) of both A tiles: zeros, once
#pragma unroll
        for (int b = 0; b < 2; ++b)
            tmem_st4(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(kACol0 + b * kACols + 100), make_uint4(0u, 0u, 0u, 0u));
        asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");

    if (warp < kGenWarps) {
        // ================= generators =================
        WarpScratch<kFCap, kFHash>& S = sm.ws[warp];
        Warp<kFCap, kFHash> W(S, lane);
        const int q = warp & 3;
        uint4* ckey = sm.carry_key[warp];
        uint8_t* croll = sm.carry_roll[warp];
        unsigned long long nleaves = 0;
        for (;;) {
            unsigned int wi = 0;
            if (lane == 0) wi = atomicAdd(work_ctr, 1u);
            wi = __shfl_sync(kFull, wi, 0);
            if ((long long)wi >= M) break;
            const long long i = (long long)wi;
            const uint32_t bword = load_board_word(after52, i, lane);
            const int me = movers[i] & 1, p = me ^ 1;                          // p: the replying player
            Node root;
            const bool ok = build_root(bword, p, lane, S.rootw, W.R, root);
            __syncwarp();
            if (!ok) { if (lane == 0) atomicOr(status, BG_STATUS_BAD_INPUT); continue; }
            const uint32_t misc = S.rootw[12];
            if (((misc >> (me ? 24 : 16)) & 0xFFu) == 15u) continue;           // A_i already won by the mover: scored by the scores kernel
            // the other side (the root mover): men on the bar / borne off, and the points of p's home board it occupies
            const uint32_t oth_bar0 = (misc >> (p ? 0 : 8)) & 0xFFu, oth_off0 = (misc >> (p ? 16 : 24)) & 0xFFu;
            const uint32_t home = p == 0 ? 0xFC0000u : 0x00003Fu;
            const uint32_t block_home = W.R.block & home, blot_home = W.R.blot & home;
            if (lane < 12) {                                                   // the other side's feature chunks, once per afterstate
                const uint32_t ow = S.rootw[(p ? 0 : 6) + (lane >> 1)] >> (16 * (lane & 1));
                const uint2 e = sm.flut.units[ow & 15u], f = sm.flut.units[(ow >> 8) & 15u];
                sm.oth_chunk[warp][lane] = make_uint4(e.x, e.y, f.x, f.y);
            }
            __syncwarp();
            int carry = 0;
            // r = 21 is the flush of the replies still waiting for a full group of 32
            for (int r = 0; r <= 21; ++r) {
                int obase = 0, n = 0;
                const bool flush = r == 21;
                if (!flush) {
                    W.overflow = false;
                    W.generate(root, kRoll21[r][0], kRoll21[r][1], obase, n);
                    if (W.overflow) {                                          // too many boards per level for this scratch: the ordinary path
                        if (lane == 0) { const unsigned int k = atomicAdd(ovf_ctr, 1u); ovf_list[k] = (int32_t)(i * 21 + r); }
                        __syncwarp();
                        continue;
                    }
                    if (n == 0) continue;
                    nleaves += (unsigned long long)n;
                } else if (carry == 0) break;
                // rows [-carry, 0) are the waiting replies, [0, n) this roll's; groups of 32 go to the tile pipeline
                int pos = -carry;
                while (n - pos >= 32 || (flush && n - pos > 0)) {
                    const int idx = pos + lane;
                    const bool valid = idx < n;
                    uint4 key = make_uint4(0u, 0u, 0u, 0u);
                    int rr = r;
                    if (valid) {
                        if (idx < 0) { key = ckey[carry + idx]; rr = (int)croll[carry + idx]; }
                        else key = S.key[obase + idx];
                    }
                    uint32_t code = kInvalid;
                    if (valid) {
                        code = (uint32_t)(i * 21 + rr);
                        if ((key.w >> 28) == 15u) {                            // p has borne off 15: the leaf is a finished game
                            // environment/backgammon_env.py:156-171, 365-405: 1 normal, 1.5 gammon, 2 backgammon
                            const bool men_home = (block_home | (blot_home & ~key.w)) != 0u;
                            const bool on_bar = oth_bar0 + (uint32_t)__popc(key.w & 0xFFFFFFu) > 0u;
                            code |= (oth_off0 != 0u ? 1u : ((men_home || on_bar) ? 3u : 2u)) << 29;
                        }
                    }
                    __syncwarp();
                    const int t = take_ticket(&sm, q, lane, tmem, S.rootw, sm.oth_chunk[warp]);
                    tile_quarter(&sm, q, lane, tmem, t, true, key, code, p, S.rootw, sm.oth_chunk[warp], oth_bar0, oth_off0);
                    pos += 32;
                }
                if (pos < 0) {                                                 // no group formed: this roll's replies join the waiting ones
                    if (lane < n) { ckey[carry + lane] = S.key[obase + lane]; croll[carry + lane] = (uint8_t)r; }
                    carry += n;
                } else {
                    const int rem = max(n - pos, 0);
                    if (lane < rem) { ckey[lane] = S.key[obase + pos + lane]; croll[lane] = (uint8_t)r; }
                    carry = rem;
                }
                __syncwarp();
            }
        }
        if (lane == 0) {
            if (nleaves) atomicAdd(leaves_ctr, nleaves);
            atomicAdd(&sm.finished, 1);
        }
        __syncwarp();
        // out of work: keep the tile pipeline complete for the classes that are still producing
        for (;;) {
            int act = 0, t = 0;                                                // 1: contribute an empty quarter (tile t), 2: all done
            if (lane == 0) {
                volatile int* tk = sm.ticket;
                const int t0 = tk[0], t1 = tk[1], t2 = tk[2], t3 = tk[3];
                const int mine = tk[q], mx = max(max(t0, t1), max(t2, t3));
                if (mine < mx) {                                               // another class is ahead: take this class's next quarter
                    // (two idle warps of a class may both decide to fill and the second one run one tile ahead of everybody: an
                    // all-empty tile is harmless, the other classes then catch up the same way)
                    if (atomicSub(&sm.sem[q], 1) > 0) { t = atomicAdd(&sm.ticket[q], 1); act = 1; }
                    else atomicAdd(&sm.sem[q], 1);
                } else if (*reinterpret_cast<volatile int*>(&sm.finished) == kGenWarps && t0 == t1 && t1 == t2 && t2 == t3) act = 2;
            }
            act = __shfl_sync(kFull, act, 0);
            t = __shfl_sync(kFull, t, 0);
            if (act == 1) tile_quarter(&sm, q, lane, tmem, t, false, make_uint4(0u, 0u, 0u, 0u), kInvalid, 0, S.rootw, sm.oth_chunk[warp], 0u, 0u);
            else if (act == 2) break;
            else __nanosleep(256);
        }
        if (lane == 0) *reinterpret_cast<volatile int*>(&sm.final_tiles) = *reinterpret_cast<volatile int*>(&sm.ticket[0]);
    } else if (warp == kMmaWarp) {
        // ================= MMA issuer =================
        const uint32_t w_addr = smem_u32(sm.W);
        for (int t = 0;; ++t) {
            const int s = t & 1;
            const uint32_t it = (uint32_t)(t >> 1);
            bool stop = false;
            while (!mbar_try_wait(&sm.a_full[s], it & 1u)) {
                if (*reinterpret_cast<volatile int*>(&sm.final_tiles) == t) { stop = true; break; }
                __nanosleep(64);
            }
            if (stop) break;
            mbar_wait(&sm.acc_empty[s], (it & 1u) ^ 1u);
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            if (lane == 0) {
#pragma unroll 1
                for (int ks = 0; ks < kKPad / 16; ++ks)                        // (rolled: code size)
                    mma_bf16_ts(tmem + (uint32_t)(s * kHidden), tmem + (uint32_t)(kACol0 + s * kACols + ks * 8),
                                make_smem_desc(w_addr + ks * 2 * 2048), kIdesc, ks > 0 ? 1u : 0u);
                umma_commit(&sm.a_empty[s]);
                umma_commit(&sm.acc_full[s]);
            }
            __syncwarp();
        }
    } else {
        // ================= epilogue =================
        const int q = warp & 3;
        for (int t = 0;; ++t) {
            const int s = t & 1;
            const uint32_t it = (uint32_t)(t >> 1);
            bool stop = false;
            while (!mbar_try_wait(&sm.acc_full[s], it & 1u)) {
                if (*reinterpret_cast<volatile int*>(&sm.final_tiles) == t) { stop = true; break; }
                __nanosleep(64);
            }
            if (stop) break;
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            const uint32_t code = sm.seg[t & 3][q * 32 + lane];
            float v0 = 0.0f, v1 = 0.0f, v2 = 0.0f, v3 = 0.0f;
            uint32_t acc[32];
            const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(s * kHidden);
#pragma unroll 1
            for (int pass = 0; pass < 4; ++pass) {                             // (rolled: code size, see build_leaf_row)
                tmem_ld32(taddr + 32 * pass, acc);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; j += 4) {                              // the bias came out of the GEMM (columns 198, 199); as K4
                    const float4 ww = *reinterpret_cast<const float4*>(&sm.wv[32 * pass + j]);
                    v0 = fmaf(ww.x, fmaxf(__uint_as_float(acc[j + 0]), 0.0f), v0);
                    v1 = fmaf(ww.y, fmaxf(__uint_as_float(acc[j + 1]), 0.0f), v1);
                    v2 = fmaf(ww.z, fmaxf(__uint_as_float(acc[j + 2]), 0.0f), v2);
                    v3 = fmaf(ww.w, fmaxf(__uint_as_float(acc[j + 3]), 0.0f), v3);
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
            mbar_arrive(&sm.acc_empty[s]);
            float v = bv + ((v0 + v1) + (v2 + v3));
            const uint32_t rc = (code >> 29) & 3u;
            if (rc) v = rc == 1u ? 1.0f : (rc == 2u ? 1.5f : 2.0f);
            const bool valid = code != kInvalid;
            uint32_t key = valid ? enc_max(v) : 0u;
            // segmented max over runs of equal code (the leaves of one (afterstate, roll) are contiguous rows)
            const uint32_t prev = __shfl_up_sync(kFull, code, 1);
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t ok = __shfl_down_sync(kFull, key, o), oc = __shfl_down_sync(kFull, code, o);
                if (lane + o < 32 && oc == code) key = max(key, ok);
            }
            if (valid && (lane == 0 || prev != code)) atomicMax(&vmax[code & 0x1FFFFFFFu], key);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tmem), "r"(512u) : "memory");
}

// vmax of the (i, r) items that went through the ordinary path: max over their reply rows' leaf values
__global__ void __launch_bounds__(256) overflow_max_kernel(const int32_t* __restrict__ list, const unsigned int* __restrict__ n_dev,
                                                           const long long* __restrict__ starts, const int32_t* __restrict__ counts,
                                                           const float* __restrict__ leaf_values, uint32_t* __restrict__ vmax) {
    const unsigned int n = *n_dev;
    const int lane = threadIdx.x & 31;
    for (unsigned int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; k < n; k += (gridDim.x * blockDim.x) >> 5) {
        const int g = list[k];
        const int c = counts[g];
        if (c <= 0) continue;
        const float* lv = leaf_values + starts[g];
        float v = -INFINITY;
        for (int j = lane; j < c; j += 32) v = fmaxf(v, lv[j]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, o));
        if (lane == 0) vmax[g] = enc_max(v);
    }
}

__device__ __forceinline__ float win_reward_row(const int8_t* b, int p) {       // backgammon_env.py:156-171,365-405
    const int o = p ^ 1;
    if (b[50 + o] != 0) return 1.0f;
    bool bgm = b[48 + o] > 0;
    const int8_t* orow = b + 24 * o;
    const int h0 = p == 0 ? 18 : 0;
#pragma unroll
    for (int i = 0; i < 6; ++i) bgm = bgm || orow[h0 + i] > 0;
    return bgm ? 2.0f : 1.5f;
}

// score_i = +win reward if movers[i] has borne off 15 in A_i, else - sum_r p_r * (vmax[i][r], or the pass value when the
// opponent has no reply); fixed roll order, separate multiply and add (same arithmetic as twoply_scores_kernel)
__global__ void __launch_bounds__(256) fused_scores_kernel(const uint32_t* __restrict__ vmax, const float* __restrict__ pass_values,
                                                           const int8_t* __restrict__ after52, const int8_t* __restrict__ movers,
                                                           long long M_cap, const unsigned long long* __restrict__ n_rows_dev,
                                                           float* __restrict__ scores) {
    long long M = M_cap;
    if (n_rows_dev) M = min(M, (long long)*n_rows_dev);
    const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (i >= M) return;
    const int8_t* a = after52 + i * kBoardBytes;
    const int me = movers[i] & 1;
    if (a[50 + me] == 15) {
        if (lane == 0) scores[i] = win_reward_row(a, me);
        return;
    }
    float term = 0.0f;
    if (lane < 21) {
        const uint32_t k = vmax[i * 21 + lane];
        const float v = k ? dec_max(k) : pass_values[i];
        const bool dbl = kRoll21[lane][0] == kRoll21[lane][1];
        term = __fmul_rn(dbl ? 1.0f / 36.0f : 2.0f / 36.0f, v);               // get_all_dice_rolls.py:19-32
    }
    float acc = 0.0f;
#pragma unroll
    for (int r = 0; r < 21; ++r) acc = __fadd_rn(acc, __shfl_sync(kFull, term, r));
    if (lane == 0) scores[i] = -acc;
}

__global__ void stats_kernel(const unsigned long long* __restrict__ n_after, long long cap, const unsigned long long* __restrict__ leaves_chip,
                             const unsigned long long* __restrict__ leaves_ovf, const unsigned int* __restrict__ ovf_items,
                             unsigned long long* __restrict__ stats) {
    stats[0] = min((unsigned long long)cap, *n_after);
    stats[1] = *leaves_chip + *leaves_ovf;
    stats[2] = *ovf_items;
    stats[3] = *leaves_ovf;
}

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace

// workspace layout of bg_twoply
struct TwoPlyPlan {
    size_t root_ws, row_players, counters, vmax, pass_v, list_a, ovf_counts, ovf_starts, ovf_ws, ovf_rows52, ovf_rowp, ovf_leaf, total;
    size_t root_ws_bytes, ovf_ws_bytes;
    long long ovf_cap_rows;
};
static TwoPlyPlan make_plan(long long B, long long cap) {
    TwoPlyPlan p{};
    const long long W = cap * 21;
    size_t o = 0;
    auto take = [&](size_t bytes) { const size_t at = o; o = align_up(o + bytes, 256); return at; };
    p.root_ws_bytes = bg_movegen_workspace_bytes(B);
    p.ovf_ws_bytes = bg_movegen_workspace_bytes(W);
    p.ovf_cap_rows = cap * 64 + 262144;
    p.root_ws = take(p.root_ws_bytes);
    p.row_players = take((size_t)cap);
    p.counters = take(256);
    p.vmax = take(sizeof(uint32_t) * (size_t)W);
    p.pass_v = take(sizeof(float) * (size_t)cap);
    p.list_a = take(sizeof(int32_t) * (size_t)W);
    p.ovf_counts = take(sizeof(int32_t) * (size_t)W);
    p.ovf_starts = take(sizeof(long long) * (size_t)W);
    p.ovf_ws = take(p.ovf_ws_bytes);
    p.ovf_rows52 = take((size_t)kBoardBytes * (size_t)p.ovf_cap_rows);
    p.ovf_rowp = take((size_t)p.ovf_cap_rows);
    p.ovf_leaf = take(sizeof(float) * (size_t)p.ovf_cap_rows);
    p.total = o;
    return p;
}

}  // namespace bg

using namespace bg;

extern "C" size_t bg_twoply_workspace_bytes(long long B, long long max_afterstates) {
    if (B < 1) B = 1;
    if (max_afterstates < 1) max_afterstates = 1;
    return make_plan(B, max_afterstates).total;
}

extern "C" size_t bg_workspace_bytes(int kind, long long B) {
    switch (kind) {
        case BG_WS_MOVEGEN: return bg_movegen_workspace_bytes(B);
        case BG_WS_POLICY: return bg_policy_workspace_bytes(B);
        case BG_WS_TWOPLY: return bg_twoply_workspace_bytes(B, BG_TWOPLY_DEFAULT_ROWS_PER_ROOT * (B > 0 ? B : 1) + 4096);
        default: return 0;
    }
}

extern "C" int bg_twoply(const int8_t* boards52, const int8_t* players, const int8_t* dice, long long B,
                         const uint16_t* w1_bf16, const float* wv, float bv, long long max_afterstates,
                         int8_t* afterstates52, float* scores, int32_t* counts, long long* starts, int32_t* best,
                         float* best_score, unsigned long long* stats, int32_t* status, void* workspace,
                         size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (B < 0 || max_afterstates < 0) return bg_set_error_msg(BG_ERR_INVALID, "bg_twoply: negative size");
    if (B == 0) return BG_OK;
    if (!boards52 || !players || !dice || !w1_bf16 || !wv || !afterstates52 || !scores || !counts || !starts || !best || !status || !workspace)
        return bg_set_error_msg(BG_ERR_INVALID, "bg_twoply: null pointer");
    if (max_afterstates < 1 || max_afterstates * 21 >= (1LL << 29))
        return bg_set_error_msg(BG_ERR_INVALID, "bg_twoply: max_afterstates out of range (1 .. 2^29 / 21)");
    const TwoPlyPlan P = make_plan(B, max_afterstates);
    if (workspace_bytes < P.total) return bg_set_error_msg(BG_ERR_INVALID, "bg_twoply: workspace too small (bg_twoply_workspace_bytes)");
    unsigned char* ws = static_cast<unsigned char*>(workspace);
    int8_t* row_players = reinterpret_cast<int8_t*>(ws + P.row_players);
    // counters (zeroed): [0] root row allocator u64, [8] overflow-slab row allocator u64, [16] leaves on chip u64,
    // [24] generator work counter u32, [28] overflow item counter u32
    unsigned long long* root_alloc = reinterpret_cast<unsigned long long*>(ws + P.counters);
    unsigned long long* ovf_alloc = root_alloc + 1;
    unsigned long long* leaves_chip = root_alloc + 2;
    unsigned int* work_ctr = reinterpret_cast<unsigned int*>(ws + P.counters + 24);
    unsigned int* ovf_ctr = work_ctr + 1;
    uint32_t* vmax = reinterpret_cast<uint32_t*>(ws + P.vmax);
    float* pass_v = reinterpret_cast<float*>(ws + P.pass_v);
    int32_t* list_a = reinterpret_cast<int32_t*>(ws + P.list_a);
    int32_t* ovf_counts = reinterpret_cast<int32_t*>(ws + P.ovf_counts);
    long long* ovf_starts = reinterpret_cast<long long*>(ws + P.ovf_starts);
    int8_t* ovf_rows = reinterpret_cast<int8_t*>(ws + P.ovf_rows52);
    int8_t* ovf_rowp = reinterpret_cast<int8_t*>(ws + P.ovf_rowp);
    float* ovf_leaf = reinterpret_cast<float*>(ws + P.ovf_leaf);
    const long long cap = max_afterstates;

    cudaError_t e = cudaMemsetAsync(ws + P.counters, 0, 256, stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(vmax, 0, sizeof(uint32_t) * (size_t)cap * 21, stream);
    if (e != cudaSuccess) return bg_set_error(e, "bg_twoply: memset");
    // 1. the roots' legal plays (K1, slab form: rows of root b at starts[b], reference order)
    int rc = movegen_run(boards52, players, dice, B, 1, 0, 2, nullptr, 0, afterstates52, cap, row_players, nullptr, nullptr, counts,
                         starts, root_alloc, status, ws + P.root_ws, P.root_ws_bytes, stream);
    if (rc != BG_OK) return rc;
    // 2. value of every afterstate with the opponent to move (the roll's value when the opponent has no reply)
    rc = mlp_value_launch(afterstates52, row_players, 0, 1, cap, nullptr, root_alloc, w1_bf16, nullptr, wv, bv, 0, pass_v, stream);
    if (rc != BG_OK) return rc;
    // 3. replies x 21 rolls, features, MLP and the per-(afterstate, roll) maximum, on chip
    const size_t smem = sizeof(FusedSmem) + 1024;
    e = cudaFuncSetAttribute(twoply_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return bg_set_error(e, "bg_twoply: cudaFuncSetAttribute");
    long long grid = bg_sm_count();
    if (grid > (cap + kGenWarps - 1) / kGenWarps) grid = (cap + kGenWarps - 1) / kGenWarps;
    twoply_fused_kernel<<<(unsigned)grid, kFusedThreads, smem, stream>>>(afterstates52, row_players, cap, root_alloc, w1_bf16, wv, bv, vmax,
                                                                         list_a, ovf_ctr, work_ctr, leaves_chip, status);
    e = cudaGetLastError();
    if (e != cudaSuccess) return bg_set_error(e, "bg_twoply: fused kernel launch");
    // 4. the (afterstate, roll) items that did not fit the on-chip scratch: K1's team tiers -> rows -> K4 -> maximum
    {
        unsigned char* ows = ws + P.ovf_ws;
        unsigned int* ctr = reinterpret_cast<unsigned int*>(ows);                 // [0] tier-1 work ctr, [1] tier-2 count, [2] tier-2 work ctr
        int32_t* list_b = reinterpret_cast<int32_t*>(ows + BG_WS_LISTS);
        e = cudaMemsetAsync(ows, 0, BG_WS_LISTS, stream);
        if (e != cudaSuccess) return bg_set_error(e, "bg_twoply: memset");
        rc = movegen_team_mid(afterstates52, row_players, nullptr, ovf_ctr, list_a, 21, 1, 2, nullptr, 0, ovf_rows, P.ovf_cap_rows, ovf_rowp,
                              nullptr, nullptr, ovf_counts, ovf_starts, ovf_alloc, status, ctr + 0, list_b, ctr + 1, stream, 128);
        if (rc != BG_OK) return rc;
        rc = movegen_team_big(afterstates52, row_players, nullptr, ctr + 1, nullptr, list_b, 21, 1, 2, nullptr, 0, ovf_rows, P.ovf_cap_rows, ovf_rowp,
                              nullptr, nullptr, ovf_counts, ovf_starts, ovf_alloc, status, ctr + 2, stream);
        if (rc != BG_OK) return rc;
        rc = mlp_value_launch(ovf_rows, ovf_rowp, 0, 0, P.ovf_cap_rows, nullptr, ovf_alloc, w1_bf16, nullptr, wv, bv, 1, ovf_leaf, stream);
        if (rc != BG_OK) return rc;
        overflow_max_kernel<<<bg_sm_count() * 2, 256, 0, stream>>>(list_a, ovf_ctr, ovf_starts, ovf_counts, ovf_leaf, vmax);
        e = cudaGetLastError();
        if (e != cudaSuccess) return bg_set_error(e, "bg_twoply: overflow max launch");
    }
    // 5. scores, and the best play of every root
    fused_scores_kernel<<<(unsigned)((cap * 32 + 255) / 256), 256, 0, stream>>>(vmax, pass_v, afterstates52, row_players, cap, root_alloc, scores);
    e = cudaGetLastError();
    if (e != cudaSuccess) return bg_set_error(e, "bg_twoply: scores launch");
    rc = bg_segment_argmax(scores, starts, counts, B, best, best_score, stream);
    if (rc != BG_OK) return rc;
    if (stats) {
        stats_kernel<<<1, 1, 0, stream>>>(root_alloc, cap, leaves_chip, ovf_alloc, ovf_ctr, stats);
        e = cudaGetLastError();
        if (e != cudaSuccess) return bg_set_error(e, "bg_twoply: stats launch");
    }
    return BG_OK;
}
