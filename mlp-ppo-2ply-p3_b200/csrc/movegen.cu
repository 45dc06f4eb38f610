// movegen.cu -- K1: batched legal-move (afterstate) generation, one warp per position.
//
// Replaces, for a whole batch resident in HBM, the reference's
//   get_all_possible_moves              src/moves/get_all_moves.py:9-94
//   handle_non_doubles / handle_doubles src/moves/handle_moves.py:109-310
//   add_unique_board / board_hash       src/moves/handle_moves.py:313-341, board/immutable_board.py:236-246
//   get_moves_with_one_die (+ helpers)  src/moves/move_logic.py:20-275, src/moves/conditions.py:7-147
//   move_checker                        src/board/immutable_board.py:42-89
//
// Algorithm (DESIGN.md "K1"): the reference walks a 2-deep (non-doubles, both die
// orders) or 4-deep (doubles) DFS and keeps the first sequence that reaches each
// distinct board.  Here the same tree is expanded LEVEL BY LEVEL: a level is an
// ordered list of distinct boards in shared memory; one lane builds one child
// (parent, move) of the next level, children are enumerated in the reference's
// lexicographic order 32 at a time, duplicates inside the 32 are found with
// __match_any_sync on the packed board, duplicates against earlier children with a
// per-warp shared-memory hash set, and survivors are compacted in order with
// ballot + popc prefix sums.  Removing a duplicate intermediate board removes a
// subtree whose leaves were all seen earlier, so the surviving leaves and their
// order (= first-occurrence order of the reference's DFS) are unchanged; the
// emitted rows are therefore in the reference's legal_moves order, action
// indices coincide, and the env's "first max_legal_moves" truncation
// (src/environment/backgammon_env.py:218-223) is a plain prefix.
#include "bg_device.cuh"
#include "bg_movegen_common.cuh"
#include "bg_movegen_warp.cuh"
#include "bg_features.cuh"
#include "bg_internal.h"

namespace bg {

// mode: 0 = count only; 1 = write rows at offsets[b]; 2 = slab (atomicAdd on *alloc, writes starts[b])
// FEATS: also write the bf16 feature row of every afterstate (optional fused K3; a separate instantiation so that
// the common one stays inside the instruction cache)
template <int CAP, int HS, bool FEATS>
__global__ void __launch_bounds__(256, 4) movegen_kernel(
    const int8_t* __restrict__ boards, const int8_t* __restrict__ players, const int8_t* __restrict__ dice,
    long long B, const unsigned int* __restrict__ nwork_dev, const int32_t* __restrict__ worklist,
    int replicate, int flip_player, int mode, const long long* __restrict__ offsets, int max_rows,
    int8_t* __restrict__ after, long long after_cap_rows, int8_t* __restrict__ row_players,
    uint16_t* __restrict__ row_feats, int32_t* __restrict__ counts_true, int32_t* __restrict__ counts, long long* __restrict__ starts,
    unsigned long long* __restrict__ alloc, int32_t* __restrict__ status,
    unsigned int* __restrict__ work_ctr, int32_t* __restrict__ overflow_list, unsigned int* __restrict__ overflow_ctr,
    int32_t* __restrict__ overflow_list_big, unsigned int* __restrict__ overflow_ctr_big) {
    static_assert(CAP * 16 >= 32 * kBoardWords * 4, "a region must be able to stage 32 output rows");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    WarpScratch<CAP, HS>& S = reinterpret_cast<WarpScratch<CAP, HS>*>(smem_raw)[warp];
    __shared__ uint32_t s_lut[32], s_desc[32];          // K3's chunk tables, for the fused feature output
    if (FEATS) { load_chunk_tables(s_lut, s_desc); __syncthreads(); }
    const long long nwork = nwork_dev ? (long long)*nwork_dev : B;

    // (Reading the work counter / the item's global loads one or two items AHEAD was tried and lost 15 us: with ~14 items per
    // warp and item costs from 3 to 80 us, items parked in a busy warp's pipeline cost more at the tail than the hidden latency
    // gains -- scripts/exp_k1_variants.py, profiles/r2_summary.md.)
    // (Also tried, and not kept: walking the items twice, doubles -- 4 levels, up to ~80 us for a lone warp -- first and the other
    // rolls after them.  With the games physically sorted doubles-first K1 takes 14 us less, but the in-kernel version (groups of 8
    // dice loads, doubles skipped in the second walk) pays that back in extra fetches: 181 us against 178.)
    struct Item { unsigned int wi; uint32_t bword; int pl, d0, d1; };
    for (;;) {
        unsigned int wi0 = 0;
        if (lane == 0) wi0 = atomicAdd(work_ctr, 1u);
        wi0 = __shfl_sync(kFull, wi0, 0);
        if ((long long)wi0 >= nwork) break;
        Item cur; cur.wi = wi0;
        {
            const long long g0 = worklist ? (long long)worklist[wi0] : (long long)wi0;
            const long long src = replicate > 1 ? g0 / replicate : g0;
            cur.bword = load_board_word(boards, src, lane);
            cur.pl = players[src];
            if (replicate > 1) { const int r = (int)(g0 - src * replicate); cur.d0 = kRoll21[r][0]; cur.d1 = kRoll21[r][1]; }
            else { cur.d0 = dice[2 * g0]; cur.d1 = dice[2 * g0 + 1]; }
        }
        const long long g = worklist ? (long long)worklist[cur.wi] : (long long)cur.wi;

        // ---- the work item, its root (13 words, coalesced) and the mover-relative view (bg_movegen_common.cuh)
        const uint32_t bword = cur.bword;
        const int player = (cur.pl ^ flip_player) & 1, d0 = cur.d0, d1 = cur.d1;
        Warp<CAP, HS> W(S, lane);
        Node root;
        const bool bad = !build_root(bword, player, lane, S.rootw, W.R, root) ||
                         d0 < 1 || d0 > 6 || d1 < 1 || d1 > 6;
        __syncwarp();

        int obase = 0, n = 0;
        if (!bad) W.generate(root, d0, d1, obase, n);

        if (bad) {
            if (lane == 0) {
                atomicOr(status, BG_STATUS_BAD_INPUT);
                if (counts_true) counts_true[g] = -1;
                if (counts) counts[g] = 0;
            }
            __syncwarp();
            continue;
        }
        if (W.overflow) {
            // too many boards for this launch's per-warp scratch: hand the position to the large-scratch pass
            if (lane == 0) {
                // a doubles position whose THIRD level (or an earlier one) already exceeds the scratch is headed for a last level of
                // several hundred boards: straight to the big-scratch tier, which then runs beside tier 1 instead of after it
                // (scripts/level_stats.py: 0.06 % of the positions, and every position that overflows tier 1 is among them)
                if (overflow_list_big && W.ovf_stage <= 2) { unsigned int k = atomicAdd(overflow_ctr_big, 1u); overflow_list_big[k] = (int32_t)g; }
                else if (overflow_list) { unsigned int k = atomicAdd(overflow_ctr, 1u); overflow_list[k] = (int32_t)g; }
                else {
                    atomicOr(status, BG_STATUS_SCRATCH_OVERFLOW);
                    if (counts_true) counts_true[g] = -1;
                    if (counts) counts[g] = 0;
                }
            }
            __syncwarp();
            continue;
        }
        const int nw = (max_rows > 0 && n > max_rows) ? max_rows : n;     // rows written (env truncation)
        // slab mode: the allocation (an atomic on one address, ~1 us) is issued here, before the counts are written, and read where the
        // rows are stored (FEATS: after the first rows have been staged)
        long long start = 0;
        unsigned long long s0 = 0;
        if (mode == 1) start = offsets[g];
        else if (mode == 2 && lane == 0) s0 = atomicAdd(alloc, (unsigned long long)nw);
        if (lane == 0) {
            if (counts_true) counts_true[g] = n;
            if (counts) counts[g] = nw;
        }
        if (mode == 2 && nw == 0 && lane == 0 && starts) starts[g] = (long long)s0;
        if (mode != 0 && nw > 0) {
            {
                uint32_t* stage = reinterpret_cast<uint32_t*>(&S.key[obase < CAP ? CAP : 0]);
                uint32_t* gout = nullptr;
                const RowContext rc = make_row_context(player, S.rootw);
#ifndef BG_K1_STAGED_OUTPUT
                if (!FEATS) {
                    // every lane writes its own row straight to global memory (13 word stores at a 52-byte stride: L2 merges the sectors;
                    // K1 is issue bound, not memory bound, and the staged copy cost more instructions than these stores cost transactions)
                    if (mode == 2) {
                        start = (long long)__shfl_sync(kFull, s0, 0);
                        if (lane == 0 && starts) starts[g] = start;
                    }
                    if (start + nw > after_cap_rows) {
                        if (lane == 0) { atomicOr(status, BG_STATUS_OUTPUT_OVERFLOW); if (counts) counts[g] = 0; }
                    } else {
                        gout = reinterpret_cast<uint32_t*>(after) + start * kBoardWords;
                        for (int r = lane; r < nw; r += 32) {
                            store_row_direct(S.key[obase + r], player, rc, S.rootw, gout + (long long)r * kBoardWords);
                            if (row_players) row_players[start + r] = (int8_t)player;
                        }
                    }
                } else
#endif
                for (int r0 = 0; r0 < nw; r0 += 32) {
                    int r = r0 + lane;
                    if (r < nw) {
                        build_row(S.key[obase + r], player, rc, S.rootw, stage + lane * kBoardWords);
                    }
                    __syncwarp();
                    if (r0 == 0) {
                        if (mode == 2) {
                            start = (long long)__shfl_sync(kFull, s0, 0);
                            if (lane == 0 && starts) starts[g] = start;
                        }
                        if (start + nw > after_cap_rows) {
                            if (lane == 0) { atomicOr(status, BG_STATUS_OUTPUT_OVERFLOW); if (counts) counts[g] = 0; }
                            break;
                        }
                        gout = reinterpret_cast<uint32_t*>(after) + start * kBoardWords;
                    }
                    int rows = min(32, nw - r0);
                    {   // copy out: 13 predicated word copies at immediate offsets (no per-iteration address arithmetic)
                        const int total = rows * kBoardWords;
                        uint32_t* gp = gout + (long long)r0 * kBoardWords + lane;
                        const uint32_t* sp = stage + lane;
#pragma unroll
                        for (int i = 0; i < kBoardWords; ++i) if (lane + 32 * i < total) gp[32 * i] = sp[32 * i];
                    }
                    if (row_players && lane < rows) row_players[start + r0 + lane] = (int8_t)player;
                    if (FEATS && row_feats) {
                        // fused K3: the 208-wide bf16 feature rows of these afterstates (mover's turn flag,
                        // ai/batching.py:72-74), 16 bytes per lane, contiguous in global memory
                        uint4* fdst = reinterpret_cast<uint4*>(row_feats + (start + r0) * (long long)BG_FEAT_LD_BF16);
                        const uint8_t* sb = reinterpret_cast<const uint8_t*>(stage);
                        // lane's chunks: c = lane, lane+32, ...; (r, k) advance by (1, +6) with carry, no division
                        int r = lane >= 26 ? 1 : 0, k = lane >= 26 ? lane - 26 : lane;
                        const int nchunk = rows * 26;
                        for (int c = lane; c < nchunk; c += 64) {
                            const int k2 = k + 6 >= 26 ? k + 6 - 26 : k + 6, r2 = r + 1 + (k + 6 >= 26 ? 1 : 0);
                            const uint4 v0 = chunk_from_desc(sb + r * kBoardBytes, player, s_desc[k], s_lut);
                            const bool has2 = c + 32 < nchunk;
                            uint4 v1 = make_uint4(0u, 0u, 0u, 0u);
                            if (has2) v1 = chunk_from_desc(sb + r2 * kBoardBytes, player, s_desc[k2], s_lut);
                            fdst[c] = v0;
                            if (has2) fdst[c + 32] = v1;
                            k = k2 + 6 >= 26 ? k2 + 6 - 26 : k2 + 6; r = r2 + 1 + (k2 + 6 >= 26 ? 1 : 0);
                        }
                    }
                    __syncwarp();
                }
            }
        }
        __syncwarp();
    }
}

template <int CAP, int HS, int WARPS>
static int launch_movegen(const int8_t* boards, const int8_t* players, const int8_t* dice, long long B,
                          const unsigned int* nwork_dev, const int32_t* worklist, int replicate, int flip_player, int mode,
                          const long long* offsets, int max_rows, int8_t* after, long long after_cap_rows,
                          int8_t* row_players, uint16_t* row_feats, int32_t* counts_true, int32_t* counts, long long* starts, unsigned long long* alloc,
                          int32_t* status, unsigned int* work_ctr, int32_t* overflow_list, unsigned int* overflow_ctr,
                          int32_t* overflow_list_big, unsigned int* overflow_ctr_big, long long grid_hint, cudaStream_t stream) {
    size_t smem = sizeof(WarpScratch<CAP, HS>) * WARPS;
    auto kern = row_feats ? movegen_kernel<CAP, HS, true> : movegen_kernel<CAP, HS, false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return bg_set_error(e, "movegen: cudaFuncSetAttribute");
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, WARPS * 32, smem);
    if (occ < 1) occ = 1;
    long long grid = (long long)bg_sm_count() * occ;
    if (grid_hint > 0 && grid > grid_hint) grid = grid_hint;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, WARPS * 32, smem, stream>>>(boards, players, dice, B, nwork_dev, worklist, replicate, flip_player, mode, offsets,
                                                       max_rows, after, after_cap_rows, row_players, row_feats,
                                                       counts_true, counts, starts, alloc, status, work_ctr, overflow_list, overflow_ctr, overflow_list_big, overflow_ctr_big);
    return bg_set_error(cudaGetLastError(), "movegen: launch");
}

}  // namespace bg

using namespace bg;

// Workspace layout (bytes; constants in bg_internal.h): [0] work_ctr0, [4] overflow_ctr A, [8] work_ctr1, [12] overflow_ctr B,
// [16] work_ctr2, [24] list B's length after tier 0 (first entry of tier 2's second pass), [28] work counter of that pass,
// [BG_WS_ROWS_AFTER_TIER0 = 40] u64 row-count snapshot after tier 0 (written by callers that fork there),
// [BG_WS_LISTS = 64 ..] overflow list A int32[B], then overflow list B int32[B]
extern "C" size_t bg_movegen_workspace_bytes(long long B) { return BG_WS_LISTS + 2 * sizeof(int32_t) * (size_t)(B > 0 ? B : 1); }

namespace {
// tier 2's own stream (forked from and joined to the caller's stream inside movegen_run), one per device and host thread
struct TierStreams { cudaStream_t stream = nullptr; cudaEvent_t tier0 = nullptr, tier2 = nullptr; int device = -1; };
thread_local TierStreams g_tier_streams[16];
TierStreams* tier_streams() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
    TierStreams& t = g_tier_streams[dev];
    if (t.device != dev) {
        if (cudaStreamCreateWithFlags(&t.stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&t.tier0, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&t.tier2, cudaEventDisableTiming) != cudaSuccess)
            return nullptr;
        t.device = dev;
    }
    return &t;
}
}  // namespace

int bg::movegen_run(const int8_t* boards, const int8_t* players, const int8_t* dice, long long B, int replicate,
                    int flip_player, int mode,
                    const long long* offsets, int max_rows, int8_t* after, long long after_cap_rows,
                    int8_t* row_players, uint16_t* row_feats, int32_t* counts_true, int32_t* counts, long long* starts,
                    unsigned long long* alloc, int32_t* status, void* workspace, size_t ws_bytes, cudaStream_t stream,
                    const MovegenTier0Hook* hook) {
    if (B < 0) return bg_set_error_msg(BG_ERR_INVALID, "movegen: negative batch");
    if (B == 0) return BG_OK;
    if (replicate < 1) replicate = 1;
    if (!boards || !players || (!dice && replicate == 1) || !status || !workspace)
        return bg_set_error_msg(BG_ERR_INVALID, "movegen: null pointer");
    if (mode != 0 && !after) return bg_set_error_msg(BG_ERR_INVALID, "movegen: null output");
    if (mode == 1 && !offsets) return bg_set_error_msg(BG_ERR_INVALID, "movegen: null offsets");
    if (mode == 2 && (!alloc || !starts)) return bg_set_error_msg(BG_ERR_INVALID, "movegen: null slab allocator");
    if (ws_bytes < bg_movegen_workspace_bytes(B)) return bg_set_error_msg(BG_ERR_INVALID, "movegen: workspace too small");
    if (B > 0x7FFFFFF0LL) return bg_set_error_msg(BG_ERR_INVALID, "movegen: batch too large");
    unsigned char* ws = static_cast<unsigned char*>(workspace);
    unsigned int* ctr = reinterpret_cast<unsigned int*>(ws);
    int32_t* list_a = reinterpret_cast<int32_t*>(ws + BG_WS_LISTS);
    int32_t* list_b = list_a + B;
    cudaError_t e = cudaMemsetAsync(ws, 0, BG_WS_LISTS, stream);
    if (e != cudaSuccess) return bg_set_error(e, "movegen: memset");
    // Tier 0: every position, BG_MOVEGEN_CAP_SMALL boards per level, 8 warps per CTA.
    // (Splitting tier 0 into 2..8 launches over ranges of games, so that K3 could encode part p's rows beside part p + 1, was tried:
    // every extra part cost ~26 us -- each launch ends with its own tail of lone warps finishing big doubles -- and the step got
    // slower, 265 -> 298 us with two parts; scripts/exp_tier0_pipeline.py in the history, profiles/r2_summary.md.)
    // Large batches: tier 0 sends the positions that are certainly huge straight to tier 2's list, and tier 2 takes those on a
    // stream of its own BESIDE tier 1 (both are latency bound: one CTA per position); a second, usually empty, tier-2 pass after
    // tier 1 takes what tier 1 itself could not hold.
    // Not when a consumer of the rows runs beside the tiers (the hook: K3 / K4 on the caller's side stream): that consumer is HBM
    // bound and needs its CTAs resident; tier 2's 165 KB CTAs beside it cost the env step 5 us more than they save (265 -> 270 us),
    // and so does the routing alone (272 us).  K1 on its own (policy rollouts, bg_movegen_*, the unfused 2-ply): 178 -> 166 us.
    const bool hooked = hook && mode == 2;
    TierStreams* ts = (B >= 4096 && !hooked) ? tier_streams() : nullptr;
    int rc = launch_movegen<BG_MOVEGEN_CAP_SMALL, 2 * BG_MOVEGEN_CAP_SMALL, 8>(
        boards, players, dice, B, nullptr, nullptr, replicate, flip_player, mode, offsets, max_rows, after,
        after_cap_rows, row_players, row_feats, counts_true, counts, starts, alloc, status, ctr + 0, list_a, ctr + 1,
        ts ? list_b : nullptr, ctr + 3, (B + 7) / 8, stream);
    if (rc != BG_OK) return rc;
    bool forked = false;
    if (ts) {
        // fork: [snapshot of list B's length] -> [tier 2 on entries [0, snapshot)] on ts->stream
        e = cudaEventRecord(ts->tier0, stream);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(ts->stream, ts->tier0, 0);
        if (e != cudaSuccess) return bg_set_error(e, "movegen: fork");
        forked = true;                                              // from here on every exit goes through the join
        e = cudaMemcpyAsync(ctr + 6, ctr + 3, sizeof(unsigned int), cudaMemcpyDeviceToDevice, ts->stream);
        rc = e == cudaSuccess ? BG_OK : bg_set_error(e, "movegen: snapshot");
        if (rc == BG_OK)
            rc = movegen_team_big(boards, players, dice, ctr + 6, nullptr, list_b, replicate, flip_player, mode, offsets, max_rows, after,
                                  after_cap_rows, row_players, row_feats, counts_true, counts, starts, alloc, status, ctr + 4, ts->stream);
    }
    // Slab mode: rows [0, *alloc) are final once tier 0 is done (the other tiers only append).  A caller that wants to
    // consume them while the latency-bound overflow tiers run gets the row count snapshot and a fork point here.
    if (rc == BG_OK && hook && mode == 2) {
        e = cudaMemcpyAsync(hook->rows_after_tier0, alloc, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, stream);
        rc = e == cudaSuccess ? hook->fn(hook->user) : bg_set_error(e, "movegen: snapshot");
    }
    // Tier 1: the (~1 %) positions whose levels did not fit: one CTA per position (movegen_team.cu),
    // BG_MOVEGEN_CAP_MID boards per level.  Work counts of tiers 1 and 2 are read from device memory, so no host
    // synchronisation is needed.
    if (rc == BG_OK)
        rc = movegen_team_mid(boards, players, dice, ctr + 1, list_a, replicate, flip_player, mode, offsets, max_rows, after,
                              after_cap_rows, row_players, row_feats, counts_true, counts, starts, alloc, status, ctr + 2, list_b,
                              ctr + 3, stream, ((hook && mode == 2) || B >= 262144) ? 128 : 0);   // 128 threads per position when something shares the GPU or the list is long (throughput, not latency)
    if (forked) {                                                   // join
        e = cudaEventRecord(ts->tier2, ts->stream);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(stream, ts->tier2, 0);
        if (e != cudaSuccess && rc == BG_OK) rc = bg_set_error(e, "movegen: join");
    }
    if (rc != BG_OK) return rc;
    // Tier 2: the rest (> BG_MOVEGEN_CAP_MID boards in a level): one big CTA per position, BG_MOVEGEN_CAP_BIG
    // boards per level.  Positions that do not fit even this raise BG_STATUS_SCRATCH_OVERFLOW (never dropped silently).
    // Forked form: entries [snapshot, end) of list B -- what tier 1 appended.
    return movegen_team_big(boards, players, dice, ctr + 3, forked ? ctr + 6 : nullptr, list_b, replicate, flip_player, mode, offsets,
                            max_rows, after, after_cap_rows, row_players, row_feats, counts_true, counts, starts, alloc, status,
                            forked ? ctr + 7 : ctr + 4, stream);
}
