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
#include "bg_internal.h"

namespace bg {

constexpr uint32_t kEmpty = 0xFFFFFFFFu;

__device__ __constant__ int8_t kRoll21[21][2] = {{1, 1}, {1, 2}, {1, 3}, {1, 4}, {1, 5}, {1, 6}, {2, 2}, {2, 3}, {2, 4}, {2, 5}, {2, 6},
                                                {3, 3}, {3, 4}, {3, 5}, {3, 6}, {4, 4}, {4, 5}, {4, 6}, {5, 5}, {5, 6}, {6, 6}};

template <int CAP, int HS>
struct WarpScratch {
    uint32_t list[2][6][CAP];   // two level lists, SoA: lo.lo lo.hi hi.lo hi.hi hit occ
    uint32_t pm[CAP];           // per parent: move mask | (special+1) << 24
    uint32_t hash[HS];          // open-addressing set of indices into the destination list
    uint16_t off[CAP + 2];      // exclusive prefix of per-parent move counts
    uint32_t rootw[kBoardWords];
    uint32_t pad;
};

template <int CAP, int HS>
struct Warp {
    WarpScratch<CAP, HS>& S;
    Root R;
    int lane;
    bool overflow;

    __device__ Warp(WarpScratch<CAP, HS>& s, int l) : S(s), lane(l), overflow(false) {}

    __device__ __forceinline__ Node load(int b, int i) const {
        Node n;
        n.lo = (unsigned long long)S.list[b][0][i] | ((unsigned long long)S.list[b][1][i] << 32);
        n.hi = (unsigned long long)S.list[b][2][i] | ((unsigned long long)S.list[b][3][i] << 32);
        n.hit = S.list[b][4][i];
        n.occ = S.list[b][5][i];
        return n;
    }
    __device__ __forceinline__ void store(int b, int i, const Node& n) {
        S.list[b][0][i] = (uint32_t)n.lo; S.list[b][1][i] = (uint32_t)(n.lo >> 32);
        S.list[b][2][i] = (uint32_t)n.hi; S.list[b][3][i] = (uint32_t)(n.hi >> 32);
        S.list[b][4][i] = n.hit; S.list[b][5][i] = n.occ;
    }
    __device__ __forceinline__ bool same(int b, int i, const Node& n) const {
        return S.list[b][0][i] == (uint32_t)n.lo && S.list[b][1][i] == (uint32_t)(n.lo >> 32) &&
               S.list[b][2][i] == (uint32_t)n.hi && S.list[b][3][i] == (uint32_t)(n.hi >> 32) &&
               S.list[b][4][i] == n.hit;
    }
    __device__ void clear_hash() {
        for (int i = lane; i < HS; i += 32) S.hash[i] = kEmpty;
        __syncwarp();
    }
    // true if a board equal to n is already in destination list b (via the hash set)
    __device__ __forceinline__ bool in_set(int b, const Node& n) const {
        uint32_t s = hash_node(n) & (HS - 1);
        for (;;) {
            uint32_t e = S.hash[s];
            if (e == kEmpty) return false;
            if (same(b, (int)e, n)) return true;
            s = (s + 1) & (HS - 1);
        }
    }
    __device__ __forceinline__ void set_insert(const Node& n, int pos) {
        uint32_t s = hash_node(n) & (HS - 1);
        while (atomicCAS(&S.hash[s], kEmpty, (uint32_t)pos) != kEmpty) s = (s + 1) & (HS - 1);
    }

    // Count the one-die moves of every parent in list pb[0..np); fills S.pm / S.off. Returns the total.
    __device__ int count_moves(int pb, int np, int d) {
        int base = 0;
        for (int i0 = 0; i0 < np; i0 += 32) {
            int i = i0 + lane;
            int cnt = 0;
            if (i < np) {
                Node n = load(pb, i);
                uint32_t mask; int special;
                one_die(n, R, d, mask, special);
                cnt = __popc(mask) + (special >= 0);
                S.pm[i] = mask | ((uint32_t)(special + 1) << 24);
            }
            int inc = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int v = __shfl_up_sync(kFull, inc, o);
                if (lane >= o) inc += v;
            }
            if (i < np) S.off[i] = (uint16_t)(base + inc - cnt);
            base += __shfl_sync(kFull, inc, 31);
        }
        if (lane == 0) S.off[np] = (uint16_t)base;
        __syncwarp();
        return base;
    }

    // Append `keep` lanes' nodes to list cb at nc in lane order; returns the new nc (or sets overflow).
    template <bool INSERT>
    __device__ __forceinline__ int append(int cb, int nc, bool keep, const Node& ch) {
        unsigned surv = __ballot_sync(kFull, keep);
        int nsurv = __popc(surv);
        if (nc + nsurv > CAP) { overflow = true; return nc; }
        if (keep) {
            int pos = nc + __popc(surv & ((1u << lane) - 1u));
            store(cb, pos, ch);
            if (INSERT) set_insert(ch, pos);
        }
        __syncwarp();
        return nc + nsurv;
    }

    // Append the children (die d) of parents pb[0..np) to list cb starting at nc, in reference
    // order; children equal to an entry already in the set of cb or to an earlier child are
    // dropped.  count_moves(pb, np, d) must have been called (total = its result).
    __device__ int expand(int pb, int np, int d, int total, int cb, int nc) {
        for (int c0 = 0; c0 < total && !overflow; c0 += 32) {
            int idx = c0 + lane;
            bool valid = idx < total;
            Node ch;
            ch.lo = ~0ull - (unsigned long long)lane; ch.hi = 0; ch.hit = 0; ch.occ = 0;   // impossible board
            if (valid) {
                int lo = 0, hi = np - 1;                       // largest parent with off[parent] <= idx
                while (lo < hi) {
                    int mid = (lo + hi + 1) >> 1;
                    if ((int)S.off[mid] <= idx) lo = mid; else hi = mid - 1;
                }
                int j = idx - (int)S.off[lo];
                uint32_t pmv = S.pm[lo];
                Node p = load(pb, lo);
                ch = apply_move(p, R, d, pmv & 0xFFFFFFu, (int)(pmv >> 24) - 1, j);
            }
            unsigned m = __match_any_sync(kFull, ch.lo) & __match_any_sync(kFull, ch.hi) &
                         __match_any_sync(kFull, ch.hit);
            bool keep = valid && (__ffs(m) - 1 == lane);
            if (keep && in_set(cb, ch)) keep = false;
            nc = append<true>(cb, nc, keep, ch);
        }
        return nc;
    }

    // Level-1 boards of `root` for die d -> list b[0..n) (distinct sources => distinct boards). Returns n <= 16.
    __device__ int first_level(const Node& root, int d, int b) {
        uint32_t mask; int special;
        one_die(root, R, d, mask, special);
        int n = __popc(mask) + (special >= 0);
        if (lane < n) store(b, lane, apply_move(root, R, d, mask, special, lane));
        __syncwarp();
        return n;
    }

    // Full generator.  On return the legal afterstates are list `ob`[from, from+n), reference order.
    __device__ void generate(const Node& root, int d0, int d1, int& ob, int& from, int& n) {
        ob = 0; from = 0; n = 0;
        if (d0 != d1) {
            const int hi = max(d0, d1), lo = min(d0, d1);          // get_all_moves.py:30
            clear_hash();
            int nF = 0;      // plays collected in list 0 (full_moves)
            int nA1 = 0;     // leading plays of length 1 contributed by pass A
            bool lenA2 = false, lenB2 = false;
            // ---- pass A: larger die first (handle_moves.py:109-200, reverse=False)
            int nA = first_level(root, hi, 1);
            if (nA) {
                int t2 = count_moves(1, nA, lo);                   // two_move_sequences_exist, :145-155
                if (t2) { nF = expand(1, nA, lo, t2, 0, 0); lenA2 = true; }
                else {                                             // singles are the plays, :192-200
                    Node c = load(1, lane < nA ? lane : 0);
                    nF = append<true>(0, 0, lane < nA, c);
                    nA1 = nA;
                    if (nA == 1) { n = 1; return; }                // skip-reverse shortcut, get_all_moves.py:43-45
                }
                if (overflow) return;
            }
            // ---- pass B: smaller die first (reverse=True), sharing full_moves / unique_boards
            int nB = first_level(root, lo, 1);
            if (nB) {
                int t2 = count_moves(1, nB, hi);
                if (t2) { nF = expand(1, nB, hi, t2, 0, nF); lenB2 = true; }
                else if (!lenA2) {                                 // only singles anywhere: union (add_unique_board)
                    Node c = load(1, lane < nB ? lane : 0);
                    bool keep = lane < nB && !in_set(0, c);
                    nF = append<true>(0, nF, keep, c);
                }
                // (singles of pass B next to length-2 plays of pass A are removed by the max filter
                //  and, being last, influence nobody's dedupe: not materialised.)
                if (overflow) return;
            }
            // ---- filter_full_moves_by_max_submoves (get_all_moves.py:73-94), applied AFTER dedupe
            if (lenB2 && !lenA2) { from = nA1; n = nF - nA1; }     // length-1 plays of pass A are dropped
            else { from = 0; n = nF; }
        } else {
            // ---- doubles (handle_moves.py:203-310): levels 1..4, output = deepest non-empty level
            const int d = d0;
            if (lane == 0) store(0, 0, root);
            __syncwarp();
            int pb = 0, np = 1;
            for (int depth = 1; depth <= 4; ++depth) {
                int t = count_moves(pb, np, d);
                if (t == 0) break;
                clear_hash();
                int nc = expand(pb, np, d, t, pb ^ 1, 0);
                if (overflow) return;
                pb ^= 1; np = nc;
                ob = pb; n = np;
            }
        }
    }
};

// mode: 0 = count only; 1 = write rows at offsets[b]; 2 = slab (atomicAdd on *alloc, writes starts[b])
template <int CAP, int HS>
__global__ void __launch_bounds__(256) movegen_kernel(
    const int8_t* __restrict__ boards, const int8_t* __restrict__ players, const int8_t* __restrict__ dice,
    long long B, const unsigned int* __restrict__ nwork_dev, const int32_t* __restrict__ worklist,
    int replicate, int flip_player, int mode, const long long* __restrict__ offsets, int max_rows,
    int8_t* __restrict__ after, long long after_cap_rows, int8_t* __restrict__ row_players,
    int32_t* __restrict__ counts_true, int32_t* __restrict__ counts, long long* __restrict__ starts,
    unsigned long long* __restrict__ alloc, int32_t* __restrict__ status,
    unsigned int* __restrict__ work_ctr, int32_t* __restrict__ overflow_list, unsigned int* __restrict__ overflow_ctr) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    WarpScratch<CAP, HS>& S = reinterpret_cast<WarpScratch<CAP, HS>*>(smem_raw)[warp];
    const long long nwork = nwork_dev ? (long long)*nwork_dev : B;

    for (;;) {
        unsigned int wi = 0;
        if (lane == 0) wi = atomicAdd(work_ctr, 1u);
        wi = __shfl_sync(kFull, wi, 0);
        if ((long long)wi >= nwork) break;
        const long long g = worklist ? (long long)worklist[wi] : (long long)wi;

        // ---- load the root (13 words, coalesced) and build the mover-relative view
        // replicate == 21: work item g is (position g / 21, sorted roll g % 21) -- the 2-ply opponent expansion
        // in the roll order of get_all_dice_rolls_tensor (moves/get_all_dice_rolls.py:19-32)
        const long long src = replicate > 1 ? g / replicate : g;
        const uint32_t* bw = reinterpret_cast<const uint32_t*>(boards + src * kBoardBytes);
        uint32_t w = lane < kBoardWords ? bw[lane] : 0u;
        if (lane < kBoardWords) S.rootw[lane] = w;
        const int player = (players[src] ^ flip_player) & 1;
        int d0, d1;
        if (replicate > 1) { const int r = (int)(g - src * replicate); d0 = kRoll21[r][0]; d1 = kRoll21[r][1]; }
        else { d0 = dice[2 * g]; d1 = dice[2 * g + 1]; }
        const int p = lane < 24 ? lane : 0;
        uint32_t ownw = __shfl_sync(kFull, w, (player ? 6 : 0) + (p >> 2));
        uint32_t oppw = __shfl_sync(kFull, w, (player ? 0 : 6) + (p >> 2));
        uint32_t misc = __shfl_sync(kFull, w, 12);
        int ownc = lane < 24 ? (int)((ownw >> (8 * (p & 3))) & 0xFFu) : 0;
        int oppc = lane < 24 ? (int)((oppw >> (8 * (p & 3))) & 0xFFu) : 0;
        int ownbar = (int)((misc >> (player ? 8 : 0)) & 0xFFu), ownoff = (int)((misc >> (player ? 24 : 16)) & 0xFFu);
        Warp<CAP, HS> W(S, lane);
        W.R.player = player;
        W.R.block = __ballot_sync(kFull, oppc >= 2) & 0xFFFFFFu;
        W.R.blot = __ballot_sync(kFull, oppc == 1) & 0xFFFFFFu;
        Node root;
        root.occ = __ballot_sync(kFull, ownc > 0) & 0xFFFFFFu;
        root.hit = 0;
        uint32_t nib = (uint32_t)(ownc & 15) << (4 * (p & 7));
        uint32_t w0 = __reduce_or_sync(kFull, (lane < 8) ? nib : 0u);
        uint32_t w1 = __reduce_or_sync(kFull, (lane >= 8 && lane < 16) ? nib : 0u);
        uint32_t w2 = __reduce_or_sync(kFull, (lane >= 16 && lane < 24) ? nib : 0u);
        root.lo = (unsigned long long)w0 | ((unsigned long long)w1 << 32);
        root.hi = (unsigned long long)w2 | ((unsigned long long)((ownbar & 15) | ((ownoff & 15) << 4)) << 32);
        int total = __reduce_add_sync(kFull, ownc) + ownbar + ownoff;
        W.R.tot15 = total == 15;
        bool bad = __any_sync(kFull, ownc > 15 || oppc > 15) || ownbar > 15 || ownoff > 15 ||
                   d0 < 1 || d0 > 6 || d1 < 1 || d1 > 6;
        __syncwarp();

        int ob = 0, from = 0, n = 0;
        if (!bad) W.generate(root, d0, d1, ob, from, n);

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
                if (overflow_list) { unsigned int k = atomicAdd(overflow_ctr, 1u); overflow_list[k] = (int32_t)g; }
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
        long long start = 0;
        if (mode == 1) start = offsets[g];
        else if (mode == 2) {
            unsigned long long s0 = 0;
            if (lane == 0) s0 = atomicAdd(alloc, (unsigned long long)nw);
            start = (long long)__shfl_sync(kFull, s0, 0);
        }
        if (lane == 0) {
            if (counts_true) counts_true[g] = n;
            if (counts) counts[g] = nw;
            if (mode == 2 && starts) starts[g] = start;
        }
        if (mode != 0 && nw > 0) {
            if (start + nw > after_cap_rows) {
                if (lane == 0) { atomicOr(status, BG_STATUS_OUTPUT_OVERFLOW); if (counts) counts[g] = 0; }
            } else {
                // stage 32 rows x 13 words in the unused list, then copy out fully coalesced
                uint32_t* stage = &S.list[ob ^ 1][0][0];
                uint32_t* gout = reinterpret_cast<uint32_t*>(after) + start * kBoardWords;
                for (int r0 = 0; r0 < nw; r0 += 32) {
                    int r = r0 + lane;
                    if (r < nw) {
                        Node nd = W.load(ob, from + r);
#pragma unroll
                        for (int k = 0; k < kBoardWords; ++k)
                            stage[lane * kBoardWords + k] = node_row_word(nd, player, S.rootw, k);
                    }
                    __syncwarp();
                    int rows = min(32, nw - r0);
                    for (int k = lane; k < rows * kBoardWords; k += 32) gout[(long long)r0 * kBoardWords + k] = stage[k];
                    if (row_players && lane < rows) row_players[start + r0 + lane] = (int8_t)player;
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
                          int8_t* row_players, int32_t* counts_true, int32_t* counts, long long* starts, unsigned long long* alloc,
                          int32_t* status, unsigned int* work_ctr, int32_t* overflow_list, unsigned int* overflow_ctr,
                          long long grid_hint, cudaStream_t stream) {
    size_t smem = sizeof(WarpScratch<CAP, HS>) * WARPS;
    auto kern = movegen_kernel<CAP, HS>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return bg_set_error(e, "movegen: cudaFuncSetAttribute");
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, WARPS * 32, smem);
    if (occ < 1) occ = 1;
    long long grid = (long long)bg_sm_count() * occ;
    if (grid_hint > 0 && grid > grid_hint) grid = grid_hint;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, WARPS * 32, smem, stream>>>(boards, players, dice, B, nwork_dev, worklist, replicate, flip_player, mode, offsets,
                                                       max_rows, after, after_cap_rows, row_players, counts_true, counts,
                                                       starts, alloc, status, work_ctr, overflow_list, overflow_ctr);
    return bg_set_error(cudaGetLastError(), "movegen: launch");
}

}  // namespace bg

using namespace bg;

// Workspace layout (bytes): [0] work_ctr u32, [4] overflow_ctr u32, [8] work_ctr2 u32, [64..] overflow_list int32[B]
extern "C" size_t bg_movegen_workspace_bytes(long long B) { return 64 + sizeof(int32_t) * (size_t)(B > 0 ? B : 1); }

int bg::movegen_run(const int8_t* boards, const int8_t* players, const int8_t* dice, long long B, int replicate,
                    int flip_player, int mode,
                    const long long* offsets, int max_rows, int8_t* after, long long after_cap_rows,
                    int8_t* row_players, int32_t* counts_true, int32_t* counts, long long* starts,
                    unsigned long long* alloc, int32_t* status, void* workspace, size_t ws_bytes, cudaStream_t stream) {
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
    unsigned int* work_ctr = reinterpret_cast<unsigned int*>(ws);
    unsigned int* overflow_ctr = reinterpret_cast<unsigned int*>(ws + 4);
    unsigned int* work_ctr2 = reinterpret_cast<unsigned int*>(ws + 8);
    int32_t* overflow_list = reinterpret_cast<int32_t*>(ws + 64);
    cudaError_t e = cudaMemsetAsync(ws, 0, 64, stream);
    if (e != cudaSuccess) return bg_set_error(e, "movegen: memset");
    int rc = launch_movegen<BG_MOVEGEN_CAP_SMALL, 2 * BG_MOVEGEN_CAP_SMALL, 8>(
        boards, players, dice, B, nullptr, nullptr, replicate, flip_player, mode, offsets, max_rows, after,
        after_cap_rows, row_players,
        counts_true, counts, starts, alloc, status, work_ctr, overflow_list, overflow_ctr, (B + 7) / 8, stream);
    if (rc != BG_OK) return rc;
    // Large-scratch pass over the (rare) positions whose levels did not fit: one warp per CTA, work count read
    // from device memory so no host synchronisation is needed.  Positions that do not fit even this scratch
    // raise BG_STATUS_SCRATCH_OVERFLOW (never dropped silently).
    return launch_movegen<BG_MOVEGEN_CAP_BIG, BG_MOVEGEN_HASH_BIG, 1>(
        boards, players, dice, B, overflow_ctr, overflow_list, replicate, flip_player, mode, offsets, max_rows, after,
        after_cap_rows,
        row_players, counts_true, counts, starts, alloc, status, work_ctr2, nullptr, nullptr, 0, stream);
}
