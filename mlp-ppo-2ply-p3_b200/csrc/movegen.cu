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
#include "bg_features.cuh"
#include "bg_internal.h"

namespace bg {

constexpr uint32_t kEmpty = 0xFFFFFFFFu;

// Per-warp scratch.  Two level lists ("regions") of CAP boards each plus one slot for the root:
// node i of region r lives at index r*CAP + i, the root at 2*CAP.  A board is a 16-byte key
// (x = points 0..7, y = points 8..15, z = points 16..23 as nibbles, w = hit mask | bar << 24 | off << 28)
// plus its occupancy mask, so loads/stores are one LDS.128/STS.128 + one 32-bit access per lane.
template <int CAP, int HS>
struct WarpScratch {
    uint4 key[2 * CAP + 1];
    uint32_t occ[2 * CAP + 1];
    uint32_t pm[CAP];           // per parent: move mask | (special+1) << 24
    uint32_t hash[HS];          // open-addressing set of node indices of the destination region
    uint16_t off[CAP + 2];      // exclusive prefix of per-parent move counts
    uint32_t rootw[kBoardWords];
    uint32_t pad[2];
};

template <int CAP, int HS>
struct Warp {
    static_assert(CAP * 16 >= 32 * kBoardWords * 4, "a region must be able to stage 32 output rows");
    static constexpr int kRoot = 2 * CAP;
    WarpScratch<CAP, HS>& S;
    Root R;
    int lane;
    bool overflow;

    __device__ Warp(WarpScratch<CAP, HS>& s, int l) : S(s), lane(l), overflow(false) {}

    __device__ __forceinline__ Node load(int i) const {
        uint4 k = S.key[i];
        Node n;
        n.lo = (unsigned long long)k.x | ((unsigned long long)k.y << 32);
        n.hi = (unsigned long long)k.z | ((unsigned long long)(k.w >> 24) << 32);
        n.hit = k.w & 0xFFFFFFu;
        const uint32_t o = S.occ[i];                                 // occupancy | last source << 24
        n.occ = o & 0xFFFFFFu;
        n.last = o >> 24;
        return n;
    }
    static __device__ __forceinline__ uint4 key_of(const Node& n) {
        return make_uint4((uint32_t)n.lo, (uint32_t)(n.lo >> 32), (uint32_t)n.hi, n.hit | ((uint32_t)(n.hi >> 32) << 24));
    }
    __device__ __forceinline__ void store(int i, const Node& n) { S.key[i] = key_of(n); S.occ[i] = n.occ | (n.last << 24); }
    __device__ __forceinline__ bool same(int i, const uint4& k) const {
        uint4 e = S.key[i];
        return e.x == k.x && e.y == k.y && e.z == k.z && e.w == k.w;
    }
    static __device__ __forceinline__ uint32_t hash_key(const uint4& k) {
        uint32_t h = k.x * 0x9E3779B1u ^ k.y * 0x85EBCA77u ^ k.z * 0xC2B2AE3Du ^ k.w * 0x27D4EB2Fu;
        return h ^ (h >> 15);
    }
    __device__ void clear_hash() {
        for (int i = lane; i < HS; i += 32) S.hash[i] = kEmpty;
        __syncwarp();
    }
    __device__ __forceinline__ bool in_set(const uint4& k) const {
        uint32_t s = hash_key(k) & (HS - 1);
        for (;;) {
            uint32_t e = S.hash[s];
            if (e == kEmpty) return false;
            if (same((int)e, k)) return true;
            s = (s + 1) & (HS - 1);
        }
    }
    __device__ __forceinline__ void set_insert(const uint4& k, int idx) {
        uint32_t s = hash_key(k) & (HS - 1);
        while (atomicCAS(&S.hash[s], kEmpty, (uint32_t)idx) != kEmpty) s = (s + 1) & (HS - 1);
    }

    // Count the one-die moves of the parents at [pbase, pbase+np); parent i plays die (i < split ? dA : dB).
    // Fills S.pm / S.off. Returns the total.
    __device__ __forceinline__ int count_moves(int pbase, int np, int split, int dA, int dB) {
        int base = 0;
        for (int i0 = 0; i0 < np; i0 += 32) {
            int i = i0 + lane;
            int cnt = 0;
            if (i < np) {
                Node n = load(pbase + i);
                uint32_t mask; int special;
                one_die(n, R, i < split ? dA : dB, mask, special);
                mask = prune_mask(mask, n, R, dA, dA == dB, i >= split);     // drop provably duplicate candidates
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

    // Append `keep` lanes' nodes to the region at cbase (currently nc entries) in lane order.
    __device__ __forceinline__ int append(int cbase, int nc, bool keep, const Node& ch, bool insert) {
        unsigned surv = __ballot_sync(kFull, keep);
        int nsurv = __popc(surv);
        if (nc + nsurv > CAP) { overflow = true; return nc; }
        if (keep) {
            int idx = cbase + nc + __popc(surv & ((1u << lane) - 1u));
            store(idx, ch);
            if (insert) set_insert(key_of(ch), idx);
        }
        __syncwarp();
        return nc + nsurv;
    }

    // Children of the parents at [pbase, pbase+np) (parent i plays die i < split ? dA : dB), candidates
    // [cfrom, total) of the count_moves numbering, in the reference's order, appended to the region at cbase
    // (nc entries so far).  use_set: drop children equal to an entry of the set or to an earlier child.
    // use_hash: also consult / feed the hash set (needed only when the stage has more than one chunk of candidates
    // or entries appended before it; a single chunk is deduped by match_any alone).
    __device__ __forceinline__ int expand(int pbase, int np, int split, int dA, int dB, int cfrom, int total, int cbase,
                                          int nc, bool use_set, bool use_hash) {
        for (int c0 = cfrom; c0 < total && !overflow; c0 += 32) {
            int idx = c0 + lane;
            bool valid = idx < total;
            Node ch;
            ch.lo = ~0ull - (unsigned long long)lane; ch.hi = 0; ch.hit = 0; ch.occ = 0; ch.last = 31u;   // impossible board
            if (valid) {
                int lo = 0, hi = np - 1;                       // largest parent with off[parent] <= idx
                while (lo < hi) {
                    int mid = (lo + hi + 1) >> 1;
                    if ((int)S.off[mid] <= idx) lo = mid; else hi = mid - 1;
                }
                int j = idx - (int)S.off[lo];
                uint32_t pmv = S.pm[lo];
                Node p = load(pbase + lo);
                ch = apply_move(p, R, lo < split ? dA : dB, pmv & 0xFFFFFFu, (int)(pmv >> 24) - 1, j);
            }
            bool keep = valid;
            if (use_set) {
                uint4 k = key_of(ch);
                unsigned m = __match_any_sync(kFull, ch.lo) &
                             __match_any_sync(kFull, (unsigned long long)k.z | ((unsigned long long)k.w << 32));
                keep = valid && (__ffs(m) - 1 == lane);
                if (use_hash && keep && in_set(k)) keep = false;
            }
            nc = append(cbase, nc, keep, ch, use_hash);
        }
        return nc;
    }

    // Full generator.  On return the legal afterstates are nodes [obase, obase+n), reference order.
    // count_moves/expand are instantiated once (code size: the kernel must stay inside the instruction cache).
    //
    // Non-doubles (handle_moves.py:109-200, get_all_moves.py:28-56): the two die orders are expanded TOGETHER.
    // Region 1 holds both first levels -- the larger-die-first boards [0,nA), then the smaller-die-first boards
    // [nA,nA+nB) (<= 16 each) -- one count_moves gives every first-level board its second-move count with the
    // OTHER die, and one expand sweeps the candidates of pass A, then those of pass B, in order: exactly the
    // reference's insertion order into full_moves, with its first-wins dedupe, in half the stages.
    __device__ void generate(const Node& root, int d0, int d1, int& obase, int& n) {
        obase = 0; n = 0;
        const bool dbl = d0 == d1;
        const int dhi = max(d0, d1), dlo = min(d0, d1);            // get_all_moves.py:30
        int pbase, np, split, dA, dB, nstages;
        int nA = 0;
        if (!dbl) {
            uint32_t mA, mB; int sA, sB;
            one_die(root, R, dhi, mA, sA);
            one_die(root, R, dlo, mB, sB);
            R.mA = mA;
            nA = __popc(mA) + (sA >= 0);
            const int nB = __popc(mB) + (sB >= 0);
            if (lane < nA) store(CAP + lane, apply_move(root, R, dhi, mA, sA, lane));
            else if (lane < nA + nB) store(CAP + lane, apply_move(root, R, dlo, mB, sB, lane - nA));
            __syncwarp();
            if (nA + nB == 0) return;
            pbase = CAP; np = nA + nB; split = nA; dA = dlo; dB = dhi; nstages = 1;
        } else {
            if (lane == 0) store(kRoot, root);
            __syncwarp();
            pbase = kRoot; np = 1; split = 0x7FFFFFFF; dA = dB = d0; nstages = 4;
        }
        for (int stage = 0; stage < nstages; ++stage) {
            const int total = count_moves(pbase, np, split, dA, dB);
            int cbase = 0, cfrom = 0, cto = total, nc0 = 0;
            bool do_expand = true, use_hash = true;
            int nA1 = 0;
            bool lenB2_only = false;
            if (dbl) {
                // doubles (handle_moves.py:203-310): stage k expands level k -> k+1, regions alternate
                if (total == 0) break;                             // dead end: the previous level is the answer
                cbase = (stage & 1) ? CAP : 0;
                use_hash = total > 32;                             // one chunk: match_any alone dedupes it
                if (use_hash) clear_hash();
            } else {
                const int tA = (int)S.off[nA];                     // two-move candidates of the larger-die-first pass
                const int tB = total - tA;
                if (tA > 0) {                                      // two_move_sequences_exist (:145-155) in pass A
                    if (tB == 0) cto = tA;                         // pass-B singles vanish in the max filter
                    use_hash = cto > 32;
                    if (use_hash) clear_hash();
                } else {
                    // pass A has only singles (:192-200): they are the first plays, in order.  (They stay in the
                    // set while pass B's two-move boards are added, as in add_unique_board; a one-move board can
                    // never equal a two-move board, but the dedupe is kept literal.)
                    clear_hash();
                    Node c = load(CAP + (lane < nA ? lane : 0));
                    nc0 = append(0, 0, lane < nA, c, true);
                    if (nA == 1) { n = 1; return; }                // skip-reverse shortcut, get_all_moves.py:43-45
                    nA1 = nA;
                    if (tB > 0) lenB2_only = true;                 // length-1 plays of pass A are dropped by the filter
                    else {                                         // only singles anywhere: union (add_unique_board)
                        do_expand = false;
                        const int nB = np - nA;
                        Node c2 = load(CAP + nA + (lane < nB ? lane : 0));
                        bool keep = lane < nB && !in_set(key_of(c2));
                        nc0 = append(0, nc0, keep, c2, true);
                    }
                }
                cfrom = tA > 0 ? 0 : tA;                           // (= 0 either way: A parents own no candidates if tA == 0)
            }
            int nc = nc0;
            if (do_expand) nc = expand(pbase, np, split, dA, dB, cfrom, cto, cbase, nc0, true, use_hash);   // the only call site
            if (overflow) return;
            if (dbl) { pbase = cbase; np = nc; obase = cbase; n = nc; }
            else {
                // filter_full_moves_by_max_submoves (get_all_moves.py:73-94), applied AFTER dedupe
                if (lenB2_only) { obase = nA1; n = nc - nA1; } else { obase = 0; n = nc; }
            }
        }
    }
};

// mode: 0 = count only; 1 = write rows at offsets[b]; 2 = slab (atomicAdd on *alloc, writes starts[b])
// FEATS: also write the bf16 feature row of every afterstate (optional fused K3; a separate instantiation so that
// the common one stays inside the instruction cache)
template <int CAP, int HS, bool FEATS>
__global__ void __launch_bounds__(256) movegen_kernel(
    const int8_t* __restrict__ boards, const int8_t* __restrict__ players, const int8_t* __restrict__ dice,
    long long B, const unsigned int* __restrict__ nwork_dev, const int32_t* __restrict__ worklist,
    int replicate, int flip_player, int mode, const long long* __restrict__ offsets, int max_rows,
    int8_t* __restrict__ after, long long after_cap_rows, int8_t* __restrict__ row_players,
    uint16_t* __restrict__ row_feats, int32_t* __restrict__ counts_true, int32_t* __restrict__ counts, long long* __restrict__ starts,
    unsigned long long* __restrict__ alloc, int32_t* __restrict__ status,
    unsigned int* __restrict__ work_ctr, int32_t* __restrict__ overflow_list, unsigned int* __restrict__ overflow_ctr) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    WarpScratch<CAP, HS>& S = reinterpret_cast<WarpScratch<CAP, HS>*>(smem_raw)[warp];
    __shared__ uint32_t s_lut[32], s_desc[32];          // K3's chunk tables, for the fused feature output
    if (FEATS) { load_chunk_tables(s_lut, s_desc); __syncthreads(); }
    const long long nwork = nwork_dev ? (long long)*nwork_dev : B;

    for (;;) {
        unsigned int wi = 0;
        if (lane == 0) wi = atomicAdd(work_ctr, 1u);
        wi = __shfl_sync(kFull, wi, 0);
        if ((long long)wi >= nwork) break;
        const long long g = worklist ? (long long)worklist[wi] : (long long)wi;

        // ---- the work item, its root (13 words, coalesced) and the mover-relative view (bg_movegen_common.cuh)
        const long long src = replicate > 1 ? g / replicate : g;
        const uint32_t bword = load_board_word(boards, src, lane);
        const WorkItem item = decode_work_item(g, src, replicate, flip_player, players, dice);
        const int player = item.player, d0 = item.d0, d1 = item.d1;
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
                // stage 32 rows x 13 words in the region that does not hold the result, then copy out coalesced
                uint32_t* stage = reinterpret_cast<uint32_t*>(&S.key[obase < CAP ? CAP : 0]);
                uint32_t* gout = reinterpret_cast<uint32_t*>(after) + start * kBoardWords;
                const RowContext rc = make_row_context(player, S.rootw);
                for (int r0 = 0; r0 < nw; r0 += 32) {
                    int r = r0 + lane;
                    if (r < nw) {
                        build_row(S.key[obase + r], player, rc, S.rootw, stage + lane * kBoardWords);
                    }
                    __syncwarp();
                    int rows = min(32, nw - r0);
                    for (int k2 = lane; k2 < rows * kBoardWords; k2 += 32) gout[(long long)r0 * kBoardWords + k2] = stage[k2];
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
                          long long grid_hint, cudaStream_t stream) {
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
                                                       counts_true, counts, starts, alloc, status, work_ctr, overflow_list, overflow_ctr);
    return bg_set_error(cudaGetLastError(), "movegen: launch");
}

}  // namespace bg

using namespace bg;

// Workspace layout (bytes; constants in bg_internal.h): [0] work_ctr0, [4] overflow_ctr A, [8] work_ctr1, [12] overflow_ctr B,
// [16] work_ctr2, [BG_WS_ROWS_AFTER_TIER0 = 40] u64 row-count snapshot after tier 0 (written by callers that fork there),
// [BG_WS_LISTS = 64 ..] overflow list A int32[B], then overflow list B int32[B]
extern "C" size_t bg_movegen_workspace_bytes(long long B) { return BG_WS_LISTS + 2 * sizeof(int32_t) * (size_t)(B > 0 ? B : 1); }

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
    int rc = launch_movegen<BG_MOVEGEN_CAP_SMALL, 2 * BG_MOVEGEN_CAP_SMALL, 8>(
        boards, players, dice, B, nullptr, nullptr, replicate, flip_player, mode, offsets, max_rows, after,
        after_cap_rows, row_players, row_feats, counts_true, counts, starts, alloc, status, ctr + 0, list_a, ctr + 1,
        (B + 7) / 8, stream);
    if (rc != BG_OK) return rc;
    // Slab mode: rows [0, *alloc) are final once tier 0 is done (tiers 1/2 only append).  A caller that wants to
    // consume them while the latency-bound overflow tiers run gets the row count snapshot and a fork point here.
    if (hook && mode == 2) {
        e = cudaMemcpyAsync(hook->rows_after_tier0, alloc, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, stream);
        if (e != cudaSuccess) return bg_set_error(e, "movegen: snapshot");
        rc = hook->fn(hook->user);
        if (rc != BG_OK) return rc;
    }
    // Tier 1: the (~1 %) positions whose levels did not fit: one CTA per position (movegen_team.cu),
    // BG_MOVEGEN_CAP_MID boards per level.  Work counts of tiers 1 and 2 are read from device memory, so no host
    // synchronisation is needed.
    rc = movegen_team_mid(boards, players, dice, ctr + 1, list_a, replicate, flip_player, mode, offsets, max_rows, after,
                          after_cap_rows, row_players, row_feats, counts_true, counts, starts, alloc, status, ctr + 2, list_b,
                          ctr + 3, stream, ((hook && mode == 2) || B >= 262144) ? 128 : 0);   // 128 threads per position when something shares the GPU or the list is long (throughput, not latency)
    if (rc != BG_OK) return rc;
    // Tier 2: the rest (> BG_MOVEGEN_CAP_MID boards in a level): one big CTA per position, BG_MOVEGEN_CAP_BIG
    // boards per level.  Positions that do not fit even this raise BG_STATUS_SCRATCH_OVERFLOW (never dropped silently).
    return movegen_team_big(boards, players, dice, ctr + 3, list_b, replicate, flip_player, mode, offsets, max_rows, after,
                            after_cap_rows, row_players, row_feats, counts_true, counts, starts, alloc, status, ctr + 4, stream);
}
