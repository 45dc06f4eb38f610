// movegen_team.cu -- K1 for the heavy tail: one CTA ("team" of T threads) per position.
//
// Same algorithm and same output as movegen.cu (level-by-level expansion in the reference's order, see the
// header of that file), for the ~1 % of positions whose levels hold more than BG_MOVEGEN_CAP_SMALL boards
// (big doubles: up to 1,654 plays in the fixtures).  A lone warp spends ~0.3 us per generated board on such a
// position with nothing to hide its latencies behind, which made the overflow pass as long as the main pass;
// here T candidates are built per iteration and the whole CTA shares the level lists.
//
// Ordered first-wins dedupe across warps (match_any only sees one warp): every candidate of a chunk publishes
// its key, then probes the hash set; an empty slot is claimed tentatively with PEND|tid, a slot pending for an
// EQUAL key is taken over with atomicMin (lowest candidate index wins), a committed entry with an equal key kills
// the candidate.  After a barrier a candidate survives iff its slot still names it; survivors are compacted in
// candidate order (ballot + per-warp counts) and commit their slots to list indices.
#include <cstdlib>
#include "bg_device.cuh"
#include "bg_movegen_common.cuh"
#include "bg_features.cuh"
#include "bg_internal.h"

namespace bg {

namespace {
constexpr uint32_t kEmptyT = 0xFFFFFFFFu;
constexpr uint32_t kPend = 0x80000000u;

template <int CAP, int HS, int T>
struct TeamScratch {
    uint4 key[2 * CAP + 1];
    uint4 ckey[T];              // keys of the current chunk's candidates
    uint32_t occ[2 * CAP + 1];
    uint32_t pm[CAP];
    uint32_t hash[HS];
    uint16_t off[CAP + 2];
    uint32_t rootw[kBoardWords];
    int warp_cnt[2][T / 32];     // alternating buffers: one barrier per team_scan
    Node root;
    Root R;
    int bcast[4];
};

template <int CAP, int HS, int T>
struct Team {
    // output rows staged per iteration: what one region (CAP keys of 16 B) can hold, at most one per thread
    static constexpr int kStageRows = (CAP * 16 / kBoardBytes) / 32 * 32 < T ? (CAP * 16 / kBoardBytes) / 32 * 32 : T;
    static_assert(kStageRows >= 32, "a region must be able to stage at least 32 output rows");
    static constexpr int kRoot = 2 * CAP;
    static constexpr int kWarps = T / 32;
    TeamScratch<CAP, HS, T>& S;
    Root R;
    int tid, lane, warp;
    int scan_buf;
    bool overflow;

    __device__ Team(TeamScratch<CAP, HS, T>& s) : S(s), tid(threadIdx.x), lane(threadIdx.x & 31), warp(threadIdx.x >> 5), scan_buf(0), overflow(false) {}

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
    static __device__ __forceinline__ bool eq(const uint4& a, const uint4& b) { return a.x == b.x && a.y == b.y && a.z == b.z && a.w == b.w; }
    static __device__ __forceinline__ uint32_t hash_key(const uint4& k) {
        uint32_t h = k.x * 0x9E3779B1u ^ k.y * 0x85EBCA77u ^ k.z * 0xC2B2AE3Du ^ k.w * 0x27D4EB2Fu;
        return h ^ (h >> 15);
    }
    __device__ void clear_hash() {
        for (int i = tid; i < HS; i += T) S.hash[i] = kEmptyT;
        __syncthreads();
    }
    // exclusive prefix over the team of `v` in thread order; *total = team sum.  ONE barrier: the per-warp counts
    // alternate between two buffers, and a buffer is rewritten only two scans later, i.e. after every thread has
    // passed the barrier of the scan in between and therefore finished reading it.
    __device__ __forceinline__ int team_scan(int v, int* total) {
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int u = __shfl_up_sync(kFull, inc, o);
            if (lane >= o) inc += u;
        }
        int* cnt = S.warp_cnt[scan_buf];
        scan_buf ^= 1;
        if (lane == 31) cnt[warp] = inc;
        __syncthreads();
        int before = 0, sum = 0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) { int c = cnt[w]; sum += c; if (w < warp) before += c; }
        *total = sum;
        return before + inc - v;
    }

    __device__ int count_moves(int pbase, int np, int d, bool dbl, bool second_of_b) {
        int base = 0;
        for (int i0 = 0; i0 < np; i0 += T) {
            int i = i0 + tid;
            int cnt = 0;
            if (i < np) {
                Node n = load(pbase + i);
                uint32_t mask; int special;
                one_die(n, R, d, mask, special);
                mask = prune_mask(mask, n, R, d, dbl, second_of_b);          // drop provably duplicate candidates (bg_device.cuh)
                cnt = __popc(mask) + (special >= 0);
                S.pm[i] = mask | ((uint32_t)(special + 1) << 24);
            }
            int sum;
            int ex = team_scan(cnt, &sum);
            if (i < np) S.off[i] = (uint16_t)(base + ex);
            base += sum;
        }
        if (tid == 0) S.off[np] = (uint16_t)base;
        __syncthreads();
        return base;
    }

    __device__ int expand(int pbase, int np, int d, int total, int cbase, int nc, bool use_set) {
        for (int c0 = 0; c0 < total; c0 += T) {
            int idx = c0 + tid;
            bool valid = idx < total;
            Node ch; ch.lo = 0; ch.hi = 0; ch.hit = 0; ch.occ = 0; ch.last = 31u;
            if (valid) {
                int lo = 0, hi = np - 1;
                while (lo < hi) {
                    int mid = (lo + hi + 1) >> 1;
                    if ((int)S.off[mid] <= idx) lo = mid; else hi = mid - 1;
                }
                int j = idx - (int)S.off[lo];
                uint32_t pmv = S.pm[lo];
                Node p = load(pbase + lo);
                ch = apply_move(p, R, d, pmv & 0xFFFFFFu, (int)(pmv >> 24) - 1, j);
            }
            const uint4 k = key_of(ch);
            bool keep = valid;
            int myslot = -1;
            if (use_set) {
                S.ckey[tid] = k;
                __syncthreads();
                if (valid) {
                    const uint32_t me = kPend | (uint32_t)tid;
                    uint32_t s = hash_key(k) & (HS - 1);
                    keep = false;
                    for (;;) {
                        uint32_t e = *reinterpret_cast<volatile uint32_t*>(&S.hash[s]);
                        if (e == kEmptyT) {
                            uint32_t old = atomicCAS(&S.hash[s], kEmptyT, me);
                            if (old == kEmptyT) { myslot = (int)s; keep = true; break; }
                            e = old;
                        }
                        if (e & kPend) {
                            if (eq(S.ckey[e & ~kPend], k)) {           // same board pending: lowest candidate index wins
                                uint32_t old = atomicMin(&S.hash[s], me);
                                if (old > me) { myslot = (int)s; keep = true; }
                                break;
                            }
                        } else if (eq(S.key[e], k)) break;             // already in the level
                        s = (s + 1) & (HS - 1);
                    }
                }
                __syncthreads();
                if (keep) keep = S.hash[myslot] == (kPend | (uint32_t)tid);
            }
            unsigned surv = __ballot_sync(kFull, keep);
            int sum;
            int wbefore = team_scan(lane == 0 ? __popc(surv) : 0, &sum);      // lane 0 carries the warp's count
            wbefore = __shfl_sync(kFull, wbefore, 0);
            if (nc + sum > CAP) { overflow = true; return nc; }
            if (keep) {
                int pos = cbase + nc + wbefore + __popc(surv & ((1u << lane) - 1u));
                S.key[pos] = k; S.occ[pos] = ch.occ | (ch.last << 24);
                if (use_set) S.hash[myslot] = (uint32_t)pos;                   // commit
            }
            nc += sum;
            // no barrier here: the next iteration's first barrier (after its ckey stores / its scan) also orders these
            // commits before any probe that could read them; nothing read in between is written above
        }
        __syncthreads();                                                       // the new level is complete
        return nc;
    }

    // same stage loop as Warp::generate (movegen.cu); every control variable is uniform across the CTA
    __device__ void generate(const Node& root, int d0, int d1, int& obase, int& n) {
        obase = 0; n = 0;
        if (tid == 0) { S.key[kRoot] = key_of(root); S.occ[kRoot] = root.occ | (31u << 24); }
        const bool dbl = d0 == d1;
        const int dhi = max(d0, d1), dlo = min(d0, d1);
        if (!dbl) clear_hash(); else __syncthreads();
        int pbase = kRoot, np = 1;
        int nF = 0, nA1 = 0;
        bool lenA2 = false, lenB2 = false;
        for (int stage = 0; stage < 4; ++stage) {
            const int d = dbl ? d0 : ((stage == 0 || stage == 3) ? dhi : dlo);
            const int total = count_moves(pbase, np, d, dbl, !dbl && stage == 3);
            const bool first = !dbl && (stage & 1) == 0;
            int cbase = 0, nc0 = 0;
            bool use_set = true, do_expand = true;
            if (dbl) {
                if (total == 0) break;
                cbase = (stage & 1) ? CAP : 0;
                clear_hash();
            } else if (first) {
                if (total == 0) { ++stage; continue; }
                cbase = CAP; use_set = false;
            } else if (total > 0) {
                nc0 = nF;
                if (stage == 1) lenA2 = true; else lenB2 = true;
            } else {
                do_expand = false;
                if (stage == 1 || !lenA2) {                        // singles are the plays (np <= 16: warp 0 does it)
                    if (warp == 0) {
                        Node c = load(CAP + (lane < np ? lane : 0));
                        const uint4 k = key_of(c);
                        bool keep = lane < np;
                        if (keep) {
                            uint32_t s = hash_key(k) & (HS - 1);
                            for (;;) {
                                uint32_t e = S.hash[s];
                                if (e == kEmptyT) break;
                                if (eq(S.key[e], k)) { keep = false; break; }
                                s = (s + 1) & (HS - 1);
                            }
                        }
                        unsigned surv = __ballot_sync(kFull, keep);
                        if (keep) {
                            int pos = nF + __popc(surv & ((1u << lane) - 1u));
                            S.key[pos] = k; S.occ[pos] = c.occ | (c.last << 24);
                            uint32_t s = hash_key(k) & (HS - 1);
                            while (atomicCAS(&S.hash[s], kEmptyT, (uint32_t)pos) != kEmptyT) s = (s + 1) & (HS - 1);
                        }
                        if (lane == 0) S.bcast[0] = nF + __popc(surv);
                    }
                    __syncthreads();
                    nF = S.bcast[0];
                    __syncthreads();
                    if (stage == 1) {
                        nA1 = np;
                        if (np == 1) { n = 1; return; }            // skip-reverse shortcut, get_all_moves.py:43-45
                    }
                }
            }
            int nc = nc0;
            if (do_expand) nc = expand(pbase, np, d, total, cbase, nc0, use_set);
            if (overflow) return;
            if (dbl) { pbase = cbase; np = nc; obase = cbase; n = nc; }
            else if (first) { pbase = CAP; np = nc; }
            else { if (do_expand) nF = nc; pbase = kRoot; np = 1; }
        }
        if (!dbl) {
            if (lenB2 && !lenA2) { obase = nA1; n = nF - nA1; }
            else { obase = 0; n = nF; }
        }
    }
};

template <int CAP, int HS, int T>
__global__ void __launch_bounds__(T) movegen_team_kernel(
    const int8_t* __restrict__ boards, const int8_t* __restrict__ players, const int8_t* __restrict__ dice,
    const unsigned int* __restrict__ nwork_dev, const unsigned int* __restrict__ first_dev, const int32_t* __restrict__ worklist,
    int replicate, int flip_player, int mode, const long long* __restrict__ offsets, int max_rows,
    int8_t* __restrict__ after, long long after_cap_rows, int8_t* __restrict__ row_players,
    uint16_t* __restrict__ row_feats, int32_t* __restrict__ counts_true, int32_t* __restrict__ counts, long long* __restrict__ starts,
    unsigned long long* __restrict__ alloc, int32_t* __restrict__ status,
    unsigned int* __restrict__ work_ctr, int32_t* __restrict__ overflow_list, unsigned int* __restrict__ overflow_ctr) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TeamScratch<CAP, HS, T>& S = *reinterpret_cast<TeamScratch<CAP, HS, T>*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long nwork = (long long)*nwork_dev, first = first_dev ? (long long)*first_dev : 0;   // list entries [first, nwork)
    __shared__ uint32_t s_lut[32], s_desc[32];
    load_chunk_tables(s_lut, s_desc);
    __syncthreads();

    for (;;) {
        if (tid == 0) S.bcast[1] = (int)atomicAdd(work_ctr, 1u);
        __syncthreads();
        const long long wi = first + (long long)(unsigned int)S.bcast[1];
        if (wi >= nwork) break;
        const long long g = (long long)worklist[wi];
        const WorkItem item = decode_work_item(g, replicate > 1 ? g / replicate : g, replicate, flip_player, players, dice);
        const int player = item.player, d0 = item.d0, d1 = item.d1;
        if (warp == 0) {                                  // root: warp-collective, shared with movegen.cu (tier 0 already rejected malformed boards)
            Root R; Node root;
            build_root(load_board_word(boards, item.src, lane), player, lane, S.rootw, R, root);
            if (d0 != d1) { uint32_t mA; int sA; one_die(root, R, max(d0, d1), mA, sA); R.mA = mA; }   // larger-die sources at the root
            if (lane == 0) { S.root = root; S.R = R; }
        }
        __syncthreads();
        Team<CAP, HS, T> W(S);
        W.R = S.R;
        const Node root = S.root;
        int obase = 0, n = 0;
        W.generate(root, d0, d1, obase, n);
        if (W.overflow) {
            if (tid == 0) {
                if (overflow_list) { unsigned int k = atomicAdd(overflow_ctr, 1u); overflow_list[k] = (int32_t)g; }
                else {
                    atomicOr(status, BG_STATUS_SCRATCH_OVERFLOW);
                    if (counts_true) counts_true[g] = -1;
                    if (counts) counts[g] = 0;
                }
            }
            __syncthreads();
            continue;
        }
        const int nw = (max_rows > 0 && n > max_rows) ? max_rows : n;
        if (tid == 0) {
            long long st = 0;
            if (mode == 1) st = offsets[g];
            else if (mode == 2) st = (long long)atomicAdd(alloc, (unsigned long long)nw);
            S.bcast[2] = (int)(st & 0xFFFFFFFFll); S.bcast[3] = (int)(st >> 32);
            if (counts_true) counts_true[g] = n;
            if (counts) counts[g] = nw;
            if (mode == 2 && starts) starts[g] = st;
        }
        __syncthreads();
        const long long start = (long long)(unsigned int)S.bcast[2] | ((long long)S.bcast[3] << 32);
        if (mode != 0 && nw > 0) {
            if (start + nw > after_cap_rows) {
                if (tid == 0) { atomicOr(status, BG_STATUS_OUTPUT_OVERFLOW); if (counts) counts[g] = 0; }
            } else {
                uint32_t* stage = reinterpret_cast<uint32_t*>(&S.key[obase < CAP ? CAP : 0]);
                uint32_t* gout = reinterpret_cast<uint32_t*>(after) + start * kBoardWords;
                const RowContext rc = make_row_context(player, S.rootw);
                constexpr int SR = Team<CAP, HS, T>::kStageRows;
                for (int r0 = 0; r0 < nw; r0 += SR) {
                    int r = r0 + tid;
                    if (tid < SR && r < nw) {
                        build_row(S.key[obase + r], player, rc, S.rootw, stage + tid * kBoardWords);
                    }
                    __syncthreads();
                    int rows = min(SR, nw - r0);
                    for (int k2 = tid; k2 < rows * kBoardWords; k2 += T) gout[(long long)r0 * kBoardWords + k2] = stage[k2];
                    if (row_players && tid < rows) row_players[start + r0 + tid] = (int8_t)player;
                    if (row_feats) {                          // fused K3 (see movegen.cu)
                        uint4* fdst = reinterpret_cast<uint4*>(row_feats + (start + r0) * (long long)BG_FEAT_LD_BF16);
                        const uint8_t* sb = reinterpret_cast<const uint8_t*>(stage);
                        for (int c = tid; c < rows * 26; c += T) {
                            const int r = c / 26, k = c - r * 26;
                            fdst[c] = chunk_from_desc(sb + r * kBoardBytes, player, s_desc[k], s_lut);
                        }
                    }
                    __syncthreads();
                }
            }
        }
        __syncthreads();
    }
}
}  // namespace

// team size override for tuning runs (threads per position); the defaults are the measured best
static int team_size_env(const char* name, int dflt) {
    const char* v = getenv(name);
    return v ? atoi(v) : dflt;
}
// ... and for the race tests (bg_set_team_threads): the cross-warp claim / atomicMin protocol must give bit-identical
// output whatever the team size
static int g_team_mid_override = 0, g_team_big_override = 0;

template <int CAP, int HS, int T>
static int launch_team(const int8_t* boards, const int8_t* players, const int8_t* dice, const unsigned int* nwork_dev,
                       const unsigned int* first_dev, const int32_t* worklist, int replicate, int flip_player, int mode, const long long* offsets,
                       int max_rows, int8_t* after, long long after_cap_rows, int8_t* row_players, uint16_t* row_feats,
                       int32_t* counts_true,
                       int32_t* counts, long long* starts, unsigned long long* alloc, int32_t* status,
                       unsigned int* work_ctr, int32_t* overflow_list, unsigned int* overflow_ctr, cudaStream_t stream) {
    size_t smem = sizeof(TeamScratch<CAP, HS, T>);
    auto kern = movegen_team_kernel<CAP, HS, T>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return bg_set_error(e, "movegen(team): cudaFuncSetAttribute");
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, T, smem);
    if (occ < 1) occ = 1;
    unsigned grid = (unsigned)(bg_sm_count() * occ);
    kern<<<grid, T, smem, stream>>>(boards, players, dice, nwork_dev, first_dev, worklist, replicate, flip_player, mode, offsets,
                                    max_rows, after, after_cap_rows, row_players, row_feats, counts_true, counts, starts,
                                    alloc, status, work_ctr, overflow_list, overflow_ctr);
    return bg_set_error(cudaGetLastError(), "movegen(team): launch");
}

int movegen_team_mid(const int8_t* boards, const int8_t* players, const int8_t* dice, const unsigned int* nwork_dev,
                     const int32_t* worklist, int replicate, int flip_player, int mode, const long long* offsets,
                     int max_rows, int8_t* after, long long after_cap_rows, int8_t* row_players, uint16_t* row_feats,
                       int32_t* counts_true,
                     int32_t* counts, long long* starts, unsigned long long* alloc, int32_t* status,
                     unsigned int* work_ctr, int32_t* overflow_list, unsigned int* overflow_ctr, cudaStream_t stream,
                     int team_threads_hint) {
#define BG_TEAM_MID(T) launch_team<BG_MOVEGEN_CAP_MID, 2 * BG_MOVEGEN_CAP_MID, T>( \
        boards, players, dice, nwork_dev, nullptr, worklist, replicate, flip_player, mode, offsets, max_rows, after, after_cap_rows, \
        row_players, row_feats, counts_true, counts, starts, alloc, status, work_ctr, overflow_list, overflow_ctr, stream)
    // 256 threads per position is the fastest alone; when an encoder runs beside the tiers (bg_update_legal_plays)
    // 128 leaves it more of the SMs' thread slots and the pair finishes sooner
    static const int t1_env = team_size_env("BG_TEAM_MID", 0);
    const int t1 = g_team_mid_override ? g_team_mid_override : (t1_env ? t1_env : (team_threads_hint ? team_threads_hint : 256));
    if (t1 == 128) return BG_TEAM_MID(128);
    if (t1 == 512) return BG_TEAM_MID(512);
    return BG_TEAM_MID(256);
#undef BG_TEAM_MID
}
int movegen_team_big(const int8_t* boards, const int8_t* players, const int8_t* dice, const unsigned int* nwork_dev,
                     const unsigned int* first_dev, const int32_t* worklist, int replicate, int flip_player, int mode, const long long* offsets,
                     int max_rows, int8_t* after, long long after_cap_rows, int8_t* row_players, uint16_t* row_feats,
                       int32_t* counts_true,
                     int32_t* counts, long long* starts, unsigned long long* alloc, int32_t* status,
                     unsigned int* work_ctr, cudaStream_t stream) {
#define BG_TEAM_BIG(T) launch_team<BG_MOVEGEN_CAP_BIG, BG_MOVEGEN_HASH_BIG, T>( \
        boards, players, dice, nwork_dev, first_dev, worklist, replicate, flip_player, mode, offsets, max_rows, after, after_cap_rows, \
        row_players, row_feats, counts_true, counts, starts, alloc, status, work_ctr, nullptr, nullptr, stream)
    static const int t2_env = team_size_env("BG_TEAM_BIG", 512);
    const int t2 = g_team_big_override ? g_team_big_override : t2_env;
    if (t2 == 1024) return BG_TEAM_BIG(1024);
    return BG_TEAM_BIG(512);
#undef BG_TEAM_BIG
}

}  // namespace bg

extern "C" int bg_set_team_threads(int mid, int big) {
    if ((mid != 0 && mid != 128 && mid != 256 && mid != 512) || (big != 0 && big != 512 && big != 1024))
        return bg_set_error_msg(BG_ERR_INVALID, "bg_set_team_threads: mid must be 0/128/256/512, big 0/512/1024");
    bg::g_team_mid_override = mid; bg::g_team_big_override = big;
    return BG_OK;
}
