// bg_movegen_warp.cuh -- K1's warp-per-position generator (level-by-level expansion with ordered first-wins dedupe), shared by
// the legal-move kernel (movegen.cu) and the fused 2-ply kernel (twoply_fused.cu).  See movegen.cu for the algorithm and
// the reference lines it replaces (moves/get_all_moves.py:9-94, moves/handle_moves.py:109-341).
#pragma once
#include "bg_device.cuh"

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
    static constexpr int kRoot = 2 * CAP;
    WarpScratch<CAP, HS>& S;
    Root R;
    int lane;
    bool overflow;
    int ovf_stage;              // the stage (0-based: level stage -> stage + 1) whose children did not fit, when overflow is set

    __device__ Warp(WarpScratch<CAP, HS>& s, int l) : S(s), lane(l), overflow(false), ovf_stage(3) {}

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
            if (overflow) { ovf_stage = dbl ? stage : 3; return; }
            if (dbl) { pbase = cbase; np = nc; obase = cbase; n = nc; }
            else {
                // filter_full_moves_by_max_submoves (get_all_moves.py:73-94), applied AFTER dedupe
                if (lenB2_only) { obase = nA1; n = nc - nA1; } else { obase = 0; n = nc; }
            }
        }
    }
};

}  // namespace bg
