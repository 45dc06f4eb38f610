// bg_device.cuh -- device-side board representation and backgammon rules for sm_100a.
//
// HBM layout of one position ("board52"): 52 bytes, int8
//   [0..23]  PLAYER1 checkers on points 0..23      (reference row 0, src/board/immutable_board.py:20-27)
//   [24..47] PLAYER2 checkers on points 0..23      (reference row 1)
//   [48,49]  bar PLAYER1, PLAYER2                  (reference row 2, cols 0,1)
//   [50,51]  borne off PLAYER1, PLAYER2            (reference row 3, cols 0,1)
//
// In registers a position is seen from the mover's side as a Node: the mover's
// 24 point counts packed one nibble per point (counts are 0..15), the mover's
// bar/off nibbles, a 24-bit occupancy mask and the set of opponent blots hit so
// far.  The opponent's men never move during the mover's turn except by being
// hit, so (root opponent rows, hit mask) reproduces the opponent side exactly.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

// The rule functions are __host__ __device__ so that tests/host_emul can run the very same code on the CPU.
#define BG_HD __host__ __device__ __forceinline__
#ifdef __CUDA_ARCH__
#define BG_FFS(x) __ffs((int)(x))
#define BG_CLZ(x) __clz((int)(x))
#define BG_POPC(x) __popc(x)
#else
#define BG_FFS(x) __builtin_ffs((int)(x))
#define BG_CLZ(x) ((x) ? __builtin_clz((unsigned)(x)) : 32)
#define BG_POPC(x) __builtin_popcount((unsigned)(x))
#endif

namespace bg {

constexpr int kBoardBytes = 52;
constexpr int kBoardWords = 13;
constexpr int kBar = 24;   // Position.BAR       (src/moves/move_types.py:34)
constexpr int kOff = 25;   // Position.BEAR_OFF  (src/moves/move_types.py:35)
constexpr unsigned kFull = 0xFFFFFFFFu;

struct Node {
    unsigned long long lo;  // own points 0..15, one nibble each
    unsigned long long hi;  // bits 0..31 own points 16..23; bits 32..35 own bar; bits 36..39 own off
    uint32_t occ;           // bit p set <=> own count at point p > 0
    uint32_t hit;           // bit p set <=> the opponent blot on point p has been hit this turn
    uint32_t last;          // source point of the sub-move that first produced this board (31: none / bar / bear-off)
};

// Per-root (warp-uniform) constants.
struct Root {
    uint32_t block;   // opponent has >= 2 men on point p      (conditions.py:24-26)
    uint32_t blot;    // opponent has exactly 1 man on point p (conditions.py:51-53)
    int player;       // mover: 0 = PLAYER1 (moves +), 1 = PLAYER2 (moves -)
    int tot15;        // own points + bar + off == 15 (needed by all_checkers_home, conditions.py:147)
    // duplicate pruning (prune_mask below)
    uint32_t mA;      // non-doubles: ordinary sources of the LARGER die at the root
    uint32_t cnt2;    // own count at point p >= 2 at the root
};

BG_HD int node_bar(const Node& n) { return (int)((n.hi >> 32) & 15ull); }
BG_HD int node_off(const Node& n) { return (int)((n.hi >> 36) & 15ull); }
BG_HD int node_count(const Node& n, int p) {
    return p < 16 ? (int)((n.lo >> (4 * p)) & 15ull) : (int)((n.hi >> (4 * (p - 16))) & 15ull);
}

// One-die move list of a node as (mask of ordinary source points, special move).
// special: -1 none, 0..23 bear-off from that point, kBar = enter from the bar.
// List order of the reference = ascending source point, then the bear-off move
// (get_moves_normal move_logic.py:47-92, get_moves_bar :95-137, get_moves_bear_off :140-255,
//  compute_board_state :258-275).
BG_HD void one_die(const Node& n, const Root& R, int d, uint32_t& mask, int& special) {
    mask = 0; special = -1;
    if (node_off(n) == 15) return;                                   // GAME_OVER (conditions.py:96-108)
    if (node_bar(n) > 0) {                                           // ON_BAR
        int e = R.player == 0 ? d - 1 : 24 - d;                      // move_logic.py:111-114
        if (!((R.block >> e) & 1u)) special = kBar;                  // conditions.py:57-78
        return;
    }
    if (R.player == 0) mask = n.occ & ~(R.block >> d) & ((1u << (24 - d)) - 1u);
    else               mask = n.occ & ~(R.block << d) & ~((1u << d) - 1u) & 0xFFFFFFu;
    const uint32_t home = R.player == 0 ? 0xFC0000u : 0x00003Fu;     // conditions.py:123-126
    if (R.tot15 && (n.occ & ~home) == 0 && n.occ != 0) {             // BEAR_OFF (all_checkers_home)
        if (R.player == 0) {
            int last = BG_FFS(n.occ) - 1;                             // farthest from exit, move_logic.py:196-201
            if (last + d >= 24) special = last;                      // :212-218
            else if ((n.occ >> (24 - d)) & 1u) special = 24 - d;     // :221-231
        } else {
            int last = 31 - BG_CLZ(n.occ);                            // :202-207
            if (last - d < 0) special = last;                        // :234-240
            else if ((n.occ >> (d - 1)) & 1u) special = d - 1;       // :243-253
        }
    }
}

// Index of the j-th (0-based) set bit of a 24-bit mask, j < popc(m).  Branch-free 12/6/3 bisection on popcounts
// plus a 2-bit-per-entry table for the last 3 bits (a data-dependent "clear lowest bit j times" loop cost ~5
// instructions per iteration of the slowest lane of the warp).
// ---------------------------------------------------------------------------------------------------------------
// Candidates that are PROVABLY duplicates of an earlier candidate of the same level can be dropped before they are
// built: that changes neither the surviving boards nor their order (the reference keeps the first sequence that
// reaches a board, handle_moves.py:313-341).  Both rules only involve ORDINARY moves (point to point), which every
// list of the reference enumerates first and by ascending source (move_logic.py:67,172-193), and whose legality
// does not depend on the board state once nobody is on the bar.
//
// Doubles.  Board P was first reached by a sequence ending with the move from point m.  A further move from s < m
// whose man does not owe its presence to m's arrival (NOT (s == dest(m) and P has exactly one man on s)) can be
// played BEFORE m: (..., s) reaches a board P'' whose first sequence is lexicographically smaller than P's, so
// P'' precedes P in the level and (P'', m) produced the same board earlier.
// Non-doubles.  Smaller-die-first board (root + lo from s2), then the larger die from s1: if s1 is a larger-die
// source at the root and the two moves use different men (s1 != s2 or two men on s2), the larger-die-first pass
// (which is enumerated first) already produced the board as (hi from s1, lo from s2).
// Everything else still goes through the exact dedupe.  Measured on 90,000 random-play positions x (their roll + six
// doubles): 98.6 % of the duplicate candidates of doubles and 90 % of those of non-doubles are dropped this way, and
// none of 99 M candidates dropped was a first occurrence; tests/test_kernel_algorithm_cpu.py runs the rule on the
// golden corpora through the host emulation.
BG_HD uint32_t prune_mask(uint32_t mask, const Node& n, const Root& R, int d, bool doubles, bool smaller_die_first_parent);

BG_HD int nth_set_bit(uint32_t m, int j) {
    int pos = 0, t;
    t = BG_POPC(m & 0xFFFu);          if (j >= t) { j -= t; pos = 12; }
    t = BG_POPC((m >> pos) & 0x3Fu);  if (j >= t) { j -= t; pos += 6; }
    t = BG_POPC((m >> pos) & 0x7u);   if (j >= t) { j -= t; pos += 3; }
    const uint32_t r = (m >> pos) & 7u;
    // entry (r, j) at bit 2*(3r+j): position of the j-th set bit of the 3-bit value r
    constexpr unsigned long long kSel3 =
        (1ull << (2 * (2 * 3 + 0))) | (1ull << (2 * (3 * 3 + 1))) | (2ull << (2 * (4 * 3 + 0))) | (2ull << (2 * (5 * 3 + 1))) |
        (1ull << (2 * (6 * 3 + 0))) | (2ull << (2 * (6 * 3 + 1))) | (1ull << (2 * (7 * 3 + 1))) | (2ull << (2 * (7 * 3 + 2)));
    return pos + (int)((kSel3 >> (2 * (r * 3 + (uint32_t)j))) & 3ull);
}

BG_HD uint32_t prune_mask(uint32_t mask, const Node& n, const Root& R, int d, bool doubles, bool smaller_die_first_parent) {
    if (n.last >= 24u) return mask;
    if (doubles) {
        uint32_t drop = mask & ((1u << n.last) - 1u);
        if (R.player != 0) {                                         // PLAYER2 lands below its source
            const int dest = (int)n.last - d;
            if (dest >= 0 && ((drop >> dest) & 1u) && node_count(n, dest) == 1) drop &= ~(1u << dest);
        }
        return mask & ~drop;
    }
    if (!smaller_die_first_parent) return mask;
    return mask & ~(R.mA & ~((1u << n.last) & ~R.cnt2));
}

// Apply move number j of the (mask, special) list with die d  (move_checker, immutable_board.py:42-89).
BG_HD Node apply_move(const Node& n, const Root& R, int d, uint32_t mask, int special, int j) {
    Node c = n;
    int nm = BG_POPC(mask);
    int s, t;
    c.last = 31u;
    if (j < nm) { s = nth_set_bit(mask, j); t = R.player == 0 ? s + d : s - d; c.last = (uint32_t)s; }
    else if (special == kBar) { s = kBar; t = R.player == 0 ? d - 1 : 24 - d; }
    else { s = special; t = kOff; }
    if (s == kBar) c.hi -= 1ull << 32;
    else {
        if (s < 16) c.lo -= 1ull << (4 * s); else c.hi -= 1ull << (4 * (s - 16));
        if (node_count(c, s) == 0) c.occ &= ~(1u << s);
    }
    if (t == kOff) c.hi += 1ull << 36;
    else {
        if (t < 16) c.lo += 1ull << (4 * t); else c.hi += 1ull << (4 * (t - 16));
        c.occ |= 1u << t;
        if (((R.blot & ~n.hit) >> t) & 1u) c.hit |= 1u << t;
    }
    return c;
}

// 4 nibbles (16 bits) -> 4 bytes
BG_HD uint32_t spread_nibbles(uint32_t x) {
    x &= 0xFFFFu;
    x = (x | (x << 8)) & 0x00FF00FFu;
    x = (x | (x << 4)) & 0x0F0F0F0Fu;
    return x;
}
// 4 bits -> 4 bytes of 0/1
BG_HD uint32_t spread_bits(uint32_t x) { return ((x & 0xFu) * 0x00204081u) & 0x01010101u; }

// Word k (0..12) of the board52 row of node n; rootw = the root's 13 words.
BG_HD uint32_t node_row_word(const Node& n, int player, const uint32_t* rootw, int k) {
    const int own0 = player == 0 ? 0 : 6, opp0 = player == 0 ? 6 : 0;
    if (k == 12) {
        uint32_t ob = (uint32_t)node_bar(n), oo = (uint32_t)node_off(n);
        uint32_t m = rootw[12];
        uint32_t pb = ((m >> (player == 0 ? 8 : 0)) & 0xFFu) + (uint32_t)BG_POPC(n.hit);   // opponent bar
        uint32_t po = (m >> (player == 0 ? 24 : 16)) & 0xFFu;                              // opponent off
        return player == 0 ? (ob | (pb << 8) | (oo << 16) | (po << 24))
                           : (pb | (ob << 8) | (po << 16) | (oo << 24));
    }
    if (k >= own0 && k < own0 + 6) {
        int q = k - own0;
        uint32_t nib = q < 4 ? (uint32_t)(n.lo >> (16 * q)) : (uint32_t)(n.hi >> (16 * (q - 4)));
        return spread_nibbles(nib);
    }
    int q = k - opp0;
    return rootw[k] - spread_bits(n.hit >> (4 * q));
}

BG_HD uint32_t hash_node(const Node& n) {
    uint32_t h = (uint32_t)n.lo * 0x9E3779B1u;
    h ^= (uint32_t)(n.lo >> 32) * 0x85EBCA77u;
    h ^= (uint32_t)n.hi * 0xC2B2AE3Du;
    h ^= (uint32_t)(n.hi >> 32) * 0x27D4EB2Fu;
    h ^= n.hit * 0x165667B1u;
    return h ^ (h >> 15);
}

// ---------------------------------------------------------------- Philox4x32-10 (Salmon et al., SC'11)
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
constexpr uint32_t kTagDice = 0x44494345u;   // "DICE"
constexpr uint32_t kTagAct = 0x41435431u;    // "ACT1"

// Draw `t` of game stream `g`: two dice in 1..6.
__device__ __forceinline__ void philox_dice(unsigned long long seed, unsigned long long g, uint32_t t, int& d0, int& d1) {
    uint32_t o[4];
    philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), t, kTagDice, (uint32_t)seed, (uint32_t)(seed >> 32), o);
    d0 = 1 + (int)__umulhi(o[0], 6u);
    d1 = 1 + (int)__umulhi(o[1], 6u);
}
__device__ __forceinline__ uint32_t philox_action(unsigned long long seed, unsigned long long g, uint32_t t, uint32_t n) {
    uint32_t o[4];
    philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), t, kTagAct, (uint32_t)seed, (uint32_t)(seed >> 32), o);
    return __umulhi(o[0], n);
}

}  // namespace bg
