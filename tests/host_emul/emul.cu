// Host emulation of K1's algorithm (level-by-level expansion with in-order dedupe), compiled from the SAME
// rule code as the kernel (mlp-ppo-2ply-p3_b200/csrc/bg_device.cuh is __host__ __device__).  Test support only:
// lets the CPU suite check the algorithm (order preservation, Q1 shortcut, max filter) without a GPU.
#include <vector>
#include <cstring>
#include "bg_device.cuh"
using namespace bg;

static bool same(const Node& a, const Node& b) { return a.lo == b.lo && a.hi == b.hi && a.hit == b.hit; }

// children of `parents` with die d appended to `out` (deduped against everything already in `out`), reference order
static int expand(const std::vector<Node>& parents, const Root& R, int d, std::vector<Node>& out, bool doubles = false,
                  bool second_of_b = false) {
    int total = 0;
    for (const Node& p : parents) {
        uint32_t mask; int special;
        one_die(p, R, d, mask, special);
        mask = prune_mask(mask, p, R, d, doubles, second_of_b);      // the kernel's duplicate pruning (bg_device.cuh)
        int cnt = BG_POPC(mask) + (special >= 0);
        total += cnt;
        for (int j = 0; j < cnt; ++j) {
            Node c = apply_move(p, R, d, mask, special, j);
            bool dup = false;
            for (const Node& o : out) if (same(o, c)) { dup = true; break; }
            if (!dup) out.push_back(c);
        }
    }
    return total;
}

extern "C" int emul_legal_moves(const int8_t* b52, int player, int d0, int d1, int8_t* out52, int cap) {
    uint32_t rootw[13];
    memcpy(rootw, b52, 52);
    const int8_t* own = b52 + (player ? 24 : 0);
    const int8_t* opp = b52 + (player ? 0 : 24);
    Root R; R.player = player; R.block = 0; R.blot = 0; R.mA = 0; R.cnt2 = 0;
    Node root; root.lo = 0; root.hi = 0; root.occ = 0; root.hit = 0; root.last = 31u;
    int total = 0;
    for (int p = 0; p < 24; ++p) {
        if (own[p] >= 2) R.cnt2 |= 1u << p;
        if (opp[p] >= 2) R.block |= 1u << p;
        if (opp[p] == 1) R.blot |= 1u << p;
        if (own[p] > 0) root.occ |= 1u << p;
        if (p < 16) root.lo |= (unsigned long long)(own[p] & 15) << (4 * p);
        else root.hi |= (unsigned long long)(own[p] & 15) << (4 * (p - 16));
        total += own[p];
    }
    int ownbar = b52[48 + player], ownoff = b52[50 + player];
    root.hi |= (unsigned long long)((ownbar & 15) | ((ownoff & 15) << 4)) << 32;
    R.tot15 = (total + ownbar + ownoff) == 15;
    if (d0 != d1) { uint32_t mA; int sA; one_die(root, R, d0 > d1 ? d0 : d1, mA, sA); R.mA = mA; }

    std::vector<Node> F;     // result
    size_t from = 0;
    if (d0 != d1) {
        int hi = d0 > d1 ? d0 : d1, lo = d0 > d1 ? d1 : d0;
        std::vector<Node> rootv{root}, L1, tmp;
        bool lenA2 = false, lenB2 = false; size_t nA1 = 0;
        expand(rootv, R, hi, L1);
        bool done = false;
        if (!L1.empty()) {
            std::vector<Node> probe;
            int t2 = 0;
            for (const Node& p : L1) { uint32_t m; int s; one_die(p, R, lo, m, s); t2 += BG_POPC(m) + (s >= 0); }
            if (t2) { expand(L1, R, lo, F); lenA2 = true; }
            else { F = L1; nA1 = L1.size(); if (L1.size() == 1) done = true; }
        }
        if (!done) {
            std::vector<Node> L1b;
            expand(rootv, R, lo, L1b);
            if (!L1b.empty()) {
                int t2 = 0;
                for (const Node& p : L1b) { uint32_t m; int s; one_die(p, R, hi, m, s); t2 += BG_POPC(m) + (s >= 0); }
                if (t2) { expand(L1b, R, hi, F, false, true); lenB2 = true; }
                else if (!lenA2) {
                    for (const Node& c : L1b) { bool dup = false; for (const Node& o : F) if (same(o, c)) dup = true; if (!dup) F.push_back(c); }
                }
            }
            if (lenB2 && !lenA2) from = nA1;
        }
    } else {
        std::vector<Node> cur{root};
        for (int depth = 1; depth <= 4; ++depth) {
            std::vector<Node> nxt;
            int t = expand(cur, R, d0, nxt, true);
            if (t == 0) break;
            cur = nxt; F = cur;
        }
    }
    int n = (int)(F.size() - from);
    for (int i = 0; i < n && i < cap; ++i) {
        uint32_t w[13];
        for (int k = 0; k < 13; ++k) w[k] = node_row_word(F[from + i], player, rootw, k);
        memcpy(out52 + 52 * (size_t)i, w, 52);
    }
    return n;
}

// Level sizes of a doubles position (distinct boards per level 1..4 and candidates built per level): sizing data for the
// kernel's per-warp scratch (scripts/level_stats.py).
extern "C" void emul_doubles_levels(const int8_t* b52, int player, int d, int* sizes4, int* cands4) {
    const int8_t* own = b52 + (player ? 24 : 0);
    const int8_t* opp = b52 + (player ? 0 : 24);
    Root R; R.player = player; R.block = 0; R.blot = 0; R.mA = 0; R.cnt2 = 0;
    Node root; root.lo = 0; root.hi = 0; root.occ = 0; root.hit = 0; root.last = 31u;
    int total = 0;
    for (int p = 0; p < 24; ++p) {
        if (own[p] >= 2) R.cnt2 |= 1u << p;
        if (opp[p] >= 2) R.block |= 1u << p;
        if (opp[p] == 1) R.blot |= 1u << p;
        if (own[p] > 0) root.occ |= 1u << p;
        if (p < 16) root.lo |= (unsigned long long)(own[p] & 15) << (4 * p);
        else root.hi |= (unsigned long long)(own[p] & 15) << (4 * (p - 16));
        total += own[p];
    }
    int ownbar = b52[48 + player], ownoff = b52[50 + player];
    root.hi |= (unsigned long long)((ownbar & 15) | ((ownoff & 15) << 4)) << 32;
    R.tot15 = (total + ownbar + ownoff) == 15;
    std::vector<Node> cur{root};
    for (int k = 0; k < 4; ++k) { sizes4[k] = 0; cands4[k] = 0; }
    for (int depth = 0; depth < 4; ++depth) {
        std::vector<Node> nxt;
        int t = expand(cur, R, d, nxt, true);
        if (t == 0) break;
        sizes4[depth] = (int)nxt.size(); cands4[depth] = t;
        cur = nxt;
    }
}
