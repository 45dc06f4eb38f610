// bg_movegen_common.cuh -- pieces shared by K1's warp-per-position kernel (movegen.cu) and its CTA-per-position
// tiers (movegen_team.cu): the warp-collective root builder and the afterstate row builder.
#pragma once
#include "bg_device.cuh"

namespace bg {

// the 21 sorted rolls in the order of get_all_dice_rolls_tensor (moves/get_all_dice_rolls.py:19-32)
__device__ __constant__ int8_t kRoll21[21][2] = {{1, 1}, {1, 2}, {1, 3}, {1, 4}, {1, 5}, {1, 6}, {2, 2}, {2, 3}, {2, 4}, {2, 5}, {2, 6},
                                                {3, 3}, {3, 4}, {3, 5}, {3, 6}, {4, 4}, {4, 5}, {4, 6}, {5, 5}, {5, 6}, {6, 6}};

// Work item g of a launch: position `src`, mover, dice.  replicate == 21: item g = (position g / 21, sorted roll g % 21),
// the 2-ply opponent expansion.
struct WorkItem { long long src; int player, d0, d1; };
__device__ __forceinline__ WorkItem decode_work_item(long long g, long long src, int replicate, int flip_player,
                                                     const int8_t* __restrict__ players, const int8_t* __restrict__ dice) {
    WorkItem it;
    it.src = src;
    it.player = (players[src] ^ flip_player) & 1;
    if (replicate > 1) { const int r = (int)(g - src * replicate); it.d0 = kRoll21[r][0]; it.d1 = kRoll21[r][1]; }
    else { it.d0 = dice[2 * g]; it.d1 = dice[2 * g + 1]; }
    return it;
}

// lane's word of a board52 row (13 words, coalesced; issue it before anything that depends on other loads)
__device__ __forceinline__ uint32_t load_board_word(const int8_t* __restrict__ boards, long long src, int lane) {
    return lane < kBoardWords ? reinterpret_cast<const uint32_t*>(boards + src * kBoardBytes)[lane] : 0u;
}
// Warp-collective: from the 13 words of a board52 row (one per lane, load_board_word) keep a copy in rootw[] (shared
// memory, for the output stage) and build the mover-relative view: Root constants (opponent block / blot masks, own
// >= 2 mask) and the root Node (own counts as nibbles, bar / off, occupancy).  Returns false (on every lane) for a
// malformed board.
__device__ __forceinline__ bool build_root(uint32_t w, int player, int lane, uint32_t* rootw, Root& R, Node& root) {
    if (lane < kBoardWords) rootw[lane] = w;
    const int p = lane < 24 ? lane : 0;
    const uint32_t ownw = __shfl_sync(kFull, w, (player ? 6 : 0) + (p >> 2));
    const uint32_t oppw = __shfl_sync(kFull, w, (player ? 0 : 6) + (p >> 2));
    const uint32_t misc = __shfl_sync(kFull, w, 12);
    const int ownc = lane < 24 ? (int)((ownw >> (8 * (p & 3))) & 0xFFu) : 0;
    const int oppc = lane < 24 ? (int)((oppw >> (8 * (p & 3))) & 0xFFu) : 0;
    const int ownbar = (int)((misc >> (player ? 8 : 0)) & 0xFFu), ownoff = (int)((misc >> (player ? 24 : 16)) & 0xFFu);
    R.player = player;
    R.block = __ballot_sync(kFull, oppc >= 2) & 0xFFFFFFu;
    R.blot = __ballot_sync(kFull, oppc == 1) & 0xFFFFFFu;
    R.cnt2 = __ballot_sync(kFull, ownc >= 2) & 0xFFFFFFu;
    R.mA = 0;
    root.occ = __ballot_sync(kFull, ownc > 0) & 0xFFFFFFu;
    root.hit = 0;
    root.last = 31u;
    const uint32_t nib = (uint32_t)(ownc & 15) << (4 * (p & 7));
    const uint32_t w0 = __reduce_or_sync(kFull, (lane < 8) ? nib : 0u);
    const uint32_t w1 = __reduce_or_sync(kFull, (lane >= 8 && lane < 16) ? nib : 0u);
    const uint32_t w2 = __reduce_or_sync(kFull, (lane >= 16 && lane < 24) ? nib : 0u);
    root.lo = (unsigned long long)w0 | ((unsigned long long)w1 << 32);
    root.hi = (unsigned long long)w2 | ((unsigned long long)((ownbar & 15) | ((ownoff & 15) << 4)) << 32);
    R.tot15 = (__reduce_add_sync(kFull, ownc) + ownbar + ownoff) == 15;
    return !(__any_sync(kFull, ownc > 15 || oppc > 15) || ownbar > 15 || ownoff > 15);
}

// The board52 row (13 words) of the level entry with key k = (points 0..7, 8..15, 16..23 as nibbles, hit mask | bar << 24
// | off << 28) of mover `player`, written to row[0..12]; rootw = the root's 13 words.
struct RowContext { int own0, opp0; uint32_t opp_bar0, opp_off0; };
__device__ __forceinline__ RowContext make_row_context(int player, const uint32_t* rootw) {
    RowContext c;
    c.own0 = player ? 6 : 0; c.opp0 = player ? 0 : 6;
    const uint32_t misc0 = rootw[12];
    c.opp_bar0 = (misc0 >> (player ? 0 : 8)) & 0xFFu; c.opp_off0 = (misc0 >> (player ? 16 : 24)) & 0xFFu;
    return c;
}
__device__ __forceinline__ void build_row(const uint4& k, int player, const RowContext& c, const uint32_t* rootw, uint32_t* row) {
    row[c.own0 + 0] = spread_nibbles(k.x);       row[c.own0 + 1] = spread_nibbles(k.x >> 16);
    row[c.own0 + 2] = spread_nibbles(k.y);       row[c.own0 + 3] = spread_nibbles(k.y >> 16);
    row[c.own0 + 4] = spread_nibbles(k.z);       row[c.own0 + 5] = spread_nibbles(k.z >> 16);
#pragma unroll
    for (int q = 0; q < 6; ++q) row[c.opp0 + q] = rootw[c.opp0 + q] - spread_bits(k.w >> (4 * q));
    const uint32_t ob = (k.w >> 24) & 15u, oo = k.w >> 28;
    const uint32_t pb = c.opp_bar0 + (uint32_t)__popc(k.w & 0xFFFFFFu);
    row[12] = player == 0 ? (ob | (pb << 8) | (oo << 16) | (c.opp_off0 << 24))
                          : (pb | (ob << 8) | (c.opp_off0 << 16) | (oo << 24));
}

// The same row written straight to global memory by its own lane (13 word stores; no staging, no dynamic register indexing)
__device__ __forceinline__ void store_row_direct(const uint4& k, int player, const RowContext& c, const uint32_t* rootw, uint32_t* __restrict__ dst) {
    uint32_t own[6], opp[6];
    own[0] = spread_nibbles(k.x); own[1] = spread_nibbles(k.x >> 16);
    own[2] = spread_nibbles(k.y); own[3] = spread_nibbles(k.y >> 16);
    own[4] = spread_nibbles(k.z); own[5] = spread_nibbles(k.z >> 16);
#pragma unroll
    for (int q = 0; q < 6; ++q) opp[q] = rootw[c.opp0 + q] - spread_bits(k.w >> (4 * q));
    const uint32_t ob = (k.w >> 24) & 15u, oo = k.w >> 28;
    const uint32_t pb = c.opp_bar0 + (uint32_t)__popc(k.w & 0xFFFFFFu);
#pragma unroll
    for (int q = 0; q < 6; ++q) { dst[q] = player ? opp[q] : own[q]; dst[6 + q] = player ? own[q] : opp[q]; }
    dst[12] = player == 0 ? (ob | (pb << 8) | (oo << 16) | (c.opp_off0 << 24))
                          : (pb | (ob << 8) | (c.opp_off0 << 16) | (oo << 24));
}

}  // namespace bg
