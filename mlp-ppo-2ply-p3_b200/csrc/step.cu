// step.cu -- K2: fused step / reward / terminal / auto-reset / dice for N resident games.
//
// Replaces BackgammonEnv.step, reset, roll_dice, pass_turn, check_for_gammon, check_for_backgammon
// (src/environment/backgammon_env.py:78-191, 245-251, 365-405) and the per-env loop + auto-reset of
// VectorizedBackgammonEnv.step (src/environment/vec_bg_env.py:28-49).  One thread per game: the
// work per game is ~120 bytes of traffic and a few dozen instructions; the legal-move refresh
// (update_legal_moves, backgammon_env.py:198-243) is K1 (movegen.cu), launched right after.
#include "bg_device.cuh"
#include "bg_internal.h"

namespace bg {

// initial position (board/immutable_board.py:25-40) as board52 words
__device__ __constant__ uint32_t kInitialWords[kBoardWords] = {
    0x00000002u, 0x00000000u, 0x05000000u, 0x00000000u, 0x00050003u, 0x00000000u,   // P1: 0:2 11:5 16:3 18:5
    0x00000000u, 0x03000500u, 0x00000000u, 0x00000005u, 0x00000000u, 0x02000000u,   // P2: 5:5 7:3 12:5 23:2
    0x00000000u};

struct DiceSrc {
    unsigned long long seed, stream;
    const int8_t* ext; long long ext_len;
    uint32_t draw;
    int32_t* status;
    __device__ __forceinline__ void roll(int& d0, int& d1) {   // backgammon_env.py:245-246
        if (ext) {
            if ((long long)draw < ext_len) { d0 = ext[2 * draw]; d1 = ext[2 * draw + 1]; }
            else { d0 = 1; d1 = 2; atomicOr(status, BG_STATUS_DICE_EXHAUSTED); }
        } else {
            philox_dice(seed, stream, draw, d0, d1);
        }
        ++draw;
    }
};

// reset(): backgammon_env.py:78-113 (the alternating-starter store at :89-91 is dead: overwritten at :99-102)
__device__ __forceinline__ void new_game(const bg_env_state& st, long long g, DiceSrc& ds) {
    if (st.match_over[g]) { st.scores[2 * g] = 0; st.scores[2 * g + 1] = 0; st.match_over[g] = 0; }   // :79-82
    uint32_t* bw = reinterpret_cast<uint32_t*>(st.boards52 + g * kBoardBytes);
#pragma unroll
    for (int k = 0; k < kBoardWords; ++k) bw[k] = kInitialWords[k];                                       // :85
    int d0, d1;
    do { ds.roll(d0, d1); } while (d0 == d1);                                                             // :94-96
    st.players[g] = (int8_t)(d0 < d1 ? 1 : 0);                                                            // :99-102
    do { ds.roll(d0, d1); } while (d0 == d1);                                                             // :105-107
    st.dice[2 * g] = (int8_t)d0; st.dice[2 * g + 1] = (int8_t)d1;
}

__global__ void __launch_bounds__(256) reset_kernel(bg_env_state st, const uint8_t* __restrict__ mask, int32_t* status) {
    long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= st.n_games) return;
    if (mask && !mask[g]) return;
    DiceSrc ds{st.seed, st.stream_base + (unsigned long long)g,
               st.ext_dice ? st.ext_dice + 2 * st.ext_len * g : nullptr, st.ext_len, st.draws[g], status};
    new_game(st, g, ds);
    st.draws[g] = ds.draw;
    if (st.game_over) st.game_over[g] = 0;
}

// actions == NULL: the uniform-random policy of bg_random_actions inside the step (same Philox draw: act_seed, global game id, act_t),
// optionally written to actions_out -- one launch less per turn of a random-policy rollout
__global__ void __launch_bounds__(256) step_kernel(bg_env_state st, const int32_t* __restrict__ actions,
                                                   bg_step_out out, int32_t* status, unsigned long long act_seed, uint32_t act_t,
                                                   int32_t* __restrict__ actions_out) {
    long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= st.n_games) return;
    const int cur = st.players[g] & 1;
    const int n = st.counts[g];
    const long long row0 = st.starts[g];                      // (read up front: one dependent round trip less before the row loads)
    int a;
    if (actions) a = actions[g];
    else {
        a = n > 0 ? (int)philox_action(act_seed, st.stream_base + (unsigned long long)g, act_t, (uint32_t)n) : 0;
        if (actions_out) actions_out[g] = a;
    }
    float reward = 0.0f;
    int done = 0, winner = -1, gs = 0, flags = 0;
    DiceSrc ds{st.seed, st.stream_base + (unsigned long long)g,
               st.ext_dice ? st.ext_dice + 2 * st.ext_len * g : nullptr, st.ext_len, st.draws[g], status};
    if (st.no_auto_reset && st.game_over[g]) {                // step on a finished game: reset, reward 0, done (:119-121)
        flags = 4; done = 1;
        st.game_over[g] = 0;
        new_game(st, g, ds);
    } else if (n == 0) {                                             // pass: backgammon_env.py:124-140
        flags = 1;
        st.players[g] = (int8_t)(cur ^ 1);
        int d0, d1; ds.roll(d0, d1);
        st.dice[2 * g] = (int8_t)d0; st.dice[2 * g + 1] = (int8_t)d1;
    } else if (a < 0 || a >= n) {                             // invalid action: :143-149 (state unchanged)
        flags = 2; reward = -1.0f;
    } else {
        const uint32_t* row = reinterpret_cast<const uint32_t*>(st.afterstates52 + (row0 + a) * kBoardBytes);
        uint32_t w[kBoardWords];
#pragma unroll
        for (int k = 0; k < kBoardWords; ++k) w[k] = row[k];                                              // :152-153
        const uint32_t misc = w[12];
        const int off_cur = (misc >> (cur ? 24 : 16)) & 0xFF;
        if (off_cur == 15) {                                  // win: :156-181
            const int opp_off = (misc >> (cur ? 16 : 24)) & 0xFF;
            const int opp_bar = (misc >> (cur ? 0 : 8)) & 0xFF;
            // opponent men inside the winner's home board (:390-398)
            uint32_t in_home = cur == 0 ? ((w[6 + 4] & 0xFFFF0000u) | w[6 + 5])       // P2 men on 18..23
                                        : (w[0] | (w[1] & 0x0000FFFFu));              // P1 men on 0..5
            const bool backgammon = opp_off == 0 && (in_home != 0 || opp_bar > 0);    // :375-405
            gs = backgammon ? 3 : (opp_off == 0 ? 2 : 1);                             // :163-171, 365-373
            reward = gs == 3 ? 2.0f : (gs == 2 ? 1.5f : 1.0f);                        // :26-28
            winner = cur; done = 1;
            int sc = st.scores[2 * g + cur] + gs;                                     // :173
            st.scores[2 * g + cur] = sc;
            if (sc >= st.match_length) st.match_over[g] = 1;                          // :178-181
            if (st.no_auto_reset) {                           // BackgammonEnv: terminal board stays, winner to move (:152-153,190)
                uint32_t* bw = reinterpret_cast<uint32_t*>(st.boards52 + g * kBoardBytes);
#pragma unroll
                for (int k = 0; k < kBoardWords; ++k) bw[k] = w[k];
                st.game_over[g] = 1;
            } else {
                new_game(st, g, ds);                          // auto-reset: vec_bg_env.py:35-36
            }
        } else {                                              // :182-188
            uint32_t* bw = reinterpret_cast<uint32_t*>(st.boards52 + g * kBoardBytes);
#pragma unroll
            for (int k = 0; k < kBoardWords; ++k) bw[k] = w[k];
            st.players[g] = (int8_t)(cur ^ 1);
            int d0, d1; ds.roll(d0, d1);
            st.dice[2 * g] = (int8_t)d0; st.dice[2 * g + 1] = (int8_t)d1;
        }
    }
    st.draws[g] = ds.draw;
    if (out.rewards) out.rewards[g] = reward;
    if (out.dones) out.dones[g] = (uint8_t)done;
    if (out.info_player) out.info_player[g] = (int8_t)cur;
    if (out.winner) out.winner[g] = (int8_t)winner;
    if (out.game_score) out.game_score[g] = (int8_t)gs;
    if (out.flags) out.flags[g] = (uint8_t)flags;
}

__global__ void __launch_bounds__(256) random_actions_kernel(const int32_t* __restrict__ counts, long long N,
                                                             unsigned long long seed, unsigned long long stream_base,
                                                             uint32_t t, int32_t* __restrict__ actions) {
    long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= N) return;
    int n = counts[g];
    actions[g] = n > 0 ? (int32_t)philox_action(seed, stream_base + (unsigned long long)g, t, (uint32_t)n) : 0;
}

}  // namespace bg

using namespace bg;

static int check_state(const bg_env_state* st, const char* who) {
    if (!st) return bg_set_error_msg(BG_ERR_INVALID, who);
    if (st->n_games < 0) return bg_set_error_msg(BG_ERR_INVALID, who);
    if (st->n_games > 0 && (!st->boards52 || !st->players || !st->dice || !st->scores || !st->draws || !st->match_over))
        return bg_set_error_msg(BG_ERR_INVALID, who);
    if (st->ext_dice && st->ext_len <= 0) return bg_set_error_msg(BG_ERR_INVALID, who);
    if (st->no_auto_reset && st->n_games > 0 && !st->game_over) return bg_set_error_msg(BG_ERR_INVALID, who);
    return BG_OK;
}

extern "C" int bg_env_reset(const bg_env_state* st, const uint8_t* mask, int32_t* status, void* stream) {
    int rc = check_state(st, "bg_env_reset: bad state");
    if (rc != BG_OK) return rc;
    if (!status) return bg_set_error_msg(BG_ERR_INVALID, "bg_env_reset: null status");
    if (st->n_games == 0) return BG_OK;
    unsigned grid = (unsigned)((st->n_games + 255) / 256);
    reset_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*st, mask, status);
    return bg_set_error(cudaGetLastError(), "bg_env_reset: launch");
}

extern "C" int bg_env_step(const bg_env_state* st, const int32_t* actions, const bg_step_out* out, int32_t* status,
                           void* stream) {
    int rc = check_state(st, "bg_env_step: bad state");
    if (rc != BG_OK) return rc;
    if (!status || !out) return bg_set_error_msg(BG_ERR_INVALID, "bg_env_step: null status/out");
    if (st->n_games == 0) return BG_OK;
    if (!actions || !st->afterstates52 || !st->starts || !st->counts)
        return bg_set_error_msg(BG_ERR_INVALID, "bg_env_step: null actions or legal-play buffers");
    unsigned grid = (unsigned)((st->n_games + 255) / 256);
    step_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*st, actions, *out, status, 0ull, 0u, nullptr);
    return bg_set_error(cudaGetLastError(), "bg_env_step: launch");
}

extern "C" int bg_env_step_random(const bg_env_state* st, unsigned long long act_seed, uint32_t t, int32_t* actions_out,
                                  const bg_step_out* out, int32_t* status, void* stream) {
    int rc = check_state(st, "bg_env_step_random: bad state");
    if (rc != BG_OK) return rc;
    if (!status || !out) return bg_set_error_msg(BG_ERR_INVALID, "bg_env_step_random: null status/out");
    if (st->n_games == 0) return BG_OK;
    if (!st->afterstates52 || !st->starts || !st->counts)
        return bg_set_error_msg(BG_ERR_INVALID, "bg_env_step_random: null legal-play buffers");
    unsigned grid = (unsigned)((st->n_games + 255) / 256);
    step_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*st, nullptr, *out, status, act_seed, t, actions_out);
    return bg_set_error(cudaGetLastError(), "bg_env_step_random: launch");
}

extern "C" int bg_copy_actions_async(int32_t* actions_dev, const int32_t* host_actions, long long n, void* stream) {
    if (n < 0 || (n > 0 && (!actions_dev || !host_actions))) return bg_set_error_msg(BG_ERR_INVALID, "bg_copy_actions_async: bad args");
    if (n == 0) return BG_OK;
    return bg_set_error(cudaMemcpyAsync(actions_dev, host_actions, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice,
                                        (cudaStream_t)stream), "bg_copy_actions_async");
}

// The rollout's record of a step's inputs (RolloutBuffer: boards52, mover, legal-play count of every game) in ONE launch.
namespace bg {
__global__ void __launch_bounds__(256) record_state_kernel(const uint32_t* __restrict__ boards, const int8_t* __restrict__ players,
                                                           const int32_t* __restrict__ counts, long long N, uint32_t* __restrict__ boards_out,
                                                           int8_t* __restrict__ players_out, int32_t* __restrict__ counts_out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nw = N * kBoardWords;
    if (i < nw) boards_out[i] = boards[i];
    if (i < N) { players_out[i] = players[i]; counts_out[i] = counts[i]; }
}
}  // namespace bg
extern "C" int bg_record_state(const bg_env_state* st, int8_t* boards52_out, int8_t* players_out, int32_t* counts_out, void* stream) {
    int rc = check_state(st, "bg_record_state: bad state");
    if (rc != BG_OK) return rc;
    if (st->n_games == 0) return BG_OK;
    if (!st->counts || !boards52_out || !players_out || !counts_out) return bg_set_error_msg(BG_ERR_INVALID, "bg_record_state: null pointer");
    const long long nw = st->n_games * kBoardWords;
    record_state_kernel<<<(unsigned)((nw + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const uint32_t*>(st->boards52), st->players, st->counts, st->n_games, reinterpret_cast<uint32_t*>(boards52_out),
        players_out, counts_out);
    return bg_set_error(cudaGetLastError(), "bg_record_state: launch");
}

extern "C" int bg_random_actions(const int32_t* counts, long long N, unsigned long long seed,
                                 unsigned long long stream_base, uint32_t t, int32_t* actions, void* stream) {
    if (N < 0 || (N > 0 && (!counts || !actions))) return bg_set_error_msg(BG_ERR_INVALID, "bg_random_actions: bad args");
    if (N == 0) return BG_OK;
    unsigned grid = (unsigned)((N + 255) / 256);
    random_actions_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(counts, N, seed, stream_base, t, actions);
    return bg_set_error(cudaGetLastError(), "bg_random_actions: launch");
}
