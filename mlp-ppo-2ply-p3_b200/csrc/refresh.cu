// refresh.cu -- update_legal_moves + get_observation for a whole batch in one call, with the encoders overlapped
// with K1's overflow tiers.
//
// Replaces BackgammonEnv.update_legal_moves (src/environment/backgammon_env.py:198-243: get_all_possible_moves +
// generate_all_board_features) and get_observation (:193-196) for N games.  K1's tier 0 (all SMs busy, issue
// bound) produces ~99 % of the rows; tiers 1/2 (the few huge doubles positions, one CTA each) are latency bound
// and leave the GPU nearly idle, so the HBM-bound encoder of the rows that are already final runs beside them
// on a second stream:
//
//   stream : [K1 tier 0] -> snapshot rows -> [tier 1] -> [tier 2] -> [K3 bf16 rows (snapshot, end)] -> join
//   side   : [K3 f32 observations] ........ -> [K3 bf16 rows [0, snapshot)] ------------------------^
#include "bg_device.cuh"
#include "bg_internal.h"

namespace {
struct ForkJoin {
    cudaEvent_t start = nullptr, tier0 = nullptr, side_done = nullptr;
    int device = -1;
};
thread_local ForkJoin g_fj[16];

ForkJoin* fork_join_events() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
    ForkJoin& f = g_fj[dev];
    if (f.device != dev) {
        if (cudaEventCreateWithFlags(&f.start, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&f.tier0, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&f.side_done, cudaEventDisableTiming) != cudaSuccess)
            return nullptr;
        f.device = dev;
    }
    return &f;
}

struct HookCtx {
    ForkJoin* fj;
    cudaStream_t stream, side;
    const int8_t* after52;
    const int8_t* row_players;
    long long cap_rows;
    const unsigned long long* rows_t0;
    uint16_t* feats;
    long long ld;
    int rc;
};

int after_tier0(void* user) {
    HookCtx* c = static_cast<HookCtx*>(user);
    cudaError_t e = cudaEventRecord(c->fj->tier0, c->stream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(c->side, c->fj->tier0, 0);
    if (e != cudaSuccess) return bg_set_error(e, "bg_update_legal_plays: fork");
    // rows [0, *rows_t0) are final: encode them beside tiers 1/2
    return bg::encode_bf16_launch(c->after52, c->row_players, 0, c->cap_rows, nullptr, c->rows_t0, c->feats, c->ld, c->side);
}
}  // namespace

extern "C" int bg_update_legal_plays(const int8_t* boards52, const int8_t* players, const int8_t* dice, long long N,
                                     int max_rows_per_board, int8_t* afterstates52, long long afterstate_capacity_rows,
                                     int8_t* row_players, int32_t* counts_true, int32_t* counts, long long* starts,
                                     unsigned long long* alloc_rows, int32_t* status, void* workspace,
                                     size_t workspace_bytes, uint16_t* features_bf16, long long features_ld,
                                     float* observations_f32, long long observations_ld, void* k1_begin_event,
                                     void* k1_end_event, void* side_stream, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_, side = (cudaStream_t)side_stream;
    if (N < 0) return bg_set_error_msg(BG_ERR_INVALID, "bg_update_legal_plays: negative batch");
    if (N == 0) return BG_OK;
    if (!counts || !row_players) return bg_set_error_msg(BG_ERR_INVALID, "bg_update_legal_plays: null counts / row_players");
    if (workspace_bytes < bg_movegen_workspace_bytes(N) || !workspace)
        return bg_set_error_msg(BG_ERR_INVALID, "bg_update_legal_plays: workspace too small");
    ForkJoin* fj = (side && side != stream) ? fork_join_events() : nullptr;
    if (!fj) {                                                     // serial form
        if (k1_begin_event) cudaEventRecord((cudaEvent_t)k1_begin_event, stream);
        int rc = bg::movegen_run(boards52, players, dice, N, 1, 0, 2, nullptr, max_rows_per_board, afterstates52,
                                 afterstate_capacity_rows, row_players, nullptr, counts_true, counts, starts, alloc_rows,
                                 status, workspace, workspace_bytes, stream);
        if (k1_end_event) cudaEventRecord((cudaEvent_t)k1_end_event, stream);
        if (rc != BG_OK) return rc;
        if (observations_f32) {
            rc = bg_encode_f32(boards52, players, 0, N, nullptr, observations_f32, observations_ld, stream);
            if (rc != BG_OK) return rc;
        }
        if (features_bf16)
            rc = bg::encode_bf16_launch(afterstates52, row_players, 0, afterstate_capacity_rows, nullptr, alloc_rows,
                                        features_bf16, features_ld, stream);
        return rc;
    }
    cudaError_t e = cudaEventRecord(fj->start, stream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(side, fj->start, 0);
    if (e != cudaSuccess) return bg_set_error(e, "bg_update_legal_plays: fork");
    // from here on `side` is forked: every exit, failures included, goes through the join below
    int rc = BG_OK;
    if (observations_f32)
        rc = bg_encode_f32(boards52, players, 0, N, nullptr, observations_f32, observations_ld, side);
    unsigned long long* rows_t0 = reinterpret_cast<unsigned long long*>(static_cast<unsigned char*>(workspace) + BG_WS_ROWS_AFTER_TIER0);
    HookCtx ctx{fj, stream, side, afterstates52, row_players, afterstate_capacity_rows, rows_t0, features_bf16, features_ld, BG_OK};
    bg::MovegenTier0Hook hook{rows_t0, after_tier0, &ctx};
    if (rc == BG_OK) {
        if (k1_begin_event) cudaEventRecord((cudaEvent_t)k1_begin_event, stream);
        rc = bg::movegen_run(boards52, players, dice, N, 1, 0, 2, nullptr, max_rows_per_board, afterstates52,
                             afterstate_capacity_rows, row_players, nullptr, counts_true, counts, starts, alloc_rows, status,
                             workspace, workspace_bytes, stream, features_bf16 ? &hook : nullptr);
        if (k1_end_event) cudaEventRecord((cudaEvent_t)k1_end_event, stream);
    }
    if (rc == BG_OK && features_bf16)                              // rows appended by tiers 1/2
        rc = bg::encode_bf16_launch(afterstates52, row_players, 0, afterstate_capacity_rows, rows_t0, alloc_rows,
                                    features_bf16, features_ld, stream);
    // join (always, so that `side` never runs ahead of the caller's stream order)
    e = cudaEventRecord(fj->side_done, side);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(stream, fj->side_done, 0);
    if (e != cudaSuccess && rc == BG_OK) rc = bg_set_error(e, "bg_update_legal_plays: join");
    return rc;
}

// ------------------------------------------------------------------------------------------------------------
// 2-ply inner step for M root afterstates: replies to all 21 rolls (K1, replicate mode) + leaf values (K4) + pass
// values (K4 with the flag flipped), with K4 on the rows that are final after K1's tier 0 overlapped with the
// overflow tiers, exactly like the encoder above.
namespace {
struct LeafCtx {
    ForkJoin* fj;
    cudaStream_t stream, side;
    const int8_t* replies52;
    const int8_t* row_players;
    long long cap_rows;
    const unsigned long long* rows_t0;
    const uint16_t* w1; const float* b1; const float* wv; float bv;
    float* leaf_values;
};
int leaves_after_tier0(void* user) {
    LeafCtx* c = static_cast<LeafCtx*>(user);
    cudaError_t e = cudaEventRecord(c->fj->tier0, c->stream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(c->side, c->fj->tier0, 0);
    if (e != cudaSuccess) return bg_set_error(e, "bg_twoply_replies_values: fork");
    return bg::mlp_value_launch(c->replies52, c->row_players, 0, 0, c->cap_rows, nullptr, c->rows_t0, c->w1, c->b1, c->wv, c->bv,
                                1, c->leaf_values, c->side);
}
}  // namespace

extern "C" int bg_twoply_replies_values(const int8_t* positions52, const int8_t* movers, long long M, int8_t* replies52,
                                        long long reply_capacity_rows, int8_t* row_players, int32_t* counts,
                                        long long* starts, unsigned long long* alloc_rows, int32_t* status,
                                        void* workspace, size_t workspace_bytes, const uint16_t* w1_bf16, const float* b1,
                                        const float* wv, float bv, float* leaf_values, float* pass_values,
                                        void* side_stream, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_, side = (cudaStream_t)side_stream;
    if (M < 0) return bg_set_error_msg(BG_ERR_INVALID, "bg_twoply_replies_values: negative batch");
    if (M == 0) return BG_OK;
    if (!counts || !row_players || !leaf_values || !pass_values || !workspace)
        return bg_set_error_msg(BG_ERR_INVALID, "bg_twoply_replies_values: null pointer");
    if (workspace_bytes < bg_movegen_workspace_bytes(M * 21))
        return bg_set_error_msg(BG_ERR_INVALID, "bg_twoply_replies_values: workspace too small");
    ForkJoin* fj = (side && side != stream) ? fork_join_events() : nullptr;
    cudaStream_t aux = fj ? side : stream;
    int rc = BG_OK;
    if (fj) {
        cudaError_t e = cudaEventRecord(fj->start, stream);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(side, fj->start, 0);
        if (e != cudaSuccess) return bg_set_error(e, "bg_twoply_replies_values: fork");
    }
    // value of the position with the opponent to move = the roll's value when the opponent has no reply
    // (after the fork every exit, failures included, goes through the join at the end)
    rc = bg::mlp_value_launch(positions52, movers, 0, 1, M, nullptr, nullptr, w1_bf16, b1, wv, bv, 0, pass_values, aux);
    unsigned long long* rows_t0 = reinterpret_cast<unsigned long long*>(static_cast<unsigned char*>(workspace) + BG_WS_ROWS_AFTER_TIER0);
    LeafCtx ctx{fj, stream, side, replies52, row_players, reply_capacity_rows, rows_t0, w1_bf16, b1, wv, bv, leaf_values};
    bg::MovegenTier0Hook hook{rows_t0, leaves_after_tier0, &ctx};
    if (rc == BG_OK)
        rc = bg::movegen_run(positions52, movers, nullptr, M * 21, 21, 1, 2, nullptr, 0, replies52, reply_capacity_rows,
                             row_players, nullptr, nullptr, counts, starts, alloc_rows, status, workspace, workspace_bytes,
                             stream, fj ? &hook : nullptr);
    if (rc == BG_OK)                                               // all rows (serial) / the rows appended by tiers 1, 2
        rc = bg::mlp_value_launch(replies52, row_players, 0, 0, reply_capacity_rows, fj ? rows_t0 : nullptr, alloc_rows, w1_bf16,
                                  b1, wv, bv, 1, leaf_values, stream);
    if (fj) {
        cudaError_t e = cudaEventRecord(fj->side_done, side);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(stream, fj->side_done, 0);
        if (e != cudaSuccess && rc == BG_OK) rc = bg_set_error(e, "bg_twoply_replies_values: join");
    }
    return rc;
}
