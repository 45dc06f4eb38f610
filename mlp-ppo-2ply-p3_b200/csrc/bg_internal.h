// bg_internal.h -- shared host-side helpers of libbg_b200.so (not part of the public ABI).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstddef>
#include "bg_b200.h"

// Per-warp shared-memory scratch of K1: boards per level (lists) before a position is handed to the
// large-scratch pass, and the size of that pass (DESIGN.md "K1 capacity").
#define BG_MOVEGEN_CAP_SMALL 128
#define BG_MOVEGEN_CAP_MID 512
#define BG_MOVEGEN_CAP_BIG 3072
#define BG_MOVEGEN_HASH_BIG 4096

// K1 workspace layout (bytes; bg_movegen_workspace_bytes): [0] work_ctr0, [4] overflow_ctr A, [8] work_ctr1, [12] overflow_ctr B,
// [16] work_ctr2, [24] list B's length after tier 0, [28] work counter of tier 2's second pass, [40] rows_after_tier0 (u64 snapshot of the slab allocator for the tier-0 fork, refresh.cu), [48] u64 left to the
// caller (the env's slab row allocator: zeroed by the same memset as the counters), [64 ..] overflow
// list A int32[B], then overflow list B int32[B]
#define BG_WS_ROWS_AFTER_TIER0 40
#define BG_WS_CALLER_ALLOC_ROWS 48   /* u64 free for the caller: the env keeps its slab row allocator here (zeroed with the counters) */
#define BG_WS_LISTS 64

int bg_set_error(cudaError_t e, const char* where);      // BG_OK if e == cudaSuccess else BG_ERR_CUDA
int bg_set_error_msg(int code, const char* msg);
int bg_sm_count();

namespace bg {
// Optional fork point of movegen_run in slab mode: after tier 0 has been launched, *rows_after_tier0 (device) receives
// the rows allocated so far and fn(user) is called on the host so the caller can enqueue consumers of those rows.
struct MovegenTier0Hook {
    unsigned long long* rows_after_tier0;
    int (*fn)(void* user);
    void* user;
};
// B work items; replicate == 21: item g = (position g/21, sorted roll g%21), player = players[g/21] ^ flip_player
int movegen_run(const int8_t* boards, const int8_t* players, const int8_t* dice, long long B, int replicate,
                int flip_player, int mode,
                const long long* offsets, int max_rows, int8_t* after, long long after_cap_rows,
                int8_t* row_players, uint16_t* row_feats, int32_t* counts_true, int32_t* counts, long long* starts, unsigned long long* alloc,
                int32_t* status, void* workspace, size_t ws_bytes, cudaStream_t stream,
                const MovegenTier0Hook* hook = nullptr);
// K3 launchers (encode.cu); rows [*row_begin_dev (0 if null), min(B, *n_rows_dev)) are encoded
int encode_bf16_launch(const int8_t* boards52, const int8_t* flags, int flag_all, long long B,
                       const unsigned long long* row_begin_dev, const unsigned long long* n_rows_dev, uint16_t* out,
                       long long ld, cudaStream_t stream);
// K4 launcher (mlp.cu); rows [*row_begin_dev (0 if null), min(B, *n_rows_dev)) are evaluated
int mlp_value_launch(const int8_t* boards52, const int8_t* flags, int flag_all, int flip_flags, long long B,
                     const unsigned long long* row_begin_dev, const unsigned long long* n_rows_dev,
                     const uint16_t* w1_bf16, const float* b1, const float* wv, float bv, int terminal_aware,
                     float* values, cudaStream_t stream);
// overflow tiers (movegen_team.cu): one CTA per position of worklist[0 .. *nwork_dev)
int movegen_team_mid(const int8_t* boards, const int8_t* players, const int8_t* dice, const unsigned int* nwork_dev,
                     const int32_t* worklist, int replicate, int flip_player, int mode, const long long* offsets,
                     int max_rows, int8_t* after, long long after_cap_rows, int8_t* row_players, uint16_t* row_feats, int32_t* counts_true,
                     int32_t* counts, long long* starts, unsigned long long* alloc, int32_t* status,
                     unsigned int* work_ctr, int32_t* overflow_list, unsigned int* overflow_ctr, cudaStream_t stream,
                     int team_threads_hint = 0);
// entries [*first_dev (0 if null), *nwork_dev) of the list
int movegen_team_big(const int8_t* boards, const int8_t* players, const int8_t* dice, const unsigned int* nwork_dev,
                     const unsigned int* first_dev, const int32_t* worklist, int replicate, int flip_player, int mode, const long long* offsets,
                     int max_rows, int8_t* after, long long after_cap_rows, int8_t* row_players, uint16_t* row_feats, int32_t* counts_true,
                     int32_t* counts, long long* starts, unsigned long long* alloc, int32_t* status,
                     unsigned int* work_ctr, cudaStream_t stream);
}
