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

int bg_set_error(cudaError_t e, const char* where);      // BG_OK if e == cudaSuccess else BG_ERR_CUDA
int bg_set_error_msg(int code, const char* msg);
int bg_sm_count();

namespace bg {
// B work items; replicate == 21: item g = (position g/21, sorted roll g%21), player = players[g/21] ^ flip_player
int movegen_run(const int8_t* boards, const int8_t* players, const int8_t* dice, long long B, int replicate,
                int flip_player, int mode,
                const long long* offsets, int max_rows, int8_t* after, long long after_cap_rows,
                int8_t* row_players, uint16_t* row_feats, int32_t* counts_true, int32_t* counts, long long* starts, unsigned long long* alloc,
                int32_t* status, void* workspace, size_t ws_bytes, cudaStream_t stream);
// overflow tiers (movegen_team.cu): one CTA per position of worklist[0 .. *nwork_dev)
int movegen_team_mid(const int8_t* boards, const int8_t* players, const int8_t* dice, const unsigned int* nwork_dev,
                     const int32_t* worklist, int replicate, int flip_player, int mode, const long long* offsets,
                     int max_rows, int8_t* after, long long after_cap_rows, int8_t* row_players, uint16_t* row_feats, int32_t* counts_true,
                     int32_t* counts, long long* starts, unsigned long long* alloc, int32_t* status,
                     unsigned int* work_ctr, int32_t* overflow_list, unsigned int* overflow_ctr, cudaStream_t stream);
int movegen_team_big(const int8_t* boards, const int8_t* players, const int8_t* dice, const unsigned int* nwork_dev,
                     const int32_t* worklist, int replicate, int flip_player, int mode, const long long* offsets,
                     int max_rows, int8_t* after, long long after_cap_rows, int8_t* row_players, uint16_t* row_feats, int32_t* counts_true,
                     int32_t* counts, long long* starts, unsigned long long* alloc, int32_t* status,
                     unsigned int* work_ctr, cudaStream_t stream);
}
