// capi.cu -- error plumbing and the thin extern "C" entry points of libbg_b200.so (include/bg_b200.h).
#include <cstdio>
#include <cstring>
#include "bg_internal.h"

static thread_local char g_err[512] = "";

int bg_set_error(cudaError_t e, const char* where) {
    if (e == cudaSuccess) return BG_OK;
    snprintf(g_err, sizeof g_err, "%s: %s", where, cudaGetErrorString(e));
    return BG_ERR_CUDA;
}
int bg_set_error_msg(int code, const char* msg) {
    snprintf(g_err, sizeof g_err, "%s", msg);
    return code;
}
int bg_sm_count() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
            sms = 148;
    }
    return sms;
}

extern "C" const char* bg_last_error(void) { return g_err; }
extern "C" int bg_version(void) { return 100; }

extern "C" int bg_movegen_count(const int8_t* boards52, const int8_t* players, const int8_t* dice, long long B,
                                int32_t* counts_true, int32_t* status, void* workspace, size_t workspace_bytes,
                                void* stream) {
    if (B > 0 && !counts_true) return bg_set_error_msg(BG_ERR_INVALID, "bg_movegen_count: null counts");
    return bg::movegen_run(boards52, players, dice, B, 1, 0, 0, nullptr, 0, nullptr, 0, nullptr, nullptr, counts_true, nullptr, nullptr,
                           nullptr, status, workspace, workspace_bytes, (cudaStream_t)stream);
}
extern "C" int bg_movegen_write(const int8_t* boards52, const int8_t* players, const int8_t* dice, long long B,
                                const long long* offsets, int max_rows_per_board, int8_t* afterstates52,
                                long long afterstate_capacity_rows, int8_t* row_players, uint16_t* row_features_bf16,
                                int32_t* counts_true, int32_t* counts, int32_t* status, void* workspace, size_t workspace_bytes, void* stream) {
    return bg::movegen_run(boards52, players, dice, B, 1, 0, 1, offsets, max_rows_per_board, afterstates52,
                           afterstate_capacity_rows, row_players, row_features_bf16, counts_true, counts, nullptr, nullptr, status,
                           workspace,
                           workspace_bytes, (cudaStream_t)stream);
}
extern "C" int bg_movegen_slab(const int8_t* boards52, const int8_t* players, const int8_t* dice, long long B,
                               int max_rows_per_board, int8_t* afterstates52, long long afterstate_capacity_rows,
                               int8_t* row_players, uint16_t* row_features_bf16, int32_t* counts_true, int32_t* counts,
                               long long* starts, unsigned long long* alloc_rows, int32_t* status, void* workspace,
                               size_t workspace_bytes, void* stream) {
    if (B > 0 && !counts) return bg_set_error_msg(BG_ERR_INVALID, "bg_movegen_slab: null counts");
    return bg::movegen_run(boards52, players, dice, B, 1, 0, 2, nullptr, max_rows_per_board, afterstates52,
                           afterstate_capacity_rows, row_players, row_features_bf16, counts_true, counts, starts, alloc_rows, status,
                           workspace,
                           workspace_bytes, (cudaStream_t)stream);
}

extern "C" int bg_movegen_replies_slab(const int8_t* positions52, const int8_t* movers, long long M,
                                       int max_rows_per_board, int8_t* replies52, long long reply_capacity_rows,
                                       int8_t* row_players, uint16_t* row_features_bf16, int32_t* counts_true,
                                       int32_t* counts, long long* starts,
                                       unsigned long long* alloc_rows, int32_t* status, void* workspace,
                                       size_t workspace_bytes, void* stream) {
    if (M > 0 && !counts) return bg_set_error_msg(BG_ERR_INVALID, "bg_movegen_replies_slab: null counts");
    return bg::movegen_run(positions52, movers, nullptr, M * 21, 21, 1, 2, nullptr, max_rows_per_board, replies52,
                           reply_capacity_rows, row_players, row_features_bf16, counts_true, counts, starts, alloc_rows, status,
                           workspace, workspace_bytes, (cudaStream_t)stream);
}
