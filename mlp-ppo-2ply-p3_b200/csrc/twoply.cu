// twoply.cu -- K5 support: segmented reductions of the 2-ply search.
//
// The reference's own 2-ply (src/moves/expect_minmax.py:1-206) is commented-out code; the search
// implemented here is the definition of SURVEY.md 8(c) on the reference's live primitives:
//   root (board, me, roll) -> afterstates A_i (K1)
//   A_i x 21 sorted opponent rolls (moves/get_all_dice_rolls.py:5-34) -> replies B_ij (K1, replicate mode)
//   leaf(B) = win reward if the opponent has borne off 15 else V(encode(B, flag=opp))          (K4)
//   score_i = +win_reward(A_i) if me has borne off 15
//           = - sum_r p_r * ( max_j leaf(B_ij)  or, with no reply, V(encode(A_i, flag=opp)) )  (here)
//   choice  = argmax_i score_i, lowest index on ties                                            (here)
#include "bg_device.cuh"
#include "bg_internal.h"

namespace bg {

__device__ __forceinline__ float win_reward52(const int8_t* b, int p) {      // backgammon_env.py:156-171,365-405
    const int o = p ^ 1;
    if (b[50 + o] != 0) return 1.0f;
    bool bgm = b[48 + o] > 0;
    const int8_t* orow = b + 24 * o;
    const int h0 = p == 0 ? 18 : 0;
#pragma unroll
    for (int i = 0; i < 6; ++i) bgm = bgm || orow[h0 + i] > 0;
    return bgm ? 2.0f : 1.5f;
}

// one warp per root afterstate
__global__ void __launch_bounds__(256) twoply_scores_kernel(
    const float* __restrict__ leaf_values, const long long* __restrict__ starts, const int32_t* __restrict__ counts,
    const float* __restrict__ pass_values, const int8_t* __restrict__ after52, const int8_t* __restrict__ movers,
    long long M, float* __restrict__ scores) {
    const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (i >= M) return;
    const int8_t* a = after52 + i * kBoardBytes;
    const int me = movers[i] & 1;
    if (a[50 + me] == 15) {
        if (lane == 0) scores[i] = win_reward52(a, me);
        return;
    }
    const float p1 = 1.0f / 36.0f, p2 = 2.0f / 36.0f;                      // get_all_dice_rolls.py:19-32
    float acc = 0.0f;
    int r = 0;
    for (int r0 = 1; r0 <= 6; ++r0)
        for (int r1 = r0; r1 <= 6; ++r1, ++r) {
            const int n = counts[i * 21 + r];
            float v;
            if (n == 0) v = pass_values[i];
            else {
                const float* lv = leaf_values + starts[i * 21 + r];
                v = -INFINITY;
                for (int j = lane; j < n; j += 32) v = fmaxf(v, lv[j]);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, o));
            }
            acc = __fadd_rn(acc, __fmul_rn(r0 == r1 ? p1 : p2, v));         // separate mul and add, fixed roll order (reproducible)
        }
    if (lane == 0) scores[i] = -acc;
}

// one warp per segment: best[b] = lowest index of the maximum, -1 for an empty segment
__global__ void __launch_bounds__(256) segment_argmax_kernel(const float* __restrict__ scores,
                                                             const long long* __restrict__ starts,
                                                             const int32_t* __restrict__ counts, long long B,
                                                             int32_t* __restrict__ best, float* __restrict__ best_score) {
    const long long b = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (b >= B) return;
    const int n = counts[b];
    const float* s = scores + starts[b];
    float bv = -INFINITY; int bi = 0x7FFFFFFF;
    for (int j = lane; j < n; j += 32) { float v = s[j]; if (v > bv) { bv = v; bi = j; } }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(kFull, bv, o); int oi = __shfl_xor_sync(kFull, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { best[b] = n > 0 ? bi : -1; if (best_score) best_score[b] = n > 0 ? bv : 0.0f; }
}

}  // namespace bg

using namespace bg;

extern "C" int bg_twoply_scores(const float* leaf_values, const long long* reply_starts, const int32_t* reply_counts,
                                const float* pass_values, const int8_t* after52, const int8_t* movers, long long M,
                                float* scores, void* stream) {
    if (M < 0) return bg_set_error_msg(BG_ERR_INVALID, "bg_twoply_scores: negative size");
    if (M == 0) return BG_OK;
    if (!leaf_values || !reply_starts || !reply_counts || !pass_values || !after52 || !movers || !scores)
        return bg_set_error_msg(BG_ERR_INVALID, "bg_twoply_scores: null pointer");
    unsigned grid = (unsigned)((M * 32 + 255) / 256);
    twoply_scores_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(leaf_values, reply_starts, reply_counts, pass_values,
                                                                 after52, movers, M, scores);
    return bg_set_error(cudaGetLastError(), "bg_twoply_scores: launch");
}

extern "C" int bg_segment_argmax(const float* scores, const long long* starts, const int32_t* counts, long long B,
                                 int32_t* best, float* best_score, void* stream) {
    if (B < 0) return bg_set_error_msg(BG_ERR_INVALID, "bg_segment_argmax: negative size");
    if (B == 0) return BG_OK;
    if (!scores || !starts || !counts || !best) return bg_set_error_msg(BG_ERR_INVALID, "bg_segment_argmax: null pointer");
    unsigned grid = (unsigned)((B * 32 + 255) / 256);
    segment_argmax_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(scores, starts, counts, B, best, best_score);
    return bg_set_error(cudaGetLastError(), "bg_segment_argmax: launch");
}
