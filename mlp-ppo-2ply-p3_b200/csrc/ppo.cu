// ppo.cu -- rollout-side helpers of the PPO loop that are not GEMMs: discounted returns / GAE per game.
//
// Replaces BackgammonPPOAgent.compute_returns (src/agent/ppo_agent.py:206-216: R = r + gamma * R walking the memory
// backwards, R = 0 at a done) for a rollout stored step-major [T][N] on the device, one thread per game, and
// generalises it to GAE(lambda) with a bootstrap value:
//     delta_t = r_t + gamma * (1 - done_t) * V_{t+1} - V_t          (V_T = last_values, or 0)
//     A_t     = delta_t + gamma * lambda * (1 - done_t) * A_{t+1}
//     ret_t   = A_t + V_t
// lambda = 1 and last_values = NULL give exactly compute_returns for each game (ret_t = sum_k gamma^k r_{t+k} up
// to the end of the game or of the rollout).  The reference walks its memory -- steps of all envs interleaved,
// src/agent/train.py:64-66 -- as ONE sequence; that is reproduced by calling this with T = T*N, N = 1.
#include "bg_device.cuh"
#include "bg_internal.h"

namespace bg {

__global__ void __launch_bounds__(256) gae_kernel(const float* __restrict__ rewards, const uint8_t* __restrict__ dones,
                                                  const float* __restrict__ values, const float* __restrict__ last_values,
                                                  int T, long long N, float gamma, float lambda,
                                                  float* __restrict__ returns, float* __restrict__ advantages) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= N) return;
    float next_v = last_values ? last_values[g] : 0.0f;
    float adv = 0.0f;
    for (int t = T - 1; t >= 0; --t) {
        const long long i = (long long)t * N + g;
        const float nd = dones[i] ? 0.0f : 1.0f;
        const float v = values ? values[i] : 0.0f;
        // separate multiplies and adds in a fixed order: bit-reproducible against the CPU restatement
        const float delta = __fadd_rn(__fadd_rn(rewards[i], __fmul_rn(__fmul_rn(gamma, nd), next_v)), -v);
        adv = __fadd_rn(delta, __fmul_rn(__fmul_rn(__fmul_rn(gamma, lambda), nd), adv));
        if (advantages) advantages[i] = adv;
        if (returns) returns[i] = __fadd_rn(adv, v);
        next_v = v;
    }
}

}  // namespace bg

extern "C" int bg_gae(const float* rewards, const uint8_t* dones, const float* values, const float* last_values, int T,
                      long long N, float gamma, float lambda, float* returns, float* advantages, void* stream) {
    if (T < 0 || N < 0) return bg_set_error_msg(BG_ERR_INVALID, "bg_gae: negative size");
    if (T == 0 || N == 0) return BG_OK;
    if (!rewards || !dones || (!returns && !advantages)) return bg_set_error_msg(BG_ERR_INVALID, "bg_gae: null pointer");
    bg::gae_kernel<<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(rewards, dones, values, last_values, T, N,
                                                                                gamma, lambda, returns, advantages);
    return bg_set_error(cudaGetLastError(), "bg_gae: launch");
}
