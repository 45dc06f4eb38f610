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
#include <cuda_bf16.h>
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

// ---------------------------------------------------------------------------------------------------------------
// The loss of one PPO epoch and its gradient w.r.t. the network outputs, in one pass over the logits.
//
// Replaces the elementwise chain of BackgammonPPOAgent.update (src/agent/ppo_agent.py:271-299) -- log(mask + 1e-45),
// softmax, Categorical.log_prob / entropy, ratio, clipped surrogate, MSE, their sum -- and its autograd backward
// (about twenty passes over the [B,500] f32 logits in torch) by: one warp per sample reads its 500 logits once, and
// writes d loss / d logits once:
//     z = logits + (slot >= n ? log(1e-45) : 0);  lp = log_softmax(z);  p = exp(lp);  H = -sum p lp
//     r = exp(lp[a] - old_lp);  pl = -min(r A, clip(r, 1-eps, 1+eps) A);  vl = (v - R)^2
//     loss = mean(pl) + c_v mean(vl) - c_e mean(H)
//     d loss / d z_j = g (1[j == a] - p_j) + (c_e / B) p_j (lp_j + H),   g = -(A r / B) [r inside the clip range or r A < clip(r) A]
//     d loss / d v   = 2 c_v (v - R) / B
// (torch.min splits the gradient of a tie in halves, which sums to the same thing: inside the range both branches are r A.)
template <typename T> struct Vec4;
template <> struct Vec4<float> {
    static __device__ __forceinline__ void load(const float* p, float (&x)[4]) { float4 v = *reinterpret_cast<const float4*>(p); x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w; }
    static __device__ __forceinline__ void store(float* p, const float (&x)[4]) { *reinterpret_cast<float4*>(p) = make_float4(x[0], x[1], x[2], x[3]); }
};
template <> struct Vec4<__nv_bfloat16> {
    static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&x)[4]) {
        uint2 v = *reinterpret_cast<const uint2*>(p);
        x[0] = __uint_as_float(v.x << 16); x[1] = __uint_as_float(v.x & 0xFFFF0000u);
        x[2] = __uint_as_float(v.y << 16); x[3] = __uint_as_float(v.y & 0xFFFF0000u);
    }
    static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&x)[4]) {
        __nv_bfloat162 a = __floats2bfloat162_rn(x[0], x[1]), b = __floats2bfloat162_rn(x[2], x[3]);
        uint2 v; v.x = *reinterpret_cast<uint32_t*>(&a); v.y = *reinterpret_cast<uint32_t*>(&b);
        *reinterpret_cast<uint2*>(p) = v;
    }
};

template <typename T>
__global__ void __launch_bounds__(256) ppo_loss_grad_kernel(
    const T* __restrict__ logits, long long ld, const float* __restrict__ values, const int32_t* __restrict__ counts,
    const int32_t* __restrict__ actions, const float* __restrict__ old_logp, const float* __restrict__ adv,
    const float* __restrict__ returns, long long B, float eps_clip, float value_coef, float entropy_coef,
    T* __restrict__ dlogits, float* __restrict__ dvalues, float* __restrict__ sums) {
    constexpr float kMaskLog = -103.27893f;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long row = (long long)blockIdx.x * 8 + warp;
    __shared__ float s_part[8][3];
    float pl = 0.0f, vl = 0.0f, ent = 0.0f;
    if (row < B) {
        const int n = counts[row], a = actions[row];
        const float invB = 1.0f / (float)B;
        float z[16];
        const T* src = logits + row * ld + 16 * lane;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float x[4] = {0.0f, 0.0f, 0.0f, 0.0f};
            if (16 * lane + 4 * q < BG_ACTIONS) Vec4<T>::load(src + 4 * q, x);            // 500 = 4 * 125: whole quads only
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int i = 16 * lane + 4 * q + e;
                z[4 * q + e] = i < BG_ACTIONS ? (i >= n ? x[e] + kMaskLog : x[e]) : -INFINITY;
            }
        }
        float m = z[0];
#pragma unroll
        for (int k = 1; k < 16; ++k) m = fmaxf(m, z[k]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(kFull, m, o));
        float ssum = 0.0f;
#pragma unroll
        for (int k = 0; k < 16; ++k) ssum += __expf(z[k] - m);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ssum += __shfl_xor_sync(kFull, ssum, o);
        const float lse = m + __logf(ssum);
        float h = 0.0f, lpa = 0.0f;
        float p[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const float lp = z[k] - lse;                       // (-inf for the 12 padding slots)
            p[k] = __expf(lp);
            if (p[k] > 0.0f) h -= p[k] * lp;
            if (16 * lane + k == a) lpa = lp;
            z[k] = lp;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { h += __shfl_xor_sync(kFull, h, o); lpa += __shfl_xor_sync(kFull, lpa, o); }
        const float A = adv[row];
        const float r = __expf(lpa - old_logp[row]);
        const float rc = fminf(fmaxf(r, 1.0f - eps_clip), 1.0f + eps_clip);
        const float s1 = r * A, s2 = rc * A;
        const bool through = (r >= 1.0f - eps_clip && r <= 1.0f + eps_clip) || s1 < s2;
        const float g = through ? -A * r * invB : 0.0f;
        const float ce = entropy_coef * invB;
        T* dst = dlogits + row * ld + 16 * lane;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float x[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int k = 4 * q + e;
                const float pk = p[k];
                float d = -g * pk + (pk > 0.0f ? ce * pk * (z[k] + h) : 0.0f);
                if (16 * lane + k == a) d += g;
                x[e] = d;
            }
            if (16 * lane + 4 * q < BG_ACTIONS) Vec4<T>::store(dst + 4 * q, x);
        }
        const float v = values[row], dv = v - returns[row];
        if (lane == 0) { dvalues[row] = 2.0f * value_coef * dv * invB; pl = -fminf(s1, s2); vl = dv * dv; ent = h; }
    }
    if (lane == 0) { s_part[warp][0] = pl; s_part[warp][1] = vl; s_part[warp][2] = ent; }
    __syncthreads();
    if (threadIdx.x < 3) {
        float t = 0.0f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += s_part[w][threadIdx.x];
        atomicAdd(&sums[threadIdx.x], t);
    }
}

}  // namespace bg

extern "C" int bg_ppo_loss_grad(const void* logits, int logits_bf16, long long ld, const float* values, const int32_t* counts,
                                const int32_t* actions, const float* old_log_probs, const float* advantages,
                                const float* returns, long long B, float eps_clip, float value_coef, float entropy_coef,
                                void* dlogits, float* dvalues, float* sums, void* stream) {
    if (B < 0 || ld < BG_ACTIONS || (ld & 3)) return bg_set_error_msg(BG_ERR_INVALID, "bg_ppo_loss_grad: bad B or ld (need ld >= 500, multiple of 4)");
    if (B == 0) return BG_OK;
    if (!logits || !values || !counts || !actions || !old_log_probs || !advantages || !returns || !dlogits || !dvalues || !sums)
        return bg_set_error_msg(BG_ERR_INVALID, "bg_ppo_loss_grad: null pointer");
    const unsigned grid = (unsigned)((B + 7) / 8);
    if (logits_bf16)
        bg::ppo_loss_grad_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(
            (const __nv_bfloat16*)logits, ld, values, counts, actions, old_log_probs, advantages, returns, B, eps_clip, value_coef,
            entropy_coef, (__nv_bfloat16*)dlogits, dvalues, sums);
    else
        bg::ppo_loss_grad_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(
            (const float*)logits, ld, values, counts, actions, old_log_probs, advantages, returns, B, eps_clip, value_coef,
            entropy_coef, (float*)dlogits, dvalues, sums);
    return bg_set_error(cudaGetLastError(), "bg_ppo_loss_grad: launch");
}

extern "C" int bg_gae(const float* rewards, const uint8_t* dones, const float* values, const float* last_values, int T,
                      long long N, float gamma, float lambda, float* returns, float* advantages, void* stream) {
    if (T < 0 || N < 0) return bg_set_error_msg(BG_ERR_INVALID, "bg_gae: negative size");
    if (T == 0 || N == 0) return BG_OK;
    if (!rewards || !dones || (!returns && !advantages)) return bg_set_error_msg(BG_ERR_INVALID, "bg_gae: null pointer");
    bg::gae_kernel<<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(rewards, dones, values, last_values, T, N,
                                                                                gamma, lambda, returns, advantages);
    return bg_set_error(cudaGetLastError(), "bg_gae: launch");
}
