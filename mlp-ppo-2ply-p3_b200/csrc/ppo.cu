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
#include <cstdlib>
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
    using Raw = float4;
    static __device__ __forceinline__ Raw zero() { return make_float4(0.0f, 0.0f, 0.0f, 0.0f); }
    static __device__ __forceinline__ Raw load_raw(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
    static __device__ __forceinline__ void unpack(const Raw& v, float (&x)[4]) { x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w; }
    static __device__ __forceinline__ void store(float* p, const float (&x)[4]) { *reinterpret_cast<float4*>(p) = make_float4(x[0], x[1], x[2], x[3]); }
};
template <> struct Vec4<__nv_bfloat16> {
    using Raw = uint2;
    static __device__ __forceinline__ Raw zero() { return make_uint2(0u, 0u); }
    static __device__ __forceinline__ Raw load_raw(const __nv_bfloat16* p) { return __ldg(reinterpret_cast<const uint2*>(p)); }
    static __device__ __forceinline__ void unpack(const Raw& v, float (&x)[4]) {
        x[0] = __uint_as_float(v.x << 16); x[1] = __uint_as_float(v.x & 0xFFFF0000u);
        x[2] = __uint_as_float(v.y << 16); x[3] = __uint_as_float(v.y & 0xFFFF0000u);
    }
    static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&x)[4]) {
        __nv_bfloat162 a = __floats2bfloat162_rn(x[0], x[1]), b = __floats2bfloat162_rn(x[2], x[3]);
        uint2 v; v.x = *reinterpret_cast<uint32_t*>(&a); v.y = *reinterpret_cast<uint32_t*>(&b);
        *reinterpret_cast<uint2*>(p) = v;
    }
};

// Element offset of (row r, column c) of a matrix with ld columns: row-major, or (blocked) the tile-blocked layout of
// ppo_gemm.cu -- tiles of 128 rows, inside a tile 16-byte chunks [c / 8][row][8].  c is a multiple of 4 wherever a 4-vector
// is accessed, so a vector never straddles a chunk.
__device__ __forceinline__ long long mat_off(long long r, int c, long long ld, int blocked) {
    return blocked ? (r >> 7) * (ld * 128) + (long long)(c >> 3) * 1024 + (r & 127) * 8 + (c & 7) : r * ld + c;
}

// Slot (column) kk = 0..15 of lane group `sub` in the packed kernel.  Row-major: 16 consecutive slots per lane.  Blocked: two whole
// 16-byte chunks per lane (slots 8 sub .. 8 sub + 7 and 64 + 8 sub ..), so that a warp's accesses cover full sectors of the
// tile-blocked layout (with 16 consecutive slots per lane every sector was fetched four times: the kernel ran 2x slower).
__device__ __forceinline__ int pk_col(int sub, int kk, int blocked) {
    return blocked ? ((kk >> 3) << 6) + 8 * sub + (kk & 7) : 16 * sub + kk;
}

// A warp per sample, persistent warps striding over the rows; lane l holds slots 128 q + 4 l + e (q, e < 4), so that every
// load / store instruction of the warp covers 256 (bf16) or 512 (f32) contiguous bytes, and the next row's logits and
// scalars are fetched before the current row is worked on (the kernel needs ~125 registers, i.e. 16 warps per SM: one
// row in flight per warp left it latency bound at 1.9 TB/s).  d loss / d logits also gives the gradient of
// action_head.bias (its column sums): each warp adds up its rows in registers, the warps of a CTA are combined through
// shared memory and one atomicAdd per column and CTA goes to dbias (nullable).
constexpr int kLossWarps = 8;

template <typename T> struct LossRow {
    typename Vec4<T>::Raw x[4];
    int n, a;
    float adv, old_logp, ret, v;
};
// Quads of 128 slots a row needs: the mask is a prefix (n legal slots), and a masked slot's probability
// exp(logit - 103.28 - lse) is below 1e-38 relative to the legal ones -- it changes neither the sums nor, in bf16 / f32,
// its own gradient (0) -- so only the quads that hold legal slots are read and computed (the policy kernel does the
// same).  Passes (n == 0: the reference's arithmetic gives a softmax over all 500 slots) and rows whose stored action
// is not a legal slot keep all four quads and the literal arithmetic.
__device__ __forceinline__ int loss_active_quads(int n, int a) { return (n <= 0 || a >= n) ? 4 : (n + 127) >> 7; }

// Rows the packed kernel takes: 1..128 legal slots (one quad) and a stored action inside the legal prefix.
__device__ __forceinline__ bool loss_row_is_packed(int n, int a) { return n >= 1 && n <= 128 && a >= 0 && a < n; }
__device__ __forceinline__ float loss_scalar(const float* p) { return __ldg(p); }
__device__ __forceinline__ float loss_scalar(const __nv_bfloat16* p) { return __bfloat162float(*p); }

template <typename T>
__device__ __forceinline__ void loss_row_load(LossRow<T>& r, long long row, int n, int a, int lane, const T* __restrict__ logits,
                                              long long ld, const float* __restrict__ values, const float* __restrict__ old_logp,
                                              const float* __restrict__ adv, const float* __restrict__ returns, int blocked) {
    const int nq = loss_active_quads(n, a);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int i0 = 128 * q + 4 * lane;
        const bool need = q < nq || (q == 3 && !values && lane == 29);               // lane 29 of quad 3: column 500 = the value
        r.x[q] = (need && i0 < ld) ? Vec4<T>::load_raw(logits + mat_off(row, i0, ld, blocked)) : Vec4<T>::zero();
    }
    r.n = n; r.a = a;
    r.adv = __ldg(adv + row); r.old_logp = __ldg(old_logp + row); r.ret = __ldg(returns + row);
    r.v = values ? __ldg(values + row) : 0.0f;
}

// The common rows (loss_row_is_packed: 94 % of a self-play batch, mean 18 legal slots), FOUR per warp: eight lanes per row,
// lane s of a row holding slots 16 s .. 16 s + 15 of the only quad that has legal slots, reductions over 8 lanes; columns
// 128 .. ld-1 of dlogits are written as zeros (column 500 = d loss / d value when the value rides in the logits).  One row per
// warp spent ~470 warp-instructions on a row whatever its mask; here it is ~85.  Same arithmetic as the general kernel below.
template <typename T> struct LossPacked {
    typename Vec4<T>::Raw x[4];
    int n, a;
    float adv, old_logp, ret, v;
};
template <typename T>
__device__ __forceinline__ void loss_packed_load(LossPacked<T>& d, long long r, long long B, int sub, const T* __restrict__ logits, long long ld,
                                                 const float* __restrict__ values, const int32_t* __restrict__ counts,
                                                 const int32_t* __restrict__ actions, const float* __restrict__ old_logp,
                                                 const float* __restrict__ adv, const float* __restrict__ returns, int value_col, int blocked) {
    d.n = 0; d.a = 0; d.adv = 0.0f; d.old_logp = 0.0f; d.ret = 0.0f; d.v = 0.0f;
#pragma unroll
    for (int k = 0; k < 4; ++k) d.x[k] = Vec4<T>::zero();
    if (r < B) {
#pragma unroll
        for (int k = 0; k < 4; ++k) d.x[k] = Vec4<T>::load_raw(logits + mat_off(r, pk_col(sub, 4 * k, blocked), ld, blocked));     // (read whatever the row's class: no dependent load)
        d.n = __ldg(counts + r); d.a = __ldg(actions + r);
        d.adv = __ldg(adv + r); d.old_logp = __ldg(old_logp + r); d.ret = __ldg(returns + r);
        d.v = values ? __ldg(values + r) : loss_scalar(logits + mat_off(r, value_col, ld, blocked));
    }
}

template <typename T>
__global__ void __launch_bounds__(kLossWarps * 32) ppo_loss_grad_packed_kernel(
    const T* __restrict__ logits, long long ld, const float* __restrict__ values, const int32_t* __restrict__ counts,
    const int32_t* __restrict__ actions, const float* __restrict__ old_logp, const float* __restrict__ adv,
    const float* __restrict__ returns, long long B, float eps_clip, float value_coef, float entropy_coef,
    T* __restrict__ dlogits, float* __restrict__ dvalues, float* __restrict__ dbias, float* __restrict__ sums, int prezeroed,
    int value_col, long long B_norm, int blocked) {
    // value_col: column of the logits / dlogits row that carries the value head when values == NULL (500 in the 512-wide
    // layout, 128 in the 144-wide class A layout of ppo_gemm.cu); B_norm: the batch size the means are taken over (this call
    // may cover only a part of the batch)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, sub = lane & 7, grp = lane >> 3;
    __shared__ float s_part[kLossWarps][3];
    __shared__ float s_col[kLossWarps][129];                         // 128 slot columns + the value column
    float pl = 0.0f, vl = 0.0f, ent = 0.0f, vsum = 0.0f;
    float colsum[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) colsum[k] = 0.0f;
    const float invB = 1.0f / (float)B_norm;
    const float ce = entropy_coef * invB;
    const long long S = (long long)gridDim.x * kLossWarps;          // persistent warps stride over groups of 4 rows
    long long g4 = (long long)blockIdx.x * kLossWarps + warp;
    LossPacked<T> nx;
    loss_packed_load(nx, g4 * 4 + grp, B, sub, logits, ld, values, counts, actions, old_logp, adv, returns, value_col, blocked);
#pragma unroll 1
    for (; g4 * 4 < B; g4 += S) {
        const LossPacked<T> cur = nx;
        loss_packed_load(nx, (g4 + S) * 4 + grp, B, sub, logits, ld, values, counts, actions, old_logp, adv, returns, value_col, blocked);
        const long long row = g4 * 4 + grp;
        const bool act = row < B && loss_row_is_packed(cur.n, cur.a);
        const int n = act ? cur.n : 1, a = act ? cur.a : 0;         // (idle lanes compute on a finite dummy row)
        float z[16];
        float m = -INFINITY;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float x[4];
            Vec4<T>::unpack(cur.x[k], x);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int i = pk_col(sub, 4 * k + e, blocked);
                z[4 * k + e] = i < n ? (act ? x[e] : 0.0f) : -INFINITY;
                m = fmaxf(m, z[4 * k + e]);
            }
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(kFull, m, o));
        float ssum = 0.0f;
#pragma unroll
        for (int k = 0; k < 16; ++k) ssum += __expf(z[k] - m);
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) ssum += __shfl_xor_sync(kFull, ssum, o);
        const float lse = m + __logf(ssum);
        float h = 0.0f, lpa = 0.0f;
        float p[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const float lp = z[k] - lse;
            p[k] = __expf(lp);
            if (p[k] > 0.0f) h -= p[k] * lp;
            if (pk_col(sub, k, blocked) == a) lpa = lp;
            z[k] = lp;
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) { h += __shfl_xor_sync(kFull, h, o); lpa += __shfl_xor_sync(kFull, lpa, o); }
        const float A = cur.adv;
        const float r = __expf(lpa - cur.old_logp);
        const float rc = fminf(fmaxf(r, 1.0f - eps_clip), 1.0f + eps_clip);
        const float s1 = r * A, s2 = rc * A;
        const bool through = (r >= 1.0f - eps_clip && r <= 1.0f + eps_clip) || s1 < s2;
        const float g = through ? -A * r * invB : 0.0f;
        const float dv = cur.v - cur.ret;
        const float dvalue = 2.0f * value_coef * dv * invB;
        if (act) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float x[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int kk = 4 * k + e;
                    const float pk = p[kk];
                    float d = -g * pk + (pk > 0.0f ? ce * pk * (z[kk] + h) : 0.0f);
                    if (pk_col(sub, kk, blocked) == a) d += g;
                    x[e] = d;
                    colsum[kk] += d;
                }
                Vec4<T>::store(dlogits + mat_off(row, pk_col(sub, 4 * k, blocked), ld, blocked), x);
            }
            // quads 1..3 and the padding of the GEMM's N: exact zeros; column 500 carries d loss / d value when the value rides there
            // (prezeroed: the caller guarantees zeros there already -- only the value column is written)
            if (!prezeroed) {
                for (int c = 128 + 4 * sub; c < ld; c += 32) {
                    float x[4] = {0.0f, 0.0f, 0.0f, 0.0f};
                    if (!values && c == value_col) { x[0] = dvalue; vsum += dvalue; }
                    Vec4<T>::store(dlogits + mat_off(row, c, ld, blocked), x);
                }
            } else if (!values && sub >= 4) {
                // the value column's whole 32-byte sector (columns 496 .. 511: a lone 8-byte store would cost a read-modify-write)
                const int c = BG_ACTIONS - 4 + 4 * (sub - 4);
                float x[4] = {0.0f, 0.0f, 0.0f, 0.0f};
                if (sub == 5) { x[0] = dvalue; vsum += dvalue; }
                if (c < ld) Vec4<T>::store(dlogits + mat_off(row, c, ld, blocked), x);
            }
            if (sub == 0) {
                if (dvalues) dvalues[row] = dvalue;
                pl -= fminf(s1, s2); vl += dv * dv; ent += h;
            }
        }
    }
    // the four rows of a warp hold the same columns: add them up, then the warps of the CTA, then one atomic per column
#pragma unroll
    for (int o = 8; o <= 16; o <<= 1) {
#pragma unroll
        for (int k = 0; k < 16; ++k) colsum[k] += __shfl_xor_sync(kFull, colsum[k], o);
        pl += __shfl_xor_sync(kFull, pl, o); vl += __shfl_xor_sync(kFull, vl, o); ent += __shfl_xor_sync(kFull, ent, o);
        vsum += __shfl_xor_sync(kFull, vsum, o);
    }
    if (lane == 0) { s_part[warp][0] = pl; s_part[warp][1] = vl; s_part[warp][2] = ent; }
    if (grp == 0) {
#pragma unroll
        for (int k = 0; k < 16; ++k) s_col[warp][pk_col(sub, k, blocked)] = colsum[k];
    }
    if (lane == ((value_col - 128) & 31) >> 2) s_col[warp][128] = vsum;   // the lane group that wrote the value column (500: group 5, 128: group 0)
    __syncthreads();
    if (threadIdx.x < 3) {
        float t = 0.0f;
#pragma unroll
        for (int w = 0; w < kLossWarps; ++w) t += s_part[w][threadIdx.x];
        atomicAdd(&sums[threadIdx.x], t);
    }
    if (dbias && threadIdx.x < 129) {
        float t = 0.0f;
#pragma unroll
        for (int w = 0; w < kLossWarps; ++w) t += s_col[w][threadIdx.x];
        if (threadIdx.x < 128) atomicAdd(&dbias[threadIdx.x], t);
        else if (!values) atomicAdd(&dbias[BG_ACTIONS], t);
    }
}

template <typename T>
__global__ void __launch_bounds__(kLossWarps * 32) ppo_loss_grad_kernel(
    const T* __restrict__ logits, long long ld, const float* __restrict__ values, const int32_t* __restrict__ counts,
    const int32_t* __restrict__ actions, const float* __restrict__ old_logp, const float* __restrict__ adv,
    const float* __restrict__ returns, long long B, float eps_clip, float value_coef, float entropy_coef,
    T* __restrict__ dlogits, float* __restrict__ dvalues, float* __restrict__ dbias, float* __restrict__ sums, int all_rows,
    long long B_norm, int blocked) {
    constexpr float kMaskLog = -103.27893f;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ float s_part[kLossWarps][3];
    __shared__ float s_col[kLossWarps][512];
    float pl = 0.0f, vl = 0.0f, ent = 0.0f;
    float colsum[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) colsum[k] = 0.0f;
    const float invB = 1.0f / (float)B_norm;
    const float ce = entropy_coef * invB;
    const long long S = (long long)gridDim.x * kLossWarps;          // persistent warps stride over blocks of 32 rows
    // This kernel handles the rows the packed kernel leaves out (loss_row_is_packed: passes, more than 128 legal slots, an
    // action outside the legal prefix -- a few per cent of a self-play batch): every lane looks at one row of the block, the
    // warp then works through the flagged rows one after the other, the next flagged row of the block being fetched while
    // the current one is worked on.
    long long blk = (long long)blockIdx.x * kLossWarps + warp;
    int nl_next = 0, al_next = 0;                                   // counts / actions one block ahead
    if (blk * 32 + lane < B) { nl_next = __ldg(counts + blk * 32 + lane); al_next = __ldg(actions + blk * 32 + lane); }
#pragma unroll 1
    for (; blk * 32 < B; blk += S) {
      const long long rl = blk * 32 + lane;
      const int nl = nl_next, al = al_next;
      if ((blk + S) * 32 + lane < B) { nl_next = __ldg(counts + (blk + S) * 32 + lane); al_next = __ldg(actions + (blk + S) * 32 + lane); }
      unsigned todo = __ballot_sync(kFull, rl < B && (all_rows || !loss_row_is_packed(nl, al)));
      LossRow<T> nx;
      if (todo) {
          const int j = __ffs(todo) - 1;
          loss_row_load(nx, blk * 32 + j, __shfl_sync(kFull, nl, j), __shfl_sync(kFull, al, j), lane, logits, ld, values, old_logp, adv, returns, blocked);
      }
#pragma unroll 1
      while (todo) {
        const long long row = blk * 32 + (__ffs(todo) - 1);
        todo &= todo - 1;
        const LossRow<T> cur = nx;
        if (todo) {
            const int j = __ffs(todo) - 1;
            loss_row_load(nx, blk * 32 + j, __shfl_sync(kFull, nl, j), __shfl_sync(kFull, al, j), lane, logits, ld, values, old_logp, adv, returns, blocked);
        }
        const int n = cur.n, a = cur.a;
        const int nq = loss_active_quads(n, a);
        float z[16], xv = 0.0f;
        float m = -INFINITY;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int i0 = 128 * q + 4 * lane;
            float x[4];
            Vec4<T>::unpack(cur.x[q], x);                                               // 500 = 4 * 125: whole quads only
            if (q == 3) xv = x[0];                                                      // lane 29: column 500
            if (q < nq) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int i = i0 + e;
                    z[4 * q + e] = i < BG_ACTIONS ? (i >= n ? x[e] + kMaskLog : x[e]) : -INFINITY;
                    m = fmaxf(m, z[4 * q + e]);
                }
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) z[4 * q + e] = -INFINITY;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(kFull, m, o));
        float ssum = 0.0f;
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (q < nq) {
#pragma unroll
                for (int e = 0; e < 4; ++e) ssum += __expf(z[4 * q + e] - m);
            }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ssum += __shfl_xor_sync(kFull, ssum, o);
        const float lse = m + __logf(ssum);
        float h = 0.0f, lpa = 0.0f;
        float p[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (q < nq) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int k = 4 * q + e;
                    const float lp = z[k] - lse;                   // (-inf for the 12 padding slots)
                    p[k] = __expf(lp);
                    if (p[k] > 0.0f) h -= p[k] * lp;
                    if (128 * q + 4 * lane + e == a) lpa = lp;
                    z[k] = lp;
                }
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) p[4 * q + e] = 0.0f;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { h += __shfl_xor_sync(kFull, h, o); lpa += __shfl_xor_sync(kFull, lpa, o); }
        const float A = cur.adv;
        const float r = __expf(lpa - cur.old_logp);
        const float rc = fminf(fmaxf(r, 1.0f - eps_clip), 1.0f + eps_clip);
        const float s1 = r * A, s2 = rc * A;
        const bool through = (r >= 1.0f - eps_clip && r <= 1.0f + eps_clip) || s1 < s2;
        const float g = through ? -A * r * invB : 0.0f;
        // value head: its own array, or (values == NULL) column 500 of the logits -- the value head computed as row 500 of
        // the action head's GEMM; its gradient then goes to column 500 of dlogits and rides through the backward GEMMs
        const float v = values ? cur.v : __shfl_sync(kFull, xv, 29);
        const float dv = v - cur.ret;
        const float dvalue = 2.0f * value_coef * dv * invB;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int i0 = 128 * q + 4 * lane;
            float x[4] = {0.0f, 0.0f, 0.0f, 0.0f};
            if (q < nq) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int k = 4 * q + e;
                    const float pk = p[k];
                    float d = -g * pk + (pk > 0.0f ? ce * pk * (z[k] + h) : 0.0f);
                    if (i0 + e == a) d += g;
                    x[e] = d;
                }
            }
            if (q == 3 && !values && i0 == BG_ACTIONS) x[0] = dvalue;
#pragma unroll
            for (int e = 0; e < 4; ++e) colsum[4 * q + e] += x[e];
            if (i0 < ld) Vec4<T>::store(dlogits + mat_off(row, i0, ld, blocked), x);          // masked quads and slots 500 .. ld-1 (padding of the GEMM's N): exact zeros
        }
        if (lane == 0) {
            if (dvalues) dvalues[row] = dvalue;
            pl -= fminf(s1, s2); vl += dv * dv; ent += h;
        }
      }
    }
    if (lane == 0) { s_part[warp][0] = pl; s_part[warp][1] = vl; s_part[warp][2] = ent; }
    if (dbias) {
#pragma unroll
        for (int k = 0; k < 16; ++k) s_col[warp][128 * (k >> 2) + 4 * lane + (k & 3)] = colsum[k];
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        float t = 0.0f;
#pragma unroll
        for (int w = 0; w < kLossWarps; ++w) t += s_part[w][threadIdx.x];
        atomicAdd(&sums[threadIdx.x], t);
    }
    if (dbias) {
        const int ncol = ld < 512 ? (int)ld : 512;
        for (int c = threadIdx.x; c < ncol; c += kLossWarps * 32) {
            float t = 0.0f;
#pragma unroll
            for (int w = 0; w < kLossWarps; ++w) t += s_col[w][c];
            atomicAdd(&dbias[c], t);
        }
    }
}

}  // namespace bg

extern "C" int bg_ppo_loss_grad(const void* logits, int flags, long long ld, const float* values, const int32_t* counts,
                                const int32_t* actions, const float* old_log_probs, const float* advantages,
                                const float* returns, long long B, float eps_clip, float value_coef, float entropy_coef,
                                void* dlogits, float* dvalues, float* dbias, float* sums, void* stream) {
    if (B < 0 || ld < BG_ACTIONS || (ld & 3)) return bg_set_error_msg(BG_ERR_INVALID, "bg_ppo_loss_grad: bad B or ld (need ld >= 500, multiple of 4)");
    if (B == 0) return BG_OK;
    if (!logits || !counts || !actions || !old_log_probs || !advantages || !returns || !dlogits || !sums || (values && !dvalues))
        return bg_set_error_msg(BG_ERR_INVALID, "bg_ppo_loss_grad: null pointer");
    if (!values && ld <= BG_ACTIONS) return bg_set_error_msg(BG_ERR_INVALID, "bg_ppo_loss_grad: values == NULL needs the value in column 500 (ld >= 504)");
    // two launches: the packed kernel for the common rows (four per warp), the general kernel for the others (BG_LOSS_GENERAL_ONLY=1:
    // every row through the general kernel -- for tests / comparison)
    const int logits_bf16 = flags & BG_LOSS_LOGITS_BF16, prezeroed = (flags & BG_LOSS_DLOGITS_PREZEROED) ? 1 : 0;
    static const int general_only = getenv("BG_LOSS_GENERAL_ONLY") ? atoi(getenv("BG_LOSS_GENERAL_ONLY")) : 0;
    const long long resident = (long long)bg_sm_count() * 2;                   // ~125 registers x 256 threads: two CTAs per SM
    const long long need_p = (B + 4 * bg::kLossWarps - 1) / (4 * bg::kLossWarps), need_g = (B + 32 * bg::kLossWarps - 1) / (32 * bg::kLossWarps);
    const unsigned grid_p = (unsigned)(need_p < resident ? need_p : resident), grid_g = (unsigned)(need_g < resident ? need_g : resident);
    cudaStream_t st = (cudaStream_t)stream;
    if (logits_bf16) {
        if (!general_only)
            bg::ppo_loss_grad_packed_kernel<__nv_bfloat16><<<grid_p, bg::kLossWarps * 32, 0, st>>>(
                (const __nv_bfloat16*)logits, ld, values, counts, actions, old_log_probs, advantages, returns, B, eps_clip, value_coef,
                entropy_coef, (__nv_bfloat16*)dlogits, dvalues, dbias, sums, prezeroed, BG_ACTIONS, B, 0);
        bg::ppo_loss_grad_kernel<__nv_bfloat16><<<grid_g, bg::kLossWarps * 32, 0, st>>>(
            (const __nv_bfloat16*)logits, ld, values, counts, actions, old_log_probs, advantages, returns, B, eps_clip, value_coef,
            entropy_coef, (__nv_bfloat16*)dlogits, dvalues, dbias, sums, general_only, B, 0);
    } else {
        if (!general_only)
            bg::ppo_loss_grad_packed_kernel<float><<<grid_p, bg::kLossWarps * 32, 0, st>>>(
                (const float*)logits, ld, values, counts, actions, old_log_probs, advantages, returns, B, eps_clip, value_coef,
                entropy_coef, (float*)dlogits, dvalues, dbias, sums, prezeroed, BG_ACTIONS, B, 0);
        bg::ppo_loss_grad_kernel<float><<<grid_g, bg::kLossWarps * 32, 0, st>>>(
            (const float*)logits, ld, values, counts, actions, old_log_probs, advantages, returns, B, eps_clip, value_coef,
            entropy_coef, (float*)dlogits, dvalues, dbias, sums, general_only, B, 0);
    }
    return bg_set_error(cudaGetLastError(), "bg_ppo_loss_grad: launch");
}

// The loss over a batch that ppo_gemm.cu keeps sorted by class, on its tile-blocked buffers: n_a class A rows in the
// 144-column layout (128 slots, value head in column 128) through the packed kernel, n_b class B rows in the 512-column layout
// (value head in column 500) through the general kernel; the per-sample arrays hold class A at [0, n_a) and class B at
// [b_offset, b_offset + n_b) (b_offset = n_a rounded up to a whole tile); every mean is taken over n_a + n_b.
extern "C" int bg_ppo_loss_grad_classes(const void* logits_a, void* dlogits_a, const void* logits_b, void* dlogits_b, long long n_a,
                                        long long n_b, long long b_offset, const int32_t* counts, const int32_t* actions,
                                        const float* old_log_probs, const float* advantages, const float* returns, float eps_clip,
                                        float value_coef, float entropy_coef, float* dbias, float* sums, int class_a_done, void* stream) {
    // class_a_done: class A has been handled by bg_ppo_logits_loss_a (the loss fused into the GEMM's epilogue); it still counts in the means
    if (n_a < 0 || n_b < 0 || b_offset < n_a) return bg_set_error_msg(BG_ERR_INVALID, "bg_ppo_loss_grad_classes: bad sizes");
    const long long B = n_a + n_b;
    if (B == 0) return BG_OK;
    if (!counts || !actions || !old_log_probs || !advantages || !returns || !sums || (n_a > 0 && !class_a_done && (!logits_a || !dlogits_a)) ||
        (n_b > 0 && (!logits_b || !dlogits_b)))
        return bg_set_error_msg(BG_ERR_INVALID, "bg_ppo_loss_grad_classes: null pointer");
    const long long resident = (long long)bg_sm_count() * 2;
    cudaStream_t st = (cudaStream_t)stream;
    if (n_a > 0 && !class_a_done) {
        const long long need = (n_a + 4 * bg::kLossWarps - 1) / (4 * bg::kLossWarps);
        bg::ppo_loss_grad_packed_kernel<__nv_bfloat16><<<(unsigned)(need < resident ? need : resident), bg::kLossWarps * 32, 0, st>>>(
            (const __nv_bfloat16*)logits_a, 144, nullptr, counts, actions, old_log_probs, advantages, returns, n_a, eps_clip, value_coef,
            entropy_coef, (__nv_bfloat16*)dlogits_a, nullptr, dbias, sums, 0, 128, B, 1);
    }
    if (n_b > 0) {
        const long long need = (n_b + 32 * bg::kLossWarps - 1) / (32 * bg::kLossWarps);
        bg::ppo_loss_grad_kernel<__nv_bfloat16><<<(unsigned)(need < resident ? need : resident), bg::kLossWarps * 32, 0, st>>>(
            (const __nv_bfloat16*)logits_b, 512, nullptr, counts + b_offset, actions + b_offset, old_log_probs + b_offset,
            advantages + b_offset, returns + b_offset, n_b, eps_clip, value_coef, entropy_coef, (__nv_bfloat16*)dlogits_b, nullptr, dbias,
            sums, 1, B, 1);
    }
    return bg_set_error(cudaGetLastError(), "bg_ppo_loss_grad_classes: launch");
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
