// ppo_gemm.cu -- N2: the GEMMs of the PPO update on the tcgen05 tensor cores, hand-written (no cuBLAS on the update path).
//
// Replaces, for one epoch over B samples, the linear algebra of BackgammonPPOAgent.update (src/agent/ppo_agent.py:268-305:
// policy_network forward, loss.backward()) around the loss kernels of ppo.cu:
//     h      = relu(x W1p^T)                    bg_ppo_gemm_nt(HIDDEN)        x = K3's rows, column 198 = 1 (bias)
//     logits = h Wap^T + b                      bg_ppo_gemm_nt(LOGITS_A / _B)
//     dpre   = (dlogits Wap) * [h > 0]          bg_ppo_gemm_nt(DPRE_A / _B)
//     dWap  += dlogits^T h                      bg_ppo_gemm_tn(GRAD_WA_A / _B)
//     dW1p  += dpre^T x                         bg_ppo_gemm_tn(GRAD_W1)
// The work follows the action mask, as in the policy kernel: the caller sorts the samples into class A (1..128 legal slots
// and the stored action among them: ~94 % of a self-play batch, mean 18 legal slots) and class B (passes -- the reference's
// arithmetic is a softmax over all 500 slots -- and > 128 slots).  Class A rows only ever touch action slots 0..127, so
// their logits / dlogits are 144 columns wide (128 slots, the value head in column 128, zero padding) instead of 512:
// a quarter of the head's FLOPs and of its HBM traffic.  Class B rows use the full 512-column layout (value in column 500).
//
// DATA LAYOUT.  Every activation matrix lives in HBM in the very layout the tensor cores read from shared memory: tiles
// of 128 rows, inside a tile 16-byte chunks ordered [column / 8][row][8 columns] ("tile-blocked"; the tcgen05 no-swizzle
// interleave).  A tile -- or a range of its chunk columns -- is one contiguous block, so a stage of the pipeline is ONE
// cp.async.bulk (TMA engine, full lines, no thread touches the data, completion on an mbarrier), and the very same
// shared-memory image serves both majors: read K-major (rows = M or N, columns = K: LBO = 2048, SBO = 128) by the forward
// products, MN-major (columns = M or N, rows = K: SBO = 2048, LBO = 128) by dlogits^T h, dpre^T x and dlogits Wap -- no
// transpose anywhere.  Epilogue stores are coalesced in this layout as well (a warp writes 512 contiguous bytes).
// (The first version staged row-major matrices with per-thread 16-byte cp.async: 2.4 .. 2.9 TB/s whatever the pipeline
// depth; profiles/r2_ppo_gemm_*.log.)  bg_ppo_gather_block builds the blocked x from K3's row-major rows and the class
// permutation once per rollout; the classes are padded to whole tiles with zero rows.
//
// Kernels are warp-specialised like K4 (mlp.cu): one producer thread (bulk copies into a ring of stage buffers), one MMA
// issuer (both operands from shared memory, accumulators in TMEM, two of them when N <= 256), eight epilogue warps.
#include <cuda_bf16.h>
#include "bg_device.cuh"
#include "bg_features.cuh"
#include "bg_tcgen05.cuh"
#include "bg_internal.h"

namespace bg {
namespace {

constexpr int kRows = 128;               // sample rows per tile = UMMA M (NT) / reduction length per stage (TN)
constexpr int kChunk = kRows * 16;       // bytes of one chunk column (8 matrix columns) of a tile

// flat parameter / gradient layout (policy_net.KEYS order, agent/policy_network.py:44-56)
constexpr int kOffW1 = 0, kOffB1 = 128 * 198, kOffWa = kOffB1 + 128, kOffBa = kOffWa + 500 * 128, kOffWv = kOffBa + 500,
              kOffBv = kOffWv + 128, kNumParams = kOffBv + 1;
static_assert(kNumParams == 90101, "flat parameter layout");

__device__ __forceinline__ void mma_bf16_ss(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
// kind::f16 instruction descriptor: D = f32, A = B = bf16; a_mn / b_mn = 1: the operand is MN-major (bits 15 / 16)
__device__ __forceinline__ uint32_t idesc_bf16(int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ uint32_t tmem_cols_for(int n) { return n <= 32 ? 32u : (n <= 64 ? 64u : (n <= 128 ? 128u : (n <= 256 ? 256u : 512u))); }
// one lane polls an mbarrier for its warp (every polling thread is shared-memory traffic the tensor core's operand fetches
// compete with: with all threads of the CTA spinning an SS MMA took ~400 cycles)
__device__ __forceinline__ void warp_wait(unsigned long long* bar, uint32_t parity, int lane) {
    if (lane == 0) mbar_wait(bar, parity);
    __syncwarp();
}
// the same for the (many) epilogue warps, with a pause between tries: sixteen lanes spinning on try_wait were two thirds of HIDDEN's
// executed instructions and shared their schedulers with the ONE thread that issues the copies and MMAs of a tile
__device__ __forceinline__ void warp_wait_relaxed(unsigned long long* bar, uint32_t parity, int lane) {
    if (lane == 0) {
        while (!mbar_try_wait(bar, parity)) __nanosleep(128);
    }
    __syncwarp();
}
// `bytes` contiguous bytes global -> shared by the TMA engine; completion is counted on `bar` (armed with expect_tx)
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}\n" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// a stage = [src, src + bytes) in pieces of at most 16 KB, all counted on the same barrier
__device__ __forceinline__ void stage_load(uint32_t dst, const unsigned char* src, uint32_t bytes, unsigned long long* bar) {
    for (uint32_t o = 0; o < bytes; o += 16384u) bulk_load(dst + o, src + o, bytes - o < 16384u ? bytes - o : 16384u, bar);
}

// ---------------------------------------------------------------------------------------------------------------
// out[tile] = epilogue( A[tile] . W^T ) for tiles [tile_begin, tile_end); every matrix tile-blocked
struct NtArgs {
    const uint16_t* A; int nc_a;               // source matrix and its chunk columns per tile (K / 8 of the full matrix)
    long long tile_begin, tile_end;
    const uint16_t* W; int w_bytes; int w_rows;   // packed weight tile (chunk layout) and its number of rows
    int N, K, KC;                              // output columns (<= 512), reduction length, K per stage (K % KC == 0, KC % 16 == 0)
    int b_mn;                                  // 0: W rows = N, columns = K (K-major B);  1: W rows = K, columns = N = 128 (MN-major B)
    int D;                                     // ring of D stage buffers (<= 8)
    int epi;                                   // 0 relu, 1 + bias, 2 * [mask > 0]
    const float* bias; const uint16_t* mask;   // mask: h, tile-blocked, 16 chunk columns per tile
    uint16_t* out; int nc_out;
    int dbg;                                   // experiment switches (bg_ppo_gemm_debug): 1 no MMAs, 2 no epilogue stores, 4 no loads, 8 A operand stays in shared memory
    int a_tmem;                                // the A tile is copied shared -> tensor memory (tcgen05.cp) and the MMAs read it from there (N <= 128, KC == K <= 208)
    // epi == 3 (LOGITS_LOSS_A): the class A loss in the epilogue -- `out` receives d loss / d logits, the logits never leave the SM
    const int32_t* counts; const int32_t* actions; const float* old_logp; const float* adv; const float* returns;
    long long n_rows;                          // real class A rows (the tiles' rows beyond it are padding)
    float eps_clip, value_coef, entropy_coef, inv_b;
    float* dbias; float* sums;                 // [512] column sums of dlogits (value head at 500), [3] policy / value / entropy sums
};

constexpr int kNtEpiWarps = 16, kNtParts = kNtEpiWarps / 4;               // warps 0-15 epilogue (four per TMEM lane quarter), 16 producer, 17 MMA issuer
constexpr int kNtThreads = 32 * (kNtEpiWarps + 2);
constexpr int kNtThreadsLoss = 32 * (kNtEpiWarps + 2 + 4);                  // + warps 18-21: converters (all sixteen epilogue warps are busy with the loss)
template <bool LOSS>
__global__ void __launch_bounds__(LOSS ? kNtThreadsLoss : kNtThreads, 1) ppo_gemm_nt_kernel(const NtArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ unsigned long long full[8], empty[8], acc_full[2], acc_empty[2], a_ready[2], w_bar;
    __shared__ uint32_t s_tmem;
    __shared__ __align__(16) float s_bias[512];
    __shared__ __align__(16) float4 s_red[LOSS ? 2 : 1][LOSS ? 4 : 1][LOSS ? kRows : 1];   // per tile parity, part, row: (max, sum exp, sum exp * z, z[action])
    __shared__ float s_vrow[LOSS ? 2 : 1][LOSS ? 2 * kRows : 1];                             // the row's value, per tile parity and group
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t stage_bytes = (uint32_t)(a.KC >> 3) * kChunk;
    if (a.epi == 1 || LOSS) for (int c = tid; c < 512; c += blockDim.x) s_bias[c] = c < a.N ? a.bias[c] : 0.0f;
    unsigned char* A0 = smem + ((a.w_bytes + 1023) & ~1023);
    const uint32_t Ws = smem_u32(smem), As0 = smem_u32(A0);
    const int nacc = a.N <= 256 ? 2 : 1;                               // accumulators in TMEM (columns 0.. and 256..)
    const bool conv = a.a_tmem && !(a.dbg & 8);
    const int conv_warp0 = LOSS ? kNtEpiWarps + 2 : 12;                // the four converter warps (one per TMEM lane quarter)
    const uint32_t a_col = LOSS ? 160u : 128u;                         // the A image's columns inside an accumulator's half of tensor memory (N <= 144)
    if (tid == 0) {
        // conv (HIDDEN): warps 12-15 carry the A tile from shared to tensor memory; they release the stage (128 arrivals) and tell the MMA
        // issuer (a_ready); the epilogue is then warps 0-11
        for (int i = 0; i < 8; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], conv ? 128 : 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], LOSS ? kNtEpiWarps * 16 : (conv ? 12 * 32 : kNtEpiWarps * 32));   // (LOSS: one group of eight warps per accumulator)
            mbar_init(&a_ready[i], 128);
        }
        mbar_init(&w_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" :: "r"(smem_u32(&s_tmem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = s_tmem;
    const long long n_tiles = a.tile_end - a.tile_begin;
    const int n_kc = a.K / a.KC;
    const long long my_tiles = n_tiles > blockIdx.x ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const long long n_steps = my_tiles * n_kc;                         // steps of this CTA: (tile, K chunk)
    auto tile_of = [&](long long t) { return a.tile_begin + blockIdx.x + t * gridDim.x; };

    if (warp == kNtEpiWarps) {
        // ================= producer: one thread, one bulk copy per stage (+ the weight tile, once) =================
        if (lane == 0) {
            mbar_expect_tx(&w_bar, (uint32_t)a.w_bytes);
            stage_load(Ws, reinterpret_cast<const unsigned char*>(a.W), (uint32_t)a.w_bytes, &w_bar);
            for (long long s = 0; s < n_steps; ++s) {
                const int b = (int)(s % a.D);
                const uint32_t it = (uint32_t)(s / a.D);
                mbar_wait(&empty[b], (it & 1u) ^ 1u);                  // the MMAs that read this buffer D steps ago are done
                if (a.dbg & 4) { mbar_arrive(&full[b]); continue; }
                const unsigned char* src = reinterpret_cast<const unsigned char*>(a.A) +
                                           ((size_t)tile_of(s / n_kc) * a.nc_a + (size_t)(s % n_kc) * (a.KC >> 3)) * kChunk;
                mbar_expect_tx(&full[b], stage_bytes);
                stage_load(As0 + (uint32_t)b * stage_bytes, src, stage_bytes, &full[b]);
            }
        }
    } else if (warp == kNtEpiWarps + 1) {
        // ================= MMA issuer =================
        warp_wait(&w_bar, 0u, lane);
        for (long long s = 0; s < n_steps; ++s) {
            const int b = (int)(s % a.D);
            const uint32_t it = (uint32_t)(s / a.D);
            const long long t = s / n_kc;
            const int kc = (int)(s % n_kc), acc = (int)(t % nacc);
            if (conv) warp_wait(&a_ready[acc], (uint32_t)(t / nacc) & 1u, lane);     // the A tile is in tensor memory
            else warp_wait(&full[b], it & 1u, lane);
            if (kc == 0) warp_wait(&acc_empty[acc], ((uint32_t)(t / nacc) & 1u) ^ 1u, lane);
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            if (lane == 0 && !(a.dbg & 1)) {
                const uint32_t Ab = As0 + (uint32_t)b * stage_bytes, D0 = tmem + (uint32_t)(acc * 256);
                if (conv) {
                    // A from tensor memory (TS MMAs: 64 cycles per K16 step against ~415 with both operands in no-swizzle shared memory).
                    // (tcgen05.cp.128x128b as the carrier was measured too: 26 copies of 2 KB cost ~4.5 k cycles per tile and the kernel
                    // stayed MMA-pipe bound at 159 us whatever the loads and stores did; the converter warps below do it in ~1 k.)
                    const uint32_t At = D0 + a_col;
                    for (int ks = 0; ks < a.KC / 16; ++ks) {
                        const uint64_t db = make_smem_desc_kmajor(Ws + (uint32_t)(ks * 2 * a.w_rows * 16), (uint32_t)(a.w_rows * 16), 128);
                        mma_bf16_ts(D0, At + (uint32_t)(ks * 8), db, idesc_bf16(128, a.N, 0, 0), ks > 0 ? 1u : 0u);
                    }
                } else
                for (int ks = 0; ks < a.KC / 16; ++ks) {
                    const int kg = kc * a.KC + ks * 16;                // first reduction index of this MMA
                    const uint64_t da = make_smem_desc_kmajor(Ab + (uint32_t)(ks * 2 * kChunk), kChunk, 128);
                    const uint32_t accum = (kc > 0 || ks > 0) ? 1u : 0u;
                    if (a.b_mn) {
                        // W rows = reduction index (action slots), columns = N = 128 hidden units: MN-major, K groups 128 B apart
                        const uint64_t db = make_smem_desc_kmajor(Ws + (uint32_t)(kg * 16), 128, (uint32_t)(a.w_rows * 16));
                        mma_bf16_ss(D0, da, db, idesc_bf16(128, a.N, 0, 1), accum);
                    } else {
                        for (int n0 = 0; n0 < a.N; n0 += 256) {
                            const int nn = a.N - n0 < 256 ? a.N - n0 : 256;
                            const uint64_t db = make_smem_desc_kmajor(Ws + (uint32_t)((kg >> 3) * a.w_rows * 16 + n0 * 16), (uint32_t)(a.w_rows * 16), 128);
                            mma_bf16_ss(D0 + (uint32_t)n0, da, db, idesc_bf16(128, nn, 0, 0), accum);
                        }
                    }
                }
            }
            if (lane == 0) {
                if (!conv) umma_commit(&empty[b]);                     // the stage buffer may be refilled
                if (kc == n_kc - 1) umma_commit(&acc_full[acc]);       // the tile's accumulator is complete
            }
            __syncwarp();
        }
    } else if (conv && warp >= conv_warp0 && warp < conv_warp0 + 4) {
        // ================= converters (HIDDEN): thread = row; its 26 chunks of the staged A tile -> registers -> tensor memory =================
        const int q = warp & 3, r = q * 32 + lane;
        for (long long t = 0; t < my_tiles; ++t) {
            const int b = (int)(t % a.D), acc = (int)(t % nacc);
            warp_wait(&full[b], (uint32_t)(t / a.D) & 1u, lane);
            if (t >= nacc) warp_wait(&acc_full[acc], (uint32_t)((t - nacc) / nacc) & 1u, lane);   // the MMAs that read these TMEM columns two tiles ago are done
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            const uint4* src = reinterpret_cast<const uint4*>(A0 + (size_t)b * stage_bytes) + r;
            const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 256) + a_col;
#pragma unroll 2
            for (int c = 0; c < a.KC / 8; ++c) tmem_st4(trow + (uint32_t)(4 * c), src[c * (kChunk / 16)]);
            asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
            mbar_arrive(&empty[b]);                                    // the stage buffer may be refilled
            mbar_arrive(&a_ready[acc]);
        }
    } else if (warp < kNtEpiWarps) {
        // ================= epilogue: thread = row (TMEM lane), the four warps of a lane quarter split the column blocks =================
        // (the epilogue is the busiest role: with eight warps HIDDEN and LOGITS_A were bound by it)
        const int q = warp & 3, part = warp >> 2;
        const int r = q * 32 + lane;                                   // row inside the tile
        const int nblk = (a.N + 31) >> 5;                              // blocks of 32 columns (the last may be 16 wide: N = 144)
        if (LOSS) {
            // ---- LOGITS_LOSS_A: logits = acc + bias (rounded to bf16 as the reference's autocast does), then the loss of
            // ppo_agent.py:271-299 and its gradient, as in ppo_loss_grad_packed_kernel (csrc/ppo.cu); d loss / d logits goes out as bf16,
            // the logits do not.  TWO GROUPS of eight warps, each on its own accumulator (tiles t = group, group + 2, ...): the phases of
            // a tile (wait for the MMAs, pass 1, exchange + barrier, pass 2) are serial, and with all sixteen warps on one tile the
            // issue slots idled at every hand-over; two tiles in different phases fill each other's gaps.  Inside a group: lane quarter
            // q (TMEM lanes 32 q ..), column part p of two; the 128 slots of a row are 16 sub-blocks of 8 (one 16-byte chunk column of
            // the blocked output), sub-block sb belongs to part sb & 1 -- so that rows with few legal slots load both parts alike --
            // and is STREAMED from tensor memory twice (pass 1: online max / sums; pass 2: the gradient): holding a row's logits in
            // registers across the exchange spilled.  The rows arrive sorted by their number of legal slots (TensorCoreUpdate.prepare),
            // so most warps stop after the first one or two of their eight sub-blocks and write zeros for the rest.
            const int grp = warp >> 3, lpart = (warp >> 2) & 1;
            float cs[8];                                               // column sums of d loss / d logits (the head biases' gradients): sub-block 2 j + lpart,
#pragma unroll                                                         // column 4 bit4 + 2 bit3 + bit2 of the lane (see the reduce-scatter below)
            for (int j = 0; j < 8; ++j) cs[j] = 0.0f;
            float pl = 0.0f, vl = 0.0f, ent = 0.0f, vsum = 0.0f;
            const float ce = a.entropy_coef * a.inv_b;
            for (long long t = grp; t < my_tiles; t += 2) {
                const int acc = grp, par = (int)((t >> 1) & 1);
                const size_t tile = (size_t)tile_of(t);
                const long long gr = (long long)tile * kRows + r;      // class A sample index
                const bool live = gr < a.n_rows;
                int n = 1, act = 0; float A = 0.0f, olp = 0.0f, ret = 0.0f;
                if (live) { n = __ldg(a.counts + gr); act = __ldg(a.actions + gr); A = __ldg(a.adv + gr); olp = __ldg(a.old_logp + gr); ret = __ldg(a.returns + gr); }
                warp_wait_relaxed(&acc_full[acc], (uint32_t)(t >> 1) & 1u, lane);
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 256);
                auto logits8 = [&](int sb, float (&z)[8]) {            // slots 8 sb .. 8 sb + 7 of this thread's row: bf16(acc + bias), -inf if illegal
                    uint32_t av[8];
                    tmem_ld8(taddr + 8 * sb, av);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int c = 8 * sb + j;
                        const float l = __bfloat162float(__float2bfloat16_rn(__uint_as_float(av[j]) + s_bias[c]));
                        z[j] = c < n ? l : -INFINITY;
                    }
                };
                float m = -INFINITY, ssum = 0.0f, tsum = 0.0f, za = 0.0f;
#pragma unroll 1
                for (int j = 0; j < 8; ++j) {
                    const int sb = 2 * j + lpart;
                    if (!__any_sync(kFull, n > 8 * sb)) break;
                    float z[8];
                    logits8(sb, z);
                    float cm = z[0];
#pragma unroll
                    for (int i = 1; i < 8; ++i) cm = fmaxf(cm, z[i]);
                    if (cm > -INFINITY) {
                        const float mn = fmaxf(m, cm);
                        const float sc = __expf(m - mn);               // 0 on the first block (m = -inf)
                        ssum *= sc; tsum *= sc;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float e = __expf(z[i] - mn);         // 0 for masked slots
                            ssum += e;
                            if (z[i] > -INFINITY) tsum = fmaf(e, z[i], tsum);
                            if (8 * sb + i == act) za = z[i];
                        }
                        m = mn;
                    }
                }
                s_red[par][2 * grp + lpart][r] = make_float4(m, ssum, tsum, za);
                if (lpart == 0) {
                    uint32_t vv[8];
                    tmem_ld8(taddr + 128, vv);
                    tmem_ld_wait();
                    s_vrow[par][grp * kRows + r] = __bfloat162float(__float2bfloat16_rn(__uint_as_float(vv[0]) + s_bias[128]));
                }
                asm volatile("bar.sync %0, %1;\n" :: "r"(1 + grp), "n"(kNtEpiWarps * 16) : "memory");
                const float4 r0 = s_red[par][2 * grp][r], r1 = s_red[par][2 * grp + 1][r];
                const float M = fmaxf(r0.x, r1.x);
                float S = 0.0f, T = 0.0f;
                if (r0.x > -INFINITY) { const float sc = __expf(r0.x - M); S = r0.y * sc; T = r0.z * sc; }
                if (r1.x > -INFINITY) { const float sc = __expf(r1.x - M); S = fmaf(r1.y, sc, S); T = fmaf(r1.z, sc, T); }
                const float ZA = r0.w + r1.w;
                const float lse = M + __logf(S);
                const float H = lse - T / S;                           // entropy = -sum p log p
                const float lpa = ZA - lse;
                const float rr = __expf(lpa - olp);
                const float rc = fminf(fmaxf(rr, 1.0f - a.eps_clip), 1.0f + a.eps_clip);
                const float s1 = rr * A, s2 = rc * A;
                const bool through = (rr >= 1.0f - a.eps_clip && rr <= 1.0f + a.eps_clip) || s1 < s2;
                const float g = through ? -A * rr * a.inv_b : 0.0f;
                const float v = s_vrow[par][grp * kRows + r];
                const float dv = v - ret;
                const float dvalue = 2.0f * a.value_coef * dv * a.inv_b;
                unsigned char* otile = reinterpret_cast<unsigned char*>(a.out) + tile * a.nc_out * kChunk + r * 16;
                const bool b4 = (lane & 16) != 0, b3 = (lane & 8) != 0, b2 = (lane & 4) != 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int sb = 2 * j + lpart;
                    uint32_t o[4] = {0u, 0u, 0u, 0u};
                    if (__any_sync(kFull, n > 8 * sb)) {
                        float z[8], d[8];
                        logits8(sb, z);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            float di = 0.0f;
                            if (live && z[i] > -INFINITY) {
                                const float lp = z[i] - lse;
                                const float pk = __expf(lp);
                                di = -g * pk + (pk > 0.0f ? ce * pk * (lp + H) : 0.0f);
                                if (8 * sb + i == act) di += g;
                            }
                            d[i] = di;
                        }
#pragma unroll
                        for (int e2 = 0; e2 < 4; ++e2) {
                            const __nv_bfloat162 pr = __floats2bfloat162_rn(d[2 * e2], d[2 * e2 + 1]);
                            o[e2] = *reinterpret_cast<const uint32_t*>(&pr);
                        }
                        // column sums over the warp's 32 rows, reduce-scatter: after three halving rounds a lane holds ONE column (4 bit4 +
                        // 2 bit3 + bit2) summed over 8 lanes, two more rounds finish it -- 9 shuffles for 8 columns instead of 40
                        float w4[4], w2[2];
#pragma unroll
                        for (int k = 0; k < 4; ++k) w4[k] = (b4 ? d[4 + k] : d[k]) + __shfl_xor_sync(kFull, b4 ? d[k] : d[4 + k], 16);
#pragma unroll
                        for (int k = 0; k < 2; ++k) w2[k] = (b3 ? w4[2 + k] : w4[k]) + __shfl_xor_sync(kFull, b3 ? w4[k] : w4[2 + k], 8);
                        float w1 = (b2 ? w2[1] : w2[0]) + __shfl_xor_sync(kFull, b2 ? w2[0] : w2[1], 4);
                        w1 += __shfl_xor_sync(kFull, w1, 2);
                        w1 += __shfl_xor_sync(kFull, w1, 1);
                        cs[j] += w1;
                    }
                    if (!(a.dbg & 2)) *reinterpret_cast<uint4*>(otile + (size_t)sb * kChunk) = make_uint4(o[0], o[1], o[2], o[3]);
                }
                asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
                mbar_arrive(&acc_empty[acc]);                          // the accumulator has been read for the last time
                if (lpart == 0) {                                      // the value head's column (128) and the zero padding (129 .. 143)
                    const float dvw = live ? dvalue : 0.0f;
                    const __nv_bfloat162 pr = __floats2bfloat162_rn(dvw, 0.0f);
                    if (!(a.dbg & 2)) {
                        *reinterpret_cast<uint4*>(otile + (size_t)16 * kChunk) = make_uint4(*reinterpret_cast<const uint32_t*>(&pr), 0u, 0u, 0u);
                        *reinterpret_cast<uint4*>(otile + (size_t)17 * kChunk) = make_uint4(0u, 0u, 0u, 0u);
                    }
                    if (live) { vsum += dvw; pl -= fminf(s1, s2); vl = fmaf(dv, dv, vl); ent += H; }
                }
            }
            // the column sums (one atomic per column and warp: the lanes with lane % 4 == 0 hold the eight columns of a sub-block) and the loss sums
            if ((lane & 3) == 0 && a.dbias) {
                const int ci = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
#pragma unroll
                for (int j = 0; j < 8; ++j) atomicAdd(a.dbias + 8 * (2 * j + lpart) + ci, cs[j]);
            }
            if (lpart == 0) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    pl += __shfl_xor_sync(kFull, pl, o); vl += __shfl_xor_sync(kFull, vl, o); ent += __shfl_xor_sync(kFull, ent, o);
                    vsum += __shfl_xor_sync(kFull, vsum, o);
                }
                if (lane == 0) {
                    atomicAdd(a.sums + 0, pl); atomicAdd(a.sums + 1, vl); atomicAdd(a.sums + 2, ent);
                    if (a.dbias) atomicAdd(a.dbias + BG_ACTIONS, vsum);
                }
            }
        } else
        for (long long t = 0; t < my_tiles; ++t) {
            const int acc = (int)(t % nacc);
            const size_t tile = (size_t)tile_of(t);
            unsigned char* otile = reinterpret_cast<unsigned char*>(a.out) + tile * a.nc_out * kChunk + r * 16;
            // relu'(h) mask of this thread's blocks: fetched BEFORE waiting for the accumulator, so that the latency of the loads
            // hides behind the MMAs (loaded after the wait, DPRE_A ran at 2.6 TB/s)
            // (the dpre ops have N = 128: exactly one block per warp)
            uint4 mk[4];
            if (a.epi == 2) {
                const unsigned char* mtile = reinterpret_cast<const unsigned char*>(a.mask) + tile * 16 * kChunk + r * 16;
#pragma unroll
                for (int j = 0; j < 4; ++j) mk[j] = __ldg(reinterpret_cast<const uint4*>(mtile + (size_t)(4 * part + j) * kChunk));
            }
            warp_wait_relaxed(&acc_full[acc], (uint32_t)(t / nacc) & 1u, lane);
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            for (int blk = part; blk < nblk; blk += (conv ? 3 : kNtParts)) {
                const int c0 = 32 * blk, w = a.N - c0 < 32 ? a.N - c0 : 32;
                uint32_t av[32];
                const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 256 + c0);
                if (w == 32) tmem_ld32(taddr, av);
                else { tmem_ld8(taddr, av); tmem_ld8(taddr + 8, av + 8); }         // w == 16
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 4; ++j) {                          // chunk column c0 / 8 + j: 8 results = 16 bytes, a warp writes 512 contiguous bytes
                    if (8 * j >= w) break;
                    const uint32_t mw[4] = {mk[j].x, mk[j].y, mk[j].z, mk[j].w};
                    uint32_t o[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float v0 = __uint_as_float(av[8 * j + 2 * e]), v1 = __uint_as_float(av[8 * j + 2 * e + 1]);
                        if (a.epi == 0) { v0 = fmaxf(v0, 0.0f); v1 = fmaxf(v1, 0.0f); }
                        else if (a.epi == 1) { const float2 bb = *reinterpret_cast<const float2*>(&s_bias[c0 + 8 * j + 2 * e]); v0 += bb.x; v1 += bb.y; }
                        else {                                                       // relu'(h): h is bf16 >= 0, so "> 0" is "!= 0"
                            if ((mw[e] & 0x0000FFFFu) == 0u) v0 = 0.0f;
                            if ((mw[e] & 0xFFFF0000u) == 0u) v1 = 0.0f;
                        }
                        const __nv_bfloat162 pr = __floats2bfloat162_rn(v0, v1);
                        o[e] = *reinterpret_cast<const uint32_t*>(&pr);
                    }
                    if (!(a.dbg & 2)) *reinterpret_cast<uint4*>(otile + (size_t)((c0 >> 3) + j) * kChunk) = make_uint4(o[0], o[1], o[2], o[3]);
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
            mbar_arrive(&acc_empty[acc]);                              // the accumulator may be overwritten
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tmem), "r"(512u) : "memory");
}

// ---------------------------------------------------------------------------------------------------------------
// G[128 x N] += A[tile]^T . B[tile] over tiles [tile_begin, tile_end) (split over the CTAs; f32 atomics into the flat gradient)
struct TnArgs {
    const uint16_t* A;                         // rows x 128 (hidden units: h or dpre), tile-blocked, 16 chunk columns per tile
    const uint16_t* B; int nc_b; int cb0; int N;   // rows x (8 nc_b) tile-blocked; chunk columns [cb0, cb0 + N / 8) are used; N % 16 == 0, <= 256
    long long tile_begin, tile_end;
    int mode;                                  // 0 dWap class A (N = 144: 128 slots, value head at column 128), 1 dWap class B (slot = col_base + n), 2 dW1p (N = 208)
    int col_base;
    float* grad;                               // flat f32 gradient (kNumParams)
    float* scratch;                            // mode 2: [199][128] f32, dW1p transposed (zeroed by the caller)
    int D;                                     // stage ring size
    int dbg;
};

constexpr int kTnThreads = 32 * 6;             // warps 0-3: epilogue at the end (lane quarters), warp 4: producer, warp 5: MMA issuer
__global__ void __launch_bounds__(kTnThreads, 1) ppo_gemm_tn_kernel(const TnArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ unsigned long long full[8], empty[8], done_bar;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t a_bytes = 16 * kChunk, b_bytes = (uint32_t)(a.N >> 3) * kChunk, stage_bytes = a_bytes + b_bytes;
    const uint32_t S0 = smem_u32(smem);
    if (tid == 0) {
        for (int i = 0; i < 8; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(&done_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    const uint32_t ncols = tmem_cols_for(a.N);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" :: "r"(smem_u32(&s_tmem)), "r"(ncols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = s_tmem;
    const long long n_tiles = a.tile_end - a.tile_begin;
    const long long n_steps = n_tiles > blockIdx.x ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int D = a.D;
    if (warp == 4) {
        // ================= producer: two bulk copies per stage (the A tile, the used chunk columns of the B tile) =================
        if (lane == 0) {
            for (long long s = 0; s < n_steps; ++s) {
                const int b = (int)(s % D);
                const uint32_t it = (uint32_t)(s / D);
                mbar_wait(&empty[b], (it & 1u) ^ 1u);
                if (a.dbg & 4) { mbar_arrive(&full[b]); continue; }
                const size_t tile = (size_t)(a.tile_begin + blockIdx.x + s * gridDim.x);
                mbar_expect_tx(&full[b], stage_bytes);
                stage_load(S0 + (uint32_t)b * stage_bytes, reinterpret_cast<const unsigned char*>(a.A) + tile * 16 * kChunk, a_bytes, &full[b]);
                stage_load(S0 + (uint32_t)b * stage_bytes + a_bytes,
                           reinterpret_cast<const unsigned char*>(a.B) + (tile * a.nc_b + a.cb0) * kChunk, b_bytes, &full[b]);
            }
        }
    } else if (warp == 5) {
        // ================= MMA issuer =================
        const uint32_t idesc = idesc_bf16(128, a.N, 1, 1);
        for (long long s = 0; s < n_steps; ++s) {
            const int b = (int)(s % D);
            const uint32_t it = (uint32_t)(s / D);
            warp_wait(&full[b], it & 1u, lane);
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            if (lane == 0) {
                if (!(a.dbg & 1)) {
                    const uint32_t Ab = S0 + (uint32_t)b * stage_bytes, Bb = Ab + a_bytes;
#pragma unroll 1
                    for (int ks = 0; ks < kRows / 16; ++ks) {
                        // both tiles are read MN-major: the reduction index is the ROW (sample); 16 rows = two 8-row groups 128 B apart
                        const uint64_t da = make_smem_desc_kmajor(Ab + (uint32_t)(ks * 256), 128, kChunk);
                        const uint64_t db = make_smem_desc_kmajor(Bb + (uint32_t)(ks * 256), 128, kChunk);
                        mma_bf16_ss(tmem, da, db, idesc, (s > 0 || ks > 0) ? 1u : 0u);
                    }
                }
                umma_commit(&empty[b]);
                if (s == n_steps - 1) umma_commit(&done_bar);
            }
            __syncwarp();
        }
    } else if (n_steps > 0) {
        // ---- epilogue (warps 0-3): accumulator lane = hidden unit m, column = n; warp w reads the lanes of its quarter
        warp_wait(&done_bar, 0u, lane);
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        const int q = warp & 3;
        const int m = q * 32 + lane;
        const int nblk = a.N >> 3;
        for (int blk = 0; blk < nblk; ++blk) {
            uint32_t acc[8];
            tmem_ld8(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(8 * blk), acc);
            tmem_ld_wait();
            if (a.dbg & 1) continue;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int n = 8 * blk + j;
                const float v = __uint_as_float(acc[j]);
                float* dst = nullptr;
                if (a.mode == 0) { if (n < 128) dst = a.grad + kOffWa + n * 128 + m; else if (n == 128) dst = a.grad + kOffWv + m; }
                else if (a.mode == 1) { const int slot = a.col_base + n; if (slot < 500) dst = a.grad + kOffWa + slot * 128 + m; else if (slot == 500) dst = a.grad + kOffWv + m; }
                else { if (n <= 198) dst = a.scratch + n * 128 + m; }       // dW1p^T: lanes = consecutive addresses (finished by grad_w1_finish_kernel)
                if (dst) atomicAdd(dst, v);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tmem), "r"(ncols) : "memory");
}

// x_blocked[row p] = x_rowmajor[perm[p]] (zeros where perm[p] < 0): K3's row-major bf16 rows (ld_src columns, 26 chunks used)
// -> the tile-blocked layout, in the class order the update works in.  One thread per 16-byte chunk; a warp reads 512
// contiguous bytes of one... (chunks of a row are contiguous in the source, rows are contiguous in the destination)
__global__ void __launch_bounds__(256) gather_block_kernel(const uint16_t* __restrict__ src, long long ld_src, const int32_t* __restrict__ perm,
                                                           long long rows_pad, int nch, int set_one_col, uint16_t* __restrict__ dst) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;        // i = (row p, chunk c8), c8 fastest
    if (i >= rows_pad * nch) return;
    const long long p = i / nch;
    const int c8 = (int)(i - p * nch);
    const int g = perm[p];
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (g >= 0) {
        v = __ldg(reinterpret_cast<const uint4*>(src + (long long)g * ld_src + 8 * c8));
        if (set_one_col >= 0 && (set_one_col >> 3) == c8) {            // the bias column: 1.0 in bf16
            uint32_t* w = reinterpret_cast<uint32_t*>(&v);
            const int e = set_one_col & 7;
            w[e >> 1] = (e & 1) ? ((w[e >> 1] & 0x0000FFFFu) | 0x3F800000u) : ((w[e >> 1] & 0xFFFF0000u) | 0x00003F80u);
        }
    }
    *reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(dst) + ((size_t)(p >> 7) * nch + c8) * kChunk + (size_t)(p & 127) * 16) = v;
}

// chunk k (features 8k .. 8k+7, reference order) of the bf16 feature row of a board held as 13 words in registers: feature_chunk_lut
// (bg_features.cuh) with the byte loads spelled as shifts of compile-time-indexed words, so that the board never goes to local memory
__device__ __forceinline__ uint4 feature_chunk_words(const uint32_t (&w)[kBoardWords], int flag, int k, const uint2* lut) {
    auto cnt = [&](int byte) -> uint32_t { return (w[byte >> 2] >> (8 * (byte & 3))) & 15u; };
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    if (k < 12) {                                    // PLAYER1 points 2k, 2k+1
        const uint2 a = lut[cnt(2 * k)], c = lut[cnt(2 * k + 1)];
        o = make_uint4(a.x, a.y, c.x, c.y);
    } else if (k == 12) {                            // bar1/2, off1/15, P2 point 0, first half of P2 point 1
        const uint2 a = lut[cnt(24)], c = lut[cnt(25)];
        o = make_uint4(bar_off_pair_bf16((int)cnt(48), (int)cnt(50)), a.x, a.y, c.x);
    } else if (k < 24) {                             // second half of P2 point q, P2 point q+1, first half of q+2
        const int q = 2 * (k - 12) - 1;
        const uint2 a = lut[cnt(24 + q)], c = lut[cnt(25 + q)], e = lut[cnt(26 + q)];
        o = make_uint4(a.y, c.x, c.y, e.x);
    } else if (k == 24) {                            // second half of P2 point 23, bar2/2, off2/15, flags, pad
        o.x = lut[cnt(47)].y;
        o.y = bar_off_pair_bf16((int)cnt(49), (int)cnt(51));
        o.z = flag == 0 ? 0x00003F80u : 0x3F800000u;
    }
    return o;
}

// K3 + the gather in one pass: the bf16 feature rows of boards52[perm[p]] (reference feature order, turn flag = flags[perm[p]]), written
// straight into the tile-blocked layout with the bias column set -- no row-major feature tensor in between (it was written by K3, read
// by the gather and never used again: 2 x 436 B per sample of traffic and 1.7 GB per 4 M samples).  One thread per row: its 26 chunks
// go to 26 chunk columns, so a warp writes 512 contiguous bytes per chunk.
__global__ void __launch_bounds__(128) encode_block_kernel(const int8_t* __restrict__ boards, const int8_t* __restrict__ flags,
                                                           const int32_t* __restrict__ perm, long long rows_pad, int set_one_col,
                                                           uint16_t* __restrict__ dst) {
    __shared__ uint2 lut[16];
    load_units_lut(lut);
    __syncthreads();
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= rows_pad) return;
    const int g = perm[p];
    uint32_t w[kBoardWords];
#pragma unroll
    for (int i = 0; i < kBoardWords; ++i) w[i] = 0u;
    int flag = 0;
    if (g >= 0) {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(boards + (long long)g * kBoardBytes);
#pragma unroll
        for (int i = 0; i < kBoardWords; ++i) w[i] = __ldg(src + i);
        flag = flags[g] & 1;
    }
    unsigned char* out = reinterpret_cast<unsigned char*>(dst) + (size_t)(p >> 7) * 26 * kChunk + (size_t)(p & 127) * 16;
#pragma unroll
    for (int k = 0; k < 26; ++k) {
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (g >= 0) {
            v = feature_chunk_words(w, flag, k, lut);
            if (set_one_col >= 0 && (set_one_col >> 3) == k) {             // the bias column: 1.0 in bf16 (no indexing into v: it stays in registers)
                const int e = set_one_col & 7;
                const uint32_t keep = (e & 1) ? 0x0000FFFFu : 0xFFFF0000u, one = (e & 1) ? 0x3F800000u : 0x00003F80u;
                if ((e >> 1) == 0) v.x = (v.x & keep) | one;
                else if ((e >> 1) == 1) v.y = (v.y & keep) | one;
                else if ((e >> 1) == 2) v.z = (v.z & keep) | one;
                else v.w = (v.w & keep) | one;
            }
        }
        *reinterpret_cast<uint4*>(out + (size_t)k * kChunk) = v;
    }
}

// dW1p^T [199][128] (scratch of GRAD_W1: with lane = hidden unit the atomics of the accumulator tile are only coalesced in this
// orientation; fc1.weight is [hidden][feature], where they hit a different sector each: 115 us per launch) -> fc1.weight, fc1.bias
__global__ void grad_w1_finish_kernel(const float* __restrict__ scratch, float* __restrict__ grad) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;               // i = m * 199 + n
    if (i >= 128 * 199) return;
    const int m = i / 199, n = i - m * 199;
    const float v = scratch[n * 128 + m];
    if (n < 198) grad[kOffW1 + m * 198 + n] = v; else grad[kOffB1 + m] = v;
}

// ---------------------------------------------------------------------------------------------------------------
// bf16 operand tiles (chunk layout [k / 8][row][8]) and f32 bias rows from the flat f32 master weights
__global__ void ppo_pack_kernel(const float* __restrict__ p, uint16_t* __restrict__ w1p, uint16_t* __restrict__ wap_a,
                                uint16_t* __restrict__ wap_b, float* __restrict__ bias_a, float* __restrict__ bias_b) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    auto bf = [](float v) { return __bfloat16_as_ushort(__float2bfloat16_rn(v)); };
    if (i < 128 * 208) {                                              // W1p: rows = hidden units, K = 208 (198 features, bias, zeros)
        const int n = i / 208, k = i - n * 208;
        const float v = k < 198 ? p[kOffW1 + n * 198 + k] : (k == 198 ? p[kOffB1 + n] : 0.0f);
        w1p[(k >> 3) * (128 * 8) + n * 8 + (k & 7)] = bf(v);
    }
    if (i < 144 * 128) {                                              // class A head: 128 slots, the value head as row 128, zeros
        const int s = i >> 7, k = i & 127;
        const float v = s < 128 ? p[kOffWa + s * 128 + k] : (s == 128 ? p[kOffWv + k] : 0.0f);
        wap_a[(k >> 3) * (144 * 8) + s * 8 + (k & 7)] = bf(v);
    }
    if (i < 512 * 128) {                                              // class B head: 500 slots, the value head as row 500, zeros
        const int s = i >> 7, k = i & 127;
        const float v = s < 500 ? p[kOffWa + s * 128 + k] : (s == 500 ? p[kOffWv + k] : 0.0f);
        wap_b[(k >> 3) * (512 * 8) + s * 8 + (k & 7)] = bf(v);
    }
    if (i < 144) bias_a[i] = i < 128 ? p[kOffBa + i] : (i == 128 ? p[kOffBv] : 0.0f);
    if (i < 512) bias_b[i] = i < 500 ? p[kOffBa + i] : (i == 500 ? p[kOffBv] : 0.0f);
}

// torch.optim.Adam (no weight decay, no amsgrad): one thread per parameter (ppo_agent.py:83,301-305)
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int n,
                            float lr, float b1, float b2, float eps, float bc1, float bc2_sqrt, float gscale) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float gi = g[i] * gscale;
    const float mi = b1 * m[i] + (1.0f - b1) * gi;
    const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    p[i] -= (lr / bc1) * mi / (sqrtf(vi) / bc2_sqrt + eps);
}

}  // namespace
}  // namespace bg

using namespace bg;

static int g_ppo_gemm_dbg = 0;
extern "C" int bg_ppo_gemm_debug(int flags) { g_ppo_gemm_dbg = flags; return BG_OK; }

extern "C" int bg_ppo_pack_weights(const float* flat_params, uint16_t* w1p, uint16_t* wap_a, uint16_t* wap_b, float* bias_a,
                                   float* bias_b, void* stream) {
    if (!flat_params || !w1p || !wap_a || !wap_b || !bias_a || !bias_b) return bg_set_error_msg(BG_ERR_INVALID, "bg_ppo_pack_weights: null pointer");
    ppo_pack_kernel<<<(512 * 128 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(flat_params, w1p, wap_a, wap_b, bias_a, bias_b);
    return bg_set_error(cudaGetLastError(), "bg_ppo_pack_weights: launch");
}

extern "C" int bg_ppo_gather_block(const uint16_t* x_rowmajor, long long ld_src, const int32_t* perm, long long rows_pad, int ncols,
                                   int set_one_col, uint16_t* x_blocked, void* stream) {
    if (rows_pad < 0 || (rows_pad & 127) || ncols <= 0 || (ncols & 7) || ld_src < ncols)
        return bg_set_error_msg(BG_ERR_INVALID, "bg_ppo_gather_block: rows_pad must be a multiple of 128, ncols of 8, ld_src >= ncols");
    if (rows_pad == 0) return BG_OK;
    if (!x_rowmajor || !perm || !x_blocked) return bg_set_error_msg(BG_ERR_INVALID, "bg_ppo_gather_block: null pointer");
    const long long n = rows_pad * (ncols >> 3);
    gather_block_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x_rowmajor, ld_src, perm, rows_pad, ncols >> 3, set_one_col, x_blocked);
    return bg_set_error(cudaGetLastError(), "bg_ppo_gather_block: launch");
}

extern "C" int bg_ppo_encode_block(const int8_t* boards52, const int8_t* flags, const int32_t* perm, long long rows_pad, int set_one_col,
                                   uint16_t* x_blocked, void* stream) {
    if (rows_pad < 0 || (rows_pad & 127) || set_one_col >= 208) return bg_set_error_msg(BG_ERR_INVALID, "bg_ppo_encode_block: rows_pad must be a multiple of 128");
    if (rows_pad == 0) return BG_OK;
    if (!boards52 || !flags || !perm || !x_blocked) return bg_set_error_msg(BG_ERR_INVALID, "bg_ppo_encode_block: null pointer");
    encode_block_kernel<<<(unsigned)((rows_pad + 127) / 128), 128, 0, (cudaStream_t)stream>>>(boards52, flags, perm, rows_pad, set_one_col, x_blocked);
    return bg_set_error(cudaGetLastError(), "bg_ppo_encode_block: launch");
}

extern "C" int bg_ppo_gemm_nt(int op, const uint16_t* A, long long tile_begin, long long tile_end, const uint16_t* W,
                              const float* bias, const uint16_t* h_mask, uint16_t* out, void* stream) {
    if (tile_begin < 0 || tile_end < tile_begin) return bg_set_error_msg(BG_ERR_INVALID, "bg_ppo_gemm_nt: bad tile range");
    if (tile_end == tile_begin) return BG_OK;
    if (!A || !W || !out) return bg_set_error_msg(BG_ERR_INVALID, "bg_ppo_gemm_nt: null pointer");
    NtArgs a{};
    a.A = A; a.tile_begin = tile_begin; a.tile_end = tile_end; a.W = W; a.bias = bias; a.mask = h_mask; a.out = out; a.dbg = g_ppo_gemm_dbg;
    switch (op) {
        // ring sizes: what fits beside the weight tile in 220 KB of shared memory
        case BG_PPO_OP_HIDDEN:   a.nc_a = 26; a.N = 128; a.K = 208; a.KC = 208; a.w_rows = 128; a.b_mn = 0; a.epi = 0; a.nc_out = 16; a.D = 3; a.a_tmem = 1; break;
        case BG_PPO_OP_LOGITS_A: a.nc_a = 16; a.N = 144; a.K = 128; a.KC = 128; a.w_rows = 144; a.b_mn = 0; a.epi = 1; a.nc_out = 18; a.D = 5; break;
        case BG_PPO_OP_LOGITS_B: a.nc_a = 16; a.N = 512; a.K = 128; a.KC = 128; a.w_rows = 512; a.b_mn = 0; a.epi = 1; a.nc_out = 64; a.D = 2; break;
        case BG_PPO_OP_DPRE_A:   a.nc_a = 18; a.N = 128; a.K = 144; a.KC = 144; a.w_rows = 144; a.b_mn = 1; a.epi = 2; a.nc_out = 16; a.D = 4; break;
        case BG_PPO_OP_DPRE_B:   a.nc_a = 64; a.N = 128; a.K = 512; a.KC = 128; a.w_rows = 512; a.b_mn = 1; a.epi = 2; a.nc_out = 16; a.D = 2; break;
        default: return bg_set_error_msg(BG_ERR_INVALID, "bg_ppo_gemm_nt: unknown op");
    }
    if (a.epi == 1 && !bias) return bg_set_error_msg(BG_ERR_INVALID, "bg_ppo_gemm_nt: the logits ops need the bias row");
    if (a.epi == 2 && !h_mask) return bg_set_error_msg(BG_ERR_INVALID, "bg_ppo_gemm_nt: the dpre ops need h");
    a.w_bytes = (op == BG_PPO_OP_HIDDEN ? 26 : 16) * a.w_rows * 16;
    const size_t smem = ((size_t)(a.w_bytes + 1023) & ~(size_t)1023) + (size_t)a.D * (size_t)(a.KC >> 3) * kChunk;
    cudaError_t e = cudaFuncSetAttribute(ppo_gemm_nt_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (e != cudaSuccess) return bg_set_error(e, "bg_ppo_gemm_nt: cudaFuncSetAttribute");
    if (smem > 220 * 1024) return bg_set_error_msg(BG_ERR_INVALID, "bg_ppo_gemm_nt: stage ring does not fit shared memory");
    const long long tiles = tile_end - tile_begin;
    long long grid = (long long)bg_sm_count();
    if (grid > tiles) grid = tiles;
    ppo_gemm_nt_kernel<false><<<(unsigned)grid, kNtThreads, smem, (cudaStream_t)stream>>>(a);
    return bg_set_error(cudaGetLastError(), "bg_ppo_gemm_nt: launch");
}

// LOGITS_A with the class A loss as its epilogue: dlogits_a = d loss / d logits of rows [0, n_a) (tiles [0, ceil(n_a / 128)); the padding
// rows of the last tile are written as zeros), dbias / sums accumulated as in bg_ppo_loss_grad_classes; means over B_norm samples.
extern "C" int bg_ppo_logits_loss_a(const uint16_t* h, long long n_a, long long B_norm, const uint16_t* wap_a, const float* bias_a,
                                    const int32_t* counts, const int32_t* actions, const float* old_log_probs, const float* advantages,
                                    const float* returns, float eps_clip, float value_coef, float entropy_coef, uint16_t* dlogits_a,
                                    float* dbias, float* sums, void* stream) {
    if (n_a < 0 || B_norm < n_a) return bg_set_error_msg(BG_ERR_INVALID, "bg_ppo_logits_loss_a: bad sizes");
    if (n_a == 0) return BG_OK;
    if (!h || !wap_a || !bias_a || !counts || !actions || !old_log_probs || !advantages || !returns || !dlogits_a || !sums)
        return bg_set_error_msg(BG_ERR_INVALID, "bg_ppo_logits_loss_a: null pointer");
    NtArgs a{};
    a.A = h; a.tile_begin = 0; a.tile_end = (n_a + kRows - 1) / kRows; a.W = wap_a; a.bias = bias_a; a.out = dlogits_a; a.dbg = g_ppo_gemm_dbg;
    a.nc_a = 16; a.N = 144; a.K = 128; a.KC = 128; a.w_rows = 144; a.b_mn = 0; a.epi = 3; a.nc_out = 18; a.D = 5; a.a_tmem = 1;
    a.counts = counts; a.actions = actions; a.old_logp = old_log_probs; a.adv = advantages; a.returns = returns; a.n_rows = n_a;
    a.eps_clip = eps_clip; a.value_coef = value_coef; a.entropy_coef = entropy_coef; a.inv_b = 1.0f / (float)B_norm;
    a.dbias = dbias; a.sums = sums;
    a.w_bytes = 16 * a.w_rows * 16;
    const size_t smem = ((size_t)(a.w_bytes + 1023) & ~(size_t)1023) + (size_t)a.D * (size_t)(a.KC >> 3) * kChunk;
    cudaError_t e = cudaFuncSetAttribute(ppo_gemm_nt_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   // + 20 KB static
    if (e != cudaSuccess) return bg_set_error(e, "bg_ppo_logits_loss_a: cudaFuncSetAttribute");
    long long grid = (long long)bg_sm_count();
    if (grid > a.tile_end) grid = a.tile_end;
    ppo_gemm_nt_kernel<true><<<(unsigned)grid, kNtThreadsLoss, smem, (cudaStream_t)stream>>>(a);
    return bg_set_error(cudaGetLastError(), "bg_ppo_logits_loss_a: launch");
}

extern "C" int bg_ppo_gemm_tn(int op, const uint16_t* A, const uint16_t* B, long long tile_begin, long long tile_end,
                              float* flat_grad, float* scratch, void* stream) {
    if (tile_begin < 0 || tile_end < tile_begin) return bg_set_error_msg(BG_ERR_INVALID, "bg_ppo_gemm_tn: bad tile range");
    if (tile_end == tile_begin) return BG_OK;
    if (!A || !B || !flat_grad) return bg_set_error_msg(BG_ERR_INVALID, "bg_ppo_gemm_tn: null pointer");
    if (op == BG_PPO_OP_GRAD_W1 && !scratch) return bg_set_error_msg(BG_ERR_INVALID, "bg_ppo_gemm_tn: GRAD_W1 needs the scratch");
    cudaError_t e = cudaFuncSetAttribute(ppo_gemm_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (e != cudaSuccess) return bg_set_error(e, "bg_ppo_gemm_tn: cudaFuncSetAttribute");
    const long long tiles = tile_end - tile_begin;
    auto launch = [&](TnArgs a) -> int {
        const size_t stage = ((size_t)16 + (size_t)(a.N >> 3)) * kChunk;
        a.D = (int)((215 * 1024) / stage);
        if (a.D > 4) a.D = 4;
        const size_t smem = (size_t)a.D * stage;
        long long grid = (long long)bg_sm_count();
        if (grid > tiles) grid = tiles;
        ppo_gemm_tn_kernel<<<(unsigned)grid, kTnThreads, smem, (cudaStream_t)stream>>>(a);
        return bg_set_error(cudaGetLastError(), "bg_ppo_gemm_tn: launch");
    };
    TnArgs a{};
    a.A = A; a.B = B; a.tile_begin = tile_begin; a.tile_end = tile_end; a.grad = flat_grad; a.scratch = scratch; a.dbg = g_ppo_gemm_dbg;
    switch (op) {
        case BG_PPO_OP_GRAD_WA_A: a.nc_b = 18; a.cb0 = 0; a.N = 144; a.mode = 0; return launch(a);
        case BG_PPO_OP_GRAD_W1: {
            a.nc_b = 26; a.cb0 = 0; a.N = 208; a.mode = 2;
            cudaError_t e2 = cudaMemsetAsync(scratch, 0, sizeof(float) * 199 * 128, (cudaStream_t)stream);
            if (e2 != cudaSuccess) return bg_set_error(e2, "bg_ppo_gemm_tn: memset");
            const int rc = launch(a);
            if (rc != BG_OK) return rc;
            grad_w1_finish_kernel<<<(128 * 199 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(scratch, flat_grad);   // (overwrites fc1.*: nobody else adds to them)
            return bg_set_error(cudaGetLastError(), "bg_ppo_gemm_tn: finish launch");
        }
        case BG_PPO_OP_GRAD_WA_B:
            for (int cb = 0; cb < 4; ++cb) {                          // four blocks of 128 action slots (value head = slot 500, in the last)
                a.nc_b = 64; a.cb0 = 16 * cb; a.N = 128; a.mode = 1; a.col_base = 128 * cb;
                const int rc = launch(a);
                if (rc != BG_OK) return rc;
            }
            return BG_OK;
        default: return bg_set_error_msg(BG_ERR_INVALID, "bg_ppo_gemm_tn: unknown op");
    }
}

extern "C" int bg_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n, float lr, float beta1,
                            float beta2, float eps, int step, float grad_scale, void* stream) {
    if (n < 0 || step < 1) return bg_set_error_msg(BG_ERR_INVALID, "bg_adam_step: bad n / step (step counts from 1)");
    if (n == 0) return BG_OK;
    if (!params || !grads || !exp_avg || !exp_avg_sq) return bg_set_error_msg(BG_ERR_INVALID, "bg_adam_step: null pointer");
    const float bc1 = 1.0f - powf(beta1, (float)step), bc2 = 1.0f - powf(beta2, (float)step);
    adam_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(params, grads, exp_avg, exp_avg_sq, (int)n, lr, beta1, beta2,
                                                                             eps, bc1, sqrtf(bc2), grad_scale);
    return bg_set_error(cudaGetLastError(), "bg_adam_step: launch");
}
