// ppo_gemm.cu -- N2: the GEMMs of the PPO update on the tcgen05 tensor cores, hand-written (no cuBLAS on the update path).
//
// Replaces, for one epoch over B samples, the linear algebra of BackgammonPPOAgent.update (src/agent/ppo_agent.py:268-305:
// policy_network forward, loss.backward()) around the loss kernel of ppo.cu:
//     h      = relu(x W1p^T)                    bg_ppo_gemm_nt(HIDDEN)        x (B,208) bf16 = K3's rows, column 198 = 1 (bias)
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
// Operands are staged in shared memory in ONE physical layout, 16-byte chunks [column / 8][row][8 columns] (the tcgen05
// no-swizzle "interleave"), which serves both majors: read as K-major (rows = M or N, columns = K: LBO = chunk stride,
// SBO = 128) for the forward GEMMs, and as MN-major (columns = M or N, rows = K: SBO = chunk stride, LBO = 128) for the
// transposed products -- dlogits^T h, dpre^T x and dlogits Wap all read the very same tiles without a transpose.
// Both operands come from shared memory (tcgen05.mma SS form), accumulators live in TMEM; the kernels are HBM-bound
// (they stream 0.3 .. 1 KB per sample).  Inside a CTA a step is serial -- cp.async of the next stage(s) is issued, the
// MMAs of the current one run, the epilogue reads TMEM -- and the phases of different CTAs overlap: two CTAs per SM.
#include <cuda_bf16.h>
#include "bg_device.cuh"
#include "bg_tcgen05.cuh"
#include "bg_internal.h"

namespace bg {
namespace {

constexpr int kGT = 256;                 // threads per CTA
constexpr int kRows = 128;               // sample rows per tile = UMMA M (NT) / UMMA K per tile (TN)
constexpr int kChunk = kRows * 16;       // bytes of one 16-byte-chunk column of a 128-row tile

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
// one full 32-byte sector per thread and instruction (STG.256, sm_100): p must be 32-byte aligned
__device__ __forceinline__ void st_global_256(void* p, const uint32_t* v) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" :: "l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]),
                 "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void cp_async_wait_but(int newest) {          // wait until at most `newest` (0, 1 or 2) commit groups are pending
    if (newest >= 2) asm volatile("cp.async.wait_group 2;\n" ::: "memory");
    else if (newest == 1) asm volatile("cp.async.wait_group 1;\n" ::: "memory");
    else asm volatile("cp.async.wait_group 0;\n" ::: "memory");
}
// one lane polls an mbarrier for its warp (every polling thread is shared-memory traffic the tensor core's operand fetches compete with)
__device__ __forceinline__ void warp_wait(unsigned long long* bar, uint32_t parity, int lane) {
    if (lane == 0) mbar_wait(bar, parity);
    __syncwarp();
}
__device__ __forceinline__ uint32_t tmem_cols_for(int n) { return n <= 32 ? 32u : (n <= 64 ? 64u : (n <= 128 ? 128u : (n <= 256 ? 256u : 512u))); }

// rows [row0, row0 + ROWS) x columns [col0, col0 + 8 nch) of a row-major bf16 matrix -> the chunk layout at `dst` (shared address).
// 16-byte accesses are coalesced per QUARTER warp (8 lanes): a quarter covers 4 rows x 2 adjacent chunks, i.e. one full
// 32-byte sector per row in global memory (8 rows x 1 chunk fetched every sector twice: measured 2x L2 traffic) and two
// 64-byte runs in shared memory (a 2-way bank conflict at most).  A warp moves 16 rows x 2 chunks per instruction.
// Rows >= row_end are zero-filled.
template <int ROWS>
__device__ __forceinline__ void stage_tile(uint32_t dst, unsigned char* dst_generic, const uint16_t* __restrict__ src, long long ld,
                                           long long row0, long long row_end, int col0, int nch, int warp, int nwarps, int lane) {
    constexpr int RB = ROWS / 16;                                      // 16-row blocks of the tile
    const int r_in = (lane >> 3) * 4 + (lane & 3), c_in = (lane >> 2) & 1;
    const uint16_t* base = src + row0 * ld + col0;
    // unit u = (row block rb, chunk pair cb), rb fastest; a warp walks units warp, warp + nwarps, ... without divisions
    int rb = warp % RB, cb = warp / RB;
    const int drb = nwarps % RB, dcb = nwarps / RB;
    for (; 2 * cb < nch; ) {
        const int r = 16 * rb + r_in, c8 = 2 * cb + c_in;
        if (c8 < nch) {
            const uint32_t off = (uint32_t)(c8 * (ROWS * 16) + r * 16);
            if (row0 + r < row_end) cp_async16_s(dst + off, base + (long long)r * ld + 8 * c8);
            else *reinterpret_cast<uint4*>(dst_generic + off) = make_uint4(0u, 0u, 0u, 0u);
        }
        rb += drb; cb += dcb;
        if (rb >= RB) { rb -= RB; ++cb; }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// C[rows x N] = epilogue( A[rows x K] . W^T ),  M = 128 rows per tile
struct NtArgs {
    const uint16_t* A; long long lda;          // source rows (global, bf16), rows [row_begin, row_end)
    long long row_begin, row_end;
    const uint16_t* W; int w_bytes; int w_rows;   // packed weight tile (chunk layout) and its number of rows
    int N, K, KC;                              // output columns (<= 512), reduction length, K per stage (K % KC == 0, KC % 16 == 0)
    int b_mn;                                  // 0: W rows = N, columns = K (K-major B);  1: W rows = K, columns = N = 128 (MN-major B)
    int D, LA;                                 // ring of D stage buffers (<= 8); LA unused (the producers keep min(D-1, 3) stages in flight)
    int epi;                                   // 0 relu, 1 + bias, 2 * [mask > 0]
    const float* bias; const uint16_t* mask; long long ldm;
    uint16_t* out; long long ldo;
    int dbg;                                   // experiment switches (bg_ppo_gemm_debug): 1 no MMAs, 2 no epilogue stores, 4 no loads
};

// Warp-specialised like K4 (mlp.cu): warps 0-7 epilogue (warp w: TMEM lane quarter w % 4, column half w / 4), warps 8-11
// producers (cp.async of the A tiles into a ring of D stage buffers, up to three stages in flight per thread), warp 12 the
// MMA issuer; two accumulators in TMEM (when N <= 256), so loading tile t+2, multiplying tile t+1 and writing out tile t
// overlap.  (Before this split a step was serial inside the CTA and the kernel ran at 1.8 .. 3 TB/s.)
constexpr int kNtEpiWarps = 8, kNtProdWarps = 4, kNtThreads = 32 * (kNtEpiWarps + kNtProdWarps + 1);
__global__ void __launch_bounds__(kNtThreads, 1) ppo_gemm_nt_kernel(const NtArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ unsigned long long full[8], empty[8], acc_full[2], acc_empty[2];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(16) float s_bias[512];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int stage_bytes = (a.KC >> 3) * kChunk;
    if (a.epi == 1) for (int c = tid; c < 512; c += blockDim.x) s_bias[c] = c < a.N ? a.bias[c] : 0.0f;
    unsigned char* Wg = smem;
    unsigned char* A0 = smem + ((a.w_bytes + 1023) & ~1023);
    const uint32_t Ws = smem_u32(Wg), As0 = smem_u32(A0);
    for (int c = tid; c < a.w_bytes / 16; c += blockDim.x) cp_async16_s(Ws + 16u * c, reinterpret_cast<const unsigned char*>(a.W) + 16 * c);
    cp_async_commit();
    cp_async_wait_but(0);
    const int nacc = a.N <= 256 ? 2 : 1;                               // accumulators in TMEM (columns 0.. and 256..)
    if (tid == 0) {
        for (int i = 0; i < 8; ++i) { mbar_init(&full[i], kNtProdWarps * 32); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], kNtEpiWarps * 32); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" :: "r"(smem_u32(&s_tmem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // the weight tile is visible to the tensor-core proxy
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = s_tmem;
    const long long n_tiles = (a.row_end - a.row_begin + kRows - 1) / kRows;
    const int n_kc = a.K / a.KC, nch = a.KC >> 3;
    const long long my_tiles = n_tiles > blockIdx.x ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const long long n_steps = my_tiles * n_kc;                         // steps of this CTA: (tile, K chunk)
    auto tile_of = [&](long long t) { return blockIdx.x + t * gridDim.x; };

    if (warp >= kNtEpiWarps && warp < kNtEpiWarps + kNtProdWarps) {
        // ================= producers =================
        const int pw = warp - kNtEpiWarps;
        const int ahead = a.D >= 4 ? 2 : (a.D == 3 ? 1 : 0);           // stages this thread keeps in flight beyond the one it completes
        for (long long s = 0; s < n_steps; ++s) {
            const int b = (int)(s % a.D);
            const uint32_t it = (uint32_t)(s / a.D);
            warp_wait(&empty[b], (it & 1u) ^ 1u, lane);                // the MMAs that read this buffer D steps ago are done
            if (!(a.dbg & 4))
                stage_tile<kRows>(As0 + (uint32_t)(b * stage_bytes), A0 + b * stage_bytes, a.A, a.lda, a.row_begin + tile_of(s / n_kc) * kRows,
                                  a.row_end, (int)(s % n_kc) * a.KC, nch, pw, kNtProdWarps, lane);
            cp_async_commit();
            if (s >= ahead) {                                          // stage s - ahead has landed: hand it to the MMA warp
                cp_async_wait_but(ahead);
                asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
                mbar_arrive(&full[(int)((s - ahead) % a.D)]);
            }
        }
        for (long long s = n_steps > ahead ? n_steps - ahead : 0; s < n_steps; ++s) {
            cp_async_wait_but((int)(n_steps - 1 - s));
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
            mbar_arrive(&full[(int)(s % a.D)]);
        }
    } else if (warp == kNtEpiWarps + kNtProdWarps) {
        // ================= MMA issuer =================
        for (long long s = 0; s < n_steps; ++s) {
            const int b = (int)(s % a.D);
            const uint32_t it = (uint32_t)(s / a.D);
            const long long t = s / n_kc;
            const int kc = (int)(s % n_kc), acc = (int)(t % nacc);
            warp_wait(&full[b], it & 1u, lane);
            if (kc == 0) warp_wait(&acc_empty[acc], ((uint32_t)(t / nacc) & 1u) ^ 1u, lane);
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            if (lane == 0 && !(a.dbg & 1)) {
                const uint32_t Ab = As0 + (uint32_t)(b * stage_bytes), D0 = tmem + (uint32_t)(acc * 256);
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
                umma_commit(&empty[b]);                                // the stage buffer may be refilled
                if (kc == n_kc - 1) umma_commit(&acc_full[acc]);       // the tile's accumulator is complete
            }
            __syncwarp();
        }
    } else if (warp < kNtEpiWarps) {
        // ================= epilogue: thread = row (TMEM lane), the two warps of a lane quarter split the columns =================
        const int q = warp & 3, part = warp >> 2;
        for (long long t = 0; t < my_tiles; ++t) {
            const int acc = (int)(t % nacc);
            warp_wait(&acc_full[acc], (uint32_t)(t / nacc) & 1u, lane);
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            const long long row = a.row_begin + tile_of(t) * kRows + q * 32 + lane;
            const bool live = row < a.row_end;
            const int nblk = (a.N + 31) >> 5;                          // blocks of 32 columns (the last may be 16 wide: N = 144)
            for (int blk = part; blk < nblk; blk += 2) {
                const int c0 = 32 * blk, w = a.N - c0 < 32 ? a.N - c0 : 32;
                uint32_t av[32];
                const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 256 + c0);
                if (w == 32) tmem_ld32(taddr, av);
                else { tmem_ld8(taddr, av); tmem_ld8(taddr + 8, av + 8); }         // w == 16
                tmem_ld_wait();
                if (live) {
                    uint32_t mk[16];
                    if (a.epi == 2) {
                        const uint4* mp = reinterpret_cast<const uint4*>(a.mask + row * a.ldm + c0);
#pragma unroll
                        for (int j = 0; j < 4; ++j) { const uint4 v = (8 * j < w) ? __ldg(mp + j) : make_uint4(0u, 0u, 0u, 0u); mk[4 * j] = v.x; mk[4 * j + 1] = v.y; mk[4 * j + 2] = v.z; mk[4 * j + 3] = v.w; }
                    }
                    uint32_t o[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        float v0 = __uint_as_float(av[2 * j]), v1 = __uint_as_float(av[2 * j + 1]);
                        if (a.epi == 0) { v0 = fmaxf(v0, 0.0f); v1 = fmaxf(v1, 0.0f); }
                        else if (a.epi == 1) { const float2 bb = *reinterpret_cast<const float2*>(&s_bias[c0 + 2 * j]); v0 += bb.x; v1 += bb.y; }
                        else {                                                       // relu'(h): h is bf16 >= 0, so "> 0" is "!= 0"
                            if ((mk[j] & 0x0000FFFFu) == 0u) v0 = 0.0f;
                            if ((mk[j] & 0xFFFF0000u) == 0u) v1 = 0.0f;
                        }
                        o[j] = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v0)) | ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v1)) << 16);
                    }
                    uint16_t* op = a.out + row * a.ldo + c0;               // rows are 256 / 288 / 1024 bytes: every block is sector aligned
#pragma unroll
                    for (int j = 0; j < 2; ++j)
                        if (16 * j < w && !(a.dbg & 2)) st_global_256(op + 16 * j, o + 8 * j);
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
// G[128 x N] += A[rows x 128]^T . B[rows x N]   (split over the sample rows across the CTAs; f32 atomics into the flat gradient)
struct TnArgs {
    const uint16_t* A; long long lda;          // rows x 128 (hidden units: h or dpre)
    const uint16_t* B; long long ldb; int N;   // rows x N  (dlogits or x); N % 16 == 0, <= 256
    long long row_begin, row_end;
    int mode;                                  // 0 dWap class A (N = 144: 128 slots, value head at column 128), 1 dWap class B (slot = col_base + n), 2 dW1p (N = 208)
    int col_base;
    float* grad;                               // flat f32 gradient (kNumParams)
    float* scratch;                            // mode 2: [199][128] f32, dW1p transposed (zeroed by the caller)
    int D;                                     // stage ring size
    int dbg;
};

constexpr int kTnRows = 64;               // sample rows per stage of the TN kernel (four MMA K-steps)
constexpr int kTnProdWarps = 4, kTnThreads = 32 * (kTnProdWarps + 1);
// warps 0-3: producers (cp.async ring of D stages, three in flight per thread), then the epilogue; warp 4: MMA issuer
__global__ void __launch_bounds__(kTnThreads, 1) ppo_gemm_tn_kernel(const TnArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ unsigned long long full[8], empty[8], done_bar;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int CS = kTnRows * 16;                                   // chunk-column stride of a stage tile
    const int a_bytes = 16 * CS, b_bytes = (a.N >> 3) * CS, stage_bytes = a_bytes + b_bytes;
    const uint32_t S0 = smem_u32(smem);
    if (tid == 0) {
        for (int i = 0; i < 8; ++i) { mbar_init(&full[i], kTnProdWarps * 32); mbar_init(&empty[i], 1); }
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
    const long long n_tiles = (a.row_end - a.row_begin + kTnRows - 1) / kTnRows;
    const long long n_steps = n_tiles > blockIdx.x ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int D = a.D;
    if (warp < kTnProdWarps) {
        // ================= producers =================
        const int ahead = D >= 4 ? 2 : (D == 3 ? 1 : 0);
        for (long long s = 0; s < n_steps; ++s) {
            const int b = (int)(s % D);
            const uint32_t it = (uint32_t)(s / D);
            warp_wait(&empty[b], (it & 1u) ^ 1u, lane);
            if (!(a.dbg & 4)) {
                const long long row0 = a.row_begin + (blockIdx.x + s * gridDim.x) * kTnRows;
                stage_tile<kTnRows>(S0 + (uint32_t)(b * stage_bytes), smem + b * stage_bytes, a.A, a.lda, row0, a.row_end, 0, 16, warp, kTnProdWarps, lane);
                stage_tile<kTnRows>(S0 + (uint32_t)(b * stage_bytes + a_bytes), smem + b * stage_bytes + a_bytes, a.B, a.ldb, row0, a.row_end, 0, a.N >> 3,
                                    warp, kTnProdWarps, lane);
            }
            cp_async_commit();
            if (s >= ahead) {
                cp_async_wait_but(ahead);
                asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
                mbar_arrive(&full[(int)((s - ahead) % D)]);
            }
        }
        for (long long s = n_steps > ahead ? n_steps - ahead : 0; s < n_steps; ++s) {
            cp_async_wait_but((int)(n_steps - 1 - s));
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
            mbar_arrive(&full[(int)(s % D)]);
        }
    } else {
        // ================= MMA issuer =================
        const uint32_t idesc = idesc_bf16(128, a.N, 1, 1);
        for (long long s = 0; s < n_steps; ++s) {
            const int b = (int)(s % D);
            const uint32_t it = (uint32_t)(s / D);
            warp_wait(&full[b], it & 1u, lane);
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            if (lane == 0) {
                if (!(a.dbg & 1)) {
                    const uint32_t Ab = S0 + (uint32_t)(b * stage_bytes), Bb = Ab + (uint32_t)a_bytes;
#pragma unroll 1
                    for (int ks = 0; ks < kTnRows / 16; ++ks) {
                        // both tiles are read MN-major: the reduction index is the ROW (sample); 16 rows = two 8-row groups 128 B apart
                        const uint64_t da = make_smem_desc_kmajor(Ab + (uint32_t)(ks * 256), 128, CS);
                        const uint64_t db = make_smem_desc_kmajor(Bb + (uint32_t)(ks * 256), 128, CS);
                        mma_bf16_ss(tmem, da, db, idesc, (s > 0 || ks > 0) ? 1u : 0u);
                    }
                }
                umma_commit(&empty[b]);
                if (s == n_steps - 1) umma_commit(&done_bar);
            }
            __syncwarp();
        }
    }
    if (n_steps > 0 && warp < 4) {
        // ---- epilogue: accumulator lane = hidden unit m, column = n; warp w reads the lanes of its quarter
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

extern "C" int bg_ppo_gemm_nt(int op, const uint16_t* A, long long row_begin, long long row_end, const uint16_t* W,
                              const float* bias, const uint16_t* h_mask, uint16_t* out, void* stream) {
    if (row_begin < 0 || row_end < row_begin) return bg_set_error_msg(BG_ERR_INVALID, "bg_ppo_gemm_nt: bad row range");
    if (row_end == row_begin) return BG_OK;
    if (!A || !W || !out) return bg_set_error_msg(BG_ERR_INVALID, "bg_ppo_gemm_nt: null pointer");
    NtArgs a{};
    a.A = A; a.row_begin = row_begin; a.row_end = row_end; a.W = W; a.bias = bias; a.mask = h_mask; a.ldm = 128; a.out = out; a.dbg = g_ppo_gemm_dbg;
    switch (op) {
        // (ring D / stages in flight LA: what fits beside the weight tile in 220 KB)
        // ring sizes: what fits beside the weight tile in 220 KB of shared memory
        case BG_PPO_OP_HIDDEN:   a.lda = 208; a.N = 128; a.K = 208; a.KC = 208; a.w_rows = 128; a.b_mn = 0; a.epi = 0; a.ldo = 128; a.D = 3; break;
        case BG_PPO_OP_LOGITS_A: a.lda = 128; a.N = 144; a.K = 128; a.KC = 128; a.w_rows = 144; a.b_mn = 0; a.epi = 1; a.ldo = 144; a.D = 5; break;
        case BG_PPO_OP_LOGITS_B: a.lda = 128; a.N = 512; a.K = 128; a.KC = 128; a.w_rows = 512; a.b_mn = 0; a.epi = 1; a.ldo = 512; a.D = 2; break;
        case BG_PPO_OP_DPRE_A:   a.lda = 144; a.N = 128; a.K = 144; a.KC = 144; a.w_rows = 144; a.b_mn = 1; a.epi = 2; a.ldo = 128; a.D = 4; break;
        case BG_PPO_OP_DPRE_B:   a.lda = 512; a.N = 128; a.K = 512; a.KC = 128; a.w_rows = 512; a.b_mn = 1; a.epi = 2; a.ldo = 128; a.D = 2; break;
        default: return bg_set_error_msg(BG_ERR_INVALID, "bg_ppo_gemm_nt: unknown op");
    }
    if (a.epi == 1 && !bias) return bg_set_error_msg(BG_ERR_INVALID, "bg_ppo_gemm_nt: the logits ops need the bias row");
    if (a.epi == 2 && !h_mask) return bg_set_error_msg(BG_ERR_INVALID, "bg_ppo_gemm_nt: the dpre ops need h");
    a.w_bytes = (op == BG_PPO_OP_HIDDEN ? 26 : 16) * a.w_rows * 16;
    const size_t smem = ((size_t)(a.w_bytes + 1023) & ~(size_t)1023) + (size_t)a.D * (size_t)(a.KC >> 3) * kChunk;
    cudaError_t e = cudaFuncSetAttribute(ppo_gemm_nt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (e != cudaSuccess) return bg_set_error(e, "bg_ppo_gemm_nt: cudaFuncSetAttribute");
    if (smem > 220 * 1024) return bg_set_error_msg(BG_ERR_INVALID, "bg_ppo_gemm_nt: stage ring does not fit shared memory");
    const long long tiles = (row_end - row_begin + kRows - 1) / kRows;
    long long grid = (long long)bg_sm_count();
    if (grid > tiles) grid = tiles;
    ppo_gemm_nt_kernel<<<(unsigned)grid, kNtThreads, smem, (cudaStream_t)stream>>>(a);
    return bg_set_error(cudaGetLastError(), "bg_ppo_gemm_nt: launch");
}

extern "C" int bg_ppo_gemm_tn(int op, const uint16_t* A, const uint16_t* B, long long row_begin, long long row_end,
                              float* flat_grad, float* scratch, void* stream) {
    if (row_begin < 0 || row_end < row_begin) return bg_set_error_msg(BG_ERR_INVALID, "bg_ppo_gemm_tn: bad row range");
    if (row_end == row_begin) return BG_OK;
    if (!A || !B || !flat_grad) return bg_set_error_msg(BG_ERR_INVALID, "bg_ppo_gemm_tn: null pointer");
    if (op == BG_PPO_OP_GRAD_W1 && !scratch) return bg_set_error_msg(BG_ERR_INVALID, "bg_ppo_gemm_tn: GRAD_W1 needs the scratch");
    cudaError_t e = cudaFuncSetAttribute(ppo_gemm_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (e != cudaSuccess) return bg_set_error(e, "bg_ppo_gemm_tn: cudaFuncSetAttribute");
    const long long tiles = (row_end - row_begin + kTnRows - 1) / kTnRows;
    auto launch = [&](TnArgs a) -> int {
        const size_t stage = ((size_t)16 + (size_t)(a.N >> 3)) * (kTnRows * 16);
        a.D = (int)((215 * 1024) / stage);
        if (a.D > 6) a.D = 6;
        const size_t smem = (size_t)a.D * stage;
        long long grid = (long long)bg_sm_count();
        if (grid > tiles) grid = tiles;
        ppo_gemm_tn_kernel<<<(unsigned)grid, kTnThreads, smem, (cudaStream_t)stream>>>(a);
        return bg_set_error(cudaGetLastError(), "bg_ppo_gemm_tn: launch");
    };
    TnArgs a{};
    a.A = A; a.lda = 128; a.B = B; a.row_begin = row_begin; a.row_end = row_end; a.grad = flat_grad; a.scratch = scratch; a.dbg = g_ppo_gemm_dbg;
    switch (op) {
        case BG_PPO_OP_GRAD_WA_A: a.ldb = 144; a.N = 144; a.mode = 0; return launch(a);
        case BG_PPO_OP_GRAD_W1: {
            a.ldb = 208; a.N = 208; a.mode = 2;
            cudaError_t e2 = cudaMemsetAsync(scratch, 0, sizeof(float) * 199 * 128, (cudaStream_t)stream);
            if (e2 != cudaSuccess) return bg_set_error(e2, "bg_ppo_gemm_tn: memset");
            const int rc = launch(a);
            if (rc != BG_OK) return rc;
            grad_w1_finish_kernel<<<(128 * 199 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(scratch, flat_grad);   // (overwrites fc1.*: nobody else adds to them)
            return bg_set_error(cudaGetLastError(), "bg_ppo_gemm_tn: finish launch");
        }
        case BG_PPO_OP_GRAD_WA_B:
            for (int cb = 0; cb < 4; ++cb) {                          // four blocks of 128 action slots (value head = slot 500, in the last)
                a.ldb = 512; a.N = 128; a.mode = 1; a.col_base = 128 * cb; a.B = B + 128 * cb;
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
