// encode.cu -- K3: the reference's 198-feature board encoding, written as f32 rows (API parity) or
// as bf16 rows padded to a multiple of 8 columns (the MLP's tensor-core input).
//
// Replaces get_board_features_batch_from_tensors (src/ai/batching.py:78-147) and
// ImmutableBoard.get_board_features (src/board/immutable_board.py:171-212):
//   [0,96)   PLAYER1 points 0..23, 4 units each: c==1 -> 1,0,0,0; c==2 -> 1,1,0,0; c>=3 -> 1,1,1,(c-3)/2
//   [96] bar1/2   [97] off1/15   [98,194) PLAYER2 points   [194] bar2/2   [195] off2/15
//   [196] 1 if flag==PLAYER1     [197] 1 if flag==PLAYER2
// HBM-bound: 52 B read, 792 B (f32) or 2*ld B (bf16) written per row.  A CTA stages ROWS boards in
// shared memory with coalesced loads, then every thread produces 16-byte (bf16) / 8-byte (f32) output
// vectors so that a warp's stores cover contiguous 512 / 256 bytes.
#include <cstdlib>
#include <cuda_bf16.h>
#include "bg_device.cuh"
#include "bg_features.cuh"
#include "bg_internal.h"

namespace bg {

constexpr int kEncRows = 128;     // boards per CTA tile

__device__ __forceinline__ float4 point_units_f32(int c) {
    float4 r;
    r.x = c >= 1 ? 1.0f : 0.0f; r.y = c >= 2 ? 1.0f : 0.0f; r.z = c >= 3 ? 1.0f : 0.0f;
    r.w = c >= 3 ? ((float)c - 3.0f) * 0.5f : 0.0f;                       // batching.py:117-119 (/2 is exact)
    return r;
}

// bf16 rows.  Output = 16-byte chunks (8 features); chunk k of a row is described by kChunkDesc[k]: byte w of it
// describes word w (2 features): bit 7 clear -> bits 0..5 = board52 byte of the point, bit 6 = which half of its 4
// units; bit 7 set -> 0 zero, 1 bar1/off1, 2 bar2/off2, 3 turn flags.
// Thread t of a CTA owns chunk column k = t % cpr for rows t / cpr, t / cpr + 16, ... of the tile, so the descriptor
// is decoded ONCE per thread (four byte offsets, four table offsets) and a chunk costs four (count byte -> units
// word) pairs of shared-memory loads and one 16-byte store; consecutive threads write consecutive chunks (rows are
// contiguous), i.e. full 512-byte runs per warp.  A staged row is 16 words: the 13 board words, then the row's
// three "special" words (bar1/off1, bar2/off2, turn flags) precomputed once per row, so that a special word is just
// another address -- no divergent branch, no constant-memory lookups.
constexpr int kBfRows = 128;
constexpr int kBfRowWords = 16;
constexpr int kBfRowsPerPass = 16;
constexpr int kBfThreads = kBfRowsPerPass * 26;       // 416 = 13 warps (ld = 208); other ld: threads beyond 16*cpr idle
constexpr int kBfUnroll = 4;

__global__ void __launch_bounds__(kBfThreads, 3) encode_bf16_kernel(const int8_t* __restrict__ boards,
                                                                 const int8_t* __restrict__ flags, int flag_all,
                                                                 long long B, const unsigned long long* __restrict__ row_begin_dev,
                                                                 const unsigned long long* __restrict__ n_rows_dev,
                                                                 uint16_t* __restrict__ out, int cpr /* ld/8 */) {
    __shared__ __align__(16) uint32_t sm[kBfRows * kBfRowWords];
    __shared__ uint32_t s_lut[32], s_desc[32];
    load_chunk_tables(s_lut, s_desc);
    if (n_rows_dev) B = min(B, (long long)*n_rows_dev);
    const long long begin = row_begin_dev ? (long long)*row_begin_dev : 0;
    // this thread's chunk column and its decoded descriptor
    const int rpp = cpr <= 26 ? kBfRowsPerPass : kBfThreads / cpr;      // rows per pass
    const int tr = threadIdx.x / cpr, tk = threadIdx.x - tr * cpr;
    const bool active = tr < rpp;
    const uint32_t desc = tk < 26 ? kChunkDesc[tk] : 0x80808080u;
    uint32_t boff[4], loff[4];                                          // byte offset in the row; table offset / special
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint32_t d = (desc >> (8 * q)) & 0xFFu;
        boff[q] = d & 63u;
        loff[q] = (d & 0x80u) ? (0x100u | (d & 3u)) : ((d >> 6) & 1u);
    }
    for (long long row0 = begin + (long long)blockIdx.x * kBfRows; row0 < B; row0 += (long long)gridDim.x * kBfRows) {
        const int rows = (int)min((long long)kBfRows, B - row0);
        const uint32_t* src = reinterpret_cast<const uint32_t*>(boards + row0 * kBoardBytes);
        for (int i = threadIdx.x; i < rows * kBoardWords; i += kBfThreads) {
            const int r = i / kBoardWords;
            sm[r * kBfRowWords + (i - r * kBoardWords)] = __ldg(src + i);
        }
        __syncthreads();
        if (threadIdx.x < rows) {                                            // the row's special words
            const int r = threadIdx.x;
            const uint32_t misc = sm[r * kBfRowWords + 12];
            const int flag = flags ? (flags[row0 + r] & 1) : flag_all;
            sm[r * kBfRowWords + 13] = bar_off_pair_bf16((int)(misc & 0xFFu), (int)((misc >> 16) & 0xFFu));
            sm[r * kBfRowWords + 14] = bar_off_pair_bf16((int)((misc >> 8) & 0xFFu), (int)(misc >> 24));
            sm[r * kBfRowWords + 15] = flag == 0 ? 0x00003F80u : 0x3F800000u;
        }
        __syncthreads();
        if (active) {
            uint4* dst = reinterpret_cast<uint4*>(out + row0 * (long long)cpr * 8) + tk;
            for (int r0 = tr; r0 < rows; r0 += rpp * kBfUnroll) {
                uint4 v[kBfUnroll];
#pragma unroll
                for (int u = 0; u < kBfUnroll; ++u) {
                    const int r = min(r0 + u * rpp, rows - 1);
                    const uint32_t* srow = sm + r * kBfRowWords;
                    const uint8_t* b = reinterpret_cast<const uint8_t*>(srow);
                    uint32_t w[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint32_t cnt = b[boff[q]] & 15u;
                        // special: code 0 -> s_lut[0] (= 0), codes 1..3 -> words 13..15 of the staged row
                        const uint32_t* a = (loff[q] & 0x100u) ? ((loff[q] & 3u) ? srow + 12 + (loff[q] & 3u) : s_lut)
                                                               : s_lut + ((cnt << 1) | loff[q]);
                        w[q] = *a;
                    }
                    v[u] = make_uint4(w[0], w[1], w[2], w[3]);
                }
#pragma unroll
                for (int u = 0; u < kBfUnroll; ++u) {
                    const int r = r0 + u * rpp;
                    if (r < rows) dst[(long long)r * cpr] = v[u];
                }
            }
        }
        __syncthreads();
    }
}

// bf16 rows, version 5 (ld = 208 only): whole feature tiles built POINT BY POINT in shared memory and handed to the copy
// engine.  The chunk-per-thread kernel above spends ~38 instructions on a 16-byte chunk (four generic count -> units lookups
// with the special words folded in) -- 76 M warp-instructions per 1.2 M rows, which, not HBM, is what bounds it (a memset
// of the same bytes runs at 7.2 TB/s).  Here a builder thread owns one of the 48 points (its board byte and its two
// destination words are thread constants), a unit of work is one byte load, one 8-byte table load and the store of the
// point's two words into the tile (~6 instructions per 8 bytes); warp 6 writes the rows' special words (bar/off pairs, turn
// flags, zero padding); the finished 13.3 KB tile (32 rows) leaves with ONE cp.async.bulk shared -> global.  Boards of the next
// tile are prefetched with cp.async, tiles are double buffered, six CTAs per SM (measured: 64 rows x 3 CTAs 102.8 us, 32 x 6
// 97.7, 32 x 7 102.8, 16 x 9 113.8 per 1.2 M rows).
constexpr int kV5Rows = 32;
constexpr int kV5Words = 104;                 // 208 bf16 = 104 words per row
constexpr int kV5Builders = 192;              // 4 rows x 48 points per pass
constexpr int kV5Threads = 224;               // + warp 6: special words
struct V5Smem {
    uint32_t tile[2][kV5Rows * kV5Words];     // 2 x 13,312 B
    uint32_t boards[2][kV5Rows * kBoardWords];
    uint2 units[16];
    uint16_t half[16], off15[16];
};

__global__ void __launch_bounds__(kV5Threads, 6) encode_bf16_v5_kernel(const int8_t* __restrict__ boards, const int8_t* __restrict__ flags,
                                                                      int flag_all, long long B,
                                                                      const unsigned long long* __restrict__ row_begin_dev,
                                                                      const unsigned long long* __restrict__ n_rows_dev,
                                                                      uint16_t* __restrict__ out) {
    extern __shared__ __align__(128) unsigned char v5_raw[];
    V5Smem& S = *reinterpret_cast<V5Smem*>(v5_raw);
    const int t = threadIdx.x;
    if (t < 16) { S.units[t] = kUnitsBf16[t]; S.half[t] = kHalfBf16[t]; S.off15[t] = kOff15Bf16[t]; }
    if (n_rows_dev) B = min(B, (long long)*n_rows_dev);
    const long long begin = row_begin_dev ? (long long)*row_begin_dev : 0;
    const long long stride = (long long)gridDim.x * kV5Rows;
    const int tr = t / 48, tp = t - tr * 48;                         // builder: rows tr, tr + 4, ...; point tp (board byte tp)
    const int dw = tp < 24 ? 2 * tp : 2 * tp + 1;                    // first destination word: P1 points at 0.., P2 points at 49..
    const int sr = t - kV5Builders;                                  // warp 6: rows sr and sr + 32
    auto prefetch = [&](long long row0, int b, int (&fl)[2]) {
        if (row0 < B) {
            const int rows = (int)min((long long)kV5Rows, B - row0);
            const uint32_t* src = reinterpret_cast<const uint32_t*>(boards + row0 * kBoardBytes);
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&S.boards[b][0]);
            for (int i = t; i < rows * kBoardWords; i += kV5Threads)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" :: "r"(dst + 4u * (uint32_t)i), "l"(src + i) : "memory");
            if (sr >= 0) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int r = sr + 32 * h;
                    fl[h] = (r < rows && flags) ? (__ldg(flags + row0 + r) & 1) : flag_all;
                }
            }
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    };
    int fl_cur[2] = {0, 0}, fl_next[2] = {0, 0};
    long long row0 = begin + (long long)blockIdx.x * kV5Rows;
    prefetch(row0, 0, fl_cur);
    for (int it = 0; row0 < B; row0 += stride, ++it) {
        const int rows = (int)min((long long)kV5Rows, B - row0);
        const int b = it & 1;
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");              // this thread's share of the tile's boards has landed
        if (t == 0) asm volatile("cp.async.bulk.wait_group.read 1;\n" ::: "memory");   // tile[b]'s previous store has left shared memory
        __syncthreads();                                                      // ... everybody's boards; the previous tile is fully built
        prefetch(row0 + stride, b ^ 1, fl_next);
        uint32_t* tl = S.tile[b];
        const uint8_t* sb = reinterpret_cast<const uint8_t*>(S.boards[b]);
        if (t < kV5Builders) {
#pragma unroll 4
            for (int r = tr; r < rows; r += 4) {
                const uint2 u = S.units[sb[r * kBoardBytes + tp] & 15u];
                tl[r * kV5Words + dw] = u.x;
                tl[r * kV5Words + dw + 1] = u.y;
            }
        } else {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int r = sr + 32 * h;
                if (r < rows) {
                    const uint32_t misc = S.boards[b][r * kBoardWords + 12];    // bar1, bar2, off1, off2
                    uint32_t* w = tl + r * kV5Words;
                    w[48] = (uint32_t)S.half[misc & 15u] | ((uint32_t)S.off15[(misc >> 16) & 15u] << 16);
                    w[97] = (uint32_t)S.half[(misc >> 8) & 15u] | ((uint32_t)S.off15[(misc >> 24) & 15u] << 16);
                    w[98] = fl_cur[h] == 0 ? 0x00003F80u : 0x3F800000u;
                    w[99] = 0u; w[100] = 0u; w[101] = 0u; w[102] = 0u; w[103] = 0u;
                }
            }
            fl_cur[0] = fl_next[0]; fl_cur[1] = fl_next[1];
        }
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");      // generic-proxy writes -> visible to the copy engine
        __syncthreads();
        if (t == 0) {
            const uint32_t s_addr = (uint32_t)__cvta_generic_to_shared(tl);
            uint16_t* g_addr = out + row0 * (long long)(2 * kV5Words);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n"
                         :: "l"(g_addr), "r"(s_addr), "r"((uint32_t)(rows * kV5Words * 4)) : "memory");
            asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
        }
    }
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    if (t == 0) asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");   // shared memory must outlive the copies
}

// f32 rows (the reference's own layout).  Thread t owns the float2 pair column j = t % 99 (features 2j, 2j+1) for rows
// t / 99, +4, ... of the tile: the column is decoded once (board byte + which half of its four units, or one of the three
// special pairs), a pair is then one byte load + one 8-byte table load and an 8-byte store; consecutive threads write
// consecutive pairs of a row (792 contiguous bytes per row).
constexpr int kF32RowsPerPass = 4;
constexpr int kF32Threads = kF32RowsPerPass * 99;     // 396

__global__ void __launch_bounds__(kF32Threads) encode_f32_kernel(const int8_t* __restrict__ boards,
                                                                 const int8_t* __restrict__ flags, int flag_all,
                                                                 long long B, const unsigned long long* __restrict__ n_rows_dev,
                                                                 float* __restrict__ out, long long ld) {
    __shared__ __align__(16) uint32_t sm[kEncRows * kBoardWords];
    __shared__ int8_t sflag[kEncRows];
    __shared__ float2 s_units[32];                                           // [(count << 1) | half]
    if (threadIdx.x < 32) {
        const float4 u = point_units_f32((int)(threadIdx.x >> 1));
        s_units[threadIdx.x] = (threadIdx.x & 1) ? make_float2(u.z, u.w) : make_float2(u.x, u.y);
    }
    if (n_rows_dev) B = min(B, (long long)*n_rows_dev);
    const int tr = threadIdx.x / 99, j = threadIdx.x - tr * 99;
    // kind 0: point pair (board byte bidx, half); 1: bar/2, off/15 of player `half`; 2: turn flags
    int kind = 0, bidx = 0, half = 0;
    if (j < 48) { bidx = j >> 1; half = j & 1; }
    else if (j == 48) { kind = 1; half = 0; }
    else if (j < 97) { bidx = 24 + ((j - 49) >> 1); half = (j - 49) & 1; }
    else if (j == 97) { kind = 1; half = 1; }
    else kind = 2;
    for (long long row0 = (long long)blockIdx.x * kEncRows; row0 < B; row0 += (long long)gridDim.x * kEncRows) {
        const int rows = (int)min((long long)kEncRows, B - row0);
        const uint32_t* src = reinterpret_cast<const uint32_t*>(boards + row0 * kBoardBytes);
        for (int i = threadIdx.x; i < rows * kBoardWords; i += kF32Threads) sm[i] = __ldg(src + i);
        for (int i = threadIdx.x; i < rows; i += kF32Threads) sflag[i] = flags ? (flags[row0 + i] & 1) : (int8_t)flag_all;
        __syncthreads();
        float* dst = out + row0 * ld + 2 * j;
#pragma unroll 4
        for (int r = tr; r < rows; r += kF32RowsPerPass) {
            const int8_t* b = reinterpret_cast<const int8_t*>(sm) + r * kBoardBytes;
            float2 v;
            if (kind == 0) {
                const int c = b[bidx];
                if ((unsigned)c < 16u) v = s_units[(c << 1) | half];
                else { const float4 u = point_units_f32(c); v = half ? make_float2(u.z, u.w) : make_float2(u.x, u.y); }
            } else if (kind == 1) {
                v = make_float2((float)b[48 + half] * 0.5f, __uint_as_float(kOff15F32[b[50 + half] & 15]));
            } else {
                v = sflag[r] == 0 ? make_float2(1.0f, 0.0f) : make_float2(0.0f, 1.0f);
            }
            *reinterpret_cast<float2*>(dst + (long long)r * ld) = v;
        }
        __syncthreads();
    }
}

}  // namespace bg

using namespace bg;

static unsigned enc_grid(long long B) {
    long long tiles = (B + kEncRows - 1) / kEncRows;
    long long g = (long long)bg_sm_count() * 8;
    return (unsigned)(tiles < g ? tiles : g);
}

extern "C" int bg_encode_f32(const int8_t* boards52, const int8_t* flags, int flag_all, long long B,
                             const unsigned long long* n_rows_dev, float* out, long long ld, void* stream) {
    if (B < 0 || ld < BG_FEATURES || (ld & 1)) return bg_set_error_msg(BG_ERR_INVALID, "bg_encode_f32: bad B or ld (need even ld >= 198)");
    if (B == 0) return BG_OK;
    if (!boards52 || !out) return bg_set_error_msg(BG_ERR_INVALID, "bg_encode_f32: null pointer");
    encode_f32_kernel<<<enc_grid(B), kF32Threads, 0, (cudaStream_t)stream>>>(boards52, flags, flag_all & 1, B, n_rows_dev, out, ld);
    return bg_set_error(cudaGetLastError(), "bg_encode_f32: launch");
}

int bg::encode_bf16_launch(const int8_t* boards52, const int8_t* flags, int flag_all, long long B,
                           const unsigned long long* row_begin_dev, const unsigned long long* n_rows_dev, uint16_t* out,
                           long long ld, cudaStream_t stream) {
    if (B < 0 || ld < 200 || (ld & 7) || ld > 3328) return bg_set_error_msg(BG_ERR_INVALID, "bg_encode_bf16: bad B or ld (need ld >= 200, multiple of 8)");
    if (B == 0) return BG_OK;
    if (!boards52 || !out) return bg_set_error_msg(BG_ERR_INVALID, "bg_encode_bf16: null pointer");
    static const int v5_env = getenv("BG_ENCODE_V5") ? atoi(getenv("BG_ENCODE_V5")) : 1;   // 0: the chunk-per-thread kernel for every ld
    if (v5_env && ld == BG_FEAT_LD_BF16 && !(reinterpret_cast<uintptr_t>(out) & 15u)) {
        const size_t smem = sizeof(V5Smem);
        cudaError_t e = cudaFuncSetAttribute(encode_bf16_v5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return bg_set_error(e, "bg_encode_bf16: cudaFuncSetAttribute");
        long long tiles = (B + kV5Rows - 1) / kV5Rows, grid = (long long)bg_sm_count() * 6;
        if (grid > tiles) grid = tiles;
        encode_bf16_v5_kernel<<<(unsigned)grid, kV5Threads, smem, stream>>>(boards52, flags, flag_all & 1, B, row_begin_dev, n_rows_dev, out);
        return bg_set_error(cudaGetLastError(), "bg_encode_bf16: launch");
    }
    const int cpr = (int)(ld / 8);
    long long tiles = (B + kBfRows - 1) / kBfRows;
    long long grid = (long long)bg_sm_count() * 3;
    if (grid > tiles) grid = tiles;
    encode_bf16_kernel<<<(unsigned)grid, kBfThreads, 0, stream>>>(boards52, flags, flag_all & 1, B, row_begin_dev, n_rows_dev, out, cpr);
    return bg_set_error(cudaGetLastError(), "bg_encode_bf16: launch");
}

extern "C" int bg_encode_bf16(const int8_t* boards52, const int8_t* flags, int flag_all, long long B,
                              const unsigned long long* n_rows_dev, uint16_t* out, long long ld, void* stream) {
    return bg::encode_bf16_launch(boards52, flags, flag_all, B, nullptr, n_rows_dev, out, ld, (cudaStream_t)stream);
}
