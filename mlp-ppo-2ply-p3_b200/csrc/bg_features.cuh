// bg_features.cuh -- exact lookup tables for the 198-feature encoding (board/immutable_board.py:171-212,
// ai/batching.py:78-147), shared by K3 (encode.cu) and K4 (mlp.cu).  Counts are 0..15 in any legal position;
// tables are indexed with (count & 15).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace bg {

// 4 bf16 units of a point with c men: c==1 -> 1,0,0,0; c==2 -> 1,1,0,0; c>=3 -> 1,1,1,(c-3)/2  (x = units 0,1; y = units 2,3)
__device__ __constant__ uint2 kUnitsBf16[16] = {
    {0x00000000u, 0x00000000u}, {0x00003F80u, 0x00000000u}, {0x3F803F80u, 0x00000000u}, {0x3F803F80u, 0x00003F80u},
    {0x3F803F80u, 0x3F003F80u}, {0x3F803F80u, 0x3F803F80u}, {0x3F803F80u, 0x3FC03F80u}, {0x3F803F80u, 0x40003F80u},
    {0x3F803F80u, 0x40203F80u}, {0x3F803F80u, 0x40403F80u}, {0x3F803F80u, 0x40603F80u}, {0x3F803F80u, 0x40803F80u},
    {0x3F803F80u, 0x40903F80u}, {0x3F803F80u, 0x40A03F80u}, {0x3F803F80u, 0x40B03F80u}, {0x3F803F80u, 0x40C03F80u}};
// RN_bf16(RN_f32(k / 15)) -- what ref.to(torch.bfloat16) gives for the borne-off feature
__device__ __constant__ uint16_t kOff15Bf16[16] = {0x0000, 0x3D89, 0x3E09, 0x3E4D, 0x3E89, 0x3EAB, 0x3ECD, 0x3EEF,
                                                   0x3F09, 0x3F1A, 0x3F2B, 0x3F3C, 0x3F4D, 0x3F5E, 0x3F6F, 0x3F80};
// k / 2 in bf16 (exact)
__device__ __constant__ uint16_t kHalfBf16[16] = {0x0000, 0x3F00, 0x3F80, 0x3FC0, 0x4000, 0x4020, 0x4040, 0x4060,
                                                  0x4080, 0x4090, 0x40A0, 0x40B0, 0x40C0, 0x40D0, 0x40E0, 0x40F0};
// RN_f32(k / 15) bit patterns (IEEE division, as torch computes borne_off.float() / 15.0)
__device__ __constant__ uint32_t kOff15F32[16] = {0x00000000u, 0x3D888889u, 0x3E088889u, 0x3E4CCCCDu, 0x3E888889u, 0x3EAAAAABu,
                                                  0x3ECCCCCDu, 0x3EEEEEEFu, 0x3F088889u, 0x3F19999Au, 0x3F2AAAABu, 0x3F3BBBBCu,
                                                  0x3F4CCCCDu, 0x3F5DDDDEu, 0x3F6EEEEFu, 0x3F800000u};

// bar/off feature pair of one side as two packed bf16
__device__ __forceinline__ uint32_t bar_off_pair_bf16(int bar, int off) {
    return (uint32_t)kHalfBf16[bar & 15] | ((uint32_t)kOff15Bf16[off & 15] << 16);
}
// copy the point-unit table into shared memory (conflict-free random access: 16 entries x 8 B = 32 banks)
__device__ __forceinline__ void load_units_lut(uint2* s_units) {
    if (threadIdx.x < 16) s_units[threadIdx.x] = kUnitsBf16[threadIdx.x];
}

// 16-byte chunk k (features 8k .. 8k+7) of the bf16 feature row of board b (52 bytes), lut = shared-memory copy
__device__ __forceinline__ uint4 feature_chunk_lut(const int8_t* b, int flag, int k, const uint2* lut) {
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    if (k < 12) {                                    // PLAYER1 points 2k, 2k+1
        uint2 a = lut[b[2 * k] & 15], c = lut[b[2 * k + 1] & 15];
        o = make_uint4(a.x, a.y, c.x, c.y);
    } else if (k == 12) {                            // bar1/2, off1/15, P2 point 0, first half of P2 point 1
        uint2 a = lut[b[24] & 15], c = lut[b[25] & 15];
        o = make_uint4(bar_off_pair_bf16(b[48], b[50]), a.x, a.y, c.x);
    } else if (k < 24) {                             // second half of P2 point q, P2 point q+1, first half of q+2
        int q = 2 * (k - 12) - 1;
        uint2 a = lut[b[24 + q] & 15], c = lut[b[25 + q] & 15], e = lut[b[26 + q] & 15];
        o = make_uint4(a.y, c.x, c.y, e.x);
    } else if (k == 24) {                            // second half of P2 point 23, bar2/2, off2/15, flags, pad
        o.x = lut[b[47] & 15].y;
        o.y = bar_off_pair_bf16(b[49], b[51]);
        o.z = flag == 0 ? 0x00003F80u : 0x3F800000u;
    }
    return o;
}


// ---- table-driven chunk builder (uniform code for every chunk of a row) ----------------------------------
// byte w of kChunkDesc[k] describes word w (2 features) of 16-byte chunk k: bit 7 clear -> bits 0..5 = board52 byte
// of the point, bit 6 = which half of its 4 units; bit 7 set -> 0 zero, 1 bar1/off1, 2 bar2/off2, 3 turn flags.
__device__ __constant__ uint32_t kChunkDesc[26] = {0x41014000u, 0x43034202u, 0x45054404u, 0x47074606u, 0x49094808u, 0x4B0B4A0Au, 0x4D0D4C0Cu, 0x4F0F4E0Eu, 0x51115010u, 0x53135212u, 0x55155414u, 0x57175616u, 0x19581881u, 0x1B5A1A59u, 0x1D5C1C5Bu, 0x1F5E1E5Du, 0x2160205Fu, 0x23622261u, 0x25642463u, 0x27662665u, 0x29682867u, 0x2B6A2A69u, 0x2D6C2C6Bu, 0x2F6E2E6Du, 0x8083826Fu, 0x80808080u};

// shared-memory copies: lut[(count << 1) | half] -> two packed bf16 units; desc[k] (k >= 26: zero chunk)
__device__ __forceinline__ void load_chunk_tables(uint32_t* s_lut, uint32_t* s_desc) {
    if (threadIdx.x < 16) { uint2 u = kUnitsBf16[threadIdx.x]; s_lut[2 * threadIdx.x] = u.x; s_lut[2 * threadIdx.x + 1] = u.y; }
    if (threadIdx.x < 32) s_desc[threadIdx.x] = threadIdx.x < 26 ? kChunkDesc[threadIdx.x] : 0x80808080u;
}
__device__ __forceinline__ uint4 chunk_from_desc(const uint8_t* b, int flag, uint32_t desc, const uint32_t* s_lut) {
    uint32_t w[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint32_t d = (desc >> (8 * q)) & 0xFFu;
        w[q] = s_lut[((b[d & 63u] & 15u) << 1) | ((d >> 6) & 1u)];
    }
    if (desc & 0x80808080u) {                                // chunks 12, 24, 25 (and padding): bar/off, flags, zeros
        const uint32_t sp1 = bar_off_pair_bf16(b[48], b[50]), sp2 = bar_off_pair_bf16(b[49], b[51]);
        const uint32_t sp3 = flag == 0 ? 0x00003F80u : 0x3F800000u;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t d = (desc >> (8 * q)) & 0xFFu;
            if (d & 0x80u) { const uint32_t code = d & 3u; w[q] = code == 1 ? sp1 : (code == 2 ? sp2 : (code == 3 ? sp3 : 0u)); }
        }
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

}  // namespace bg
