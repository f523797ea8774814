// Variant P decoder pieces shared by the fused renderer (render_fused.cu) and the occupancy-grid builder (density.cu).
#pragma once
#include "common.cuh"

namespace ssdnerf {

// ------------------------------------------------------------------------------------------------
// Variant P decoder: weights in shared memory, one sample per lane.
// blob layout (floats), see ssdnerf_b200/decoder_pack.py:
//   W1[18][64] (row k = plane*6 + c) | b1[64] | Wd[64] | bd,0,0,0 | Wdir[16][64] | bdir[64] | Wc[3][64] | bc[3],0 | sat,0,0,0
// ------------------------------------------------------------------------------------------------
struct DecP {
    static constexpr int C = 6, CPAD = 8, KF = 18, HID = 64;
    static constexpr int OFF_W1 = 0, OFF_B1 = OFF_W1 + KF * HID, OFF_WD = OFF_B1 + HID, OFF_BD = OFF_WD + HID,
                         OFF_WDIR = OFF_BD + 4, OFF_BDIR = OFF_WDIR + 16 * HID, OFF_WC = OFF_BDIR + HID,
                         OFF_BC = OFF_WC + 3 * HID, OFF_SAT = OFF_BC + 4, BLOB = OFF_SAT + 4;
};

constexpr int kWarpsPerCta = 4;
constexpr int kCtaThreads = kWarpsPerCta * 32;

struct BitfieldLoader {
    const uint8_t* __restrict__ g;
    __device__ __forceinline__ uint32_t operator()(uint32_t byte) const { return __ldg(g + byte); }
};

// one 32-byte texel (8 floats, 6 used) as ONE 256-bit read-only load (LDG.E.256, sm_100): half the load instructions and half the
// L1 wavefronts of two 128-bit loads -- the gather is the main client of the L1 data pipe (profiles/r02_render_p_analysis.txt)
struct Texel8 { float4 lo, hi; };
__device__ __forceinline__ Texel8 ldg_texel8(const float* __restrict__ p) {
    Texel8 t;
    asm("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
        : "=f"(t.lo.x), "=f"(t.lo.y), "=f"(t.lo.z), "=f"(t.lo.w), "=f"(t.hi.x), "=f"(t.hi.y), "=f"(t.hi.z), "=f"(t.hi.w) : "l"(p));
    return t;
}

// bilinear gather of one plane, fp32 channels-last with 8 floats per texel (6 used)
__device__ __forceinline__ void gather_plane_p(const float* __restrict__ plane, uint32_t Hp, uint32_t Wp,
                                               float u, float v, float* __restrict__ f) {
    // grid_sample(align_corners=False, padding_mode='border'): unnormalise, clip, bilinear
    float ix = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(u, 1.0f), (float)Wp), 1.0f), 0.5f);
    float iy = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(v, 1.0f), (float)Hp), 1.0f), 0.5f);
    ix = fminf((float)(Wp - 1), fmaxf(ix, 0.0f));
    iy = fminf((float)(Hp - 1), fmaxf(iy, 0.0f));
    const float fx0 = floorf(ix), fy0 = floorf(iy);
    const int x0 = (int)fx0, y0 = (int)fy0;
    const int x1 = min(x0 + 1, (int)Wp - 1), y1 = min(y0 + 1, (int)Hp - 1);
    const float wx1 = ix - fx0, wy1 = iy - fy0, wx0 = (fx0 + 1.0f) - ix, wy0 = (fy0 + 1.0f) - iy;
    const float nw = wx0 * wy0, ne = wx1 * wy0, sw = wx0 * wy1, se = wx1 * wy1;
    const Texel8 ta = ldg_texel8(plane + ((size_t)y0 * Wp + x0) * 8), tb = ldg_texel8(plane + ((size_t)y0 * Wp + x1) * 8);
    const Texel8 tc = ldg_texel8(plane + ((size_t)y1 * Wp + x0) * 8), td = ldg_texel8(plane + ((size_t)y1 * Wp + x1) * 8);
    const float4 a0 = ta.lo, a1 = ta.hi, b0 = tb.lo, b1 = tb.hi, c0 = tc.lo, c1 = tc.hi, d0 = td.lo, d1 = td.hi;
    f[0] = a0.x * nw + b0.x * ne + c0.x * sw + d0.x * se;
    f[1] = a0.y * nw + b0.y * ne + c0.y * sw + d0.y * se;
    f[2] = a0.z * nw + b0.z * ne + c0.z * sw + d0.z * se;
    f[3] = a0.w * nw + b0.w * ne + c0.w * sw + d0.w * se;
    f[4] = a1.x * nw + b1.x * ne + c1.x * sw + d1.x * se;
    f[5] = a1.y * nw + b1.y * ne + c1.y * sw + d1.y * se;
}


}  // namespace ssdnerf
