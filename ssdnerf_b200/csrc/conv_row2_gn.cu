// Fused GroupNorm(32) [+ NormWithEmbedding scale / shift] + SiLU + 3x3 convolution for the UNet's 128 x 128 level (128 output channels).
//
// Replaces on the reference path (mmgen DenoisingResBlock.forward as used by lib/models/architecture/ddpm/modules.py:51-110):
//     conv3x3( SiLU( GroupNorm(x) * (1 + scale) + shift ) )                      -- three kernels and a 2 x 67 MB round trip in between.
// The normalised activation never exists in global memory: eight loader warps read the RAW NHWC fp16 input rows (LDG.128), apply the
// per-(image, channel) affine a x + b in fp32 (coefficients built exactly like k_gn_apply's, GroupNorm statistics from the quad sums the
// producing GEMM's epilogue emitted), SiLU as h + h tanh(h), h = u / 2, on packed halves (tanh.approx.f16x2), and store the result into the K-major SWIZZLE_128B operand layout tcgen05.mma reads
// (16-byte chunk index XOR (row & 7)).  Zero padding is written as literal zeros (it pads the activation, not the raw input).
//
// Operand staging follows conv_row2.cu (tile = two image rows -> two TMEM accumulators per weight tile; a 130-pixel row serves the three
// horizontal taps through shifted descriptors) and adds vertical reuse: the 4 input rows y0-1 .. y0+2 of a 64-channel chunk are
// produced ONCE and serve all 9 taps (row r feeds accumulator a at tap ky = r - a).  Rows live in an 8-slot ring (17 KB each) and are
// released as soon as their last tap has been issued (row 0 after ky = 0, row 1 after ky = 1, rows 2, 3 after ky = 2), which leaves
// room for 4 weight stages.  Staged bytes per 128 x 128 x 64 product: 3.6 KB activations + 8 KB weights (generic tile kernel: 32 KB).
// The transform costs half a MUFU per element (33 k elements per (tile, chunk)); with fp32 ex2 + rcp SiLU (2 MUFU per element, 6 instructions)
// the eight loader warps were the bottleneck of the kernel (110 us vs 90 us for the two-pass composition).
#include "common.cuh"
#include "tc_common.cuh"
#include "conv_row_epilogue.cuh"
#include "../../include/ssdnerf_b200.h"
#include <cuda_fp16.h>
#include <cstdlib>

namespace ssdnerf {
using namespace tc;

constexpr int kGnThreads = 448;                          // warp 0 weight producer, 1 MMA, 2..5 epilogue, 6..13 activation loaders
constexpr int kGnLoaders = 256;
constexpr int kGnRowBytes = 130 * 128;                   // one activation row: 130 pixels (x = -1 .. 128) x 64 halves
constexpr int kGnRowSlot = 17 * 1024;                    // slot stride (1024-aligned)
constexpr int kGnBSlot = kRwN * 128;                     // 16 KB weight tile
constexpr int kGnMaxC = 384;
template <bool PAIR>
constexpr size_t gn_smem() {
    return (size_t)(PAIR ? 8 : 6) * kGnRowSlot + (size_t)(PAIR ? 8 * (kGnBSlot / 2) : 6 * kGnBSlot) + 1024 /*align*/ + 512 /*barriers*/ + 512 /*qacc*/ +
           512 /*bias*/ + 8 * 2048 /*epilogue staging*/;
}

struct ConvGnParams {
    uint32_t B, H;
    const __half* x1; uint32_t C1;      // raw inputs, NHWC [B][H][128][C]
    const __half* x2; uint32_t C2;      // optional second input (skip concat along channels)
    const float2* coef;                 // [B][C1 + C2] per-(image, channel) affine {a, b} of GroupNorm (+ scale / shift), from k_gn_coef
    const float* bias; const __half* residual; __half* out; float* qstats;
    unsigned long long* prof;           // optional debug counters: [0] MMA wait rows, [1] MMA wait weights, [2] MMA wait TMEM, [3] MMA total,
                                        // [4] loader wait free row, [5] loader total, [6] weight producer wait, [7] weight producer total
};

// PAIR: CTA pair with one tcgen05.mma.cta_group::2 (M = 256) per (tap, accumulator) for two consecutive tiles, as in conv_row2.cu: half the
// MMA instructions and half the staged weight bytes per SM; the leader's MMA waits for the loader warps of BOTH CTAs (remote arrivals).
// Single CTA: 6 rows + 6 weight tiles; pair: 8 rows + 8 half tiles (ring splits 6/6, 7/5, 8/4 measured identical).
template <bool PAIR>
__global__ void __launch_bounds__(kGnThreads, 1)
k_conv_row2_gn(const __grid_constant__ CUtensorMap mapB, const ConvGnParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    constexpr int kGnRowSlots = PAIR ? 8 : 6, kGnBStages = PAIR ? 8 : 6, kBSlot = PAIR ? kGnBSlot / 2 : kGnBSlot;
    uint8_t* sR = smem;                                               // activation row ring
    uint8_t* sB = smem + kGnRowSlots * kGnRowSlot;
    uint64_t* fullR = reinterpret_cast<uint64_t*>(sB + kGnBStages * kBSlot);
    uint64_t* emptyR = fullR + kGnRowSlots;
    uint64_t* fullB = emptyR + kGnRowSlots;
    uint64_t* emptyB = fullB + kGnBStages;
    uint64_t* tfull = emptyB + kGnBStages;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
    float* qacc = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(fullR) + 512);      // [32 quads][2]  (36 mbarriers + TMEM slot < 512 B)
    float* sbias = qacc + 128;                                                             // [128]
    uint8_t* sstage = reinterpret_cast<uint8_t*>(sbias + 128);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t tiles_per_img = p.H / 2, total_tiles = p.B * tiles_per_img;
    const uint32_t kc1 = p.C1 / 64, KC = (p.C1 + p.C2) / 64;
    const uint32_t crank = PAIR ? cluster_ctarank() : 0u;
    const uint32_t tile0 = PAIR ? cluster_id_x() * 2u + crank : blockIdx.x;
    const uint32_t tstep = PAIR ? num_clusters_x() * 2u : gridDim.x;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&mapB);
        for (int i = 0; i < kGnRowSlots; ++i) { mbar_init(&fullR[i], (kGnLoaders / 32) * (PAIR ? 2 : 1)); mbar_init(&emptyR[i], 1); }   // every loader warp (of both CTAs) arrives once per row
        for (int i = 0; i < kGnBStages; ++i) { mbar_init(&fullB[i], 1); mbar_init(&emptyB[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 128 * (PAIR ? 2 : 1)); }
        fence_mbar_init();
    }
    if (warp == 1) { if (PAIR) tmem_alloc2(tmem_slot, 512); else tmem_alloc(tmem_slot, 512); }
    for (int i = threadIdx.x; i < 64; i += kGnThreads) qacc[i] = 0.0f;
    for (int i = threadIdx.x; i < kRwN; i += kGnThreads) sbias[i] = p.bias ? __ldg(p.bias + i) : 0.0f;
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();
    pdl_wait();

    if (warp == 0) {   // ---------------- weight producer (TMA): for every (tile, chunk): 9 taps in (ky, kx) order
        uint32_t sb = 0, pb = 0;
        long long pw = 0; const long long q0 = clock64();
        for (uint32_t tile = tile0; tile < total_tiles; tile += tstep) {
            for (uint32_t j = 0; j < KC; ++j) {
                for (uint32_t tap = 0; tap < 9; ++tap) {
                    if (p.prof) { const long long c_ = clock64(); mbar_wait(&emptyB[sb], pb ^ 1); pw += clock64() - c_; } else mbar_wait(&emptyB[sb], pb ^ 1);
                    if (!PAIR) {
                        mbar_expect_tx_w(&fullB[sb], kGnBSlot);
                        tma_load_4d_w(sB + sb * kBSlot, &mapB, &fullB[sb], (int)(j * 64), 0, (int)tap, 0);
                    } else {   // own half of the weight tile, bytes credited to the leader's barrier
                        if (crank == 0) mbar_expect_tx_w(&fullB[sb], kGnBSlot);
                        tma_load_4d_2cta_w(sB + sb * kBSlot, &mapB, mapa_u32(smem_u32(&fullB[sb]), 0), (int)(j * 64), (int)(crank * 64), (int)tap, 0);
                    }
                    if (++sb == kGnBStages) { sb = 0; pb ^= 1; }
                }
            }
        }
        if (p.prof && lane == 0) { atomicAdd(p.prof + 6, (unsigned long long)pw); atomicAdd(p.prof + 7, (unsigned long long)(clock64() - q0)); }
        __syncwarp();
    } else if (warp == 1) {   // ---------------- MMA issuer (converged warp, elected lane; pair: the leader issues for both SMs)
      if (!PAIR || crank == 0) {
        constexpr uint32_t idesc = make_idesc_f16(PAIR ? 256 : 128, kRwN);
        constexpr uint16_t kBoth = 3;
        uint32_t sb = 0, pb = 0, acc = 0, acc_phase = 0, rctr = 0;           // rctr: rows consumed so far (ring position of row 0 of the chunk)
        long long wr = 0, wb = 0, wt = 0; const long long m0 = clock64();
#define GN_TIMED(acc_var, stmt) do { if (p.prof) { const long long c_ = clock64(); stmt; acc_var += clock64() - c_; } else { stmt; } } while (0)
        for (uint32_t tile = tile0; tile < total_tiles; tile += tstep) {
            if (PAIR) GN_TIMED(wt, mbar_wait_cluster(&tempty[acc], acc_phase ^ 1)); else GN_TIMED(wt, mbar_wait(&tempty[acc], acc_phase ^ 1));
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * 256;
            uint32_t started = 0;
            for (uint32_t j = 0; j < KC; ++j) {
                uint32_t slot[4], rphase[4];
#pragma unroll
                for (uint32_t r = 0; r < 4; ++r) { slot[r] = (rctr + r) % kGnRowSlots; rphase[r] = ((rctr + r) / kGnRowSlots) & 1u; }
#pragma unroll
                for (uint32_t ky = 0; ky < 3; ++ky) {
                    if (ky == 0) { GN_TIMED(wr, mbar_wait_cluster(&fullR[slot[0]], rphase[0])); GN_TIMED(wr, mbar_wait_cluster(&fullR[slot[1]], rphase[1])); }
                    else GN_TIMED(wr, mbar_wait_cluster(&fullR[slot[ky + 1]], rphase[ky + 1]));
                    for (uint32_t kx = 0; kx < 3; ++kx) {
                        GN_TIMED(wb, mbar_wait(&fullB[sb], pb));
                        tc_fence_after();
                        const uint64_t b_desc = make_desc_sw128(smem_u32(sB + sb * kBSlot));
#pragma unroll
                        for (uint32_t a = 0; a < 2; ++a) {
                            const uint64_t a_desc = make_desc_sw128(smem_u32(sR + slot[ky + a] * kGnRowSlot) + kx * 128);
#pragma unroll
                            for (uint32_t k = 0; k < 4; ++k) {
                                if (PAIR) umma_f16_2cta_w(d_tmem + a * kRwN, a_desc + 2 * k, b_desc + 2 * k, idesc, (started | k) != 0);
                                else umma_f16_w(d_tmem + a * kRwN, a_desc + 2 * k, b_desc + 2 * k, idesc, (started | k) != 0);
                            }
                        }
                        started = 1;
                        if (PAIR) umma_commit_2cta_w(&emptyB[sb], kBoth); else umma_commit_w(&emptyB[sb]);
                        if (++sb == kGnBStages) { sb = 0; pb ^= 1; }
                    }
                    // rows whose last tap has been issued go back to the loaders
                    if (PAIR) umma_commit_2cta_w(&emptyR[slot[ky]], kBoth); else umma_commit_w(&emptyR[slot[ky]]);
                    if (ky == 2) { if (PAIR) umma_commit_2cta_w(&emptyR[slot[3]], kBoth); else umma_commit_w(&emptyR[slot[3]]); }
                }
                rctr += 4;
            }
            if (PAIR) umma_commit_2cta_w(&tfull[acc], kBoth); else umma_commit_w(&tfull[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (p.prof && lane == 0) {
            atomicAdd(p.prof + 0, (unsigned long long)wr); atomicAdd(p.prof + 1, (unsigned long long)wb); atomicAdd(p.prof + 2, (unsigned long long)wt);
            atomicAdd(p.prof + 3, (unsigned long long)(clock64() - m0));
        }
      }
        __syncwarp();
    } else if (warp < 6) {   // ---------------- epilogue warps 2..5 (conv_row_epilogue.cuh): hidden behind the ~2x longer main loop of a tile
        const RowEpiArgs ea{p.H, p.residual, p.out, p.qstats};
        conv_row_epilogue<4, PAIR>(ea, warp, lane, total_tiles, tiles_per_img, tmem_base, tfull, tempty, sstage, sbias, qacc, tile0, tstep);
    } else {   // ---------------- activation loaders (warps 6..13): raw rows -> GroupNorm affine + SiLU -> swizzled operand rows
        const uint32_t lt = threadIdx.x - 192u;                       // 0..255
        const uint32_t c8 = lt & 7u, p0 = lt >> 3;                    // 16-byte chunk (8 channels) within the 64-channel row; pixels p0 + 32 n
        const uint32_t C = p.C1 + p.C2;
        // this CTA's rows form ONE stream over (tile, 64-channel chunk, row y0-1 .. y0+2); raw loads run two rows and the affine
        // coefficients one chunk ahead of the transform, across tile boundaries
        const uint32_t my_tiles = tile0 < total_tiles ? (total_tiles - tile0 + tstep - 1) / tstep : 0;
        const uint32_t rpt = KC * 4u, n_rows = my_tiles * rpt;
        // stream position = (tile, row-within-tile w = 4 j + r), advanced without divisions
        struct Cur { uint32_t tile, w; };
        auto advance = [&](Cur& c) { if (++c.w == rpt) { c.w = 0; c.tile += tstep; } };
        auto issue_row = [&](const Cur& c, uint4* v) {                // raw 16-byte chunks of the row for pixels p0 + 32 n (zero outside the image)
            const uint32_t j = c.w >> 2, r = c.w & 3u;
            const uint32_t b = c.tile / tiles_per_img, y0 = (c.tile - b * tiles_per_img) * 2;
            const bool first = j < kc1;
            const __half* src = first ? p.x1 : p.x2;
            const uint32_t Cs = first ? p.C1 : p.C2, cl = (first ? j : j - kc1) * 64u + c8 * 8u;
            const int y = (int)(y0 + r) - 1;
            const bool row_ok = y >= 0 && y < (int)p.H;
            const __half* rowp = src + (((size_t)b * p.H + (row_ok ? y : 0)) * kRwW) * Cs + cl;
#pragma unroll
            for (int n = 0; n < 5; ++n) {
                const int x = (int)(p0 + 32u * n) - 1;
                v[n] = (row_ok && (unsigned)x < (unsigned)kRwW) ? __ldg(reinterpret_cast<const uint4*>(rowp + (size_t)x * Cs)) : make_uint4(0, 0, 0, 0);
            }
        };
        auto issue_coef = [&](const Cur& c, float4* cf) {             // {a, b} of this thread's 8 channels for the chunk the row belongs to
            const uint32_t b = c.tile / tiles_per_img;
            const float4* cp = reinterpret_cast<const float4*>(p.coef + (size_t)b * C + (c.w >> 2) * 64u + c8 * 8u);
#pragma unroll
            for (int k = 0; k < 4; ++k) cf[k] = __ldg(cp + k);
        };
        uint4 vcur[5], vn1[5], vn2[5];
        float4 cn[4];
        float ca[8], cb[8];
        Cur cu{tile0, 0}, cl2{tile0, 0}, cc{tile0, 0};      // transform cursor, load cursor (2 rows ahead), coefficient cursor (next chunk)
        if (n_rows) {
            issue_coef(cc, cn);
            issue_row(cl2, vcur); advance(cl2);
            if (n_rows > 1) { issue_row(cl2, vn1); advance(cl2); }
        }
        uint32_t rctr = 0;
        long long lw = 0; const long long l0 = clock64();
        for (uint32_t g = 0; g < n_rows; ++g) {
            const uint32_t r = cu.w & 3u;
            const uint32_t b = cu.tile / tiles_per_img, y0 = (cu.tile - b * tiles_per_img) * 2;
            if (r == 0) {
#pragma unroll
                for (int k = 0; k < 4; ++k) { ca[2 * k] = 0.5f * cn[k].x; cb[2 * k] = 0.5f * cn[k].y; ca[2 * k + 1] = 0.5f * cn[k].z; cb[2 * k + 1] = 0.5f * cn[k].w; }
                cc = cu; cc.w += 3; advance(cc);                      // first row of the next chunk
            }
            if (r == 1 && g + 3 < n_rows) issue_coef(cc, cn);
            if (g + 2 < n_rows) { issue_row(cl2, vn2); advance(cl2); }
            const uint32_t slot = rctr % kGnRowSlots, ph = (rctr / kGnRowSlots) & 1u;
            if (p.prof) { const long long c_ = clock64(); mbar_wait(&emptyR[slot], ph ^ 1); lw += clock64() - c_; } else mbar_wait(&emptyR[slot], ph ^ 1);
            const int y = (int)(y0 + r) - 1;
            const bool row_ok = y >= 0 && y < (int)p.H;
            const uint32_t dst = smem_u32(sR + slot * kGnRowSlot);
            // branch-free transform of the 5 chunks (independent dependency chains the scheduler can interleave); padding pixels are forced to zero
#pragma unroll
            for (int n = 0; n < 5; ++n) {
                const uint32_t px = p0 + 32u * n;
                const bool ok = row_ok && (unsigned)((int)px - 1) < (unsigned)kRwW;
                const __half2* h = reinterpret_cast<const __half2*>(&vcur[n]);
                uint4 o;
                uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    // affine in fp32 (half of it: ca / cb carry the factor 0.5), SiLU(u) = h + h tanh(h) with h = u / 2 on packed halves:
                    // one MUFU per two elements instead of four; result within 2 fp16 ulps of the fp32 evaluation of k_gn_apply
                    const float2 tt = __half22float2(h[k]);
                    const __half2 hh = __floats2half2_rn(fmaf(tt.x, ca[2 * k], cb[2 * k]), fmaf(tt.y, ca[2 * k + 1], cb[2 * k + 1]));
                    uint32_t hv = *reinterpret_cast<const uint32_t*>(&hh), tv;
                    asm("tanh.approx.f16x2 %0, %1;" : "=r"(tv) : "r"(hv));
                    const __half2 yy = __hfma2(hh, *reinterpret_cast<const __half2*>(&tv), hh);
                    ow[k] = ok ? *reinterpret_cast<const uint32_t*>(&yy) : 0u;
                }
                if (n < 4 || p0 < 2u) sts128(dst + px * 128u + ((c8 ^ (px & 7u)) << 4), o);
            }
            fence_proxy_async_smem();            // generic-proxy stores -> visible to the tensor core's async-proxy reads
            __syncwarp();
            if (lane == 0) { if (PAIR) mbar_arrive_cluster(mapa_u32(smem_u32(&fullR[slot]), 0)); else mbar_arrive(&fullR[slot]); }
            ++rctr;
            advance(cu);
#pragma unroll
            for (int n = 0; n < 5; ++n) { vcur[n] = vn1[n]; vn1[n] = vn2[n]; }
        }
        if (p.prof && lt == 0) { atomicAdd(p.prof + 4, (unsigned long long)lw); atomicAdd(p.prof + 5, (unsigned long long)(clock64() - l0)); }
    }
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();
    if (warp == 1) { tc_fence_after(); if (PAIR) tmem_dealloc2(tmem_base, 512); else tmem_dealloc(tmem_base, 512); }
}

// per-(image, channel) GroupNorm affine, k_gn_apply arithmetic: y = x * a + b,
//   a = rstd * gamma * (1 + scale), b = (beta - mean * rstd * gamma) * (1 + scale) + shift;  grid = images, one thread per channel (looped)
__global__ void k_gn_coef(const float* __restrict__ q1, uint32_t C1, const float* __restrict__ q2, uint32_t C2, uint32_t HW,
                          const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ scale_shift,
                          long long ss_batch_stride, float eps, float2* __restrict__ coef) {
    pdl_trigger();
    pdl_wait();
    __shared__ float2 s_mr[32];
    const uint32_t b = blockIdx.x, C = C1 + C2, cpg = C / 32u;
    const float inv_n = 1.0f / ((float)HW * (float)cpg);
    if (threadIdx.x < 32u) {
        const uint32_t q1n = C1 / 4, q2n = C2 / 4, nq = cpg / 4;
        float sm = 0.0f, sq = 0.0f;
        for (uint32_t i = 0; i < nq; ++i) {
            const uint32_t qi = threadIdx.x * nq + i;
            const float2 t = __ldg(reinterpret_cast<const float2*>(qi < q1n ? q1 + ((size_t)b * q1n + qi) * 2 : q2 + ((size_t)b * q2n + (qi - q1n)) * 2));
            sm += t.x; sq += t.y;
        }
        const float mean = sm * inv_n;
        s_mr[threadIdx.x] = make_float2(mean, rsqrtf(fmaxf(sq * inv_n - mean * mean, 0.0f) + eps));
    }
    __syncthreads();
    const float* ss = scale_shift ? scale_shift + (size_t)b * ss_batch_stride : nullptr;
    for (uint32_t c = threadIdx.x; c < C; c += blockDim.x) {
        const float2 mr = s_mr[c / cpg];
        const float ak = mr.y * __ldg(gamma + c);
        const float bk = __ldg(beta + c) - mr.x * ak;
        const float sc = ss ? 1.0f + __ldg(ss + c) : 1.0f, sh = ss ? __ldg(ss + C + c) : 0.0f;
        coef[(size_t)b * C + c] = make_float2(ak * sc, fmaf(bk, sc, sh));
    }
}

int make_map_4d_box(CUtensorMap* m, const void* base, uint64_t K, uint64_t e1, uint64_t e2, uint64_t e3, uint64_t s1, uint64_t s2, uint64_t s3,
                    uint32_t x1, uint32_t x2, uint32_t x3);

}  // namespace ssdnerf

using namespace ssdnerf;

extern "C" int ssdnerf_conv3x3_gn_f16(const ssdnerf_conv_gn_args* a, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!a || !a->x1 || !a->q1 || !a->gamma || !a->beta || !a->w || !a->out || !a->coef_workspace)
        return set_error_msg(SSDNERF_ERR_ARG, "conv3x3_gn: x1, q1, gamma, beta, w, out and coef_workspace are required");
    if (a->B == 0 || a->H == 0) return 0;
    if (a->H % 2) return set_error_msg(SSDNERF_ERR_ARG, "conv3x3_gn: H must be even (tiles are row pairs)");
    if (a->C1 == 0 || a->C1 % 64 || a->C2 % 64 || (a->x2 && !a->q2)) return set_error_msg(SSDNERF_ERR_ARG, "conv3x3_gn: channel counts must be multiples of 64; x2 needs q2");
    const uint32_t C = a->C1 + (a->x2 ? a->C2 : 0);
    if (C > (uint32_t)kGnMaxC || (C / 32) % 4) return set_error_msg(SSDNERF_ERR_ARG, "conv3x3_gn: at most 384 input channels, channels per group a multiple of 4");
    if (a->w_rows < 128) return set_error_msg(SSDNERF_ERR_ARG, "conv3x3_gn: packed weight needs >= 128 rows per tap");
    if (((uintptr_t)a->x1 | (uintptr_t)a->x2 | (uintptr_t)a->out | (uintptr_t)a->residual | (uintptr_t)a->w | (uintptr_t)a->coef_workspace) & 15u)
        return set_error_msg(SSDNERF_ERR_ARG, "conv3x3_gn: tensors must be 16-byte aligned");
    ConvGnParams p{};
    p.B = a->B; p.H = a->H; p.x1 = (const __half*)a->x1; p.C1 = a->C1; p.x2 = (const __half*)a->x2; p.C2 = a->x2 ? a->C2 : 0;
    p.coef = (const float2*)a->coef_workspace; p.bias = a->bias; p.prof = (unsigned long long*)a->debug_cycles; p.residual = (const __half*)a->residual; p.out = (__half*)a->out; p.qstats = a->qstats;
    int dev = 0, sms = 0;
    SSDNERF_CUDA_OK(cudaGetDevice(&dev));
    SSDNERF_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    SSDNERF_CUDA_OK(launch_pdl(k_gn_coef, dim3(a->B), dim3(128), 0, stream, a->q1, a->C1, a->x2 ? a->q2 : (const float*)nullptr, p.C2, a->H * 128u,
                               a->gamma, a->beta, a->scale_shift, a->ss_batch_stride, a->eps, (float2*)a->coef_workspace));
    SSDNERF_LAUNCH_OK();
    const uint32_t total = p.B * (p.H / 2);
    static int pair_env = -1;
    if (pair_env < 0) { const char* e = getenv("SSDNERF_GN_PAIR"); pair_env = e ? atoi(e) : 0; }     // CTA-pair variant: opt-in until validated on hardware
    const bool pair = pair_env && (p.H / 2) % 2 == 0 && total >= 2;
    CUtensorMap mB;
    if (int e = make_map_4d_box(&mB, a->w, C, a->w_rows, 9, 1, (uint64_t)C * 2, (uint64_t)a->w_rows * C * 2, (uint64_t)9 * a->w_rows * C * 2,
                                pair ? kRwN / 2 : kRwN, 1, 1)) return e;
    if (pair) {
        static DeviceOnce attr;
        if (attr.first()) {
            SSDNERF_CUDA_OK(cudaFuncSetAttribute(k_conv_row2_gn<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gn_smem<true>()));
        }
        const uint32_t clusters = (total / 2 < (uint32_t)sms / 2) ? total / 2 : (uint32_t)sms / 2;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(clusters * 2); cfg.blockDim = dim3(kGnThreads); cfg.dynamicSmemBytes = gn_smem<true>(); cfg.stream = stream;
        cudaLaunchAttribute at[2];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
        cfg.attrs = at; cfg.numAttrs = 2;
        SSDNERF_CUDA_OK(cudaLaunchKernelEx(&cfg, k_conv_row2_gn<true>, mB, p));
    } else {
        static DeviceOnce attr;
        if (attr.first()) {
            SSDNERF_CUDA_OK(cudaFuncSetAttribute(k_conv_row2_gn<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gn_smem<false>()));
        }
        SSDNERF_CUDA_OK(launch_pdl(k_conv_row2_gn<false>, dim3(total < (uint32_t)sms ? total : (uint32_t)sms), dim3(kGnThreads), gn_smem<false>(), stream, mB, p));
    }
    SSDNERF_LAUNCH_OK();
    return 0;
}
