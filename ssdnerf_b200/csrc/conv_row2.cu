// 3x3 convolution, 128-pixel-wide images, 128 output channels (the UNet's 128 x 128 level): row-pair kernel with horizontal halo reuse.
//
// The generic implicit-GEMM kernel (gemm_tc.cu) stages, per 64-channel k-block and per tap, a 128 x 64 activation tile and a
// 128 x 64 weight tile for ONE 128 x 128 x 64 product: 32 KB of shared memory traffic per 2.1 MFLOP, and its operand pipeline
// (shared-memory stages x bytes per stage / load round trip) saturates at ~58 B/clk/SM -- half of what the tensor pipe could eat
// at N = 128 (profiles/r01_gemm_pipeline_*.txt).  This kernel makes every staged byte do 2.4x more work:
//   * a tile is TWO image rows (y0, y0+1) of one image -> two TMEM accumulators share every weight tile;
//   * per (ky, 64-channel chunk) ONE activation box of 2 rows x 130 pixels (x = -1 .. 128, zero-filled out of bounds by TMA) serves
//     the three horizontal taps: tap kx reads rows [kx, kx + 128) of the box through a shifted shared-memory descriptor
//     (K-major SWIZZLE_128B, start address advanced by kx rows of 128 B; the matrix-base-offset field stays 0 -- measured: the
//     128B swizzle phase of a row is taken from its absolute shared-memory address, exactly as TMA wrote it).
// Staged bytes per (ky, chunk): 33 KB activations + 3 x 16 KB weights for 6 products of 128 x 128 x 64 (13.5 KB per product).
// Roles as in gemm_tc.cu: warp 0 TMA producer, warp 1 MMA issuer (converged warp, elected lane), warps 2..9 epilogue
// (bias, residual, fp16 store, fused GroupNorm quad statistics).
#include "common.cuh"
#include "tc_common.cuh"
#include "../../include/ssdnerf_b200.h"
#include <cuda_fp16.h>

namespace ssdnerf {
using namespace tc;

constexpr int kRwThreads = 320, kRwEpi = 256;
constexpr int kRwW = 128, kRwN = 128;                   // image width (pixels per row) and output channels
constexpr int kRwABox = 2 * 130 * 128;                  // bytes of one activation box: 2 rows x 130 pixels x 64 halves
constexpr int kRwASlot = 33 * 1024;                     // slot stride (1024-aligned)
constexpr int kRwBSlot = kRwN * 128;                    // 16 KB weight tile
constexpr int kRwAStages = 2, kRwBStages = 7;
constexpr size_t kRwSmem = (size_t)kRwAStages * kRwASlot + (size_t)kRwBStages * kRwBSlot + 1024 /*align*/ + 256 /*barriers*/ + 512 /*qacc*/ + 512 /*bias*/ + 8 * 2048 /*epilogue staging*/;

struct ConvRowParams {
    uint32_t B, H;                // images, rows (H even)
    uint32_t kc1, kc2;            // 64-channel chunks from input 1 / input 2 (skip concat)
    const float* bias;            // [128] or NULL
    const __half* residual;       // NHWC [B][H][128][128] or NULL
    __half* out;                  // NHWC [B][H][128][128]
    float* qstats;                // optional [B][32][2]
    unsigned long long* prof;
};

__global__ void __launch_bounds__(kRwThreads, 1)
k_conv_row2(const __grid_constant__ CUtensorMap mapA1, const __grid_constant__ CUtensorMap mapA2, const __grid_constant__ CUtensorMap mapB,
            const ConvRowParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;
    uint8_t* sB = smem + kRwAStages * kRwASlot;
    uint64_t* fullA = reinterpret_cast<uint64_t*>(sB + kRwBStages * kRwBSlot);
    uint64_t* emptyA = fullA + kRwAStages;
    uint64_t* fullB = emptyA + kRwAStages;
    uint64_t* emptyB = fullB + kRwBStages;
    uint64_t* tfull = emptyB + kRwBStages;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
    float* qacc = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(fullA) + 256);      // [32 quads][2]
    float* sbias = qacc + 128;                                                             // [128]
    uint8_t* sstage = reinterpret_cast<uint8_t*>(sbias + 128);                              // [8 warps][32 rows][64 B]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t tiles_per_img = p.H / 2, total_tiles = p.B * tiles_per_img;
    const uint32_t KC = p.kc1 + p.kc2;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&mapA1); prefetch_tmap(&mapA2); prefetch_tmap(&mapB);
        for (int i = 0; i < kRwAStages; ++i) { mbar_init(&fullA[i], 1); mbar_init(&emptyA[i], 1); }
        for (int i = 0; i < kRwBStages; ++i) { mbar_init(&fullB[i], 1); mbar_init(&emptyB[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], kRwEpi); }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    for (int i = threadIdx.x; i < 64; i += kRwThreads) qacc[i] = 0.0f;
    for (int i = threadIdx.x; i < kRwN; i += kRwThreads) sbias[i] = p.bias ? __ldg(p.bias + i) : 0.0f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();
    pdl_wait();

    if (warp == 0) {   // ---------------- TMA producer
        uint32_t sa = 0, pa = 0, sb = 0, pb = 0;
        for (uint32_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const uint32_t b = tile / tiles_per_img, y0 = (tile - b * tiles_per_img) * 2;
            for (uint32_t ky = 0; ky < 3; ++ky) {
                for (uint32_t j = 0; j < KC; ++j) {
                    mbar_wait(&emptyA[sa], pa ^ 1);
                    mbar_expect_tx_w(&fullA[sa], kRwABox);
                    const bool first = j < p.kc1;
                    tma_load_4d_w(sA + sa * kRwASlot, first ? &mapA1 : &mapA2, &fullA[sa], (int)((first ? j : j - p.kc1) * 64), -1,
                                  (int)(y0 + ky) - 1, (int)b);
                    if (++sa == kRwAStages) { sa = 0; pa ^= 1; }
                    for (uint32_t kx = 0; kx < 3; ++kx) {
                        mbar_wait(&emptyB[sb], pb ^ 1);
                        mbar_expect_tx_w(&fullB[sb], kRwBSlot);
                        tma_load_4d_w(sB + sb * kRwBSlot, &mapB, &fullB[sb], (int)(j * 64), 0, (int)(ky * 3 + kx), 0);
                        if (++sb == kRwBStages) { sb = 0; pb ^= 1; }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {   // ---------------- MMA issuer
        constexpr uint32_t idesc = make_idesc_f16(128, kRwN);
        uint32_t sa = 0, pa = 0, sb = 0, pb = 0, acc = 0, acc_phase = 0;
        long long wf = 0; const long long mt0 = clock64();
        for (uint32_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            mbar_wait(&tempty[acc], acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * 256;
            uint32_t started = 0;
            for (uint32_t ky = 0; ky < 3; ++ky) {
                for (uint32_t j = 0; j < KC; ++j) {
                    if (p.prof) { const long long c = clock64(); mbar_wait(&fullA[sa], pa); wf += clock64() - c; } else mbar_wait(&fullA[sa], pa);
                    const uint32_t a_base = smem_u32(sA + sa * kRwASlot);
                    for (uint32_t kx = 0; kx < 3; ++kx) {
                        if (p.prof) { const long long c = clock64(); mbar_wait(&fullB[sb], pb); wf += clock64() - c; } else mbar_wait(&fullB[sb], pb);
                        tc_fence_after();
                        const uint64_t b_desc = make_desc_sw128(smem_u32(sB + sb * kRwBSlot));
#pragma unroll
                        for (uint32_t a = 0; a < 2; ++a) {
                            const uint32_t row0 = a * 130 + kx;
                            const uint64_t a_desc = make_desc_sw128(a_base + row0 * 128);
#pragma unroll
                            for (uint32_t k = 0; k < 4; ++k) umma_f16_w(d_tmem + a * kRwN, a_desc + 2 * k, b_desc + 2 * k, idesc, (started | k) != 0);
                        }
                        started = 1;
                        umma_commit_w(&emptyB[sb]);
                        if (++sb == kRwBStages) { sb = 0; pb ^= 1; }
                    }
                    umma_commit_w(&emptyA[sa]);
                    if (++sa == kRwAStages) { sa = 0; pa ^= 1; }
                }
            }
            umma_commit_w(&tfull[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (p.prof && lane == 0) { atomicAdd(p.prof + 2, (unsigned long long)wf); atomicAdd(p.prof + 4, (unsigned long long)(clock64() - mt0)); }
        __syncwarp();
    } else {   // ---------------- epilogue warps 2..9: TMEM lane quarter = warp % 4 (pixels), column half = (warp - 2) / 4
        const uint32_t q = (uint32_t)warp & 3u, hsel = (uint32_t)(warp - 2) >> 2;
        const uint32_t x = q * 32 + (uint32_t)lane;
        uint32_t acc = 0, acc_phase = 0;
        for (uint32_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const uint32_t b = tile / tiles_per_img, y0 = (tile - b * tiles_per_img) * 2;
            // global accesses re-mapped through a per-warp staging tile: one warp instruction = 8 pixels x 64 contiguous bytes (see gemm_tc.cu)
            const uint32_t wst = smem_u32(sstage) + (uint32_t)(warp - 2) * 2048u;     // shared-space address of this warp's staging tile
            const uint32_t pc = (uint32_t)lane & 3u;
            uint32_t st_own[4], st_map[4];     // swizzled byte offsets: own row (lane) piece g / re-mapped row (lane >> 2) + 8 i piece pc
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                st_own[g] = wst + (uint32_t)lane * 64u + (((uint32_t)g ^ (((uint32_t)lane >> 1) & 3u)) << 4);
                const uint32_t r = ((uint32_t)lane >> 2) + 8u * g;
                st_map[g] = wst + r * 64u + ((pc ^ ((r >> 1) & 3u)) << 4);
            }
            const size_t offq = (((size_t)b * p.H + y0) * kRwW + q * 32) * kRwN + hsel * 64;      // pixel q*32 of row y0, this warp's column half
            uint4 rcur[4], rnext[4];
            auto fetch_res = [&](size_t off, uint4* r) {      // off: element offset of pixel q*32 for the wanted (row, chunk)
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    r[i] = p.residual ? __ldg(reinterpret_cast<const uint4*>(p.residual + off + (size_t)(((uint32_t)lane >> 2) + 8u * i) * kRwN + pc * 8))
                                      : make_uint4(0, 0, 0, 0);
            };
            fetch_res(offq, rcur);
            mbar_wait(&tfull[acc], acc_phase);
            tc_fence_after();
#pragma unroll
            for (int it = 0; it < 4; ++it) {                  // (row a, 32-column chunk ci)
                const uint32_t a = it >> 1, ci = it & 1;
                const uint32_t c0 = hsel * 64 + ci * 32;
                const size_t off = offq + (size_t)a * kRwW * kRwN + ci * 32;
                uint32_t v[32];
                tmem_ld32(tmem_base + ((q * 32u) << 16) + acc * 256 + a * kRwN + c0, v);
                if (it + 1 < 4) fetch_res(offq + (size_t)((it + 1) >> 1) * kRwW * kRwN + ((it + 1) & 1) * 32, rnext);
                uint4 rrow[4];
                if (p.residual) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) sts128(st_map[i], rcur[i]);
                    __syncwarp();
#pragma unroll
                    for (int g = 0; g < 4; ++g) rrow[g] = lds128(st_own[g]);
                    __syncwarp();
                }
                tmem_ld_wait();
                float f[32];
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    const float4 bv = *reinterpret_cast<const float4*>(sbias + c0 + 4 * g);
                    f[4 * g] = __uint_as_float(v[4 * g]) + bv.x; f[4 * g + 1] = __uint_as_float(v[4 * g + 1]) + bv.y;
                    f[4 * g + 2] = __uint_as_float(v[4 * g + 2]) + bv.z; f[4 * g + 3] = __uint_as_float(v[4 * g + 3]) + bv.w;
                }
                if (p.residual) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const __half2* h2 = reinterpret_cast<const __half2*>(&rrow[g]);
#pragma unroll
                        for (int i = 0; i < 4; ++i) { const float2 t = __half22float2(h2[i]); f[8 * g + 2 * i] += t.x; f[8 * g + 2 * i + 1] += t.y; }
                    }
                }
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    uint4 o;
                    __half2* h2 = reinterpret_cast<__half2*>(&o);
#pragma unroll
                    for (int i = 0; i < 4; ++i) h2[i] = __floats2half2_rn(f[8 * g + 2 * i], f[8 * g + 2 * i + 1]);
                    sts128(st_own[g], o);
                }
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint32_t r = ((uint32_t)lane >> 2) + 8u * i;
                    *reinterpret_cast<uint4*>(p.out + off + (size_t)r * kRwN + pc * 8) = lds128(st_map[i]);
                }
                __syncwarp();
                if (p.qstats) {   // fused GroupNorm quad statistics (same reduce-scatter as gemm_tc.cu)
                    float sv[16];
#pragma unroll
                    for (int q4 = 0; q4 < 8; ++q4) {
                        float su = 0.0f, sq = 0.0f;
#pragma unroll
                        for (int e = 0; e < 4; ++e) { const float xv = f[4 * q4 + e]; su += xv; sq = fmaf(xv, xv, sq); }
                        sv[q4] = su; sv[8 + q4] = sq;
                    }
#pragma unroll
                    for (int m = 16, half = 8; m >= 2; m >>= 1, half >>= 1) {
                        const bool upper = (lane & m) != 0;
#pragma unroll
                        for (int i = 0; i < half; ++i) {
                            const float send = upper ? sv[i] : sv[i + half];
                            const float recv = __shfl_xor_sync(0xffffffffu, send, m);
                            sv[i] = (upper ? sv[i + half] : sv[i]) + recv;
                        }
                    }
                    sv[0] += __shfl_xor_sync(0xffffffffu, sv[0], 1);
                    if ((lane & 1) == 0) {
                        const int idx = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
                        atomicAdd(qacc + (c0 / 4 + (idx & 7)) * 2 + (idx >> 3), sv[0]);
                    }
                }
                if (it + 1 < 4) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) rcur[g] = rnext[g];
                }
            }
            tc_fence_before();
            mbar_arrive(&tempty[acc]);
            if (p.qstats) {
                asm volatile("bar.sync 1, 256;" ::: "memory");
                const uint32_t et = threadIdx.x - 64;
                if (et < 64) {
                    const float val = qacc[et];
                    if (val != 0.0f) atomicAdd(p.qstats + (size_t)b * 64 + et, val);
                    qacc[et] = 0.0f;
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// tensor-map helper shared with gemm_tc.cu
int make_map_4d_box(CUtensorMap* m, const void* base, uint64_t K, uint64_t e1, uint64_t e2, uint64_t e3, uint64_t s1, uint64_t s2, uint64_t s3,
                    uint32_t x1, uint32_t x2, uint32_t x3);

// a: validated by ssdnerf_gemm_f16 (taps == 9, d1 == 128, b1 == 128, n == 128, fp16 output, dense NHWC output strides)
int conv_row2_launch(const ssdnerf_gemm_args* a, int sms, cudaStream_t stream) {
    ConvRowParams p{};
    p.B = a->d3; p.H = a->d2; p.kc1 = a->k1 / 64; p.kc2 = a->a2 ? a->k2 / 64 : 0;
    p.bias = a->bias_n; p.residual = (const __half*)a->residual; p.out = (__half*)a->out; p.qstats = a->qstats;
    p.prof = (unsigned long long*)a->debug_cycles;
    const uint64_t ktot = (uint64_t)a->k1 + (a->a2 ? a->k2 : 0);
    CUtensorMap mA1, mA2, mB;
    if (int e = make_map_4d_box(&mA1, a->a1, a->k1, a->d1, a->d2, a->d3, a->a1_strides[0], a->a1_strides[1], a->a1_strides[2], 130, 2, 1)) return e;
    if (a->a2) {
        if (int e = make_map_4d_box(&mA2, a->a2, a->k2, a->d1, a->d2, a->d3, a->a2_strides[0], a->a2_strides[1], a->a2_strides[2], 130, 2, 1)) return e;
    } else {
        mA2 = mA1;
    }
    if (int e = make_map_4d_box(&mB, a->b, ktot, a->n_rows_b ? a->n_rows_b : a->n, a->bx2 ? a->bx2 : 1, a->bx3 ? a->bx3 : 1, a->b_strides[0],
                                a->b_strides[1], a->b_strides[2], kRwN, 1, 1)) return e;
    static bool attr = false;
    if (!attr) {
        SSDNERF_CUDA_OK(cudaFuncSetAttribute(k_conv_row2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRwSmem));
        attr = true;
    }
    const uint32_t total = p.B * (p.H / 2);
    SSDNERF_CUDA_OK(launch_pdl(k_conv_row2, dim3(total < (uint32_t)sms ? total : (uint32_t)sms), dim3(kRwThreads), kRwSmem, stream, mA1, mA2, mB, p));
    SSDNERF_LAUNCH_OK();
    return 0;
}

}  // namespace ssdnerf
