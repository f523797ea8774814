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
#include "conv_row_epilogue.cuh"
#include "../../include/ssdnerf_b200.h"
#include <cuda_fp16.h>
#include <cstdlib>

namespace ssdnerf {
using namespace tc;

constexpr int kRwThreads = 320, kRwEpi = 256;
constexpr int kRwABox = 2 * 130 * 128;                  // bytes of one activation box: 2 rows x 130 pixels x 64 halves
constexpr int kRwASlot = 33 * 1024;                     // slot stride (1024-aligned)
constexpr int kRwBSlot = kRwN * 128;                    // 16 KB weight tile
constexpr int kRwAStages = 2, kRwBStages = 7;          // pair mode: 3 activation slots, 8 half-size weight stages
constexpr size_t kRwSmemTail = 1024 /*align*/ + 256 /*barriers*/ + 512 /*qacc*/ + 512 /*bias*/ + 8 * 2048 /*epilogue staging*/;
constexpr size_t kRwSmem = (size_t)kRwAStages * kRwASlot + (size_t)kRwBStages * kRwBSlot + kRwSmemTail;           // single CTA
constexpr size_t kRwSmemPair = (size_t)3 * kRwASlot + (size_t)8 * (kRwBSlot / 2) + kRwSmemTail;                    // per CTA of a pair

struct ConvRowParams {
    uint32_t B, H;                // images, rows (H even)
    uint32_t kc1, kc2;            // 64-channel chunks from input 1 / input 2 (skip concat)
    const float* bias;            // [128] or NULL
    const __half* residual;       // NHWC [B][H][128][128] or NULL
    __half* out;                  // NHWC [B][H][128][128]
    float* qstats;                // optional [B][32][2]
    unsigned long long* prof;
};

// PAIR: two CTAs (a cluster of 2 = the SMs of one TPC) work on two consecutive tiles with ONE tcgen05.mma.cta_group::2 of M = 256 per
// (tap, accumulator): the leader's MMA reads each CTA's own 128 activation rows and half of the weight tile from each CTA, so the number
// of issued MMA instructions -- what paces the single-CTA kernel at ~100 cycles per instruction against 64 cycles of execution
// (profiles/r01_gemm_pipeline_prof.txt) -- and the staged weight bytes per SM both halve.
template <bool PAIR>
__global__ void __launch_bounds__(kRwThreads, 1)
k_conv_row2(const __grid_constant__ CUtensorMap mapA1, const __grid_constant__ CUtensorMap mapA2, const __grid_constant__ CUtensorMap mapB,
            const ConvRowParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    constexpr int kAStages = PAIR ? 3 : kRwAStages, kBStages = PAIR ? 8 : kRwBStages, kBSlot = PAIR ? kRwBSlot / 2 : kRwBSlot;
    uint8_t* sA = smem;
    uint8_t* sB = smem + kAStages * kRwASlot;
    uint64_t* fullA = reinterpret_cast<uint64_t*>(sB + kBStages * kBSlot);
    uint64_t* emptyA = fullA + kAStages;
    uint64_t* fullB = emptyA + kAStages;
    uint64_t* emptyB = fullB + kBStages;
    uint64_t* tfull = emptyB + kBStages;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
    float* qacc = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(fullA) + 256);      // [32 quads][2]
    float* sbias = qacc + 128;                                                             // [128]
    uint8_t* sstage = reinterpret_cast<uint8_t*>(sbias + 128);                              // [8 warps][32 rows][64 B]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t tiles_per_img = p.H / 2, total_tiles = p.B * tiles_per_img;
    const uint32_t KC = p.kc1 + p.kc2;
    const uint32_t crank = PAIR ? cluster_ctarank() : 0u;
    // this CTA's tiles: single CTA: blockIdx.x, + gridDim.x, ...; pair: cluster c owns tiles 2c + rank, stride 2 x #clusters
    const uint32_t tile0 = PAIR ? cluster_id_x() * 2u + crank : blockIdx.x;
    const uint32_t tstep = PAIR ? num_clusters_x() * 2u : gridDim.x;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&mapA1); prefetch_tmap(&mapA2); prefetch_tmap(&mapB);
        for (int i = 0; i < kAStages; ++i) { mbar_init(&fullA[i], 1); mbar_init(&emptyA[i], 1); }
        for (int i = 0; i < kBStages; ++i) { mbar_init(&fullB[i], 1); mbar_init(&emptyB[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], kRwEpi * (PAIR ? 2 : 1)); }   // pair: both epilogues free the leader's
        fence_mbar_init();
    }
    if (warp == 1) { if (PAIR) tmem_alloc2(tmem_slot, 512); else tmem_alloc(tmem_slot, 512); }
    for (int i = threadIdx.x; i < 64; i += kRwThreads) qacc[i] = 0.0f;
    for (int i = threadIdx.x; i < kRwN; i += kRwThreads) sbias[i] = p.bias ? __ldg(p.bias + i) : 0.0f;
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();        // the peer's barriers exist before any remote arrive / TMA completion targets them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();
    pdl_wait();

    if (warp == 0) {   // ---------------- TMA producer (pair: own activation box + own half of the weight tile, bytes credited to the leader)
        uint32_t sa = 0, pa = 0, sb = 0, pb = 0;
        for (uint32_t tile = tile0; tile < total_tiles; tile += tstep) {
            const uint32_t b = tile / tiles_per_img, y0 = (tile - b * tiles_per_img) * 2;
            for (uint32_t ky = 0; ky < 3; ++ky) {
                for (uint32_t j = 0; j < KC; ++j) {
                    mbar_wait(&emptyA[sa], pa ^ 1);
                    const bool first = j < p.kc1;
                    const int ak = (int)((first ? j : j - p.kc1) * 64);
                    if (!PAIR) {
                        mbar_expect_tx_w(&fullA[sa], kRwABox);
                        tma_load_4d_w(sA + sa * kRwASlot, first ? &mapA1 : &mapA2, &fullA[sa], ak, -1, (int)(y0 + ky) - 1, (int)b);
                    } else {
                        if (crank == 0) mbar_expect_tx_w(&fullA[sa], 2 * kRwABox);
                        tma_load_4d_2cta_w(sA + sa * kRwASlot, first ? &mapA1 : &mapA2, mapa_u32(smem_u32(&fullA[sa]), 0), ak, -1, (int)(y0 + ky) - 1, (int)b);
                    }
                    if (++sa == kAStages) { sa = 0; pa ^= 1; }
                    for (uint32_t kx = 0; kx < 3; ++kx) {
                        mbar_wait(&emptyB[sb], pb ^ 1);
                        if (!PAIR) {
                            mbar_expect_tx_w(&fullB[sb], kRwBSlot);
                            tma_load_4d_w(sB + sb * kBSlot, &mapB, &fullB[sb], (int)(j * 64), 0, (int)(ky * 3 + kx), 0);
                        } else {
                            if (crank == 0) mbar_expect_tx_w(&fullB[sb], kRwBSlot);
                            tma_load_4d_2cta_w(sB + sb * kBSlot, &mapB, mapa_u32(smem_u32(&fullB[sb]), 0), (int)(j * 64), (int)(crank * 64), (int)(ky * 3 + kx), 0);
                        }
                        if (++sb == kBStages) { sb = 0; pb ^= 1; }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {   // ---------------- MMA issuer (pair: the leader issues for both SMs)
      if (!PAIR || crank == 0) {
        constexpr uint32_t idesc = make_idesc_f16(PAIR ? 256 : 128, kRwN);
        constexpr uint16_t kBoth = 3;
        uint32_t sa = 0, pa = 0, sb = 0, pb = 0, acc = 0, acc_phase = 0;
        long long wf = 0; const long long mt0 = clock64();
        for (uint32_t tile = tile0; tile < total_tiles; tile += tstep) {
            if (PAIR) mbar_wait_cluster(&tempty[acc], acc_phase ^ 1); else mbar_wait(&tempty[acc], acc_phase ^ 1);
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
                        const uint64_t b_desc = make_desc_sw128(smem_u32(sB + sb * kBSlot));
#pragma unroll
                        for (uint32_t a = 0; a < 2; ++a) {
                            const uint32_t row0 = a * 130 + kx;
                            const uint64_t a_desc = make_desc_sw128(a_base + row0 * 128);
#pragma unroll
                            for (uint32_t k = 0; k < 4; ++k) {
                                if (PAIR) umma_f16_2cta_w(d_tmem + a * kRwN, a_desc + 2 * k, b_desc + 2 * k, idesc, (started | k) != 0);
                                else umma_f16_w(d_tmem + a * kRwN, a_desc + 2 * k, b_desc + 2 * k, idesc, (started | k) != 0);
                            }
                        }
                        started = 1;
                        if (PAIR) umma_commit_2cta_w(&emptyB[sb], kBoth); else umma_commit_w(&emptyB[sb]);
                        if (++sb == kBStages) { sb = 0; pb ^= 1; }
                    }
                    if (PAIR) umma_commit_2cta_w(&emptyA[sa], kBoth); else umma_commit_w(&emptyA[sa]);
                    if (++sa == kAStages) { sa = 0; pa ^= 1; }
                }
            }
            if (PAIR) umma_commit_2cta_w(&tfull[acc], kBoth); else umma_commit_w(&tfull[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (p.prof && lane == 0) { atomicAdd(p.prof + 2, (unsigned long long)wf); atomicAdd(p.prof + 4, (unsigned long long)(clock64() - mt0)); }
      }
        __syncwarp();
    } else {   // ---------------- epilogue warps 2..9 (conv_row_epilogue.cuh)
        const RowEpiArgs ea{p.H, p.residual, p.out, p.qstats};
        conv_row_epilogue<8, PAIR>(ea, warp, lane, total_tiles, tiles_per_img, tmem_base, tfull, tempty, sstage, sbias, qacc, tile0, tstep);
    }
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();        // nobody exits while the leader's MMAs may still read its shared memory
    if (warp == 1) { tc_fence_after(); if (PAIR) tmem_dealloc2(tmem_base, 512); else tmem_dealloc(tmem_base, 512); }
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
    const uint32_t total = p.B * (p.H / 2);
    // CTA pairs need an even tile count per image (pairs never straddle images); SSDNERF_ROW2_PAIR=0 forces the single-CTA kernel
    static int pair_env = -1;
    if (pair_env < 0) { const char* e = getenv("SSDNERF_ROW2_PAIR"); pair_env = e ? atoi(e) : 1; }
    const bool pair = pair_env && (p.H / 2) % 2 == 0 && total >= 2;
    if (int e = make_map_4d_box(&mB, a->b, ktot, a->n_rows_b ? a->n_rows_b : a->n, a->bx2 ? a->bx2 : 1, a->bx3 ? a->bx3 : 1, a->b_strides[0],
                                a->b_strides[1], a->b_strides[2], pair ? kRwN / 2 : kRwN, 1, 1)) return e;
    if (pair) {
        static DeviceOnce attr;
        if (attr.first()) {
            SSDNERF_CUDA_OK(cudaFuncSetAttribute(k_conv_row2<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRwSmemPair));
        }
        const uint32_t clusters = (total / 2 < (uint32_t)sms / 2) ? total / 2 : (uint32_t)sms / 2;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(clusters * 2); cfg.blockDim = dim3(kRwThreads); cfg.dynamicSmemBytes = kRwSmemPair; cfg.stream = stream;
        cudaLaunchAttribute at[2];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
        cfg.attrs = at; cfg.numAttrs = 2;
        SSDNERF_CUDA_OK(cudaLaunchKernelEx(&cfg, k_conv_row2<true>, mA1, mA2, mB, p));
    } else {
        static DeviceOnce attr;
        if (attr.first()) {
            SSDNERF_CUDA_OK(cudaFuncSetAttribute(k_conv_row2<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRwSmem));
        }
        SSDNERF_CUDA_OK(launch_pdl(k_conv_row2<false>, dim3(total < (uint32_t)sms ? total : (uint32_t)sms), dim3(kRwThreads), kRwSmem, stream, mA1, mA2, mB, p));
    }
    SSDNERF_LAUNCH_OK();
    return 0;
}

}  // namespace ssdnerf
