// Fused inference renderer, variant P (shipped configs: 3x6 channels, hidden 64, dir_net) on the tensor cores.
//
// Same structure as render_tc.cu (one CTA = 128 rays = 128 TMEM lanes) but sized for the tiny decoder and kept at
// fp32-class accuracy: the 18 -> 64 base layer runs as a SPLIT-PRECISION fp16 tcgen05 GEMM
//     base_x = F_hi W_hi^T + F_hi W_lo^T + F_lo W_hi^T          (x = hi + lo, hi = fp16(x), lo = fp16(x - hi))
// (K = 32: 3 planes x 8 padded channels + a constant-one column that carries the bias; 6 MMAs of 128x64x16 per
// iteration), which removes the 1152 FMAs + 288 shared-memory weight loads per sample of the CUDA-core version
// (render_fused.cu, kept as SSDNERF_DEC_P_SIMT for A/B).  The bilinear gather is cooperative: four lanes fetch the
// 64 contiguous bytes of one (sample, plane, row) texel pair, so a warp-level 16-byte load touches ~8 lines, not 32.
// Heads (density 64->1, colour 64->3 with the per-ray dir_net(SH16) vector), exp / sigmoid and the compositor stay in
// the registers of the thread that owns the ray.
#include "common.cuh"
#include "render_common.cuh"
#include "dec_p.cuh"
#include "tc_common.cuh"
#include "../../include/ssdnerf_b200.h"

namespace ssdnerf {
using namespace tc;

constexpr int kPThreads = 128;
constexpr uint32_t kPA_LBO = 128 * 16, kPA_SBO = 128, kPA_BYTES = 4 * kPA_LBO;     // A: 128 rows x 32 K fp16 (4 K-chunks)
constexpr uint32_t kPW_LBO = 64 * 16, kPW_BYTES = 4 * kPW_LBO;                      // W: 64 rows x 32 K fp16
constexpr uint32_t kPTmemCols = 64;
constexpr int kOneCol = 24;        // K index of the constant-one column (bias row of W)

struct SmemPT {
    alignas(128) uint8_t a_hi[kPA_BYTES];
    alignas(128) uint8_t a_lo[kPA_BYTES];
    alignas(128) uint8_t w_hi[kPW_BYTES];
    alignas(128) uint8_t w_lo[kPW_BYTES];
    float4 wdir[16][DecP::HID / 4];
    float bdir[DecP::HID];
    float4 heads[DecP::HID];                 // {wd, wc0, wc1, wc2}[o]
    float dirf[DecP::HID][kPThreads];        // per-ray dir_net(SH16(d)); column = thread
    float bd, bc[3], sat;
    alignas(8) uint64_t mma_bar;
    uint32_t tmem_slot;
    uint32_t tile;
};

__device__ __forceinline__ void split_h(float x, __half& hi, __half& lo) {
    hi = __float2half_rn(x);
    lo = __float2half_rn(x - __half2float(hi));
}

__global__ void __launch_bounds__(kPThreads, 2) k_render_ptc(RenderParams p, int mode) {
    extern __shared__ uint8_t smem_raw[];
    SmemPT& s = *reinterpret_cast<SmemPT*>(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    {   // ---- one-time set-up
        const float* blob = p.blob;
        __half* whi = reinterpret_cast<__half*>(s.w_hi);
        __half* wlo = reinterpret_cast<__half*>(s.w_lo);
        for (int i = tid; i < DecP::HID * 32; i += kPThreads) {
            const int n = i >> 5, k = i & 31;                      // W[n][k], k = plane*8 + c | 24 = bias | rest 0
            float v = 0.0f;
            const int pl = k >> 3, c = k & 7;
            if (pl < 3 && c < DecP::C) v = __ldg(blob + DecP::OFF_W1 + (pl * DecP::C + c) * DecP::HID + n);
            else if (k == kOneCol) v = __ldg(blob + DecP::OFF_B1 + n);
            __half hi, lo;
            split_h(v, hi, lo);
            const int off = ((k >> 3) * kPW_LBO + n * 16) / 2 + (k & 7);
            whi[off] = hi; wlo[off] = lo;
        }
        float* wdir = reinterpret_cast<float*>(s.wdir);
        for (int i = tid; i < 16 * DecP::HID; i += kPThreads) wdir[i] = __ldg(blob + DecP::OFF_WDIR + i);
        for (int i = tid; i < DecP::HID; i += kPThreads) {
            s.bdir[i] = __ldg(blob + DecP::OFF_BDIR + i);
            s.heads[i] = make_float4(__ldg(blob + DecP::OFF_WD + i), __ldg(blob + DecP::OFF_WC + i), __ldg(blob + DecP::OFF_WC + DecP::HID + i),
                                     __ldg(blob + DecP::OFF_WC + 2 * DecP::HID + i));
        }
        if (tid == 0) {
            s.bd = __ldg(blob + DecP::OFF_BD);
            s.bc[0] = __ldg(blob + DecP::OFF_BC); s.bc[1] = __ldg(blob + DecP::OFF_BC + 1); s.bc[2] = __ldg(blob + DecP::OFF_BC + 2);
            s.sat = __ldg(blob + DecP::OFF_SAT);
            mbar_init(&s.mma_bar, 1);
            fence_mbar_init();
        }
        // A tiles: zero everything, then the constant-one column (hi = 1, lo = 0) of every row
        for (int i = tid; i < (int)(kPA_BYTES / 16); i += kPThreads) {
            reinterpret_cast<uint4*>(s.a_hi)[i] = make_uint4(0, 0, 0, 0);
            reinterpret_cast<uint4*>(s.a_lo)[i] = make_uint4(0, 0, 0, 0);
        }
        __syncthreads();
        reinterpret_cast<__half*>(s.a_hi)[((kOneCol >> 3) * kPA_LBO + (tid >> 3) * kPA_SBO + (tid & 7) * 16) / 2 + (kOneCol & 7)] = __float2half(1.0f);
        if (warp == 0) tmem_alloc(&s.tmem_slot, kPTmemCols);
        fence_proxy_async_smem();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }
    const uint32_t tmem = s.tmem_slot;
    const uint32_t lane_base = ((uint32_t)warp * 32u) << 16;
    const uint32_t ahi = smem_u32(s.a_hi), alo = smem_u32(s.a_lo), whi_a = smem_u32(s.w_hi), wlo_a = smem_u32(s.w_lo);
    constexpr uint32_t idesc = make_idesc_f16(128, DecP::HID);
    uint32_t bar_phase = 0;
    unsigned long long pc[6] = {0, 0, 0, 0, 0, 0};
    const bool prof = p.prof != nullptr && tid == 0;
#define PTC_T(i) do { if (prof) { const long long _n = clock64(); pc[i] += (unsigned long long)(_n - t_prev); t_prev = _n; } } while (0)
    long long t_prev = clock64();

    const uint32_t warp_tiles_per_scene = div_up(p.rays_per_scene, 32u);
    const uint32_t cta_tiles_per_scene = div_up(warp_tiles_per_scene, 4u);
    const uint32_t total_tiles = cta_tiles_per_scene * p.num_scenes;
    uint32_t* tile_counter = p.counters + mode;
    const int sub = lane & 3, quad = lane >> 2;
    const int xs = sub >> 1, hh = sub & 1;            // which texel of the pair (x0 / x1), which channel half (0-3 / 4-7)

    for (;;) {
        if (tid == 0) s.tile = atomicAdd(tile_counter, 1u);
        __syncthreads();
        const uint32_t tile = s.tile;
        __syncthreads();
        if (tile >= total_tiles) break;
        const uint32_t scene = tile / cta_tiles_per_scene;
        const uint32_t wtile = (tile - scene * cta_tiles_per_scene) * 4u + (uint32_t)warp;
        const uint32_t n = (wtile < warp_tiles_per_scene) ? ray_in_tile(p, wtile, lane) : 0xffffffffu;
        const bool valid = n < p.rays_per_scene;
        const size_t gidx = (size_t)scene * p.rays_per_scene + (valid ? n : 0);

        uint32_t cap = p.hard_cap;
        bool active = valid;
        if (mode == 1) {
            cap = p.budget[scene];
            active = valid && (uint32_t)p.count_buf[gidx] > cap;
            if (!__syncthreads_or(active)) continue;
        }

        Ray r;
        make_ray(p, scene, valid ? n : 0, r);
        float near, far;
        near_far_aabb(r, p.aabb, p.min_near, near, far);
        MarchCfg c = p.cfg;
        if (p.dt_gamma) c.dt_gamma = __ldg(p.dt_gamma + scene);

        {   // per-ray view-direction features: dirf = Wdir^T SH16(d) + bdir
            float sh[16];
            sh16(r.dx, r.dy, r.dz, sh);
#pragma unroll 4
            for (int o4 = 0; o4 < DecP::HID / 4; ++o4) {
                float a0 = s.bdir[4 * o4], a1 = s.bdir[4 * o4 + 1], a2 = s.bdir[4 * o4 + 2], a3 = s.bdir[4 * o4 + 3];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float4 w = s.wdir[j][o4];
                    a0 = fmaf(sh[j], w.x, a0); a1 = fmaf(sh[j], w.y, a1); a2 = fmaf(sh[j], w.z, a2); a3 = fmaf(sh[j], w.w, a3);
                }
                s.dirf[4 * o4][tid] = a0; s.dirf[4 * o4 + 1][tid] = a1; s.dirf[4 * o4 + 2][tid] = a2; s.dirf[4 * o4 + 3][tid] = a3;
            }
        }

        const float* planes = reinterpret_cast<const float*>(p.planes) + (size_t)scene * 3 * p.plane_h * p.plane_w * DecP::CPAD;
        const size_t plane_stride = (size_t)p.plane_h * p.plane_w * DecP::CPAD;
        BitfieldLoader grid{p.bitfield + (size_t)scene * (p.cfg.H * p.cfg.H * p.cfg.H / 8) * p.cfg.C};
        int32_t* trace = p.voxel_trace ? p.voxel_trace + gidx * p.trace_cap : nullptr;
        const float Wf = (float)p.plane_w, Hf = (float)p.plane_h;
        const int Wm1 = (int)p.plane_w - 1, Hm1 = (int)p.plane_h - 1;

        float t = near;
        float ws = 0.0f, dep = 0.0f, cr = 0.0f, cg = 0.0f, cb = 0.0f;
        uint32_t ns = 0;
        bool alive = active, tbreak = false;
        for (;;) {
            // ---- phase 1: next occupied sample of this thread's ray
            bool has = false;
            float x = 0.0f, y = 0.0f, z = 0.0f, dt = 0.0f; uint32_t vi = 0;
            while (alive && !has) {
                if (!(t < far) || ns >= cap) { alive = false; break; }
                has = probe(c, r, grid, t, x, y, z, dt, vi);
            }
            if (!__syncthreads_or(has)) break;
            PTC_T(0);   // probe + vote

            // ---- phase 2: cooperative gather -> split fp16 feature rows in shared memory.
            // All 24 texel loads of the iteration are issued before any of them is consumed (one L2 round trip per iteration
            // instead of twelve dependent ones): 2a fetch the sample positions, 2b issue loads, 2c blend / exchange / store.
            const uint32_t has_mask = __ballot_sync(0xffffffffu, has);
            float4 la[4][3], lc[4][3];
            float w0s[4][3], w1s[4][3];
#pragma unroll
            for (int round = 0; round < 4; ++round) {
                const int src = round * 8 + quad;
                const float sx = __shfl_sync(0xffffffffu, x, src);
                const float sy = __shfl_sync(0xffffffffu, y, src);
                const float sz = __shfl_sync(0xffffffffu, z, src);
                const bool on = (has_mask >> src) & 1u;
#pragma unroll
                for (int pl = 0; pl < 3; ++pl) {
                    const float u = (pl == 2) ? sy : sx, v = (pl == 0) ? sy : sz;     // planes (x,y) (x,z) (y,z)
                    float ix = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(u, 1.0f), Wf), 1.0f), 0.5f);
                    float iy = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(v, 1.0f), Hf), 1.0f), 0.5f);
                    ix = fminf((float)Wm1, fmaxf(ix, 0.0f));
                    iy = fminf((float)Hm1, fmaxf(iy, 0.0f));
                    const float fx0 = floorf(ix), fy0 = floorf(iy);
                    const int x0 = (int)fx0, y0 = (int)fy0;
                    const int xsel = xs ? min(x0 + 1, Wm1) : x0, y1 = min(y0 + 1, Hm1);
                    const float wx = xs ? (ix - fx0) : ((fx0 + 1.0f) - ix);
                    w0s[round][pl] = wx * ((fy0 + 1.0f) - iy);
                    w1s[round][pl] = wx * (iy - fy0);
                    la[round][pl] = make_float4(0.f, 0.f, 0.f, 0.f);
                    lc[round][pl] = la[round][pl];
                    if (on) {
                        const float* base = planes + pl * plane_stride + hh * 4;
                        la[round][pl] = __ldg(reinterpret_cast<const float4*>(base + ((size_t)y0 * p.plane_w + xsel) * DecP::CPAD));
                        lc[round][pl] = __ldg(reinterpret_cast<const float4*>(base + ((size_t)y1 * p.plane_w + xsel) * DecP::CPAD));
                    }
                }
            }
#pragma unroll
            for (int round = 0; round < 4; ++round) {
                const int src = round * 8 + quad;
                const bool on = (has_mask >> src) & 1u;
                const uint32_t row = (uint32_t)warp * 32u + (uint32_t)src;
                const uint32_t row_off = (row >> 3) * kPA_SBO + (row & 7) * 16 + hh * 8;
#pragma unroll
                for (int pl = 0; pl < 3; ++pl) {
                    const float4 a = la[round][pl], cc = lc[round][pl];
                    const float w0 = w0s[round][pl], w1 = w1s[round][pl];
                    float f0 = a.x * w0 + cc.x * w1, f1 = a.y * w0 + cc.y * w1, f2 = a.z * w0 + cc.z * w1, f3 = a.w * w0 + cc.w * w1;
                    f0 += __shfl_xor_sync(0xffffffffu, f0, 2);
                    f1 += __shfl_xor_sync(0xffffffffu, f1, 2);
                    f2 += __shfl_xor_sync(0xffffffffu, f2, 2);
                    f3 += __shfl_xor_sync(0xffffffffu, f3, 2);
                    if (on && xs == 0) {
                        __half h0, l0, h1, l1, h2, l2, h3, l3;
                        split_h(f0, h0, l0); split_h(f1, h1, l1); split_h(f2, h2, l2); split_h(f3, h3, l3);
                        uint2 vh, vl;
                        vh.x = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
                        vh.y = (uint32_t)__half_as_ushort(h2) | ((uint32_t)__half_as_ushort(h3) << 16);
                        vl.x = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
                        vl.y = (uint32_t)__half_as_ushort(l2) | ((uint32_t)__half_as_ushort(l3) << 16);
                        *reinterpret_cast<uint2*>(s.a_hi + pl * kPA_LBO + row_off) = vh;
                        *reinterpret_cast<uint2*>(s.a_lo + pl * kPA_LBO + row_off) = vl;
                    }
                }
            }
            PTC_T(1);   // gather + stores
            fence_proxy_async_smem();
            tc_fence_before();
            __syncthreads();
            PTC_T(2);   // fence + barrier

            // ---- base layer on the tensor cores (3 split-precision products, K = 32 each)
            if (tid == 0) {
                tc_fence_after();
#pragma unroll
                for (uint32_t k = 0; k < 2; ++k)
                    umma_f16(tmem, make_desc_nosw(ahi + k * 2 * kPA_LBO, kPA_LBO, kPA_SBO), make_desc_nosw(whi_a + k * 2 * kPW_LBO, kPW_LBO, 128), idesc, k != 0);
#pragma unroll
                for (uint32_t k = 0; k < 2; ++k)
                    umma_f16(tmem, make_desc_nosw(ahi + k * 2 * kPA_LBO, kPA_LBO, kPA_SBO), make_desc_nosw(wlo_a + k * 2 * kPW_LBO, kPW_LBO, 128), idesc, 1);
#pragma unroll
                for (uint32_t k = 0; k < 2; ++k)
                    umma_f16(tmem, make_desc_nosw(alo + k * 2 * kPA_LBO, kPA_LBO, kPA_SBO), make_desc_nosw(whi_a + k * 2 * kPW_LBO, kPW_LBO, 128), idesc, 1);
                umma_commit(&s.mma_bar);
            }
            mbar_wait(&s.mma_bar, bar_phase); bar_phase ^= 1;
            tc_fence_after();
            PTC_T(3);   // MMA issue + completion wait

            // ---- heads in registers
            float sd = s.bd, o_r = s.bc[0], o_g = s.bc[1], o_b = s.bc[2];
#pragma unroll 1
            for (int c0 = 0; c0 < DecP::HID; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(tmem + lane_base + c0, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float bx = __uint_as_float(v[i]);
                    const float4 hw = s.heads[c0 + i];
                    sd = fmaf(silu_f(bx), hw.x, sd);
                    const float h = silu_f(bx + s.dirf[c0 + i][tid]);
                    o_r = fmaf(h, hw.y, o_r); o_g = fmaf(h, hw.z, o_g); o_b = fmaf(h, hw.w, o_b);
                }
            }
            tc_fence_before();
            PTC_T(4);   // TMEM loads + heads

            // ---- composite (raymarching.cu:865-897 arithmetic)
            if (has) {
                const float sigma = __expf(sd);
                const float k1 = 1.0f + 2.0f * s.sat;
                const float sr = sigmoid_f(o_r) * k1 - s.sat, sg = sigmoid_f(o_g) * k1 - s.sat, sb = sigmoid_f(o_b) * k1 - s.sat;
                const float alpha = 1.0f - __expf(-sigma * dt);
                const float T = 1.0f - ws;
                const float w = alpha * T;
                ws += w;
                dep = __fmaf_rn(w, t, dep);
                cr = __fmaf_rn(w, sr, cr); cg = __fmaf_rn(w, sg, cg); cb = __fmaf_rn(w, sb, cb);
                if (trace && ns < p.trace_cap) trace[ns] = (int32_t)vi;
                ++ns;
                if (T < p.T_thresh) { alive = false; tbreak = true; }
                else t = __fadd_rn(t, dt);
            }
            PTC_T(5);   // composite
        }
        if (active) {
            p.weights_sum[gidx] = ws;
            if (p.depth) p.depth[gidx] = dep;
            p.image[3 * gidx] = cr; p.image[3 * gidx + 1] = cg; p.image[3 * gidx + 2] = cb;
            if (p.rgb_blend) {
                const float k = p.bg_color * (1.0f - ws);
                p.rgb_blend[3 * gidx] = cr + k; p.rgb_blend[3 * gidx + 1] = cg + k; p.rgb_blend[3 * gidx + 2] = cb + k;
            }
            if (trace) for (uint32_t i = ns; i < p.trace_cap; ++i) trace[i] = -1;
            p.count_buf[gidx] = (int32_t)ns;
            if (mode == 0 && p.hist) {
                const uint32_t L = tbreak ? ns - 1 : ns;
                atomicAdd(p.hist + (size_t)scene * p.hist_bins + min(L, p.hist_bins - 1), 1u);
            }
        }
    }
    if (prof) for (int i = 0; i < 6; ++i) atomicAdd(p.prof + i, pc[i]);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, kPTmemCols); }
}

int render_ptc_launch(const RenderParams& p, int emulate_schedule, uint32_t* hist, int sms, cudaStream_t stream) {
    const size_t smem = sizeof(SmemPT) + 128;
    static DeviceOnce attr_set;
    if (attr_set.first()) {
        SSDNERF_CUDA_OK(cudaFuncSetAttribute(k_render_ptc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    int occ = 0;
    SSDNERF_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_render_ptc, kPThreads, smem));
    if (occ < 1) return set_error_msg(SSDNERF_ERR_CUDA, "render_fwd: variant P (tensor-core) kernel does not fit on this device");
    if (occ > 8) occ = 8;   // 64 TMEM columns per CTA
    const uint32_t total_tiles = div_up(div_up(p.rays_per_scene, 32u), 4u) * p.num_scenes;
    const uint32_t grid = (uint32_t)min((uint64_t)sms * occ, (uint64_t)total_tiles);
    k_render_ptc<<<grid, kPThreads, smem, stream>>>(p, 0);
    SSDNERF_LAUNCH_OK();
    if (emulate_schedule) {
        if (int e = launch_schedule(hist, p.hist_bins, p.num_scenes, p.rays_per_scene, p.max_steps, p.budget, stream)) return e;
        k_render_ptc<<<grid, kPThreads, smem, stream>>>(p, 1);
        SSDNERF_LAUNCH_OK();
    }
    return 0;
}

}  // namespace ssdnerf
