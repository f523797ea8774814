// Fused inference renderer, variant S (TriPlaneDecoder class defaults: 3x32 channels, hidden 128,
// colour net (128 + SH16) -> 128 -> 3; lib/models/decoders/triplane_decoder.py:24-39,119-179).
//
// One CTA = 128 rays = 128 TMEM lanes.  Every iteration each ray advances to its next occupied sample
// (bit-exact stepping, common.cuh), the CTA gathers 128 x 96 bilinear features (fp16 channels-last planes,
// four lanes cooperate on one sample so a warp-level 16-byte load touches 8 lines instead of 32), and the
// two hidden layers run on the tensor cores:
//     GEMM1  [128 x 96 ] x W1^T [96  x 128]            -> TMEM, + b1, SiLU, fp16 -> A operand of GEMM2
//     GEMM2  [128 x 144] x W2^T [144 x 144]            -> TMEM   (K = 128 base_act + 16 SH; N = 128 hidden + 1
//                                                         density pre-activation + 15 zero columns)
// issued by one thread with tcgen05.mma (fp16 x fp16 -> fp32), accumulators read back with tcgen05.ld; the
// 128->3 output layer, exp / sigmoid and the compositor stay in registers of the thread that owns the ray.
// Operands live in shared memory in the un-swizzled K-major canonical layout (8x16-byte core matrices).
#include "common.cuh"
#include "render_common.cuh"
#include "tc_common.cuh"
#include "../../include/ssdnerf_b200.h"

namespace ssdnerf {
using namespace tc;

struct DecS {
    static constexpr int C = 32, KF = 96, HID = 128, K2 = 144, N2 = 144;
    static constexpr int OFF_W1 = 0, OFF_B1 = OFF_W1 + HID * KF, OFF_WD = OFF_B1 + HID, OFF_BD = OFF_WD + HID,
                         OFF_WC0 = OFF_BD + 4, OFF_BC0 = OFF_WC0 + HID * K2, OFF_WC2 = OFF_BC0 + HID,
                         OFF_BC2 = OFF_WC2 + 3 * HID, OFF_SAT = OFF_BC2 + 4, BLOB = OFF_SAT + 4;
};
size_t dec_s_blob_floats() { return DecS::BLOB; }

constexpr int kSThreads = 128;
// A operand: 18 K-chunks (8 halves each); chunk stride padded by 32 B so the 4-lanes-per-sample stores are conflict-free
constexpr uint32_t kA_LBO = 2048 + 32, kA_SBO = 128, kA_BYTES = 18 * kA_LBO;
constexpr uint32_t kW1_LBO = DecS::HID * 16, kW1_BYTES = (DecS::KF / 8) * kW1_LBO;   // 12 x 2048
constexpr uint32_t kW2_LBO = DecS::N2 * 16, kW2_BYTES = (DecS::K2 / 8) * kW2_LBO;    // 18 x 2304
constexpr uint32_t kTmemCols = 256;

struct SmemS {
    alignas(128) uint8_t a[kA_BYTES];
    alignas(128) uint8_t w1[kW1_BYTES];
    alignas(128) uint8_t w2[kW2_BYTES];
    float b1[DecS::HID];
    float b2[DecS::HID];
    float wc2[3][DecS::HID];
    float bd, bc2[3], sat;
    alignas(8) uint64_t mma_bar;
    uint32_t tmem_slot;
    uint32_t tile;
};

__device__ __forceinline__ float tanh_approx(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float silu_fast(float x) {   // x * sigmoid(x) = 0.5x * (1 + tanh(0.5x))
    const float h = 0.5f * x;
    return fmaf(h, tanh_approx(h), h);
}

struct BitfieldLoaderS {
    const uint8_t* __restrict__ g;
    __device__ __forceinline__ uint32_t operator()(uint32_t byte) const { return __ldg(g + byte); }
};

// one plane of the cooperative gather: this lane owns channels [8*sub, 8*sub+8) of sample (u, v)
__device__ __forceinline__ uint4 gather_plane_s(const __half* __restrict__ plane, uint32_t Hp, uint32_t Wp, float u, float v, int sub) {
    float ix = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(u, 1.0f), (float)Wp), 1.0f), 0.5f);
    float iy = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(v, 1.0f), (float)Hp), 1.0f), 0.5f);
    ix = fminf((float)(Wp - 1), fmaxf(ix, 0.0f));
    iy = fminf((float)(Hp - 1), fmaxf(iy, 0.0f));
    const float fx0 = floorf(ix), fy0 = floorf(iy);
    const int x0 = (int)fx0, y0 = (int)fy0;
    const int x1 = min(x0 + 1, (int)Wp - 1), y1 = min(y0 + 1, (int)Hp - 1);
    const float wx1 = ix - fx0, wy1 = iy - fy0, wx0 = (fx0 + 1.0f) - ix, wy0 = (fy0 + 1.0f) - iy;
    const float nw = wx0 * wy0, ne = wx1 * wy0, sw = wx0 * wy1, se = wx1 * wy1;
    const uint4* base = reinterpret_cast<const uint4*>(plane) + sub;   // 4 x 16 B per texel
    const uint4 a = __ldg(base + ((size_t)y0 * Wp + x0) * 4);
    const uint4 b = __ldg(base + ((size_t)y0 * Wp + x1) * 4);
    const uint4 c = __ldg(base + ((size_t)y1 * Wp + x0) * 4);
    const uint4 d = __ldg(base + ((size_t)y1 * Wp + x1) * 4);
    const __half2* ha = reinterpret_cast<const __half2*>(&a);
    const __half2* hb = reinterpret_cast<const __half2*>(&b);
    const __half2* hc = reinterpret_cast<const __half2*>(&c);
    const __half2* hd = reinterpret_cast<const __half2*>(&d);
    uint4 o;
    __half2* ho = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 fa = __half22float2(ha[i]), fb = __half22float2(hb[i]), fc = __half22float2(hc[i]), fd = __half22float2(hd[i]);
        const float r0 = fa.x * nw + fb.x * ne + fc.x * sw + fd.x * se;
        const float r1 = fa.y * nw + fb.y * ne + fc.y * sw + fd.y * se;
        ho[i] = __floats2half2_rn(r0, r1);
    }
    return o;
}

__global__ void __launch_bounds__(kSThreads, 2) k_render_s(RenderParams p, int mode) {
    extern __shared__ uint8_t smem_raw[];
    SmemS& s = *reinterpret_cast<SmemS*>(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // ---- one-time set-up: weights -> fp16 UMMA layout, barrier, TMEM
    {
        const float* blob = p.blob;
        __half* w1 = reinterpret_cast<__half*>(s.w1);
        for (int i = tid; i < DecS::HID * DecS::KF; i += kSThreads) {
            const int n = i / DecS::KF, k = i - n * DecS::KF;
            w1[((k >> 3) * kW1_LBO + n * 16) / 2 + (k & 7)] = __float2half_rn(__ldg(blob + DecS::OFF_W1 + i));
        }
        __half* w2 = reinterpret_cast<__half*>(s.w2);
        for (int i = tid; i < DecS::N2 * DecS::K2; i += kSThreads) {
            const int n = i / DecS::K2, k = i - n * DecS::K2;
            float v = 0.0f;
            if (n < DecS::HID) v = __ldg(blob + DecS::OFF_WC0 + n * DecS::K2 + k);
            else if (n == DecS::HID && k < DecS::HID) v = __ldg(blob + DecS::OFF_WD + k);
            w2[((k >> 3) * kW2_LBO + n * 16) / 2 + (k & 7)] = __float2half_rn(v);
        }
        for (int i = tid; i < DecS::HID; i += kSThreads) {
            s.b1[i] = __ldg(blob + DecS::OFF_B1 + i);
            s.b2[i] = __ldg(blob + DecS::OFF_BC0 + i);
            s.wc2[0][i] = __ldg(blob + DecS::OFF_WC2 + i);
            s.wc2[1][i] = __ldg(blob + DecS::OFF_WC2 + DecS::HID + i);
            s.wc2[2][i] = __ldg(blob + DecS::OFF_WC2 + 2 * DecS::HID + i);
        }
        if (tid == 0) {
            s.bd = __ldg(blob + DecS::OFF_BD);
            s.bc2[0] = __ldg(blob + DecS::OFF_BC2); s.bc2[1] = __ldg(blob + DecS::OFF_BC2 + 1); s.bc2[2] = __ldg(blob + DecS::OFF_BC2 + 2);
            s.sat = __ldg(blob + DecS::OFF_SAT);
            mbar_init(&s.mma_bar, 1);
            fence_mbar_init();
        }
        // zero the A tile once so rows of rays without a sample never hold NaN bit patterns
        for (int i = tid; i < (int)(kA_BYTES / 16); i += kSThreads) reinterpret_cast<uint4*>(s.a)[i] = make_uint4(0, 0, 0, 0);
        if (warp == 0) tmem_alloc(&s.tmem_slot, kTmemCols);
        fence_proxy_async_smem();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }
    const uint32_t tmem = s.tmem_slot;
    const uint32_t lane_base = ((uint32_t)warp * 32u) << 16;
    const uint32_t a_addr = smem_u32(s.a), w1_addr = smem_u32(s.w1), w2_addr = smem_u32(s.w2);
    constexpr uint32_t idesc1 = make_idesc_f16(128, DecS::HID);
    constexpr uint32_t idesc2 = make_idesc_f16(128, DecS::N2);
    uint32_t bar_phase = 0;

    const uint32_t warp_tiles_per_scene = div_up(p.rays_per_scene, 32u);
    const uint32_t cta_tiles_per_scene = div_up(warp_tiles_per_scene, 4u);
    const uint32_t total_tiles = cta_tiles_per_scene * p.num_scenes;
    uint32_t* tile_counter = p.counters + mode;
    const int sub = lane & 3, quad = lane >> 2;

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

        {   // SH16 of the ray direction -> K chunks 16, 17 of this thread's A row (constant along the ray)
            float sh[16];
            sh16(r.dx, r.dy, r.dz, sh);
            uint4 o0, o1;
            __half2* h0 = reinterpret_cast<__half2*>(&o0);
            __half2* h1 = reinterpret_cast<__half2*>(&o1);
#pragma unroll
            for (int i = 0; i < 4; ++i) { h0[i] = __floats2half2_rn(sh[2 * i], sh[2 * i + 1]); h1[i] = __floats2half2_rn(sh[8 + 2 * i], sh[9 + 2 * i]); }
            *reinterpret_cast<uint4*>(s.a + 16 * kA_LBO + tid * 16) = o0;
            *reinterpret_cast<uint4*>(s.a + 17 * kA_LBO + tid * 16) = o1;
        }

        const __half* planes = reinterpret_cast<const __half*>(p.planes) + (size_t)scene * 3 * p.plane_h * p.plane_w * DecS::C;
        const size_t plane_stride = (size_t)p.plane_h * p.plane_w * DecS::C;
        BitfieldLoaderS grid{p.bitfield + (size_t)scene * (p.cfg.H * p.cfg.H * p.cfg.H / 8) * p.cfg.C};
        int32_t* trace = p.voxel_trace ? p.voxel_trace + gidx * p.trace_cap : nullptr;

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

            // ---- phase 2: cooperative gather, 4 lanes per sample, 8 samples per round
            const uint32_t has_mask = __ballot_sync(0xffffffffu, has);
#pragma unroll 1
            for (int round = 0; round < 4; ++round) {
                const int src = round * 8 + quad;
                const float sx = __shfl_sync(0xffffffffu, x, src);
                const float sy = __shfl_sync(0xffffffffu, y, src);
                const float sz = __shfl_sync(0xffffffffu, z, src);
                if (has_mask & (1u << src)) {
                    const uint32_t row = (uint32_t)warp * 32u + (uint32_t)src;
                    uint8_t* dst = s.a + (row >> 3) * kA_SBO + (row & 7) * 16;
                    *reinterpret_cast<uint4*>(dst + (0 + sub) * kA_LBO) = gather_plane_s(planes, p.plane_h, p.plane_w, sx, sy, sub);
                    *reinterpret_cast<uint4*>(dst + (4 + sub) * kA_LBO) = gather_plane_s(planes + plane_stride, p.plane_h, p.plane_w, sx, sz, sub);
                    *reinterpret_cast<uint4*>(dst + (8 + sub) * kA_LBO) = gather_plane_s(planes + 2 * plane_stride, p.plane_h, p.plane_w, sy, sz, sub);
                }
            }
            fence_proxy_async_smem();
            tc_fence_before();
            __syncthreads();

            // ---- GEMM1: base_x = F[128x96] W1^T
            if (tid == 0) {
                tc_fence_after();
#pragma unroll
                for (uint32_t k = 0; k < DecS::KF / 16; ++k)
                    umma_f16(tmem, make_desc_nosw(a_addr + k * 2 * kA_LBO, kA_LBO, kA_SBO),
                             make_desc_nosw(w1_addr + k * 2 * kW1_LBO, kW1_LBO, 128), idesc1, k != 0);
                umma_commit(&s.mma_bar);
            }
            mbar_wait(&s.mma_bar, bar_phase); bar_phase ^= 1;
            tc_fence_after();
            // base_act = SiLU(base_x + b1) -> fp16 -> K chunks 0..15 of this thread's row
            {
                uint8_t* dst = s.a + tid * 16;
#pragma unroll 1
                for (int c0 = 0; c0 < DecS::HID; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld32(tmem + lane_base + c0, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        uint4 o;
                        __half2* ho = reinterpret_cast<__half2*>(&o);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int nn = c0 + 8 * g + 2 * i;
                            ho[i] = __floats2half2_rn(silu_fast(__uint_as_float(v[8 * g + 2 * i]) + s.b1[nn]),
                                                      silu_fast(__uint_as_float(v[8 * g + 2 * i + 1]) + s.b1[nn + 1]));
                        }
                        *reinterpret_cast<uint4*>(dst + ((c0 >> 3) + g) * kA_LBO) = o;
                    }
                }
            }
            fence_proxy_async_smem();
            tc_fence_before();
            __syncthreads();

            // ---- GEMM2: [base_act | SH16] (K = 144) x W2^T -> 128 hidden + density
            if (tid == 0) {
                tc_fence_after();
#pragma unroll
                for (uint32_t k = 0; k < DecS::K2 / 16; ++k)
                    umma_f16(tmem, make_desc_nosw(a_addr + k * 2 * kA_LBO, kA_LBO, kA_SBO),
                             make_desc_nosw(w2_addr + k * 2 * kW2_LBO, kW2_LBO, 128), idesc2, k != 0);
                umma_commit(&s.mma_bar);
            }
            mbar_wait(&s.mma_bar, bar_phase); bar_phase ^= 1;
            tc_fence_after();
            float o_r = s.bc2[0], o_g = s.bc2[1], o_b = s.bc2[2];
#pragma unroll 1
            for (int c0 = 0; c0 < DecS::HID; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(tmem + lane_base + c0, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float h = silu_fast(__uint_as_float(v[i]) + s.b2[c0 + i]);
                    o_r = fmaf(h, s.wc2[0][c0 + i], o_r);
                    o_g = fmaf(h, s.wc2[1][c0 + i], o_g);
                    o_b = fmaf(h, s.wc2[2][c0 + i], o_b);
                }
            }
            float sd;
            {
                uint32_t v[16];
                tmem_ld16(tmem + lane_base + DecS::HID, v);
                tmem_ld_wait();
                sd = __uint_as_float(v[0]) + s.bd;
            }
            tc_fence_before();

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
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, kTmemCols); }
}

int render_s_launch(const RenderParams& p, int emulate_schedule, uint32_t* hist, int sms, cudaStream_t stream) {
    const size_t smem = sizeof(SmemS) + 128;
    static DeviceOnce attr_set;
    if (attr_set.first()) {
        SSDNERF_CUDA_OK(cudaFuncSetAttribute(k_render_s, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    int occ = 0;
    SSDNERF_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_render_s, kSThreads, smem));
    if (occ < 1) return set_error_msg(SSDNERF_ERR_CUDA, "render_fwd: variant S kernel does not fit on this device");
    if (occ > 2) occ = 2;   // 256 TMEM columns per CTA
    const uint32_t total_tiles = div_up(div_up(p.rays_per_scene, 32u), 4u) * p.num_scenes;
    const uint32_t grid = (uint32_t)min((uint64_t)sms * occ, (uint64_t)total_tiles);
    k_render_s<<<grid, kSThreads, smem, stream>>>(p, 0);
    SSDNERF_LAUNCH_OK();
    if (emulate_schedule) {
        if (int e = launch_schedule(hist, p.hist_bins, p.num_scenes, p.rays_per_scene, p.max_steps, p.budget, stream)) return e;
        k_render_s<<<grid, kSThreads, smem, stream>>>(p, 1);
        SSDNERF_LAUNCH_OK();
    }
    return 0;
}

}  // namespace ssdnerf
