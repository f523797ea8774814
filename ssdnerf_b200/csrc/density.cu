// Occupancy-grid builder: one launch decodes the density of every (jittered) voxel centre of every scene,
// applies the EMA-max update in place (morton order) and emits per-block partial sums of clamp(grid, 0);
// a one-block reduction turns them into the threshold min(mean, density_thresh); a third launch packs bits.
//
// Replaces lib/models/autodecoders/base_nerf.py:318-389 (update_extra_state, full-update branch) which runs
// arange/meshgrid/cat, morton3D (K3), rand_like, grid_sample, 2 Linear, exp, scatter, where/maximum, clamp, mean
// (with a device->host sync for the threshold) and packbits (K5) as ~25 separate launches per iteration.
#include "common.cuh"
#include "dec_p.cuh"
#include "../../include/ssdnerf_b200.h"

namespace ssdnerf {

constexpr int kDenThreads = 256;

struct SmemDenP {
    float4 w1[DecP::KF][DecP::HID / 4];
    float b1[DecP::HID];
    float wd[DecP::HID];
    float bd;
    float red[kDenThreads / 32];
};

template <typename G>
__device__ __forceinline__ float grid_load(const G* g, size_t i);
template <> __device__ __forceinline__ float grid_load<float>(const float* g, size_t i) { return g[i]; }
template <> __device__ __forceinline__ float grid_load<__half>(const __half* g, size_t i) { return __half2float(g[i]); }

template <typename G>
__global__ void __launch_bounds__(kDenThreads) k_density_update_p(const float* __restrict__ planes, uint32_t Hp, uint32_t Wp,
                                                                 const float* __restrict__ blob, uint32_t num_scenes, uint32_t Gs,
                                                                 float bound, const float* __restrict__ jitter, float decay,
                                                                 G* __restrict__ grid, float* __restrict__ partials) {
    __shared__ SmemDenP s;
    {
        float* w1 = reinterpret_cast<float*>(s.w1);
        for (int i = threadIdx.x; i < DecP::KF * DecP::HID; i += kDenThreads) w1[i] = __ldg(blob + DecP::OFF_W1 + i);
        for (int i = threadIdx.x; i < DecP::HID; i += kDenThreads) {
            s.b1[i] = __ldg(blob + DecP::OFF_B1 + i);
            s.wd[i] = __ldg(blob + DecP::OFF_WD + i);
        }
        if (threadIdx.x == 0) s.bd = __ldg(blob + DecP::OFF_BD);
    }
    __syncthreads();
    const uint32_t G3 = Gs * Gs * Gs;
    const size_t gid = (size_t)blockIdx.x * kDenThreads + threadIdx.x;   // over scenes x voxels (ij-meshgrid order)
    float contrib = 0.0f;
    if (gid < (size_t)num_scenes * G3) {
        const uint32_t scene = (uint32_t)(gid / G3), v = (uint32_t)(gid - (size_t)scene * G3);
        const uint32_t k = v % Gs, j = (v / Gs) % Gs, i = v / (Gs * Gs);
        // xyz = (coord - (G-1)/2) * (2*bound/G) + (rand * 2*half - half), half = bound/G   (base_nerf.py:341-344)
        const float scale = 2.0f * bound / (float)Gs, half_w = bound / (float)Gs, mid = ((float)Gs - 1.0f) / 2.0f;
        float x = __fmul_rn(__fsub_rn((float)i, mid), scale);
        float y = __fmul_rn(__fsub_rn((float)j, mid), scale);
        float z = __fmul_rn(__fsub_rn((float)k, mid), scale);
        if (jitter) {
            const float* jt = jitter + (size_t)v * 3;
            x = __fadd_rn(x, __fsub_rn(__fmul_rn(__ldg(jt), 2.0f * half_w), half_w));
            y = __fadd_rn(y, __fsub_rn(__fmul_rn(__ldg(jt + 1), 2.0f * half_w), half_w));
            z = __fadd_rn(z, __fsub_rn(__fmul_rn(__ldg(jt + 2), 2.0f * half_w), half_w));
        }
        // density-only decode (triplane_decoder.py:119-160 with density_only=True)
        const float* pl = planes + (size_t)scene * 3 * Hp * Wp * DecP::CPAD;
        const size_t plane_stride = (size_t)Hp * Wp * DecP::CPAD;
        float f[DecP::KF];
        gather_plane_p(pl, Hp, Wp, x, y, f);
        gather_plane_p(pl + plane_stride, Hp, Wp, x, z, f + 6);
        gather_plane_p(pl + 2 * plane_stride, Hp, Wp, y, z, f + 12);
        float sd = s.bd;
#pragma unroll 4
        for (int o4 = 0; o4 < DecP::HID / 4; ++o4) {
            float a0 = s.b1[4 * o4], a1 = s.b1[4 * o4 + 1], a2 = s.b1[4 * o4 + 2], a3 = s.b1[4 * o4 + 3];
#pragma unroll
            for (int kk = 0; kk < DecP::KF; ++kk) {
                const float4 w = s.w1[kk][o4];
                a0 = fmaf(f[kk], w.x, a0); a1 = fmaf(f[kk], w.y, a1); a2 = fmaf(f[kk], w.z, a2); a3 = fmaf(f[kk], w.w, a3);
            }
            sd = fmaf(silu_f(a0), s.wd[4 * o4], sd);
            sd = fmaf(silu_f(a1), s.wd[4 * o4 + 1], sd);
            sd = fmaf(silu_f(a2), s.wd[4 * o4 + 2], sd);
            sd = fmaf(silu_f(a3), s.wd[4 * o4 + 3], sd);
        }
        const float sigma = __expf(sd);
        // tmp_grid[:, morton] = sigma.clamp(max=finfo.max).to(dtype); grid = where(valid, max(grid*decay, tmp), grid)
        const size_t gi = (size_t)scene * G3 + morton3D(i, j, k);
        float nv;
        if (sizeof(G) == 2) {
            const float tmp = __half2float(__float2half_rn(fminf(sigma, 65504.0f)));
            const float old = grid_load(grid, gi);
            nv = old;
            if (old >= 0.0f && tmp >= 0.0f) nv = fmaxf(__half2float(__float2half_rn(old * decay)), tmp);
            reinterpret_cast<__half*>(grid)[gi] = __float2half_rn(nv);
        } else {
            const float tmp = fminf(sigma, FLT_MAX);
            const float old = grid_load(grid, gi);
            nv = old;
            if (old >= 0.0f && tmp >= 0.0f) nv = fmaxf(__fmul_rn(old, decay), tmp);
            reinterpret_cast<float*>(grid)[gi] = nv;
        }
        contrib = fmaxf(nv, 0.0f);
    }
    // block partial sum of clamp(grid, min=0)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
    if ((threadIdx.x & 31) == 0) s.red[threadIdx.x >> 5] = contrib;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.0f;
        for (int w = 0; w < kDenThreads / 32; ++w) t += s.red[w];
        partials[blockIdx.x] = t;
    }
}

// thresh = min(mean(clamp(grid, 0)) over the whole batch, density_thresh)   (base_nerf.py:381-386)
__global__ void k_density_thresh(const float* __restrict__ partials, uint32_t n_partials, float count, float density_thresh,
                                 int round_mean_to_half, float* __restrict__ thresh, float* __restrict__ thresh_out) {
    __shared__ double red[32];
    double acc = 0.0;
    for (uint32_t i = threadIdx.x; i < n_partials; i += blockDim.x) acc += (double)partials[i];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (uint32_t w = 0; w < blockDim.x / 32; ++w) t += red[w];
        float mean = (float)(t / (double)count);
        if (round_mean_to_half) mean = __half2float(__float2half_rn(mean));   // torch.mean of an fp16 tensor returns fp16
        const float th = fminf(mean, density_thresh);
        *thresh = th;
        if (thresh_out) *thresh_out = th;
    }
}

template <typename G>
__global__ void k_density_pack(const G* __restrict__ grid, uint32_t N, const float* __restrict__ thresh, uint8_t* __restrict__ bitfield) {
    const uint32_t n = threadIdx.x + blockIdx.x * blockDim.x;
    if (n >= N) return;
    const float th = __ldg(thresh);
    uint32_t bits = 0;
    if (sizeof(G) == 2) {
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(grid) + n);
        const __half2* h = reinterpret_cast<const __half2*>(&a);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 f = __half22float2(h[i]);
            bits |= (f.x > th) ? (1u << (2 * i)) : 0u;
            bits |= (f.y > th) ? (1u << (2 * i + 1)) : 0u;
        }
    } else {
        const float4 a = __ldg(reinterpret_cast<const float4*>(grid) + 2 * (size_t)n);
        const float4 b = __ldg(reinterpret_cast<const float4*>(grid) + 2 * (size_t)n + 1);
        bits = (a.x > th) | ((a.y > th) << 1) | ((a.z > th) << 2) | ((a.w > th) << 3) | ((b.x > th) << 4) | ((b.y > th) << 5) |
               ((b.z > th) << 6) | ((b.w > th) << 7);
    }
    bitfield[n] = (uint8_t)bits;
}

}  // namespace ssdnerf

using namespace ssdnerf;

extern "C" {

size_t ssdnerf_density_workspace_bytes(uint32_t num_scenes, uint32_t grid_size) {
    const size_t blocks = ((size_t)num_scenes * grid_size * grid_size * grid_size + kDenThreads - 1) / kDenThreads;
    return 16 + blocks * sizeof(float);
}

int ssdnerf_density_update(int variant, const void* planes, uint32_t plane_h, uint32_t plane_w, const float* decoder_blob,
                           uint32_t num_scenes, uint32_t grid_size, float bound, const float* jitter, float decay,
                           void* density_grid, int grid_is_half, void* workspace, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (variant != SSDNERF_DEC_P && variant != SSDNERF_DEC_P_SIMT && variant != SSDNERF_DEC_P_TC && variant != SSDNERF_DEC_P_MMA)
        return set_error_msg(SSDNERF_ERR_ARG, "density_update: only decoder variant P (shipped configs) is implemented");
    if (!planes || !decoder_blob || !density_grid || !workspace) return set_error_msg(SSDNERF_ERR_ARG, "density_update: NULL argument");
    if (grid_size == 0 || grid_size > 1024 || (grid_size & (grid_size - 1))) return set_error_msg(SSDNERF_ERR_ARG, "density_update: grid_size must be a power of two");
    if (num_scenes == 0) return 0;
    const size_t total = (size_t)num_scenes * grid_size * grid_size * grid_size;
    const uint32_t blocks = (uint32_t)((total + kDenThreads - 1) / kDenThreads);
    float* partials = reinterpret_cast<float*>((unsigned char*)workspace + 16);
    if (grid_is_half)
        k_density_update_p<__half><<<blocks, kDenThreads, 0, stream>>>((const float*)planes, plane_h, plane_w, decoder_blob, num_scenes,
                                                                       grid_size, bound, jitter, decay, (__half*)density_grid, partials);
    else
        k_density_update_p<float><<<blocks, kDenThreads, 0, stream>>>((const float*)planes, plane_h, plane_w, decoder_blob, num_scenes,
                                                                      grid_size, bound, jitter, decay, (float*)density_grid, partials);
    SSDNERF_LAUNCH_OK();
    return 0;
}

int ssdnerf_density_pack(const void* density_grid, int grid_is_half, uint32_t num_scenes, uint32_t grid_size, float density_thresh,
                         uint8_t* bitfield, float* thresh_out, void* workspace, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!density_grid || !bitfield || !workspace) return set_error_msg(SSDNERF_ERR_ARG, "density_pack: NULL argument");
    if (num_scenes == 0) return 0;
    const size_t total = (size_t)num_scenes * grid_size * grid_size * grid_size;
    const uint32_t blocks = (uint32_t)((total + kDenThreads - 1) / kDenThreads);
    float* thresh = reinterpret_cast<float*>(workspace);
    const float* partials = reinterpret_cast<const float*>((unsigned char*)workspace + 16);
    k_density_thresh<<<1, 1024, 0, stream>>>(partials, blocks, (float)total, density_thresh, grid_is_half, thresh, thresh_out);
    SSDNERF_LAUNCH_OK();
    const uint32_t nbytes = (uint32_t)(total / 8);
    if (grid_is_half) k_density_pack<__half><<<div_up(nbytes, 256u), 256, 0, stream>>>((const __half*)density_grid, nbytes, thresh, bitfield);
    else k_density_pack<float><<<div_up(nbytes, 256u), 256, 0, stream>>>((const float*)density_grid, nbytes, thresh, bitfield);
    SSDNERF_LAUNCH_OK();
    return 0;
}

}  // extern "C"
