// Per-op C-ABI entry points mirroring the reference's two pybind modules one-to-one
// (`_raymarching`: lib/ops/raymarching/src/bindings.cpp:5-18, raymarching.h:7-18;
//  `_shencoder`:   lib/ops/shencoder/src/bindings.cpp:5-6, shencoder.h:9,12).
// They let the reference's own Python driver (VolumeRenderer.forward) run op-by-op on this
// library for A/B testing.  Differences by design: raw pointers + explicit stream, kernels
// enqueue on the CALLER's stream (the reference uses the legacy default stream), every
// launch is error-checked, fp32 only (the reference's Python side casts to fp32 anyway).
#include "common.cuh"
#include "../../include/ssdnerf_b200.h"

namespace ssdnerf {

constexpr int kThreads = 128;

struct GlobalGrid {
    const uint8_t* __restrict__ g;
    __device__ __forceinline__ uint32_t operator()(uint32_t byte) const { return __ldg(g + byte); }
};

// ---------------------------------------------------------------- K1
__global__ void k_near_far(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                           const float* __restrict__ aabb, uint32_t N, float min_near,
                           float* __restrict__ nears, float* __restrict__ fars) {
    const uint32_t n = threadIdx.x + blockIdx.x * blockDim.x;
    if (n >= N) return;
    Ray r; ray_load(r, rays_o + 3 * (size_t)n, rays_d + 3 * (size_t)n);
    float near, far;
    near_far_aabb(r, aabb, min_near, near, far);
    nears[n] = near; fars[n] = far;
}

// ---------------------------------------------------------------- K2 (never called by the model; API completeness)
__global__ void k_sph_from_ray(const float* __restrict__ rays_o, const float* __restrict__ rays_d, float radius,
                               uint32_t N, float* __restrict__ coords) {
    const uint32_t n = threadIdx.x + blockIdx.x * blockDim.x;
    if (n >= N) return;
    const float ox = rays_o[3*n], oy = rays_o[3*n+1], oz = rays_o[3*n+2];
    const float dx = rays_d[3*n], dy = rays_d[3*n+1], dz = rays_d[3*n+2];
    const float A = dx * dx + dy * dy + dz * dz;
    const float B = ox * dx + oy * dy + oz * dz;
    const float Cc = ox * ox + oy * oy + oz * oz - radius * radius;
    const float t = (-B + sqrtf(B * B - A * Cc)) / A;
    const float x = ox + t * dx, y = oy + t * dy, z = oz + t * dz;
    const float theta = atan2f(sqrtf(x * x + z * z), y);
    const float phi = atan2f(z, x);
    coords[2*n] = 2 * theta * 0.3183098861837907f - 1;
    coords[2*n+1] = phi * 0.3183098861837907f;
}

// ---------------------------------------------------------------- K3 / K4
__global__ void k_morton3D(const int* __restrict__ coords, uint32_t N, int* __restrict__ indices) {
    const uint32_t n = threadIdx.x + blockIdx.x * blockDim.x;
    if (n >= N) return;
    indices[n] = (int)morton3D((uint32_t)coords[3*n], (uint32_t)coords[3*n+1], (uint32_t)coords[3*n+2]);
}
__global__ void k_morton3D_invert(const int* __restrict__ indices, uint32_t N, int* __restrict__ coords) {
    const uint32_t n = threadIdx.x + blockIdx.x * blockDim.x;
    if (n >= N) return;
    const int ind = indices[n];
    coords[3*n]   = (int)morton3D_invert((uint32_t)(ind >> 0));
    coords[3*n+1] = (int)morton3D_invert((uint32_t)(ind >> 1));
    coords[3*n+2] = (int)morton3D_invert((uint32_t)(ind >> 2));
}

// ---------------------------------------------------------------- K5: one thread packs 8 densities (two 16 B loads for f32, one for f16)
template <typename T> struct Vec8;
template <> struct Vec8<float> {
    static __device__ __forceinline__ void load(const float* p, float* v) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(p));
        const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
};
template <> struct Vec8<__half> {
    static __device__ __forceinline__ void load(const __half* p, float* v) {
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(p));
        const __half2* h = reinterpret_cast<const __half2*>(&a);
#pragma unroll
        for (int i = 0; i < 4; ++i) { const float2 f = __half22float2(h[i]); v[2*i] = f.x; v[2*i+1] = f.y; }
    }
};
template <typename T>
__global__ void k_packbits(const T* __restrict__ grid, uint32_t N, float thresh, uint8_t* __restrict__ bitfield) {
    const uint32_t n = threadIdx.x + blockIdx.x * blockDim.x;
    if (n >= N) return;
    float v[8];
    Vec8<T>::load(grid + 8 * (size_t)n, v);
    uint32_t bits = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) bits |= (v[i] > thresh) ? (1u << i) : 0u;
    bitfield[n] = (uint8_t)bits;
}

// ---------------------------------------------------------------- K9
__global__ void k_march_rays(uint32_t n_alive, uint32_t n_step, const int* __restrict__ rays_alive,
                             const float* __restrict__ rays_t, const float* __restrict__ rays_o,
                             const float* __restrict__ rays_d, MarchCfg c, const uint8_t* __restrict__ grid,
                             const float* __restrict__ fars, float* __restrict__ xyzs, float* __restrict__ dirs,
                             float* __restrict__ deltas, const float* __restrict__ noises) {
    const uint32_t n = threadIdx.x + blockIdx.x * blockDim.x;
    if (n >= n_alive) return;
    const int index = rays_alive[n];
    const float noise = noises ? noises[n] : 0.0f;
    Ray r; ray_load(r, rays_o + 3 * (size_t)index, rays_d + 3 * (size_t)index);
    xyzs += (size_t)n * n_step * 3; dirs += (size_t)n * n_step * 3; deltas += (size_t)n * n_step * 2;
    float t = rays_t[index];
    const float far = fars[index];
    t = __fmaf_rn(clampf(__fmul_rn(t, c.dt_gamma), c.dt_min, c.dt_max), noise, t);
    GlobalGrid gl{grid};
    uint32_t step = 0;
    while (t < far && step < n_step) {
        float x, y, z, dt; uint32_t vi;
        if (probe(c, r, gl, t, x, y, z, dt, vi)) {
            xyzs[0] = x; xyzs[1] = y; xyzs[2] = z;
            dirs[0] = r.dx; dirs[1] = r.dy; dirs[2] = r.dz;
            deltas[0] = dt; deltas[1] = t;
            t = __fadd_rn(t, dt);
            xyzs += 3; dirs += 3; deltas += 2; ++step;
        }
    }
}

// ---------------------------------------------------------------- K10
__global__ void k_composite_rays(uint32_t n_alive, uint32_t n_step, float T_thresh, int* __restrict__ rays_alive,
                                 float* __restrict__ rays_t, const float* __restrict__ sigmas,
                                 const float* __restrict__ rgbs, const float* __restrict__ deltas,
                                 float* __restrict__ weights_sum, float* __restrict__ depth, float* __restrict__ image) {
    const uint32_t n = threadIdx.x + blockIdx.x * blockDim.x;
    if (n >= n_alive) return;
    const int index = rays_alive[n];
    sigmas += (size_t)n * n_step; rgbs += (size_t)n * n_step * 3; deltas += (size_t)n * n_step * 2;
    float ws = weights_sum[index], d = depth[index];
    float r = image[3*(size_t)index], g = image[3*(size_t)index+1], b = image[3*(size_t)index+2];
    uint32_t step = 0;
    while (step < n_step) {
        if (deltas[0] == 0) break;
        const float alpha = 1.0f - __expf(-sigmas[0] * deltas[0]);
        const float T = 1.0f - ws;
        const float w = alpha * T;
        ws += w;
        d = __fmaf_rn(w, deltas[1], d);
        r = __fmaf_rn(w, rgbs[0], r); g = __fmaf_rn(w, rgbs[1], g); b = __fmaf_rn(w, rgbs[2], b);
        if (T < T_thresh) break;
        sigmas++; rgbs += 3; deltas += 2; step++;
    }
    if (step < n_step) rays_alive[n] = -1;
    else rays_t[index] = deltas[-1] + deltas[-2];
    weights_sum[index] = ws; depth[index] = d;
    image[3*(size_t)index] = r; image[3*(size_t)index+1] = g; image[3*(size_t)index+2] = b;
}

// ---------------------------------------------------------------- K6 (two passes, global atomics for offsets)
__global__ void k_march_rays_train(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                   const uint8_t* __restrict__ grid, MarchCfg c, uint32_t max_steps, uint32_t N, uint32_t M,
                                   const float* __restrict__ nears, const float* __restrict__ fars,
                                   float* __restrict__ xyzs, float* __restrict__ dirs, float* __restrict__ deltas,
                                   int* __restrict__ rays, int* __restrict__ counter, const float* __restrict__ noises) {
    const uint32_t n = threadIdx.x + blockIdx.x * blockDim.x;
    if (n >= N) return;
    Ray r; ray_load(r, rays_o + 3 * (size_t)n, rays_d + 3 * (size_t)n);
    const float near = nears[n], far = fars[n], noise = noises ? noises[n] : 0.0f;
    const float t0 = __fmaf_rn(clampf(__fmul_rn(near, c.dt_gamma), c.dt_min, c.dt_max), noise, near);
    GlobalGrid gl{grid};
    float t = t0; uint32_t num_steps = 0;
    while (t < far && num_steps < max_steps) {
        float x, y, z, dt; uint32_t vi;
        if (probe(c, r, gl, t, x, y, z, dt, vi)) { num_steps++; t = __fadd_rn(t, dt); }
    }
    const uint32_t point_index = atomicAdd(counter, (int)num_steps);
    const uint32_t ray_index = atomicAdd(counter + 1, 1);
    rays[3*ray_index] = (int)n; rays[3*ray_index+1] = (int)point_index; rays[3*ray_index+2] = (int)num_steps;
    if (num_steps == 0 || point_index + num_steps > M) return;
    xyzs += (size_t)point_index * 3; dirs += (size_t)point_index * 3; deltas += (size_t)point_index * 2;
    t = t0; uint32_t step = 0;
    while (t < far && step < num_steps) {
        float x, y, z, dt; uint32_t vi;
        if (probe(c, r, gl, t, x, y, z, dt, vi)) {
            xyzs[0] = x; xyzs[1] = y; xyzs[2] = z;
            dirs[0] = r.dx; dirs[1] = r.dy; dirs[2] = r.dz;
            deltas[0] = dt; deltas[1] = t;
            t = __fadd_rn(t, dt);
            xyzs += 3; dirs += 3; deltas += 2; ++step;
        }
    }
}

// ---------------------------------------------------------------- K7
__global__ void k_composite_train_fwd(const float* __restrict__ sigmas, const float* __restrict__ rgbs,
                                      const float* __restrict__ deltas, const int* __restrict__ rays,
                                      uint32_t M, uint32_t N, float T_thresh,
                                      float* __restrict__ weights_sum, float* __restrict__ depth, float* __restrict__ image) {
    const uint32_t n = threadIdx.x + blockIdx.x * blockDim.x;
    if (n >= N) return;
    const uint32_t index = rays[3*n], offset = rays[3*n+1], num_steps = rays[3*n+2];
    if (num_steps == 0 || offset + num_steps > M) {
        weights_sum[index] = 0; depth[index] = 0;
        image[3*(size_t)index] = 0; image[3*(size_t)index+1] = 0; image[3*(size_t)index+2] = 0;
        return;
    }
    sigmas += offset; rgbs += (size_t)offset * 3; deltas += (size_t)offset * 2;
    float T = 1.0f, r = 0, g = 0, b = 0, ws = 0, d = 0;
    for (uint32_t step = 0; step < num_steps; ++step) {
        const float alpha = 1.0f - __expf(-sigmas[0] * deltas[0]);
        const float w = alpha * T;
        r = __fmaf_rn(w, rgbs[0], r); g = __fmaf_rn(w, rgbs[1], g); b = __fmaf_rn(w, rgbs[2], b);
        d = __fmaf_rn(w, deltas[1], d);
        ws += w;
        T *= 1.0f - alpha;
        if (T < T_thresh) break;
        sigmas++; rgbs += 3; deltas += 2;
    }
    weights_sum[index] = ws; depth[index] = d;
    image[3*(size_t)index] = r; image[3*(size_t)index+1] = g; image[3*(size_t)index+2] = b;
}

// ---------------------------------------------------------------- K8
__global__ void k_composite_train_bwd(const float* __restrict__ grad_ws, const float* __restrict__ grad_image,
                                      const float* __restrict__ sigmas, const float* __restrict__ rgbs,
                                      const float* __restrict__ deltas, const int* __restrict__ rays,
                                      const float* __restrict__ weights_sum, const float* __restrict__ image,
                                      uint32_t M, uint32_t N, float T_thresh,
                                      float* __restrict__ grad_sigmas, float* __restrict__ grad_rgbs) {
    const uint32_t n = threadIdx.x + blockIdx.x * blockDim.x;
    if (n >= N) return;
    const uint32_t index = rays[3*n], offset = rays[3*n+1], num_steps = rays[3*n+2];
    if (num_steps == 0 || offset + num_steps > M) return;
    const float gws = grad_ws[index];
    const float gi0 = grad_image[3*(size_t)index], gi1 = grad_image[3*(size_t)index+1], gi2 = grad_image[3*(size_t)index+2];
    const float r_final = image[3*(size_t)index], g_final = image[3*(size_t)index+1], b_final = image[3*(size_t)index+2];
    const float ws_final = weights_sum[index];
    sigmas += offset; rgbs += (size_t)offset * 3; deltas += (size_t)offset * 2;
    grad_sigmas += offset; grad_rgbs += (size_t)offset * 3;
    float T = 1.0f, r = 0, g = 0, b = 0, ws = 0;
    for (uint32_t step = 0; step < num_steps; ++step) {
        const float alpha = 1.0f - __expf(-sigmas[0] * deltas[0]);
        const float w = alpha * T;
        r = __fmaf_rn(w, rgbs[0], r); g = __fmaf_rn(w, rgbs[1], g); b = __fmaf_rn(w, rgbs[2], b);
        ws += w;
        T *= 1.0f - alpha;
        if (T < T_thresh) break;
        grad_rgbs[0] = gi0 * w; grad_rgbs[1] = gi1 * w; grad_rgbs[2] = gi2 * w;
        grad_sigmas[0] = deltas[0] * (gi0 * (T * rgbs[0] - (r_final - r)) +
                                      gi1 * (T * rgbs[1] - (g_final - g)) +
                                      gi2 * (T * rgbs[2] - (b_final - b)) +
                                      gws * (1 - ws_final));
        sigmas++; rgbs += 3; deltas += 2; grad_sigmas++; grad_rgbs += 3;
    }
}

// ---------------------------------------------------------------- K11 / K12 (degree <= 4; the model only uses 4)
__global__ void k_sh_fwd(const float* __restrict__ inputs, float* __restrict__ outputs, uint32_t B, uint32_t degree,
                         float* __restrict__ dy_dx) {
    const uint32_t b = threadIdx.x + blockIdx.x * blockDim.x;
    if (b >= B) return;
    const float x = inputs[3*b], y = inputs[3*b+1], z = inputs[3*b+2];
    float o[16];
    sh16(x, y, z, o);
    const uint32_t C2 = degree * degree;
    for (uint32_t i = 0; i < C2; ++i) outputs[(size_t)b * C2 + i] = o[i];
    if (dy_dx) {
        // Jacobian of the 16 basis functions (shencoder.cu:124-190), rows dx | dy | dz
        const float xy = x * y, xz = x * z, yz = y * z, x2 = x * x, y2 = y * y, z2 = z * z;
        float jx[16], jy[16], jz[16];
        jx[0] = 0; jy[0] = 0; jz[0] = 0;
        jx[1] = 0; jy[1] = -0.48860251190291992f; jz[1] = 0;
        jx[2] = 0; jy[2] = 0; jz[2] = 0.48860251190291992f;
        jx[3] = -0.48860251190291992f; jy[3] = 0; jz[3] = 0;
        jx[4] = 1.0925484305920792f * y; jy[4] = 1.0925484305920792f * x; jz[4] = 0;
        jx[5] = 0; jy[5] = -1.0925484305920792f * z; jz[5] = -1.0925484305920792f * y;
        jx[6] = 0; jy[6] = 0; jz[6] = 1.8923493915151202f * z;
        jx[7] = -1.0925484305920792f * z; jy[7] = 0; jz[7] = -1.0925484305920792f * x;
        jx[8] = 1.0925484305920792f * x; jy[8] = -1.0925484305920792f * y; jz[8] = 0;
        jx[9] = -3.5402615395598609f * xy; jy[9] = -1.7701307697799304f * x2 + 1.7701307697799304f * y2; jz[9] = 0;
        jx[10] = 2.8906114426405538f * yz; jy[10] = 2.8906114426405538f * xz; jz[10] = 2.8906114426405538f * xy;
        jx[11] = 0; jy[11] = 0.45704579946446572f - 2.2852289973223288f * z2; jz[11] = -4.5704579946446566f * yz;
        jx[12] = 0; jy[12] = 0; jz[12] = 5.597644988851731f * z2 - 1.1195289977703462f;
        jx[13] = 0.45704579946446572f - 2.2852289973223288f * z2; jy[13] = 0; jz[13] = -4.5704579946446566f * xz;
        jx[14] = 2.8906114426405538f * xz; jy[14] = -2.8906114426405538f * yz; jz[14] = 1.4453057213202769f * x2 - 1.4453057213202769f * y2;
        jx[15] = -1.7701307697799304f * x2 + 1.7701307697799304f * y2; jy[15] = 3.5402615395598609f * xy; jz[15] = 0;
        float* dx = dy_dx + (size_t)b * 3 * C2;
        for (uint32_t i = 0; i < C2; ++i) { dx[i] = jx[i]; dx[C2 + i] = jy[i]; dx[2 * C2 + i] = jz[i]; }
    }
}
__global__ void k_sh_bwd(const float* __restrict__ grad, uint32_t B, uint32_t degree, const float* __restrict__ dy_dx,
                         float* __restrict__ grad_inputs) {
    const uint32_t t = threadIdx.x + blockIdx.x * blockDim.x;
    const uint32_t b = t / 3, d = t - b * 3;
    if (b >= B) return;
    const uint32_t C2 = degree * degree;
    float acc = 0;
    for (uint32_t ch = 0; ch < C2; ++ch) acc += grad[(size_t)b * C2 + ch] * dy_dx[(size_t)b * 3 * C2 + d * C2 + ch];
    grad_inputs[(size_t)b * 3 + d] += acc;
}

}  // namespace ssdnerf

using namespace ssdnerf;
#define LAUNCH_1D(kernel, N, stream, ...)                                                   \
    do {                                                                                    \
        if ((N) > 0) {                                                                      \
            kernel<<<div_up((uint32_t)(N), kThreads), kThreads, 0, (cudaStream_t)(stream)>>>(__VA_ARGS__); \
            SSDNERF_LAUNCH_OK();                                            \
        }                                                                                   \
    } while (0)

extern "C" {

int ssdnerf_near_far_from_aabb(const float* rays_o, const float* rays_d, const float* aabb, uint32_t N, float min_near,
                               float* nears, float* fars, void* stream) {
    LAUNCH_1D(k_near_far, N, stream, rays_o, rays_d, aabb, N, min_near, nears, fars);
    return 0;
}
int ssdnerf_sph_from_ray(const float* rays_o, const float* rays_d, float radius, uint32_t N, float* coords, void* stream) {
    LAUNCH_1D(k_sph_from_ray, N, stream, rays_o, rays_d, radius, N, coords);
    return 0;
}
int ssdnerf_morton3D(const int* coords, uint32_t N, int* indices, void* stream) {
    LAUNCH_1D(k_morton3D, N, stream, coords, N, indices);
    return 0;
}
int ssdnerf_morton3D_invert(const int* indices, uint32_t N, int* coords, void* stream) {
    LAUNCH_1D(k_morton3D_invert, N, stream, indices, N, coords);
    return 0;
}
int ssdnerf_packbits(const void* grid, int grid_is_half, uint32_t N, float thresh, uint8_t* bitfield, void* stream) {
    if (((uintptr_t)grid & 15u) != 0) return set_error_msg(SSDNERF_ERR_ARG, "packbits: grid must be 16-byte aligned");
    if (grid_is_half) LAUNCH_1D(k_packbits<__half>, N, stream, (const __half*)grid, N, thresh, bitfield);
    else LAUNCH_1D(k_packbits<float>, N, stream, (const float*)grid, N, thresh, bitfield);
    return 0;
}
int ssdnerf_march_rays_train(const float* rays_o, const float* rays_d, const uint8_t* grid, float bound, float dt_gamma,
                             uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M, const float* nears,
                             const float* fars, float* xyzs, float* dirs, float* deltas, int* rays, int* counter,
                             const float* noises, void* stream) {
    if (C < 1 || H < 1 || max_steps < 1) return set_error_msg(SSDNERF_ERR_ARG, "march_rays_train: C, H, max_steps must be >= 1");
    const MarchCfg c = make_march_cfg(bound, dt_gamma, max_steps, C, H);
    LAUNCH_1D(k_march_rays_train, N, stream, rays_o, rays_d, grid, c, max_steps, N, M, nears, fars, xyzs, dirs, deltas, rays, counter, noises);
    return 0;
}
int ssdnerf_composite_rays_train_forward(const float* sigmas, const float* rgbs, const float* deltas, const int* rays,
                                         uint32_t M, uint32_t N, float T_thresh, float* weights_sum, float* depth,
                                         float* image, void* stream) {
    LAUNCH_1D(k_composite_train_fwd, N, stream, sigmas, rgbs, deltas, rays, M, N, T_thresh, weights_sum, depth, image);
    return 0;
}
int ssdnerf_composite_rays_train_backward(const float* grad_weights_sum, const float* grad_image, const float* sigmas,
                                          const float* rgbs, const float* deltas, const int* rays, const float* weights_sum,
                                          const float* image, uint32_t M, uint32_t N, float T_thresh, float* grad_sigmas,
                                          float* grad_rgbs, void* stream) {
    LAUNCH_1D(k_composite_train_bwd, N, stream, grad_weights_sum, grad_image, sigmas, rgbs, deltas, rays, weights_sum, image,
              M, N, T_thresh, grad_sigmas, grad_rgbs);
    return 0;
}
int ssdnerf_march_rays(uint32_t n_alive, uint32_t n_step, const int* rays_alive, const float* rays_t, const float* rays_o,
                       const float* rays_d, float bound, float dt_gamma, uint32_t max_steps, uint32_t C, uint32_t H,
                       const uint8_t* grid, const float* nears, const float* fars, float* xyzs, float* dirs, float* deltas,
                       const float* noises, void* stream) {
    (void)nears;
    if (C < 1 || H < 1 || max_steps < 1) return set_error_msg(SSDNERF_ERR_ARG, "march_rays: C, H, max_steps must be >= 1");
    const MarchCfg c = make_march_cfg(bound, dt_gamma, max_steps, C, H);
    LAUNCH_1D(k_march_rays, n_alive, stream, n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, c, grid, fars, xyzs, dirs, deltas, noises);
    return 0;
}
int ssdnerf_composite_rays(uint32_t n_alive, uint32_t n_step, float T_thresh, int* rays_alive, float* rays_t,
                           const float* sigmas, const float* rgbs, const float* deltas, float* weights_sum, float* depth,
                           float* image, void* stream) {
    LAUNCH_1D(k_composite_rays, n_alive, stream, n_alive, n_step, T_thresh, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image);
    return 0;
}
int ssdnerf_sh_encode_forward(const float* inputs, float* outputs, uint32_t B, uint32_t D, uint32_t C, int calc_grad_inputs,
                              float* dy_dx, void* stream) {
    if (D != 3) return set_error_msg(SSDNERF_ERR_ARG, "sh_encode: input dim must be 3");
    if (C < 1 || C > 4) return set_error_msg(SSDNERF_ERR_ARG, "sh_encode: degree must be in [1, 4] (the model uses 4)");
    LAUNCH_1D(k_sh_fwd, B, stream, inputs, outputs, B, C, calc_grad_inputs ? dy_dx : nullptr);
    return 0;
}
int ssdnerf_sh_encode_backward(const float* grad, const float* inputs, uint32_t B, uint32_t D, uint32_t C, const float* dy_dx,
                               float* grad_inputs, void* stream) {
    (void)inputs;
    if (D != 3) return set_error_msg(SSDNERF_ERR_ARG, "sh_encode: input dim must be 3");
    if (C < 1 || C > 4) return set_error_msg(SSDNERF_ERR_ARG, "sh_encode: degree must be in [1, 4]");
    LAUNCH_1D(k_sh_bwd, B * 3, stream, grad, B, C, dy_dx, grad_inputs);
    return 0;
}

}  // extern "C"
