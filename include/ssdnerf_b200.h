/*
 * ssdnerf_b200.h -- C ABI of libssdnerf_b200.so: B200 (sm_100a) kernels for SSDNeRF's two hot paths.
 *
 * Conventions (all entry points):
 *   - plain pointers + sizes, no torch / ATen types; every pointer is a DEVICE pointer owned by the
 *     caller unless the parameter name ends in `_host`;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); kernels are enqueued
 *     asynchronously on it and nothing synchronises;
 *   - no hidden allocations: scratch is caller-provided (see the *_workspace_bytes queries);
 *   - return 0 on success, a negative SSDNERF_ERR_* code otherwise; ssdnerf_last_error() returns a
 *     thread-local description of the last failure.
 *
 * "replaces:" lines cite the reference interface (Lakonik/SSDNeRF @ b9d195d) each function stands in for.
 */
#ifndef SSDNERF_B200_H_
#define SSDNERF_B200_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* exported symbols (the library is built with -fvisibility=hidden) */
#define SSDNERF_API __attribute__((visibility("default")))

#define SSDNERF_OK 0
#define SSDNERF_ERR_CUDA (-1)   /* a CUDA runtime / driver call failed */
#define SSDNERF_ERR_ARG (-2)    /* invalid argument (shape, alignment, unsupported variant) */
#define SSDNERF_ERR_ARCH (-3)   /* device is not sm_100 */

SSDNERF_API const char* ssdnerf_last_error(void);
/* library version and the SM architecture it was compiled for (100) */
SSDNERF_API int ssdnerf_version(void);
SSDNERF_API int ssdnerf_compiled_arch(void);
/* number of CUDA kernels this library has launched (or recorded into a capturing stream) in this process */
SSDNERF_API unsigned long long ssdnerf_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * 1. Legacy per-op entry points == the reference's pybind FFI, one to one.
 *    replaces: lib/ops/raymarching/src/raymarching.h:7-18 + bindings.cpp:5-18 (module `_raymarching`)
 *              lib/ops/shencoder/src/shencoder.h:9,12 + bindings.cpp:5-6    (module `_shencoder`)
 *    Same argument order and meaning; at::Tensor -> pointer; fp32 only; caller allocates (and, where
 *    the reference's Python wrapper does, zero-fills) every output.
 * ---------------------------------------------------------------------------------------------- */
/* replaces: near_far_from_aabb (raymarching.cu:148-156, kernel :92-145) */
SSDNERF_API int ssdnerf_near_far_from_aabb(const float* rays_o, const float* rays_d, const float* aabb, uint32_t N, float min_near,
                               float* nears, float* fars, void* stream);
/* replaces: sph_from_ray (raymarching.cu:200-208) -- unused by the model, kept for API completeness */
SSDNERF_API int ssdnerf_sph_from_ray(const float* rays_o, const float* rays_d, float radius, uint32_t N, float* coords, void* stream);
/* replaces: morton3D / morton3D_invert (raymarching.cu:229-232, :257-260) */
SSDNERF_API int ssdnerf_morton3D(const int* coords, uint32_t N, int* indices, void* stream);
SSDNERF_API int ssdnerf_morton3D_invert(const int* indices, uint32_t N, int* coords, void* stream);
/* replaces: packbits (raymarching.cu:292-300); N = number of OUTPUT bytes; grid is fp32 or fp16 (grid_is_half) */
SSDNERF_API int ssdnerf_packbits(const void* grid, int grid_is_half, uint32_t N, float thresh, uint8_t* bitfield, void* stream);
/* replaces: march_rays_train (raymarching.cu:484-492, kernel :312-482); counter = int[2] {points, rays} */
SSDNERF_API int ssdnerf_march_rays_train(const float* rays_o, const float* rays_d, const uint8_t* grid, float bound, float dt_gamma,
                             uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M, const float* nears,
                             const float* fars, float* xyzs, float* dirs, float* deltas, int* rays, int* counter,
                             const float* noises, void* stream);
/* replaces: composite_rays_train_forward / _backward (raymarching.cu:584-592, :690-698) */
SSDNERF_API int ssdnerf_composite_rays_train_forward(const float* sigmas, const float* rgbs, const float* deltas, const int* rays,
                                         uint32_t M, uint32_t N, float T_thresh, float* weights_sum, float* depth,
                                         float* image, void* stream);
SSDNERF_API int ssdnerf_composite_rays_train_backward(const float* grad_weights_sum, const float* grad_image, const float* sigmas,
                                          const float* rgbs, const float* deltas, const int* rays, const float* weights_sum,
                                          const float* image, uint32_t M, uint32_t N, float T_thresh, float* grad_sigmas,
                                          float* grad_rgbs, void* stream);
/* replaces: march_rays (raymarching.cu:815-822, kernel :706-812); noises may be NULL (== zeros) */
SSDNERF_API int ssdnerf_march_rays(uint32_t n_alive, uint32_t n_step, const int* rays_alive, const float* rays_t, const float* rays_o,
                       const float* rays_d, float bound, float dt_gamma, uint32_t max_steps, uint32_t C, uint32_t H,
                       const uint8_t* grid, const float* nears, const float* fars, float* xyzs, float* dirs, float* deltas,
                       const float* noises, void* stream);
/* replaces: composite_rays (raymarching.cu:916-922, kernel :826-913) */
SSDNERF_API int ssdnerf_composite_rays(uint32_t n_alive, uint32_t n_step, float T_thresh, int* rays_alive, float* rays_t,
                           const float* sigmas, const float* rgbs, const float* deltas, float* weights_sum, float* depth,
                           float* image, void* stream);
/* replaces: sh_encode_forward / sh_encode_backward (shencoder.cu:386-399, :416-440); degree C <= 4 */
SSDNERF_API int ssdnerf_sh_encode_forward(const float* inputs, float* outputs, uint32_t B, uint32_t D, uint32_t C, int calc_grad_inputs,
                              float* dy_dx, void* stream);
SSDNERF_API int ssdnerf_sh_encode_backward(const float* grad, const float* inputs, uint32_t B, uint32_t D, uint32_t C, const float* dy_dx,
                               float* grad_inputs, void* stream);

/* ------------------------------------------------------------------------------------------------
 * 2. Fused triplane renderer (inference).
 *    replaces, as ONE launch sequence with no host synchronisation:
 *      lib/models/decoders/base_volume_renderer.py:79-123  (eval branch of VolumeRenderer.forward:
 *          K1 near/far + host loop of march_rays -> point_decode -> composite_rays -> compaction)
 *      lib/models/decoders/triplane_decoder.py:119-179     (TriPlaneDecoder.point_decode)
 *      lib/core/utils/nerf_utils.py:17-61                  (get_cam_rays, when rays are generated in-kernel)
 *      lib/models/autodecoders/base_nerf.py:520-523        (image + bg * (1 - weights_sum), optional)
 * ---------------------------------------------------------------------------------------------- */

/* Decoder variants (template instantiations). */
#define SSDNERF_DEC_P 0 /* shipped configs: base 3*6->64, density 64->1, dir_net 16->64, color 64->3
                           (configs/paper_cfgs/ssdnerf_cars_uncond.py:40-51) */
#define SSDNERF_DEC_P_SIMT 2 /* decoder P on the CUDA cores in plain fp32 (csrc/render_fused.cu) */
#define SSDNERF_DEC_P_TC 3   /* decoder P with the base layer as a split-precision fp16 tcgen05 GEMM (csrc/render_ptc.cu);
                                SSDNERF_DEC_P selects the fastest P kernel (currently SSDNERF_DEC_P_MMA2) */
#define SSDNERF_DEC_P_MMA 4  /* decoder P, warp-synchronous: per-warp split-precision mma.sync base layer (csrc/render_p2.cu) */
#define SSDNERF_DEC_P_MMA2 7 /* decoder P, warp-synchronous v2: one exponential per hidden unit shared by both branches, dir_net on the
                                tensor cores, one reciprocal per four sigmoids (csrc/render_p3.cu); what SSDNERF_DEC_P selects */
#define SSDNERF_DEC_S_TC 6   /* decoder S, CTA-synchronous tcgen05 kernel (csrc/render_tc.cu) */
#define SSDNERF_DEC_S_MMA 5  /* decoder S, warp-synchronous mma.sync kernel (csrc/render_s2.cu); SSDNERF_DEC_S selects the faster one */
#define SSDNERF_DEC_S 1 /* TriPlaneDecoder class defaults: base 3*32->128, density 128->1, color (128+16)->128->3
                           (lib/models/decoders/triplane_decoder.py:24-39) */

/* Size in floats of the packed fp32 decoder-weight blob for a variant (layout: ssdnerf_b200/decoder_pack.py). */
SSDNERF_API size_t ssdnerf_decoder_blob_floats(int variant);

/* Re-layout one batch of triplanes for the gather:
 *   code  fp32 [B][3][C][Hp][Wp]  (reference layout, triplane_decoder.py:123)
 *   -> planes [B][3][Hp][Wp][Cpad] channels-last, fp32 (variant P, Cpad = 8) or fp16 (variant S, Cpad = 32). */
SSDNERF_API size_t ssdnerf_planes_bytes(int variant, uint32_t B, uint32_t Hp, uint32_t Wp);
SSDNERF_API int ssdnerf_pack_planes(int variant, const float* code, uint32_t B, uint32_t C, uint32_t Hp, uint32_t Wp, void* planes,
                        void* stream);

/* Stand-alone point decode (variant P): sigma [M] (and rgb [M][3] when `rgbs` != NULL) at M points given as the concatenation
 * of per-scene point lists; scene_offsets [B+1] (int64, device) holds the prefix sums, scene_offsets[B] == M.
 *   replaces: lib/models/decoders/triplane_decoder.py:104-117 (xyz_transform), :119-179 (point_decode), :181-184 (point_density_decode) */
SSDNERF_API int ssdnerf_point_decode(int variant, const void* planes, uint32_t plane_h, uint32_t plane_w, const float* decoder_blob,
                                     const float* xyzs, const float* dirs, const long long* scene_offsets, uint32_t num_scenes,
                                     unsigned long long num_points, float* sigmas, float* rgbs, void* stream);

typedef struct ssdnerf_render_args {
    int variant;              /* SSDNERF_DEC_* */
    uint32_t num_scenes;      /* B */
    uint32_t rays_per_scene;  /* N (explicit rays) or V*h*w (camera mode) */
    /* --- rays: either explicit ... */
    const float* rays_o;      /* [B][N][3] or NULL */
    const float* rays_d;      /* [B][N][3] or NULL */
    /* --- ... or generated in-kernel from cameras (nerf_utils.py:17-61) */
    const float* poses;       /* [B][V][4][4] row-major c2w, or NULL */
    const float* intrinsics;  /* [B][V][4] = fx, fy, cx, cy, or NULL */
    uint32_t num_views, img_h, img_w;
    /* --- scene */
    const void* planes;       /* from ssdnerf_pack_planes */
    uint32_t plane_h, plane_w;
    const uint8_t* bitfield;  /* [B][H^3/8], morton order, LSB first */
    uint32_t grid_size;       /* H */
    const float* decoder_blob;/* packed weights, shared by all scenes */
    const float* dt_gamma;    /* [B] or NULL (== 0) */
    float bound, min_near, T_thresh, bg_color;
    uint32_t max_steps;
    int emulate_schedule;     /* 1: reproduce the reference host loop's per-scene sample budget exactly
                                    (n_step = clamp(N / n_alive, 1, 8) quanta until step >= max_steps) */
    /* --- outputs, [B][N] / [B][N][3]; any may be NULL except image + weights_sum */
    float* weights_sum;
    float* depth;
    float* image;             /* un-blended, == TriPlaneDecoder.forward()['image'] */
    float* rgb_blend;         /* image + bg_color * (1 - weights_sum), optional */
    int32_t* num_samples;     /* samples composited per ray, optional */
    int32_t* voxel_trace;     /* optional [B][N][trace_cap] occupancy-bit index of every composited sample (-1 padded) */
    uint32_t trace_cap;
    void* debug_phase_cycles; /* optional uint64[8]: per-phase clock64 totals of thread 0 of every CTA (SSDNERF_DEC_P_TC only) */
    /* --- scratch */
    void* workspace;          /* >= ssdnerf_render_workspace_bytes(...) bytes, 16-byte aligned */
    size_t workspace_bytes;
} ssdnerf_render_args;

SSDNERF_API size_t ssdnerf_render_workspace_bytes(uint32_t num_scenes, uint32_t rays_per_scene, uint32_t max_steps);
SSDNERF_API int ssdnerf_render_fwd(const ssdnerf_render_args* args, void* stream);

/* ------------------------------------------------------------------------------------------------
 * 2b. Fused differentiable renderer (train / guidance branch), variant P.
 *    replaces: lib/models/decoders/base_volume_renderer.py:59-77 (march_rays_train -> point_decode ->
 *              batch_composite_rays_train), lib/ops/raymarching/raymarching.py:200-395 and the autograd
 *              graph of triplane_decoder.py:119-179 for the gradient w.r.t. the triplane code;
 *              lib/models/autodecoders/base_nerf.py:276-296 (pixel MSE + RegLoss terms of BaseNeRF.loss).
 *    forward : per ray K6 march (perturbed start t0 = near + clamp(near*dt_gamma) * noise), decode, K7
 *              compositing; writes weights_sum / depth / image [B][N] (and the per-ray sample count).
 *    backward: re-marches the same samples, applies K8 with the saved weights_sum / image and
 *              accumulates d(loss)/d(planes) into grad_planes (channels-last, layout of ssdnerf_pack_planes,
 *              caller zero-fills); ssdnerf_unpack_plane_grads converts to the [B][3][6][H][W] code layout.
 *    The decoder weights receive a gradient only when grad_decoder_blob is given (frozen decoder otherwise, diffusion_nerf.py:273).
 * ---------------------------------------------------------------------------------------------- */
typedef struct ssdnerf_render_train_args {
    int variant;               /* SSDNERF_DEC_P */
    uint32_t num_scenes, rays_per_scene;
    const float* rays_o;       /* [B][N][3] */
    const float* rays_d;       /* [B][N][3] */
    const float* noises;       /* [B][N] uniform [0,1) (perturb=True) or NULL */
    const void* planes;        /* from ssdnerf_pack_planes(SSDNERF_DEC_P, ...) */
    uint32_t plane_h, plane_w;
    const uint8_t* bitfield;   /* [B][H^3/8] */
    uint32_t grid_size;
    const float* decoder_blob;
    const float* dt_gamma;     /* [B] or NULL */
    float bound, min_near, T_thresh;
    uint32_t max_steps;
    float* weights_sum;        /* [B][N]    forward: out, backward: in (saved) */
    float* depth;              /* [B][N]    forward: out (optional) */
    float* image;              /* [B][N][3] forward: out, backward: in (saved) */
    int32_t* num_samples;      /* [B][N]    forward: out (optional) */
    const float* grad_ws;      /* [B][N]    backward: in (optional) */
    const float* grad_image;   /* [B][N][3] backward: in */
    float* grad_planes;        /* [B][3][H][W][8] fp32, backward: accumulated into */
    uint32_t* counter;         /* 4 bytes of device scratch (tile counter) */
    float* grad_decoder_blob;  /* backward, optional: [ssdnerf_decoder_blob_floats(SSDNERF_DEC_P)] fp32, accumulated into --
                                  d(loss)/d(decoder weights) in the layout of decoder_blob (trainable decoder:
                                  lib/models/autodecoders/multiscene_nerf.py:203-207 loss.backward() + decoder optimizer step) */
} ssdnerf_render_train_args;

/* lib/core/utils/nerf_utils.py:17-61 get_cam_rays: poses [B][V][4][4] c2w, intrinsics [B][V][4] -> rays_o / rays_d [B][V][h][w][3] */
SSDNERF_API int ssdnerf_cam_rays(const float* poses, const float* intrinsics, uint32_t B, uint32_t V, uint32_t h, uint32_t w,
                     float* rays_o, float* rays_d, void* stream);
SSDNERF_API int ssdnerf_render_train_fwd(const ssdnerf_render_train_args* args, void* stream);
SSDNERF_API int ssdnerf_render_train_bwd(const ssdnerf_render_train_args* args, void* stream);
/* grad_code[b][p][c][y][x] (=|+=) grad_planes[b][p][y][x][c] + reg_coef * code[...]   (code may be NULL) */
SSDNERF_API int ssdnerf_unpack_plane_grads(const float* grad_planes, const float* code, float reg_coef, uint32_t B, uint32_t Hp, uint32_t Wp,
                               int accumulate, float* grad_code, void* stream);
/* out = image + bg * (1 - ws);  *loss += coef_loss * sum (out - target)^2;  grad_image = coef_grad * (out - target);
 * grad_ws = -bg * sum_c grad_image.  (MSELoss 'mean' * weights are folded into the two coefficients by the caller.) */
SSDNERF_API int ssdnerf_mse_render_loss(const float* image, const float* weights_sum, const float* target, uint64_t rays, float bg_color,
                            float coef_loss, float coef_grad, float* out_rgb /* optional */, float* grad_image, float* grad_ws,
                            float* loss /* [1], accumulated */, void* stream);

/* ------------------------------------------------------------------------------------------------
 * 3. Occupancy-grid builder.
 *    replaces: lib/models/autodecoders/base_nerf.py:318-389 (update_extra_state, full-update branch:
 *              jittered voxel centres -> morton3D -> point_density_decode -> EMA max -> mean -> packbits)
 *    Two launches per iteration: ssdnerf_density_update (decode + max + per-block partial sums) then
 *    ssdnerf_density_pack (threshold = min(mean, density_thresh), bit pack).  No host sync.
 * ---------------------------------------------------------------------------------------------- */
SSDNERF_API size_t ssdnerf_density_workspace_bytes(uint32_t num_scenes, uint32_t grid_size);
SSDNERF_API int ssdnerf_density_update(int variant, const void* planes, uint32_t plane_h, uint32_t plane_w, const float* decoder_blob,
                           uint32_t num_scenes, uint32_t grid_size, float bound,
                           const float* jitter,   /* [G^3][3] uniform [0,1) in ij-meshgrid order (== torch.rand_like of
                                                     base_nerf.py:344), shared by all scenes; NULL = voxel centres */
                           float decay, void* density_grid, int grid_is_half, void* workspace, void* stream);
SSDNERF_API int ssdnerf_density_pack(const void* density_grid, int grid_is_half, uint32_t num_scenes, uint32_t grid_size,
                         float density_thresh, uint8_t* bitfield, float* thresh_out /* [1], optional */,
                         void* workspace, void* stream);

/* ------------------------------------------------------------------------------------------------
 * 4. UNet building blocks of the DDIM loop (tcgen05 tensor-core GEMM / implicit-GEMM convolution +
 *    memory-bound glue kernels).  Activations are NHWC fp16; accumulation is fp32.
 *    replaces: cuDNN/cuBLAS calls under lib/models/architecture/ddpm/denoising.py:191-216 and
 *              modules.py:28-48 (+ mmgen 0.7.2 DenoisingResBlock / NormWithEmbedding / QKVAttention /
 *              DenoisingDownsample / DenoisingUpsample forwards), and the DDIM algebra of
 *              lib/models/diffusions/gaussian_diffusion.py:180-240,264-293.
 * ---------------------------------------------------------------------------------------------- */
typedef struct ssdnerf_gemm_args {
    /* D[m, n] = alpha * sum_{tap,k} A_tap[m, k] * B[tap][n, k] + bias_n[n] + residual[m, n]
     * A: fp16, viewed as {K, d1, d2, d3} with K contiguous and byte strides a*_strides[0..2] for d1, d2, d3;
     *    an output tile is 128 rows = a (b1 x b2 x b3) box of (d1, d2, d3); taps = 9 shifts the box origin by
     *    (kx-1, ky-1) in (d1, d2) with zero fill outside (3x3 convolution, padding 1, stride 1).
     *    Optional second source a2 is concatenated after a1 along K (channel concat of the UNet skip). */
    const void* a1; uint64_t a1_strides[3]; uint32_t k1;
    const void* a2; uint64_t a2_strides[3]; uint32_t k2;
    uint32_t d1, d2, d3, b1, b2, b3;
    uint32_t taps;
    /* B: fp16 {K = k1 + k2, n_rows_b, bx2, bx3}, K contiguous, byte strides b_strides[0..2];
     *    conv / plain: coordinate 2 = tap; b_batched: coordinates (2, 3) = the tile's (d2, d3) tile indices */
    const void* b; uint64_t b_strides[3]; uint32_t n, n_rows_b, bx2, bx3; uint32_t b_batched;
    uint32_t bn;              /* N tile: 0 = auto, else 64 / 128 / 256 */
    uint32_t cluster;         /* 0 = auto, 1 = no cluster, 2 = CTA pairs along M with TMA multicast of the B tile */
    float alpha;
    const float* bias_n;      /* [n] fp32 or NULL */
    const void* residual;     /* fp16, addressed like out, or NULL */
    void* out; uint32_t out_f32; long long so1, so2, so3; /* element strides of d1, d2, d3; columns contiguous */
    /* optional fused GroupNorm statistics of the output: qstats [images][n/4][2] += {sum, sum of squares} of every 4-channel quad
     * (caller zero-fills); image of a row = index along d3 (stats_hw == 0) or (index along d1) / stats_hw (flattened rows) */
    float* qstats; uint32_t stats_hw;
    void* debug_cycles;      /* optional uint64[8] device counters (pipeline wait cycles per role, summed over CTAs); NULL in production */
    uint32_t algo;           /* 0 = auto, 1 = generic tile kernel, 2 = row-pair 3x3 convolution (128-pixel rows, 128 output channels) */
    /* generalised K-slabs: taps in [1, 9] with tap_offsets[2t], [2t+1] = shift of slab t in (d1, d2) (NULL: the 3x3 / 1x1 defaults) --
     * e.g. the four 2x2-tap phase convolutions a nearest-x2 upsample + 3x3 convolution decomposes into;
     * a_stride 2 = stride-2 convolution: (d1, d2) are output extents, a1 / a2 describe the (2 d1 x 2 d2) input read at every 2nd pixel */
    const int8_t* tap_offsets;   /* HOST pointer, 2 * taps entries, or NULL */
    uint32_t a_stride;           /* 0 / 1: dense; 2: stride-2 convolution */
} ssdnerf_gemm_args;
SSDNERF_API int ssdnerf_gemm_f16(const ssdnerf_gemm_args* args, void* stream);

/* x fp32 [B,C,H,W] -> fp16 [B,H,W,Cpad] (zero-padded channels): UNet input layout */
SSDNERF_API int ssdnerf_nchw_to_nhwc_f16(const float* x, uint32_t B, uint32_t C, uint32_t H, uint32_t W, uint32_t Cpad, void* out, void* stream);
/* GroupNorm over the channel concat of x1 [B,HW,C1] and optional x2 [B,HW,C2] (fp16 NHWC):
 * stats [B][groups][2] += {sum, sum of squares} (caller zero-fills); apply: y = GN(x)*gamma+beta, optionally
 * y = y*(1+scale)+shift with scale_shift = [scale(C) | shift(C)] per sample (NormWithEmbedding, use_scale_shift_norm)
 * and SiLU; writes the concatenated fp16 result. */
SSDNERF_API int ssdnerf_gn_stats(const void* x1, uint32_t C1, const void* x2, uint32_t C2, uint32_t B, uint32_t HW, uint32_t groups,
                                 float* stats, void* stream);
/* same as ssdnerf_gn_apply with the statistics given per 4-channel quad (as emitted by ssdnerf_gemm_f16's qstats):
 * q1 [B][C1/4][2], q2 [B][C2/4][2] */
SSDNERF_API int ssdnerf_gn_apply_q(const void* x1, uint32_t C1, const void* x2, uint32_t C2, uint32_t B, uint32_t HW, uint32_t groups,
                                   const float* q1, const float* q2, const float* gamma, const float* beta, const float* scale_shift,
                                   long long ss_batch_stride, float eps, int do_silu, void* out, void* stream);
SSDNERF_API int ssdnerf_gn_apply(const void* x1, uint32_t C1, const void* x2, uint32_t C2, uint32_t B, uint32_t HW, uint32_t groups,
                                 const float* stats, const float* gamma, const float* beta, const float* scale_shift,
                                 long long ss_batch_stride, float eps, int do_silu, void* out, void* stream);
/* [B,H,W,C] -> [B,H/2,W/2,9C] patches of a 3x3 stride-2 pad-1 convolution (DenoisingDownsample), K index = tap*C + c */
SSDNERF_API int ssdnerf_im2col_s2(const void* x, uint32_t B, uint32_t H, uint32_t W, uint32_t C, void* out, void* stream);
/* nearest-neighbour x2 (DenoisingUpsample) */
SSDNERF_API int ssdnerf_upsample2x(const void* x, uint32_t B, uint32_t H, uint32_t W, uint32_t C, void* out, void* stream);
/* Fused GroupNorm(32) (+ NormWithEmbedding scale/shift) + SiLU + 3x3 convolution, 128-pixel-wide images, 128 output channels
 * (the UNet's 128 x 128 level).  replaces: mmgen DenoisingResBlock's `conv(act(norm(x)))` pairs as used by
 * lib/models/architecture/ddpm/modules.py:51-110 -- GroupNorm apply + SiLU + Conv2d, without materialising the normalised activation.
 * x1 (+ x2, channel concat) are the RAW NHWC fp16 tensors; q1 / q2 their quad statistics as emitted by ssdnerf_gemm_f16 (qstats). */
typedef struct ssdnerf_conv_gn_args {
    const void* x1; uint32_t C1;          /* [B][H][128][C1] fp16 */
    const void* x2; uint32_t C2;          /* optional second input [B][H][128][C2] */
    uint32_t B, H;                         /* H even */
    const float* q1; const float* q2;      /* [B][C/4][2] */
    const float* gamma; const float* beta; /* [C1 + C2] */
    const float* scale_shift;              /* optional [B][...]: row b holds scale[C] | shift[C] at scale_shift + b * ss_batch_stride */
    long long ss_batch_stride;
    float eps;
    const void* w; uint32_t w_rows;        /* packed fp16 weight [9][w_rows >= 128][C1 + C2] */
    const float* bias;                     /* [128] or NULL */
    const void* residual;                  /* [B][H][128][128] fp16 or NULL */
    void* out;                             /* [B][H][128][128] fp16 */
    float* qstats;                         /* optional [B][32][2] quad statistics of the output (caller zero-fills) */
    void* coef_workspace;                  /* device scratch, B * (C1 + C2) * 8 bytes, 16-byte aligned (per-image affine table) */
    void* debug_cycles;                    /* optional uint64[8] pipeline wait counters (debug); NULL in production */
} ssdnerf_conv_gn_args;
SSDNERF_API int ssdnerf_conv3x3_gn_f16(const ssdnerf_conv_gn_args* args, void* stream);
/* fused attention: out[b][t][h*ch + d] = softmax_s(scale * q[b,t,h,:] . k[b,s,h,:]) v[b,s,h,d], scores kept on chip (flash-style);
 * qkv fp16 [B][T][3*heads*ch] with the legacy head layout (modules.py:36-48), ch in {64, 128}, T % 64 == 0 */
SSDNERF_API int ssdnerf_flash_attn(const void* qkv, uint32_t B, uint32_t T, uint32_t heads, uint32_t ch, float scale, void* out, void* stream);
/* P = softmax(S) along the last axis; S fp32 [rows][T] -> P fp16 */
SSDNERF_API int ssdnerf_softmax_rows(const float* S, uint32_t rows, uint32_t T, void* P, void* stream);
/* Vt[b][h][c][t] = qkv[b][t][h*3ch + 2ch + c] (legacy head layout of modules.py:36-48) */
SSDNERF_API int ssdnerf_transpose_v(const void* qkv, uint32_t B, uint32_t T, uint32_t heads, uint32_t ch, void* vt, void* stream);
/* One DDIM step of the V-parameterisation (gaussian_diffusion.py:198-230,264-293), x_t fp32 [B,C,H,W] updated in place:
 *   x0 = clamp(c0*x_t - c1*v);  eps = (x_t - c0*x0)/c1;  x_prev = c2*x0 + c3*eps,  coef[step] = {c0,c1,c2,c3}
 * v fp32 NHWC [B,H,W,Cv]; step index read from *step_ptr (device) or 0; optionally writes x0 and the next UNet input. */
SSDNERF_API int ssdnerf_ddim_update(float* x_t, const float* v, uint32_t B, uint32_t C, uint32_t H, uint32_t W, uint32_t Cv,
                                    const float* coef, const int* step_ptr, int clip, float clip_lo, float clip_hi, float* x0_out,
                                    void* next_in, uint32_t Cpad, void* stream);
/* device-side step bookkeeping of the graph-replayed DDIM loop: *step_ptr = value (set) or += value; and
 * dst[c][:] = table[*step_ptr][:] for c < copies (selects the per-step time-embedding projections) */
SSDNERF_API int ssdnerf_step_counter(int* step_ptr, int value, int set, void* stream);
SSDNERF_API int ssdnerf_select_row(const float* table, uint32_t row_elems, const int* step_ptr, float* dst, uint32_t copies, void* stream);

/* ------------------------------------------------------------------------------------------------
 * 4b. UNet input-gradient pass (frozen weights): the memory-bound kernels between the data-gradient GEMMs.
 *     replaces: torch.autograd through the denoiser in lib/models/diffusions/gaussian_diffusion.py:193-216
 *       (guidance with grad_through_unet=True) and lib/models/autodecoders/diffusion_nerf.py:356-372 (val_optim:
 *       loss.backward() of the diffusion prior w.r.t. the code).  Convolution / linear data gradients reuse ssdnerf_gemm_f16 on
 *       transposed, tap-flipped weights.  Gradients are fp16 NHWC under a device-side loss scale.
 * ---------------------------------------------------------------------------------------------- */
typedef struct ssdnerf_gn_bwd_args {
    const void* x1; uint32_t C1;           /* raw GroupNorm input, source 1: fp16 [B][HW][C1] */
    const void* x2; uint32_t C2;           /* source 2 of a channel concat or NULL */
    uint32_t B, HW, groups;
    const float* stats; const float* stats2; int quad_stats;   /* forward statistics as for ssdnerf_gn_apply / ssdnerf_gn_apply_q */
    const float* gamma; const float* beta; /* [C1+C2] */
    const float* scale_shift; long long ss_batch_stride;       /* NormWithEmbedding (1 + scale, shift) rows or NULL */
    float eps; int do_silu;
    const void* dy;                        /* gradient w.r.t. the (SiLU'd) normalised output, fp16 [B][HW][C1+C2] */
    const void* add;                       /* optional gradient added to the result (shortcut / residual), fp16 [B][HW][C1+C2] */
    float* group_sums;                     /* scratch [B][groups][2] */
    void* dx1; void* dx2;                  /* out: fp16 [B][HW][C1], [B][HW][C2] */
    float* channel_sums;                   /* optional out [B][C1+C2][2]: sum_p dy', sum_p dy' * xhat with dy' = d loss / d (xhat*g' + b')
                                              (after the SiLU derivative): d gamma, d beta, d scale, d shift are linear in them */
} ssdnerf_gn_bwd_args;
SSDNERF_API int ssdnerf_gn_bwd(const ssdnerf_gn_bwd_args* args, void* stream);
/* dS = P * (dP - rowsum(P * dP)): softmax backward over attention rows; P fp16, dP fp32, dS fp16, [rows][T] */
SSDNERF_API int ssdnerf_softmax_bwd_rows(const void* P, const float* dP, uint32_t rows, uint32_t T, void* dS, void* stream);
/* n2 x n1 batched fp16 transposes: dst[((b2*n1 + b1)*cols + c)*rows + r] = src[b2*stride2 + b1*stride1 + r*row_stride + c] (elements) */
SSDNERF_API int ssdnerf_transpose_f16(const void* src, void* dst, uint32_t rows, uint32_t cols, long long row_stride, long long stride1,
                                      uint32_t n1, long long stride2, uint32_t n2, void* stream);
/* data gradient of ssdnerf_im2col_s2 (+ optional add [B][H][W][C]): dcol fp16 [B][H/2][W/2][9C] -> dx fp16 [B][H][W][C] */
SSDNERF_API int ssdnerf_col2im_s2(const void* dcol, uint32_t B, uint32_t H, uint32_t W, uint32_t C, const void* add, void* dx, void* stream);
/* data gradient of ssdnerf_upsample2x: dup fp16 [B][2H][2W][C] -> dx fp16 [B][H][W][C] */
SSDNERF_API int ssdnerf_sum2x2(const void* dup, uint32_t B, uint32_t H, uint32_t W, uint32_t C, void* dx, void* stream);
/* dst += src over n fp16 elements (n % 8 == 0) */
SSDNERF_API int ssdnerf_add_f16(void* dst, const void* src, unsigned long long n, void* stream);
/* loss scale of an incoming gradient: scale[0] = target / max|g| (1 when g == 0), scale[1] = 1 / scale[0]; g fp32 [n] */
SSDNERF_API int ssdnerf_grad_scale(const float* g, unsigned long long n, float target, float* scale, void* stream);
/* g fp32 [B][C][H][W] * scale[0] -> fp16 [B][H][W][Cpad];  dx fp32 [B][H][W][Cpad] * scale[1] -> fp32 [B][C][H][W] */
SSDNERF_API int ssdnerf_grad_nchw_to_nhwc_f16(const float* g, uint32_t B, uint32_t C, uint32_t H, uint32_t W, uint32_t Cpad, const float* scale,
                                              void* out, void* stream);
SSDNERF_API int ssdnerf_grad_nhwc_to_nchw_f32(const float* dx, uint32_t B, uint32_t C, uint32_t H, uint32_t W, uint32_t Cpad, const float* scale,
                                              float* out, void* stream);


/* ------------------------------------------------------------------------------------------------
 * 4c. Weight gradients of the UNet's convolutions / linear layers (training of the denoiser).
 *     replaces: autograd of mmgen's conv / linear modules under lib/models/autodecoders/diffusion_nerf.py:113-121
 *               (loss_diffusion.backward(); optimizer['diffusion'].step())
 *     dw[co][tap][dw_c0 + ci] += sum over (b, y, x) of gy[b][y][x][gy_c0 + co] * X[b][y*stride + dy - 1][x*stride + dx - 1][x_c0 + ci]
 *     (tap = (dy+1)*3 + (dx+1); taps == 1: no offset / padding; up: X is read through a nearest x2 upsample).  fp16 operands
 *     (mma.sync tensor cores, tiles transposed by ldmatrix.trans), fp32 accumulation, split over the pixel axis; dw is ACCUMULATED.
 * ---------------------------------------------------------------------------------------------- */
typedef struct ssdnerf_wgrad_args {
    const void* gy; uint32_t gy_stride, gy_c0;     /* fp16 [batch*out_h*out_w][gy_stride] */
    const void* x; uint32_t x_stride, x_c0;        /* fp16 [batch][in_h][in_w][x_stride] */
    float* dw; uint32_t dw_stride, dw_c0;          /* fp32 [cout][taps][dw_stride] */
    uint32_t batch, out_h, out_w, in_h, in_w;
    uint32_t cout, cin;                            /* multiples of 64 */
    uint32_t taps;                                 /* 1 or 9 */
    uint32_t stride;                               /* 1 or 2 */
    int up;                                        /* 1: input is nearest-upsampled x2 before the convolution */
    uint32_t ksplit;                               /* 0 = choose */
} ssdnerf_wgrad_args;
SSDNERF_API int ssdnerf_conv_wgrad_f16(const ssdnerf_wgrad_args* args, void* stream);
/* out[c] += sum_r src[r][c0 + c]   (bias gradients); src fp16 [rows][stride], channels % 8 == 0 */
/* in-place inverted dropout on fp16 data with a counter-based mask (same seed -> same mask: the backward re-applies it to the gradient);
 * replaces nn.Dropout of DenoisingResBlock (mmgen, used by lib/models/architecture/diffusion/modules.py:51-110) in training */
SSDNERF_API int ssdnerf_dropout_f16(void* x, unsigned long long n, unsigned long long seed, float p_drop, void* stream);
SSDNERF_API int ssdnerf_colsum_f16(const void* src, uint64_t rows, uint32_t stride, uint32_t c0, uint32_t channels, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SSDNERF_B200_H_ */
