// Weight gradients of the denoising UNet's convolutions / linear layers (training: lib/models/autodecoders/diffusion_nerf.py:66-189,
// `loss_diffusion.backward()` + `optimizer['diffusion'].step()`).
//
//   dW[co][tap][ci] = sum over output pixels (b, y, x) of  gY[b][y][x][co] * X[b][y*s + dy - 1][x*s + dx - 1][ci]
//
// i.e. a GEMM  dW = gY^T . im2col(X)  whose reduction dimension is the PIXEL axis (B*H*W = 262144 at the 128x128 level) and whose
// output is tiny (Cout x 9 Cin).  Both operands live in HBM as NHWC fp16 with the reduction index as the ROW index, the opposite
// of what the forward / data-gradient implicit GEMM (gemm_tc.cu) wants, so instead of transposing gigabytes through HBM this kernel
// stages [64 pixels][64 channels] tiles of both tensors with cp.async exactly as they lie in memory and lets ldmatrix.trans do
// the transposition on the way into the mma.sync fragments.  The 3x3 taps, the stride-2 convolution of the downsample and the
// nearest-x2 upsample in front of the up-convolution are address arithmetic of the X tile loader (zero fill outside the image).
// Split-K over the pixel axis fills the SMs; partial sums are added to the fp32 output with red.global.add.f32.
#include "common.cuh"
#include "../../include/ssdnerf_b200.h"

namespace ssdnerf {

struct WgradParams {
    const __half* gy; uint32_t gy_stride, gy_c0;     // [P][gy_stride], this launch uses channels [gy_c0, gy_c0 + Cout)
    const __half* x; uint32_t x_stride, x_c0;        // [B][Hi][Wi][x_stride], channels [x_c0, x_c0 + Cin)
    float* dw; uint32_t dw_stride, dw_c0;            // [Cout][taps][dw_stride], written at input-channel offset dw_c0
    uint32_t B, Ho, Wo, Hi, Wi, Cout, Cin, taps, stride, up, ksplit;
};

constexpr int kWgTile = 64;          // pixels per stage, output channels and input channels per CTA
constexpr int kWgRow = 72;           // padded smem row (halves): 144-byte pitch keeps ldmatrix conflict-free and 16-byte aligned

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, uint32_t src_bytes) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(s), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(s));
}
__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(128) k_wgrad_f16(const WgradParams p) {
    __shared__ __align__(16) __half sG[2][kWgTile][kWgRow];
    __shared__ __align__(16) __half sX[2][kWgTile][kWgRow];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t co0 = blockIdx.x * kWgTile, ci0 = blockIdx.y * kWgTile;
    const uint32_t tap = blockIdx.z % p.taps, split = blockIdx.z / p.taps;
    const int dy = p.taps == 9 ? (int)(tap / 3) - 1 : 0, dx = p.taps == 9 ? (int)(tap % 3) - 1 : 0;
    const uint32_t HWo = p.Ho * p.Wo, P = p.B * HWo;
    const uint32_t chunks = P / kWgTile;
    const uint32_t c_begin = (uint32_t)((uint64_t)chunks * split / p.ksplit), c_end = (uint32_t)((uint64_t)chunks * (split + 1) / p.ksplit);
    const int Hv = p.up ? (int)p.Hi * 2 : (int)p.Hi, Wv = p.up ? (int)p.Wi * 2 : (int)p.Wi;

    // this thread's 4 (row, 16-byte segment) slots of each tile
    const int seg = tid & 7, row0 = tid >> 3;                       // rows row0 + 16 i
    auto issue = [&](uint32_t chunk, int buf) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = row0 + 16 * i;
            const uint32_t pix = chunk * kWgTile + r;
            cp_async16(&sG[buf][r][seg * 8], p.gy + (size_t)pix * p.gy_stride + p.gy_c0 + co0 + seg * 8, 16);
            const uint32_t b = pix / HWo, rem = pix - b * HWo;
            const uint32_t y = rem / p.Wo, x = rem - y * p.Wo;
            int sy = (int)(y * p.stride) + dy, sx = (int)(x * p.stride) + dx;
            const bool ok = sy >= 0 && sy < Hv && sx >= 0 && sx < Wv;
            if (p.up) { sy >>= 1; sx >>= 1; }
            const __half* src = ok ? p.x + (((size_t)b * p.Hi + sy) * p.Wi + sx) * p.x_stride + p.x_c0 + ci0 + seg * 8 : p.x;
            cp_async16(&sX[buf][r][seg * 8], src, ok ? 16u : 0u);
        }
        cp_async_commit();
    };

    float acc[2][4][4];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[a][b][c] = 0.0f;

    const int wm = (warp & 1) * 32, wn = (warp >> 1) * 32;        // warp tile: 32 output channels x 32 input channels
    const int lj = lane >> 3, lr = lane & 7;                       // ldmatrix: lane supplies row lr of matrix lj
    if (c_begin < c_end) issue(c_begin, 0);
    for (uint32_t c = c_begin; c < c_end; ++c) {
        const int buf = (int)((c - c_begin) & 1);
        if (c + 1 < c_end) { issue(c + 1, buf ^ 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
        __syncthreads();
#pragma unroll
        for (int ks = 0; ks < kWgTile / 16; ++ks) {
            uint32_t a[2][4], bfr[2][4];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)       // A = gY^T: matrices (m 0-7,k 0-7) (m 8-15,k 0-7) (m 0-7,k 8-15) (m 8-15,k 8-15); stored rows are k
                ldmatrix_x4_trans(a[mt], &sG[buf][ks * 16 + (lj >> 1) * 8 + lr][wm + mt * 16 + (lj & 1) * 8]);
#pragma unroll
            for (int np = 0; np < 2; ++np)       // B = X: matrices (k 0-7,n 0-7) (k 8-15,n 0-7) (k 0-7,n 8-15) (k 8-15,n 8-15)
                ldmatrix_x4_trans(bfr[np], &sX[buf][ks * 16 + (lj & 1) * 8 + lr][wn + np * 16 + (lj >> 1) * 8]);
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) mma_16816(acc[mt][nt], a[mt], bfr[nt >> 1][(nt & 1) * 2], bfr[nt >> 1][(nt & 1) * 2 + 1]);
        }
        __syncthreads();
    }
    // accumulator (row g / g + 8, column pair 2 (lane % 4)) -> dW[co][tap][ci]
    const int g = lane >> 2, q = (lane & 3) * 2;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint32_t co = co0 + wm + mt * 16 + g + h * 8, ci = ci0 + wn + nt * 8 + q;
                float* dst = p.dw + ((size_t)co * p.taps + tap) * p.dw_stride + p.dw_c0 + ci;
                atomicAdd(dst, acc[mt][nt][2 * h]);
                atomicAdd(dst + 1, acc[mt][nt][2 * h + 1]);
            }
}

// out[c] += sum over rows of src[r][c0 + c]   (bias gradients: column sums of gY);  grid.x = row chunks, blockDim.x = (C / 8) * pr
__global__ void __launch_bounds__(256) k_colsum_f16(const __half* __restrict__ src, uint32_t rows, uint32_t stride, uint32_t c0, uint32_t C,
                                                    uint32_t rows_per_block, float* __restrict__ out) {
    extern __shared__ float s_sum[];                // [C]
    const uint32_t cv = C / 8, v = threadIdx.x % cv, lane_r = threadIdx.x / cv, rstep = blockDim.x / cv;
    for (uint32_t i = threadIdx.x; i < C; i += blockDim.x) s_sum[i] = 0.0f;
    __syncthreads();
    float a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = 0.0f;
    const uint32_t r0 = blockIdx.x * rows_per_block, r1 = min(r0 + rows_per_block, rows);
    for (uint32_t r = r0 + lane_r; r < r1; r += rstep) {
        const uint4 t = __ldg(reinterpret_cast<const uint4*>(src + (size_t)r * stride + c0 + v * 8));
        const __half2* h = reinterpret_cast<const __half2*>(&t);
#pragma unroll
        for (int i = 0; i < 4; ++i) { const float2 f = __half22float2(h[i]); a[2 * i] += f.x; a[2 * i + 1] += f.y; }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) atomicAdd(&s_sum[v * 8 + k], a[k]);
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(out + i, s_sum[i]);
}

// Dropout of the ResBlocks' second convolution input (modules.py:84-90, nn.Dropout between SiLU and the conv), in place on fp16
// data.  The keep mask is a pure function of (seed, element index), so the backward and the weight-gradient pass regenerate it
// instead of storing it: 16 random bits per element from two splitmix64 rounds per 8-element vector.
__device__ __forceinline__ uint64_t splitmix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__global__ void k_dropout_f16(__half* __restrict__ x, size_t nvec, uint64_t seed, uint32_t thresh16, float scale) {
    const size_t i = threadIdx.x + (size_t)blockIdx.x * blockDim.x;
    if (i >= nvec) return;
    uint4 v = reinterpret_cast<uint4*>(x)[i];
    __half2* h = reinterpret_cast<__half2*>(&v);
    const uint64_t r0 = splitmix64(seed ^ (2 * i)), r1 = splitmix64(seed ^ (2 * i + 1));
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint64_t r = k < 2 ? r0 : r1;
        const uint32_t a = (uint32_t)(r >> (32 * (k & 1))) & 0xffffu, b = (uint32_t)(r >> (32 * (k & 1) + 16)) & 0xffffu;
        const float2 f = __half22float2(h[k]);
        h[k] = __floats2half2_rn(a >= thresh16 ? f.x * scale : 0.0f, b >= thresh16 ? f.y * scale : 0.0f);
    }
    reinterpret_cast<uint4*>(x)[i] = v;
}

}  // namespace ssdnerf

using namespace ssdnerf;

extern "C" {

int ssdnerf_conv_wgrad_f16(const ssdnerf_wgrad_args* a, void* stream) {
    if (!a) return set_error_msg(SSDNERF_ERR_ARG, "conv_wgrad: args is NULL");
    if (!a->gy || !a->x || !a->dw) return set_error_msg(SSDNERF_ERR_ARG, "conv_wgrad: gy, x and dw are required");
    if (a->cout == 0 || a->cin == 0 || a->cout % kWgTile || a->cin % kWgTile)
        return set_error_msg(SSDNERF_ERR_ARG, "conv_wgrad: cout and cin must be non-zero multiples of 64");
    if (a->taps != 1 && a->taps != 9) return set_error_msg(SSDNERF_ERR_ARG, "conv_wgrad: taps must be 1 or 9");
    if (a->stride != 1 && a->stride != 2) return set_error_msg(SSDNERF_ERR_ARG, "conv_wgrad: stride must be 1 or 2");
    if (a->up && a->stride != 1) return set_error_msg(SSDNERF_ERR_ARG, "conv_wgrad: upsampled input implies stride 1");
    const uint64_t P = (uint64_t)a->batch * a->out_h * a->out_w;
    if (P == 0 || P % kWgTile) return set_error_msg(SSDNERF_ERR_ARG, "conv_wgrad: batch * out_h * out_w must be a non-zero multiple of 64");
    if ((a->gy_stride | a->gy_c0 | a->x_stride | a->x_c0) % 8 || ((uintptr_t)a->gy & 15u) || ((uintptr_t)a->x & 15u))
        return set_error_msg(SSDNERF_ERR_ARG, "conv_wgrad: fp16 tensors must be 16-byte aligned with channel strides / offsets in multiples of 8");
    if (a->gy_c0 + a->cout > a->gy_stride || a->x_c0 + a->cin > a->x_stride || a->dw_c0 + a->cin > a->dw_stride)
        return set_error_msg(SSDNERF_ERR_ARG, "conv_wgrad: channel window exceeds the row stride");
    const uint32_t vh = a->up ? a->in_h * 2 : a->in_h, vw = a->up ? a->in_w * 2 : a->in_w;
    if (a->out_h * a->stride != vh || a->out_w * a->stride != vw)
        return set_error_msg(SSDNERF_ERR_ARG, "conv_wgrad: output size * stride must equal the (upsampled) input size");
    WgradParams p{};
    p.gy = (const __half*)a->gy; p.gy_stride = a->gy_stride; p.gy_c0 = a->gy_c0;
    p.x = (const __half*)a->x; p.x_stride = a->x_stride; p.x_c0 = a->x_c0;
    p.dw = a->dw; p.dw_stride = a->dw_stride; p.dw_c0 = a->dw_c0;
    p.B = a->batch; p.Ho = a->out_h; p.Wo = a->out_w; p.Hi = a->in_h; p.Wi = a->in_w;
    p.Cout = a->cout; p.Cin = a->cin; p.taps = a->taps; p.stride = a->stride; p.up = a->up ? 1 : 0;
    int dev = 0, sms = 0;
    SSDNERF_CUDA_OK(cudaGetDevice(&dev));
    SSDNERF_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    // split the pixel axis until the grid covers ~4 CTAs per SM (each split keeps at least 8 stages of work)
    const uint32_t base = (a->cout / kWgTile) * (a->cin / kWgTile) * a->taps, chunks = (uint32_t)(P / kWgTile);
    uint32_t ksplit = a->ksplit;
    if (ksplit == 0) {
        ksplit = 1;
        while (base * ksplit < (uint32_t)sms * 4 && chunks / (ksplit * 2) >= 8) ksplit *= 2;
    }
    if (ksplit > chunks) ksplit = chunks;
    if ((uint64_t)a->taps * ksplit > 65535) return set_error_msg(SSDNERF_ERR_ARG, "conv_wgrad: ksplit too large");
    p.ksplit = ksplit;
    dim3 grid(a->cout / kWgTile, a->cin / kWgTile, a->taps * ksplit);
    k_wgrad_f16<<<grid, 128, 0, (cudaStream_t)stream>>>(p);
    SSDNERF_LAUNCH_OK();
    return 0;
}

int ssdnerf_colsum_f16(const void* src, uint64_t rows, uint32_t stride, uint32_t c0, uint32_t channels, float* out, void* stream) {
    if (rows == 0 || channels == 0) return 0;
    if (!src || !out) return set_error_msg(SSDNERF_ERR_ARG, "colsum: NULL buffer");
    if (channels % 8 || channels > 2048 || (stride | c0) % 8 || ((uintptr_t)src & 15u))
        return set_error_msg(SSDNERF_ERR_ARG, "colsum: channels must be a multiple of 8 (<= 2048), stride / offset multiples of 8, src 16-byte aligned");
    if (rows > 0xffffffffull) return set_error_msg(SSDNERF_ERR_ARG, "colsum: too many rows");
    const uint32_t cv = channels / 8;
    uint32_t threads = 256;
    if (cv > threads) threads = cv;                                   // cv <= 256
    threads = threads / cv * cv;
    const uint32_t rpb = 512;
    k_colsum_f16<<<(uint32_t)((rows + rpb - 1) / rpb), threads, channels * sizeof(float), (cudaStream_t)stream>>>(
        (const __half*)src, (uint32_t)rows, stride, c0, channels, rpb, out);
    SSDNERF_LAUNCH_OK();
    return 0;
}

int ssdnerf_dropout_f16(void* x, unsigned long long n, unsigned long long seed, float p_drop, void* stream) {
    if (n == 0 || p_drop <= 0.0f) return 0;
    if (!x || ((uintptr_t)x & 15u) || n % 8) return set_error_msg(SSDNERF_ERR_ARG, "dropout: x must be 16-byte aligned with n % 8 == 0");
    if (!(p_drop < 1.0f)) return set_error_msg(SSDNERF_ERR_ARG, "dropout: p must be in [0, 1)");
    const uint32_t thresh = (uint32_t)lrintf(p_drop * 65536.0f);
    const float scale = 65536.0f / (float)(65536u - thresh);
    const size_t nvec = (size_t)(n / 8);
    k_dropout_f16<<<(uint32_t)((nvec + 255) / 256), 256, 0, (cudaStream_t)stream>>>((__half*)x, nvec, (uint64_t)seed, thresh, scale);
    SSDNERF_LAUNCH_OK();
    return 0;
}

}  // extern "C"
