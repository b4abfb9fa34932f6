// dense_bf16_misc.cu - the memory-bound companions of the tensor-core convolutions (NHWC bf16):
// stem im2col (7x7/2, 3 channels -> one 192-wide K block row per output pixel), 3x3/2 max-pool,
// GroupNorm statistics and apply (+ReLU, + the FPN top-down nearest-2x add).  All HBM-bound: 16-byte
// vector accesses, grids sized in multiples of the SM count.
#include <cuda_bf16.h>

#include "common.cuh"

namespace orp {
namespace {

__device__ __forceinline__ uint32_t pack2(float a, float b)
{
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&v);
}
__device__ __forceinline__ float2 unpack2(uint32_t u)
{
    return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&u));
}

// img: NCHW fp32 [N,3,H,W] (the layout the reference feeds its backbone) -> bf16 [N,Ho,Wo,192]
// with k = (kh*7 + kw)*3 + c for k < 147 and zeros above: conv1 (resnet.py:495) becomes a 1x1
// convolution over 192 channels on the tensor cores.
// One block = 64 consecutive output pixels of one output row: the 7 x 133 x 3 input patch is staged in shared
// memory with coalesced loads, then every thread emits whole 16-byte chunks (8 k-values) - consecutive threads
// write consecutive chunks of the same 384-byte row, so both sides of the kernel move full cache lines.
constexpr int kStemPix = 64;
__global__ void __launch_bounds__(256)
stem_im2col_kernel(const float *__restrict__ img, int N, int H, int W, int Ho, int Wo, __nv_bfloat16 *__restrict__ out)
{
    constexpr int PW = 2 * kStemPix + 5;                 // 133 input columns
    __shared__ float s_p[3 * 7 * PW];
    __shared__ __align__(16) int s_koff[192];
    const int t = threadIdx.x;
    const int wblocks = (Wo + kStemPix - 1) / kStemPix;
    const int wb = blockIdx.x % wblocks;
    const int oh = (blockIdx.x / wblocks) % Ho, n = blockIdx.x / (wblocks * Ho);
    const int ow0 = wb * kStemPix, x0 = ow0 * 2 - 3, y0 = oh * 2 - 3;
    if (t < 192) {
        const int tap = t / 3, c = t - tap * 3, kh = tap / 7, kw = tap - kh * 7;
        s_koff[t] = t < 147 ? (c * 7 + kh) * PW + kw : -1;
    }
#pragma unroll
    for (int r = 0; r < 21; ++r) {                        // (channel, patch row): warp-uniform decode, coalesced along x
        const int c = r / 7, py = r - c * 7, yy = y0 + py;
        const bool rok = yy >= 0 && yy < H;
        const float *src = img + (((size_t)n * 3 + c) * H + (rok ? yy : 0)) * W;
        if (t < PW) {
            const int xx = x0 + t;
            s_p[r * PW + t] = (rok && xx >= 0 && xx < W) ? __ldg(src + xx) : 0.f;
        }
    }
    __syncthreads();
    const int npix = min(kStemPix, Wo - ow0);
    __nv_bfloat16 *orow = out + (((size_t)n * Ho + oh) * Wo + ow0) * 192;
    for (int item = t; item < npix * 24; item += 256) {
        const int px = item / 24, ck = item - px * 24;
        const int4 o0 = *reinterpret_cast<const int4 *>(&s_koff[ck * 8]), o1 = *reinterpret_cast<const int4 *>(&s_koff[ck * 8 + 4]);
        const int off[8] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w};
        const float *pb = s_p + px * 2;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = off[j] >= 0 ? pb[off[j]] : 0.f;
        *reinterpret_cast<uint4 *>(orow + (size_t)item * 8) = make_uint4(pack2(v[0], v[1]), pack2(v[2], v[3]), pack2(v[4], v[5]), pack2(v[6], v[7]));
    }
}

__global__ void __launch_bounds__(256)
maxpool3x3s2_bf16_kernel(const __nv_bfloat16 *__restrict__ x, int N, int H, int W, int C, int Ho, int Wo,
                         __nv_bfloat16 *__restrict__ y)
{
    const int c8 = C / 8;
    const size_t total = (size_t)N * Ho * Wo * c8;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % c8) * 8;
        const size_t pix = i / c8;
        const int ow = (int)(pix % Wo), oh = (int)((pix / Wo) % Ho), n = (int)(pix / ((size_t)Wo * Ho));
        float m[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
#pragma unroll
        for (int dh = 0; dh < 3; ++dh)
#pragma unroll
            for (int dw = 0; dw < 3; ++dw) {
                const int ih = oh * 2 - 1 + dh, iw = ow * 2 - 1 + dw;
                if (ih < 0 || ih >= H || iw < 0 || iw >= W) continue;
                const uint4 u = *reinterpret_cast<const uint4 *>(x + (((size_t)n * H + ih) * W + iw) * C + c);
                const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float2 f = unpack2(uu[k]);
                    m[2 * k] = fmaxf(m[2 * k], f.x);
                    m[2 * k + 1] = fmaxf(m[2 * k + 1], f.y);
                }
            }
        *reinterpret_cast<uint4 *>(y + pix * C + c) = make_uint4(pack2(m[0], m[1]), pack2(m[2], m[3]), pack2(m[4], m[5]), pack2(m[6], m[7]));
    }
}

// GroupNorm statistics of a bf16 NHWC tensor with C = 256, 32 groups (8 channels = one 16-byte vector
// = one lane): grid (slabs, N); every warp strides over the pixels of its slab, lane l owns group l.
__global__ void __launch_bounds__(256)
gn_stats_bf16_kernel(const __nv_bfloat16 *__restrict__ x, int HW, int slab, double *__restrict__ stats)
{
    __shared__ float s_sum[8][32], s_sq[8][32];
    const int n = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int p0 = blockIdx.x * slab, p1 = min(HW, p0 + slab);
    const __nv_bfloat16 *base = x + (size_t)n * HW * 256 + lane * 8;
    float s = 0.f, q = 0.f;
    for (int p = p0 + warp; p < p1; p += 8) {
        const uint4 u = *reinterpret_cast<const uint4 *>(base + (size_t)p * 256);
        const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float2 f = unpack2(uu[k]);
            s += f.x + f.y;
            q += f.x * f.x + f.y * f.y;
        }
    }
    s_sum[warp][lane] = s;
    s_sq[warp][lane] = q;
    __syncthreads();
    if (warp == 0) {
        double ds = 0, dq = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) { ds += (double)s_sum[w][lane]; dq += (double)s_sq[w][lane]; }
        atomicAdd(&stats[((size_t)n * 32 + lane) * 2], ds);
        atomicAdd(&stats[((size_t)n * 32 + lane) * 2 + 1], dq);
    }
}

__global__ void __launch_bounds__(256)
gn_apply_bf16_kernel(const __nv_bfloat16 *__restrict__ x, int N, int H, int W, const double *__restrict__ stats,
                     const float *__restrict__ gamma, const float *__restrict__ beta, float eps, int relu,
                     const __nv_bfloat16 *__restrict__ up, __nv_bfloat16 *__restrict__ y)
{
    const size_t total = (size_t)N * H * W * 32;
    const double cnt = (double)H * W * 8;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int g = (int)(i & 31);
        const size_t pix = i >> 5;
        const int n = (int)(pix / ((size_t)H * W));
        const double sm = stats[((size_t)n * 32 + g) * 2], sq = stats[((size_t)n * 32 + g) * 2 + 1];
        const double mean = sm / cnt;
        double var = sq / cnt - mean * mean;
        var = var < 0 ? 0 : var;
        const float rstd = (float)(1.0 / sqrt(var + (double)eps)), mu = (float)mean;
        const uint4 u = reinterpret_cast<const uint4 *>(x)[i];
        const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
        float o[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float2 f = unpack2(uu[k]);
            o[2 * k] = f.x;
            o[2 * k + 1] = f.y;
        }
        const float4 g0 = *reinterpret_cast<const float4 *>(gamma + g * 8), g1 = *reinterpret_cast<const float4 *>(gamma + g * 8 + 4);
        const float4 b0 = *reinterpret_cast<const float4 *>(beta + g * 8), b1 = *reinterpret_cast<const float4 *>(beta + g * 8 + 4);
        const float ga[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const float be[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            o[j] = (o[j] - mu) * rstd * ga[j] + be[j];
            if (relu) o[j] = fmaxf(o[j], 0.f);
        }
        if (up) {
            const int hw = (int)(pix % ((size_t)H * W));
            const int h = hw / W, w = hw - h * W;
            const int Hu = (H + 1) / 2, Wu = (W + 1) / 2;              // F.interpolate(size=prev_shape, mode='nearest'): src = floor(dst * in / out)
            const uint4 v = *reinterpret_cast<const uint4 *>(up + (((size_t)n * Hu + (h * Hu) / H) * Wu + (w * Wu) / W) * 256 + g * 8);
            const uint32_t vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float2 f = unpack2(vv[k]);
                o[2 * k] += f.x;
                o[2 * k + 1] += f.y;
            }
        }
        reinterpret_cast<uint4 *>(y)[i] = make_uint4(pack2(o[0], o[1]), pack2(o[2], o[3]), pack2(o[4], o[5]), pack2(o[6], o[7]));
    }
}

int grid_for(size_t items, int threads)
{
    size_t g = (items + threads - 1) / threads;
    const size_t cap = 148 * 16;
    return (int)(g < cap ? (g ? g : 1) : cap);
}

}  // namespace
}  // namespace orp

using namespace orp;

extern "C" int orp_stem_im2col_bf16(const float *img_nchw, int N, int H, int W, void *out, void *stream)
{
    if (!img_nchw || !out || N <= 0) return fail(ORP_EINVAL, "stem_im2col_bf16: bad arguments");
    int rc = ensure_device();
    if (rc) return rc;
    const int Ho = (H + 6 - 7) / 2 + 1, Wo = (W + 6 - 7) / 2 + 1;
    const int wblocks = (Wo + kStemPix - 1) / kStemPix;
    stem_im2col_kernel<<<(unsigned)((size_t)N * Ho * wblocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        img_nchw, N, H, W, Ho, Wo, static_cast<__nv_bfloat16 *>(out));
    ORP_LAUNCHED();
    return ORP_OK;
}

extern "C" int orp_maxpool3x3s2_bf16(const void *x, int N, int H, int W, int C, void *y, void *stream)
{
    if (!x || !y || C % 8) return fail(ORP_EINVAL, "maxpool3x3s2_bf16: bad arguments");
    int rc = ensure_device();
    if (rc) return rc;
    const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
    const size_t total = (size_t)N * Ho * Wo * (C / 8);
    maxpool3x3s2_bf16_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16 *>(x), N, H, W, C, Ho, Wo, static_cast<__nv_bfloat16 *>(y));
    ORP_LAUNCHED();
    return ORP_OK;
}

extern "C" int orp_gn_stats_bf16(const void *x, int N, int HW, int C, int groups, double *stats, void *stream)
{
    if (!x || !stats || C != 256 || groups != 32) return fail(ORP_EINVAL, "gn_stats_bf16: needs C=256, 32 groups");
    int rc = ensure_device();
    if (rc) return rc;
    int slabs = ceil_div(HW, 64);
    const int maxs = (148 * 4 + N - 1) / N;
    if (slabs > maxs) slabs = maxs;
    const int slab = ceil_div(HW, slabs);
    slabs = ceil_div(HW, slab);
    gn_stats_bf16_kernel<<<dim3(slabs, N), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16 *>(x), HW, slab, stats);
    ORP_LAUNCHED();
    return ORP_OK;
}

extern "C" int orp_gn_apply_bf16(const void *x, int N, int H, int W, int C, const double *stats, int groups,
                                 const float *gamma, const float *beta, float eps, int relu, const void *up_src, void *y,
                                 void *stream)
{
    if (!x || !y || !stats || !gamma || !beta || C != 256 || groups != 32) return fail(ORP_EINVAL, "gn_apply_bf16: needs C=256, 32 groups");
    int rc = ensure_device();
    if (rc) return rc;
    const size_t total = (size_t)N * H * W * 32;
    gn_apply_bf16_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16 *>(x), N, H, W, stats, gamma, beta, eps, relu,
        static_cast<const __nv_bfloat16 *>(up_src), static_cast<__nv_bfloat16 *>(y));
    ORP_LAUNCHED();
    return ORP_OK;
}
