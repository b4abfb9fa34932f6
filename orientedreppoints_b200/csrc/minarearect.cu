// minarearect.cu - 9-point sets -> minimum-area rectangles on sm_100a.
//
// Replaces minareabbox_cuda (mmdet/ops/minarearect/src/minarearect_kernel.cu:470-505), which runs one
// thread per set on the legacy stream, then copies the result to the host, loops over it and uploads it
// again.  Here: one thread per set on the caller's stream, inputs staged through shared memory with
// coalesced 128-bit loads (72 B in / 32 B out per set = 104 B algorithmic, HBM-bound), the result stays
// on the device and the `* stride + centre` affine of orientedreppoints_head.py:748-749 can be fused in.
//
// Arithmetic: the reference's mixed precision, operation for operation, with the round-to-nearest
// intrinsics so nothing is contracted (see oracle/oracle_minarearect.c for the statement of the
// three deliberate differences: double-evaluated cos, no FMA, bounded gift-wrapping loops).
#include "common.cuh"

namespace orp {

namespace {

struct P2f { float x, y; };

__device__ __forceinline__ int sgn8(float d) { return (int)(d > 1E-8f) - (int)(d < -1E-8f); }
__device__ __forceinline__ bool near_pt(P2f a, P2f b)
{
    return sgn8(__fsub_rn(a.x, b.x)) == 0 && sgn8(__fsub_rn(a.y, b.y)) == 0;
}
__device__ __forceinline__ float sqdist(P2f a, P2f b)
{
    const float dx = __fsub_rn(a.x, b.x), dy = __fsub_rn(a.y, b.y);
    return __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
}
__device__ __forceinline__ float cosr(float a) { return (float)cos((double)a); }
__device__ __forceinline__ double turn(P2f base, P2f cand, P2f cur)
{
    const double l = __dmul_rn(__dsub_rn((double)cand.x, (double)base.x), __dsub_rn((double)cur.y, (double)base.y));
    const double r = __dmul_rn(__dsub_rn((double)cur.x, (double)base.x), __dsub_rn((double)cand.y, (double)base.y));
    return __dsub_rn(l, r);
}

constexpr int kPts = 9;
constexpr int kCap = 12;

__device__ __forceinline__ int wrap_chain(const P2f *P, P2f pmax, int imax, int dir, int *stack)
{
    int top = 0, k = 0;
    stack[0] = 0;
    while (k != imax && top < kCap - 1) {
        P2f pk = pmax;
        k = imax;
        const P2f base = P[stack[top]];
        for (int i = 1; i < kPts; ++i) {
            const double s = turn(base, P[i], pk);
            const bool take = dir > 0 ? (s > 0) : (s < 0);
            if (take || (s == 0 && sqdist(base, P[i]) > sqdist(base, pk))) { pk = P[i]; k = i; }
        }
        stack[++top] = k;
    }
    return top;
}

__device__ void minrect_one(const float *in, float *out, int32_t *map)
{
    const float pi_f = 3.1415926f;
    const float hp = pi_f / 2;
    P2f P[2 * kCap], in0[kPts];
#pragma unroll
    for (int i = 0; i < kPts; ++i) { P[i].x = in[2 * i]; P[i].y = in[2 * i + 1]; in0[i] = P[i]; }
    // lowest (y, then x) point swapped into slot 0 while scanning; highest tracked alongside
    P2f pmax = P[0];
    int imax = 0;
    for (int i = 0; i < kPts; ++i) {
        if (P[i].y < P[0].y || (P[i].y == P[0].y && P[i].x < P[0].x)) { P2f t = P[0]; P[0] = P[i]; P[i] = t; }
        if (i == 0) { pmax = P[0]; imax = 0; }
        if (P[i].y > pmax.y || (P[i].y == pmax.y && P[i].x > pmax.x)) { pmax = P[i]; imax = i; }
    }
    if (imax == 0) { imax = 1; pmax = P[1]; }
    int s1[kCap], s2[kCap];
    const int top1 = wrap_chain(P, pmax, imax, +1, s1);
    const int top2 = wrap_chain(P, pmax, imax, -1, s2);
    const int nh = top1 + top2;
    P2f ring[2 * kCap + 1];
    for (int i = 0; i < nh; ++i) ring[i] = (i <= top1) ? P[s1[i]] : P[s2[top2 - (i - top1)]];
    ring[nh] = ring[0];
    if (map) {
        for (int i = 0; i < kPts; ++i) {
            int found = -1;
            if (i < nh)
                for (int j = 0; j < kPts; ++j)
                    if (near_pt(ring[i], in0[j])) { found = j; break; }
            map[i] = found;
        }
    }
    const int m = nh + 1, ne = nh;
    float uniq[2 * kCap];
    int nu = 0;
    for (int i = 0; i < ne; ++i) {
        const float ex = __fsub_rn(ring[i + 1].x, ring[i].x), ey = __fsub_rn(ring[i + 1].y, ring[i].y);
        float a = (float)atan2((double)ey, (double)ex);
        if (a >= 0) {
            a = (float)fmod((double)a, (double)pi_f / 2);
        } else {
            const float q1 = __fsub_rn(__fdiv_rn(a, hp), 1.0f);
            const int k = (int)q1;
            a = __fsub_rn(a, __fmul_rn((float)k, hp));
        }
        bool seen = false;
        if (i > 0)
            for (int j = 0; j < nu; ++j) seen = seen || (a == uniq[j]);
        if (i == 0 || !seen) uniq[nu++] = a;
    }
    float minarea = 1e12f;
    float best_a = 0.f, bxmin = 0.f, bymin = 0.f, bxmax = 0.f, bymax = 0.f;
    for (int u = 0; u < nu; ++u) {
        const float a = uniq[u];
        const float r00 = cosr(a), r01 = cosr(__fsub_rn(a, hp)), r10 = cosr(__fadd_rn(a, hp)), r11 = r00;
        float xmin = 1e12f, ymin = 1e12f, xmax = -1e12f, ymax = -1e12f;
        for (int j = 0; j < m; ++j) {
            const float rx = __fadd_rn(__fadd_rn(0.0f, __fmul_rn(r00, ring[j].x)), __fmul_rn(r01, ring[j].y));
            const float ry = __fadd_rn(__fadd_rn(0.0f, __fmul_rn(r10, ring[j].x)), __fmul_rn(r11, ring[j].y));
            if (!(isinf(rx) || isnan(rx))) { if (rx < xmin) xmin = rx; if (rx > xmax) xmax = rx; }
            if (!(isinf(ry) || isnan(ry))) { if (ry < ymin) ymin = ry; if (ry > ymax) ymax = ry; }
        }
        const float area = __fmul_rn(__fsub_rn(xmax, xmin), __fsub_rn(ymax, ymin));
        if (area < minarea) { minarea = area; best_a = a; bxmin = xmin; bymin = ymin; bxmax = xmax; bymax = ymax; }
    }
    const float r00 = cosr(best_a), r01 = cosr(__fsub_rn(best_a, hp)), r10 = cosr(__fadd_rn(best_a, hp)), r11 = r00;
    const float cx[4] = {bxmax, bxmin, bxmin, bxmax};
    const float cy[4] = {bymin, bymin, bymax, bymax};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        out[2 * c] = __fadd_rn(__fadd_rn(0.0f, __fmul_rn(cx[c], r00)), __fmul_rn(cy[c], r10));
        out[2 * c + 1] = __fadd_rn(__fadd_rn(0.0f, __fmul_rn(cx[c], r01)), __fmul_rn(cy[c], r11));
    }
}

constexpr int kThreads = 128;

__global__ void __launch_bounds__(kThreads)
minarearect_kernel(const float *__restrict__ pts, int n, float *__restrict__ out,
                   int32_t *__restrict__ hull_map, float scale, const float *__restrict__ center)
{
    __shared__ float s_in[kThreads * 18];
    __shared__ float s_out[kThreads * 8];
    const int base = blockIdx.x * kThreads;
    const int cnt = min(kThreads, n - base);
    // coalesced staging: the block's rows are one contiguous span of cnt*18 floats
    {
        const float *src = pts + (size_t)base * 18;
        const int total = cnt * 18;
        const bool al16 = ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
        if (al16) {
            const int nv = total >> 2;
            const float4 *s4 = reinterpret_cast<const float4 *>(src);
            float4 *d4 = reinterpret_cast<float4 *>(s_in);
            for (int t = threadIdx.x; t < nv; t += kThreads) d4[t] = s4[t];
            for (int t = (nv << 2) + threadIdx.x; t < total; t += kThreads) s_in[t] = src[t];
        } else {
            for (int t = threadIdx.x; t < total; t += kThreads) s_in[t] = src[t];
        }
    }
    __syncthreads();
    if (threadIdx.x < cnt) {
        float in[18], o[8];
        int32_t map[9];
#pragma unroll
        for (int k = 0; k < 18; ++k) in[k] = s_in[threadIdx.x * 18 + k];   // stride 18: conflict-free (gcd(18,32)=2 -> 2-way)
        minrect_one(in, o, hull_map ? map : nullptr);
        if (center) {
            const float cxv = center[(size_t)(base + threadIdx.x) * 2], cyv = center[(size_t)(base + threadIdx.x) * 2 + 1];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                o[2 * c] = __fadd_rn(__fmul_rn(o[2 * c], scale), cxv);
                o[2 * c + 1] = __fadd_rn(__fmul_rn(o[2 * c + 1], scale), cyv);
            }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) s_out[threadIdx.x * 8 + k] = o[k];
        if (hull_map)
#pragma unroll
            for (int k = 0; k < 9; ++k) hull_map[(size_t)(base + threadIdx.x) * 9 + k] = map[k];
    }
    __syncthreads();
    {
        float *dst = out + (size_t)base * 8;   // 32 B per row: always 16 B aligned if `out` is
        const int total = cnt * 8;
        if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
            float4 *d4 = reinterpret_cast<float4 *>(dst);
            const float4 *s4 = reinterpret_cast<const float4 *>(s_out);
            for (int t = threadIdx.x; t < (total >> 2); t += kThreads) d4[t] = s4[t];
        } else {
            for (int t = threadIdx.x; t < total; t += kThreads) dst[t] = s_out[t];
        }
    }
}

}  // namespace
}  // namespace orp

extern "C" int orp_minarearect(const float *pts, int n, float *out, int32_t *hull_map, float scale,
                               const float *center, void *stream)
{
    using namespace orp;
    if (n < 0 || (n > 0 && (!pts || !out))) return fail(ORP_EINVAL, "orp_minarearect: bad arguments");
    int rc = ensure_device();
    if (rc) return rc;
    if (n == 0) return ORP_OK;
    minarearect_kernel<<<ceil_div(n, kThreads), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        pts, n, out, hull_map, scale, center);
    ORP_LAUNCHED();
    return ORP_OK;
}
