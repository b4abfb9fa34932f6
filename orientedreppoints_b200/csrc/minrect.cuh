// minrect.cuh - device routine shared by minarearect.cu and head_post.cu: one 9-point set -> hull ->
// minimum-area rectangle, the arithmetic of mmdet/ops/minarearect/src/minarearect_kernel.cu:52-452
// (see oracle/oracle_minarearect.c for the three documented deviations).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace orp {
namespace mr {

struct P2f { float x, y; };

static __device__ __forceinline__ int sgn8(float d) { return (int)(d > 1E-8f) - (int)(d < -1E-8f); }
static __device__ __forceinline__ bool near_pt(P2f a, P2f b)
{
    return sgn8(__fsub_rn(a.x, b.x)) == 0 && sgn8(__fsub_rn(a.y, b.y)) == 0;
}
static __device__ __forceinline__ float sqdist(P2f a, P2f b)
{
    const float dx = __fsub_rn(a.x, b.x), dy = __fsub_rn(a.y, b.y);
    return __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
}
static __device__ __forceinline__ float cosr(float a) { return (float)cos((double)a); }
static __device__ __forceinline__ double turn(P2f base, P2f cand, P2f cur)
{
    const double l = __dmul_rn(__dsub_rn((double)cand.x, (double)base.x), __dsub_rn((double)cur.y, (double)base.y));
    const double r = __dmul_rn(__dsub_rn((double)cur.x, (double)base.x), __dsub_rn((double)cand.y, (double)base.y));
    return __dsub_rn(l, r);
}

constexpr int kPts = 9;
constexpr int kCap = 12;

static __device__ __forceinline__ int wrap_chain(const P2f *P, P2f pmax, int imax, int dir, int *stack)
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

static __device__ __noinline__ void minrect_one(const float *in, float *out, int32_t *map)
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


}  // namespace mr
}  // namespace orp
