/*
 * TEST INFRASTRUCTURE ONLY (oracle).  Never imported by the product path.
 *
 * CPU restatement of mmdet/ops/minarearect/src/minarearect_kernel.cu:52-452
 * (minBoundingRect :52-211, Jarvis_and_index :215-341, Findminbox :343-452): 9 points ->
 * gift-wrapped hull -> smallest enclosing rectangle over the hull-edge directions -> 4 corners,
 * plus the hull-vertex -> input-point index map the reference calls points_to_convex_ind.
 *
 * PINNED BY THE REFERENCE'S OWN DEVICE CODE: the reference implementation exists only as CUDA that includes
 * THC/THC.h (not buildable as CUDA against torch 2.11, no CPU twin, no test), but oracle/build_ref.py compiles its
 * __device__ functions as host C++ and tests/golden/device_ops_ref.npz holds their outputs: hull index maps
 * identical, rectangles identical or within 1e-6 (cos, see below) except near-ties of the min-area argmin (same area).
 * Also pinned by properties (cv2.minAreaRect area agreement, containment, analytic cases).
 *
 * Arithmetic follows the reference's mixed precision: fp32 storage, fp64 cross products in
 * the hull (:267-270), fp64 atan2/fmod (:79-83), fp32 for the negative-angle reduction (:85-87)
 * and for the rotation / min / max / area (:113-209), pi = 3.1415926f (:67,346).  Deliberate,
 * documented differences: (1) cos is evaluated in double and rounded to float - the reference
 * calls cos(float), i.e. CUDA cosf (<=1 ulp), which cannot be reproduced on a CPU; (2) no FMA
 * contraction (nvcc's default contraction choices for the reference are unknown); (3) the two
 * gift-wrapping loops are bounded (the reference spins forever on NaN input).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define NPTS 9
#define HP_CAP 12

typedef struct { float x, y; } fpt;

static inline int sgn8(float d) { return (int)(d > 1E-8f) - (int)(d < -1E-8f); }
static inline int near_pt(fpt a, fpt b) { return sgn8(a.x - b.x) == 0 && sgn8(a.y - b.y) == 0; }
static inline float sqdist(fpt a, fpt b)
{
    float dx = a.x - b.x, dy = a.y - b.y;
    float xx = dx * dx, yy = dy * dy;
    return xx + yy;
}
static inline float cosr(float a) { return (float)cos((double)a); }

/* orientation of (cand - base) against (cur - base), in double like :267-270 */
static inline double turn(fpt base, fpt cand, fpt cur)
{
    double l = ((double)cand.x - (double)base.x) * ((double)cur.y - (double)base.y);
    double r = ((double)cur.x - (double)base.x) * ((double)cand.y - (double)base.y);
    return l - r;
}

/* one gift-wrapping chain from slot 0 to slot imax; dir=+1 right chain (:255-281),
 * dir=-1 left chain (:288-312).  Returns the number of pushes (`top`). */
static int wrap_chain(const fpt *P, int n, fpt pmax, int imax, int dir, int *stack)
{
    int top = 0, k = 0;
    stack[0] = 0;
    while (k != imax && top < HP_CAP - 1) {
        fpt pk = pmax;
        k = imax;
        fpt base = P[stack[top]];
        for (int i = 1; i < n; ++i) {
            double s = turn(base, P[i], pk);
            int take = dir > 0 ? (s > 0) : (s < 0);
            if (take || (s == 0 && sqdist(base, P[i]) > sqdist(base, pk))) { pk = P[i]; k = i; }
        }
        stack[++top] = k;
    }
    return top;
}

/* Jarvis_and_index: P is reordered in place to the hull; returns hull size */
static int hull9(fpt *P, int n, int *map)
{
    fpt in0[NPTS];
    memcpy(in0, P, sizeof(fpt) * (size_t)n);
    fpt pmax = P[0];
    int imax = 0;
    for (int i = 0; i < n; ++i) {
        if (P[i].y < P[0].y || (P[i].y == P[0].y && P[i].x < P[0].x)) { fpt t = P[0]; P[0] = P[i]; P[i] = t; }
        if (i == 0) { pmax = P[0]; imax = 0; }
        if (P[i].y > pmax.y || (P[i].y == pmax.y && P[i].x > pmax.x)) { pmax = P[i]; imax = i; }
    }
    if (imax == 0) { imax = 1; pmax = P[1]; }

    int s1[HP_CAP], s2[HP_CAP];
    int top1 = wrap_chain(P, n, pmax, imax, +1, s1);
    int top2 = wrap_chain(P, n, pmax, imax, -1, s2);
    fpt right[HP_CAP], left[HP_CAP];
    for (int i = 0; i <= top1; ++i) right[i] = P[s1[i]];
    for (int i = top2 - 1; i >= 0; --i) left[i] = P[s2[i]];
    int nh = top1 + top2;
    fpt H[2 * HP_CAP];
    for (int i = 0; i < nh; ++i) H[i] = (i <= top1) ? right[i] : left[top2 - (i - top1)];
    for (int i = 0; i < nh; ++i) P[i] = H[i];
    for (int i = 0; i < nh && i < NPTS; ++i)
        for (int j = 0; j < n; ++j)
            if (near_pt(P[i], in0[j])) { map[i] = j; break; }
    return nh;
}

/* minBoundingRect over the CLOSED ring ring[0..m-1] (ring[m-1] == ring[0]) */
static void best_rect(const fpt *ring, int m, float *best /* angle,xmin,ymin,xmax,ymax */)
{
    const float pi_f = 3.1415926f;
    const float hp = pi_f / 2;
    float ang[2 * HP_CAP], uniq[2 * HP_CAP];
    int ne = m - 1, nu = 0;
    for (int i = 0; i < ne; ++i) {
        float ex = ring[i + 1].x - ring[i].x, ey = ring[i + 1].y - ring[i].y;
        float a = (float)atan2((double)ey, (double)ex);
        if (a >= 0) {
            a = (float)fmod((double)a, (double)pi_f / 2);
        } else {
            float q = a / hp;
            float q1 = q - 1;
            int k = (int)q1;
            float t = (float)k * hp;
            a = a - t;
        }
        ang[i] = a;
    }
    uniq[nu++] = ang[0];
    for (int i = 1; i < ne; ++i) {
        int seen = 0;
        for (int j = 0; j < nu; ++j) seen += (ang[i] == uniq[j]);
        if (!seen) uniq[nu++] = ang[i];
    }
    float minarea = 1e12f;
    for (int u = 0; u < nu; ++u) {
        float a = uniq[u];
        float r00 = cosr(a), r01 = cosr(a - hp), r10 = cosr(a + hp), r11 = r00;
        float xmin = 1e12f, ymin = 1e12f, xmax = -1e12f, ymax = -1e12f;
        for (int j = 0; j < m; ++j) {
            float px = r00 * ring[j].x, py = r01 * ring[j].y;
            float rx = (0.0f + px) + py;
            float qx = r10 * ring[j].x, qy = r11 * ring[j].y;
            float ry = (0.0f + qx) + qy;
            if (!(isinf(rx) || isnan(rx))) { if (rx < xmin) xmin = rx; if (rx > xmax) xmax = rx; }
            if (!(isinf(ry) || isnan(ry))) { if (ry < ymin) ymin = ry; if (ry > ymax) ymax = ry; }
        }
        float dx = xmax - xmin, dy = ymax - ymin;
        float area = dx * dy;
        if (area < minarea) {
            minarea = area;
            best[0] = a; best[1] = xmin; best[2] = ymin; best[3] = xmax; best[4] = ymax;
        }
    }
}

/* one point set: in[18] = x0,y0,...,x8,y8 -> out[8], map[9] (-1 padded), returns hull size */
int orc_minarearect_one(const float *in, float *out, int *map)
{
    const float pi_f = 3.1415926f;
    const float hp = pi_f / 2;
    fpt P[2 * HP_CAP];
    for (int i = 0; i < NPTS; ++i) { P[i].x = in[2 * i]; P[i].y = in[2 * i + 1]; map[i] = -1; }
    int nh = hull9(P, NPTS, map);
    fpt ring[2 * HP_CAP + 1];
    for (int i = 0; i < nh; ++i) ring[i] = P[i];
    ring[nh] = P[0];
    float best[5] = {0, 0, 0, 0, 0};
    best_rect(ring, nh + 1, best);
    float a = best[0], xmin = best[1], ymin = best[2], xmax = best[3], ymax = best[4];
    float r00 = cosr(a), r01 = cosr(a - hp), r10 = cosr(a + hp), r11 = r00;
    const float cx[4] = {xmax, xmin, xmin, xmax};
    const float cy[4] = {ymin, ymin, ymax, ymax};
    for (int c = 0; c < 4; ++c) {
        float a0 = cx[c] * r00, b0 = cy[c] * r10;
        out[2 * c] = (0.0f + a0) + b0;
        float a1 = cx[c] * r01, b1 = cy[c] * r11;
        out[2 * c + 1] = (0.0f + a1) + b1;
    }
    return nh;
}

void orc_minarearect(const float *pts18, int n, float *out8, int *hull_map9, int *hull_n)
{
    for (int i = 0; i < n; ++i) {
        int map[NPTS];
        int nh = orc_minarearect_one(pts18 + 18 * (size_t)i, out8 + 8 * (size_t)i, map);
        if (hull_map9) memcpy(hull_map9 + 9 * (size_t)i, map, sizeof(map));
        if (hull_n) hull_n[i] = nh;
    }
}
