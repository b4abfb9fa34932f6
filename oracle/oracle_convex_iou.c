/*
 * TEST INFRASTRUCTURE ONLY (oracle).  Never imported by the product path.
 *
 * CPU restatement of convex_iou forward:
 *   mmdet/ops/iou/src/convex_iou_kernel.cu:139-266  Jarvis_and_index (gift wrapping of the 9 points, fp64)
 *   mmdet/ops/iou/src/convex_iou_kernel.cu:62-137   area / lineCross / polygon_cut / intersectArea / intersectAreaO
 *   mmdet/ops/iou/src/convex_iou_kernel.cu:268-294  devrIoU: IoU(hull(9 points), quadrilateral), returned as float
 *
 * PINNED: the source is CUDA-only (includes THC/THC.h, not buildable as CUDA against torch 2.11) and has no CPU twin or
 * test, but oracle/build_ref.py compiles its __device__ functions as host C++; this restatement reproduces devrIoU bit for
 * bit (tests/golden/device_ops_ref.npz, tests/test_oracle_golden.py).  The clipping core is the polyiou algorithm whose fp64 instantiation IS pinned
 * bit-for-bit against the compiled DOTA_devkit/polyiou.cpp; the only textual difference (no fabs() on the clipped
 * triangle's area, :124-127) is selected by FAN_SIGNED_AREA.  The whole function is pinned by property: it agrees
 * with cv2.convexHull + cv2.intersectConvexConvex to 1e-6 (tests/test_oracle_golden.py).
 * Every fp operation is separately rounded (-ffp-contract=off); the reference's nvcc build may contract a*b-c*d into
 * FMAs, which is not reproducible on a CPU - the CUDA kernel here uses the same never-contracted sequence.
 */
#include <math.h>
#include <string.h>

#define REAL double
#define FN(n) cx_##n
#define EPSV 1E-8
#define FAN_SIGNED_AREA
#include "polyclip_body.inc"
#undef REAL
#undef FN
#undef EPSV

static inline double cx_dis(cx_pt a, cx_pt b)
{
    double dx = a.x - b.x, dy = a.y - b.y;
    double l = dx * dx, r = dy * dy;
    return l + r;
}

/* convex_iou_kernel.cu:139-266; P: n points in, hull ring out (counter-clockwise from the lowest point); returns the
 * hull size.  Chains are bounded (the reference spins forever on NaN input). */
static int cx_hull(cx_pt *P, int n)
{
    cx_pt pmax = P[0];
    int imax = 0;
    for (int i = 0; i < n; ++i) {
        if (P[i].y < P[0].y || (P[i].y == P[0].y && P[i].x < P[0].x)) { cx_pt t = P[0]; P[0] = P[i]; P[i] = t; }
        if (i == 0) { pmax = P[0]; imax = 0; }
        if (P[i].y > pmax.y || (P[i].y == pmax.y && P[i].x > pmax.x)) { pmax = P[i]; imax = i; }
    }
    if (imax == 0) { imax = 1; pmax = P[1]; }
    int st1[24], st2[24], top1 = 0, top2 = 0;
    for (int dir = 0; dir < 2; ++dir) {
        int *st = dir ? st2 : st1, top = 0, k = 0;
        st[0] = 0;
        while (k != imax && top < 20) {
            cx_pt pk = pmax;
            k = imax;
            for (int i = 1; i < n; ++i) {
                double s = cx_cross3(P[st[top]], P[i], pk);
                int take = dir ? (s < 0) : (s > 0);
                if (take || (s == 0 && cx_dis(P[st[top]], P[i]) > cx_dis(P[st[top]], pk))) { pk = P[i]; k = i; }
            }
            st[++top] = k;
        }
        if (dir) top2 = top; else top1 = top;
    }
    cx_pt out[48];
    int nh = top1 + top2;
    for (int i = 0; i < nh; ++i) out[i] = (i <= top1) ? P[st1[i]] : P[st2[top2 - (i - top1)]];
    memcpy(P, out, sizeof(cx_pt) * (size_t)nh);
    return nh;
}

static void cx_reverse(cx_pt *v, int n)
{
    for (int i = 0, j = n - 1; i < j; ++i, --j) { cx_pt t = v[i]; v[i] = v[j]; v[j] = t; }
}

/* devrIoU, convex_iou_kernel.cu:268-294 */
float orc_convex_iou_one(const float *pts18, const float *quad8)
{
    cx_pt A[50], B[8];
    for (int i = 0; i < 9; ++i) { A[i].x = (double)pts18[2 * i]; A[i].y = (double)pts18[2 * i + 1]; }
    int n1 = cx_hull(A, 9);
    for (int i = 0; i < 4; ++i) { B[i].x = (double)quad8[2 * i]; B[i].y = (double)quad8[2 * i + 1]; }
    int n2 = 4;
    if (cx_ring_area(A, n1) < 0) cx_reverse(A, n1);
    if (cx_ring_area(B, n2) < 0) cx_reverse(B, n2);
    A[n1] = A[0]; B[n2] = B[0];
    double inter = 0;
    for (int i = 0; i < n1; ++i)
        for (int j = 0; j < n2; ++j) {
            double t = cx_fan_pair(A[i], A[i + 1], B[j], B[j + 1]);
            inter = inter + t;
        }
    double sp = cx_ring_area(A, n1), sq = cx_ring_area(B, n2);
    double uni = fabs(sp) + fabs(sq);
    uni = uni - inter;
    return (float)(inter / uni);
}

/* convex_iou_kernel.cu:297-312: out[i * K + j] */
void orc_convex_iou(const float *pts18, int n, const float *quads8, int k, float *out)
{
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < k; ++j) out[(size_t)i * k + j] = orc_convex_iou_one(pts18 + 18 * (size_t)i, quads8 + 8 * (size_t)j);
}

int orc_convex_hull9(const float *pts18, double *ring /* >= 2*9 doubles */)
{
    cx_pt A[50];
    for (int i = 0; i < 9; ++i) { A[i].x = (double)pts18[2 * i]; A[i].y = (double)pts18[2 * i + 1]; }
    int n = cx_hull(A, 9);
    for (int i = 0; i < n && i < 9; ++i) { ring[2 * i] = A[i].x; ring[2 * i + 1] = A[i].y; }
    return n;
}
