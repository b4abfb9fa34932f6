/*
 * TEST INFRASTRUCTURE ONLY (oracle).  Never imported by the product path; only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.
 *
 * CPU restatement of the rotated-IoU / rotated-NMS family of the reference:
 *   - fp64 quad IoU            DOTA_devkit/polyiou.cpp:108-128
 *   - fp32 quad IoU (rnms)     mmdet/ops/nms/src/rnms_cpu.cpp:121-163 == rnms_kernel.cu:131-147
 *   - fp32 quad IoU (poly_nms) DOTA_devkit/poly_nms_gpu/poly_nms_kernel.cu:192-212 (zero-union guard)
 *   - (cx,cy,w,h,theta)->quad  DOTA_devkit/poly_nms_gpu/poly_overlaps_kernel.cu:280-297
 *   - greedy NMS drivers       DOTA_devkit/ResultMerge.py:18-41, ResultMerge_multi_process.py:60-121,
 *                              rnms_kernel.cu:204-265, poly_nms.pyx:9-24
 *
 * Pinned (tests/test_oracle_vs_reference.py) bit-for-bit against the reference's own
 * sources compiled in the authoring container (oracle/build_ref.py -> oracle/_ref/).
 * Build: see oracle/build_oracle.py  (gcc -O2 -ffp-contract=off -fno-fast-math).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---------------- fp64 instantiation (polyiou.cpp) ---------------- */
#define REAL double
#define FN(n) d_##n
#define EPSV 1E-8
#include "polyclip_body.inc"
#undef REAL
#undef FN
#undef EPSV

/* ---------------- fp32 instantiation (rnms / poly_nms) ---------------- */
#define REAL float
#define FN(n) f_##n
#define EPSV 1E-8f
#include "polyclip_body.inc"
#undef REAL
#undef FN
#undef EPSV

/* polyiou.cpp:108-128 */
double orc_iou_poly_f64(const double *p, const double *q)
{
    d_pair_res r = d_quad_pair(p, q);
    double uni = r.area_p + r.area_q - r.inter;
    return r.inter / uni;
}

/* rnms_cpu.cpp:121-163: no guard, 0/0 -> NaN */
float orc_iou_rnms_f32(const float *p, const float *q)
{
    f_pair_res r = f_quad_pair(p, q);
    float uni = r.area_p + r.area_q - r.inter;
    return r.inter / uni;
}

/* poly_nms_kernel.cu:192-212: union==0 -> (inter+1)/(union+1) */
float orc_iou_polynms_f32(const float *p, const float *q)
{
    f_pair_res r = f_quad_pair(p, q);
    float uni = r.area_p + r.area_q - r.inter;
    if (uni == 0) return (r.inter + 1) / (uni + 1);
    return r.inter / uni;
}

void orc_iou_poly_f64_pairs(const double *p, const double *q, int n, double *out)
{
    for (int i = 0; i < n; ++i) out[i] = orc_iou_poly_f64(p + 8 * i, q + 8 * i);
}

void orc_iou_rnms_f32_pairs(const float *p, const float *q, int n, float *out)
{
    for (int i = 0; i < n; ++i) out[i] = orc_iou_rnms_f32(p + 8 * i, q + 8 * i);
}

/* N x K matrix, fp64 */
void orc_iou_poly_f64_matrix(const double *p, int n, const double *q, int k, double *out)
{
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < k; ++j) out[(size_t)i * k + j] = orc_iou_poly_f64(p + 8 * i, q + 8 * j);
}

/* poly_overlaps_kernel.cu:280-297.  The reference mixes float and double:
 * `w / 2.0` is double, `cs * (w / 2.0)` is double, the sum is rounded to float on
 * store.  cs/ss are float cos/sin of the float angle. */
static void rotbox_to_quad_f32(const float *b, float *quad)
{
    float cs = cosf(b[4]);
    float ss = sinf(b[4]);
    float w = b[2], h = b[3];
    float xc = b[0], yc = b[1];
    double hw = w / 2.0, hh = h / 2.0, nhw = -w / 2.0, nhh = -h / 2.0;
    quad[0] = (float)(xc + cs * hw - ss * nhh);
    quad[2] = (float)(xc + cs * hw - ss * hh);
    quad[4] = (float)(xc + cs * nhw - ss * hh);
    quad[6] = (float)(xc + cs * nhw - ss * nhh);
    quad[1] = (float)(yc + ss * hw + cs * nhh);
    quad[3] = (float)(yc + ss * hw + cs * hh);
    quad[5] = (float)(yc + ss * nhw + cs * hh);
    quad[7] = (float)(yc + ss * nhw + cs * nhh);
}

void orc_rotbox_to_quad_f32(const float *boxes5, int n, float *quads8)
{
    for (int i = 0; i < n; ++i) rotbox_to_quad_f32(boxes5 + 5 * i, quads8 + 8 * i);
}

/* poly_overlaps_kernel.cu:330-352, N x K fp32 with the zero-union guard */
void orc_poly_overlaps_f32(const float *boxes5, int n, const float *query5, int k, float *out)
{
    float *qa = (float *)malloc(sizeof(float) * 8 * (size_t)(n > 0 ? n : 1));
    float *qb = (float *)malloc(sizeof(float) * 8 * (size_t)(k > 0 ? k : 1));
    orc_rotbox_to_quad_f32(boxes5, n, qa);
    orc_rotbox_to_quad_f32(query5, k, qb);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < k; ++j)
            out[(size_t)i * k + j] = orc_iou_polynms_f32(qa + 8 * i, qb + 8 * j);
    free(qa);
    free(qb);
}

/* ---------------------------------------------------------------------------
 * score ordering.  The reference orders with scores.argsort()[::-1]
 * (ResultMerge.py:28, poly_nms.pyx:19) or a descending torch sort
 * (rnms_kernel.cu:208); both leave ties implementation-defined.  The oracle (and
 * the CUDA path) define ties as: higher score first, lower original index first.
 * ------------------------------------------------------------------------- */
typedef struct { double s; int i; } sidx;
static int cmp_sidx(const void *a, const void *b)
{
    const sidx *x = (const sidx *)a, *y = (const sidx *)b;
    if (x->s > y->s) return -1;
    if (x->s < y->s) return 1;
    return (x->i > y->i) - (x->i < y->i);
}
static int *order_desc(const void *dets, int n, int is_f64)
{
    sidx *t = (sidx *)malloc(sizeof(sidx) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; ++i) {
        t[i].s = is_f64 ? ((const double *)dets)[9 * (size_t)i + 8] : (double)((const float *)dets)[9 * (size_t)i + 8];
        t[i].i = i;
    }
    qsort(t, (size_t)n, sizeof(sidx), cmp_sidx);
    int *o = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; ++i) o[i] = t[i].i;
    free(t);
    return o;
}

/* ResultMerge.py:18-41 (py_cpu_nms_poly): fp64, survivors of a kept box are those
 * with `iou <= thresh` (a NaN IoU therefore suppresses), keep list in score order.
 * dets: [n,9] double (x1..y4, score).  Returns number kept. */
int orc_nms_poly_f64(const double *dets, int n, double thresh, int *keep)
{
    int *ord = order_desc(dets, n, 1);
    unsigned char *dead = (unsigned char *)calloc((size_t)(n > 0 ? n : 1), 1);
    int nk = 0;
    for (int a = 0; a < n; ++a) {
        if (dead[a]) continue;
        int i = ord[a];
        keep[nk++] = i;
        for (int b = a + 1; b < n; ++b) {
            if (dead[b]) continue;
            double v = orc_iou_poly_f64(dets + 9 * (size_t)i, dets + 9 * (size_t)ord[b]);
            if (!(v <= thresh)) dead[b] = 1;
        }
    }
    free(ord);
    free(dead);
    return nk;
}

/* ResultMerge_multi_process.py:60-121 (py_cpu_nms_poly_fast): same, but a pair is
 * only clipped when the axis-aligned hulls overlap with positive area
 * (`hbb_ovr > 0`, :96); otherwise the recorded overlap is hbb_ovr (<= 0 or NaN). */
int orc_nms_poly_fast_f64(const double *dets, int n, double thresh, int *keep)
{
    int *ord = order_desc(dets, n, 1);
    unsigned char *dead = (unsigned char *)calloc((size_t)(n > 0 ? n : 1), 1);
    double *bb = (double *)malloc(sizeof(double) * 5 * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; ++i) {
        const double *d = dets + 9 * (size_t)i;
        double x1 = d[0], x2 = d[0], y1 = d[1], y2 = d[1];
        for (int k = 1; k < 4; ++k) {
            if (d[2 * k] < x1) x1 = d[2 * k];
            if (d[2 * k] > x2) x2 = d[2 * k];
            if (d[2 * k + 1] < y1) y1 = d[2 * k + 1];
            if (d[2 * k + 1] > y2) y2 = d[2 * k + 1];
        }
        bb[5 * i] = x1; bb[5 * i + 1] = y1; bb[5 * i + 2] = x2; bb[5 * i + 3] = y2;
        bb[5 * i + 4] = (x2 - x1 + 1) * (y2 - y1 + 1);            /* :68 */
    }
    int nk = 0;
    for (int a = 0; a < n; ++a) {
        if (dead[a]) continue;
        int i = ord[a];
        keep[nk++] = i;
        for (int b = a + 1; b < n; ++b) {
            if (dead[b]) continue;
            int j = ord[b];
            double xx1 = fmax(bb[5 * i], bb[5 * j]), yy1 = fmax(bb[5 * i + 1], bb[5 * j + 1]);
            double xx2 = fmin(bb[5 * i + 2], bb[5 * j + 2]), yy2 = fmin(bb[5 * i + 3], bb[5 * j + 3]);
            double w = fmax(0.0, xx2 - xx1), h = fmax(0.0, yy2 - yy1);
            double hin = w * h;
            double v = hin / (bb[5 * i + 4] + bb[5 * j + 4] - hin);
            if (v > 0) v = orc_iou_poly_f64(dets + 9 * (size_t)i, dets + 9 * (size_t)j);
            if (!(v <= thresh)) dead[b] = 1;
        }
    }
    free(ord);
    free(dead);
    free(bb);
    return nk;
}

/* rnms_kernel.cu:149-265 / poly_nms_kernel.cu:214-329: fp32 greedy NMS with the
 * `iou > thr` predicate (NaN keeps).  guard=0 -> rnms IoU, guard=1 -> poly_nms IoU.
 * keep_sel receives kept original indices in selection (score) order - what
 * poly_gpu_nms returns; sort ascending for rnms (rnms_kernel.cu:261-264). */
int orc_nms_f32(const float *dets, int n, float thresh, int guard, int *keep_sel)
{
    int *ord = order_desc(dets, n, 0);
    unsigned char *dead = (unsigned char *)calloc((size_t)(n > 0 ? n : 1), 1);
    int nk = 0;
    for (int a = 0; a < n; ++a) {
        if (dead[a]) continue;
        int i = ord[a];
        keep_sel[nk++] = i;
        for (int b = a + 1; b < n; ++b) {
            if (dead[b]) continue;
            const float *p = dets + 9 * (size_t)i, *q = dets + 9 * (size_t)ord[b];
            float v = guard ? orc_iou_polynms_f32(p, q) : orc_iou_rnms_f32(p, q);
            if (v > thresh) dead[b] = 1;
        }
    }
    free(ord);
    free(dead);
    return nk;
}
