// TEST INFRASTRUCTURE ONLY.  Compiles the reference's own DOTA_devkit/polyiou.cpp (textually
// included from where it lies under /root/reference - never copied into this repo) and
// exposes its iou_poly through a C ABI so the oracle can be pinned against it with ctypes.
// Built by oracle/build_ref.py into oracle/_ref/ (git-ignored).
#ifndef REF_POLYIOU_CPP
#error "pass -DREF_POLYIOU_CPP=\"<path to reference DOTA_devkit/polyiou.cpp>\""
#endif
#include REF_POLYIOU_CPP

extern "C" double ref_iou_poly(const double *p, const double *q)
{
    std::vector<double> P(p, p + 8), Q(q, q + 8);
    return iou_poly(P, Q);
}

extern "C" void ref_iou_poly_pairs(const double *p, const double *q, int n, double *out)
{
    for (int i = 0; i < n; ++i) out[i] = ref_iou_poly(p + 8 * i, q + 8 * i);
}
