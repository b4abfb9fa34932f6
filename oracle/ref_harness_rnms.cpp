// TEST INFRASTRUCTURE ONLY.  Second translation unit linked next to the reference's
// UNMODIFIED mmdet/ops/nms/src/rnms_cpu.cpp (compiled from /root/reference by
// oracle/build_ref.py).  rotate_iou() there has external linkage (rnms_cpu.cpp:121);
// this file only re-exports it with a C ABI so the fp32 oracle can be pinned bit-for-bit.
float rotate_iou(float const x11, float const y11, float const x12, float const y12,
                 float const x13, float const y13, float const x14, float const y14,
                 float const x21, float const y21, float const x22, float const y22,
                 float const x23, float const y23, float const x24, float const y24);

extern "C" float ref_rotate_iou(const float *p, const float *q)
{
    return rotate_iou(p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7],
                      q[0], q[1], q[2], q[3], q[4], q[5], q[6], q[7]);
}

extern "C" void ref_rotate_iou_pairs(const float *p, const float *q, int n, float *out)
{
    for (int i = 0; i < n; ++i) out[i] = ref_rotate_iou(p + 8 * i, q + 8 * i);
}
