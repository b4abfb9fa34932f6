// TEST INFRASTRUCTURE ONLY.  Linked next to the reference's UNMODIFIED
// mmdet/ops/box_iou_rotated/src/box_iou_rotated_cpu.cpp (compiled from /root/reference by oracle/build_ref.py):
// re-exports box_iou_rotated_cpu (box_iou_rotated.h:7-9) with a C ABI so golden vectors can be minted through ctypes.
#include <torch/extension.h>

at::Tensor box_iou_rotated_cpu(const at::Tensor& boxes1, const at::Tensor& boxes2);

extern "C" void ref_box_iou_rotated(const float *b1, int n, const float *b2, int m, float *out)
{
    auto t1 = torch::from_blob(const_cast<float *>(b1), {n, 5}, torch::kFloat32).clone();
    auto t2 = torch::from_blob(const_cast<float *>(b2), {m, 5}, torch::kFloat32).clone();
    auto r = box_iou_rotated_cpu(t1, t2).contiguous();
    memcpy(out, r.data_ptr<float>(), sizeof(float) * (size_t)n * m);
}
