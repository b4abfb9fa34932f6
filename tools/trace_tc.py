"""per-launch trace of the tensor-core convolutions of one dense step (ORP_TC_TRACE=1 prints every launch)
    ORP_TC_TRACE=1 python tools/trace_tc.py <tiles> [f16x3|bf16] [depth]"""
import sys, time, torch
sys.path.insert(0, '.')
from orientedreppoints_b200.weights import random_state_dict
from orientedreppoints_b200.detector import OrientedRepPointsDetector
from orientedreppoints_b200 import _lib
dev = torch.device('cuda'); B = int(sys.argv[1])
prec = sys.argv[2] if len(sys.argv) > 2 else 'f16x3'
depth = (sys.argv[3] if len(sys.argv) > 3 else '50')
if depth == 'swin_tiny':
    from orientedreppoints_b200.swin import random_swin_state_dict
    sd = random_swin_state_dict(0)
else:
    depth = int(depth)
    sd = random_state_dict(depth, 0, True)
det = OrientedRepPointsDetector(sd, depth, dev, prec, test_cfg=dict(score_thr=0.0))
img = torch.randint(0, 256, (B, 1024, 1024, 3), dtype=torch.uint8, device=dev)
for _ in range(3): det.forward_dense(img)
torch.cuda.synchronize(); _lib.set_timing(True); _lib.tc_timing_collect()
det.forward_dense(img); torch.cuda.synchronize()
ms, n, fl = _lib.tc_timing_collect()
_lib.set_timing(False)
print("%s %s %d tiles: %d tc launches %.3f ms, %.1f TFLOP/s algorithmic" % (prec, depth, B, n, ms, fl / ms / 1e9))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
det.capture(img.shape, img.dtype)
for _ in range(3): det.simple_test(img, return_tensors="padded")
torch.cuda.synchronize(); e0.record()
for _ in range(10): det.simple_test(img, return_tensors="padded")
e1.record(); torch.cuda.synchronize()
t = e0.elapsed_time(e1) / 10
print("%s whole step (graph + post): %.3f ms -> %.1f tiles/s" % (prec, t, B / t * 1e3))
if hasattr(det.eng, "overflow_count"): print("overflow count", det.eng.overflow_count())
