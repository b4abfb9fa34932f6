import sys, torch
sys.path.insert(0,'.')
from orientedreppoints_b200.weights import random_state_dict
from orientedreppoints_b200.detector import OrientedRepPointsDetector
from orientedreppoints_b200 import _lib
dev=torch.device('cuda'); B=int(sys.argv[1])
det=OrientedRepPointsDetector(random_state_dict(50,0,True),50,dev,'bf16')
img=torch.randn(B,3,1024,1024,device=dev)
for _ in range(3): det.forward_dense(img)
torch.cuda.synchronize(); _lib.set_timing(True); _lib.tc_timing_collect()
det.forward_dense(img); torch.cuda.synchronize()
print(_lib.tc_timing_collect())
