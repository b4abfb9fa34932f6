import sys, torch
sys.path.insert(0,'.')
from orientedreppoints_b200.detector import ConvLayer
from orientedreppoints_b200.engine_tc import EngineTC
dev=torch.device('cuda'); e=EngineTC(dev)
cin,cout,k,s,p,h,w,n = [int(v) for v in sys.argv[1:9]]
res = int(sys.argv[9]) if len(sys.argv)>9 else 1
g=torch.Generator().manual_seed(0)
wt=torch.randn(cout,cin,k,k,generator=g)*0.05; b=torch.randn(cout,generator=g)
L=ConvLayer(wt,b,s,p,dev)
x=torch.randn(n,h,w,cin,device=dev).bfloat16()
ho=(h+2*p-k)//s+1; wo=(w+2*p-k)//s+1
r=torch.randn(n,ho,wo,cout,device=dev).bfloat16() if res else None
for _ in range(3): y=e.conv(x,L,relu=True,residual=r)
torch.cuda.synchronize()
e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): y=e.conv(x,L,relu=True,residual=r)
e1.record(); torch.cuda.synchronize()
ms=e0.elapsed_time(e1)/10
fl=2.0*n*ho*wo*cout*cin*k*k
print(f"{ms*1e3:.1f} us  {fl/ms/1e9:.1f} TFLOP/s  bytes/s {(x.numel()+2*n*ho*wo*cout)*2/ms/1e6:.1f} GB/s")
