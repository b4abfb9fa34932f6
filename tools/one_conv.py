"""one convolution layer on the tensor-core engine, timed with CUDA events (and the target of single-kernel ncu captures)
    python tools/one_conv.py cin cout k s p h w n [res=1] [prec=f16x3|bf16] [deform=0]"""
import sys, torch
sys.path.insert(0, '.')
from orientedreppoints_b200.detector import ConvLayer
from orientedreppoints_b200.engine_tc import EngineTC, EngineTCSplit
dev = torch.device('cuda', 0)
cin, cout, k, s, p, h, w, n = [int(v) for v in sys.argv[1:9]]
res = int(sys.argv[9]) if len(sys.argv) > 9 else 1
prec = sys.argv[10] if len(sys.argv) > 10 else 'f16x3'
deform = int(sys.argv[11]) if len(sys.argv) > 11 else 0
e = EngineTCSplit(dev) if prec == 'f16x3' else EngineTC(dev)
g = torch.Generator().manual_seed(0)
wt = torch.randn(cout, cin, k, k, generator=g) * 0.05; b = torch.randn(cout, generator=g)
L = ConvLayer(wt, None if deform else b, s, p, dev)
x = e.from_float(torch.randn(n, h, w, cin, device=dev))
ho = (h + 2 * p - k) // s + 1; wo = (w + 2 * p - k) // s + 1
r = e.from_float(torch.randn(n, ho, wo, cout, device=dev)) if res else None
off = (torch.randn(n, ho, wo, 2 * k * k, device=dev) * (float(sys.argv[12]) if len(sys.argv) > 12 else 1.0)).contiguous() if deform else None
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
def run():
    if deform:
        return e.deform_conv_multi([x], [off], L, relu=True)[0]
    return e.conv(x, L, relu=True, residual=r)
for _ in range(3): y = run()
torch.cuda.synchronize()
ts = []
for i in range(10):
    flush.fill_(i)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); y = run(); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ms = sorted(ts)[len(ts) // 2]
fl = 2.0 * n * ho * wo * cout * cin * k * k
by = (x.numel() + y.numel() + (r.numel() if res else 0)) * 2
print(f"{prec} deform={deform} {ms*1e3:.1f} us  {fl/ms/1e9:.1f} TFLOP/s algorithmic  {by/ms/1e6:.1f} GB/s (in+out+res bytes, L2 flushed)")
