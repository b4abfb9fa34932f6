python - <<'PY'
import torch
d=torch.device('cuda',0)
x=torch.empty(1<<30,dtype=torch.uint8,device=d); y=torch.empty(1<<30,dtype=torch.uint8,device=d)
def t(f,n=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n
ms=t(lambda: x.fill_(1)); print("fill 1GiB  %.1f us  %.2f TB/s write" % (ms*1e3, (1<<30)/ms/1e9))
ms=t(lambda: y.copy_(x)); print("copy 1GiB  %.1f us  %.2f TB/s r+w" % (ms*1e3, 2*(1<<30)/ms/1e9))
ms=t(lambda: x.sum()); print("sum 1GiB (uint8)  %.1f us  %.2f TB/s read" % (ms*1e3, (1<<30)/ms/1e9))
PY
python tools/one_conv.py 64 256 1 1 0 256 256 16 0
python tools/one_conv.py 64 256 1 1 0 256 256 16 1
python tools/one_conv.py 128 512 1 1 0 128 128 16 1
python tools/one_conv.py 256 1024 1 1 0 64 64 16 1
python tools/one_conv.py 256 1024 1 1 0 64 64 16 0
ncu --set full --clock-control none -k regex:conv_tc -s 5 -c 1 -o gpurun_out/r2_exp_1x1_nores python tools/one_conv.py 64 256 1 1 0 256 256 16 0 > /dev/null 2>&1
ncu --set full --clock-control none -k regex:conv_tc -s 5 -c 1 -o gpurun_out/r2_exp_1x1_l3 python tools/one_conv.py 256 1024 1 1 0 64 64 16 1 > /dev/null 2>&1
