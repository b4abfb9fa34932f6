import numpy as np, torch, sys, time
sys.path.insert(0,'.')
from oracle import pyoracle as po
from orientedreppoints_b200 import _lib
from orientedreppoints_b200.ops import rnms_indices
dev=torch.device('cuda')
def run(d, thr=0.1, reps=5):
    dt=torch.from_numpy(d).to(dev)
    keep,cnt=rnms_indices(dt,thr,return_count_tensor=True); torch.cuda.synchronize()
    ts=[]
    for _ in range(reps):
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record(); keep,cnt=rnms_indices(dt,thr,return_count_tensor=True); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    st=_lib.last_nms_stats()
    return min(ts), int(cnt.item()), st
for n in (1000,10000,20000,50000,100000,200000):
    for variant in ('dense','const'):
        ext = 1024.0 if variant=='dense' else 1024.0*np.sqrt(n/1000.0)
        d=po.gen_rotated_boxes(n,seed=1,extent=ext)
        ms,k,st=run(d)
        pairs=n*(n-1)/2
        print(f"N={n:7d} {variant:5s} ms={ms:9.3f} kept={k:6d} Mpairs/s={pairs/ms/1e3:12.1f} swept={st['pairs_total']:.3e} aabb={st['pairs_aabb']:.3e} clip={st['pairs_clipped']:.3e} fp64={st['pairs_fp64']} edges={st['edges']:.3e} rounds={st['rounds']}", flush=True)
d=po.gen_clustered_boxes(2000,50,seed=2)
ms,k,st=run(d,0.1); n=len(d)
print(f"clustered N={n} ms={ms:.3f} kept={k} swept={st['pairs_total']:.3e} aabb={st['pairs_aabb']:.3e} clip={st['pairs_clipped']:.3e} fp64={st['pairs_fp64']} edges={st['edges']:.3e} rounds={st['rounds']}")
