import sys, time, torch
sys.path.insert(0,'.')
from orientedreppoints_b200.weights import random_state_dict
from orientedreppoints_b200.detector import OrientedRepPointsDetector, STRIDES
from orientedreppoints_b200.core.get_bboxes import get_bboxes
dev=torch.device('cuda')
sd=random_state_dict(50,0,True)
det=OrientedRepPointsDetector(sd,50,dev,'bf16',test_cfg=dict(score_thr=0.0))
def timeit(fn, reps=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    t0=time.perf_counter(); e0.record()
    for _ in range(reps): fn()
    e1.record(); t1=time.perf_counter(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/reps, (t1-t0)*1e3/reps
for B in (1,4,8):
    img=torch.randn(B,3,1024,1024,device=dev)
    te=timeit(lambda: det.forward_dense(img))
    det.capture(img.shape)
    tg=timeit(lambda: det.forward_dense_graph(img))
    outs,_=det.forward_dense_graph(img)
    metas=[dict(scale_factor=1.0)]*B
    tp=timeit(lambda: get_bboxes([o[0] for o in outs],[o[2] for o in outs],STRIDES,metas,det.test_cfg,True), reps=5)
    print(f"B={B}: eager dense gpu {te[0]:.3f} ms (cpu issue {te[1]:.3f}), graph dense {tg[0]:.3f} ms (cpu {tg[1]:.3f}), post {tp[0]:.3f} (cpu {tp[1]:.3f}); mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB", flush=True)
    det._g_shape=None; del det._graph; torch.cuda.empty_cache()
