"""smoke() leg for the dense path: one small tile through the tensor-core detector on cuda:0, dense
outputs checked against the PyTorch fp32 re-declaration of the reference graph (oracle/torch_reference.py,
test infrastructure - imported here only because __graft_entry__.smoke() is allowed to use the checker)."""
import torch


def run(dev):
    from oracle import torch_reference as tr
    from .detector import OrientedRepPointsDetector
    from .weights import random_state_dict
    sd = random_state_dict(50, seed=0, reference_init=False)
    det = OrientedRepPointsDetector(sd, 50, dev, "bf16", test_cfg=dict(score_thr=0.02))
    img = torch.randn(1, 3, 256, 256, generator=torch.Generator().manual_seed(7)).to(dev)
    outs, feats = det.forward_dense(img)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    with torch.no_grad():
        ref_outs, ref_feats = tr.forward_dense({k: v.to(dev) for k, v in sd.items()}, img)
    for lvl in range(5):
        a = feats[lvl].float().permute(0, 3, 1, 2)
        assert float((a - ref_feats[lvl]).abs().max()) < 0.08 * float(ref_feats[lvl].abs().max()), lvl
        for k in range(3):
            a, b = outs[lvl][k].permute(0, 3, 1, 2), ref_outs[lvl][k]
            assert float((a - b).abs().max()) < 0.1 * max(1.0, float(b.abs().max())), (lvl, k)
    res = det.simple_test(img)
    assert len(res) == 1 and len(res[0]) == 15
    print("smoke_model ok: bf16 tensor-core graph within tolerance of the fp32 torch graph; %d detections"
          % sum(len(a) for a in res[0]))
