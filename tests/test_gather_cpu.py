"""CPU, world_size 2, gloo: the detection all-gather and the dataset-order interleave
(replacement of mmdet/apis/test.py:117-147 collect_results_gpu)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from orientedreppoints_b200 import gather
    cap, tiles = 16, 3
    g = torch.Generator().manual_seed(rank)
    counts = torch.tensor([(5 * rank + 3 * t) % cap for t in range(tiles)], dtype=torch.int32)
    dets = torch.zeros(tiles, cap, 27)
    labels = torch.full((tiles, cap), -1, dtype=torch.int64)
    for t in range(tiles):
        k = int(counts[t])
        dets[t, :k] = torch.randn(k, 27, generator=g)
        labels[t, :k] = torch.randint(0, 15, (k,), generator=g)
    buf, cnt = gather.pack(dets, labels, counts)
    ab, ac = gather.all_gather_detections(buf, cnt)
    out = gather.interleave(ab, ac, dataset_len=5)       # 6 slots, dataset of 5: last one is sampler padding
    # numpy payloads are pickled by value (torch tensors travel as shared-memory handles that die with the worker)
    q.put((rank, [(d.numpy().copy(), l.numpy().copy()) for d, l in out], dets.numpy(), labels.numpy(), counts.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_all_gather_interleave_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda x: x[0])
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    per_rank = {r: (d, l, c) for r, _, d, l, c in res}
    for r, out, _, _, _ in res:
        assert len(out) == 5
        for i, (d, l) in enumerate(out):
            src_rank, slot = i % world, i // world
            sd, sl, sc = per_rank[src_rank]
            k = int(sc[slot])
            assert d.shape == (k, 27) and (d == sd[slot, :k]).all() and (l == sl[slot, :k]).all()
