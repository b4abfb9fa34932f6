"""Gathering per-tile detections across ranks: the replacement of collect_results_gpu
(mmdet/apis/test.py:117-147: pickle -> uint8 tensor -> all_gather(shape) -> all_gather(padded bytes) ->
unpickle) by ONE all_gather_into_tensor of a fixed-layout buffer.

Layout per tile: float32 [max_per_img, 28] = reppoints(18) | box(8) | score | label, zero padded, plus an
int32 count.  Rank r holds tiles r, r+world, r+2*world, ... (DistributedSampler(shuffle=False) takes
indices[rank::world], mmdet/datasets/loader/sampler.py:26-32); `interleave` restores dataset order exactly as
`zip(*part_list)` + truncation does at apis/test.py:143-147.
"""
import torch
import torch.distributed as dist


def pack(dets, labels, counts):
    """padded (dets [T,cap,27], labels [T,cap], counts [T]) -> (buf [T,cap,28], counts)"""
    buf = torch.cat([dets, labels.to(dets.dtype).unsqueeze(-1)], dim=2)
    return buf.contiguous(), counts.to(torch.int32).contiguous()


def all_gather_detections(buf, counts, group=None):
    """-> (all_buf [world,T,cap,28], all_counts [world,T]); a single collective for the payload."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return buf.unsqueeze(0), counts.unsqueeze(0)
    all_buf = torch.empty((world,) + tuple(buf.shape), dtype=buf.dtype, device=buf.device)
    all_cnt = torch.empty((world,) + tuple(counts.shape), dtype=counts.dtype, device=counts.device)
    if buf.is_cuda:
        dist.all_gather_into_tensor(all_buf, buf, group=group)
        dist.all_gather_into_tensor(all_cnt, counts, group=group)
    else:   # gloo (CPU tests)
        lb = [torch.empty_like(buf) for _ in range(world)]
        lc = [torch.empty_like(counts) for _ in range(world)]
        dist.all_gather(lb, buf, group=group)
        dist.all_gather(lc, counts, group=group)
        all_buf, all_cnt = torch.stack(lb), torch.stack(lc)
    return all_buf, all_cnt


def interleave(all_buf, all_counts, dataset_len):
    """[world,T,...] -> per-tile list in dataset order (tile i lives on rank i % world, slot i // world)"""
    world, t = all_counts.shape
    out = []
    for i in range(min(dataset_len, world * t)):
        r, s = i % world, i // world
        k = int(all_counts[r, s])
        rows = all_buf[r, s, :k]
        out.append((rows[:, :27], rows[:, 27].long()))
    return out
