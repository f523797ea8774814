"""Scene sharding across ranks for the two hot paths.

Scenes are independent units (noise -> DDIM -> occupancy grid -> render): rank r takes scenes r, r+W, r+2W, ... exactly like the
reference's `DistributedSampler(shuffle=False)` (lib/datasets/samplers/distributed_sampler.py:53-59, `indices[rank::world]`).
There is no collective on the data path; `gather_scene_outputs` is the eval-side exchange (the reference all-gathers Inception
features inside mmgen's FID.feed; we gather whatever per-scene tensor the caller produced) and works with NCCL or gloo."""
import torch
import torch.distributed as dist


def scene_indices(num_scenes, rank=None, world_size=None):
    """indices of the scenes this rank processes (strided; ragged tails allowed)"""
    if rank is None:
        rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    return list(range(rank, num_scenes, world_size))


def gather_scene_outputs(local, num_scenes, group=None):
    """local: tensor [n_local, ...] for `scene_indices(num_scenes)` in order -> [num_scenes, ...] on every rank (scene order)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n_max = (num_scenes + world - 1) // world
    pad = local.new_zeros((n_max,) + tuple(local.shape[1:]))
    pad[:local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad.contiguous(), group=group)
    out = local.new_zeros((num_scenes,) + tuple(local.shape[1:]))
    for r in range(world):
        idx = scene_indices(num_scenes, r, world)
        out[idx] = bufs[r][:len(idx)]
    return out
