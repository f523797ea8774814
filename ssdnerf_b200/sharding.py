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


# ---------------------------------------------------------------------------------------------------------------------
# view sharding (secondary mode): few scenes, many views -- broadcast the scene, shard the views
# ---------------------------------------------------------------------------------------------------------------------
def view_range(num_views, rank=None, world_size=None):
    """contiguous [lo, hi) range of views rendered by this rank (sizes differ by at most one)"""
    if rank is None:
        rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    return (num_views * rank) // world_size, (num_views * (rank + 1)) // world_size


def max_views_per_rank(num_views, world_size):
    return max(view_range(num_views, r, world_size)[1] - view_range(num_views, r, world_size)[0] for r in range(world_size))


def broadcast_scene(code, density_bitfield, src=0, group=None):
    """rank `src` owns the scene(s): code [B,3,C,H,W] fp32 (1.2 - 6.3 MB per scene) and the occupancy bitfield (32 KB) go to every rank"""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(code, src, group=group)
        dist.broadcast(density_bitfield, src, group=group)
    return code, density_bitfield


def gather_views(local_images, num_views, group=None, out=None, padded=None):
    """local_images [v_local, ...] (this rank's `view_range`) -> [num_views, ...] on every rank, view order preserved.
    `padded` / `out`: optional preallocated [v_max, ...] / [world * v_max, ...] buffers (NCCL all_gather_into_tensor needs equal sizes)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local_images
    world = dist.get_world_size(group)
    v_max = max_views_per_rank(num_views, world)
    if padded is None:
        padded = local_images.new_zeros((v_max,) + tuple(local_images.shape[1:]))
    padded[:local_images.shape[0]].copy_(local_images)
    if out is None:
        out = local_images.new_empty((world * v_max,) + tuple(local_images.shape[1:]))
    dist.all_gather_into_tensor(out, padded, group=group)
    if num_views == world * v_max:
        return out
    return torch.cat([out[r * v_max: r * v_max + (view_range(num_views, r, world)[1] - view_range(num_views, r, world)[0])] for r in range(world)])
