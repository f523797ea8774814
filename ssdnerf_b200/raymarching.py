"""Host-side mirror of the reference's ``lib.ops.raymarching`` package, on top of the C ABI.

Same callables, argument order and return values as lib/ops/raymarching/raymarching.py (cited per
function) so the reference's own driver code (VolumeRenderer.forward) runs unchanged on top of it.
Differences: kernels run on the caller's current stream; inputs must already be CUDA tensors (the
reference silently `.cuda()`s them); errors raise instead of being undefined behaviour.
"""
from itertools import groupby

import torch
from torch.autograd import Function

from . import _lib as N


def _f32(t):
    return t.contiguous().float()


def _call(name, *args):
    N.check(getattr(N.lib(), name)(*args))


def near_far_from_aabb(rays_o, rays_d, aabb, min_near=0.2):
    """raymarching.py:20-55. rays_o/d [N,3] (any leading shape), aabb [6] -> nears, fars [N]."""
    N.require_cuda(rays_o, rays_d, aabb)
    rays_o = _f32(rays_o).view(-1, 3)
    rays_d = _f32(rays_d).view(-1, 3)
    aabb = _f32(aabb)
    n = rays_o.shape[0]
    nears = torch.empty(n, dtype=torch.float32, device=rays_o.device)
    fars = torch.empty(n, dtype=torch.float32, device=rays_o.device)
    _call('ssdnerf_near_far_from_aabb', N.ptr(rays_o), N.ptr(rays_d), N.ptr(aabb), N.c_u32(n), N.c_f32(min_near),
          N.ptr(nears), N.ptr(fars), N.stream_ptr())
    return nears, fars


def batch_near_far_from_aabb(rays_o, rays_d, aabb, min_near=0.2):
    """raymarching.py:58-82: tensor (B,N,3) or list of per-scene (N_i,3)."""
    if isinstance(rays_o, torch.Tensor):
        assert rays_o.size() == rays_d.size()
        num_scenes, num_rays, _ = rays_o.size()
        nears, fars = near_far_from_aabb(rays_o.reshape(-1, 3), rays_d.reshape(-1, 3), aabb, min_near)
        return nears.reshape(num_scenes, num_rays), fars.reshape(num_scenes, num_rays)
    if len(rays_o) == 1:
        nears, fars = near_far_from_aabb(rays_o[0], rays_d[0], aabb, min_near)
        return [nears], [fars]
    sizes = [r.size(0) for r in rays_o]
    nears, fars = near_far_from_aabb(torch.cat(rays_o, dim=0), torch.cat(rays_d, dim=0), aabb, min_near)
    return nears.split(sizes), fars.split(sizes)


def sph_from_ray(rays_o, rays_d, radius):
    """raymarching.py:85-115."""
    N.require_cuda(rays_o, rays_d)
    rays_o = _f32(rays_o).view(-1, 3)
    rays_d = _f32(rays_d).view(-1, 3)
    n = rays_o.shape[0]
    coords = torch.empty(n, 2, dtype=torch.float32, device=rays_o.device)
    _call('ssdnerf_sph_from_ray', N.ptr(rays_o), N.ptr(rays_d), N.c_f32(radius), N.c_u32(n), N.ptr(coords), N.stream_ptr())
    return coords


def morton3D(coords):
    """raymarching.py:118-139: int coords [N,3] -> int32 indices [N]."""
    N.require_cuda(coords)
    coords = coords.int().contiguous()
    n = coords.shape[0]
    indices = torch.empty(n, dtype=torch.int32, device=coords.device)
    _call('ssdnerf_morton3D', N.ptr(coords), N.c_u32(n), N.ptr(indices), N.stream_ptr())
    return indices


def morton3D_invert(indices):
    """raymarching.py:142-163."""
    N.require_cuda(indices)
    indices = indices.int().contiguous()
    n = indices.shape[0]
    coords = torch.empty(n, 3, dtype=torch.int32, device=indices.device)
    _call('ssdnerf_morton3D_invert', N.ptr(indices), N.c_u32(n), N.ptr(coords), N.stream_ptr())
    return coords


def packbits(grid, thresh, bitfield=None):
    """raymarching.py:166-194: grid [C, H^3] (fp32; fp16 is read natively instead of being up-cast) -> uint8 [C*H^3/8]."""
    N.require_cuda(grid)
    if grid.dtype not in (torch.float16, torch.float32):
        grid = grid.float()
    grid = grid.contiguous()
    n = grid.numel() // 8
    if bitfield is None:
        bitfield = torch.empty(n, dtype=torch.uint8, device=grid.device)
    _call('ssdnerf_packbits', N.ptr(grid), N.c_int(int(grid.dtype == torch.float16)), N.c_u32(n), N.c_f32(float(thresh)),
          N.ptr(bitfield), N.stream_ptr())
    return bitfield


def march_rays_train(rays_o, rays_d, bound, density_bitfield, C, H, nears, fars, step_counter=None, mean_count=-1,
                     perturb=False, align=-1, force_all_rays=False, dt_gamma=0, max_steps=1024, noises=None):
    """raymarching.py:200-285. Extra kwarg `noises` injects the perturbation tensor for parity tests."""
    N.require_cuda(rays_o, rays_d, density_bitfield, nears, fars)
    rays_o = _f32(rays_o).view(-1, 3)
    rays_d = _f32(rays_d).view(-1, 3)
    density_bitfield = density_bitfield.contiguous()
    dev = rays_o.device
    n = rays_o.shape[0]
    M = n * max_steps
    if not force_all_rays and mean_count > 0:
        if align > 0:
            mean_count += align - mean_count % align
        M = mean_count
    xyzs = torch.zeros(M, 3, dtype=torch.float32, device=dev)
    dirs = torch.zeros(M, 3, dtype=torch.float32, device=dev)
    deltas = torch.zeros(M, 2, dtype=torch.float32, device=dev)
    rays = torch.empty(n, 3, dtype=torch.int32, device=dev)
    if step_counter is None:
        step_counter = torch.zeros(2, dtype=torch.int32, device=dev)
    if noises is None:
        noises = torch.rand(n, dtype=torch.float32, device=dev) if perturb else torch.zeros(n, dtype=torch.float32, device=dev)
    nears, fars, noises = _f32(nears), _f32(fars), _f32(noises)   # keep temporaries alive across the call
    _call('ssdnerf_march_rays_train', N.ptr(rays_o), N.ptr(rays_d), N.ptr(density_bitfield), N.c_f32(bound),
          N.c_f32(float(dt_gamma)), N.c_u32(max_steps), N.c_u32(n), N.c_u32(C), N.c_u32(H), N.c_u32(M),
          N.ptr(nears), N.ptr(fars), N.ptr(xyzs), N.ptr(dirs), N.ptr(deltas), N.ptr(rays),
          N.ptr(step_counter), N.ptr(noises), N.stream_ptr())
    if force_all_rays or mean_count <= 0:
        m = int(step_counter[0].item())
        if align > 0:
            m += align - m % align
        xyzs, dirs, deltas = xyzs[:m], dirs[:m], deltas[:m]
    return xyzs, dirs, deltas, rays


class _composite_rays_train(Function):
    """raymarching.py:288-343."""

    @staticmethod
    def forward(ctx, sigmas, rgbs, deltas, rays, T_thresh=1e-4):
        N.require_cuda(sigmas, rgbs, deltas, rays)
        sigmas = _f32(sigmas)
        rgbs = _f32(rgbs)
        deltas = _f32(deltas)
        rays = rays.contiguous()
        M, n = sigmas.shape[0], rays.shape[0]
        dev = sigmas.device
        weights_sum = torch.empty(n, dtype=torch.float32, device=dev)
        depth = torch.empty(n, dtype=torch.float32, device=dev)
        image = torch.empty(n, 3, dtype=torch.float32, device=dev)
        _call('ssdnerf_composite_rays_train_forward', N.ptr(sigmas), N.ptr(rgbs), N.ptr(deltas), N.ptr(rays), N.c_u32(M),
              N.c_u32(n), N.c_f32(T_thresh), N.ptr(weights_sum), N.ptr(depth), N.ptr(image), N.stream_ptr())
        ctx.save_for_backward(sigmas, rgbs, deltas, rays, weights_sum, depth, image)
        ctx.dims = [M, n, T_thresh]
        return weights_sum, depth, image

    @staticmethod
    def backward(ctx, grad_weights_sum, grad_depth, grad_image):
        sigmas, rgbs, deltas, rays, weights_sum, depth, image = ctx.saved_tensors
        M, n, T_thresh = ctx.dims
        grad_weights_sum = _f32(grad_weights_sum)
        grad_image = _f32(grad_image)
        grad_sigmas = torch.zeros_like(sigmas)
        grad_rgbs = torch.zeros_like(rgbs)
        _call('ssdnerf_composite_rays_train_backward', N.ptr(grad_weights_sum), N.ptr(grad_image), N.ptr(sigmas), N.ptr(rgbs),
              N.ptr(deltas), N.ptr(rays), N.ptr(weights_sum), N.ptr(image), N.c_u32(M), N.c_u32(n), N.c_f32(T_thresh),
              N.ptr(grad_sigmas), N.ptr(grad_rgbs), N.stream_ptr())
        return grad_sigmas, grad_rgbs, None, None, None


composite_rays_train = _composite_rays_train.apply


def _all_equal(iterable):
    g = groupby(iterable)
    return next(g, True) and not next(g, False)


def batch_composite_rays_train(sigmas, rgbs, deltas, rays, num_points, T_thresh=1e-4):
    """raymarching.py:349-395: per-scene lists of deltas / rays over one concatenated point list."""
    num_scenes = len(deltas)
    if num_scenes == 1:
        weights_sum, depth, image = composite_rays_train(sigmas, rgbs, deltas[0], rays[0], T_thresh)
        return weights_sum[None], depth[None], image[None]
    deltas_ = torch.cat(deltas, dim=0)
    rays_, num_rays = [], []
    ray_off = point_off = 0
    for ray_single, npts in zip(rays, num_points):
        rays_.append(torch.stack([ray_single[:, 0] + ray_off, ray_single[:, 1] + point_off, ray_single[:, 2]], dim=-1))
        ray_off += ray_single.size(0)
        point_off += npts
        num_rays.append(ray_single.size(0))
    rays_ = torch.cat(rays_, dim=0).int()
    weights_sum_, depth_, image_ = composite_rays_train(sigmas, rgbs, deltas_, rays_, T_thresh)
    if _all_equal(num_rays):
        return (weights_sum_.reshape(num_scenes, num_rays[0]), depth_.reshape(num_scenes, num_rays[0]),
                image_.reshape(num_scenes, num_rays[0], 3))
    return weights_sum_.split(num_rays, dim=0), depth_.split(num_rays, dim=0), image_.split(num_rays, dim=0)


def march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, density_bitfield, C, H, near, far,
               align=-1, perturb=False, dt_gamma=0, max_steps=1024, noises=None):
    """raymarching.py:402-460."""
    N.require_cuda(rays_alive, rays_t, rays_o, rays_d, density_bitfield, near, far)
    rays_o = _f32(rays_o).view(-1, 3)
    rays_d = _f32(rays_d).view(-1, 3)
    dev = rays_o.device
    M = n_alive * n_step
    if align > 0:
        M += align - (M % align)
    xyzs = torch.zeros(M, 3, dtype=torch.float32, device=dev)
    dirs = torch.zeros(M, 3, dtype=torch.float32, device=dev)
    deltas = torch.zeros(M, 2, dtype=torch.float32, device=dev)
    if noises is None and perturb:
        noises = torch.rand(n_alive, dtype=torch.float32, device=dev)
    rays_alive, density_bitfield = rays_alive.contiguous(), density_bitfield.contiguous()
    near, far = _f32(near), _f32(far)
    _call('ssdnerf_march_rays', N.c_u32(n_alive), N.c_u32(n_step), N.ptr(rays_alive), N.ptr(rays_t),
          N.ptr(rays_o), N.ptr(rays_d), N.c_f32(bound), N.c_f32(float(dt_gamma)), N.c_u32(max_steps), N.c_u32(C), N.c_u32(H),
          N.ptr(density_bitfield), N.ptr(near), N.ptr(far), N.ptr(xyzs), N.ptr(dirs), N.ptr(deltas),
          N.ptr(noises), N.stream_ptr())
    return xyzs, dirs, deltas


def composite_rays(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image, T_thresh=1e-2):
    """raymarching.py:463-489: in-place update of weights_sum / depth / image / rays_alive / rays_t."""
    N.require_cuda(rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image)
    sigmas, rgbs = _f32(sigmas), _f32(rgbs)
    _call('ssdnerf_composite_rays', N.c_u32(n_alive), N.c_u32(n_step), N.c_f32(T_thresh), N.ptr(rays_alive), N.ptr(rays_t),
          N.ptr(sigmas), N.ptr(rgbs), N.ptr(deltas), N.ptr(weights_sum), N.ptr(depth), N.ptr(image),
          N.stream_ptr())
    return tuple()
