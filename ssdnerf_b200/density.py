"""Occupancy-grid builder front-end (C ABI section 3): replaces ``BaseNeRF.update_extra_state`` / ``get_density``
(lib/models/autodecoders/base_nerf.py:318-401) with two launches per iteration and no host sync."""
import torch

from . import _lib as N


def _workspace(num_scenes, grid_size, device):
    nbytes = N.lib().ssdnerf_density_workspace_bytes(N.c_u32(num_scenes), N.c_u32(grid_size))
    return torch.empty(nbytes, dtype=torch.uint8, device=device)


def update_extra_state(variant, planes, plane_hw, blob, density_grid, density_bitfield, jitter=None,
                       density_thresh=0.01, decay=0.9, grid_size=64, bound=1.0, workspace=None, thresh_out=None):
    """In-place full update of `density_grid` [B, G^3] (fp16 or fp32, morton order) and `density_bitfield` [B, G^3/8].

    jitter: the `torch.rand_like(xyzs)` tensor of base_nerf.py:344, shape [G^3, 3] (ij-meshgrid order), drawn on the
    device if None."""
    N.require_cuda(planes, blob, density_grid, density_bitfield)
    B = density_grid.shape[0]
    dev = density_grid.device
    assert density_grid.is_contiguous() and density_bitfield.is_contiguous()
    assert density_grid.dtype in (torch.float16, torch.float32)
    if jitter is None:
        jitter = torch.rand(grid_size ** 3, 3, device=dev)
    jitter = jitter.contiguous().float()
    if workspace is None:
        workspace = _workspace(B, grid_size, dev)
    is_half = int(density_grid.dtype == torch.float16)
    L = N.lib()
    N.check(L.ssdnerf_density_update(N.c_int(variant), N.ptr(planes), N.c_u32(plane_hw[0]), N.c_u32(plane_hw[1]), N.ptr(blob),
                                     N.c_u32(B), N.c_u32(grid_size), N.c_f32(bound), N.ptr(jitter), N.c_f32(decay),
                                     N.ptr(density_grid), N.c_int(is_half), N.ptr(workspace), N.stream_ptr()))
    N.check(L.ssdnerf_density_pack(N.ptr(density_grid), N.c_int(is_half), N.c_u32(B), N.c_u32(grid_size),
                                   N.c_f32(density_thresh), N.ptr(density_bitfield), N.ptr(thresh_out), N.ptr(workspace),
                                   N.stream_ptr()))


def get_density(variant, planes, plane_hw, blob, num_scenes, density_thresh=0.01, density_step=8, grid_size=64, bound=1.0,
                jitters=None, grid_dtype=torch.float16):
    """base_nerf.py:391-401: zero-initialised grid, `density_step` full updates with decay 1.0."""
    dev = planes.device
    grid = torch.zeros(num_scenes, grid_size ** 3, dtype=grid_dtype, device=dev)
    bitfield = torch.zeros(num_scenes, grid_size ** 3 // 8, dtype=torch.uint8, device=dev)
    ws = _workspace(num_scenes, grid_size, dev)
    for i in range(density_step):
        update_extra_state(variant, planes, plane_hw, blob, grid, bitfield, None if jitters is None else jitters[i],
                           density_thresh, 1.0, grid_size, bound, ws)
    return grid, bitfield
