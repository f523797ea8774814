"""Edge cases of the fused paths: empty and ragged inputs, rays that miss the volume, an empty occupancy grid, one-step budgets,
and the loud failures the C ABI promises (no CPU fallback, argument validation)."""
import numpy as np
import pytest
import torch

from oracle import render_port as rp
from oracle import train_port as tp
from tests.common import config1

pytestmark = pytest.mark.gpu


def _setup(cuda, variant='P', res=64):
    from ssdnerf_b200 import renderer as R
    vid = R.DEC_P if variant == 'P' else R.DEC_S
    code, poses, intr = config1(variant, seed=3, res=res)
    params = rp.make_decoder_params(variant, 3)
    blob = R.pack_decoder_blob(params, vid, device=cuda)
    planes = R.pack_planes(code.to(cuda), vid)
    return R, vid, code, params, blob, planes, poses, intr


@pytest.mark.parametrize('variant', ['P', 'S'])
def test_ragged_ray_count_and_misses(cuda, variant):
    """77 rays (not a multiple of the 32-ray warp tile), a third of them pointing away from the volume: per-ray sample counts exact,
    floats within the variant's tolerance, missed rays exactly zero / background"""
    R, vid, code, params, blob, planes, poses, intr = _setup(cuda, variant)
    ro, rd = rp.get_cam_rays(poses[0], intr[0], 64, 64)
    g = torch.Generator().manual_seed(0)
    sel = torch.randperm(64 * 64, generator=g)[:77]
    ro, rd = ro.reshape(-1, 3)[sel].clone(), rd.reshape(-1, 3)[sel].clone()
    rd[::3] = -rd[::3]                                              # these rays leave the box behind the camera
    bf = rp.sphere_bitfield()
    ref = rp.render_eval_scene(params, ro.numpy(), rd.numpy(), code[0], bf, max_steps=64, return_trace=True)
    out = R.render_fwd(vid, planes, (128, 128), torch.from_numpy(bf)[None].to(cuda), blob, rays_o=ro[None].to(cuda), rays_d=rd[None].to(cuda),
                       max_steps=64)
    cnt = out['num_samples'][0].cpu().numpy()
    assert np.array_equal(cnt, np.array([len(t) for t in ref['trace']], np.int32))
    tol = dict(rtol=2e-4, atol=2e-5) if variant == 'P' else dict(rtol=0, atol=1e-3)
    np.testing.assert_allclose(out['image'][0].cpu().numpy(), ref['image'], **tol)
    np.testing.assert_allclose(out['weights_sum'][0].cpu().numpy(), ref['weights_sum'], **tol)
    miss = cnt == 0
    assert miss.sum() >= 77 // 3
    assert float(out['image'][0].cpu()[torch.from_numpy(miss)].abs().max()) == 0.0
    assert torch.equal(out['rgb'][0].cpu()[torch.from_numpy(miss)], torch.ones(int(miss.sum()), 3))        # bg_color = 1


def test_empty_grid_and_empty_batches(cuda):
    R, vid, code, params, blob, planes, poses, intr = _setup(cuda)
    zero_bits = torch.zeros(1, 64 ** 3 // 8, dtype=torch.uint8, device=cuda)
    out = R.render_fwd(vid, planes, (128, 128), zero_bits, blob, poses=poses.to(cuda), intrinsics=intr.to(cuda), img_hw=(32, 32))
    assert int(out['num_samples'].abs().sum()) == 0 and float(out['weights_sum'].abs().max()) == 0.0 and float(out['depth'].abs().max()) == 0.0
    assert torch.equal(out['rgb'].cpu(), torch.ones(1, 32 * 32, 3))
    # zero rays / zero scenes: no launch, empty outputs
    e = R.render_fwd(vid, planes, (128, 128), zero_bits, blob, rays_o=torch.zeros(1, 0, 3, device=cuda), rays_d=torch.zeros(1, 0, 3, device=cuda))
    assert e['image'].shape == (1, 0, 3) and e['num_samples'].shape == (1, 0)
    from ssdnerf_b200 import raymarching as rm
    n, f = rm.near_far_from_aabb(torch.zeros(0, 3, device=cuda), torch.zeros(0, 3, device=cuda), torch.tensor([-1., -1, -1, 1, 1, 1], device=cuda), 0.2)
    assert n.numel() == 0 and f.numel() == 0
    tr = R.render_train_fwd(planes, (128, 128), zero_bits, blob, torch.zeros(1, 0, 3, device=cuda), torch.zeros(1, 0, 3, device=cuda))
    assert tr['image'].shape == (1, 0, 3)


def test_single_step_budget_and_train_ragged(cuda):
    """max_steps = 1 (one sample per ray, the coarsest legal march) and a 45-ray train-branch batch with gradient vs the oracle"""
    R, vid, code, params, blob, planes, poses, intr = _setup(cuda, res=16)
    ro, rd = rp.get_cam_rays(poses[0], intr[0], 16, 16)
    ro, rd = ro.reshape(-1, 3), rd.reshape(-1, 3)
    bf = rp.sphere_bitfield()
    bft = torch.from_numpy(bf)[None].to(cuda)
    ref = rp.render_eval_scene(params, ro.numpy(), rd.numpy(), code[0], bf, max_steps=1, return_trace=True)
    out = R.render_fwd(vid, planes, (128, 128), bft, blob, rays_o=ro[None].to(cuda), rays_d=rd[None].to(cuda), max_steps=1)
    cnt1 = np.array([len(t) for t in ref['trace']], np.int32)
    assert cnt1.max() >= 1 and np.array_equal(out['num_samples'][0].cpu().numpy(), cnt1)
    np.testing.assert_allclose(out['image'][0].cpu().numpy(), ref['image'], rtol=2e-4, atol=2e-5)
    sel = torch.linspace(0, 255, 45).long()                        # 45 rays spread over the image (not a multiple of 32), most hit the sphere
    ro45, rd45 = ro[sel].contiguous(), rd[sel].contiguous()
    noises = torch.rand(1, 45, generator=torch.Generator().manual_seed(1))
    c = code.clone().double().requires_grad_(True)
    ws, _, img = tp.render_train_scene(params, c[0], ro45.numpy(), rd45.numpy(), bf, noises[0].numpy())
    gi = torch.randn(45, 3, generator=torch.Generator().manual_seed(2))
    gref, = torch.autograd.grad((img * gi.double()).sum() + ws.sum(), c)
    fwd = R.render_train_fwd(planes, (128, 128), bft, blob, ro45[None].to(cuda), rd45[None].to(cuda), noises=noises.to(cuda))
    np.testing.assert_allclose(fwd['image'][0].cpu().numpy(), img.detach().numpy(), rtol=2e-4, atol=2e-5)
    grad = R.render_train_bwd(planes, (128, 128), bft, blob, ro45[None].to(cuda), rd45[None].to(cuda), fwd['weights_sum'], fwd['image'],
                              torch.ones(1, 45, device=cuda), gi[None].to(cuda), noises=noises.to(cuda))
    assert float(gref.norm()) > 0
    rel = float((grad.cpu().double() - gref).norm() / gref.norm())
    assert rel < 1e-3, rel


def test_loud_failures(cuda):
    """no CPU fallback and argument validation: wrong device, wrong channel count, bad grid size, misaligned K"""
    from ssdnerf_b200 import _lib as N
    from ssdnerf_b200 import unet_ops as U
    R, vid, code, params, blob, planes, poses, intr = _setup(cuda)
    bits = torch.zeros(1, 64 ** 3 // 8, dtype=torch.uint8, device=cuda)
    with pytest.raises(N.SSDNeRFNativeError):
        R.pack_planes(code, vid)                                                       # CPU tensor
    with pytest.raises(N.SSDNeRFNativeError):
        R.pack_planes(torch.randn(1, 3, 5, 128, 128, device=cuda), vid)                 # variant P needs 6 channels
    with pytest.raises(N.SSDNeRFNativeError):
        R.render_fwd(vid, planes, (128, 128), bits, blob, poses=poses.to(cuda), intrinsics=intr.to(cuda), img_hw=(32, 32), grid_size=48)
    with pytest.raises(N.SSDNeRFNativeError):
        U.linear_f16(torch.randn(128, 96, device=cuda).half(), torch.randn(64, 96, device=cuda).half())   # K not a multiple of 64 (caught by the C ABI)
    with pytest.raises(N.SSDNeRFNativeError):
        U.flash_attn(torch.randn(1, 96, 3 * 64, device=cuda).half(), 1, 0.125)          # T not a multiple of 64
