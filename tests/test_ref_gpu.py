"""Parity against the REFERENCE'S OWN CUDA kernels (lib/ops/raymarching, lib/ops/shencoder compiled unmodified except
-std=c++17 into oracle/_ref by oracle/build_ref.sh) on the same GPU, same inputs.  Integer / index outputs and the
marcher's floats are compared bit-exactly; compositor floats to 1e-6 (both use MUFU.EX2 based __expf).
Skipped when oracle/_ref was not built (it is built in the container that has /root/reference and travels to the GPU box)."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import render_port as rp
from tests.common import spiral_poses

pytestmark = pytest.mark.gpu
_REF_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'oracle', '_ref')


@pytest.fixture(scope='module')
def ref():
    if not os.path.isdir(_REF_DIR) or not any(f.startswith('_raymarching') for f in os.listdir(_REF_DIR)):
        pytest.skip('oracle/_ref not built')
    sys.path.insert(0, _REF_DIR)
    import _raymarching
    import _shencoder
    return _raymarching, _shencoder


def _bits(t):
    return t.detach().cpu().contiguous().numpy().view(np.uint32)


def _rays(cuda, res=64, V=2):
    poses = torch.from_numpy(spiral_poses(V))
    intr = torch.tensor([res * 131.25 / 128, res * 131.25 / 128, res / 2, res / 2]).expand(V, 4).contiguous()
    ro, rd = rp.get_cam_rays(poses, intr, res, res)
    return ro.reshape(-1, 3).contiguous().to(cuda), rd.reshape(-1, 3).contiguous().to(cuda)


def test_utils_bit_exact(cuda, ref):
    rmref, _ = ref
    from ssdnerf_b200 import raymarching as rm
    ro, rd = _rays(cuda)
    N = ro.shape[0]
    aabb = torch.tensor([-1., -1, -1, 1, 1, 1], device=cuda)
    n_ref, f_ref = torch.empty(N, device=cuda), torch.empty(N, device=cuda)
    rmref.near_far_from_aabb(ro, rd, aabb, N, 0.2, n_ref, f_ref)
    n, f = rm.near_far_from_aabb(ro, rd, aabb, 0.2)
    assert np.array_equal(_bits(n), _bits(n_ref)) and np.array_equal(_bits(f), _bits(f_ref))
    g = torch.Generator().manual_seed(0)
    coords = torch.randint(0, 128, (50000, 3), generator=g, dtype=torch.int32).to(cuda)
    idx_ref = torch.empty(50000, dtype=torch.int32, device=cuda)
    rmref.morton3D(coords, 50000, idx_ref)
    assert torch.equal(rm.morton3D(coords), idx_ref)
    grid = torch.rand(2, 64 ** 3, generator=g).to(cuda)
    for gq in (grid, grid.half()):
        bits_ref = torch.empty(2 * 64 ** 3 // 8, dtype=torch.uint8, device=cuda)
        rmref.packbits(gq.contiguous(), 2 * 64 ** 3 // 8, 0.41, bits_ref)
        assert torch.equal(rm.packbits(gq, 0.41), bits_ref)


@pytest.mark.parametrize('dt_gamma', [0.0, 0.0078125])
def test_march_and_composite_bit_exact(cuda, ref, dt_gamma):
    rmref, _ = ref
    from ssdnerf_b200 import raymarching as rm
    ro, rd = _rays(cuda)
    N = ro.shape[0]
    aabb = torch.tensor([-1., -1, -1, 1, 1, 1], device=cuda)
    nears, fars = rm.near_far_from_aabb(ro, rd, aabb, 0.2)
    bf = torch.from_numpy(rp.sphere_bitfield()).to(cuda)
    alive = torch.arange(N, dtype=torch.int32, device=cuda)
    rays_t = nears.clone()
    n_step = 3
    M = N * n_step
    M += 128 - M % 128
    outs_ref = [torch.zeros(M, 3, device=cuda), torch.zeros(M, 3, device=cuda), torch.zeros(M, 2, device=cuda)]
    rmref.march_rays(N, n_step, alive, rays_t, ro, rd, 1.0, dt_gamma, 256, 1, 64, bf, nears, fars, *outs_ref, torch.zeros(N, device=cuda))
    outs = rm.march_rays(N, n_step, alive, rays_t, ro, rd, 1.0, bf, 1, 64, nears, fars, align=128, dt_gamma=dt_gamma, max_steps=256)
    for a, b in zip(outs, outs_ref):
        assert np.array_equal(_bits(a), _bits(b))
    g = torch.Generator().manual_seed(1)
    sig = (torch.rand(M, generator=g) * 30).to(cuda)
    rgb = torch.rand(M, 3, generator=g).to(cuda)
    st_ref = [alive.clone(), rays_t.clone(), torch.zeros(N, device=cuda), torch.zeros(N, device=cuda), torch.zeros(N, 3, device=cuda)]
    st = [t.clone() for t in st_ref]
    rmref.composite_rays(N, n_step, 1e-4, st_ref[0], st_ref[1], sig, rgb, outs_ref[2], st_ref[2], st_ref[3], st_ref[4])
    rm.composite_rays(N, n_step, st[0], st[1], sig, rgb, outs[2], st[2], st[3], st[4], 1e-4)
    assert torch.equal(st[0], st_ref[0]) and np.array_equal(_bits(st[1]), _bits(st_ref[1]))
    for a, b in zip(st[2:], st_ref[2:]):
        torch.testing.assert_close(a, b, rtol=1e-6, atol=1e-7)


def test_march_train_and_composite_train(cuda, ref):
    rmref, _ = ref
    from ssdnerf_b200 import raymarching as rm
    ro, rd = _rays(cuda, res=32)
    N = ro.shape[0]
    aabb = torch.tensor([-1., -1, -1, 1, 1, 1], device=cuda)
    nears, fars = rm.near_far_from_aabb(ro, rd, aabb, 0.2)
    bf = torch.from_numpy(rp.sphere_bitfield()).to(cuda)
    g = torch.Generator().manual_seed(2)
    noises = torch.rand(N, generator=g).to(cuda)
    M = N * 256
    r_xyz, r_dir, r_del = torch.zeros(M, 3, device=cuda), torch.zeros(M, 3, device=cuda), torch.zeros(M, 2, device=cuda)
    r_rays, r_cnt = torch.empty(N, 3, dtype=torch.int32, device=cuda), torch.zeros(2, dtype=torch.int32, device=cuda)
    rmref.march_rays_train(ro, rd, bf, 1.0, 0.0, 256, N, 1, 64, M, nears, fars, r_xyz, r_dir, r_del, r_rays, r_cnt, noises)
    xyz, dirs, deltas, rays = rm.march_rays_train(ro, rd, 1.0, bf, 1, 64, nears, fars, perturb=True, force_all_rays=True,
                                                  dt_gamma=0.0, max_steps=256, noises=noises)
    a, b = rays.cpu().numpy(), r_rays.cpu().numpy()
    a, b = a[np.argsort(a[:, 0])], b[np.argsort(b[:, 0])]
    assert np.array_equal(a[:, [0, 2]], b[:, [0, 2]]) and int(r_cnt[0]) == xyz.shape[0]
    xyz_n, rx_n, del_n, rd_n = xyz.cpu().numpy(), r_xyz.cpu().numpy(), deltas.cpu().numpy(), r_del.cpu().numpy()
    for i in range(0, N, 5):
        c = a[i, 2]
        assert np.array_equal(xyz_n[a[i, 1]:a[i, 1] + c].view(np.uint32), rx_n[b[i, 1]:b[i, 1] + c].view(np.uint32))
        assert np.array_equal(del_n[a[i, 1]:a[i, 1] + c].view(np.uint32), rd_n[b[i, 1]:b[i, 1] + c].view(np.uint32))
    # compositing forward/backward on identical inputs
    m = xyz.shape[0]
    sig, rgb = (torch.rand(m, generator=g) * 20).to(cuda), torch.rand(m, 3, generator=g).to(cuda)
    ws_r, dep_r, img_r = torch.empty(N, device=cuda), torch.empty(N, device=cuda), torch.empty(N, 3, device=cuda)
    rmref.composite_rays_train_forward(sig, rgb, deltas, rays, m, N, 1e-4, ws_r, dep_r, img_r)
    sig_g, rgb_g = sig.clone().requires_grad_(True), rgb.clone().requires_grad_(True)
    ws, dep, img = rm.composite_rays_train(sig_g, rgb_g, deltas, rays, 1e-4)
    torch.testing.assert_close(ws, ws_r, rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(img, img_r, rtol=1e-6, atol=1e-7)
    gws, gimg = torch.rand(N, generator=g).to(cuda), torch.rand(N, 3, generator=g).to(cuda)
    gs_r, gc_r = torch.zeros(m, device=cuda), torch.zeros(m, 3, device=cuda)
    rmref.composite_rays_train_backward(gws, gimg, sig, rgb, deltas, rays, ws_r, img_r, m, N, 1e-4, gs_r, gc_r)
    torch.autograd.backward([ws, img], [gws, gimg])
    torch.testing.assert_close(sig_g.grad, gs_r, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(rgb_g.grad, gc_r, rtol=1e-6, atol=1e-7)


def test_sh_encode_bit_exact(cuda, ref):
    _, shref = ref
    from ssdnerf_b200.shencoder import sh_encode
    g = torch.Generator().manual_seed(3)
    d = torch.nn.functional.normalize(torch.randn(10000, 3, generator=g), dim=-1).to(cuda)
    out_ref, dy_ref = torch.empty(10000, 16, device=cuda), torch.empty(10000, 48, device=cuda)
    shref.sh_encode_forward(d, out_ref, 10000, 3, 4, True, dy_ref)
    out = sh_encode(d, 4, False)
    assert np.array_equal(_bits(out), _bits(out_ref))


def _reference_eval_loop(rmref, shref, params, ro, rd, code_single, bf, max_steps=256, T_thresh=1e-4):
    """the reference's eval branch (base_volume_renderer.py:79-123) driven with the reference's own kernels + torch decode"""
    dev = ro.device
    N = ro.shape[0]
    aabb = torch.tensor([-1., -1, -1, 1, 1, 1], device=dev)
    nears, fars = torch.empty(N, device=dev), torch.empty(N, device=dev)
    rmref.near_far_from_aabb(ro, rd, aabb, N, 0.2, nears, fars)
    ws, dep, img = torch.zeros(N, device=dev), torch.zeros(N, device=dev), torch.zeros(N, 3, device=dev)
    alive = torch.arange(N, dtype=torch.int32, device=dev)
    rays_t = nears.clone()
    p = {k: v.to(dev) for k, v in params.items()}
    step = 0
    F = torch.nn.functional
    while step < max_steps:
        n_alive = alive.shape[0]
        if n_alive == 0:
            break
        n_step = min(max(N // n_alive, 1), 8)
        M = n_alive * n_step
        M += 128 - M % 128
        xyzs, dirs, deltas = torch.zeros(M, 3, device=dev), torch.zeros(M, 3, device=dev), torch.zeros(M, 2, device=dev)
        rmref.march_rays(n_alive, n_step, alive, rays_t, ro, rd, 1.0, 0.0, max_steps, 1, 64, bf, nears, fars, xyzs, dirs, deltas,
                         torch.zeros(n_alive, device=dev))
        pc = F.grid_sample(code_single, rp.xyz_transform(xyzs), mode='bilinear', padding_mode='border', align_corners=False).squeeze(-2)
        pc = pc.permute(2, 1, 0).reshape(M, -1)
        base_x = F.linear(pc, p['base_net.0.weight'], p['base_net.0.bias'])
        sig = torch.exp(F.linear(F.silu(base_x), p['density_net.0.weight'], p['density_net.0.bias'])).squeeze(-1)
        sh = torch.empty(M, 16, device=dev)
        shref.sh_encode_forward(dirs.contiguous(), sh, M, 3, 4, False, torch.empty(1, device=dev))
        rgb = torch.sigmoid(F.linear(F.silu(base_x + F.linear(sh, p['dir_net.0.weight'], p['dir_net.0.bias'])),
                                     p['color_net.0.weight'], p['color_net.0.bias'])) * 1.002 - 0.001
        rmref.composite_rays(n_alive, n_step, T_thresh, alive, rays_t, sig.contiguous(), rgb.contiguous(), deltas, ws, dep, img)
        alive = alive[alive >= 0]
        step += n_step
    return ws, dep, img


def test_fused_renderer_vs_reference_pipeline(cuda, ref):
    """whole eval renderer: reference kernels + PyTorch decode (fp32, TF32 off) vs the fused kernel, variant P"""
    rmref, shref = ref
    from ssdnerf_b200 import renderer as R
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator().manual_seed(4)
    code = torch.randn(1, 3, 6, 128, 128, generator=g).clamp(-2, 2).to(cuda)
    params = rp.make_decoder_params('P', 5)
    ro, rd = _rays(cuda, res=64, V=2)
    bf = torch.from_numpy(rp.sphere_bitfield()).to(cuda)
    ws_r, dep_r, img_r = _reference_eval_loop(rmref, shref, params, ro, rd, code[0], bf)
    out = R.render_fwd(R.DEC_P, R.pack_planes(code, R.DEC_P), (128, 128), bf[None], R.pack_decoder_blob(params, R.DEC_P, device=cuda),
                       rays_o=ro[None], rays_d=rd[None], max_steps=256)
    torch.testing.assert_close(out['weights_sum'][0], ws_r, rtol=2e-4, atol=2e-5)
    torch.testing.assert_close(out['image'][0], img_r, rtol=2e-4, atol=2e-5)
    torch.testing.assert_close(out['depth'][0], dep_r, rtol=2e-4, atol=1e-4)
