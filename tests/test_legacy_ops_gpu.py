"""GPU parity of the per-op C-ABI entry points against the CPU oracle (bit-exact for integer outputs)."""
import numpy as np
import pytest
import torch

import oracle as orc
from oracle import render_port as rp
from tests.common import config1

pytestmark = pytest.mark.gpu


def _rays(res=64):
    code, poses, intr = config1('P', res=res)
    ro, rd = rp.get_cam_rays(poses[0], intr[0], res, res)
    return ro.reshape(-1, 3).numpy().copy(), rd.reshape(-1, 3).numpy().copy()


def test_near_far_bit_exact(cuda):
    from ssdnerf_b200 import raymarching as rm
    ro, rd = _rays()
    # add axis-parallel and missing rays
    rd[0] = [1, 0, 0]; rd[1] = [0, -1, 0]; ro[2] = [5, 5, 5]; rd[2] = [0, 0, 1]
    aabb = np.array([-1, -1, -1, 1, 1, 1], np.float32)
    n_ref, f_ref = orc.near_far_from_aabb(ro, rd, aabb, 0.2)
    n, f = rm.near_far_from_aabb(torch.from_numpy(ro).to(cuda), torch.from_numpy(rd).to(cuda), torch.from_numpy(aabb).to(cuda), 0.2)
    assert np.array_equal(n.cpu().numpy().view(np.uint32), n_ref.view(np.uint32))
    assert np.array_equal(f.cpu().numpy().view(np.uint32), f_ref.view(np.uint32))


def test_morton_roundtrip_and_packbits(cuda):
    from ssdnerf_b200 import raymarching as rm
    g = torch.Generator().manual_seed(1)
    coords = torch.randint(0, 128, (100000, 3), generator=g, dtype=torch.int32)
    idx = rm.morton3D(coords.to(cuda))
    assert np.array_equal(idx.cpu().numpy(), orc.morton3D(coords.numpy()))
    back = rm.morton3D_invert(idx)
    assert torch.equal(back.cpu(), coords)
    grid = torch.rand(2, 64 ** 3, generator=g)
    for dt in (torch.float32, torch.float16):
        gq = grid.to(dt)
        bits = rm.packbits(gq.to(cuda), 0.37)
        assert np.array_equal(bits.cpu().numpy(), orc.packbits(gq.float().numpy().reshape(-1), 0.37))


@pytest.mark.parametrize('dt_gamma', [0.0, 0.01])
def test_march_and_composite_rays(cuda, dt_gamma):
    from ssdnerf_b200 import raymarching as rm
    ro, rd = _rays()
    N = ro.shape[0]
    bf = rp.sphere_bitfield()
    aabb = np.array([-1, -1, -1, 1, 1, 1], np.float32)
    nears, fars = orc.near_far_from_aabb(ro, rd, aabb, 0.2)
    alive = np.arange(N, dtype=np.int32)
    rays_t = nears.copy()
    n_step = 4
    xyz_r, dir_r, del_r = orc.march_rays(N, n_step, alive, rays_t, ro, rd, 1.0, bf, 1, 64, nears, fars, align=128,
                                         dt_gamma=dt_gamma, max_steps=256)
    T = lambda a: torch.from_numpy(a).to(cuda)
    xyz, dirs, deltas = rm.march_rays(N, n_step, T(alive), T(rays_t), T(ro), T(rd), 1.0, T(bf), 1, 64, T(nears), T(fars),
                                      align=128, dt_gamma=dt_gamma, max_steps=256)
    for a, b in ((xyz, xyz_r), (dirs, dir_r), (deltas, del_r)):
        assert np.array_equal(a.cpu().numpy().view(np.uint32), b.view(np.uint32))
    # composite with random sigmas / rgbs
    g = torch.Generator().manual_seed(3)
    sig = torch.rand(xyz_r.shape[0], generator=g) * 20
    rgb = torch.rand(xyz_r.shape[0], 3, generator=g)
    ws = np.zeros(N, np.float32); dep = np.zeros(N, np.float32); img = np.zeros((N, 3), np.float32)
    alive_r, t_r = alive.copy(), rays_t.copy()
    orc.composite_rays(N, n_step, alive_r, t_r, sig.numpy(), rgb.numpy(), del_r, ws, dep, img, 1e-4)
    ws_g, dep_g, img_g = torch.zeros(N, device=cuda), torch.zeros(N, device=cuda), torch.zeros(N, 3, device=cuda)
    alive_g, t_g = T(alive), T(rays_t)
    rm.composite_rays(N, n_step, alive_g, t_g, sig.to(cuda), rgb.to(cuda), deltas, ws_g, dep_g, img_g, 1e-4)
    assert np.array_equal(alive_g.cpu().numpy(), alive_r)
    assert np.array_equal(t_g.cpu().numpy().view(np.uint32), t_r.view(np.uint32))
    np.testing.assert_allclose(ws_g.cpu().numpy(), ws, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(img_g.cpu().numpy(), img, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(dep_g.cpu().numpy(), dep, rtol=1e-5, atol=1e-6)


def test_march_train_and_composite_train(cuda):
    from ssdnerf_b200 import raymarching as rm
    ro, rd = _rays(32)
    N = ro.shape[0]
    bf = rp.sphere_bitfield()
    aabb = np.array([-1, -1, -1, 1, 1, 1], np.float32)
    nears, fars = orc.near_far_from_aabb(ro, rd, aabb, 0.2)
    g = torch.Generator().manual_seed(5)
    noises = torch.rand(N, generator=g).numpy()
    xyz_r, dir_r, del_r, rays_r = orc.march_rays_train(ro, rd, 1.0, bf, 1, 64, nears, fars, dt_gamma=0.0, max_steps=256,
                                                       noises=noises, align=128)
    T = lambda a: torch.from_numpy(a).to(cuda)
    xyz, dirs, deltas, rays = rm.march_rays_train(T(ro), T(rd), 1.0, T(bf), 1, 64, T(nears), T(fars), perturb=True, align=128,
                                                  force_all_rays=True, dt_gamma=0.0, max_steps=256, noises=T(noises))
    rays = rays.cpu().numpy(); xyz = xyz.cpu().numpy(); deltas_np = deltas.cpu().numpy()
    # offsets are atomics-ordered on the GPU: compare per ray
    order = np.argsort(rays[:, 0])
    rays_s = rays[order]
    assert np.array_equal(rays_s[:, 0], rays_r[:, 0])
    assert np.array_equal(rays_s[:, 2], rays_r[:, 2])
    assert xyz.shape == xyz_r.shape
    for i in range(0, N, 7):
        o_g, o_r, c = rays_s[i, 1], rays_r[i, 1], rays_r[i, 2]
        assert np.array_equal(xyz[o_g:o_g + c].view(np.uint32), xyz_r[o_r:o_r + c].view(np.uint32))
        assert np.array_equal(deltas_np[o_g:o_g + c].view(np.uint32), del_r[o_r:o_r + c].view(np.uint32))
    # forward / backward compositing on the GPU's own layout vs oracle on the same layout
    M = xyz.shape[0]
    sig = (torch.rand(M, generator=g) * 15)
    rgb = torch.rand(M, 3, generator=g)
    ws_r, dep_r, img_r = orc.composite_rays_train_forward(sig.numpy(), rgb.numpy(), deltas_np, rays, 1e-4)
    sig_g = sig.to(cuda).requires_grad_(True); rgb_g = rgb.to(cuda).requires_grad_(True)
    ws, dep, img = rm.composite_rays_train(sig_g, rgb_g, deltas, T(rays), 1e-4)
    np.testing.assert_allclose(ws.detach().cpu().numpy(), ws_r, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(img.detach().cpu().numpy(), img_r, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(dep.detach().cpu().numpy(), dep_r, rtol=1e-5, atol=1e-6)
    gws = torch.rand(N, generator=g); gimg = torch.rand(N, 3, generator=g)
    (ws * gws.to(cuda)).sum().add((img * gimg.to(cuda)).sum()).backward()
    gs_r, gc_r = orc.composite_rays_train_backward(gws.numpy(), gimg.numpy(), sig.numpy(), rgb.numpy(), deltas_np, rays, ws_r, img_r, 1e-4)
    np.testing.assert_allclose(sig_g.grad.cpu().numpy(), gs_r, rtol=2e-4, atol=1e-5)
    np.testing.assert_allclose(rgb_g.grad.cpu().numpy(), gc_r, rtol=1e-5, atol=1e-6)


def test_sh_encode(cuda):
    from ssdnerf_b200.shencoder import SHEncoder
    g = torch.Generator().manual_seed(2)
    d = torch.nn.functional.normalize(torch.randn(5000, 3, generator=g), dim=-1)
    ref = orc.sh_encode(d.numpy(), 4)
    enc = SHEncoder()
    out = enc(d.to(cuda))
    np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=1e-5, atol=1e-6)
    # Jacobian vs double-precision finite differences
    dd = d[:64].double()
    x = dd.clone().to(cuda).float().requires_grad_(True)
    w = torch.randn(16, generator=g).to(cuda)
    (enc(x) * w).sum().backward()
    eps = 1e-4
    num = torch.zeros(64, 3, dtype=torch.float64)
    for k in range(3):
        e = torch.zeros(3, dtype=torch.float64); e[k] = eps
        fp = torch.from_numpy(orc.sh_encode((dd + e).float().numpy(), 4)).double()
        fm = torch.from_numpy(orc.sh_encode((dd - e).float().numpy(), 4)).double()
        num[:, k] = ((fp - fm) / (2 * eps) * w.cpu().double()).sum(-1)
    np.testing.assert_allclose(x.grad.cpu().double().numpy(), num.numpy(), rtol=5e-2, atol=5e-2)
