"""Fused differentiable renderer (C-ABI section 2b) and the guidance path built on it.

Forward (weights_sum / depth / image): vs the CPU oracle (rtol 2e-4 / atol 2e-5, fp32 MLP with fast intrinsics) and vs this
library's per-op composition (march_rays_train -> point_decode -> composite_rays_train), whose kernels are bit-exact with the
reference's own (tests/test_ref_gpu.py).  Per-ray sample counts are integers: exact.
Backward (d loss / d code): vs float64 autograd of the oracle chain, relative L2 <= 1e-3; vs per-op autograd relative L2 <= 1e-4
(both scatter with fp32 atomics, so not bit-identical)."""
import math

import numpy as np
import pytest
import torch

from oracle import render_port as rp
from oracle import train_port as tp
from oracle import unet_port as up
from tests.common import spiral_poses
from tests.per_op_train import per_op_train_render

pytestmark = pytest.mark.gpu

TOL_P = dict(rtol=2e-4, atol=2e-5)


def _rel_l2(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _case(B=2, n_rays=512, seed=5, res=32, density_bias=1.5):
    g = torch.Generator().manual_seed(seed)
    code = (torch.randn(B, 3, 6, 128, 128, generator=g) * 0.7).clamp(-2, 2)
    poses = torch.from_numpy(spiral_poses(B))
    f = 131.25 * res / 128
    intr = torch.tensor([f, f, res / 2, res / 2]).expand(B, 4)
    ros, rds = [], []
    for b in range(B):
        ro, rd = rp.get_cam_rays(poses[b], intr[b], res, res)
        sel = torch.randperm(res * res, generator=g)[:n_rays]
        ros.append(ro.reshape(-1, 3)[sel]); rds.append(rd.reshape(-1, 3)[sel])
    rays_o, rays_d = torch.stack(ros), torch.stack(rds)
    params = rp.make_decoder_params('P', seed)
    params['density_net.0.bias'] = params['density_net.0.bias'] + density_bias      # opaque enough for early termination
    bf = np.stack([rp.sphere_bitfield(radius=0.6), rp.sphere_bitfield(radius=0.45)][:B])
    noises = torch.rand(B, n_rays, generator=g)
    dt_gamma = torch.tensor([0.0, 0.004][:B])
    target = torch.rand(B, n_rays, 3, generator=g)
    return code, rays_o, rays_d, params, bf, noises, dt_gamma, target


def _decoder(params, cuda, frozen=True):
    import ssdnerf_b200 as S
    dec = S.build_module(dict(type='TriPlaneDecoder', base_layers=[18, 64], density_layers=[64, 1], color_layers=[64, 3],
                              use_dir_enc=True, dir_layers=[16, 64], max_steps=256))
    sd = dec.state_dict(); sd.update(params); dec.load_state_dict(sd)
    dec = dec.to(cuda).train()
    dec.requires_grad_(not frozen)
    return dec


def test_cam_rays_match_oracle(cuda):
    from ssdnerf_b200 import renderer as R
    poses = torch.from_numpy(spiral_poses(3))[None]
    intr = torch.tensor([[[40.0, 41.0, 15.5, 16.25]]]).expand(1, 3, 4).contiguous()
    ro, rd = R.get_cam_rays(poses.to(cuda), intr.to(cuda), 24, 32)
    ro_ref, rd_ref = rp.get_cam_rays(poses, intr, 24, 32)
    assert torch.equal(ro.cpu(), ro_ref)
    np.testing.assert_allclose(rd.cpu().numpy(), rd_ref.numpy(), rtol=1e-6, atol=1e-7)      # <= 1 ulp (torch matmul order)


@pytest.mark.parametrize('T_thresh', [1e-4, 0.2])
def test_train_forward_vs_oracle_and_per_op(cuda, T_thresh):
    from ssdnerf_b200 import renderer as R
    code, rays_o, rays_d, params, bf, noises, dt_gamma, _ = _case()
    B = code.shape[0]
    blob = R.pack_decoder_blob(params, R.DEC_P, device=cuda)
    planes = R.pack_planes(code.to(cuda), R.DEC_P)
    bft = torch.from_numpy(bf).to(cuda)
    out = R.render_train_fwd(planes, (128, 128), bft, blob, rays_o.to(cuda), rays_d.to(cuda), noises=noises.to(cuda),
                             dt_gamma=dt_gamma.to(cuda), T_thresh=T_thresh, want_counts=True)
    n_break = 0
    for b in range(B):
        ws, depth, img = tp.render_train_scene(params, code[b], rays_o[b].numpy(), rays_d[b].numpy(), bf[b], noises[b].numpy(),
                                               dt_gamma=float(dt_gamma[b]), T_thresh=T_thresh, dtype=torch.float32)
        np.testing.assert_allclose(out['weights_sum'][b].cpu().numpy(), ws.numpy(), **TOL_P)
        np.testing.assert_allclose(out['image'][b].cpu().numpy(), img.numpy(), **TOL_P)
        np.testing.assert_allclose(out['depth'][b].cpu().numpy(), depth.numpy(), rtol=2e-4, atol=1e-4)
        n_break += int((ws > 1 - T_thresh).sum())
    assert T_thresh < 1e-3 or n_break > 20                               # the early-termination branch is exercised
    # per-op composition (tests/per_op_train.py: same march / composite kernels as the reference, torch decode)
    with torch.no_grad():
        ref = per_op_train_render(params, rays_o.to(cuda), rays_d.to(cuda), code.to(cuda), bft, dt_gamma.tolist(), noises.to(cuda), T_thresh=T_thresh)
    for k in ('weights_sum', 'image'):
        np.testing.assert_allclose(out[k].cpu().numpy(), ref[k].cpu().numpy(), **TOL_P)
    np.testing.assert_allclose(out['depth'].cpu().numpy(), ref['depth'].cpu().numpy(), rtol=2e-4, atol=1e-4)


@pytest.mark.parametrize('T_thresh', [1e-4, 0.2])
def test_train_backward_vs_oracle_and_per_op(cuda, T_thresh):
    code, rays_o, rays_d, params, bf, noises, dt_gamma, target = _case(n_rays=256)
    B, n = rays_o.shape[:2]
    g = torch.Generator().manual_seed(1)
    g_img, g_ws = torch.randn(B, n, 3, generator=g), torch.randn(B, n, generator=g)
    # oracle: float64 autograd of sum(image * g_img + ws * g_ws)
    cref = code.clone().double().requires_grad_(True)
    tot = 0
    for b in range(B):
        ws, _, img = tp.render_train_scene(params, cref[b], rays_o[b].numpy(), rays_d[b].numpy(), bf[b], noises[b].numpy(),
                                           dt_gamma=float(dt_gamma[b]), T_thresh=T_thresh)
        tot = tot + (img * g_img[b].double()).sum() + (ws * g_ws[b].double()).sum()
    grad_ref, = torch.autograd.grad(tot, cref)
    bft = torch.from_numpy(bf).to(cuda)
    grads = {}
    for fused in (True, False):
        dec = _decoder(params, cuda)
        c = code.to(cuda).requires_grad_(True)
        if fused:
            out = dec(rays_o.to(cuda), rays_d.to(cuda), c, bft, 64, dt_gamma=dt_gamma.tolist(), perturb=noises.to(cuda), T_thresh=T_thresh)
        else:
            out = per_op_train_render(params, rays_o.to(cuda), rays_d.to(cuda), c, bft, dt_gamma.tolist(), noises.to(cuda), T_thresh=T_thresh)
        loss = (out['image'] * g_img.to(cuda)).sum() + (out['weights_sum'] * g_ws.to(cuda)).sum()
        grads[fused], = torch.autograd.grad(loss, c)
    assert float(grad_ref.abs().max()) > 0
    assert _rel_l2(grads[True], grad_ref) < 1e-3, _rel_l2(grads[True], grad_ref)
    assert _rel_l2(grads[True], grads[False]) < 1e-4, _rel_l2(grads[True], grads[False])
    # untouched texels stay exactly zero
    assert bool(((grads[True] == 0) == (grads[False] == 0)).float().mean() > 0.999)


def _model(cuda, params, test_cfg, unet_sd=None, spec=None):
    import ssdnerf_b200 as S
    unet_cfg = dict(type='DenoisingUnetMod', image_size=128, in_channels=18, base_channels=64, channels_cfg=[1, 2],
                    resblocks_per_downsample=1, dropout=0.0, use_scale_shift_norm=True, num_heads=2, attention_res=[])
    model = S.build_model(dict(
        type='DiffusionNeRF', code_size=(3, 6, 128, 128), code_reshape=(18, 128, 128), grid_size=64, bg_color=1,
        decoder_use_ema=False, diffusion_use_ema=False, freeze_decoder=True,
        pixel_loss=dict(type='MSELoss', loss_weight=20.0), reg_loss=dict(type='RegLoss', power=2, loss_weight=3e-3),
        diffusion=dict(type='GaussianDiffusion', num_timesteps=1000, betas_cfg=dict(type='linear'), denoising=unet_cfg),
        decoder=dict(type='TriPlaneDecoder', base_layers=[18, 64], density_layers=[64, 1], color_layers=[64, 3], use_dir_enc=True,
                     dir_layers=[16, 64], max_steps=256)), test_cfg=test_cfg)
    dsd = model.decoder.state_dict(); dsd.update(params); model.decoder.load_state_dict(dsd)
    if unet_sd is not None:
        model.diffusion.denoising.load_state_dict(unet_sd)
    return model.to(cuda).eval()


def test_loss_fused_vs_oracle_and_module_composition(cuda):
    """BaseNeRF.loss: fused (render + blend + MSE + RegLoss as one op) == oracle == reference-style module composition"""
    code, rays_o, rays_d, params, bf, noises, dt_gamma, target = _case(n_rays=256)
    B, n = rays_o.shape[:2]
    cfg = dict(loss_coef=0.1 / (128 * 128))
    model = _model(cuda, params, cfg)
    dec = model.decoder.train()
    bft = torch.from_numpy(bf).to(cuda)
    res = {}
    for fused in (True, False):
        c = code.to(cuda).requires_grad_(True)
        if fused:
            rgb, loss, ld = model.loss(dec, c, bft, target.to(cuda), rays_o.to(cuda), rays_d.to(cuda), dt_gamma.to(cuda), scale_num_ray=n, cfg=cfg,
                                       perturb=noises.to(cuda))
        else:       # base_nerf.py:276-296 composed from the per-op render + the loss modules
            out = per_op_train_render(params, rays_o.to(cuda), rays_d.to(cuda), c, bft, dt_gamma.tolist(), noises.to(cuda))
            rgb = out['image'] + 1.0 * (1 - out['weights_sum'].unsqueeze(-1))
            scale = 1 - math.exp(-cfg['loss_coef'] * n)
            pl = model.pixel_loss(rgb, target.to(cuda)) * (scale * 3)
            rl = model.reg_loss(c)
            loss, ld = pl + rl, dict(pixel_loss=pl, reg_loss=rl)
        grad, = torch.autograd.grad(loss * B, c)
        res[fused] = (float(loss), grad, rgb.detach(), {k: float(v) for k, v in ld.items()})
    loss_ref, grad_ref, rgb_ref = tp.render_loss_grad(params, code, rays_o.numpy(), rays_d.numpy(), target.numpy(), bf, noises=noises.numpy(),
                                                      dt_gamma=dt_gamma.numpy(), bg_color=1.0, pixel_weight=20.0, loss_coef=cfg['loss_coef'],
                                                      scale_num_ray=n, reg_weight=3e-3)
    assert set(res[True][3]) == {'pixel_loss', 'reg_loss'}
    assert abs(res[True][0] - float(loss_ref)) <= 2e-4 * abs(float(loss_ref))
    assert abs(res[True][0] - res[False][0]) <= 2e-5 * abs(res[False][0])
    np.testing.assert_allclose(res[True][2].cpu().numpy(), rgb_ref.numpy(), **TOL_P)
    assert _rel_l2(res[True][1], grad_ref * B) < 1e-3
    assert _rel_l2(res[True][1], res[False][1]) < 1e-4


@pytest.mark.parametrize('through_unet', [False, True])
def test_guided_ddim_matches_oracle(cuda, through_unet):
    """val_guide on a small UNet: render-loss guidance (gradient w.r.t. x_0, or w.r.t. x_t THROUGH the denoiser -- the reference
    default, gaussian_diffusion.py:213-216 -- on the native input-gradient pass) + langevin steps vs the oracle chain.  The UNet runs
    fp16 tensor-core GEMMs and guidance feeds its output back, so the comparison is relative-L2 on the code."""
    params = rp.make_decoder_params('P', 11)
    params['density_net.0.bias'] = params['density_net.0.bias'] + 1.5
    spec = up.unet_spec(image_size=128, in_channels=18, base_channels=64, channels_cfg=(1, 2), resblocks_per_downsample=1,
                        attention_res=(), num_heads=2)
    sd = up.random_state_dict(spec, seed=4, std=0.03)
    test_cfg = dict(num_timesteps=4, clip_range=[-2, 2], density_thresh=0.1, n_inverse_rays=2 ** 10, loss_coef=0.1 / (32 * 32),
                    guidance_gain=0.4 * (2 ** 10), snr_weight_power=0.25, grad_through_unet=through_unet, langevin_steps=1, langevin_delta=0.4,
                    langevin_t_range=[0, 600], dt_gamma_scale=0.5)
    model = _model(cuda, params, test_cfg, sd, spec)
    B, res = 1, 32
    g = torch.Generator().manual_seed(2)
    noise = torch.randn(B, 3, 6, 128, 128, generator=g)
    poses = torch.from_numpy(spiral_poses(1))[None]                          # [B,1,4,4]
    f = 131.25 * res / 128
    intr = torch.tensor([f, f, res / 2, res / 2]).expand(B, 1, 4).contiguous()
    cond_imgs = torch.rand(B, 1, res, res, 3, generator=g)

    # deterministic replicas of the random draws: langevin noise, perturb offsets and the occupancy jitter are injected
    lang = [torch.randn(B, 18, 128, 128, generator=g) for _ in range(8)]
    pert = [torch.rand(B, res * res, generator=g) for _ in range(16)]
    jit = [torch.rand(64 ** 3, 3, generator=g) for _ in range(16)]
    it = dict(p=iter(pert), j=iter(jit))
    orig_loss, orig_ues = model.loss, model.update_extra_state
    model.loss = lambda *a, **k: orig_loss(*a, **dict(k, perturb=next(it['p']).to(cuda)))
    model.update_extra_state = lambda *a, **k: orig_ues(*a, **dict(k, jitter=next(it['j']).to(cuda)))
    data = dict(cond_imgs=cond_imgs.to(cuda), cond_intrinsics=intr.to(cuda), cond_poses=poses.to(cuda), noise=noise.to(cuda))
    code_gpu, grid_gpu, bits_gpu = model.val_guide(data, langevin_noises=iter([n.to(cuda) for n in lang]))
    # oracle chain
    ro, rd = rp.get_cam_rays(poses, intr, res, res)
    dtg = (0.5 / intr[..., :2].mean(dim=(-2, -1))).numpy()
    grid = torch.zeros(B, 64 ** 3)
    it2 = dict(p=iter(pert), j=iter(jit))

    class _RenderLoss(torch.autograd.Function):
        """oracle render loss as an autograd node (value + analytic gradient from oracle/train_port.py)"""

        @staticmethod
        def forward(ctx, x0):
            code_pred = x0.detach().reshape(B, 3, 6, 128, 128)
            bits, _ = rp.update_extra_state(params, code_pred, grid, next(it2['j']), density_thresh=0.1, decay=0.9)
            with torch.enable_grad():          # autograd.Function.forward runs with grad mode off; the oracle differentiates internally
                loss, grad, _ = tp.render_loss_grad(params, code_pred, ro.reshape(B, -1, 3).numpy(), rd.reshape(B, -1, 3).numpy(),
                                                    cond_imgs.reshape(B, -1, 3).numpy(), bits, noises=next(it2['p']).numpy(), dt_gamma=dtg, bg_color=1.0,
                                                    pixel_weight=20.0, loss_coef=test_cfg['loss_coef'], scale_num_ray=res * res, reg_weight=3e-3)
            ctx.save_for_backward((grad * B).reshape(x0.shape).float())
            return torch.as_tensor(float(loss) * B)

        @staticmethod
        def backward(ctx, g_out):
            return ctx.saved_tensors[0] * g_out

    dv = up.diffusion_vars(up.linear_betas())
    ref = up.ddim_sample_guided(lambda x, t: up.unet_forward(sd, spec, x, t), noise.reshape(B, 18, 128, 128), dv, test_cfg,
                                grad_guide_fn=_RenderLoss.apply, langevin_noises=iter(lang))
    rel = _rel_l2(code_gpu.reshape(B, 18, 128, 128), ref)
    print('guided ddim (through_unet=%s) rel l2 %.3e' % (through_unet, rel))
    assert rel < 3e-3, rel
    # guidance actually moved the sample: the unguided run differs by much more than the tolerance
    model.loss, model.update_extra_state = orig_loss, orig_ues
    cfg0 = dict(test_cfg, langevin_steps=0)
    model.diffusion.test_cfg = cfg0
    plain = model.diffusion(noise.reshape(B, 18, 128, 128).to(cuda), return_loss=False)
    assert _rel_l2(plain, ref) > 5 * rel


@pytest.mark.parametrize('T_thresh', [1e-4, 0.2])
def test_decoder_weight_gradients_vs_oracle(cuda, T_thresh):
    """trainable decoder (stage-1 training, multiscene_nerf.py:203-207): d loss / d decoder weights from the same fused backward
    launch vs float64 autograd of the oracle chain w.r.t. its decoder parameters -- relative L2 <= 1e-3 per parameter; the code
    gradient produced alongside must equal the frozen-decoder launch's (same arithmetic, atomics reorder only)."""
    from ssdnerf_b200 import renderer as R
    code, rays_o, rays_d, params, bf, noises, dt_gamma, target = _case(n_rays=256)
    B, n = rays_o.shape[:2]
    g = torch.Generator().manual_seed(2)
    g_img, g_ws = torch.randn(B, n, 3, generator=g), torch.randn(B, n, generator=g)
    pref = {k: torch.as_tensor(v).double().requires_grad_(True) for k, v in params.items()}
    tot = 0
    for b in range(B):
        ws, _, img = tp.render_train_scene(pref, code[b].double(), rays_o[b].numpy(), rays_d[b].numpy(), bf[b], noises[b].numpy(),
                                           dt_gamma=float(dt_gamma[b]), T_thresh=T_thresh)
        tot = tot + (img * g_img[b].double()).sum() + (ws * g_ws[b].double()).sum()
    names = list(R.DEC_P_PARAM_ORDER)
    gref = dict(zip(names, torch.autograd.grad(tot, [pref[k] for k in names])))
    bft = torch.from_numpy(bf).to(cuda)
    code_grads = {}
    for frozen in (False, True):
        dec = _decoder(params, cuda, frozen=frozen)
        c = code.to(cuda).requires_grad_(True)
        out = dec(rays_o.to(cuda), rays_d.to(cuda), c, bft, 64, dt_gamma=dt_gamma.tolist(), perturb=noises.to(cuda), T_thresh=T_thresh)
        loss = (out['image'] * g_img.to(cuda)).sum() + (out['weights_sum'] * g_ws.to(cuda)).sum()
        loss.backward()
        code_grads[frozen] = c.grad.clone()
        if not frozen:
            got = dict(dec.named_parameters())
            for k in names:
                err = _rel_l2(got[k].grad, gref[k])
                print(f'T_thresh {T_thresh} {k}: rel L2 {err:.2e} (|ref| {float(gref[k].norm()):.3e})')
                assert got[k].grad.shape == gref[k].shape and float(gref[k].abs().max()) > 0
                assert err < 1e-3, (k, err)
        else:
            assert all(p.grad is None for p in dec.parameters())
    assert _rel_l2(code_grads[False], code_grads[True]) < 1e-5


def test_loss_with_trainable_decoder_vs_oracle(cuda):
    """BaseNeRF.loss on its fused route (MSELoss + RegLoss power 2) with a TRAINABLE decoder: loss, code gradient and decoder-weight
    gradients vs the float64 oracle chain"""
    from ssdnerf_b200 import renderer as R
    code, rays_o, rays_d, params, bf, noises, dt_gamma, target = _case(n_rays=256)
    model = _model(cuda, params, dict())
    model.decoder.requires_grad_(True)
    model.decoder.train()
    kw = dict(loss_coef=0.5 / (128 * 128), scale_num_ray=128 * 128, pixel_weight=20.0, reg_weight=3e-3)
    pref = {k: torch.as_tensor(v).double().requires_grad_(True) for k, v in params.items()}
    cref = code.clone().double().requires_grad_(True)
    loss_ref, _ = tp.render_loss(pref, cref, rays_o.numpy(), rays_d.numpy(), target.numpy(), bf, noises=noises.numpy(), dt_gamma=dt_gamma.numpy(),
                                 bg_color=1.0, **kw)
    names = list(R.DEC_P_PARAM_ORDER)
    grads_ref = torch.autograd.grad(loss_ref, [cref] + [pref[k] for k in names])
    c = code.to(cuda).requires_grad_(True)
    _, loss, _ = model.loss(model.decoder, c, torch.from_numpy(bf).to(cuda), target.to(cuda), rays_o.to(cuda), rays_d.to(cuda),
                            dt_gamma=dt_gamma.to(cuda), scale_num_ray=128 * 128, cfg=dict(loss_coef=0.5 / (128 * 128)), perturb=noises.to(cuda))
    loss.backward()
    assert abs(float(loss) - float(loss_ref)) < 1e-4 * abs(float(loss_ref))
    assert _rel_l2(c.grad, grads_ref[0]) < 1e-3
    got = dict(model.decoder.named_parameters())
    for k, gr in zip(names, grads_ref[1:]):
        assert _rel_l2(got[k].grad, gr) < 1e-3, (k, _rel_l2(got[k].grad, gr))
