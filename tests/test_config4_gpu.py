"""BASELINE config 4 (`ssdnerf_chairs_recons1v`: cond_mode='guide_optim') through the public entry point, and its code-optimisation
half (`val_optim`: diffusion-prior gradient through the UNet + inner render-loss steps with Adam) against the oracle chain."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import render_port as rp
from oracle import train_port as tp
from oracle import unet_port as up
from tests.common import GOLDEN, spiral_poses

pytestmark = pytest.mark.gpu


def _rel_l2(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def test_reference_config4_runs_unchanged(cuda):
    """the reference's own config (resolved fixture), full-size UNet, 8 -> 2 scenes and fewer steps so the test stays short:
    guided DDIM + langevin through the UNet backward, then val_optim, then the 251 -> 3 view render"""
    import ssdnerf_b200 as S
    c = json.load(open(os.path.join(GOLDEN, 'reference_configs.json')))['configs/paper_cfgs/ssdnerf_chairs_recons1v.py']
    assert c['test_cfg']['cond_mode'] == 'guide_optim' and 'grad_through_unet' not in c['test_cfg']        # reference default: through the UNet
    test_cfg = dict(c['test_cfg'], num_timesteps=3, langevin_steps=1, n_inverse_steps=2, extra_scene_step=1)
    torch.manual_seed(0)
    model = S.build_model(c['model'], train_cfg=c['train_cfg'], test_cfg=test_cfg)
    g = torch.Generator().manual_seed(0)
    for mod in (model.diffusion_ema.denoising,):
        for name, p in mod.named_parameters():
            if p.dim() > 1:
                p.data.copy_(torch.randn(p.shape, generator=g) * 0.02)
    model = model.to(cuda).eval()
    assert model.diffusion_ema.ddpm_loss.weight_scale == 1.0            # test_cfg.override_cfg applied by eval()
    B, res = 2, 128
    poses = torch.from_numpy(spiral_poses(4))[None].repeat(B, 1, 1, 1).to(cuda)
    intr = torch.tensor([131.25, 131.25, 64.0, 64.0]).expand(B, 4, 4).contiguous().to(cuda)
    # synthetic conditioning view: render of a random triplane
    code0 = (torch.randn(B, 3, 6, 128, 128, generator=g) * 0.5).to(cuda)
    with torch.no_grad():
        _, bits0 = model.get_density(model.decoder_ema, code0, cfg=dict(density_thresh=0.1))
        img0, _ = model.render(model.decoder_ema, code0, bits0, res, res, intr[:, :1].contiguous(), poses[:, :1].contiguous(), cfg=model.test_cfg)
    data = dict(scene_id=[0, 1], scene_name=['a', 'b'], cond_imgs=img0, cond_intrinsics=intr[:, :1].contiguous(), cond_poses=poses[:, :1].contiguous(),
                test_poses=poses[:, 1:].contiguous(), test_intrinsics=intr[:, 1:].contiguous(),
                noise=torch.randn(B, 3, 6, 128, 128, generator=g).to(cuda))
    out = model.val_step(data)
    assert out['num_samples'] == B and out['pred_imgs'].shape == (B, 3, 3, res, res)
    assert torch.isfinite(out['pred_imgs']).all() and float(out['pred_imgs'].std()) > 0
    # the other public branches of diffusion_nerf.py:406-469
    model.test_cfg['cond_mode'] = 'guide'
    assert model.val_step(data)['pred_imgs'].shape == (B, 3, 3, res, res)
    model.test_cfg['cond_mode'] = 'optim'
    assert model.val_step(data)['pred_imgs'].shape == (B, 3, 3, res, res)
    stored = [dict(param=dict(code=code0[i].cpu(), density_grid=torch.zeros(64 ** 3).half(), density_bitfield=bits0[i].cpu())) for i in range(B)]
    o2 = model.val_step(dict(scene_id=[0, 1], scene_name=['a', 'b'], code=stored, test_poses=poses[:, :1].contiguous(), test_intrinsics=intr[:, :1].contiguous()))
    ref8 = torch.round(img0.permute(0, 1, 4, 2, 3).clamp(0, 1) * 255) / 255
    assert torch.equal(o2['pred_imgs'], ref8)


def test_val_optim_matches_oracle_chain(cuda):
    """diffusion-prior gradient (UNet input-gradient pass, SNR-weighted v-loss) + `extra_scene_step`+1 inner render-loss Adam steps that
    start from that prior gradient, 2 outer steps, on a small UNet; timestep draws, noises, perturb offsets and jitter injected."""
    import ssdnerf_b200 as S
    params = rp.make_decoder_params('P', 11)
    params['density_net.0.bias'] = params['density_net.0.bias'] + 1.5
    spec = up.unet_spec(image_size=128, in_channels=18, base_channels=64, channels_cfg=(1, 2), resblocks_per_downsample=1, attention_res=(), num_heads=2)
    sd = up.random_state_dict(spec, seed=4, std=0.03)
    B, res = 1, 32
    test_cfg = dict(clip_range=[-2, 2], density_thresh=0.1, n_inverse_rays=2 ** 10, loss_coef=0.1 / (res * res), dt_gamma_scale=0.5,
                    n_inverse_steps=2, extra_scene_step=1, optimizer=dict(type='Adam', lr=0.005, weight_decay=0.), lr_scheduler=dict(type='ExponentialLR', gamma=0.998))
    unet_cfg = dict(type='DenoisingUnetMod', image_size=128, in_channels=18, base_channels=64, channels_cfg=[1, 2], resblocks_per_downsample=1,
                    dropout=0.0, use_scale_shift_norm=True, num_heads=2, attention_res=[])
    model = S.build_model(dict(
        type='DiffusionNeRF', code_size=(3, 6, 128, 128), code_reshape=(18, 128, 128), code_activation=dict(type='TanhCode', scale=2), grid_size=64,
        bg_color=1, decoder_use_ema=False, diffusion_use_ema=False, freeze_decoder=True,
        pixel_loss=dict(type='MSELoss', loss_weight=20.0), reg_loss=dict(type='RegLoss', power=2, loss_weight=3e-3),
        diffusion=dict(type='GaussianDiffusion', num_timesteps=1000, betas_cfg=dict(type='linear'), denoising=unet_cfg,
                       timestep_sampler=dict(type='SNRWeightedTimeStepSampler', power=0.25),
                       ddpm_loss=dict(type='DDPMMSELossMod', rescale_mode='timestep_weight', data_info=dict(pred='v_t_pred', target='v_t'),
                                      weight_scale=4.0, scale_norm=True)),
        decoder=dict(type='TriPlaneDecoder', base_layers=[18, 64], density_layers=[64, 1], color_layers=[64, 3], use_dir_enc=True,
                     dir_layers=[16, 64], max_steps=256)), test_cfg=test_cfg)
    dsd = model.decoder.state_dict(); dsd.update(params); model.decoder.load_state_dict(dsd)
    model.diffusion.denoising.load_state_dict(sd)
    model = model.to(cuda).eval()
    model.diffusion.ddpm_loss.norm_factor.fill_(0.5)
    g = torch.Generator().manual_seed(2)
    poses = torch.from_numpy(spiral_poses(1))[None]
    f = 131.25 * res / 128
    intr = torch.tensor([f, f, res / 2, res / 2]).expand(B, 1, 4).contiguous()
    cond_imgs = torch.rand(B, 1, res, res, 3, generator=g)
    code0_ = torch.randn(B, 3, 6, 128, 128, generator=g) * 0.3
    ts = [torch.tensor([300]), torch.tensor([700])]
    noises = [torch.randn(B, 18, 128, 128, generator=g) for _ in range(2)]
    pert = [torch.rand(B, res * res, generator=g) for _ in range(8)]
    jit = [torch.rand(64 ** 3, 3, generator=g) for _ in range(8)]
    it = dict(p=iter(pert), j=iter(jit), t=iter(ts), n=iter(noises))
    orig_loss, orig_ues, orig_ft = model.loss, model.update_extra_state, model.diffusion.forward_train
    model.loss = lambda *a, **k: orig_loss(*a, **dict(k, perturb=next(it['p']).to(cuda)))
    model.update_extra_state = lambda *a, **k: orig_ues(*a, **dict(k, jitter=next(it['j']).to(cuda)))
    model.diffusion.forward_train = lambda x0, **k: orig_ft(x0, **dict(k, t=next(it['t']), noise=next(it['n']).to(cuda)))
    code_gpu, grid_gpu, bits_gpu = model.val_optim(dict(cond_imgs=cond_imgs.to(cuda), cond_intrinsics=intr.to(cuda), cond_poses=poses.to(cuda)),
                                                   code_=code0_.clone().to(cuda).requires_grad_(True))
    # ---- oracle chain (CPU, fp32): same optimiser class, gradients from the oracle UNet autograd and the oracle render loss
    dv = up.diffusion_vars(up.linear_betas())
    w_t = up.snr_weighted_loss_weight(dv, 0.25, 'V')
    ro, rd = rp.get_cam_rays(poses, intr, res, res)
    dtg = (0.5 / intr[..., :2].mean(dim=(-2, -1))).numpy()
    grid = torch.zeros(B, 64 ** 3, dtype=torch.float16)
    code_ = code0_.clone().requires_grad_(True)
    opt = torch.optim.Adam([code_], lr=0.005, weight_decay=0.)
    sch = torch.optim.lr_scheduler.ExponentialLR(opt, gamma=0.998)
    it2 = dict(p=iter(pert), j=iter(jit))
    bits = None
    for outer in range(2):
        opt.zero_grad()
        x0 = (code_.tanh() * 2).reshape(B, 18, 128, 128)
        t, eps = ts[outer], noises[outer]
        mean, std = float(np.float32(dv['sqrt_alphas_bar'][int(t)])), float(np.float32(dv['sqrt_one_minus_alphas_bar'][int(t)]))
        x_t = x0 * mean + eps * std
        v = up.unet_forward(sd, spec, x_t, t.expand(B))
        loss = ((v - (mean * eps - std * x0)).square().flatten(1).mean(1) * 0.5 * w_t[t] * 4.0).mean() / 0.5
        loss.backward()
        prior = code_.grad.clone()
        for inner in range(2):                                       # extra_scene_step + 1
            code = (code_.detach().tanh() * 2)
            if inner % 16 == 0:
                bits, _ = rp.update_extra_state(params, code, grid, next(it2['j']), density_thresh=0.1, decay=0.9)
            _, gcode, _ = tp.render_loss_grad(params, code, ro.reshape(B, -1, 3).numpy(), rd.reshape(B, -1, 3).numpy(), cond_imgs.reshape(B, -1, 3).numpy(),
                                              bits, noises=next(it2['p']).numpy(), dt_gamma=dtg, bg_color=1.0, pixel_weight=20.0,
                                              loss_coef=test_cfg['loss_coef'], scale_num_ray=res * res, reg_weight=3e-3)
            # d loss / d code_ = d loss / d code * 2 (1 - tanh^2); the render gradient ADDS to the prior gradient already in .grad
            code_.grad = prior + gcode.float() * 2 * (1 - code_.detach().tanh() ** 2)
            opt.step(); sch.step()
    ref = (code_.detach().tanh() * 2)
    moved = _rel_l2(ref, code0_.tanh() * 2)
    rel = _rel_l2(code_gpu, ref)
    print('val_optim rel l2 %.3e (the optimisation moved the code by %.3e)' % (rel, moved))
    assert rel < 3e-3 and rel < 0.1 * moved
