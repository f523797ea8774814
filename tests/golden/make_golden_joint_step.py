"""Pins for the HOST LOGIC of the joint training steps: the reference's OWN `MultiSceneNeRF.train_step` (stage 1) and
`DiffusionNeRF.train_step` (single stage: denoiser + decoder + latents), built through their real constructors and executed from
/root/reference on CPU -- with `load_cache` / `save_cache` / `inverse_code` / `loss_decoder` / `loss` / `ray_sample` / `get_raybatch_inds` /
`mean_ema_update` / `TanhCode` / `TVLoss` / `GaussianDiffusion.forward_train` / `DDPMMSELossMod`, mmcv / mmgen stubbed.  The volume renderer
is replaced ON BOTH SIDES by tests/common.py:ToyDecoder (pure torch) and the occupancy-grid update by a no-op, because the reference's
need its CUDA extension; everything else is the reference's code.  Results -> tests/golden/reference_joint_step_v1.npz, replayed by
tests/test_reference_pin_cpu.py::test_joint_train_steps_match_reference_execution.

    python tests/golden/make_golden_joint_step.py          (needs /root/reference)
"""
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tests.common import ToyDecoder  # noqa: E402
from tests.golden import make_golden_ref as G  # noqa: E402
from tests.golden.make_golden_cache import load_reference_cache_code  # noqa: E402
from tests.golden.make_golden_train_step import LOSS_CFG, PROBE  # noqa: E402

CODE1, GRID = (3, 6, 8, 8), 8
TRAIN1 = dict(optimizer=dict(type='Adam', lr=0.01, weight_decay=0.0), n_decoder_rays=40, n_inverse_rays=48, extra_scene_step=3,
              loss_coef=0.01, dt_gamma_scale=0.5, density_thresh=0.1)
MODEL1 = dict(code_size=CODE1, grid_size=GRID, code_activation=dict(type='TanhCode', scale=2), bg_color=1, init_from_mean=True,
              pixel_loss=dict(type='MSELoss', loss_weight=20.0), reg_loss=dict(type='TVLoss', power=1.5, loss_weight=1.0),
              decoder=dict(type='ToyDecoder'), decoder_use_ema=False, cache_size=3,
              cache_16bit=True)      # 16-bit cache: on CPU an fp32 cache entry would be aliased by the reference's .to() and break its own save_cache
CODE2 = (3, 6, 16, 16)
TRAIN2 = dict(TRAIN1, optimizer=dict(type='Adam', lr=0.005, weight_decay=0.0), extra_scene_step=2)


class MSELoss(nn.Module):
    """mmgen.models.losses.MSELoss [mmgen-memory]: loss_weight * mean squared error"""

    def __init__(self, loss_weight=1.0, reduction='mean', **kw):
        super().__init__()
        self.loss_weight, self.reduction = loss_weight, reduction

    def forward(self, pred, target, weight=None, avg_factor=None, **kw):
        assert weight is None and avg_factor is None and self.reduction == 'mean'
        return self.loss_weight * torch.nn.functional.mse_loss(pred, target)


def views(B, V, res, seed):
    """conditioning views: look-at poses on a circle, pinhole intrinsics, random target images"""
    from tests.common import spiral_poses
    g = torch.Generator().manual_seed(seed)
    poses = torch.from_numpy(spiral_poses(V))[None].repeat(B, 1, 1, 1).float()
    f = 131.25 * res / 128
    intr = torch.tensor([f, f, res / 2, res / 2]).expand(B, V, 4).contiguous()
    return torch.rand(B, V, res, res, 3, generator=g), poses, intr


def load_all():
    mods, den, gd, sm = G.load_reference()
    msn, base = load_reference_cache_code()
    misc = sys.modules['ref_misc']

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m
    mod('skimage'); mod('mmgen.core'); mod('mmgen.core.registry', METRICS=G.MODULES)
    mod('mmgen.core.evaluation'); mod('mmgen.core.evaluation.metrics', FID=type('FID', (), {}))
    metrics = G._load('lib/core/evaluation/metrics.py', 'ref_metrics')
    nu = sys.modules['ref_nerf_utils']
    core = sys.modules['reflib.core']
    real = dict(eval_psnr=metrics.eval_psnr, rgetattr=misc.rgetattr, rsetattr=misc.rsetattr, module_requires_grad=misc.module_requires_grad,
                get_cam_rays=nu.get_cam_rays)
    for k, v in real.items():
        setattr(core, k, v)
    for m in (base, msn):            # names these modules bound at import time while the stubs were still None
        for k, v in real.items():
            if hasattr(m, k):
                setattr(m, k, v)
    sys.modules['mmgen.models.builder'].MODELS = G.MODULES
    dn = G._load('lib/models/autodecoders/diffusion_nerf.py', 'reflib.models.autodecoders.diffusion_nerf')
    G.MODULES.register_module(name='MSELoss', module=MSELoss)
    G.MODULES.register_module(name='ToyDecoder', module=ToyDecoder)
    return msn, dn, den


def record(out, tag, model, res, decoder, extra=None):
    lv = res['log_vars']
    out[f'{tag}_log_keys'] = np.array(sorted(lv.keys()))
    out[f'{tag}_log_vals'] = np.array([lv[k] for k in sorted(lv.keys())], np.float64)
    out[f'{tag}_num_samples'] = np.array(res['num_samples'])
    for k, v in decoder.state_dict().items():
        out[f'{tag}_dec_{k}'] = v.numpy().copy()
    out[f'{tag}_init_code'] = model.init_code.numpy().copy()
    for sid, e in model.cache.items():
        out[f'{tag}_cache{sid}_filled'] = np.array(e is not None)
        if e is not None:
            out[f'{tag}_cache{sid}_code'] = e['param']['code_'].float().numpy().copy()
            out[f'{tag}_cache{sid}_step'] = np.array(float(e['optimizer']['state'][0]['step']))
            out[f'{tag}_cache{sid}_exp_avg'] = e['optimizer']['state'][0]['exp_avg'].float().numpy().copy()
    for k, v in (extra or {}).items():
        out[f'{tag}_{k}'] = v


def run_reference():
    msn, dn, den = load_all()
    out = {}
    # ---------------- stage 1: MultiSceneNeRF.train_step
    torch.manual_seed(0)
    m1 = msn.MultiSceneNeRF(**MODEL1, train_cfg=dict(TRAIN1), test_cfg=dict())
    m1.update_extra_state = lambda *a, **k: None
    m1.train()
    imgs, poses, intr = views(3, 2, 8, 40)
    out['s1_imgs'], out['s1_poses'], out['s1_intr'] = imgs.numpy(), poses.numpy(), intr.numpy()
    opt = dict(decoder=torch.optim.Adam(m1.decoder.parameters(), lr=1e-3))
    torch.manual_seed(123)
    for it, ids in enumerate(([2, 0], [0, 1])):
        data = dict(scene_id=ids, scene_name=[f's{i}' for i in ids], cond_imgs=imgs[ids], cond_poses=poses[ids], cond_intrinsics=intr[ids])
        res = m1.train_step(data, opt)
        record(out, f's1_it{it}', m1, res, m1.decoder)
    # ---------------- single stage: DiffusionNeRF.train_step
    torch.manual_seed(0)
    m2 = dn.DiffusionNeRF(**dict(MODEL1, code_size=CODE2), code_reshape=(18, 16, 16), freeze_decoder=False, diffusion_use_ema=False,
                          diffusion=dict(type='GaussianDiffusion', denoising=dict(type='DenoisingUnetMod', **G.UNET_CFG),
                                         betas_cfg=dict(type='linear'), num_timesteps=1000, denoising_mean_mode='V',
                                         timestep_sampler=dict(type='SNRWeightedTimeStepSampler', power=0.5), ddpm_loss=dict(LOSS_CFG)),
                          train_cfg=dict(TRAIN2), test_cfg=dict())
    unet = m2.diffusion.denoising
    unet.load_state_dict(G.seeded_state_dict(unet, seed=11))
    m2.update_extra_state = lambda *a, **k: None
    m2.train()
    g = torch.Generator().manual_seed(5)
    ts = [torch.tensor([12, 870]), torch.tensor([400, 3])]
    noises = [torch.randn(2, 18, 16, 16, generator=g) for _ in range(2)]
    out['s2_t'], out['s2_noise'] = torch.stack(ts).numpy(), torch.stack(noises).numpy()
    opt = dict(diffusion=torch.optim.SGD(m2.diffusion.parameters(), lr=0.05), decoder=torch.optim.Adam(m2.decoder.parameters(), lr=1e-3))
    torch.manual_seed(321)
    for it, ids in enumerate(([2, 0], [0, 1])):
        m2.diffusion.sampler = lambda n, it=it: ts[it]
        sys.modules['ref_gaussian_diffusion']._get_noise_batch = lambda *a, it=it, **k: noises[it]
        data = dict(scene_id=ids, scene_name=[f's{i}' for i in ids], cond_imgs=imgs[ids], cond_poses=poses[ids], cond_intrinsics=intr[ids])
        res = m2.train_step(data, opt)
        sd = unet.state_dict()
        extra = {f'unet_{k}': sd[k].numpy().copy() for k in PROBE}
        extra['unet_checksum'] = np.array(sum(float(v.double().abs().sum()) for v in sd.values()))
        extra['norm_factor'] = m2.diffusion.ddpm_loss.norm_factor.numpy().copy()
        record(out, f's2_it{it}', m2, res, m2.decoder, extra)
    return out


if __name__ == '__main__':
    res = run_reference()
    np.savez_compressed(os.path.join(HERE, 'reference_joint_step_v1.npz'), **res)
    for k in sorted(res):
        if k.endswith('log_keys') or k.endswith('log_vals'):
            print(k, res[k])
    print('wrote', len(res), 'arrays,', os.path.getsize(os.path.join(HERE, 'reference_joint_step_v1.npz')), 'bytes')
