"""Pins for the HOST LOGIC of the evaluation-side code paths of `DiffusionNeRF` (BASELINE config 4, `cond_mode='guide_optim'`): the
reference's OWN `val_guide` (render-loss guided DDIM + langevin through the denoiser), `val_optim` (diffusion-prior gradient + inner
render-loss Adam steps with a learning-rate schedule) and `val_uncond` (save_intermediates list branch + `n_inverse_steps` prior-only
refinement), executed from /root/reference on CPU through the real `DiffusionNeRF` constructor.  As in make_golden_joint_step.py the volume
renderer is tests/common.py:ToyDecoder and the occupancy-grid calls (`update_extra_state`, `get_density`) are no-ops ON BOTH SIDES; timestep
draws and noises are injected.  -> tests/golden/reference_val_v1.npz, replayed by
tests/test_reference_pin_cpu.py::test_val_paths_match_reference_execution.

    python tests/golden/make_golden_val.py          (needs /root/reference)
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tests.golden import make_golden_ref as G  # noqa: E402
from tests.golden.make_golden_joint_step import MODEL1, load_all, views  # noqa: E402
from tests.golden.make_golden_train_step import LOSS_CFG  # noqa: E402

TEST_CFG = dict(num_timesteps=4, clip_range=[-2, 2], guidance_gain=50.0, snr_weight_power=0.25, langevin_steps=1, langevin_delta=0.4,
                n_inverse_rays=48, n_inverse_steps=3, extra_scene_step=2, optimizer=dict(type='Adam', lr=0.005, weight_decay=0.0),
                lr_scheduler=dict(type='ExponentialLR', gamma=0.9), loss_coef=0.01, dt_gamma_scale=0.5, density_thresh=0.1)


def run_reference():
    msn, dn, den = load_all()
    out = {}
    torch.manual_seed(0)
    m = dn.DiffusionNeRF(**dict(MODEL1, code_size=(3, 6, 16, 16), cache_size=0, cache_16bit=False, init_from_mean=False),
                         code_reshape=(18, 16, 16), freeze_decoder=True, diffusion_use_ema=False,
                         diffusion=dict(type='GaussianDiffusion', denoising=dict(type='DenoisingUnetMod', **G.UNET_CFG),
                                        betas_cfg=dict(type='linear'), num_timesteps=1000, denoising_mean_mode='V',
                                        timestep_sampler=dict(type='SNRWeightedTimeStepSampler', power=0.25), ddpm_loss=dict(LOSS_CFG)),
                         train_cfg=dict(), test_cfg=dict(TEST_CFG))
    unet = m.diffusion.denoising
    unet.load_state_dict(G.seeded_state_dict(unet, seed=11))
    m.update_extra_state = lambda *a, **k: None
    m.get_density = lambda decoder, code, cfg=dict(): (torch.zeros(code.size(0), 8 ** 3), torch.zeros(code.size(0), 8 ** 3 // 8, dtype=torch.uint8))
    m.eval()
    m.diffusion.ddpm_loss.norm_factor.fill_(0.6)
    imgs, poses, intr = views(2, 2, 8, 41)
    g = torch.Generator().manual_seed(9)
    noise = torch.randn(2, 3, 6, 16, 16, generator=g)
    lang = [torch.randn(2, 18, 16, 16, generator=g) for _ in range(2 * 3)]
    ts = [torch.tensor([30, 600]), torch.tensor([999, 250]), torch.tensor([5, 480]), torch.tensor([700, 90]), torch.tensor([350, 351]),
          torch.tensor([64, 820])]
    eps = [torch.randn(2, 18, 16, 16, generator=g) for _ in range(6)]
    code0 = torch.randn(2, 3, 6, 16, 16, generator=g) * 0.3
    out.update(imgs=imgs.numpy(), poses=poses.numpy(), intr=intr.numpy(), noise=noise.numpy(), lang=torch.stack(lang).numpy(),
               ts=torch.stack(ts).numpy(), eps=torch.stack(eps).numpy(), code0=code0.numpy())
    data = dict(scene_id=[0, 1], scene_name=['a', 'b'], cond_imgs=imgs, cond_poses=poses, cond_intrinsics=intr, noise=noise)
    gdm = sys.modules['ref_gaussian_diffusion']
    # ---- val_guide: 4 DDIM steps, guidance through the UNet (reference default), 1 langevin step per timestep
    it = iter(lang)
    gdm._get_noise_batch = lambda *a, **k: next(it)
    torch.manual_seed(77)
    with torch.no_grad():          # the reference's test loop (lib/apis/test.py) calls val_step under no_grad; pred_x_0 re-enables grad itself
        code, grid, bits = m.val_guide(data)
    out['guide_code'] = code.detach().numpy().copy()
    # ---- val_optim: 3 outer steps x (prior gradient + 3 inner render steps), explicit start latent
    it_t, it_e = iter(ts), iter(eps)
    m.diffusion.sampler = lambda n: next(it_t)
    gdm._get_noise_batch = lambda *a, **k: next(it_e)
    torch.manual_seed(78)
    with torch.no_grad():
        code, grid, bits = m.val_optim(data, code_=code0.clone().requires_grad_(True))
    out['optim_code'] = code.detach().numpy().copy()
    # ---- val_uncond: intermediates list + prior-only refinement of the last entry (n_inverse_steps = 3)
    it = iter(lang)
    it_t, it_e = iter(ts[3:]), iter(eps[3:])
    calls = dict(n=0)

    def noise_batch(*a, **k):       # langevin draws during sampling first, then the forward_train noises of the refinement
        calls['n'] += 1
        return next(it) if calls['n'] <= 4 else next(it_e)
    gdm._get_noise_batch = noise_batch
    torch.manual_seed(79)
    with torch.no_grad():
        codes, grids, bitss = m.val_uncond(dict(scene_id=[0, 1], noise=noise), save_intermediates=True)
    out['uncond_len'] = np.array(len(codes))
    out['uncond_first'], out['uncond_last'] = codes[0].detach().numpy().copy(), codes[-1].detach().numpy().copy()
    out['uncond_noise_calls'] = np.array(calls['n'])
    return out


if __name__ == '__main__':
    res = run_reference()
    np.savez_compressed(os.path.join(HERE, 'reference_val_v1.npz'), **res)
    print({k: (v.shape if v.ndim else v.item()) for k, v in res.items()})
    print('rms', {k: float(np.sqrt((res[k] ** 2).mean())) for k in ('guide_code', 'optim_code', 'uncond_last')})
