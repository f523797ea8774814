"""Plugin surface on CPU: registry names, config loading, state-dict keys, host-side DDIM tables (no GPU compute)."""
import os

import numpy as np
import pytest
import torch

import ssdnerf_b200 as S
from oracle import unet_port as up

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference'


def test_registry_names():
    for n in ('TriPlaneDecoder', 'GaussianDiffusion', 'DenoisingUnetMod', 'TanhCode', 'IdentityCode'):
        assert n in S.MODULES
    assert 'DiffusionNeRF' in S.MODELS


def test_build_from_repo_config_and_state_dict_keys():
    cfg = S.Config.fromfile(os.path.join(ROOT, 'configs', 'cars_uncond_b200.py'))
    m = S.build_model(cfg.model, test_cfg=cfg.test_cfg)
    sd = m.diffusion.denoising.state_dict()
    ref = up.random_state_dict(up.unet_spec())
    assert set(sd) == set(ref) and all(sd[k].shape == ref[k].shape for k in ref)
    assert sum(v.numel() for v in sd.values()) == 122434194
    dec = m.decoder.state_dict()
    assert tuple(dec['base_net.0.weight'].shape) == (64, 18) and tuple(dec['dir_net.0.weight'].shape) == (64, 16)
    assert tuple(dec['color_net.0.weight'].shape) == (3, 64) and tuple(dec['aabb'].shape) == (6,)
    # EMA twins selected at inference exist (diffusion_nerf.py:192-193)
    assert hasattr(m, 'decoder_ema') and hasattr(m, 'diffusion_ema')
    assert m.code_diff_pr(torch.zeros(2, 3, 6, 128, 128)).shape == (2, 18, 128, 128)
    assert m.code_diff_pr_inv(torch.zeros(2, 18, 128, 128)).shape == (2, 3, 6, 128, 128)


@pytest.mark.skipif(not os.path.isdir(REF), reason='reference checkout only exists in the build container')
@pytest.mark.parametrize('cfg_rel', ['configs/paper_cfgs/ssdnerf_cars_uncond.py', 'configs/paper_cfgs/ssdnerf_chairs_recons1v.py',
                                     'configs/paper_cfgs/ssdnerf_abotables_uncond.py'])
def test_reference_configs_build_unchanged(cfg_rel):
    cfg = S.Config.fromfile(os.path.join(REF, cfg_rel))
    m = S.build_model(cfg.model, train_cfg=cfg.get('train_cfg'), test_cfg=cfg.get('test_cfg'))
    assert type(m).__name__ == 'DiffusionNeRF'
    assert m.diffusion.test_cfg['num_timesteps'] == cfg.test_cfg['num_timesteps']


def test_ddim_tables_match_oracle():
    from ssdnerf_b200.diffusion import GaussianDiffusion
    from ssdnerf_b200.unet import DenoisingUnetMod
    unet = DenoisingUnetMod(image_size=32, in_channels=18, base_channels=64, channels_cfg=[1, 2], resblocks_per_downsample=1,
                            num_heads=2, attention_res=[16], use_scale_shift_norm=True)
    d = GaussianDiffusion(unet, betas_cfg=dict(type='linear'), test_cfg=dict(num_timesteps=50))
    dv = up.diffusion_vars(up.linear_betas())
    np.testing.assert_array_equal(d.alphas_bar, dv['alphas_bar'])
    np.testing.assert_array_equal(d.tilde_betas_t, dv['tilde_betas_t'])
    ts = d.ddim_timesteps(50)
    assert ts[0] == 999 and ts[-1] == 19 and len(ts) == 50
    coef = d.ddim_coefficients(ts)
    assert coef.shape == (50, 4) and coef[-1, 2] == 1.0          # alpha_bar_prev = 1 at the last step (t_prev = -1)
    # time embedding of the module equals the oracle restatement
    sd = {k: v.clone() for k, v in unet.state_dict().items()}
    e1 = unet.embedding(ts)
    e2 = up.time_embedding(sd, ts.float(), 64)
    torch.testing.assert_close(e1, e2)


def test_unsupported_paths_fail_loudly():
    cfg = S.Config.fromfile(os.path.join(ROOT, 'configs', 'cars_uncond_b200.py'))
    m = S.build_model(cfg.model, test_cfg=cfg.test_cfg)
    with pytest.raises(NotImplementedError):
        m.train_step({}, None)
    with pytest.raises(S._lib.SSDNeRFNativeError):
        m.render(m.decoder, torch.zeros(1, 3, 6, 128, 128), torch.zeros(1, 32768, dtype=torch.uint8), 8, 8,
                 torch.zeros(1, 1, 4), torch.zeros(1, 1, 4, 4))
