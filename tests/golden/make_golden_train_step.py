"""Pin for the stage-2 training step: the reference's OWN `DiffusionNeRF.train_step` (lib/models/autodecoders/diffusion_nerf.py:66-189)
+ `GaussianDiffusion.forward_train` + `DDPMMSELossMod` (training mode: running norm factor) + `BaseNeRF.load_scene`, executed from
/root/reference on CPU (mmcv / mmgen stubbed, the small fixture UNet of make_golden_ref.py), two consecutive iterations with an SGD
optimizer on the denoiser.  Timestep draw and noise are injected (fixed tensors) on both sides.  Written to
tests/golden/reference_train_step_v1.npz; replayed by tests/test_reference_pin_cpu.py::test_stage2_train_step_matches_reference_execution
with the oracle UNet standing in for the CUDA engine.

    python tests/golden/make_golden_train_step.py          (needs /root/reference)
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tests.golden import make_golden_ref as G  # noqa: E402
from tests.golden.make_golden_cache import load_reference_cache_code  # noqa: E402

LOSS_CFG = dict(type='DDPMMSELossMod', rescale_mode='timestep_weight', data_info=dict(pred='v_t_pred', target='v_t'), weight_scale=4.0,
                scale_norm=True, loss_name='loss_ddpm_mse')
PROBE = ['in_blocks.0.0.weight', 'time_embedding.blocks.0.weight', 'in_blocks.1.0.norm_with_embedding.embedding_layer.1.weight',
         'in_blocks.1.0.conv_1.0.weight', 'mid_blocks.1.qkv.weight', 'mid_blocks.1.proj.bias', 'out.conv.weight']


def run_reference():
    mods, den, gd, sm = G.load_reference()
    msn, base = load_reference_cache_code()
    core = sys.modules['reflib.core']
    core.rgetattr = core.module_requires_grad = None
    sys.modules['mmgen.models.builder'].MODELS = G.MODULES
    dn = G._load('lib/models/autodecoders/diffusion_nerf.py', 'reflib.models.autodecoders.diffusion_nerf')
    torch.manual_seed(0)
    unet = den.DenoisingUnetMod(**G.UNET_CFG)
    unet.load_state_dict(G.seeded_state_dict(unet, seed=11))
    dl = gd.GaussianDiffusion(denoising=unet, betas_cfg=dict(type='linear'), num_timesteps=1000, denoising_mean_mode='V',
                              timestep_sampler=dict(type='SNRWeightedTimeStepSampler', power=0.5), ddpm_loss=dict(LOSS_CFG))
    dl.train()
    g = torch.Generator().manual_seed(5)
    codes = [torch.tanh(torch.randn(3, 6, 16, 16, generator=g)) * 1.5 for _ in range(2)]
    ts = [torch.tensor([12, 870]), torch.tensor([400, 3])]
    noises = [torch.randn(2, 18, 16, 16, generator=g) for _ in range(2)]
    fake = types.SimpleNamespace(diffusion=dl, decoder=None, decoder_ema=None, freeze_decoder=True, decoder_use_ema=True, train_cfg=dict(),
                                 image_cond=False, autocast_dtype=None, code_permute=None, code_reshape=(18, 16, 16), code_size=(3, 6, 16, 16),
                                 code_activation=None, parameters=lambda: iter([torch.zeros(1)]))
    fake.load_scene = types.MethodType(base.BaseNeRF.load_scene, fake)
    fake.code_diff_pr = types.MethodType(dn.DiffusionNeRF.code_diff_pr, fake)
    opt = dict(diffusion=torch.optim.SGD(dl.parameters(), lr=0.05))
    data = dict(scene_id=[0, 1], scene_name=['a', 'b'], code=[dict(param=dict(code=c)) for c in codes])
    out = dict(codes=torch.stack(codes).numpy(), t=torch.stack(ts).numpy(), noise=torch.stack(noises).numpy(), probe_names=np.array(PROBE),
               lr=np.array(0.05))
    for it in range(2):
        dl.sampler = lambda n, it=it: ts[it]
        sys.modules['ref_gaussian_diffusion']._get_noise_batch = lambda *a, it=it, **k: noises[it]
        res = dn.DiffusionNeRF.train_step(fake, data, opt)
        lv = res['log_vars']
        out[f'it{it}_log_keys'] = np.array(sorted(lv.keys()))
        out[f'it{it}_loss'] = np.array(lv['loss_ddpm_mse'])
        out[f'it{it}_num_samples'] = np.array(res['num_samples'])
        out[f'it{it}_norm_factor'] = dl.ddpm_loss.norm_factor.numpy().copy()
        sd = unet.state_dict()
        for k in PROBE:
            out[f'it{it}_{k}'] = sd[k].numpy().copy()
        out[f'it{it}_param_checksum'] = np.array(sum(float(v.double().abs().sum()) for v in sd.values()))
    return out


if __name__ == '__main__':
    res = run_reference()
    np.savez_compressed(os.path.join(HERE, 'reference_train_step_v1.npz'), **res)
    print({k: (v.shape if v.ndim else v.item()) for k, v in res.items() if not k.startswith('it') or 'loss' in k or 'norm' in k or 'keys' in k})
