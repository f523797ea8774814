"""End-to-end plugin path on a small model: DiffusionNeRF.val_step (DDIM -> occupancy grid -> fused render) vs the oracle chain.
The UNet runs fp16 tensor-core GEMMs, so the comparison is statistical (image-level), not bit-level."""
import pytest
import torch

from oracle import render_port as rp
from oracle import unet_port as up
from tests.common import spiral_poses

pytestmark = pytest.mark.gpu


def test_val_step_small_model(cuda):
    import ssdnerf_b200 as S
    unet_cfg = dict(type='DenoisingUnetMod', image_size=32, in_channels=18, base_channels=128, channels_cfg=[1, 2], resblocks_per_downsample=1,
                    dropout=0.0, use_scale_shift_norm=True, num_heads=2, attention_res=[16])
    model = S.build_model(dict(
        type='DiffusionNeRF', code_size=(3, 6, 32, 32), code_reshape=(18, 32, 32), grid_size=64, bg_color=1, decoder_use_ema=False,
        diffusion_use_ema=False, freeze_decoder=False,
        diffusion=dict(type='GaussianDiffusion', num_timesteps=1000, betas_cfg=dict(type='linear'), denoising=unet_cfg),
        decoder=dict(type='TriPlaneDecoder', base_layers=[18, 64], density_layers=[64, 1], color_layers=[64, 3], use_dir_enc=True,
                     dir_layers=[16, 64], max_steps=256)),
        test_cfg=dict(img_size=(32, 32), num_timesteps=6, clip_range=[-2, 2], density_thresh=0.1, density_step=2))
    spec = up.unet_spec(image_size=32, in_channels=18, base_channels=128, channels_cfg=(1, 2), resblocks_per_downsample=1,
                        attention_res=(16,), num_heads=2)
    sd = up.random_state_dict(spec, seed=9, std=0.03)
    model.diffusion.denoising.load_state_dict(sd)
    params = rp.make_decoder_params('P', 7)
    params['density_net.0.bias'] = params['density_net.0.bias'] + 1.0
    dsd = model.decoder.state_dict()
    dsd.update(params)
    model.decoder.load_state_dict(dsd)
    model = model.to(cuda).eval()
    g = torch.Generator().manual_seed(10)
    B, V = 2, 2
    noise = torch.randn(B, 3, 6, 32, 32, generator=g)
    poses = torch.from_numpy(spiral_poses(V))[None].repeat(B, 1, 1, 1)
    intr = torch.tensor([32 * 131.25 / 128, 32 * 131.25 / 128, 16.0, 16.0]).expand(B, V, 4).contiguous()
    out = model.val_step(dict(scene_id=[0, 1], scene_name=['a', 'b'], noise=noise.to(cuda), test_poses=poses.to(cuda),
                              test_intrinsics=intr.to(cuda)))
    imgs = out['pred_imgs'].cpu()                       # [B, V, 3, 32, 32], 8-bit rounded
    assert out['num_samples'] == B and imgs.shape == (B, V, 3, 32, 32)
    assert float(imgs.min()) >= 0 and float(imgs.max()) <= 1 and torch.allclose(imgs * 255, torch.round(imgs * 255), atol=1e-4)
    # oracle chain on the CPU: DDIM (fp32) -> code -> occupancy grid without jitter is not available (device RNG), so take the
    # GPU's code + bitfield and check the RENDER against the oracle, and the DDIM against the oracle separately
    dv = up.diffusion_vars(up.linear_betas())
    code_ref = up.ddim_sample(lambda x, t: up.unet_forward(sd, spec, x, t), noise.reshape(B, 18, 32, 32), dv, num_timesteps=6)
    code_gpu, grid, bitfield = model.val_uncond(dict(scene_id=[0, 1], noise=noise.to(cuda)))
    rel = float((code_gpu.cpu().reshape(B, 18, 32, 32) - code_ref).norm() / code_ref.norm())
    assert rel < 1e-2, rel
    for b in range(B):
        rgb, depth, _ = rp.render_image(params, code_gpu[b].cpu(), bitfield[b].cpu().numpy(), poses[b], intr[b], 32, 32, max_steps=256)
        ref = torch.round(torch.from_numpy(rgb).permute(0, 3, 1, 2).clamp(0, 1) * 255) / 255
        img, _ = model.render(model.decoder, code_gpu[b:b + 1], bitfield[b:b + 1], 32, 32, intr[b:b + 1].to(cuda), poses[b:b + 1].to(cuda))
        got = torch.round(img[0].permute(0, 3, 1, 2).clamp(0, 1).cpu() * 255) / 255
        assert (got - ref).abs().mean().item() < 2e-3          # in-kernel ray generation differs from torch by <= 1 ulp per ray
