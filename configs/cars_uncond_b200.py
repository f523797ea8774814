# Model / test settings of the reference's `ssdnerf_cars_uncond` (configs/paper_cfgs/ssdnerf_cars_uncond.py:3-88),
# restated because /root/reference is not present on the GPU box. Training-only keys are omitted.
name = 'cars_uncond_b200'

_unet = dict(type='DenoisingUnetMod', image_size=128, in_channels=18, base_channels=128, channels_cfg=[1, 2, 2, 4, 4],
             resblocks_per_downsample=2, dropout=0.0, use_scale_shift_norm=True, downsample_conv=True, upsample_conv=True,
             num_heads=4, attention_res=[32, 16, 8])

_decoder = dict(type='TriPlaneDecoder', interp_mode='bilinear', base_layers=[6 * 3, 64], density_layers=[64, 1],
                color_layers=[64, 3], use_dir_enc=True, dir_layers=[16, 64], activation='silu', sigma_activation='trunc_exp',
                sigmoid_saturation=0.001, max_steps=256)

model = dict(
    type='DiffusionNeRF', code_size=(3, 6, 128, 128), code_reshape=(18, 128, 128), code_activation=dict(type='TanhCode', scale=2),
    grid_size=64, diffusion=dict(type='GaussianDiffusion', num_timesteps=1000, betas_cfg=dict(type='linear'), denoising=_unet),
    decoder=_decoder, decoder_use_ema=True, freeze_decoder=False, bg_color=1)

test_cfg = dict(img_size=(128, 128), num_timesteps=50, clip_range=[-2, 2], density_thresh=0.1)
