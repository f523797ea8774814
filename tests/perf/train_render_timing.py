"""Guidance-step render cost (BASELINE config 4 shape): B scenes x 128x128 rays, loss + d loss / d code.
fused = ssdnerf_render_train_fwd + mse_render_loss + render_train_bwd;  per-op = march_rays_train -> torch point_decode -> composite
(the reference's composition, on this library's kernels)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import ssdnerf_b200 as S
from ssdnerf_b200 import renderer as R

dev = torch.device('cuda:0')
B, res = int(os.environ.get('B', 8)), 128
torch.manual_seed(0)
model = S.build_model(dict(
    type='DiffusionNeRF', code_size=(3, 6, 128, 128), code_reshape=(18, 128, 128), grid_size=64, bg_color=1, decoder_use_ema=False,
    diffusion_use_ema=False, freeze_decoder=True, pixel_loss=dict(type='MSELoss', loss_weight=20.0),
    reg_loss=dict(type='RegLoss', power=2, loss_weight=3e-3),
    diffusion=dict(type='GaussianDiffusion', num_timesteps=1000, betas_cfg=dict(type='linear'),
                   denoising=dict(type='DenoisingUnetMod', image_size=128, in_channels=18, base_channels=64, channels_cfg=[1, 2],
                                  resblocks_per_downsample=1, use_scale_shift_norm=True, num_heads=2, attention_res=[])),
    decoder=dict(type='TriPlaneDecoder', base_layers=[18, 64], density_layers=[64, 1], color_layers=[64, 3], use_dir_enc=True,
                 dir_layers=[16, 64], max_steps=256)), test_cfg=dict(loss_coef=0.1 / (128 * 128), density_thresh=0.1)).to(dev)
dec = model.decoder
with torch.no_grad():
    dec.density_net[0].bias += 1.0
dec.train()
code = (torch.randn(B, 3, 6, 128, 128, device=dev) * 0.7).clamp(-2, 2)
from tests.common import spiral_poses
poses = torch.from_numpy(spiral_poses(B))[:, None].to(dev)
f = 131.25
intr = torch.tensor([f, f, 64.0, 64.0], device=dev).expand(B, 1, 4).contiguous()
rays_o, rays_d = R.get_cam_rays(poses, intr, res, res)
rays_o, rays_d = rays_o.reshape(B, -1, 3), rays_d.reshape(B, -1, 3)
target = torch.rand(B, res * res, 3, device=dev)
_, bits = model.get_density(dec, code, cfg=dict(density_thresh=0.1))
dt_gamma = torch.full((B,), 0.5 / f, device=dev)
cnt = R.render_train_fwd(R.pack_planes(code, R.DEC_P), (128, 128), bits, dec.packed_blob(), rays_o, rays_d, dt_gamma=dt_gamma, want_counts=True)['num_samples']
samples = int(cnt.sum())


def step(fused):
    dec.fused_train = fused
    c = code.clone().requires_grad_(True)
    _, loss, _ = model.loss(dec, c, bits, target, rays_o, rays_d, dt_gamma if fused else dt_gamma.tolist(), scale_num_ray=res * res,
                            cfg=model.test_cfg)
    g, = torch.autograd.grad(loss, c)
    return g


out = dict(B=B, rays=B * res * res, samples=samples)
for fused in (True, False):
    for _ in range(3):
        step(fused)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n):
        step(fused)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    out['fused_ms' if fused else 'per_op_ms'] = ms
    out['fused_Msamples_s' if fused else 'per_op_Msamples_s'] = samples / ms / 1e3
out['speedup'] = out['per_op_ms'] / out['fused_ms']
print(json.dumps(out))
