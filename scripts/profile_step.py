"""One bench step (DDIM + density + render) for profiling under ncu, without bench.py's repetitions:
    python scripts/profile_step.py [ddim_steps] [views] [S]
warm-up step, then one more step between cudaProfilerStart / Stop (use `ncu --profile-from-start off`)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import IMG, build_model, orbit_poses

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
V = int(sys.argv[2]) if len(sys.argv) > 2 else 251
variant_s = len(sys.argv) > 3 and sys.argv[3] == 'S'
dev = torch.device('cuda:0')
model, _ = build_model(dev, test_cfg_update=dict(num_timesteps=steps))
model.diffusion_ema.test_cfg['num_timesteps'] = steps
B = 16
g = torch.Generator().manual_seed(1234)
noise = torch.randn(B, *model.code_size, generator=g).to(dev)
poses = orbit_poses(V)[None].repeat(B, 1, 1, 1).contiguous().to(dev)
intr = torch.tensor([131.25, 131.25, 64.0, 64.0]).expand(B, V, 4).contiguous().to(dev)


def step():
    code = model.code_diff_pr_inv(model.diffusion_ema(model.code_diff_pr(noise), return_loss=False)).contiguous()
    _, bitfield = model.get_density(model.decoder_ema, code, cfg=model.test_cfg)
    return model.render(model.decoder_ema, code, bitfield, IMG, IMG, intr, poses, cfg=model.test_cfg)


def step_s():
    import ssdnerf_b200 as S
    from ssdnerf_b200 import density as Dm, renderer as R
    dec_s = S.build_module(dict(type='TriPlaneDecoder', max_steps=256)).to(dev).eval()
    code_s = torch.randn(B, 3, 32, 128, 128, generator=torch.Generator().manual_seed(77)).clamp(-2, 2).to(dev)
    planes_s = R.pack_planes(code_s, R.DEC_S)
    _, bits_s = Dm.get_density(R.DEC_S, planes_s, (128, 128), dec_s.packed_blob(), B, density_thresh=0.1, grid_size=64, bound=1.0)
    return lambda: R.render_fwd(R.DEC_S, planes_s, (128, 128), bits_s, dec_s.packed_blob(), poses=poses, intrinsics=intr, img_hw=(IMG, IMG), want_counts=False)


fn = step_s() if variant_s else step
fn(); torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
fn(); torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print('done')
