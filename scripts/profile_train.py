"""Training-side kernels for a bounded ncu capture (use `ncu --profile-from-start off`):
    python scripts/profile_train.py
between cudaProfilerStart / Stop: three weight-gradient GEMMs of the UNet at their real shapes (128x128 level 128->128 3x3, 32x32 level
512->512 3x3, 16x16 level 512->512 1x1; batch 8 = stage2_cars_uncond's samples_per_gpu) and one fused renderer backward with / without
decoder-weight gradients at the stage-1 config's ray batch (4 scenes x 4096 rays)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import IMG, orbit_poses, reference_config

import ssdnerf_b200 as S
from ssdnerf_b200 import renderer as R
from ssdnerf_b200 import unet_ops as U

dev = torch.device('cuda:0')
g = torch.Generator().manual_seed(0)
B = 8
cases = [(128, 128, 128, 9), (32, 512, 512, 9), (16, 512, 512, 1)]
ten = []
for H, cin, cout, taps in cases:
    ten.append((torch.randn(B, H, H, cout, generator=g).half().to(dev), torch.randn(B, H, H, cin, generator=g).half().to(dev),
                torch.zeros(cout, taps, cin, device=dev), cout, cin, taps))

cfg = reference_config('configs/paper_cfgs/stage1_cars_recons16v.py')
model = S.build_model(dict(cfg['model'], cache_size=0), train_cfg={k: v for k, v in cfg['train_cfg'].items() if k != 'cache_load_from'},
                      test_cfg=cfg['test_cfg']).to(dev).train()
Bs, V, n = 4, 16, 4096
poses = orbit_poses(V)[None].repeat(Bs, 1, 1, 1).contiguous().to(dev)
intr = torch.tensor([131.25, 131.25, 64.0, 64.0]).expand(Bs, V, 4).contiguous().to(dev)
code = (torch.randn(Bs, 3, 6, 128, 128, generator=g) * 0.5).to(dev)
with torch.no_grad():
    _, bits = model.get_density(model.decoder, code, cfg=dict(density_thresh=0.1))
ro, rd = R.get_cam_rays(poses, intr, IMG, IMG)
sel = torch.randint(0, V * IMG * IMG, (Bs, n), device=dev)
ro = ro.reshape(Bs, -1, 3).gather(1, sel[..., None].expand(-1, -1, 3)).contiguous()
rd = rd.reshape(Bs, -1, 3).gather(1, sel[..., None].expand(-1, -1, 3)).contiguous()
planes, blob = R.pack_planes(code, R.DEC_P), model.decoder.packed_blob()
dtg = torch.full((Bs,), 0.5 / 131.25, device=dev)
fw = R.render_train_fwd(planes, (128, 128), bits, blob, ro, rd, dt_gamma=dtg)
gi = torch.randn(Bs, n, 3, device=dev)


def run():
    for gy, x, dw, cout, cin, taps in ten:
        U.conv_wgrad(gy, x, dw, cout, cin, taps=taps)
    for want in (False, True):
        R.render_train_bwd(planes, (128, 128), bits, blob, ro, rd, fw['weights_sum'], fw['image'], None, gi, dt_gamma=dtg, want_decoder_grad=want)


run(); torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
run(); torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print('done')
