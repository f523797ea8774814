"""Generates tests/golden/oracle_v1.npz from the CPU oracle (run in the build container: python tests/golden/make_golden.py).

The reference ships no golden vectors and cannot run without a GPU, so these fixtures freeze the ORACLE
(regression pin); the oracle itself is cross-checked against the reference's own kernels on the GPU box (tests/test_ref_gpu.py)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle as orc  # noqa: E402
from oracle import render_port as rp  # noqa: E402
from oracle import unet_port as up  # noqa: E402
from tests.common import config1  # noqa: E402


def main():
    out = {}
    code, poses, intr = config1('P', res=16)
    ro, rd = rp.get_cam_rays(poses[0], intr[0], 16, 16)
    ro, rd = ro.reshape(-1, 3).numpy(), rd.reshape(-1, 3).numpy()
    out['rays_o'], out['rays_d'] = ro, rd
    aabb = np.array([-1, -1, -1, 1, 1, 1], np.float32)
    out['nears'], out['fars'] = orc.near_far_from_aabb(ro, rd, aabb, 0.2)
    bf = rp.sphere_bitfield()
    out['sphere_bitfield_crc'] = np.array([int(bf.astype(np.uint64).sum()), int(np.bitwise_xor.reduce(bf))], np.int64)
    tr, ts, cnt = orc.trace_rays(ro, rd, out['nears'], out['fars'], 1.0, bf, 1, 64, 0.0, 256, cap=64)
    out['trace'], out['trace_t'], out['trace_counts'] = tr, ts, cnt
    g = torch.Generator().manual_seed(0)
    coords = torch.randint(0, 128, (64, 3), generator=g, dtype=torch.int32).numpy()
    out['morton_coords'], out['morton_idx'] = coords, orc.morton3D(coords)
    d = torch.nn.functional.normalize(torch.randn(32, 3, generator=g), dim=-1).numpy()
    out['sh_dirs'], out['sh_out'] = d, orc.sh_encode(d, 4)
    params = rp.make_decoder_params('P', 0)
    ref = rp.render_eval_scene(params, ro, rd, code[0], bf, max_steps=256)
    out['render_image'], out['render_ws'], out['render_depth'] = ref['image'], ref['weights_sum'], ref['depth']
    dv = up.diffusion_vars(up.linear_betas())
    out['alphas_bar_50'] = dv['alphas_bar'][up.ddim_timesteps(1000, 50).numpy()]
    small = up.unet_spec(image_size=16, in_channels=18, base_channels=64, channels_cfg=(1, 2), resblocks_per_downsample=1,
                         attention_res=(8,), num_heads=2)
    sd = up.random_state_dict(small, seed=0, std=0.04)
    x = torch.randn(1, 18, 16, 16, generator=g)
    out['unet_x'] = x.numpy()
    out['unet_y'] = up.unet_forward(sd, small, x, torch.tensor([500])).numpy()
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'oracle_v1.npz'), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == '__main__':
    main()
