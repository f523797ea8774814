"""Generates tests/golden/oracle_train_v1.npz: regression pin of the train / guidance-branch oracle (oracle/train_port.py + the C march).
Run in the build container: python tests/golden/make_golden_train.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle as orc  # noqa: E402
from oracle import render_port as rp  # noqa: E402
from oracle import train_port as tp  # noqa: E402
from tests.common import config1  # noqa: E402


def case():
    code, poses, intr = config1('P', seed=5, res=16)
    ro, rd = rp.get_cam_rays(poses[0], intr[0], 16, 16)
    sel = np.linspace(0, 255, 40).astype(np.int64)
    ro, rd = ro.reshape(-1, 3).numpy()[sel], rd.reshape(-1, 3).numpy()[sel]
    params = rp.make_decoder_params('P', 5)
    params['density_net.0.bias'] = params['density_net.0.bias'] + 1.0
    bf = rp.sphere_bitfield()
    rng = np.random.default_rng(5)
    noises = rng.random(40).astype(np.float32)
    target = rng.random((1, 40, 3)).astype(np.float32)
    return code, ro, rd, params, bf, noises, target


def compute():
    code, ro, rd, params, bf, noises, target = case()
    out = dict(rays_o=ro, rays_d=rd, noises=noises, target=target)
    aabb = np.array([-1, -1, -1, 1, 1, 1], np.float32)
    nears, fars = orc.near_far_from_aabb(ro, rd, aabb, 0.2)
    xyzs, dirs, deltas, rays = orc.march_rays_train(ro, rd, 1.0, bf, 1, 64, nears, fars, dt_gamma=0.004, max_steps=256, noises=noises)
    out['march_rays'] = rays
    out['march_deltas_head'] = deltas[:64]
    ws, depth, img = tp.render_train_scene(params, code[0], ro, rd, bf, noises, dt_gamma=0.004, T_thresh=0.05, dtype=torch.float64)
    out['ws'], out['depth'], out['image'] = ws.numpy(), depth.numpy(), img.numpy()
    loss, grad, rgb = tp.render_loss_grad(params, code.double(), ro[None], rd[None], target, [bf], noises=noises[None], dt_gamma=np.array([0.004]),
                                          bg_color=1.0, pixel_weight=20.0, loss_coef=0.1 / 256, scale_num_ray=40, reg_weight=3e-3, T_thresh=0.05)
    g = grad.numpy()
    flat = np.abs(g).reshape(-1)
    top = np.argsort(-flat)[:32]
    out['loss'] = np.array([float(loss)])
    out['grad_top_idx'], out['grad_top_val'] = top.astype(np.int64), g.reshape(-1)[top]
    out['grad_abs_sum_per_plane'] = np.abs(g).sum(axis=(0, 2, 3, 4))
    out['out_rgb'] = rgb.numpy()
    return out


if __name__ == '__main__':
    out = compute()
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'oracle_train_v1.npz'), **out)
    print({k: np.asarray(v).shape for k, v in out.items()}, float(out['loss'][0]), int((out['ws'] > 0).sum()))
