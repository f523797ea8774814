"""Shared seeded inputs for the parity tests (SURVEY.md §8d config 1 and friends)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

# demo/camera_spiral_cars/pose/000000.txt of the reference (the only real camera data available offline),
# committed here because /root/reference does not exist on the GPU box.
POSE0 = np.array([
    -0.9129449725151062, -2.6523304086367716e-07, -0.40808194875717163, 0.5305066704750061,
    0.40808194875717163, -4.6272722897811036e-07, -0.9129451513290405, 1.186828851699829,
    1.2454208331291738e-07, -0.9999998807907104, 1.967826221971336e-07, 1.7763558229607138e-14,
    -0.0, 0.0, -0.0, 1.0], dtype=np.float32).reshape(4, 4)


def spiral_poses(num, radius=2.6, seed=0):
    """Deterministic look-at cameras on a sphere of `radius` (stand-in for the 251-pose demo spiral)."""
    poses = []
    for i in range(num):
        phi = 2 * np.pi * i / max(num, 1) + 0.3
        theta = np.pi / 2 - 0.45 * np.sin(1.7 * phi)
        pos = radius * np.array([np.sin(theta) * np.cos(phi), np.sin(theta) * np.sin(phi), np.cos(theta)])
        fwd = -pos / np.linalg.norm(pos)
        up = np.array([0.0, 0.0, 1.0])
        right = np.cross(fwd, up); right /= np.linalg.norm(right)
        down = np.cross(fwd, right)
        c2w = np.eye(4)
        c2w[:3, 0], c2w[:3, 1], c2w[:3, 2], c2w[:3, 3] = right, down, fwd, pos
        poses.append(c2w)
    return np.stack(poses).astype(np.float32)


def config1(variant='S', seed=0, res=64):
    """single random-init triplane, res x res render from the demo pose (translation x2), 32 samples/ray."""
    g = torch.Generator().manual_seed(seed)
    C = 32 if variant == 'S' else 6
    code = torch.randn(1, 3, C, 128, 128, generator=g).clamp(-2, 2)
    pose = POSE0.copy()
    pose[:3, 3] *= 2
    poses = torch.from_numpy(pose)[None, None]                      # [1,1,4,4]
    f = 131.25 * res / 128
    intr = torch.tensor([[[f, f, res / 2, res / 2]]], dtype=torch.float32)   # [1,1,4]
    return code, poses, intr


def rel_err(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-12))


def seeded_weights(keys, shapes, seed):
    """State-dict tensors drawn from one seeded CPU generator in key order (GroupNorm gains ~1, other vectors ~0.1 N(0,1),
    matrices / kernels N(0, 1.5/sqrt(fan_in))).  tests/golden/make_golden_ref.py loads exactly these into the reference's
    own modules, so the committed fixtures need only the key / shape lists and the seed."""
    import math
    g = torch.Generator().manual_seed(int(seed))
    sd = {}
    for k, shape in zip(keys, shapes):
        shape = tuple(int(s) for s in shape)
        if len(shape) == 1 and (k.endswith('norm.weight') or k.endswith('.0.weight') or k.endswith('gn.weight')):
            sd[k] = 1 + 0.1 * torch.randn(shape, generator=g)
        elif len(shape) == 1:
            sd[k] = 0.1 * torch.randn(shape, generator=g)
        else:
            sd[k] = torch.randn(shape, generator=g) * (1.5 / math.sqrt(int(np.prod(shape[1:]))))
    return sd


def parse_shapes(arr):
    return [tuple(int(x) for x in s.split(',')) if s else () for s in arr.tolist()]


class ToyDecoder(torch.nn.Module):
    """Pure-torch stand-in for the volume renderer behind the decoder interface (`forward(rays_o, rays_d, code, density_bitfield, grid_size,
    dt_gamma=, perturb=, return_loss=) -> dict(weights_sum, depth, image[, decoder_reg_loss])`).  Test infrastructure for HOST-LOGIC pins only:
    tests/golden/make_golden_joint_step.py runs the reference's own `train_step`s with it on CPU and tests/test_reference_pin_cpu.py replays
    the same sequence through ssdnerf_b200's.  It is differentiable w.r.t. the code and its own weights; rays select code texels."""

    def __init__(self, channels=18, seed=3):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.lin = torch.nn.Linear(channels, 3)
        self.head = torch.nn.Linear(channels, 1)
        with torch.no_grad():
            for p in self.parameters():
                p.copy_(torch.randn(p.shape, generator=g) * 0.3)
        self.bound, self.min_near, self.max_steps = 1, 0.2, 8

    def forward(self, rays_o, rays_d, code, density_bitfield, grid_size, dt_gamma=0, perturb=False, T_thresh=1e-4, return_loss=False):
        B = code.size(0)
        feat = code.reshape(B, code.size(1) * code.size(2), -1)                         # [B, 18, h*w]
        idx = ((rays_d[..., 0] * 7.3 + rays_d[..., 2] * 1.9 + rays_o[..., 1] * 3.1).abs() * 1000).long() % feat.size(-1)   # [B, N]
        f = feat.gather(2, idx[:, None, :].expand(-1, feat.size(1), -1)).transpose(1, 2)                                     # [B, N, 18]
        ws = torch.sigmoid(self.head(f)).squeeze(-1)
        out = dict(weights_sum=ws, depth=ws.detach() * 2.0, image=torch.sigmoid(self.lin(f)) * ws.unsqueeze(-1))
        if return_loss:
            out.update(decoder_reg_loss=None)
        return out
