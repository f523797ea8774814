"""`DiffusionNeRF` -- the registered model of the reference's configs, inference side.

Plugin surface of lib/models/autodecoders/{base_nerf,multiscene_nerf,diffusion_nerf}.py for the two accelerated hot
paths: `val_uncond` (noise -> DDIM -> code -> occupancy grid), `get_density`, `render`, `val_step`, `code_diff_pr(_inv)`;
the same constructor kwargs and `outputs_dict = {log_vars, num_samples, pred_imgs}` contract (diffusion_nerf.py:464-469).
Training (`train_step`), guided reconstruction (`val_guide` / `val_optim`) and metric plumbing are outside this
round's scope and raise NotImplementedError with a pointer to SURVEY.md §8(f).
"""
from copy import deepcopy

import torch
import torch.nn as nn

from . import _lib as N
from . import density as D
from . import renderer as R
from .registry import MODELS, MODULES, build_module


@MODULES.register_module()
class TanhCode(nn.Module):
    """base_nerf.py:25-37"""

    def __init__(self, scale=1.0, eps=1e-5):
        super().__init__()
        self.scale, self.eps = scale, eps

    def forward(self, code_):
        return code_.tanh() if self.scale == 1 else code_.tanh() * self.scale

    def inverse(self, code):
        return code.div(self.scale).clamp(min=-1 + self.eps, max=1 - self.eps).atanh()


@MODULES.register_module()
class IdentityCode(nn.Module):
    """base_nerf.py:40-48"""

    @staticmethod
    def forward(code_):
        return code_

    @staticmethod
    def inverse(code):
        return code


@MODELS.register_module()
class DiffusionNeRF(nn.Module):

    def __init__(self, code_size=(3, 8, 64, 64), code_activation=dict(type='TanhCode', scale=2), grid_size=64,
                 decoder=dict(type='TriPlaneDecoder'), decoder_use_ema=False, bg_color=1, pixel_loss=None, reg_loss=None,
                 update_extra_interval=16, use_lpips_metric=True, init_from_mean=False, init_scale=1e-4, mean_ema_momentum=0.001,
                 mean_scale=1.0, train_cfg=dict(), test_cfg=dict(), pretrained=None, cache_size=0, cache_16bit=False,
                 diffusion=dict(type='GaussianDiffusion'), diffusion_use_ema=True, freeze_decoder=True, image_cond=False,
                 code_permute=None, code_reshape=None, autocast_dtype=None):
        super().__init__()
        self.code_size = tuple(code_size)
        self.code_activation = build_module(code_activation)
        self.grid_size = grid_size
        self.decoder = build_module(decoder)
        self.decoder_use_ema = decoder_use_ema
        if self.decoder_use_ema:
            self.decoder_ema = deepcopy(self.decoder)
        self.bg_color = bg_color
        self.train_cfg = deepcopy(train_cfg) if train_cfg is not None else dict()
        self.test_cfg = deepcopy(test_cfg) if test_cfg is not None else dict()
        diffusion = dict(diffusion)
        diffusion.update(train_cfg=self.train_cfg, test_cfg=self.test_cfg)
        self.diffusion = build_module(diffusion)
        self.diffusion_use_ema = diffusion_use_ema
        if self.diffusion_use_ema:
            self.diffusion_ema = deepcopy(self.diffusion)
        self.freeze_decoder = freeze_decoder
        if self.freeze_decoder:
            self.decoder.requires_grad_(False)
            if self.decoder_use_ema:
                self.decoder_ema.requires_grad_(False)
        self.image_cond = image_cond
        self.code_permute = code_permute
        self.code_reshape = code_reshape
        self.code_reshape_inv = [self.code_size[axis] for axis in self.code_permute] if code_permute is not None else self.code_size
        self.code_permute_inv = [self.code_permute.index(axis) for axis in range(len(self.code_permute))] \
            if code_permute is not None else None
        self.autocast_dtype = autocast_dtype
        # training-only collaborators are kept as configuration
        self._unbuilt = dict(pixel_loss=pixel_loss, reg_loss=reg_loss, cache_size=cache_size, cache_16bit=cache_16bit)

    # ------------------------------------------------------------------ code <-> diffusion layout (diffusion_nerf.py:50-64)
    def code_diff_pr(self, code):
        code_diff = code
        if self.code_permute is not None:
            code_diff = code_diff.permute([0] + [axis + 1 for axis in self.code_permute])
        if self.code_reshape is not None:
            code_diff = code_diff.reshape(code.size(0), *self.code_reshape)
        return code_diff

    def code_diff_pr_inv(self, code_diff):
        code = code_diff
        if self.code_reshape is not None:
            code = code.reshape(code.size(0), *self.code_reshape_inv)
        if self.code_permute_inv is not None:
            code = code.permute([0] + [axis + 1 for axis in self.code_permute_inv])
        return code

    # ------------------------------------------------------------------ occupancy grid (base_nerf.py:391-401)
    def get_density(self, decoder, code, cfg=dict(), jitters=None):
        variant = decoder.fused_variant()
        planes = R.pack_planes(code, variant)
        return D.get_density(variant, planes, tuple(code.shape[-2:]), decoder.packed_blob(), code.size(0),
                             density_thresh=cfg.get('density_thresh', 0.01), density_step=cfg.get('density_step', 8),
                             grid_size=self.grid_size, bound=float(decoder.bound), jitters=jitters)

    # ------------------------------------------------------------------ render (base_nerf.py:494-533)
    def render(self, decoder, code, density_bitfield, h, w, intrinsics, poses, cfg=dict()):
        """-> image [B,V,h,w,3] (background-blended), depth [B,V,h,w]; rays are generated inside the kernel."""
        N.require_cuda(code, density_bitfield, intrinsics, poses)
        num_scenes, num_imgs = poses.shape[0], poses.shape[1]
        dt_gamma_scale = cfg.get('dt_gamma_scale', 0.0)
        dt_gamma = None
        if dt_gamma_scale != 0:
            dt_gamma = dt_gamma_scale * 2 / (intrinsics[..., 0] + intrinsics[..., 1]).mean(dim=-1)
        variant = decoder.fused_variant()
        planes = R.pack_planes(code, variant)
        poses44 = poses
        if poses.shape[-2] == 3:
            poses44 = torch.cat([poses, poses.new_tensor([0, 0, 0, 1]).expand(*poses.shape[:-2], 1, 4)], dim=-2)
        out = R.render_fwd(variant, planes, tuple(code.shape[-2:]), density_bitfield.reshape(num_scenes, -1), decoder.packed_blob(),
                           poses=poses44, intrinsics=intrinsics, img_hw=(h, w), grid_size=self.grid_size, bound=float(decoder.bound),
                           min_near=float(decoder.min_near), max_steps=int(decoder.max_steps), bg_color=float(self.bg_color),
                           dt_gamma=dt_gamma, want_counts=False)
        return out['rgb'].reshape(num_scenes, num_imgs, h, w, 3), out['depth'].reshape(num_scenes, num_imgs, h, w)

    # ------------------------------------------------------------------ unconditional generation (diffusion_nerf.py:191-239)
    @torch.no_grad()
    def val_uncond(self, data, show_pbar=False, **kwargs):
        diffusion = self.diffusion_ema if self.diffusion_use_ema else self.diffusion
        decoder = self.decoder_ema if self.decoder_use_ema else self.decoder
        num_batches = len(data['scene_id'])
        noise = data.get('noise', None)
        if noise is None:
            noise = torch.randn((num_batches, *self.code_size), device=next(self.parameters()).device)
        if self.test_cfg.get('n_inverse_steps', 0) > 0:
            raise NotImplementedError('n_inverse_steps > 0 needs the diffusion loss backward (SURVEY.md §8 f1)')
        code_out = diffusion(self.code_diff_pr(noise), return_loss=False, show_pbar=show_pbar, **kwargs)
        code = self.code_diff_pr_inv(code_out).contiguous()
        density_grid, density_bitfield = self.get_density(decoder, code, cfg=self.test_cfg)
        return code, density_grid, density_bitfield

    def val_guide(self, data, **kwargs):
        raise NotImplementedError('guided reconstruction needs the fused train-mode renderer backward (SURVEY.md §8 a10 / f1)')

    def val_optim(self, data, **kwargs):
        raise NotImplementedError('code optimisation is outside the accelerated hot paths (SURVEY.md §8 f1)')

    def train_step(self, data, optimizer, running_status=None):
        raise NotImplementedError('training is outside the accelerated hot paths (SURVEY.md §8 f2)')

    def eval_and_viz(self, data, decoder, code, density_bitfield, viz_dir=None, cfg=dict()):
        """base_nerf.py:535-553 (render + clamp + 8-bit rounding); metrics against ground truth are out of scope."""
        h, w = cfg['img_size']
        image, depth = self.render(decoder, code, density_bitfield, h, w, data['test_intrinsics'], data['test_poses'], cfg=cfg)
        num_scenes, num_imgs = image.shape[:2]
        pred_imgs = image.permute(0, 1, 4, 2, 3).reshape(num_scenes * num_imgs, 3, h, w).clamp(min=0, max=1)
        pred_imgs = torch.round(pred_imgs * 255) / 255
        return dict(), pred_imgs.reshape(num_scenes, num_imgs, 3, h, w)

    @torch.no_grad()
    def val_step(self, data, viz_dir=None, viz_dir_guide=None, **kwargs):
        """diffusion_nerf.py:406-469 for the unconditional branch."""
        decoder = self.decoder_ema if self.decoder_use_ema else self.decoder
        if 'code' in data or 'cond_imgs' in data:
            raise NotImplementedError('val_step: only unconditional generation is accelerated in this round')
        code, density_grid, density_bitfield = self.val_uncond(data, **kwargs)
        if 'test_poses' in data:
            log_vars, pred_imgs = self.eval_and_viz(data, decoder, code, density_bitfield, viz_dir=viz_dir, cfg=self.test_cfg)
        else:
            log_vars, pred_imgs = dict(), None
        return dict(log_vars=log_vars, num_samples=len(data['scene_id']), pred_imgs=pred_imgs)
