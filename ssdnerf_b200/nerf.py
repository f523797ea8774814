"""`DiffusionNeRF` -- the registered model of the reference's configs, inference side.

Plugin surface of lib/models/autodecoders/{base_nerf,multiscene_nerf,diffusion_nerf}.py for the two accelerated hot
paths: `val_uncond` (noise -> DDIM -> code -> occupancy grid), `get_density`, `render`, `val_step`, `code_diff_pr(_inv)`;
the same constructor kwargs and `outputs_dict = {log_vars, num_samples, pred_imgs}` contract (diffusion_nerf.py:464-469).
Guided generation (`val_guide`, `loss`, `ray_sample`, `get_raybatch_inds`, `update_extra_state`) runs on the fused
differentiable renderer with the guidance gradient taken w.r.t. x_0 (`grad_through_unet=False`); back-propagating through
the UNet, `val_optim` and `train_step` are SURVEY.md §8(f) rows and raise NotImplementedError.
"""
import math
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


@MODULES.register_module()
class MSELoss(nn.Module):
    """mmgen.models.losses.pixelwise_loss.MSELoss [mmgen-memory]: mean squared error * loss_weight."""

    def __init__(self, loss_weight=1.0, reduction='mean', **kwargs):
        super().__init__()
        assert reduction == 'mean'
        self.loss_weight = loss_weight

    def forward(self, pred, target, **kwargs):
        return torch.nn.functional.mse_loss(pred, target) * self.loss_weight


@MODULES.register_module()
class RegLoss(nn.Module):
    """lib/models/losses/reg_loss.py:7-30"""

    def __init__(self, power=1, loss_weight=1.0):
        super().__init__()
        self.power, self.loss_weight = power, loss_weight

    def forward(self, tensor, **kwargs):
        return (tensor.abs().mean() if self.power == 1 else (tensor.abs() ** self.power).mean()) * self.loss_weight


class _RenderMSELoss(torch.autograd.Function):
    """pixel + reg terms of BaseNeRF.loss as ONE differentiable op on the fused renderer: forward = render_train_fwd +
    mse_render_loss (which also leaves d/d image, d/d weights_sum), backward = render_train_bwd (+ RegLoss gradient)."""

    @staticmethod
    def forward(ctx, code, rays_o, rays_d, target, bitfield, blob, noises, dt_gamma, cfg, bg_color, pix_coef, reg_weight):
        planes = R.pack_planes(code, R.DEC_P)
        hw = tuple(code.shape[-2:])
        out = R.render_train_fwd(planes, hw, bitfield, blob, rays_o, rays_d, noises=noises, dt_gamma=dt_gamma, **cfg)
        n_el = target.numel()
        loss, g_image, g_ws, rgb = R.mse_render_loss(out['image'], out['weights_sum'], target, bg_color, pix_coef / n_el,
                                                     2.0 * pix_coef / n_el, want_rgb=True)
        reg = None
        if reg_weight is not None:
            reg = code.square().mean() * reg_weight
        ctx.save_for_backward(planes, rays_o, rays_d, bitfield, blob, noises, dt_gamma, out['weights_sum'], out['image'], g_image, g_ws,
                              code if reg_weight is not None else None)
        ctx.hw, ctx.cfg, ctx.reg_weight = hw, cfg, reg_weight
        ctx.mark_non_differentiable(rgb)
        return loss.reshape(()), (reg if reg is not None else loss.new_zeros(())), rgb

    @staticmethod
    def backward(ctx, g_pix, g_reg, g_rgb):
        planes, rays_o, rays_d, bitfield, blob, noises, dt_gamma, ws, image, g_image, g_ws, code = ctx.saved_tensors
        grad = R.render_train_bwd(planes, ctx.hw, bitfield, blob, rays_o, rays_d, ws, image, g_ws, g_image, noises=noises,
                                  dt_gamma=dt_gamma, **ctx.cfg)
        if code is not None:      # RegLoss(power=2): d/d code = 2 w code / numel, with its own upstream gradient
            return (torch.addcmul(grad * g_pix, code, g_reg, value=2.0 * ctx.reg_weight / code.numel()),) + (None,) * 11
        return (grad * g_pix,) + (None,) * 11


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
        self.pixel_loss = build_module(pixel_loss if pixel_loss is not None else dict(type='MSELoss'))   # base_nerf.py:90-92
        self.reg_loss = build_module(reg_loss) if reg_loss is not None else None
        # training-only collaborators are kept as configuration
        self._unbuilt = dict(cache_size=cache_size, cache_16bit=cache_16bit)

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

    # ------------------------------------------------------------------ guidance collaborators (base_nerf.py:231-296, 318-389)
    @staticmethod
    def ray_sample(cond_rays_o, cond_rays_d, cond_imgs, n_samples, sample_inds=None):
        """base_nerf.py:231-261"""
        device = cond_rays_o.device
        num_scenes, num_imgs, h, w, _ = cond_rays_o.size()
        num_scene_pixels = num_imgs * h * w
        rays_o = cond_rays_o.reshape(num_scenes, num_scene_pixels, 3)
        rays_d = cond_rays_d.reshape(num_scenes, num_scene_pixels, 3)
        target_rgbs = cond_imgs.reshape(num_scenes, num_scene_pixels, 3)
        if num_scene_pixels > n_samples:
            if sample_inds is None:
                sample_inds = torch.stack([torch.randperm(num_scene_pixels, device=device)[:n_samples] for _ in range(num_scenes)], dim=0)
            scene_arange = torch.arange(num_scenes, device=device)[:, None]
            rays_o, rays_d, target_rgbs = rays_o[scene_arange, sample_inds], rays_d[scene_arange, sample_inds], target_rgbs[scene_arange, sample_inds]
        return rays_o, rays_d, target_rgbs

    @staticmethod
    def get_raybatch_inds(cond_imgs, n_inverse_rays):
        """base_nerf.py:263-274"""
        device = cond_imgs.device
        num_scenes, num_imgs, h, w, _ = cond_imgs.size()
        num_scene_pixels = num_imgs * h * w
        if num_scene_pixels > n_inverse_rays:
            raybatch_inds = torch.stack([torch.randperm(num_scene_pixels, device=device) for _ in range(num_scenes)], dim=0)
            raybatch_inds = raybatch_inds.split(n_inverse_rays, dim=1)
            return raybatch_inds, len(raybatch_inds)
        return None, None

    def loss(self, decoder, code, density_bitfield, target_rgbs, rays_o, rays_d, dt_gamma=0.0, return_decoder_loss=False,
             scale_num_ray=1.0, cfg=dict(), perturb=True, **kwargs):
        """base_nerf.py:276-296.  With a frozen shipped-config decoder and the stock MSELoss / RegLoss(power=2) the whole
        chain (render, blend, MSE, reg) is one fused differentiable op; anything else composes the decoder's differentiable
        forward with the loss modules like the reference.  `perturb` may be a [B,N] tensor of start offsets (tests)."""
        scale = 1 - math.exp(-cfg['loss_coef'] * scale_num_ray) if 'loss_coef' in cfg else 1
        fusable = (isinstance(self.pixel_loss, MSELoss) and (self.reg_loss is None or self.reg_loss.power == 2)
                   and not return_decoder_loss and isinstance(rays_o, torch.Tensor) and decoder.training
                   and decoder._fused_train_ok(rays_o, code, self.grid_size))
        if fusable:
            num_scenes = rays_o.size(0)
            rays_o = rays_o.reshape(num_scenes, -1, 3).contiguous().float()
            rays_d = rays_d.reshape(num_scenes, -1, 3).contiguous().float()
            if isinstance(dt_gamma, (int, float)):
                dtg = None if dt_gamma == 0 else torch.full((num_scenes,), float(dt_gamma), device=rays_o.device)
            else:
                dtg = torch.as_tensor(dt_gamma, dtype=torch.float32, device=rays_o.device).reshape(num_scenes).contiguous()
            if isinstance(perturb, torch.Tensor):
                noises = perturb.reshape(num_scenes, -1).contiguous().float()
            else:
                noises = torch.rand(num_scenes, rays_o.size(1), device=rays_o.device) if perturb else None
            rcfg = dict(grid_size=int(self.grid_size), bound=float(decoder.bound), min_near=float(decoder.min_near),
                        max_steps=int(decoder.max_steps), T_thresh=1e-4)
            pixel_loss, reg_loss, out_rgbs = _RenderMSELoss.apply(
                code, rays_o, rays_d, target_rgbs.reshape(num_scenes, -1, 3), density_bitfield.reshape(num_scenes, -1).contiguous(),
                decoder.packed_blob(), noises, dtg, rcfg, float(self.bg_color), float(self.pixel_loss.loss_weight * scale * 3),
                None if self.reg_loss is None else float(self.reg_loss.loss_weight))
            loss, loss_dict = pixel_loss, dict(pixel_loss=pixel_loss)
            if self.reg_loss is not None:
                loss = loss + reg_loss
                loss_dict.update(reg_loss=reg_loss)
            return out_rgbs, loss, loss_dict
        outputs = decoder(rays_o, rays_d, code, density_bitfield, self.grid_size, dt_gamma=dt_gamma, perturb=perturb,
                          return_loss=return_decoder_loss)
        out_weights = outputs['weights_sum']
        out_rgbs = outputs['image'] + self.bg_color * (1 - out_weights.unsqueeze(-1))
        pixel_loss = self.pixel_loss(out_rgbs, target_rgbs, **kwargs) * (scale * 3)
        loss, loss_dict = pixel_loss, dict(pixel_loss=pixel_loss)
        if self.reg_loss is not None:
            reg_loss = self.reg_loss(code, **kwargs)
            loss = loss + reg_loss
            loss_dict.update(reg_loss=reg_loss)
        if return_decoder_loss and outputs['decoder_reg_loss'] is not None:
            loss = loss + outputs['decoder_reg_loss']
            loss_dict.update(decoder_reg_loss=outputs['decoder_reg_loss'])
        return out_rgbs, loss, loss_dict

    def update_extra_state(self, decoder, code, density_grid, density_bitfield, iter_density, density_thresh=0.01, decay=0.9,
                           S=128, jitter=None):
        """base_nerf.py:318-389, full-update branch (iter_density < 16; the partial update only occurs in training)."""
        if iter_density >= 16:
            raise NotImplementedError('partial occupancy-grid update (iter_density >= 16) is a training-only branch (SURVEY.md §8 f2)')
        with torch.no_grad():
            variant = decoder.fused_variant()
            planes = R.pack_planes(code.detach(), variant)
            D.update_extra_state(variant, planes, tuple(code.shape[-2:]), decoder.packed_blob(), density_grid, density_bitfield,
                                 jitter=jitter, density_thresh=density_thresh, decay=decay, grid_size=self.grid_size,
                                 bound=float(decoder.bound))

    # ------------------------------------------------------------------ guided generation (diffusion_nerf.py:241-311)
    def val_guide(self, data, **kwargs):
        """Render-loss guided DDIM. The gradient is taken w.r.t. x_0 (`test_cfg.grad_through_unet=False`); the reference's
        default of differentiating through the UNet is SURVEY.md §8 f1 and raises in `GaussianDiffusion.pred_x_0`."""
        diffusion = self.diffusion_ema if self.diffusion_use_ema else self.diffusion
        decoder = self.decoder_ema if self.decoder_use_ema else self.decoder
        device = next(self.parameters()).device
        cond_imgs, cond_intrinsics, cond_poses = data['cond_imgs'], data['cond_intrinsics'], data['cond_poses']
        N.require_cuda(cond_imgs, cond_intrinsics, cond_poses)
        num_scenes, num_imgs, h, w, _ = cond_imgs.size()
        cond_rays_o, cond_rays_d = R.get_cam_rays(cond_poses, cond_intrinsics, h, w)
        dt_gamma_scale = self.test_cfg.get('dt_gamma_scale', 0.0)
        dt_gamma = dt_gamma_scale / cond_intrinsics[..., :2].mean(dim=(-2, -1))
        if self.image_cond:
            raise NotImplementedError('image-conditioned (concat_cond) denoisers are not used by the shipped configs')
        decoder_training_prev = decoder.training
        decoder.train(True)
        frozen = [(p, p.requires_grad) for m in (diffusion, decoder) for p in m.parameters()]
        for p, _ in frozen:
            p.requires_grad_(False)
        try:
            n_inverse_rays = self.test_cfg.get('n_inverse_rays', 4096)
            raybatch_inds, num_raybatch = self.get_raybatch_inds(cond_imgs, n_inverse_rays)
            density_grid = torch.zeros((num_scenes, self.grid_size ** 3), device=device)
            density_bitfield = torch.zeros((num_scenes, self.grid_size ** 3 // 8), dtype=torch.uint8, device=device)
            state = dict(step=0)

            def grad_guide_fn(x_0_pred):
                code_pred = self.code_diff_pr_inv(x_0_pred)
                self.update_extra_state(decoder, code_pred, density_grid, density_bitfield, 0,
                                        density_thresh=self.test_cfg.get('density_thresh', 0.01))
                inds = raybatch_inds[state['step'] % num_raybatch] if raybatch_inds is not None else None
                rays_o, rays_d, target_rgbs = self.ray_sample(cond_rays_o, cond_rays_d, cond_imgs, n_inverse_rays, sample_inds=inds)
                _, loss, _ = self.loss(decoder, code_pred, density_bitfield, target_rgbs, rays_o, rays_d, dt_gamma,
                                       scale_num_ray=target_rgbs.size(1), cfg=self.test_cfg)
                state['step'] += 1
                return loss * num_scenes

            noise = data.get('noise', None)
            if noise is None:
                noise = torch.randn((num_scenes, *self.code_size), device=device)
            code = diffusion(self.code_diff_pr(noise), return_loss=False, grad_guide_fn=grad_guide_fn, **kwargs)
        finally:
            for p, rg in frozen:
                p.requires_grad_(rg)
            decoder.train(decoder_training_prev)
        return self.code_diff_pr_inv(code), density_grid, density_bitfield

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
