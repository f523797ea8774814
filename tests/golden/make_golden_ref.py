"""Generates tests/golden/reference_v1.npz by RUNNING THE REFERENCE'S OWN PYTHON (build container only:
`python tests/golden/make_golden_ref.py`; /root/reference does not exist on the GPU box, the fixtures are committed).

The reference cannot be imported as a package here (mmcv / mmgen / matplotlib / lpips are not installed and `np.cumproduct`
is gone from NumPy 2), so this script loads three of its files BY PATH with the missing third-party names stubbed in
`sys.modules`:

  executed from /root/reference, unmodified ........................................................ [reference]
    lib/models/architecture/ddpm/denoising.py : DenoisingUnetMod.__init__ (:14-187: block list, channel bookkeeping,
        attention placement, skip-channel stack) and DenoisingUnetMod.forward (:191-216: time embedding, skip push /
        pop + concat order, output head)
    lib/models/architecture/ddpm/modules.py   : MultiHeadAttentionMod.__init__/forward (:12-48: qkv reshape to heads --
        the "legacy" head layout -- and the residual), DenoisingResBlockMod.__init__ (:51-110: conv_1 / norm_with_embedding /
        conv_2 / shortcut construction => the state-dict keys), DenoisingDownsampleMod / DenoisingUpsampleMod.__init__ (:113-129)
    lib/models/diffusions/gaussian_diffusion.py : linear / cosine schedules (:64-110), get_betas (:112-129),
        prepare_diffusion_vars (:131-154, with `np.cumproduct = np.cumprod` patched in), pred_x_0 (:180-240, incl. both
        guidance branches), p_sample_langevin (:242-262), p_sample_ddim (:264-293), ddim_sample (:295-331), q_sample (:166-178)
    lib/models/diffusions/sampler.py          : SNRWeightedTimeStepSampler.__init__ (:15-46: the per-timestep loss weights)
    lib/core/utils/nerf_utils.py              : get_ray_directions / get_rays / get_cam_rays (:17-61) (mcubes stubbed, unused)
    lib/ops/activation.py                     : _trunc_exp forward / backward (:8-23)
    lib/models/losses/reg_loss.py, tv_loss.py : RegLoss, TVLoss (mmgen `weighted_loss` decorator stubbed: mean reduction)
    lib/models/autodecoders/base_nerf.py      : TanhCode, NormalizedTanhCode (:25-76), BaseNeRF.ray_sample / get_raybatch_inds (:231-274) with
        seeded CPU randperm draws (matplotlib / lpips / trimesh / mmcv.runner / lib.core / lib.ops stubbed: none of them is touched)
    lib/models/losses/ddpm_loss.py            : DDPMMSELossMod (:12-131: 0.5 * flat-mean MSE, weight[t] * weight_scale rescale, norm_factor)
        driven by gaussian_diffusion.py forward_train / loss (:404-450) with the timestep draw and the noise injected

  supplied by the stubs below, restated from mmgen 0.7.2 / mmcv-full 1.6.0 FROM MEMORY ............... [mmgen-memory]
    mmgen.models.architectures.ddpm.modules : TimeEmbedding (sinusoidal cos|sin -> Linear -> SiLU -> Linear), EmbedSequential,
        DenoisingResBlock.forward / forward_shortcut / init_weights, NormWithEmbedding, MultiHeadAttention.QKVAttention /
        init_weights, DenoisingDownsample.forward, DenoisingUpsample.forward
    mmgen.models.architectures.ddpm.denoising : DenoisingUnet.init_weights (zero conv_2 / out / proj)
    mmgen.models.diffusions.utils : var_to_tensor, _get_noise_batch;  mmgen.models.architectures.common.get_module_device
    mmgen.models.diffusions.UniformTimeStepSampler (base class only; its sampling is not exercised)
    mmgen.models.losses.ddpm_loss : DDPMLoss base (rescale_mode='timestep_weight' from sampler.weight, reduce), mse_loss('flatmean'), reduce_loss
    mmcv.cnn.bricks : build_norm_layer (GN, eps 1e-5), build_activation_layer (SiLU), ConvModule(order=('norm','act','conv'))
    mmgen.models.builder : MODULES registry / build_module

So the fixtures pin: the UNet WIRING and key layout, the attention head layout, and ALL of the diffusion algebra to the
reference's own code; the inner forwards of the mmgen-inherited blocks remain memory restatements (same text as SURVEY.md
Appendix B) and are therefore only as good as that memory -- this is said again in oracle/unet_port.py and DESIGN.md.
"""
import importlib.util
import math
import os
import sys
import types
from functools import partial

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

REF = '/root/reference'
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


# ----------------------------------------------------------------------------- stubs [mmgen-memory]
class _Registry:
    def __init__(self):
        self.d = {}

    def register_module(self, name=None, force=False, module=None):
        def deco(cls):
            self.d[name or cls.__name__] = cls
            return cls
        return deco(module) if module is not None else deco

    def get(self, k):
        return self.d[k]


MODULES = _Registry()


def build_module(cfg, default_args=None):
    if isinstance(cfg, nn.Module):
        return cfg
    args = dict(cfg)
    if default_args:
        for k, v in default_args.items():
            args.setdefault(k, v)
    return MODULES.get(args.pop('type'))(**args)


def build_norm_layer(cfg, num_features, postfix=''):
    assert cfg['type'] == 'GN'
    return 'gn' + str(postfix), nn.GroupNorm(cfg['num_groups'], num_features, eps=cfg.get('eps', 1e-5))


def build_activation_layer(cfg):
    assert cfg['type'] == 'SiLU'
    return nn.SiLU()


def constant_init(m, val, bias=0):
    nn.init.constant_(m.weight, val)
    if getattr(m, 'bias', None) is not None:
        nn.init.constant_(m.bias, bias)


class ConvModule(nn.Module):
    """mmcv ConvModule restricted to order=('norm','act','conv') with GN + SiLU, as denoising.py:178-187 uses it."""

    def __init__(self, in_channels, out_channels, kernel_size, padding=0, groups=1, act_cfg=None, norm_cfg=None, bias=True,
                 order=('conv', 'norm', 'act')):
        super().__init__()
        assert tuple(order) == ('norm', 'act', 'conv')
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, padding=padding, groups=groups, bias=bias)
        self.norm_name, norm = build_norm_layer(norm_cfg, in_channels)      # norm before conv => in_channels
        self.add_module(self.norm_name, norm)
        self.activate = build_activation_layer(act_cfg)

    def forward(self, x):
        return self.conv(self.activate(getattr(self, self.norm_name)(x)))


class EmbedSequential(nn.Sequential):
    def forward(self, x, y):
        for layer in self:
            x = layer(x, y) if isinstance(layer, DenoisingResBlock) else layer(x)
        return x


class TimeEmbedding(nn.Module):
    def __init__(self, in_channels, embedding_channels, embedding_mode='sin', embedding_cfg=None, act_cfg=dict(type='SiLU', inplace=False)):
        super().__init__()
        self.blocks = nn.Sequential(nn.Linear(in_channels, embedding_channels), build_activation_layer(act_cfg),
                                    nn.Linear(embedding_channels, embedding_channels))
        cfg = dict(dim=in_channels)
        if embedding_cfg is not None:
            cfg.update(embedding_cfg)
        assert embedding_mode.upper() == 'SIN'
        self.embedding_fn = partial(self.sinusodial_embedding, **cfg)

    @staticmethod
    def sinusodial_embedding(timesteps, dim, max_period=10000):
        half = dim // 2
        freqs = torch.exp(-np.log(max_period) * torch.arange(start=0, end=half, dtype=torch.float32) / half).to(device=timesteps.device)
        args = timesteps[:, None].float() * freqs[None]
        embedding = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
        if dim % 2:
            embedding = torch.cat([embedding, torch.zeros_like(embedding[:, :1])], dim=-1)
        return embedding

    def forward(self, t):
        return self.blocks(self.embedding_fn(t))


@MODULES.register_module()
class NormWithEmbedding(nn.Module):
    def __init__(self, in_channels, embedding_channels, norm_cfg=dict(type='GN', num_groups=32), act_cfg=dict(type='SiLU', inplace=False),
                 use_scale_shift=True):
        super().__init__()
        self.use_scale_shift = use_scale_shift
        _, self.norm = build_norm_layer(norm_cfg, in_channels)
        self.embedding_layer = nn.Sequential(build_activation_layer(act_cfg),
                                             nn.Linear(embedding_channels, in_channels * 2 if use_scale_shift else in_channels))

    def forward(self, x, y):
        embedding = self.embedding_layer(y)[:, :, None, None]
        if self.use_scale_shift:
            scale, shift = torch.chunk(embedding, 2, dim=1)
            x = self.norm(x)
            x = x * (1 + scale) + shift
        else:
            x = self.norm(x + embedding)
        return x


class DenoisingResBlock(nn.Module):
    def forward_shortcut(self, x):
        return self.shortcut(x) if self.learnable_shortcut else x

    def forward(self, x, y):
        shortcut = self.forward_shortcut(x)
        x = self.conv_1(x)
        x = self.norm_with_embedding(x, y)
        x = self.conv_2(x)
        return x + shortcut

    def init_weights(self):
        constant_init(self.conv_2[-1], 0)


class MultiHeadAttention(nn.Module):
    @staticmethod
    def QKVAttention(qkv):
        channel = qkv.shape[1] // 3
        q, k, v = torch.chunk(qkv, 3, dim=1)
        scale = 1 / np.sqrt(np.sqrt(channel))
        weight = torch.einsum('bct,bcs->bts', q * scale, k * scale)
        weight = torch.softmax(weight.float(), dim=-1).type(weight.dtype)
        return torch.einsum('bts,bcs->bct', weight, v)

    def init_weights(self):
        constant_init(self.proj, 0)


class DenoisingDownsample(nn.Module):
    def forward(self, x):
        return self.downsample(x)


class DenoisingUpsample(nn.Module):
    def forward(self, x):
        x = F.interpolate(x, scale_factor=2, mode='nearest')
        if self.with_conv:
            x = self.conv(x)
        return x


class DenoisingUnet(nn.Module):
    def init_weights(self, pretrained=None):
        assert pretrained is None
        for n, m in self.named_modules():
            if isinstance(m, nn.Conv2d) and ('conv_2' in n or ('out' in n and 'out_blocks' not in n)):
                constant_init(m, 0)
            if isinstance(m, nn.Conv1d) and 'proj' in n:
                constant_init(m, 0)


class UniformTimeStepSampler:
    def __init__(self, num_timesteps):
        self.num_timesteps = num_timesteps
        self.prob = [1 / num_timesteps for _ in range(num_timesteps)]


class DDPMLoss(nn.Module):
    """mmgen DDPMLoss restricted to what DDPMMSELossMod uses: rescale_mode None | 'timestep_weight', log collection dropped"""

    def __init__(self, rescale_mode=None, rescale_cfg=None, log_cfgs=None, weight=None, sampler=None, reduction='mean', loss_name=None):
        super().__init__()
        self.log_vars, self.reduction, self._loss_name, self.sampler = dict(), reduction, loss_name, sampler
        if rescale_mode is None:
            self.rescale_fn = lambda loss, t: loss
        else:
            assert rescale_mode == 'timestep_weight'
            w = weight if weight is not None else sampler.weight.clone()
            self.rescale_fn = partial(self.timestep_weight_rescale, weight=w)

    def collect_log(self, loss, timesteps):
        self.log_vars = {self._loss_name: float(loss.mean())}


def mse_loss(pred, target, reduction='flatmean'):
    assert reduction == 'flatmean'
    return F.mse_loss(pred, target, reduction='none').flatten(1).mean(dim=1)


def reduce_loss(loss, reduction):
    return dict(mean=loss.mean, sum=loss.sum, none=lambda: loss)[reduction]()


def get_module_device(module):
    return next(module.parameters()).device


def var_to_tensor(var, index, target_shape=None, device=None):
    var_indexed = torch.from_numpy(var)[index].float()
    if device is not None:
        var_indexed = var_indexed.to(device)
    while len(var_indexed.shape) < len(target_shape):
        var_indexed = var_indexed[..., None]
    return var_indexed


def _get_noise_batch(noise, image_shape, num_timesteps=0, num_batches=0, timesteps_noise=False):
    assert noise is None and not timesteps_noise
    return torch.randn((num_batches, *image_shape))


class _ProgressBar:
    def __init__(self, *a, **k):
        pass

    def update(self):
        pass


def _install_stubs():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m
    mod('mmcv', ProgressBar=_ProgressBar)
    mod('mmcv.cnn')
    mod('mmcv.cnn.bricks', build_activation_layer=build_activation_layer, build_norm_layer=build_norm_layer)
    mod('mmcv.cnn.bricks.conv_module', ConvModule=ConvModule)
    mod('mmgen')
    mod('mmgen.models', MODULES=MODULES)
    mod('mmgen.models.builder', MODULES=MODULES, build_module=build_module)
    mod('mmgen.models.architectures')
    mod('mmgen.models.architectures.common', get_module_device=get_module_device)
    mod('mmgen.models.architectures.ddpm')
    mod('mmgen.models.architectures.ddpm.modules', MultiHeadAttention=MultiHeadAttention, DenoisingResBlock=DenoisingResBlock,
        DenoisingDownsample=DenoisingDownsample, DenoisingUpsample=DenoisingUpsample, TimeEmbedding=TimeEmbedding,
        EmbedSequential=EmbedSequential)
    mod('mmgen.models.architectures.ddpm.denoising', DenoisingUnet=DenoisingUnet)
    mod('mmgen.models.diffusions', UniformTimeStepSampler=UniformTimeStepSampler)
    mod('mmgen.models.diffusions.utils', var_to_tensor=var_to_tensor, _get_noise_batch=_get_noise_batch)
    mod('mmgen.models.losses')
    mod('mmgen.models.losses.ddpm_loss', DDPMLoss=DDPMLoss, mse_loss=mse_loss, reduce_loss=reduce_loss)
    # the relative import `from ...core import reduce_mean` of lib/models/losses/ddpm_loss.py: single-process identity
    for name in ('reflib', 'reflib.models', 'reflib.models.losses'):
        pkg = mod(name)
        pkg.__path__ = []
    mod('reflib.core', reduce_mean=lambda x: x)
    if not hasattr(np, 'cumproduct'):
        np.cumproduct = np.cumprod          # removed in NumPy 2; gaussian_diffusion.py:134 calls it


def _load(relpath, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, relpath))
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    spec.loader.exec_module(m)
    return m


def weighted_loss(fn):
    """mmgen.models.losses.utils.weighted_loss [mmgen-memory]: elementwise weight, then 'mean' (or sum / avg_factor) reduction"""
    def wrapper(*args, weight=None, reduction='mean', avg_factor=None, **kwargs):
        loss = fn(*args, **kwargs)
        if weight is not None:
            loss = loss * weight
        if avg_factor is not None:
            return loss.sum() / avg_factor
        return loss.mean() if reduction == 'mean' else (loss.sum() if reduction == 'sum' else loss)
    return wrapper


def load_reference_host_helpers():
    """nerf_utils.py, activation.py, reg_loss.py, tv_loss.py, base_nerf.py of the reference, executed from /root/reference"""
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m
    mod('mcubes')
    nu = _load('lib/core/utils/nerf_utils.py', 'ref_nerf_utils')
    act = _load('lib/ops/activation.py', 'ref_activation')
    mod('mmgen.models.losses.utils', weighted_loss=weighted_loss)
    reg = _load('lib/models/losses/reg_loss.py', 'ref_reg_loss')
    tv = _load('lib/models/losses/tv_loss.py', 'ref_tv_loss')
    # base_nerf.py: heavy imports that its code-activation classes and static ray helpers never touch
    mod('matplotlib'); mod('matplotlib.pyplot'); mod('lpips'); mod('trimesh')
    mod('mmcv.runner', load_checkpoint=None)
    sys.modules['mmcv'].runner = sys.modules['mmcv.runner']
    names = ('custom_meshgrid', 'eval_psnr', 'eval_ssim_skimage', 'rgetattr', 'rsetattr', 'extract_geometry', 'module_requires_grad')
    core = sys.modules['reflib.core']
    for n in names:
        setattr(core, n, None)
    core.get_cam_rays = nu.get_cam_rays
    mod('lib'); mod('lib.ops', morton3D=None, morton3D_invert=None, packbits=None)
    for name in ('reflib.models.autodecoders',):
        sys.modules[name] = types.ModuleType(name); sys.modules[name].__path__ = []
    base = _load('lib/models/autodecoders/base_nerf.py', 'reflib.models.autodecoders.base_nerf')
    return nu, act, reg, tv, base


def host_helper_fixtures(out):
    nu, act, reg, tv, base = load_reference_host_helpers()
    g = torch.Generator().manual_seed(21)
    # cameras -> rays
    R_ = torch.linalg.qr(torch.randn(2, 2, 3, 3, generator=g))[0]
    c2w = torch.cat([torch.cat([R_, torch.randn(2, 2, 3, 1, generator=g)], dim=-1), torch.tensor([0., 0, 0, 1]).expand(2, 2, 1, 4)], dim=-2)
    intr = torch.tensor([[[40.0, 41.0, 15.5, 16.25], [131.25, 131.25, 64.0, 64.0]]]).repeat(2, 1, 1)
    ro, rd = nu.get_cam_rays(c2w, intr, 6, 5)
    out['cam_c2w'], out['cam_intr'], out['cam_rays_o'], out['cam_rays_d'] = c2w.numpy(), intr.numpy(), ro.contiguous().numpy(), rd.numpy()
    # trunc_exp forward / backward
    x = torch.tensor([-30.0, -14.0, -1.0, 0.0, 0.5, 3.0, 14.0, 30.0], requires_grad=True)
    y = act.trunc_exp(x)
    y.backward(torch.ones_like(y))
    out['trunc_exp_x'], out['trunc_exp_y'], out['trunc_exp_grad'] = x.detach().numpy(), y.detach().numpy(), x.grad.numpy()
    # latent regularisers
    code = torch.randn(2, 3, 6, 8, 8, generator=g)
    out['loss_code'] = code.numpy()
    out['reg_loss_p2'] = np.array(float(reg.RegLoss(power=2, loss_weight=3e-3)(code)))
    out['reg_loss_p1'] = np.array(float(reg.RegLoss(power=1, loss_weight=0.5)(code)))
    out['tv_loss_p15'] = np.array(float(tv.TVLoss(power=1.5, loss_weight=1.0)(code)))
    # code activations
    z = torch.randn(3, 6, 4, 4, generator=g) * 1.5
    out['act_z'] = z.numpy()
    t2 = base.TanhCode(scale=2)
    out['tanh2_fwd'], out['tanh2_inv'] = t2(z).numpy(), t2.inverse(t2(z) * 0.9).numpy()
    nt = base.NormalizedTanhCode(mean=0.0, std=0.5, clip_range=2)
    nt.running_mean.fill_(0.1); nt.running_var.fill_(0.3)
    out['ntanh_fwd'], out['ntanh_inv'] = nt(z).numpy(), nt.inverse(nt(z) * 0.9).numpy()
    nt.train()
    nt(z, update_stats=True)
    out['ntanh_running_mean_after'], out['ntanh_running_var_after'] = nt.running_mean.numpy().copy(), nt.running_var.numpy().copy()
    # ray batches: seeded CPU randperm draws
    imgs = torch.rand(2, 3, 4, 4, 3, generator=g)
    rays_o, rays_d = torch.rand(2, 3, 4, 4, 3, generator=g), torch.rand(2, 3, 4, 4, 3, generator=g)
    out['rb_imgs'], out['rb_rays_o'], out['rb_rays_d'] = imgs.numpy(), rays_o.numpy(), rays_d.numpy()
    torch.manual_seed(123)
    inds, nb = base.BaseNeRF.get_raybatch_inds(imgs, 20)
    out['rb_inds'], out['rb_num'] = torch.cat(list(inds), dim=1).numpy(), np.array(nb)
    torch.manual_seed(321)
    o, d, t = base.BaseNeRF.ray_sample(rays_o, rays_d, imgs, 10)
    out['rs_o'], out['rs_d'], out['rs_t'] = o.numpy(), d.numpy(), t.numpy()
    o, d, t = base.BaseNeRF.ray_sample(rays_o, rays_d, imgs, 20, sample_inds=inds[1])
    out['rs_o_given'], out['rs_t_given'] = o.numpy(), t.numpy()


def load_reference():
    """-> (modules.py, denoising.py, gaussian_diffusion.py, sampler.py) of the reference, executed from /root/reference."""
    _install_stubs()
    mods = _load('lib/models/architecture/ddpm/modules.py', 'ref_ddpm_modules')
    den = _load('lib/models/architecture/ddpm/denoising.py', 'ref_ddpm_denoising')
    gd = _load('lib/models/diffusions/gaussian_diffusion.py', 'ref_gaussian_diffusion')
    sm = _load('lib/models/diffusions/sampler.py', 'ref_sampler')
    _load('lib/models/losses/ddpm_loss.py', 'reflib.models.losses.ddpm_loss')
    return mods, den, gd, sm


# ----------------------------------------------------------------------------- fixtures
UNET_CFG = dict(image_size=16, in_channels=18, base_channels=32, channels_cfg=[1, 2, 2], resblocks_per_downsample=2, dropout=0.0,
                use_scale_shift_norm=True, downsample_conv=True, upsample_conv=True, num_heads=2, attention_res=[8, 4])


class _NullLoss(nn.Module):
    def __init__(self, sampler=None, **kw):
        super().__init__()
        self.log_vars = {}


def seeded_state_dict(model, seed):
    """every tensor of the reference module's own state dict drawn from a seeded generator, in key order
    (tests/common.py:seeded_weights regenerates the same tensors from the committed key / shape lists)"""
    from tests.common import seeded_weights
    sd = model.state_dict()
    return seeded_weights(list(sd.keys()), [tuple(v.shape) for v in sd.values()], seed)


def main():
    mods, den, gd, sm = load_reference()
    MODULES.register_module(name='NullLoss', module=_NullLoss)
    out = {}
    torch.manual_seed(0)
    # ---- UNet built by the reference constructor; forward = the reference's forward + stubbed block bodies
    unet = den.DenoisingUnetMod(**UNET_CFG)
    sd = seeded_state_dict(unet, seed=11)
    unet.load_state_dict(sd)
    unet.eval()
    keys = list(sd.keys())
    out['unet_keys'] = np.array(keys)
    out['unet_shapes'] = np.array([','.join(map(str, sd[k].shape)) for k in keys])
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 18, 16, 16, generator=g)
    t = torch.tensor([999, 17])
    with torch.no_grad():
        out['unet_x'], out['unet_t'] = x.numpy(), t.numpy()
        out['unet_y'] = unet(x, t).numpy()
    # d v . r / d x_t through the reference module (the guidance path differentiates the UNet w.r.t. its input)
    xr = x.clone().requires_grad_(True)
    r = torch.randn(2, 18, 16, 16, generator=g)
    (unet(xr, t) * r).sum().backward()
    out['unet_r'], out['unet_dx'] = r.numpy(), xr.grad.numpy()
    # ---- full-size constructor: key list + shapes only (the released checkpoints' layout)
    full = den.DenoisingUnetMod(image_size=128, in_channels=18, base_channels=128, channels_cfg=[1, 2, 2, 4, 4], resblocks_per_downsample=2,
                                dropout=0.0, use_scale_shift_norm=True, downsample_conv=True, upsample_conv=True, num_heads=4,
                                attention_res=[32, 16, 8])
    fsd = full.state_dict()
    out['full_keys'] = np.array(list(fsd.keys()))
    out['full_shapes'] = np.array([','.join(map(str, v.shape)) for v in fsd.values()])
    out['full_numel'] = np.array(sum(v.numel() for v in fsd.values()), np.int64)
    full_do = den.DenoisingUnetMod(image_size=128, in_channels=18, base_channels=128, channels_cfg=[1, 2, 2, 4, 4], resblocks_per_downsample=2,
                                   dropout=0.1, use_scale_shift_norm=True, num_heads=4, attention_res=[32, 16, 8])
    out['full_dropout_keys'] = np.array(list(full_do.state_dict().keys()))
    del full, full_do
    # ---- diffusion tables + algebra
    test_cfg = dict(num_timesteps=10, clip_range=[-2, 2], guidance_gain=37.5, snr_weight_power=0.25, langevin_steps=2, langevin_delta=0.4)
    diff = gd.GaussianDiffusion(denoising=unet, ddpm_loss=dict(type='NullLoss'), betas_cfg=dict(type='linear'), num_timesteps=1000,
                                timestep_sampler=dict(type='SNRWeightedTimeStepSampler', power=0.25), denoising_mean_mode='V',
                                test_cfg=test_cfg)
    for name in ('betas', 'alphas_bar', 'alphas_bar_prev', 'sqrt_alphas_bar', 'sqrt_one_minus_alphas_bar', 'sqrt_recip_alplas_bar',
                 'sqrt_recipm1_alphas_bar', 'tilde_betas_t', 'log_tilde_betas_t_clipped', 'tilde_mu_t_coef1', 'tilde_mu_t_coef2'):
        out['lin_' + name] = np.asarray(getattr(diff, name), np.float64)
    out['snr_weight_p025_V'] = diff.sampler.weight.numpy()
    cosd = gd.GaussianDiffusion(denoising=unet, ddpm_loss=dict(type='NullLoss'), betas_cfg=dict(type='cosine'), num_timesteps=1000,
                                timestep_sampler=dict(type='SNRWeightedTimeStepSampler', power=0.5, mode='V'))
    out['cos_alphas_bar'] = np.asarray(cosd.alphas_bar, np.float64)
    out['snr_weight_p05_V'] = cosd.sampler.weight.numpy()
    x_t = torch.randn(2, 18, 16, 16, generator=g) * 1.3
    out['x_t'] = x_t.numpy()
    with torch.no_grad():
        # unguided pred_x_0 / p_sample_ddim / p_sample_langevin on fixed tensors
        x0, v = diff.pred_x_0(x_t.clone(), torch.tensor(600), cfg=test_cfg)
        out['pred_x0_t600'], out['pred_v_t600'] = x0.numpy(), v.numpy()
        xp, x0b = diff.p_sample_ddim(x_t.clone(), torch.tensor(600), torch.tensor(500), cfg=test_cfg)
        out['ddim_prev_600_500'] = xp.numpy()
        xp, _ = diff.p_sample_ddim(x_t.clone(), torch.tensor(99), -1, cfg=test_cfg)
        out['ddim_prev_99_last'] = xp.numpy()
        xp, _ = diff.p_sample_ddim(x_t.clone(), torch.tensor(600), torch.tensor(500), noise=r, cfg=dict(test_cfg, eta=0.7))
        out['ddim_prev_600_500_eta07'] = xp.numpy()
        xl = diff.p_sample_langevin(x_t.clone(), torch.tensor(500), noise=r, cfg=test_cfg)
        out['langevin_500'] = xl.numpy()
        # whole 10-step loop without langevin
        diff.test_cfg = dict(test_cfg, langevin_steps=0)
        out['ddim10'] = diff.ddim_sample(x_t.clone()).numpy()
        inter = diff.ddim_sample(x_t.clone(), save_intermediates=True)
        out['ddim10_intermediates_0_1_2_3_18_19'] = torch.stack([inter[i] for i in (0, 1, 2, 3, 18, 19)]).numpy()
        diff.test_cfg = test_cfg
    # guidance: quadratic pull towards a target, both gradient routes (gaussian_diffusion.py:213-222)
    target = torch.randn(2, 18, 16, 16, generator=g)
    out['guide_target'] = target.numpy()

    def guide(x0):
        return 0.5 * ((x0 - target) ** 2).mean() * x0.size(0)

    for through in (True, False):
        cfg = dict(test_cfg, grad_through_unet=through)
        with torch.no_grad():
            x0, v = diff.pred_x_0(x_t.clone(), torch.tensor(600), grad_guide_fn=guide, cfg=cfg, update_denoising_output=True)
            tag = 'thru' if through else 'x0'
            out[f'guided_x0_{tag}'], out[f'guided_v_{tag}'] = x0.numpy(), v.numpy()
    # guided loop with langevin correction (noise injected through a patched generator: fixed tensors instead of randn)
    g2 = torch.Generator().manual_seed(77)          # tests regenerate these 18 tensors from the seed
    noises = [torch.randn(2, 18, 16, 16, generator=g2) for _ in range(2 * 9)]
    out['langevin_noise_seed'] = np.array(77)
    it = iter(noises)
    sys.modules['ref_gaussian_diffusion']._get_noise_batch = lambda *a, **k: next(it)
    with torch.no_grad():
        diff.test_cfg = dict(test_cfg, langevin_t_range=[0, 1000])
        out['guided_langevin_ddim10'] = diff.ddim_sample(x_t.clone(), grad_guide_fn=guide).numpy()
    # q_sample
    eps = torch.randn(2, 18, 16, 16, generator=g)
    xq, mean, std = diff.q_sample(target, torch.tensor([10, 900]), noise=eps)
    out['q_eps'], out['q_sample'] = eps.numpy(), xq.numpy()
    host_helper_fixtures(out)
    out['unet_weight_seed'] = np.array(11)
    out['unet_weight_checksum'] = np.array(sum(float(v.double().sum()) for v in sd.values()))
    # ---- diffusion loss as val_optim uses it (chairs_recons1v settings: SNR power 0.25, v-target, weight_scale, scale_norm)
    dl = gd.GaussianDiffusion(denoising=unet, betas_cfg=dict(type='linear'), num_timesteps=1000, denoising_mean_mode='V',
                              timestep_sampler=dict(type='SNRWeightedTimeStepSampler', power=0.25),
                              ddpm_loss=dict(type='DDPMMSELossMod', rescale_mode='timestep_weight', data_info=dict(pred='v_t_pred', target='v_t'),
                                             weight_scale=4.0, scale_norm=True, loss_name='loss_ddpm_mse'))
    dl.eval()
    dl.ddpm_loss.norm_factor.fill_(0.37)
    t_fix = torch.tensor([12, 870])
    dl.sampler = lambda n: t_fix
    sys.modules['ref_gaussian_diffusion']._get_noise_batch = lambda *a, **k: eps
    x0r = target.clone().requires_grad_(True)
    for p_ in dl.parameters():
        p_.requires_grad_(False)
    loss, _ = dl.forward_train(x0r, cfg=dict(clip_range=[-2, 2]))
    loss.backward()
    out['train_t'], out['train_loss'], out['train_dx0'] = t_fix.numpy(), np.array(float(loss)), x0r.grad.numpy()
    np.savez_compressed(os.path.join(HERE, 'reference_v1.npz'), **out)
    print({k: (v.shape if hasattr(v, 'shape') else v) for k, v in out.items()})


if __name__ == '__main__':
    main()
