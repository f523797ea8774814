"""PyTorch fp32 restatement of the reference's denoising UNet and DDIM sampler -- TEST INFRASTRUCTURE ONLY.

Constructor logic: lib/models/architecture/ddpm/denoising.py:106-187, modules.py:12-129 (reference).
DDIM algebra: lib/models/diffusions/gaussian_diffusion.py:64-154 (schedules), :180-240 (pred_x_0), :242-331.

PINNING (tests/test_reference_pin_cpu.py against tests/golden/reference_v1.npz, which tests/golden/make_golden_ref.py
produced by EXECUTING the reference's own denoising.py / modules.py / gaussian_diffusion.py / sampler.py with mmcv / mmgen
stubbed): state-dict keys + shapes of the full-size model, the forward wiring (skip stack, concat order, attention head
layout, time embedding path), d out / d x_t, every schedule table (bit-exact float64), pred_x_0 / p_sample_ddim /
p_sample_langevin / ddim_sample incl. guidance through the UNet and w.r.t. x_0, and the SNR loss weights.
STILL [mmgen-memory] (mmgen 0.7.2 is not under /root/reference and not installed): the bodies of
TimeEmbedding, DenoisingResBlock.forward, NormWithEmbedding, MultiHeadAttention.QKVAttention and
DenoisingDownsample/Upsample.forward -- the stubs that stood in for them when the fixtures were generated are the same
restatement, so for these bodies the fixtures pin self-consistency only.

Every function is device-agnostic torch: the GPU parity tests may run it on the CUDA device with TF32 disabled
(`fp32_reference_mode()`) where the CPU would take minutes (full-size 50-step chains).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------- structure
def unet_spec(image_size=128, in_channels=18, base_channels=128, channels_cfg=(1, 2, 2, 4, 4), resblocks_per_downsample=2,
              attention_res=(32, 16, 8), num_heads=4, embedding_channels=-1):
    """Walks denoising.py:106-187 and returns the block list: dicts with type in
    {'conv_in','res','attn','down','up'} grouped per EmbedSequential, plus state-dict key prefixes."""
    emb_ch = base_channels * 4 if embedding_channels == -1 else embedding_channels
    attention_scale = [image_size // int(r) for r in attention_res]
    spec = dict(emb_ch=emb_ch, base=base_channels, in_channels=in_channels, num_heads=num_heads, in_blocks=[], mid=[], out_blocks=[])
    scale = 1
    spec['in_blocks'].append([dict(type='conv_in', key='in_blocks.0.0', cin=in_channels, cout=base_channels)])
    in_list = [base_channels]
    cin = base_channels
    for level, factor in enumerate(channels_cfg):
        cin = base_channels if level == 0 else base_channels * channels_cfg[level - 1]
        cout = base_channels * factor
        for _ in range(resblocks_per_downsample):
            i = len(spec['in_blocks'])
            layers = [dict(type='res', key=f'in_blocks.{i}.0', cin=cin, cout=cout)]
            cin = cout
            if scale in attention_scale:
                layers.append(dict(type='attn', key=f'in_blocks.{i}.1', c=cin))
            in_list.append(cin)
            spec['in_blocks'].append(layers)
        if level != len(channels_cfg) - 1:
            i = len(spec['in_blocks'])
            spec['in_blocks'].append([dict(type='down', key=f'in_blocks.{i}.0', c=cin)])
            in_list.append(cin)
            scale *= 2
    spec['mid'] = [dict(type='res', key='mid_blocks.0', cin=cin, cout=cin), dict(type='attn', key='mid_blocks.1', c=cin),
                   dict(type='res', key='mid_blocks.2', cin=cin, cout=cin)]
    for level, factor in enumerate(channels_cfg[::-1]):
        for idx in range(resblocks_per_downsample + 1):
            j = len(spec['out_blocks'])
            skip = in_list.pop()
            layers = [dict(type='res', key=f'out_blocks.{j}.0', cin=cin + skip, cout=base_channels * factor, skip=skip)]
            cin = base_channels * factor
            if scale in attention_scale:
                layers.append(dict(type='attn', key=f'out_blocks.{j}.{len(layers)}', c=cin))
            if level != len(channels_cfg) - 1 and idx == resblocks_per_downsample:
                layers.append(dict(type='up', key=f'out_blocks.{j}.{len(layers)}', c=cin))
                scale //= 2
            spec['out_blocks'].append(layers)
    spec['out_c'] = cin
    return spec


def random_state_dict(spec, seed=0, std=0.02, nonzero_last=True):
    """Random weights under the reference's state-dict keys. The reference zero-initialises
    every conv_2 and attention proj (UNet ~ identity); `nonzero_last` draws them N(0, std) so every path is exercised."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def conv(key, cout, cin, k):
        sd[key + '.weight'] = torch.randn(cout, cin, k, k, generator=g) * std
        sd[key + '.bias'] = torch.randn(cout, generator=g) * std

    def lin(key, cout, cin):
        sd[key + '.weight'] = torch.randn(cout, cin, generator=g) * std * 2
        sd[key + '.bias'] = torch.randn(cout, generator=g) * std

    def gn(key, c):
        sd[key + '.weight'] = 1 + 0.1 * torch.randn(c, generator=g)
        sd[key + '.bias'] = 0.1 * torch.randn(c, generator=g)

    emb = spec['emb_ch']
    lin('time_embedding.blocks.0', emb, spec['base'])
    lin('time_embedding.blocks.2', emb, emb)

    def block(b):
        k = b['key']
        if b['type'] == 'conv_in':
            conv(k, b['cout'], b['cin'], 3)
        elif b['type'] == 'res':
            gn(k + '.conv_1.0', b['cin'])
            conv(k + '.conv_1.2', b['cout'], b['cin'], 3)
            gn(k + '.norm_with_embedding.norm', b['cout'])
            lin(k + '.norm_with_embedding.embedding_layer.1', 2 * b['cout'], emb)
            conv(k + '.conv_2.1', b['cout'], b['cout'], 3)
            if b['cin'] != b['cout']:
                conv(k + '.shortcut', b['cout'], b['cin'], 1)
        elif b['type'] == 'attn':
            gn(k + '.norm', b['c'])
            sd[k + '.qkv.weight'] = torch.randn(3 * b['c'], b['c'], 1, generator=g) * std * 2
            sd[k + '.qkv.bias'] = torch.randn(3 * b['c'], generator=g) * std
            sd[k + '.proj.weight'] = torch.randn(b['c'], b['c'], 1, generator=g) * std * 2
            sd[k + '.proj.bias'] = torch.randn(b['c'], generator=g) * std
        elif b['type'] == 'down':
            conv(k + '.downsample', b['c'], b['c'], 3)
        elif b['type'] == 'up':
            conv(k + '.conv', b['c'], b['c'], 3)

    for layers in spec['in_blocks']:
        for b in layers:
            block(b)
    for b in spec['mid']:
        block(b)
    for layers in spec['out_blocks']:
        for b in layers:
            block(b)
    gn('out.gn', spec['out_c'])
    conv('out.conv', spec['in_channels'], spec['out_c'], 3)
    return sd


# ----------------------------------------------------------------------------- forward
def time_embedding(sd, t, base):
    """[mmgen-memory] TimeEmbedding: sinusoidal (cos | sin) -> Linear -> SiLU -> Linear."""
    half = base // 2
    freqs = torch.exp(-math.log(10000) * torch.arange(half, dtype=torch.float32) / half).to(t.device)
    args = t[:, None].float() * freqs[None]
    e = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    e = F.linear(e, sd['time_embedding.blocks.0.weight'], sd['time_embedding.blocks.0.bias'])
    return F.linear(F.silu(e), sd['time_embedding.blocks.2.weight'], sd['time_embedding.blocks.2.bias'])


def _gn(sd, key, x):
    return F.group_norm(x, 32, sd[key + '.weight'], sd[key + '.bias'], eps=1e-5)


def res_block(sd, b, x, emb):
    """[mmgen-memory] DenoisingResBlock.forward with NormWithEmbedding(use_scale_shift=True)."""
    k = b['key']
    sc = F.conv2d(x, sd[k + '.shortcut.weight'], sd[k + '.shortcut.bias']) if b['cin'] != b['cout'] else x
    h = F.conv2d(F.silu(_gn(sd, k + '.conv_1.0', x)), sd[k + '.conv_1.2.weight'], sd[k + '.conv_1.2.bias'], padding=1)
    e = F.linear(F.silu(emb), sd[k + '.norm_with_embedding.embedding_layer.1.weight'],
                 sd[k + '.norm_with_embedding.embedding_layer.1.bias'])[:, :, None, None]
    scale, shift = torch.chunk(e, 2, dim=1)
    h = _gn(sd, k + '.norm_with_embedding.norm', h) * (1 + scale) + shift
    h = F.conv2d(F.silu(h), sd[k + '.conv_2.1.weight'], sd[k + '.conv_2.1.bias'], padding=1)
    return h + sc


def attention(sd, b, x, num_heads):
    """lib/models/architecture/ddpm/modules.py:28-48 (groups=1) + [mmgen-memory] QKVAttention."""
    k = b['key']
    bsz, c, *spatial = x.shape
    xf = x.reshape(bsz, c, -1)
    T = xf.size(-1)
    qkv = F.conv1d(F.group_norm(xf, 32, sd[k + '.norm.weight'], sd[k + '.norm.bias'], eps=1e-5), sd[k + '.qkv.weight'], sd[k + '.qkv.bias'])
    qkv = qkv.reshape(bsz * num_heads, -1, T)
    ch = qkv.shape[1] // 3
    q, kk, v = torch.chunk(qkv, 3, dim=1)
    s = 1 / math.sqrt(math.sqrt(ch))
    w = torch.einsum('bct,bcs->bts', q * s, kk * s)
    w = torch.softmax(w.float(), dim=-1).type(w.dtype)
    h = torch.einsum('bts,bcs->bct', w, v).reshape(bsz, -1, T)
    h = F.conv1d(h, sd[k + '.proj.weight'], sd[k + '.proj.bias'])
    return (h + xf).reshape(bsz, c, *spatial)


def _run_layers(sd, spec, layers, h, emb):
    for b in layers:
        if b['type'] == 'conv_in':
            h = F.conv2d(h, sd[b['key'] + '.weight'], sd[b['key'] + '.bias'], padding=1)
        elif b['type'] == 'res':
            h = res_block(sd, b, h, emb)
        elif b['type'] == 'attn':
            h = attention(sd, b, h, spec['num_heads'])
        elif b['type'] == 'down':
            h = F.conv2d(h, sd[b['key'] + '.downsample.weight'], sd[b['key'] + '.downsample.bias'], stride=2, padding=1)
        elif b['type'] == 'up':
            h = F.conv2d(F.interpolate(h, scale_factor=2, mode='nearest'), sd[b['key'] + '.conv.weight'], sd[b['key'] + '.conv.bias'], padding=1)
    return h


def unet_forward(sd, spec, x_t, t, num_timesteps=1000):
    """denoising.py:191-216 (use_rescale_timesteps=True => t * 1000 / num_timesteps)."""
    t = t.float() * (1000.0 / num_timesteps)
    emb = time_embedding(sd, t, spec['base'])
    h, hs = x_t, []
    for layers in spec['in_blocks']:
        h = _run_layers(sd, spec, layers, h, emb)
        hs.append(h)
    h = _run_layers(sd, spec, spec['mid'], h, emb)
    for layers in spec['out_blocks']:
        h = _run_layers(sd, spec, layers, torch.cat([h, hs.pop()], dim=1), emb)
    h = F.silu(_gn(sd, 'out.gn', h))
    return F.conv2d(h, sd['out.conv.weight'], sd['out.conv.bias'], padding=1)


# ----------------------------------------------------------------------------- diffusion
def linear_betas(T=1000, beta_0=1e-4, beta_T=2e-2):
    """gaussian_diffusion.py:64-82."""
    scale = 1000 / T
    return np.linspace(scale * beta_0, scale * beta_T, T, dtype=np.float64)


def diffusion_vars(betas):
    """gaussian_diffusion.py:131-154 (np.cumproduct -> np.cumprod)."""
    alphas = 1.0 - betas
    ab = np.cumprod(alphas, axis=0)
    ab_prev = np.append(1.0, ab[:-1])
    return dict(betas=betas, alphas_bar=ab, alphas_bar_prev=ab_prev, sqrt_alphas_bar=np.sqrt(ab),
                sqrt_one_minus_alphas_bar=np.sqrt(1.0 - ab), tilde_betas_t=betas * (1 - ab_prev) / (1 - ab))


def ddim_timesteps(T=1000, num=50):
    """gaussian_diffusion.py:302-304."""
    return torch.arange(start=T - 1, end=-1, step=-(T / num)).long()


def ddim_sample(denoise_fn, noise, dv, num_timesteps=50, T=1000, clip_range=(-2, 2), eta=0.0, clip_denoised=True):
    """gaussian_diffusion.py:295-331 + :264-293 + :180-240, V-parameterisation, no guidance.
    denoise_fn(x_t, t[B] long) -> v."""
    x_t = noise
    ts = ddim_timesteps(T, num_timesteps)
    for step, t in enumerate(ts):
        t_prev = ts[step + 1] if step + 1 < len(ts) else -1
        ab_prev = dv['alphas_bar'][t_prev] if t_prev >= 0 else dv['alphas_bar_prev'][0]
        tilde_beta = dv['tilde_betas_t'][t]
        sa = x_t.new_tensor(dv['sqrt_alphas_bar'])[t].reshape(-1, 1, 1, 1)
        s1 = x_t.new_tensor(dv['sqrt_one_minus_alphas_bar'])[t].reshape(-1, 1, 1, 1)
        v = denoise_fn(x_t, t.expand(x_t.size(0)))
        x0 = sa * x_t - s1 * v
        if clip_denoised:
            x0 = x0.clamp(*clip_range)
        eps = (x_t - dv['sqrt_alphas_bar'][t] * x0) / dv['sqrt_one_minus_alphas_bar'][t]
        x_t = np.sqrt(ab_prev) * x0 + np.sqrt(1 - ab_prev - tilde_beta * eta ** 2) * eps
    return x_t


def fp32_reference_mode():
    """strict fp32 for the oracle when it runs on a CUDA device (the reference default would allow TF32 convs)"""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = False


def state_dict_to(sd, device):
    return {k: v.to(device) for k, v in sd.items()}


def pred_x_0(denoise_fn, x_t, t, dv, grad_guide_fn=None, clip_denoised=True, clip_range=(-1, 1), guidance_gain=1.0,
             grad_through_unet=True, snr_weight_power=0.5, update_denoising_output=False):
    """gaussian_diffusion.py:180-240, V parameterisation. t: int; denoise_fn(x, t[B]) -> v."""
    B = x_t.size(0)
    tt = torch.full((B,), int(t), dtype=torch.long, device=x_t.device)
    sa = float(np.float32(dv['sqrt_alphas_bar'][int(t)]))        # the reference builds the table in x_t's dtype (:190-191)
    s1 = float(np.float32(dv['sqrt_one_minus_alphas_bar'][int(t)]))
    if grad_guide_fn is None:
        with torch.no_grad():
            v = denoise_fn(x_t, tt)
            x0 = sa * x_t - s1 * v
            if clip_denoised:
                x0 = x0.clamp(*clip_range)
        return x0, v
    with torch.enable_grad():
        if grad_through_unet:
            x_in = x_t.detach().requires_grad_(True)
            v = denoise_fn(x_in, tt)
            x0 = sa * x_in - s1 * v
            if clip_denoised:
                x0 = x0.clamp(*clip_range)
            grad = torch.autograd.grad(grad_guide_fn(x0), x_in)[0]
        else:
            with torch.no_grad():
                v = denoise_fn(x_t, tt)
                x0 = sa * x_t - s1 * v
                if clip_denoised:
                    x0 = x0.clamp(*clip_range)
            x0 = x0.detach().requires_grad_(True)
            grad = torch.autograd.grad(grad_guide_fn(x0), x0)[0]
    x0 = x0.detach() - grad * ((s1 ** (2 - snr_weight_power * 2)) * (sa ** (snr_weight_power * 2 - 1)) * guidance_gain)
    if clip_denoised:
        x0 = x0.clamp(*clip_range)
    v = v.detach()
    if update_denoising_output:
        v = (sa * x_t - x0) / s1
    return x0, v


def ddim_sample_guided(denoise_fn, noise, dv, cfg, T=1000, grad_guide_fn=None, langevin_noises=None):
    """gaussian_diffusion.py:295-331 with guidance and langevin correction steps (:242-262); `cfg` = test_cfg dict;
    `langevin_noises` = iterator of the tensors the reference would draw with `_get_noise_batch`."""
    x_t = noise
    ts = [int(t) for t in ddim_timesteps(T, cfg.get('num_timesteps', T))]
    kw = dict(clip_denoised=cfg.get('clip_denoised', True), clip_range=cfg.get('clip_range', [-1, 1]), guidance_gain=cfg.get('guidance_gain', 1.0),
              grad_through_unet=cfg.get('grad_through_unet', True), snr_weight_power=cfg.get('snr_weight_power', 0.5))
    eta = cfg.get('eta', 0)
    lsteps, lrange, ldelta = cfg.get('langevin_steps', 0), cfg.get('langevin_t_range', [0, 1000]), cfg.get('langevin_delta', 0.1)
    for step, t in enumerate(ts):
        t_prev = ts[step + 1] if step + 1 < len(ts) else -1
        ab_prev = dv['alphas_bar'][t_prev] if t_prev >= 0 else dv['alphas_bar_prev'][0]
        x0, _ = pred_x_0(denoise_fn, x_t, t, dv, grad_guide_fn=grad_guide_fn, **kw)
        eps = (x_t - dv['sqrt_alphas_bar'][t] * x0) / dv['sqrt_one_minus_alphas_bar'][t]
        x_t = np.sqrt(ab_prev) * x0 + np.sqrt(1 - ab_prev - dv['tilde_betas_t'][t] * eta ** 2) * eps
        if lsteps > 0 and lrange[0] < t_prev < lrange[1]:
            for _ in range(lsteps):
                sigma = dv['sqrt_one_minus_alphas_bar'][t_prev]
                x0, _ = pred_x_0(denoise_fn, x_t, t_prev, dv, grad_guide_fn=grad_guide_fn, **kw)
                eps = (x_t - dv['sqrt_alphas_bar'][t_prev] * x0) / sigma
                x_t = x_t - 0.5 * ldelta * sigma * eps + math.sqrt(ldelta) * sigma * next(langevin_noises)
    return x_t


def snr_weighted_loss_weight(dv, power, mode='V', min=-1, max=-1, bias=0, prob_power=0.0):
    """lib/models/diffusions/sampler.py:15-46 -> float32 weight[T]"""
    mean, std = dv['sqrt_alphas_bar'], dv['sqrt_one_minus_alphas_bar']
    weight_x = (mean / std) ** (2 * power) + bias
    if min > 0:
        weight_x = weight_x.clip(min=min)
    if max > 0:
        weight_x = weight_x.clip(max=max)
    raw = dict(EPS=weight_x * (std / mean) ** 2, START_X=weight_x, V=weight_x * std ** 2)[mode]
    prob = raw ** prob_power
    prob = prob / prob.sum()
    return torch.from_numpy(raw / (prob * len(mean))).to(torch.float)
