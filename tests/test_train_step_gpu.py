"""Stage-1 auto-decoder training step (`MultiSceneNeRF.train_step`, multiscene_nerf.py:185-252) built from the reference's own
config (`stage1_cars_recons16v`, resolved fixture): per-scene latents with private Adam state in the scene cache, shared decoder
trained through the fused differentiable renderer's weight gradients."""
import json
import os

import pytest
import torch

from tests.common import GOLDEN, spiral_poses

pytestmark = pytest.mark.gpu


def _stage1(cuda, name='configs/paper_cfgs/stage1_cars_recons16v.py', **train_over):
    import ssdnerf_b200 as S
    c = json.load(open(os.path.join(GOLDEN, 'reference_configs.json')))[name]
    assert c['model']['type'] == 'MultiSceneNeRF' and c['model']['reg_loss']['type'] == 'TVLoss'
    train_cfg = {k: v for k, v in c['train_cfg'].items() if k != 'cache_load_from'}
    train_cfg.update(train_over)
    torch.manual_seed(0)
    model = S.build_model(dict(c['model'], cache_size=6), train_cfg=train_cfg, test_cfg=c['test_cfg']).to(cuda)
    return model, c


def _scenes(model, cuda, B, V, res, seed):
    """synthetic multi-view targets: renders of random triplanes with the model's initial decoder"""
    g = torch.Generator().manual_seed(seed)
    poses = torch.from_numpy(spiral_poses(V))[None].repeat(B, 1, 1, 1).to(cuda)
    f = 131.25 * res / 128
    intr = torch.tensor([f, f, res / 2, res / 2]).expand(B, V, 4).contiguous().to(cuda)
    code = (torch.randn(B, 3, 6, 128, 128, generator=g) * 0.5).to(cuda)
    with torch.no_grad():
        _, bits = model.get_density(model.decoder, code, cfg=dict(density_thresh=0.1))
        imgs, _ = model.render(model.decoder, code, bits, res, res, intr, poses, cfg=dict(dt_gamma_scale=0.5))
    return imgs.clamp(0, 1), poses, intr


def test_stage1_train_step(cuda):
    model, c = _stage1(cuda, extra_scene_step=2, n_decoder_rays=1024, n_inverse_rays=1024)
    model.train()
    B, V, res = 2, 4, 64
    imgs, poses, intr = _scenes(model, cuda, B, V, res, 0)
    # perturb the decoder so that there is something to learn back
    with torch.no_grad():
        for p in model.decoder.parameters():
            p.add_(torch.randn_like(p) * 0.05)
    w0 = {k: v.detach().clone() for k, v in model.decoder.named_parameters()}
    opt = dict(decoder=torch.optim.Adam(model.decoder.parameters(), lr=1e-3))
    data = dict(scene_id=[1, 4], scene_name=['s1', 's4'], cond_imgs=imgs, cond_poses=poses, cond_intrinsics=intr)
    losses = []
    for it in range(6):
        out = model.train_step(data, opt)
        lv = out['log_vars']
        assert out['num_samples'] == B and set(lv) >= {'pixel_loss', 'reg_loss', 'loss', 'train_psnr', 'code_rms'}
        assert all(map(lambda v: v == v and abs(v) < 1e9, lv.values())), lv
        losses.append(lv['loss'])
    print('stage-1 losses', ['%.4f' % v for v in losses])
    assert losses[-1] < losses[0]
    # the decoder moved, every parameter received a gradient
    for k, p in model.decoder.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all() and float(p.grad.abs().max()) > 0, k
        assert not torch.equal(p.detach(), w0[k]), k
    # scene cache: only the visited scenes are filled; latent + occupancy state + private Adam state with the right step count
    assert sorted(k for k, v in model.cache.items() if v is not None) == [1, 4] and model.cache_loaded
    e = model.cache[4]
    assert e['scene_name'] == 's4' and e['param']['code_'].shape == (3, 6, 128, 128) and e['param']['code_'].dtype == torch.float32
    assert e['param']['density_bitfield'].dtype == torch.uint8 and int(e['param']['density_bitfield'].count_nonzero()) > 0
    st = e['optimizer']['state'][0]
    assert int(st['step']) == 6 * (2 + 1)                     # (extra_scene_step + joint step) per visit
    assert st['exp_avg'].shape == (3, 6, 128, 128) and float(st['exp_avg_sq'].sum()) > 0
    # a new scene starts from the running mean code (init_from_mean), not from the visited scenes' state
    assert float(model.init_code.abs().max()) > 0


def test_stage1_16bit_cache_and_scene_files(cuda, tmp_path):
    """`stage1_cars_recons16v_16bit_filesystem`: fp16 latents / bf16 Adam moments in the cache, write-through scene files readable as
    `cache_load_from` of a fresh model"""
    import ssdnerf_b200 as S
    name = 'configs/new_cfgs/stage1_cars_recons16v_16bit_filesystem.py'
    model, c = _stage1(cuda, name=name, extra_scene_step=0, n_decoder_rays=512, save_dir=str(tmp_path))
    assert c['model']['cache_16bit'] and c['model']['num_file_writers'] > 0
    model.train()
    B, V, res = 2, 2, 32
    imgs, poses, intr = _scenes(model, cuda, B, V, res, 1)
    opt = dict(decoder=torch.optim.Adam(model.decoder.parameters(), lr=1e-3))
    data = dict(scene_id=[0, 1], scene_name=['a', 'b'], cond_imgs=imgs, cond_poses=poses, cond_intrinsics=intr)
    for _ in range(2):
        model.train_step(data, opt)
    model.scene_cache.flush()
    e = model.cache[0]
    assert e['param']['code_'].dtype == torch.float16 and e['optimizer']['state'][0]['exp_avg'].dtype == torch.bfloat16
    assert e['param']['density_grid'].dtype == torch.float16 and int(e['optimizer']['state'][0]['step']) == 2
    files = sorted(os.listdir(tmp_path))
    assert files == ['a.pth', 'b.pth']
    rec = torch.load(os.path.join(tmp_path, 'a.pth'), map_location='cpu')
    assert set(rec) >= {'scene_id', 'scene_name', 'param', 'optimizer'} and set(rec['param']) == {'code_', 'density_grid', 'density_bitfield'}
    # resume: a fresh model loads the directory and continues with the stored latent and Adam state
    train_cfg = dict(model.train_cfg, cache_load_from=str(tmp_path))
    train_cfg.pop('save_dir')
    m2 = S.build_model(dict(c['model'], cache_size=2), train_cfg=train_cfg, test_cfg=c['test_cfg']).to(cuda).train()
    m2.decoder.load_state_dict(model.decoder.state_dict())
    codes, opts, grid, bits = m2.load_cache(data)
    torch.cuda.synchronize()
    assert torch.equal(codes[0].detach().cpu(), rec['param']['code_'].float()) and codes[0].requires_grad
    assert int(opts[0].state_dict()['state'][0]['step']) == 2 and opts[0].state_dict()['state'][0]['exp_avg'].dtype == torch.float32
    assert torch.equal(bits[0].cpu(), rec['param']['density_bitfield'])
    out = m2.train_step(data, dict(decoder=torch.optim.Adam(m2.decoder.parameters(), lr=1e-3)))
    assert out['log_vars']['loss'] == out['log_vars']['loss']


def _diffusion_model(cuda, name, cache_size=0, **train_over):
    import ssdnerf_b200 as S
    c = json.load(open(os.path.join(GOLDEN, 'reference_configs.json')))[name]
    train_cfg = {k: v for k, v in c['train_cfg'].items() if k != 'cache_load_from'}
    train_cfg.update(train_over)
    torch.manual_seed(0)
    model = S.build_model(dict(c['model'], cache_size=cache_size), train_cfg=train_cfg, test_cfg=c['test_cfg'])
    g = torch.Generator().manual_seed(0)
    for p in model.diffusion.denoising.parameters():          # mmgen zero-initialises the last conv of every block: give them values
        if p.dim() > 1:
            p.data.copy_(torch.randn(p.shape, generator=g) * 0.02)
    return model.to(cuda).train(), c


def test_stage2_train_step_trains_the_denoiser(cuda):
    """`stage2_cars_uncond` (reference config): stored scenes, only the diffusion optimizer; full-size UNet, 2 scenes"""
    model, c = _diffusion_model(cuda, 'configs/paper_cfgs/stage2_cars_uncond.py')
    assert 'optimizer' not in model.train_cfg and model.freeze_decoder
    g = torch.Generator().manual_seed(1)
    stored = [dict(param=dict(code=torch.tanh(torch.randn(3, 6, 128, 128, generator=g)) * 0.8, density_grid=torch.zeros(64 ** 3).half(),
                              density_bitfield=torch.zeros(64 ** 3 // 8, dtype=torch.uint8))) for _ in range(2)]
    data = dict(scene_id=[0, 1], scene_name=['a', 'b'], code=stored)
    unet = model.diffusion.denoising
    w0 = {k: v.detach().clone() for k, v in unet.named_parameters()}
    opt = dict(diffusion=torch.optim.Adam(model.diffusion.parameters(), lr=1e-4))
    losses = []
    for _ in range(3):
        out = model.train_step(data, opt)
        assert out['num_samples'] == 2 and 'loss_ddpm_mse' in out['log_vars']
        losses.append(out['log_vars']['loss_ddpm_mse'])
    assert all(v == v and v < 1e6 for v in losses), losses
    missing = [k for k, p in unet.named_parameters() if p.grad is None or not torch.isfinite(p.grad).all() or float(p.grad.abs().max()) == 0]
    assert not missing, missing[:8]
    assert all(not torch.equal(p.detach(), w0[k]) for k, p in unet.named_parameters())
    assert all(p.grad is None for p in model.decoder.parameters())
    # the EMA copy is not touched by train_step (the reference updates it from a runner hook)
    assert model.diffusion_ema is not model.diffusion


def test_single_stage_train_step(cuda):
    """`ssdnerf_cars_uncond` (reference config): latents (scene cache) + decoder + denoiser in one iteration; fewer inner steps / rays"""
    model, c = _diffusion_model(cuda, 'configs/paper_cfgs/ssdnerf_cars_uncond.py', cache_size=4, extra_scene_step=2, n_decoder_rays=1024,
                                n_inverse_rays=1024)
    assert not model.freeze_decoder and 'optimizer' in model.train_cfg
    B, V, res = 2, 3, 64
    imgs, poses, intr = _scenes(model, cuda, B, V, res, 3)
    opt = dict(diffusion=torch.optim.Adam(model.diffusion.parameters(), lr=1e-4), decoder=torch.optim.Adam(model.decoder.parameters(), lr=1e-3))
    data = dict(scene_id=[0, 3], scene_name=['s0', 's3'], cond_imgs=imgs, cond_poses=poses, cond_intrinsics=intr)
    dec0 = {k: v.detach().clone() for k, v in model.decoder.named_parameters()}
    unet0 = model.diffusion.denoising.in_blocks[0][0].weight.detach().clone()
    for _ in range(2):
        out = model.train_step(data, opt)
    lv = out['log_vars']
    assert set(lv) >= {'loss_ddpm_mse', 'pixel_loss', 'loss_decoder', 'train_psnr', 'code_rms'} and all(v == v for v in lv.values()), lv
    assert all(not torch.equal(p.detach(), dec0[k]) for k, p in model.decoder.named_parameters())
    assert not torch.equal(model.diffusion.denoising.in_blocks[0][0].weight.detach(), unet0)
    e = model.cache[3]
    assert e is not None and int(e['optimizer']['state'][0]['step']) == 2 * (2 + 1)
    # the diffusion prior reached the latents: with x_t_detach unset the code receives a gradient from the diffusion loss
    assert float(e['optimizer']['state'][0]['exp_avg'].abs().sum()) > 0
