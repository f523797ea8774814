"""Pins for the stage-1 trainer's scene cache: the reference's OWN `MultiSceneNeRF.load_cache` / `save_cache` (+ `out_dict_to`,
`optimizer_state_to`, `load_tensor_to_dict`, `optimizer_state_copy`, `optimizer_set_state` of lib/core/utils/misc.py) executed from
/root/reference on CPU with mmcv / mmgen stubbed, driven through a fixed visit sequence; the tensors they leave in the cache and hand back
are stored in tests/golden/reference_cache_v1.npz.  `tests/test_scene_cache_cpu.py::test_cache_matches_reference_execution` replays the
sequence on `ssdnerf_b200.MultiSceneNeRF`.

    python tests/golden/make_golden_cache.py          (needs /root/reference; the GPU box only reads the committed .npz)

Sequence (2 scenes, latent 3x2x4x4, grid 8^3, cache_16bit=True, Adam lr 0.01 then 0.02):
  visit 1: fresh scenes -> 2 Adam steps on seeded gradients -> save_cache          (first-time path: out_dict_to)
  visit 2: load_cache (lr changed to 0.02 in train_cfg) -> 1 Adam step -> save_cache (in-place path: load_tensor_to_dict / optimizer_state_copy)
  visit 3: load_cache only
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tests.golden import make_golden_ref as G  # noqa: E402

CODE_SIZE, GRID = (3, 2, 4, 4), 8


def seeded(shape, seed, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale


def visit_inputs():
    """everything random the sequence consumes, by name"""
    d = {}
    for s in range(2):
        init = seeded(CODE_SIZE, 10 + s)
        init.view(-1)[0] = 1e6 if s == 0 else -1e6          # beyond the fp16 range: exercises the saturating cast
        d[f'init_{s}'] = init
        for k in range(3):
            d[f'grad_{s}_{k}'] = seeded(CODE_SIZE, 100 + 10 * s + k, 0.5)
        d[f'grid_{s}'] = seeded((GRID ** 3,), 200 + s).abs().half()
        d[f'bits_{s}'] = (seeded((GRID ** 3 // 8,), 300 + s) * 64).abs().clamp(max=255).to(torch.uint8)
    return d


def load_reference_cache_code():
    G._install_stubs()
    nu, act, reg, tv, base = G.load_reference_host_helpers()

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m
    mod('mmcv.parallel', MMDistributedDataParallel=type('MMDistributedDataParallel', (), {}))
    misc = G._load('lib/core/utils/misc.py', 'ref_misc')
    core = sys.modules['reflib.core']
    for n in ('optimizer_state_to', 'load_tensor_to_dict', 'optimizer_state_copy', 'optimizer_set_state'):
        setattr(core, n, getattr(misc, n))
    sys.modules['mmcv.runner'].get_dist_info = lambda: (0, 1)
    sys.modules['mmcv'].print_log = lambda *a, **k: None
    sys.modules['mmgen.models.builder'].MODELS = G.MODULES
    msn = G._load('lib/models/autodecoders/multiscene_nerf.py', 'reflib.models.autodecoders.multiscene_nerf')
    return msn, base


def run_reference():
    msn, base = load_reference_cache_code()
    inp = visit_inputs()
    cls = msn.MultiSceneNeRF
    fresh = iter(())

    fake = types.SimpleNamespace(
        cache={0: None, 1: None}, cache_loaded=False, cache_16bit=True, num_file_writers=0, is_file_writers_initialized=False,
        train_cfg=dict(optimizer=dict(type='Adam', lr=0.01, weight_decay=0.0)),
        get_init_density_grid=lambda n, device=None: torch.zeros(GRID ** 3, dtype=torch.float16),
        get_init_density_bitfield=lambda n, device=None: torch.zeros(GRID ** 3 // 8, dtype=torch.uint8),
        build_optimizer=base.BaseNeRF.build_optimizer, code_activation=None, parameters=lambda: iter([torch.zeros(1)]))
    fake.get_init_code_ = lambda n, device=None: next(fresh)
    scene_id, names = [1, 0], ['b', 'a']
    out = {k: v.float().numpy() if v.is_floating_point() else v.numpy() for k, v in inp.items()}

    def record(tag):
        for sid in (0, 1):
            e = fake.cache[sid]
            out[f'{tag}_s{sid}_code'] = e['param']['code_'].float().numpy().copy()
            out[f'{tag}_s{sid}_code_dtype'] = np.array(str(e['param']['code_'].dtype))
            out[f'{tag}_s{sid}_grid'] = e['param']['density_grid'].float().numpy().copy()
            out[f'{tag}_s{sid}_grid_dtype'] = np.array(str(e['param']['density_grid'].dtype))
            out[f'{tag}_s{sid}_bits'] = e['param']['density_bitfield'].numpy().copy()
            st = e['optimizer']['state'][0]
            out[f'{tag}_s{sid}_step'] = np.array(float(st['step']))
            out[f'{tag}_s{sid}_exp_avg'] = st['exp_avg'].float().numpy().copy()
            out[f'{tag}_s{sid}_exp_avg_sq'] = st['exp_avg_sq'].float().numpy().copy()
            out[f'{tag}_s{sid}_moment_dtype'] = np.array(str(st['exp_avg'].dtype))
            out[f'{tag}_s{sid}_keys'] = np.array(sorted(e.keys()))
            out[f'{tag}_s{sid}_name'] = np.array(e['scene_name'])

    def visit(tag, n_steps, k0, save=True):
        nonlocal fresh
        fresh = iter([inp[f'init_{s}'].clone().requires_grad_(True) for s in scene_id])
        codes, opts, grid, bits = cls.load_cache(fake, dict(scene_id=scene_id, scene_name=names))
        for i, s in enumerate(scene_id):
            out[f'{tag}_loaded_s{s}_code'] = codes[i].detach().numpy().copy()
            out[f'{tag}_loaded_s{s}_lr'] = np.array(opts[i].param_groups[0]['lr'])
            sd = opts[i].state_dict()['state']
            out[f'{tag}_loaded_s{s}_has_state'] = np.array(len(sd) > 0)
            if len(sd) > 0:
                out[f'{tag}_loaded_s{s}_exp_avg'] = sd[0]['exp_avg'].float().numpy().copy()
                out[f'{tag}_loaded_s{s}_state_dtype'] = np.array(str(sd[0]['exp_avg'].dtype))
                out[f'{tag}_loaded_s{s}_step'] = np.array(float(sd[0]['step']))
        out[f'{tag}_loaded_grid'] = grid.float().numpy().copy()
        out[f'{tag}_loaded_bits'] = bits.numpy().copy()
        for k in range(n_steps):
            for i, s in enumerate(scene_id):
                codes[i].grad = inp[f'grad_{s}_{k0 + k}'].clone()
                opts[i].step()
        if save:
            grid = torch.stack([inp[f'grid_{s}'] for s in scene_id]) * (1 if tag == 'v1' else 2)
            bits = torch.stack([inp[f'bits_{s}'] for s in scene_id])
            cls.save_cache(fake, codes, opts, grid, bits, scene_id, names)
            record(tag)

    visit('v1', 2, 0)
    fake.train_cfg = dict(optimizer=dict(type='Adam', lr=0.02, weight_decay=0.0))
    visit('v2', 1, 2)
    visit('v3', 0, 0, save=False)
    return out


if __name__ == '__main__':
    res = run_reference()
    np.savez_compressed(os.path.join(HERE, 'reference_cache_v1.npz'), **res)
    print('wrote reference_cache_v1.npz:', len(res), 'arrays,', os.path.getsize(os.path.join(HERE, 'reference_cache_v1.npz')), 'bytes')
