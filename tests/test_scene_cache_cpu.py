"""Scene cache of the stage-1 trainer (multiscene_nerf.py:19-183): host logic, no GPU needed."""
import json
import os

import numpy as np
import torch

from tests.common import GOLDEN


def _model(cache_size=5, **over):
    import ssdnerf_b200 as S
    c = json.load(open(os.path.join(GOLDEN, 'reference_configs.json')))['configs/paper_cfgs/stage1_cars_recons16v.py']
    train_cfg = {k: v for k, v in c['train_cfg'].items() if k != 'cache_load_from'}
    train_cfg.update(over.pop('train_cfg', {}))
    torch.manual_seed(0)
    return S.build_model(dict(c['model'], cache_size=cache_size, **over), train_cfg=train_cfg, test_cfg=c['test_cfg'])


def test_rank_shards_follow_the_reference_split():
    from ssdnerf_b200.scene_cache import SceneCache, shard_bounds
    # multiscene_nerf.py:49-50: np.round(np.linspace(0, cache_size, ws + 1)) -- python-side restatement of the contract
    for size, ws in [(2458, 8), (7, 3), (5, 5), (3, 4)]:
        b = shard_bounds(size, ws)
        owned = [set(SceneCache(size, r, ws).entries) for r in range(ws)]
        assert set().union(*owned) == set(range(size)) and sum(map(len, owned)) == size
        for r in range(ws):
            assert owned[r] == set(range(int(b[r]), int(b[r + 1])))
            assert abs(len(owned[r]) - size / ws) <= 1
    assert SceneCache(0).entries is None


def test_load_step_save_roundtrip_and_in_place_reuse():
    m = _model()
    data = dict(scene_id=[3, 0], scene_name=['c', 'a'])
    codes, opts, grid, bits = m.load_cache(data)
    assert len(codes) == 2 and codes[0].shape == (3, 6, 128, 128) and codes[0].requires_grad and codes[0].is_leaf
    assert grid.shape == (2, 64 ** 3) and grid.dtype == torch.float16 and bits.shape == (2, 64 ** 3 // 8) and bits.dtype == torch.uint8
    assert isinstance(opts[0], torch.optim.Adam) and opts[0].param_groups[0]['lr'] == m.train_cfg['optimizer']['lr']
    for c_, o in zip(codes, opts):
        c_.grad = torch.ones_like(c_)
        o.step()
    grid[0, :10] = 1.0
    bits[1, :4] = 255
    m.save_cache(codes, opts, grid, bits, data['scene_id'], data['scene_name'])
    e3 = m.cache[3]
    assert m.cache[1] is None and e3['scene_id'] == 3 and e3['scene_name'] == 'c'
    assert torch.equal(e3['param']['code_'], codes[0].detach()) and e3['param']['code_'] is not codes[0].data
    assert int(e3['optimizer']['state'][0]['step']) == 1
    keep = e3['param']['code_']
    # second visit: state restored, Adam continues (step 2), cached host tensors overwritten in place
    codes2, opts2, grid2, bits2 = m.load_cache(data)
    assert torch.equal(codes2[0].detach(), codes[0].detach()) and torch.equal(grid2, grid) and torch.equal(bits2, bits)
    sd = opts2[0].state_dict()['state'][0]
    assert int(sd['step']) == 1 and torch.equal(sd['exp_avg'], opts[0].state_dict()['state'][0]['exp_avg'])
    codes2[0].grad = torch.ones_like(codes2[0])
    opts2[0].step()
    m.save_cache(codes2, opts2, grid2, bits2, data['scene_id'], data['scene_name'])
    assert m.cache[3]['param']['code_'] is keep and torch.equal(keep, codes2[0].detach())
    assert int(m.cache[3]['optimizer']['state'][0]['step']) == 2


def test_16bit_storage_saturates_and_files_resume(tmp_path):
    m = _model(cache_size=2, cache_16bit=True, train_cfg=dict(save_dir=str(tmp_path)))
    data = dict(scene_id=[0, 1], scene_name=['a', 'b'])
    codes, opts, grid, bits = m.load_cache(data)
    with torch.no_grad():
        codes[0].view(-1)[0] = 1e9                                   # beyond fp16 range: stored as the largest finite fp16
    for c_, o in zip(codes, opts):
        c_.grad = torch.full_like(c_, 3.0)
        o.step()
    m.save_cache(codes, opts, grid, bits, data['scene_id'], data['scene_name'])
    m.scene_cache.flush()
    e = m.cache[0]
    assert e['param']['code_'].dtype == torch.float16 and torch.isfinite(e['param']['code_']).all()
    assert float(e['param']['code_'].view(-1)[0]) == torch.finfo(torch.float16).max
    assert e['optimizer']['state'][0]['exp_avg'].dtype == torch.bfloat16 and e['param']['density_grid'].dtype == torch.float16
    assert sorted(os.listdir(tmp_path)) == ['a.pth', 'b.pth']
    m2 = _model(cache_size=2, train_cfg=dict(cache_load_from=str(tmp_path)))
    codes2, opts2, _, _ = m2.load_cache(data)
    assert codes2[1].dtype == torch.float32 and torch.equal(codes2[1].detach(), codes[1].detach().half().float())
    assert opts2[1].state_dict()['state'][0]['exp_avg'].dtype == torch.float32 and int(opts2[1].state_dict()['state'][0]['step']) == 1


def test_stored_activated_code_is_inverted():
    """scene records that only hold the ACTIVATED code (`code`, as save_scene writes them) are accepted through the activation's inverse"""
    m = _model(cache_size=0)
    code = torch.tanh(torch.randn(3, 6, 128, 128) * 0.3) * 2 * 0.9
    rec = dict(param=dict(code=code, density_grid=torch.zeros(64 ** 3).half(), density_bitfield=torch.zeros(64 ** 3 // 8, dtype=torch.uint8)))
    codes, opts, _, _ = m.load_cache(dict(scene_id=[7], scene_name=['x'], code=[rec]))
    np.testing.assert_allclose(m.code_activation(codes[0]).detach().numpy(), code.numpy(), atol=1e-5)
    assert m.cache is None


def test_cache_matches_reference_execution():
    """replay of tests/golden/make_golden_cache.py's visit sequence: every tensor the REFERENCE's own load_cache / save_cache left in its
    cache or handed back (fixture produced by executing multiscene_nerf.py + misc.py) vs `ssdnerf_b200.MultiSceneNeRF`"""
    import ssdnerf_b200 as S
    ref = np.load(os.path.join(GOLDEN, 'reference_cache_v1.npz'))
    CODE_SIZE, GRID = (3, 2, 4, 4), 8
    m = S.build_model(dict(type='MultiSceneNeRF', code_size=CODE_SIZE, grid_size=GRID, cache_size=2, cache_16bit=True,
                           decoder=dict(type='TriPlaneDecoder')), train_cfg=dict(optimizer=dict(type='Adam', lr=0.01, weight_decay=0.0)))
    scene_id, names = [1, 0], ['b', 'a']
    fresh = iter(())
    m.get_init_code_ = lambda n, device=None: next(fresh)

    def check(tag):
        for sid in (0, 1):
            e = m.cache[sid]
            assert str(e['param']['code_'].dtype) == str(ref[f'{tag}_s{sid}_code_dtype']) == 'torch.float16'
            assert str(e['param']['density_grid'].dtype) == str(ref[f'{tag}_s{sid}_grid_dtype'])
            assert np.array_equal(e['param']['code_'].float().numpy(), ref[f'{tag}_s{sid}_code'])          # incl. the saturated +-65504
            assert np.array_equal(e['param']['density_grid'].float().numpy(), ref[f'{tag}_s{sid}_grid'])
            assert np.array_equal(e['param']['density_bitfield'].numpy(), ref[f'{tag}_s{sid}_bits'])
            st = e['optimizer']['state'][0]
            assert float(st['step']) == float(ref[f'{tag}_s{sid}_step'])
            assert str(st['exp_avg'].dtype) == str(ref[f'{tag}_s{sid}_moment_dtype']) == 'torch.bfloat16'
            assert np.array_equal(st['exp_avg'].float().numpy(), ref[f'{tag}_s{sid}_exp_avg'])
            assert np.array_equal(st['exp_avg_sq'].float().numpy(), ref[f'{tag}_s{sid}_exp_avg_sq'])
            assert sorted(e.keys()) == list(ref[f'{tag}_s{sid}_keys']) and e['scene_name'] == str(ref[f'{tag}_s{sid}_name'])

    def visit(tag, n_steps, k0, save=True):
        nonlocal fresh
        fresh = iter([torch.from_numpy(ref[f'init_{s}']).clone().requires_grad_(True) for s in scene_id])
        codes, opts, grid, bits = m.load_cache(dict(scene_id=scene_id, scene_name=names))
        for i, s in enumerate(scene_id):
            assert np.array_equal(codes[i].detach().numpy(), ref[f'{tag}_loaded_s{s}_code']), (tag, s)
            assert codes[i].requires_grad and codes[i].dtype == torch.float32
            assert opts[i].param_groups[0]['lr'] == float(ref[f'{tag}_loaded_s{s}_lr'])      # the CURRENT config's lr, not the cached one
            sd = opts[i].state_dict()['state']
            assert (len(sd) > 0) == bool(ref[f'{tag}_loaded_s{s}_has_state'])
            if len(sd) > 0:
                assert str(sd[0]['exp_avg'].dtype) == str(ref[f'{tag}_loaded_s{s}_state_dtype']) == 'torch.float32'
                assert np.array_equal(sd[0]['exp_avg'].numpy(), ref[f'{tag}_loaded_s{s}_exp_avg'])
                assert float(sd[0]['step']) == float(ref[f'{tag}_loaded_s{s}_step'])
        assert np.array_equal(grid.float().numpy(), ref[f'{tag}_loaded_grid']) and np.array_equal(bits.numpy(), ref[f'{tag}_loaded_bits'])
        for k in range(n_steps):
            for i, s in enumerate(scene_id):
                codes[i].grad = torch.from_numpy(ref[f'grad_{s}_{k0 + k}']).clone()
                opts[i].step()
        if save:
            grid = torch.stack([torch.from_numpy(ref[f'grid_{s}']).half() for s in scene_id]) * (1 if tag == 'v1' else 2)
            bits = torch.stack([torch.from_numpy(ref[f'bits_{s}']) for s in scene_id])
            m.save_cache(codes, opts, grid, bits, scene_id, names)
            check(tag)

    visit('v1', 2, 0)
    m.train_cfg = dict(optimizer=dict(type='Adam', lr=0.02, weight_decay=0.0))
    visit('v2', 1, 2)
    visit('v3', 0, 0, save=False)
