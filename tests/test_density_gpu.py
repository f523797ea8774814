"""Occupancy-grid builder vs the oracle's restatement of base_nerf.py:318-401 with injected jitter."""
import numpy as np
import pytest
import torch

from oracle import render_port as rp

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('grid_dtype', [torch.float16, torch.float32])
def test_get_density_matches_oracle(cuda, grid_dtype):
    from ssdnerf_b200 import renderer as R, density as D
    g = torch.Generator().manual_seed(21)
    B = 2
    code = torch.randn(B, 3, 6, 128, 128, generator=g).clamp(-2, 2)
    params = rp.make_decoder_params('P', 4)
    params['density_net.0.bias'] = params['density_net.0.bias'] - 2.5   # mean density ~0.1: threshold branch matters
    rands = [torch.rand(64 ** 3, 3, generator=g) for _ in range(3)]
    grid_ref, bf_ref = rp.get_density(params, code, rands, density_thresh=0.1, grid_dtype=grid_dtype)
    blob = R.pack_decoder_blob(params, R.DEC_P, device=cuda)
    planes = R.pack_planes(code.to(cuda), R.DEC_P)
    grid, bf = D.get_density(R.DEC_P, planes, (128, 128), blob, B, density_thresh=0.1, density_step=3,
                             jitters=[r.to(cuda) for r in rands], grid_dtype=grid_dtype)
    gr, gg = grid_ref.float().numpy(), grid.float().cpu().numpy()
    # sigma = exp(MLP): fp32 round-off, then (for fp16 grids) rounding to half may flip the last bit
    tol = 2e-3 if grid_dtype == torch.float16 else 2e-4
    np.testing.assert_allclose(gg, gr, rtol=tol, atol=1e-6)
    # bits may differ only where the density sits within round-off of the threshold
    diff = np.unpackbits(bf.cpu().numpy() ^ bf_ref, axis=-1).sum()
    assert diff <= 1e-4 * B * 64 ** 3, diff
    assert 0.02 < np.unpackbits(bf_ref).mean() < 0.98


def test_get_density_variant_S(cuda):
    """class-default decoder (3x32 fp16 planes, hidden 128): grid values within fp16 tolerance of the fp32 oracle"""
    from ssdnerf_b200 import renderer as R, density as D
    g = torch.Generator().manual_seed(22)
    code = torch.randn(1, 3, 32, 128, 128, generator=g).clamp(-2, 2)
    params = rp.make_decoder_params('S', 6)
    params['density_net.0.bias'] = params['density_net.0.bias'] - 2.0
    rands = [torch.rand(64 ** 3, 3, generator=g) for _ in range(2)]
    grid_ref, bf_ref = rp.get_density(params, code, rands, density_thresh=0.1, grid_dtype=torch.float32)
    blob = R.pack_decoder_blob(params, R.DEC_S, device=cuda)
    planes = R.pack_planes(code.to(cuda), R.DEC_S)
    grid, bf = D.get_density(R.DEC_S, planes, (128, 128), blob, 1, density_thresh=0.1, density_step=2, jitters=[r.to(cuda) for r in rands],
                             grid_dtype=torch.float32)
    gr, gg = grid_ref.numpy(), grid.cpu().numpy()
    rel = np.abs(gg - gr) / np.maximum(gr, 1e-3)
    assert np.median(rel) < 2e-3 and rel.max() < 3e-2, (np.median(rel), rel.max())
    diff = np.unpackbits(bf.cpu().numpy() ^ bf_ref, axis=-1).mean()
    assert diff < 5e-3, diff
