"""Pin for the occupancy-grid oracle: the reference's OWN `BaseNeRF.update_extra_state` (full-update branch) and `get_density`
(lib/models/autodecoders/base_nerf.py:318-401) + its `TriPlaneDecoder.point_density_decode` + its `morton3D` / `packbits` wrappers,
executed from /root/reference on CPU over the kernel-exact C backend of make_golden_decoder.py.  -> tests/golden/reference_density_v1.npz,
against which tests/test_reference_pin_cpu.py checks `render_port.get_density` / `update_extra_state` (the restatement the GPU density
kernels are measured against).  The jitter is torch's seeded stream on both sides.

    python tests/golden/make_golden_density.py          (needs /root/reference)
"""
import ctypes
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle as orc  # noqa: E402
from oracle import render_port as rp  # noqa: E402
from tests.golden import make_golden_ref as G  # noqa: E402
from tests.golden.make_golden_decoder import CFG, P, load_reference_decoder  # noqa: E402

u32, f32 = ctypes.c_uint32, ctypes.c_float


def run_reference():
    tri = load_reference_decoder()
    L = orc.lib()
    rm = sys.modules['_raymarching']
    rm.morton3D = lambda coords, N, out: L.orc_morton3D(P(coords.contiguous()), u32(N), P(out))
    rm.morton3D_invert = lambda ind, N, out: L.orc_morton3D_invert(P(ind.contiguous()), u32(N), P(out))

    def packbits(grid, N, thresh, bitfield):
        g32 = grid.float().contiguous()                 # the extension dispatches on the grid's dtype and compares in fp32
        L.orc_packbits(P(g32), u32(N), f32(float(thresh)), P(bitfield))
    rm.packbits = packbits
    torch.Tensor.cuda = lambda self, *a, **k: self
    ray = sys.modules['ref_raymarching']
    nu, act, reg, tv, base = G.load_reference_host_helpers()
    base.morton3D, base.morton3D_invert, base.packbits, base.custom_meshgrid = ray.morton3D, ray.morton3D_invert, ray.packbits, nu.custom_meshgrid
    out = {}
    g = torch.Generator().manual_seed(21)
    code = (torch.randn(2, 3, 6, 128, 128, generator=g) * 0.7).clamp(-2, 2)
    params = rp.make_decoder_params('P', 6)
    params['density_net.0.bias'] = params['density_net.0.bias'] - 2.5        # densities straddle the 0.1 / mean thresholds: mixed bitfields
    dec = tri.TriPlaneDecoder(**CFG['P'])
    dec.load_state_dict({k: torch.as_tensor(v) for k, v in params.items()}, strict=False)
    dec.eval()
    fake = types.SimpleNamespace(grid_size=64, parameters=lambda: iter([torch.zeros(1)]))
    for name in ('get_init_density_grid', 'get_init_density_bitfield', 'update_extra_state', 'get_density'):
        setattr(fake, name, types.MethodType(getattr(base.BaseNeRF, name), fake))
    torch.manual_seed(5)
    grid, bits = fake.get_density(dec, code, cfg=dict(density_thresh=0.1, density_step=3))
    out['gd_grid_sub'], out['gd_grid_sum'] = grid.float().numpy()[:, ::37].copy(), np.array(float(grid.double().sum()))
    out['gd_grid_dtype'], out['gd_bits'] = np.array(str(grid.dtype)), bits.numpy()
    # two EMA updates with the training defaults (decay 0.9; thresh 0.08 so that the bitfield is mixed) on an fp32 grid (val_guide's own allocation, diffusion_nerf.py:279)
    grid2 = torch.zeros(2, 64 ** 3)
    bits2 = torch.zeros(2, 64 ** 3 // 8, dtype=torch.uint8)
    torch.manual_seed(6)
    for i in range(2):
        fake.update_extra_state(dec, code * (1.0 if i == 0 else 0.5), grid2, bits2, i, density_thresh=0.08)
        out[f'ue_grid_sub_{i}'], out[f'ue_grid_sum_{i}'] = grid2.numpy()[:, ::37].copy(), np.array(float(grid2.double().sum()))
        out[f'ue_bits_{i}'] = bits2.numpy().copy()
    return out


if __name__ == '__main__':
    res = run_reference()
    path = os.path.join(HERE, 'reference_density_v1.npz')
    np.savez_compressed(path, **res)
    print({k: (v.shape, float(np.asarray(v, np.float64).mean()) if v.ndim else v) for k, v in res.items()}, os.path.getsize(path), 'bytes')
    print('occupied fraction', {k: float(np.unpackbits(v).mean()) for k, v in res.items() if 'bits' in k})
