"""ad-hoc timing of the fused renderer (not the bench contract; see bench.py)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from oracle import render_port as rp
from ssdnerf_b200 import renderer as R
from tests.common import spiral_poses

dev = torch.device('cuda:0')
variant = sys.argv[1] if len(sys.argv) > 1 else 'P'
B, V, res = int(os.environ.get('B', 16)), int(os.environ.get('V', 8)), 128
vid = {'P': R.DEC_P, 'P_SIMT': R.DEC_P_SIMT, 'P_TC': R.DEC_P_TC, 'P_MMA': R.DEC_P_MMA, 'P_MMA2': R.DEC_P_MMA2, 'S': R.DEC_S, 'S_MMA': R.DEC_S_MMA, 'S_TC': R.DEC_S_TC}[variant]
C = 32 if variant[0] == 'S' else 6
g = torch.Generator().manual_seed(0)
code = torch.randn(B, 3, C, 128, 128, generator=g).clamp(-2, 2).to(dev)
params = rp.make_decoder_params(variant[0], 0)
blob = R.pack_decoder_blob(params, vid, device=dev)
planes = R.pack_planes(code, vid)
import numpy as _np
bf = torch.from_numpy(rp.sphere_bitfield() if os.environ.get('GRID', 'sphere') == 'sphere' else _np.full(64**3//8, 255, _np.uint8))[None].repeat(B, 1).to(dev)
poses = torch.from_numpy(spiral_poses(V))[None].repeat(B, 1, 1, 1).to(dev)
intr = torch.tensor([131.25, 131.25, 64, 64]).expand(B, V, 4).contiguous().to(dev)
for emu in (1,):
    for _ in range(3):
        out = R.render_fwd(vid, planes, (128, 128), bf, blob, poses=poses, intrinsics=intr, img_hw=(res, res), emulate_schedule=bool(emu))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        out = R.render_fwd(vid, planes, (128, 128), bf, blob, poses=poses, intrinsics=intr, img_hw=(res, res), emulate_schedule=bool(emu))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    ns = out['num_samples'].sum().item()
    rays = B * V * res * res
    if variant == 'P_MMA2':
        prof = torch.zeros(8, dtype=torch.int64, device=dev)
        R.render_fwd(vid, planes, (128, 128), bf, blob, poses=poses, intrinsics=intr, img_hw=(res, res), emulate_schedule=False, debug_phase_cycles=prof)
        pc = prof.cpu().numpy()
        print('decode iterations', int(pc[0]), 'samples', int(pc[1]), 'lane utilisation %.3f' % (pc[1] / (32.0 * max(pc[0], 1))))
    if variant == 'P_TC':
        prof = torch.zeros(8, dtype=torch.int64, device=dev)
        R.render_fwd(vid, planes, (128, 128), bf, blob, poses=poses, intrinsics=intr, img_hw=(res, res), emulate_schedule=False, debug_phase_cycles=prof)
        pc = prof.cpu().numpy()[:6].astype(float)
        print('phase share [probe, gather, fence+bar, mma wait, heads, composite]:', (pc / pc.sum()).round(3), 'cycles per CTA-thread0 total', pc.sum() / (148 * 2))
    print(f'variant {variant} emulate={emu}: {ms:.3f} ms  rays/s={rays/ms*1e3:.3e}  samples={ns} samples/s={ns/ms*1e3:.3e} mean rgb={out["rgb"].mean().item():.4f}')
