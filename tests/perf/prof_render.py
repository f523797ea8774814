"""small render workload for ncu captures: python scripts/prof_render.py <P|P_TC|S> """
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from oracle import render_port as rp
from ssdnerf_b200 import renderer as R
from tests.common import spiral_poses
dev = torch.device('cuda:0')
variant = sys.argv[1]
vid = {'P': R.DEC_P, 'P_TC': R.DEC_P_TC, 'P_MMA': R.DEC_P_MMA, 'S': R.DEC_S, 'S_MMA': R.DEC_S_MMA, 'S_TC': R.DEC_S_TC}[variant]
C = 32 if variant[0] == 'S' else 6
B, V = 4, 8
g = torch.Generator().manual_seed(0)
code = torch.randn(B, 3, C, 128, 128, generator=g).clamp(-2, 2).to(dev)
params = rp.make_decoder_params(variant[0], 0)
blob = R.pack_decoder_blob(params, vid, device=dev)
planes = R.pack_planes(code, vid)
bf = torch.from_numpy(np.full(64 ** 3 // 8, 255, np.uint8))[None].repeat(B, 1).to(dev)
poses = torch.from_numpy(spiral_poses(V))[None].repeat(B, 1, 1, 1).to(dev)
intr = torch.tensor([131.25, 131.25, 64, 64]).expand(B, V, 4).contiguous().to(dev)
for _ in range(2):
    out = R.render_fwd(vid, planes, (128, 128), bf, blob, poses=poses, intrinsics=intr, img_hw=(128, 128), emulate_schedule=False)
torch.cuda.synchronize()
print(variant, out['num_samples'].sum().item())
