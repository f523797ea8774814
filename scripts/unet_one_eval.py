"""one UNet evaluation at B=16 (for ncu launch lists / captures)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import build_model
dev = torch.device('cuda:0')
model, cfg = build_model(dev)
unet = model.diffusion_ema.denoising
B = int(os.environ.get('B', 16))
x = torch.randn(B, 18, 128, 128, device=dev)
t = torch.full((B,), 500, device=dev, dtype=torch.long)
n = int(os.environ.get('N', 2))
for _ in range(n):
    v = unet(x, t)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
v = unet(x, t)
e1.record(); torch.cuda.synchronize()
print('unet eval ms', e0.elapsed_time(e1), float(v.std()))
