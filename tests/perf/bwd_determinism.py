"""diagnostic: run-to-run differences of the UNet forward and of the input-gradient pass (same inputs)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from oracle import unet_port as up
from ssdnerf_b200.unet import DenoisingUnetMod
SMALL = dict(image_size=32, in_channels=18, base_channels=64, channels_cfg=[1, 2, 2], resblocks_per_downsample=1, num_heads=2, attention_res=[16, 8], use_scale_shift_norm=True)
dev = torch.device('cuda:0')
spec = up.unet_spec(**{k: v for k, v in SMALL.items() if k != 'use_scale_shift_norm'})
sd = up.random_state_dict(spec, seed=1, std=0.04)
m = DenoisingUnetMod(**SMALL); m.load_state_dict(sd); m = m.to(dev).eval().requires_grad_(False)
g = torch.Generator().manual_seed(12)
x = torch.randn(3, 18, 32, 32, generator=g).to(dev); r = (torch.randn(3, 18, 32, 32, generator=g) * 1e-4).to(dev)
t = torch.tensor([999, 400, 19]).to(dev)
rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
eng = m.engine(3, dev)
eng.set_embedding(m.embedding(t))
outs = []
for i in range(3):
    eng.load_input_nchw(x); outs.append(eng.forward_nhwc(save=True).clone())
print('forward run-to-run', rel(outs[1], outs[0]), rel(outs[2], outs[0]))
gs = [eng.backward_nchw(r).clone() for _ in range(3)]
print('backward (same saved forward) run-to-run', rel(gs[1], gs[0]), rel(gs[2], gs[0]))
# per-record gradient comparison between two backward runs
import ssdnerf_b200.unet_ops as U
def snap():
    return {k: v.clone() for k, v in eng.bufs.items() if isinstance(k, tuple) and k and k[0] == 'bwd' and v.dtype in (torch.float16, torch.float32)}
eng.backward_nchw(r); a = snap(); eng.backward_nchw(r); b = snap()
bad = sorted(((rel(b[k].float(), a[k].float()) if a[k].float().norm() > 0 else 0.0), str(k)) for k in a)
print('largest buffer differences:'); [print('  %.3e %s' % e) for e in bad[-12:]]
