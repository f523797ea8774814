"""diagnostic: which forward buffers differ run to run (same input)?  usage: fwd_determinism.py [small|full]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from oracle import unet_port as up
from ssdnerf_b200.unet import DenoisingUnetMod
which = sys.argv[1] if len(sys.argv) > 1 else 'small'
if which == 'small':
    cfg = dict(image_size=32, in_channels=18, base_channels=64, channels_cfg=[1, 2, 2], resblocks_per_downsample=1, num_heads=2, attention_res=[16, 8], use_scale_shift_norm=True)
    std, B, res = 0.04, 3, 32
else:
    cfg = dict(image_size=128, in_channels=18, base_channels=128, channels_cfg=[1, 2, 2, 4, 4], resblocks_per_downsample=2, num_heads=4, attention_res=[32, 16, 8], use_scale_shift_norm=True)
    std, B, res = 0.02, 2, 128
dev = torch.device('cuda:0')
spec = up.unet_spec(**{k: v for k, v in cfg.items() if k != 'use_scale_shift_norm'})
sd = up.random_state_dict(spec, seed=1, std=std)
m = DenoisingUnetMod(**cfg); m.load_state_dict(sd); m = m.to(dev).eval().requires_grad_(False)
g = torch.Generator().manual_seed(12)
x = torch.randn(B, 18, res, res, generator=g).to(dev)
t = torch.tensor([999, 400, 19][:B]).to(dev)
rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))
eng = m.engine(B, dev)
eng.set_embedding(m.embedding(t))
def run(save):
    eng.load_input_nchw(x); v = eng.forward_nhwc(save=save).clone(); torch.cuda.synchronize()
    return v, {k: b.clone() for k, b in eng.bufs.items()}, eng.qarena[:eng.qoff].clone()
for save in (True, False):
    v0, b0, q0 = run(save); v1, b1, q1 = run(save)
    print(which, 'save' if save else 'shared-scratch', 'forward run-to-run rel', rel(v1, v0), 'qarena rel', rel(q1, q0), 'max abs', float((q1 - q0).abs().max()))
    diffs = [(rel(b1[k].float(), b0[k].float()), str(k)) for k in b0 if b0[k].dtype in (torch.float16, torch.float32)]
    order = list(b0.keys())
    first = [(str(k), rel(b1[k].float(), b0[k].float())) for k in order][:60]
    for k, e in first:
        if e > 0: print('   %.3e %s' % (e, k))
