"""diagnostic: decoder weight gradients of the fused backward vs fp64 oracle autograd, broken down per parameter / row / column"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import train_port as tp  # noqa: E402
from tests.test_train_render_gpu import _case, _decoder, _rel_l2  # noqa: E402
from ssdnerf_b200 import renderer as R  # noqa: E402

cuda = torch.device('cuda:0')
T_thresh = 1e-4
code, rays_o, rays_d, params, bf, noises, dt_gamma, target = _case(n_rays=256)
B, n = rays_o.shape[:2]
g = torch.Generator().manual_seed(2)
g_img, g_ws = torch.randn(B, n, 3, generator=g), torch.randn(B, n, generator=g)
pref = {k: torch.as_tensor(v).double().requires_grad_(True) for k, v in params.items()}
cref = code.double().requires_grad_(True)
tot = 0
for b in range(B):
    ws, _, img = tp.render_train_scene(pref, cref[b], rays_o[b].numpy(), rays_d[b].numpy(), bf[b], noises[b].numpy(),
                                       dt_gamma=float(dt_gamma[b]), T_thresh=T_thresh)
    tot = tot + (img * g_img[b].double()).sum() + (ws * g_ws[b].double()).sum()
names = list(R.DEC_P_PARAM_ORDER)
gr = torch.autograd.grad(tot, [cref] + [pref[k] for k in names])
gref = dict(zip(names, gr[1:]))
bft = torch.from_numpy(bf).to(cuda)
dec = _decoder(params, cuda, frozen=False)
c = code.to(cuda).requires_grad_(True)
out = dec(rays_o.to(cuda), rays_d.to(cuda), c, bft, 64, dt_gamma=dt_gamma.tolist(), perturb=noises.to(cuda), T_thresh=T_thresh)
loss = (out['image'] * g_img.to(cuda)).sum() + (out['weights_sum'] * g_ws.to(cuda)).sum()
loss.backward()
print('code grad rel', _rel_l2(c.grad, gr[0]))
got = dict(dec.named_parameters())
for k in names:
    a, r = got[k].grad.double().cpu(), gref[k]
    print(k, 'rel', _rel_l2(a, r), 'ratio of norms', float(a.norm() / r.norm()), 'cos', float((a * r).sum() / a.norm() / r.norm()))
a, r = got['base_net.0.weight'].grad.double().cpu(), gref['base_net.0.weight']
print('W1 per input column (c*3+plane):', [round(float((a[:, j] - r[:, j]).norm() / r[:, j].norm()), 3) for j in range(18)])
print('W1 per output row first 16:', [round(float((a[j] - r[j]).norm() / r[j].norm()), 3) for j in range(16)])
print('W1 elementwise ratio sample:', (a / r)[:3, :6])
a, r = got['dir_net.0.weight'].grad.double().cpu(), gref['dir_net.0.weight']
print('Wdir per input column:', [round(float((a[:, j] - r[:, j]).norm() / r[:, j].norm()), 3) for j in range(16)])
