"""Where does the fp16 error of one UNet evaluation come from?  CPU experiment on the fp32 oracle with fp16 rounding injected at
(A) GEMM operands (activations + weights), (B) conv outputs that feed a GroupNorm (h1, qkv, attention out), (C) the residual stream
(block outputs).  Prints rel-L2 vs the un-rounded fp32 forward.  TEST INFRASTRUCTURE (imports oracle/)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torch.nn.functional as F
from oracle import unet_port as up

torch.set_num_threads(os.cpu_count())
rt = lambda x: x.half().float()
FLAGS = dict(A=False, B=False, C=False, TF32=False)

_conv2d, _conv1d, _linear = F.conv2d, F.conv1d, F.linear


def tf32(x):
    return (x.view(torch.int32) + 0x1000 & ~0x1FFF).view(torch.float32) if False else (x.contiguous().view(torch.int32).add(0x1000).bitwise_and(~0x1FFF)).view(torch.float32)


def conv2d(x, w, b=None, **kw):
    if FLAGS['A']:
        x, w = rt(x), rt(w)
    if FLAGS['TF32']:
        x, w = tf32(x), tf32(w)
    y = _conv2d(x, w, b, **kw)
    return y


def conv1d(x, w, b=None, **kw):
    if FLAGS['A']:
        x, w = rt(x), rt(w)
    return _conv1d(x, w, b, **kw)


class Fp:  # proxy for torch.nn.functional inside unet_port
    def __getattr__(self, k):
        if k == 'conv2d': return conv2d
        if k == 'conv1d': return conv1d
        return getattr(F, k)


up.F = Fp()
_res, _attn = up.res_block, up.attention


def res_block(sd, b, x, emb):
    k = b['key']
    sc = up.F.conv2d(x, sd[k + '.shortcut.weight'], sd[k + '.shortcut.bias']) if b['cin'] != b['cout'] else x
    h = up.F.conv2d(F.silu(up._gn(sd, k + '.conv_1.0', x)), sd[k + '.conv_1.2.weight'], sd[k + '.conv_1.2.bias'], padding=1)
    if FLAGS['B']:
        h = rt(h)
    e = F.linear(F.silu(emb), sd[k + '.norm_with_embedding.embedding_layer.1.weight'], sd[k + '.norm_with_embedding.embedding_layer.1.bias'])[:, :, None, None]
    scale, shift = torch.chunk(e, 2, dim=1)
    h = up._gn(sd, k + '.norm_with_embedding.norm', h) * (1 + scale) + shift
    key2 = k + '.conv_2.1'
    h = up.F.conv2d(F.silu(h), sd[key2 + '.weight'], sd[key2 + '.bias'], padding=1)
    y = h + sc
    return rt(y) if FLAGS['C'] else y


def attention(sd, b, x, nh):
    y = _attn(sd, b, x, nh)
    return rt(y) if FLAGS['C'] else y


up.res_block, up.attention = res_block, attention

small = '--small' in sys.argv
spec = up.unet_spec(image_size=64, base_channels=64) if small else up.unet_spec()
sd = up.random_state_dict(spec, seed=0, std=0.02)
g = torch.Generator().manual_seed(1)
x = torch.randn(1, 18, 64 if small else 128, 64 if small else 128, generator=g)
t = torch.tensor([500])
with torch.no_grad():
    t0 = time.time(); ref = up.unet_forward(sd, spec, x, t); print('fp32 forward %.1fs' % (time.time() - t0), flush=True)
    for name, fl in (('A operands', dict(A=True)), ('B h1 storage', dict(B=True)), ('C residual stream', dict(C=True)), ('A+B', dict(A=True, B=True)),
                     ('A+B+C', dict(A=True, B=True, C=True)), ('TF32 convs (reference default)', dict(TF32=True))):
        FLAGS.update(A=False, B=False, C=False, TF32=False); FLAGS.update(fl)
        y = up.unet_forward(sd, spec, x, t)
        print(f'{name:32s} rel-L2 {float((y - ref).norm() / ref.norm()):.3e}', flush=True)
