"""Mirror of lib/ops/activation.py:8-44 (TruncExp): exp forward in fp32, gradient g * clamp(exp x, 1e-6, 1e6).

Inside the fused renderers the activation is folded into the decoder epilogue; this module serves callers
that build a decoder out of nn.Modules."""
import torch
import torch.nn as nn
from torch.autograd import Function


class _trunc_exp(Function):
    @staticmethod
    def forward(ctx, x):
        exp_x = torch.exp(x.float())
        ctx.save_for_backward(exp_x)
        return exp_x

    @staticmethod
    def backward(ctx, g):
        return g * ctx.saved_tensors[0].clamp(min=1e-6, max=1e6)


trunc_exp = _trunc_exp.apply


class TruncExp(nn.Module):
    @staticmethod
    def forward(x):
        return _trunc_exp.apply(x)
