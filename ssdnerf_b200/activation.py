"""Density activation of the module path: sigma = exp(x) with a bounded derivative.

API mirror of the reference's `lib/ops/activation.py` (`trunc_exp`, `TruncExp`, cited in SURVEY.md §8 a20): the forward value is the
plain fp32 exponential, the backward multiplies the incoming gradient by the exponential limited to [1e-6, 1e6] so a saturated density
neither kills nor explodes the gradient.  The fused renderers (csrc/render_*.cu, csrc/render_train.cu) fold the same rule into their
decoder epilogue / backward; this module only serves callers that compose a decoder out of nn.Modules (per-op train branch).
"""
import torch
from torch import nn

_GRAD_FLOOR, _GRAD_CEIL = 1e-6, 1e6


class _BoundedGradExp(torch.autograd.Function):
    """y = exp(x) in float32;  dy/dx := min(max(y, 1e-6), 1e6)"""

    @staticmethod
    def forward(x):
        return x.to(torch.float32).exp()

    @staticmethod
    def setup_context(ctx, inputs, output):
        ctx.save_for_backward(output)

    @staticmethod
    def backward(ctx, grad_out):
        (y,) = ctx.saved_tensors
        return grad_out * torch.clamp(y, _GRAD_FLOOR, _GRAD_CEIL)


def trunc_exp(x):
    return _BoundedGradExp.apply(x)


class TruncExp(nn.Module):
    """nn.Module wrapper (registered under 'trunc_exp' in TriPlaneDecoder.activation_dict)"""

    def forward(self, x):
        return trunc_exp(x)
