"""Mirror of the reference's ``lib.ops.shencoder`` (lib/ops/shencoder/sphere_harmonics.py:15-87).

The fused renderer evaluates SH16 once per RAY in registers (ssdnerf_b200/csrc/common.cuh `sh16`); this
module is the stand-alone op for callers that use ``SHEncoder`` directly (e.g. TriPlaneDecoder.point_decode).
"""
import torch
import torch.nn as nn
from torch.autograd import Function

from . import _lib as N


class _sh_encoder(Function):
    @staticmethod
    def forward(ctx, inputs, degree, calc_grad_inputs=False):
        N.require_cuda(inputs)
        inputs = inputs.contiguous().float()
        B, input_dim = inputs.shape
        output_dim = degree ** 2
        outputs = torch.empty(B, output_dim, dtype=torch.float32, device=inputs.device)
        dy_dx = torch.empty(B, input_dim * output_dim, dtype=torch.float32, device=inputs.device) if calc_grad_inputs \
            else torch.empty(1, dtype=torch.float32, device=inputs.device)
        N.check(N.lib().ssdnerf_sh_encode_forward(N.ptr(inputs), N.ptr(outputs), N.c_u32(B), N.c_u32(input_dim),
                                                  N.c_u32(degree), N.c_int(int(calc_grad_inputs)), N.ptr(dy_dx),
                                                  N.stream_ptr()))
        ctx.save_for_backward(inputs, dy_dx)
        ctx.dims = [B, input_dim, degree]
        ctx.calc_grad_inputs = calc_grad_inputs
        return outputs

    @staticmethod
    def backward(ctx, grad):
        if not ctx.calc_grad_inputs:
            return None, None, None
        grad = grad.contiguous().float()
        inputs, dy_dx = ctx.saved_tensors
        B, input_dim, degree = ctx.dims
        grad_inputs = torch.zeros_like(inputs)
        N.check(N.lib().ssdnerf_sh_encode_backward(N.ptr(grad), N.ptr(inputs), N.c_u32(B), N.c_u32(input_dim),
                                                   N.c_u32(degree), N.ptr(dy_dx), N.ptr(grad_inputs), N.stream_ptr()))
        return grad_inputs, None, None


sh_encode = _sh_encoder.apply


class SHEncoder(nn.Module):
    """sphere_harmonics.py:61-87 (degree <= 4 here; the model uses 4 -> 16 features)."""

    def __init__(self, input_dim=3, degree=4):
        super().__init__()
        self.input_dim = input_dim
        self.degree = degree
        self.output_dim = degree ** 2
        assert self.input_dim == 3, 'SH encoder only support input dim == 3'
        assert 0 < self.degree <= 4, 'this build supports SH degree in [1, 4]'

    def __repr__(self):
        return f'SHEncoder: input_dim={self.input_dim} degree={self.degree}'

    def forward(self, inputs, size=1):
        inputs = inputs / size
        prefix_shape = list(inputs.shape[:-1])
        inputs = inputs.reshape(-1, self.input_dim)
        outputs = sh_encode(inputs, self.degree, inputs.requires_grad)
        return outputs.reshape(prefix_shape + [self.output_dim])
