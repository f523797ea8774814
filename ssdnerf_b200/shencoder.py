"""Spherical-harmonics direction encoding as a stand-alone op (C ABI §1: ssdnerf_sh_encode_forward / _backward).

API contract kept from the reference's ``lib.ops.shencoder`` (lib/ops/shencoder/sphere_harmonics.py:61-87): the module name
``SHEncoder(input_dim=3, degree=4)``, its ``output_dim`` attribute, ``forward(inputs, size=1)`` and the functional
``sh_encode(inputs, degree, calc_grad_inputs)``.  The fused renderers never call this: they evaluate the 16 basis values once
per RAY in registers (csrc/common.cuh `sh16`); this op serves direct users of the encoder.
"""
import torch
import torch.nn as nn

from . import _lib as N

MAX_DEGREE = 4


def _launch_forward(dirs, degree, want_jacobian):
    """dirs fp32 [M,3] on the GPU -> (basis [M, degree^2], jacobian [M, 3*degree^2] | 1-element placeholder)"""
    M = dirs.shape[0]
    n_basis = degree * degree
    basis = dirs.new_empty(M, n_basis)
    jac = dirs.new_empty(M, 3 * n_basis) if want_jacobian else dirs.new_empty(1)
    N.check(N.lib().ssdnerf_sh_encode_forward(N.ptr(dirs), N.ptr(basis), N.c_u32(M), N.c_u32(3), N.c_u32(degree),
                                              N.c_int(1 if want_jacobian else 0), N.ptr(jac), N.stream_ptr()))
    return basis, jac


class _SHBasis(torch.autograd.Function):
    """basis values with an optional gradient w.r.t. the directions (the kernel stores d basis / d dir in forward)"""

    @staticmethod
    def forward(ctx, dirs, degree, want_dir_grad):
        N.require_cuda(dirs)
        dirs = dirs.contiguous().float()
        basis, jac = _launch_forward(dirs, degree, want_dir_grad)
        ctx.degree, ctx.want_dir_grad = degree, want_dir_grad
        if want_dir_grad:
            ctx.save_for_backward(dirs, jac)
        return basis

    @staticmethod
    def backward(ctx, grad_basis):
        if not ctx.want_dir_grad:
            return None, None, None
        dirs, jac = ctx.saved_tensors
        grad_dirs = torch.zeros_like(dirs)
        N.check(N.lib().ssdnerf_sh_encode_backward(N.ptr(grad_basis.contiguous().float()), N.ptr(dirs), N.c_u32(dirs.shape[0]), N.c_u32(3),
                                                   N.c_u32(ctx.degree), N.ptr(jac), N.ptr(grad_dirs), N.stream_ptr()))
        return grad_dirs, None, None


def sh_encode(inputs, degree, calc_grad_inputs=False):
    return _SHBasis.apply(inputs, degree, calc_grad_inputs)


class SHEncoder(nn.Module):
    def __init__(self, input_dim=3, degree=4):
        super().__init__()
        if input_dim != 3:
            raise ValueError('SH encoder only support input dim == 3')
        if not 1 <= degree <= MAX_DEGREE:
            raise ValueError(f'this build evaluates SH up to degree {MAX_DEGREE} (the model uses 4 -> 16 features), got {degree}')
        self.input_dim, self.degree, self.output_dim = input_dim, degree, degree * degree

    def extra_repr(self):
        return f'input_dim={self.input_dim}, degree={self.degree}'

    def forward(self, inputs, size=1):
        flat = (inputs / size).reshape(-1, self.input_dim)
        out = sh_encode(flat, self.degree, flat.requires_grad)
        return out.reshape(*inputs.shape[:-1], self.output_dim)
