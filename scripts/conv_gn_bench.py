"""fused GroupNorm + SiLU + conv (csrc/conv_row2_gn.cu) vs GroupNorm-apply pass + row-pair convolution, 128x128 level, B=16 (ncu for the times)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ssdnerf_b200 import _lib as N
from ssdnerf_b200 import unet_ops as U
dev = torch.device('cuda:0')
B, H, W, Cout = 16, 128, 128, 128
for C1, C2 in [(128, 0), (128, 128)]:
    C = C1 + C2
    x1 = torch.randn(B, H, W, C1, device=dev).half()
    x2 = torch.randn(B, H, W, C2, device=dev).half() if C2 else None
    q = lambda x: torch.stack([x.float().view(B, -1, x.shape[-1] // 4, 4).sum(dim=(1, 3)), (x.float() ** 2).view(B, -1, x.shape[-1] // 4, 4).sum(dim=(1, 3))], dim=-1).contiguous()
    q1, q2 = q(x1), (q(x2) if C2 else None)
    gamma, beta = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    wp = U.pack_conv_weight(torch.randn(Cout, C, 3, 3) * 0.02).to(dev)
    bias = torch.randn(Cout, device=dev)
    out = torch.empty(B, H, W, Cout, dtype=torch.float16, device=dev)
    y = torch.empty(B, H, W, C, dtype=torch.float16, device=dev)
    qo = torch.zeros(B, Cout // 4, 2, device=dev)

    def fused():
        U.conv3x3_gn_f16(x1, q1, gamma, beta, wp, bias=bias, x2=x2, q2=q2, out=out, qstats=qo)

    def split():
        N.check(N.lib().ssdnerf_gn_apply_q(N.ptr(x1), N.c_u32(C1), N.ptr(x2), N.c_u32(C2), N.c_u32(B), N.c_u32(H * W), N.c_u32(32), N.ptr(q1), N.ptr(q2),
                                           N.ptr(gamma), N.ptr(beta), None, N.c_longlong(0), N.c_f32(1e-5), N.c_int(1), N.ptr(y), N.stream_ptr()))
        U.conv3x3_f16(y, wp, Cout, bias=bias, out=out, qstats=qo)
    for name, fn in (('fused', fused), ('split', split)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record(); torch.cuda.synchronize()
        print(f'Cin={C1}+{C2} {name}: {e0.elapsed_time(e1) / 10 * 1e3:7.1f} us', flush=True)
    prof = torch.zeros(8, dtype=torch.int64, device=dev)
    U.GEMM_PROF = prof
    fused()
    torch.cuda.synchronize()
    U.GEMM_PROF = None
    c = [v / 148 / 1e3 for v in prof.tolist()]
    print(f'  per-CTA kcycles: MMA wait rows {c[0]:.1f}, wait weights {c[1]:.1f}, wait TMEM {c[2]:.1f}, total {c[3]:.1f} | loader wait free row {c[4]:.1f}, total {c[5]:.1f} | '
          f'weight producer wait {c[6]:.1f}, total {c[7]:.1f}', flush=True)
