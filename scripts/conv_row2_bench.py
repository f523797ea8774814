"""row-pair 3x3 convolution kernel (algo 2) vs the generic tile kernel (algo 1) on the UNet's 128x128-level shapes"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ssdnerf_b200 import unet_ops as U
dev = torch.device('cuda:0')
B, H, W, Cout = 16, 128, 128, 128
for C1, C2 in [(128, 0), (128, 128), (256, 128), (64, 0)]:
    x = torch.randn(B, H, W, C1, device=dev).half()
    x2 = torch.randn(B, H, W, C2, device=dev).half() if C2 else None
    wp = U.pack_conv_weight(torch.randn(Cout, C1 + C2, 3, 3) * 0.02).to(dev)
    bias = torch.randn(Cout, device=dev)
    out = torch.empty(B, H, W, Cout, dtype=torch.float16, device=dev)
    q = torch.zeros(B, Cout // 4, 2, device=dev)
    flops = 2.0 * B * H * W * Cout * (C1 + C2) * 9
    line = []
    for algo in (1, 2):
        for _ in range(3):
            U.conv3x3_f16(x, wp, Cout, bias=bias, x2=x2, out=out, qstats=q, algo=algo)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            U.conv3x3_f16(x, wp, Cout, bias=bias, x2=x2, out=out, qstats=q, algo=algo)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 10 * 1e3
        line.append(f'algo {algo}: {us:7.1f} us {flops / us / 1e6:6.0f} TFLOP/s')
    print(f'Cin={C1}+{C2}: ' + ' | '.join(line), flush=True)
