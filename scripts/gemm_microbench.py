"""conv / GEMM micro-benchmark of k_gemm_tc on representative UNet layer shapes (B=16)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ssdnerf_b200 import unet_ops as U
dev = torch.device('cuda:0')
B = 16
shapes = [  # H, Cin, Cout
    (128, 128, 128), (128, 256, 128), (64, 256, 256), (64, 512, 256), (32, 256, 256), (32, 512, 256), (16, 512, 512), (16, 1024, 512), (8, 512, 512), (8, 1024, 512)]
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
for H, Cin, Cout in shapes:
    x = torch.randn(B, H, H, Cin, device=dev).half()
    wp = U.pack_conv_weight(torch.randn(Cout, Cin, 3, 3) * 0.02).to(dev)
    out = torch.empty(B, H, H, Cout, dtype=torch.float16, device=dev)
    flops = 2.0 * B * H * H * Cout * Cin * 9
    res = []
    for bn, cl in [(256, 1), (256, 2), (128, 1), (128, 2), (64, 1), (0, 0)]:
        if bn > 64 and bn // 2 >= Cout:
            continue
        try:
            for _ in range(3):
                U.conv3x3_f16(x, wp, Cout, out=out, bn=bn, cluster=cl)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 20
            e0.record()
            for _ in range(n):
                U.conv3x3_f16(x, wp, Cout, out=out, bn=bn, cluster=cl)
            e1.record(); torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / n * 1e3
            res.append(f'bn{bn}/cl{cl}: {us:7.1f}us {flops / us / 1e6:6.0f}TF')
        except Exception as ex:
            res.append(f'bn{bn}/cl{cl}: ERR {str(ex)[:40]}')
    print(f'H={H:3d} Cin={Cin:4d} Cout={Cout:3d} GF={flops/1e9:6.1f} | ' + ' | '.join(res), flush=True)
