"""representative UNet convs for an ncu --set full capture: row-pair kernel (128x128 level), CTA-pair tile kernel (64x64 level), single-CTA tile
kernel (16x16 level), fused attention (32x32 level)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ssdnerf_b200 import unet_ops as U
dev = torch.device('cuda:0')
for (H, Cin, Cout) in [(128, 128, 128), (64, 256, 256), (16, 512, 512)]:
    x = torch.randn(16, H, H, Cin, device=dev).half()
    wp = U.pack_conv_weight(torch.randn(Cout, Cin, 3, 3) * 0.02).to(dev)
    b = torch.randn(Cout, device=dev)
    q = torch.zeros(16, Cout // 4, 2, device=dev)
    for _ in range(2):
        U.conv3x3_f16(x, wp, Cout, bias=b, qstats=q)
qkv = torch.randn(16, 1024, 768, device=dev).half()
for _ in range(2):
    U.flash_attn(qkv, 4, 0.125)
torch.cuda.synchronize()
