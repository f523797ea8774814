"""two representative UNet convs for an ncu --set full capture"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ssdnerf_b200 import unet_ops as U
dev = torch.device('cuda:0')
for (H, Cin, Cout) in [(64, 256, 256), (128, 128, 128)]:
    x = torch.randn(16, H, H, Cin, device=dev).half()
    wp = U.pack_conv_weight(torch.randn(Cout, Cin, 3, 3) * 0.02).to(dev)
    for _ in range(2):
        U.conv3x3_f16(x, wp, Cout)
torch.cuda.synchronize()
