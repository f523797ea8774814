"""short-K GEMMs (attention qkv projection, 1x1 shortcut convolution) for an ncu --set full capture"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ssdnerf_b200 import unet_ops as U
dev = torch.device('cuda:0')
for M, Nn, K in [(16384, 768, 256), (262144, 128, 256)]:
    a = torch.randn(M, K, device=dev).half()
    w = U.pack_linear_weight(torch.randn(Nn, K) * 0.05).to(dev)
    bias = torch.randn(Nn, device=dev)
    out = torch.empty(M, Nn, dtype=torch.float16, device=dev)
    for _ in range(2):
        U.linear_f16(a, w, bias=bias, out=out, n=Nn)
torch.cuda.synchronize()
