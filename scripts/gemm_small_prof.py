"""pipeline wait counters for the short-K GEMMs of the attention blocks / 1x1 convolutions"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ssdnerf_b200 import unet_ops as U
dev = torch.device('cuda:0')
names = ['prod wait-empty', 'prod total', 'mma wait-full', 'mma wait-tmem-empty', 'mma total', 'epi wait-tmem-full', 'epi total']
for M, Nn, K, res, qs, cfgs in [(16384, 768, 256, False, False, [(256, 0), (256, 1), (128, 1)]), (16384, 256, 256, True, True, [(256, 1), (128, 1), (64, 1)]),
                                (4096, 1536, 512, False, False, [(128, 0), (256, 1)]), (1024, 512, 512, True, True, [(64, 1), (128, 1)]),
                                (262144, 128, 256, False, True, [(128, 1)])]:
    a = torch.randn(M, K, device=dev).half()
    w = U.pack_linear_weight(torch.randn(Nn, K) * 0.05).to(dev)
    bias = torch.randn(Nn, device=dev)
    r = torch.randn(M, Nn, device=dev).half() if res else None
    out = torch.empty(M, Nn, dtype=torch.float16, device=dev)
    T = 1024 if M >= 16384 else 64
    for bn, cl in cfgs:
        q = torch.zeros(M // T, Nn // 4, 2, device=dev) if qs else None
        kw = dict(bias=bias, residual=r, out=out, bn=bn, cluster=cl, n=Nn, qstats=q, stats_hw=T if qs else 0)
        U.GEMM_PROF = None
        for _ in range(3):
            U.linear_f16(a, w, **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            U.linear_f16(a, w, **kw)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        prof = torch.zeros(8, dtype=torch.int64, device=dev)
        U.GEMM_PROF = prof
        U.linear_f16(a, w, **kw)
        torch.cuda.synchronize()
        U.GEMM_PROF = None
        c = prof.tolist()
        tiles = (M // 128) * ((Nn + bn - 1) // bn)
        ctas = min(148, tiles)
        print(f'M={M} N={Nn} K={K} bn={bn} cluster={cl} res={int(res)} qstats={int(qs)}: {us:6.1f} us back-to-back; {tiles} tiles on <= {ctas} CTAs; '
              + ', '.join(f'{nm} {c[i] / ctas / 1e3:.1f}k' for i, nm in enumerate(names)), flush=True)
