"""Where does the GEMM pipeline wait?  Per-role clock64 totals (debug counters of ssdnerf_gemm_args.debug_cycles) for a few conv shapes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ssdnerf_b200 import unet_ops as U
dev = torch.device('cuda:0')
B = 16
names = ['prod wait-empty', 'prod total', 'mma wait-full', 'mma wait-tmem-empty', 'mma total', 'epi wait-tmem-full', 'epi total']
for H, Cin, Cout, cfgs in [(128, 128, 128, [(128, 1), (128, 2), (128, 4), (128, 8)]), (64, 256, 256, [(256, 1), (256, 2), (256, 4), (256, 8)]), (64, 512, 256, [(256, 1), (256, 8)]), (128, 256, 128, [(128, 1), (128, 8)]), (8, 512, 512, [(64, 1), (128, 1)])]:
    x = torch.randn(B, H, H, Cin, device=dev).half()
    wp = U.pack_conv_weight(torch.randn(Cout, Cin, 3, 3) * 0.02).to(dev)
    out = torch.empty(B, H, H, Cout, dtype=torch.float16, device=dev)
    bias = torch.randn(Cout, device=dev)
    for bn, cl in cfgs:
        for qs in (True,):
            q = torch.zeros(B, Cout // 4, 2, device=dev) if qs else None
            U.GEMM_PROF = None
            for _ in range(3):
                U.conv3x3_f16(x, wp, Cout, out=out, bn=bn, cluster=cl, bias=bias, qstats=q)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                U.conv3x3_f16(x, wp, Cout, out=out, bn=bn, cluster=cl, bias=bias, qstats=q)
            e1.record(); torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / 10 * 1e3
            prof = torch.zeros(8, dtype=torch.int64, device=dev)
            U.GEMM_PROF = prof
            U.conv3x3_f16(x, wp, Cout, out=out, bn=bn, cluster=cl, bias=bias, qstats=q)
            torch.cuda.synchronize()
            U.GEMM_PROF = None
            c = prof.tolist()
            tiles = (B * H * H // 128) * ((Cout + bn - 1) // bn)
            ctas = min(148, tiles)
            kblocks = 9 * Cin // 64
            print(f'H={H} Cin={Cin} Cout={Cout} bn={bn} cl={cl} qstats={int(qs)}: {us:7.1f} us; {tiles} tiles x {kblocks} k-blocks on {ctas} CTAs')
            n_mma = ctas // 2 if cl == 2 else ctas
            for i, nm in enumerate(names):
                div = n_mma if nm.startswith('mma') else ctas
                print(f'    {nm:22s} {c[i] / div / 1e3:9.1f} kcyc per CTA')
            per_kb = c[4] / n_mma / (tiles / ctas * kblocks) if n_mma else 0
            print(f'    MMA-thread cycles per k-block: {per_kb:.0f} (of which waiting for operands {c[2] / n_mma / (tiles / ctas * kblocks):.0f})', flush=True)
