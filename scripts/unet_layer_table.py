"""Per-GEMM table of one UNet evaluation: run under ncu (warm launch list) and merge with the Python-side shape log.

  ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/unet_layers.csv \
      python scripts/unet_layer_table.py run
  python scripts/unet_layer_table.py merge > profiles/rNN_unet_layer_table.txt
"""
import csv, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
OUT = 'gpurun_out'

if sys.argv[1] == 'run':
    import torch
    from bench import build_model
    from ssdnerf_b200 import unet_ops
    dev = torch.device('cuda:0')
    model, cfg = build_model(dev)
    unet = model.diffusion_ema.denoising
    B = int(os.environ.get('B', 16))
    x = torch.randn(B, 18, 128, 128, device=dev)
    t = torch.full((B,), 500, device=dev, dtype=torch.long)
    for _ in range(2):
        unet(x, t)
    torch.cuda.synchronize()
    unet_ops.GEMM_LOG = []
    unet(x, t)
    torch.cuda.synchronize()
    os.makedirs(OUT, exist_ok=True)
    json.dump(unet_ops.GEMM_LOG, open(os.path.join(OUT, 'unet_gemm_log.json'), 'w'))
else:
    log = json.load(open(os.path.join(OUT, 'unet_gemm_log.json')))
    rows = list(csv.reader(open(os.path.join(OUT, 'unet_layers.csv'), errors='ignore')))
    hdr, L = None, []
    for r in rows:
        if 'Kernel Name' in r:
            hdr = r; continue
        if hdr is None or len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        try:
            v = float(d['Metric Value'].replace(',', ''))
        except ValueError:
            continue
        L.append((d['Kernel Name'], d['Grid Size'], v / 1e3))
    idx = max(i for i, x in enumerate(L) if 'k_nchw_to_nhwc' in x[0])
    ev = L[idx:]
    gem = [x for x in ev if 'k_gemm_tc' in x[0] or 'k_conv_row2' in x[0]]      # k_conv_row2 and k_conv_row2_gn (fused GroupNorm + SiLU)
    assert len(gem) == len(log), (len(gem), len(log))
    tot = sum(v for _, _, v in ev)
    print(f'one UNet evaluation (B=16), warm per-kernel durations: {tot:.1f} us over {len(ev)} launches; GEMM launches: {len(gem)}')
    other = {}
    for k, g, v in ev:
        if 'k_gemm_tc' not in k and 'k_conv_row2' not in k:
            nm = k.split('(')[0].replace('void ', '').replace('ssdnerf::', '')[:32]
            o = other.setdefault(nm, [0, 0.0]); o[0] += 1; o[1] += v
    for k, v in sorted(other.items(), key=lambda kv: -kv[1][1]):
        print(f'  glue {k:34s} {v[0]:4d} launches {v[1]:9.1f} us')
    agg = {}
    for (k, g, v), a in zip(gem, log):
        key = (a['M'], a['N'], a['K'] * a['taps'], a['taps'], a['batched'], ('gn+row2' if 'row2_gn' in k else (k.split('<')[1].split('>')[0] if 'k_gemm_tc' in k else 'row2')), g)
        e = agg.setdefault(key, [0, 0.0]); e[0] += 1; e[1] += v
    print(f'{"M":>8} {"N":>5} {"Ktot":>6} taps bat  tile<BN,CL> grid        n    us/launch  TFLOP/s   total us  lost-vs-1.3PF us')
    tl = 0
    for key, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        M, N_, K, taps, bat, tile, grid = key
        fl = 2.0 * M * N_ * K
        tf = fl / (us / n * 1e-6) / 1e12
        lost = us - n * fl / 1.3e15 * 1e6
        tl += lost
        print(f'{M:8d} {N_:5d} {K:6d} {taps:4d} {bat:3d}  <{tile:>7s}> {grid:12s} {n:3d} {us / n:10.1f} {tf:9.1f} {us:10.1f} {lost:10.1f}')
    print(f'GEMM total {sum(v for _, _, v in gem):.1f} us; time above a 1.3 PFLOP/s floor: {tl:.1f} us')
