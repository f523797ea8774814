#!/usr/bin/env bash
# warm-cache per-kernel durations of UNet evaluations (no cache flush between kernels, clocks untouched)
mkdir -p gpurun_out
N=3 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/unet_launches_warm.csv python scripts/unet_one_eval.py > gpurun_out/unet_warm.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open("gpurun_out/unet_launches_warm.csv", errors="ignore")))
hdr = None; L = []
for r in rows:
    if "Kernel Name" in r: hdr = r; continue
    if hdr is None or len(r) != len(hdr): continue
    d = dict(zip(hdr, r))
    try: v = float(d["Metric Value"].replace(",", ""))
    except: continue
    L.append((d["Kernel Name"], d["Grid Size"], v / 1e3))
# keep only the last evaluation: find the last k_nchw_to_nhwc launch
idx = max(i for i, x in enumerate(L) if "k_nchw_to_nhwc" in x[0])
ev = L[idx:]
agg = collections.defaultdict(lambda: [0, 0.0])
for k, g, v in ev:
    name = k.split("(")[0].replace("void ", "").replace("ssdnerf::", "")[:40]
    agg[name][0] += 1; agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print("one UNet eval, warm per-kernel sum: %.1f us over %d launches" % (tot, len(ev)))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
    print(f"{v[1]:9.1f} us {v[0]:4d} {100*v[1]/tot:5.1f}%  {k}")
g = collections.defaultdict(lambda: [0, 0.0])
for k, gs, v in ev:
    if "k_gemm_tc" in k:
        g[(k.split("(")[0][-20:], gs)][0] += 1; g[(k.split("(")[0][-20:], gs)][1] += v
for k, v in sorted(g.items(), key=lambda kv: -kv[1][1])[:14]:
    print("gemm", k, v[0], round(v[1], 1))
PY
