"""profiles/*.ncu-rep (one `ncu --set full` capture of a render kernel inside the bench command) -> profiles/r02_render_traffic.json:
dram__bytes_read.sum + dram__bytes_write.sum per launch, with the launch's ray count, so bench.py can report `roofline.traffic` from a
measurement instead of a constant.  usage: python scripts/ncu_traffic.py profiles/r02_render_p3.ncu-rep k_render_p3 <rays_per_launch>"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rep, kernel, rays = sys.argv[1], sys.argv[2], int(sys.argv[3])
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], stdout=subprocess.PIPE, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    best = None
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        if kernel in d['Kernel Name']:
            dur = float(d['gpu__time_duration.sum'].replace(',', ''))
            if best is None or dur > best[0]:
                best = (dur, d)
    if best is None:
        raise SystemExit(f'no launch of {kernel} in {rep}')
    d = best[1]
    u = dict(zip(hdr, units))

    def to_bytes(key):
        v = float(d[key].replace(',', ''))
        return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u[key]]
    rd, wr = to_bytes('dram__bytes_read.sum'), to_bytes('dram__bytes_write.sum')
    path = os.path.join(ROOT, 'profiles', 'r02_render_traffic.json')
    data = json.load(open(path)) if os.path.exists(path) else {}
    data[kernel] = dict(dram_bytes_per_launch=rd + wr, dram_bytes_read=rd, dram_bytes_write=wr, rays_per_launch=rays,
                        duration_under_ncu=f"{d['gpu__time_duration.sum']} {u['gpu__time_duration.sum']}",
                        source=f'{os.path.relpath(rep, ROOT)}: ncu --set full --clock-control none, longest launch of {kernel}')
    json.dump(data, open(path, 'w'), indent=1)
    print(json.dumps(data[kernel], indent=1))


if __name__ == '__main__':
    main()
