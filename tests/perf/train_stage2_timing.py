"""Stage-2 (denoiser) training step at the reference config's batch, one GPU: ms per train_step + a coarse split (forward+loss, backward
incl. weight gradients, optimizer step).   python tests/perf/train_stage2_timing.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench  # noqa: E402


def main():
    dev = torch.device('cuda:0')
    stream = torch.cuda.current_stream(dev)

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(steps):
            fn()
        b.record(stream)
        torch.cuda.synchronize()
        return a.elapsed_time(b) / steps
    r = bench.train_stage2_measurement(dev, timed, 0, steps=6)
    print(json.dumps(dict(ms_per_train_step=r['ms'], triplanes_per_sec=r['scenes'] / r['ms'] * 1e3, **r['info'])))


if __name__ == '__main__':
    main()
