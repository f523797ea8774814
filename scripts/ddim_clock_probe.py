"""SM clock / power while the DDIM loop runs back to back (is the stage power-capped?)"""
import sys, os, subprocess, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import build_model
dev = torch.device('cuda:0')
model, cfg = build_model(dev)
diffusion = model.diffusion_ema
noise = torch.randn(16, 18, 128, 128, device=dev)
for _ in range(2):
    diffusion(noise, return_loss=False)
torch.cuda.synchronize()
lines = []
proc = subprocess.Popen(['nvidia-smi', '-i', '0', '--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap', '--format=csv,noheader,nounits', '-lms', '100'],
                        stdout=subprocess.PIPE, text=True)
threading.Thread(target=lambda: [lines.append(l.strip()) for l in proc.stdout], daemon=True).start()
time.sleep(0.5)
n = int(os.environ.get('N', 12))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
per = []
for i in range(n):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); diffusion(noise, return_loss=False); b.record()
    per.append((a, b))
torch.cuda.synchronize()
time.sleep(0.3)
proc.terminate()
print('ddim ms per call:', ' '.join(f'{a.elapsed_time(b):.1f}' for a, b in per))
print('clock MHz, power W, power-cap:', ' | '.join(lines))
