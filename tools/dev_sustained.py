"""Long sustained run (default 200 steps ~ 4.6 s) of the bench workload with NVML sampling: steady-state ms/step,
SM clock, board power and throttle reasons — what the 10-step bench region converges to."""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pynvml
from densepose_torchscript_b200 import synth
from densepose_torchscript_b200.config import BUILTIN
from densepose_torchscript_b200.engine import Engine

N = int(os.environ.get("STEPS", "200"))
spec = BUILTIN["densepose_rcnn_R_50_FPN_s1x"]
eng = Engine(spec, synth.make_state_dict(spec, 0))
B, H, W = 8, 800, 1333
imgs = torch.stack([synth.synthetic_image(H, W, seed=100 + i) for i in range(B)]).cuda()
sess = eng.session(B, H, W, False)
for _ in range(3):
    sess.run(imgs)
torch.cuda.synchronize()
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
print("power limit W:", pynvml.nvmlDeviceGetPowerManagementLimit(h) / 1000.0, "enforced:", pynvml.nvmlDeviceGetEnforcedPowerLimit(h) / 1000.0, flush=True)
samples, stop = [], False


def loop():
    while not stop:
        samples.append((time.perf_counter(), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
                        pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0, pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)))
        time.sleep(0.02)


t = threading.Thread(target=loop, daemon=True)
t.start()
evs = [torch.cuda.Event(enable_timing=True) for _ in range(N // 10 + 1)]
t0 = time.perf_counter()
evs[0].record()
for i in range(N):
    sess.run(imgs)
    if (i + 1) % 10 == 0:
        evs[(i + 1) // 10].record()
torch.cuda.synchronize()
stop = True
t.join()
per10 = [evs[i].elapsed_time(evs[i + 1]) / 10 for i in range(N // 10)]
print("ms/step per block of 10:", [round(x, 2) for x in per10], flush=True)
tail = per10[len(per10) // 2:]
print(f"steady state (second half): {sum(tail) / len(tail):.2f} ms/step = {B / (sum(tail) / len(tail)) * 1e3:.1f} images/s")
sm = sorted(s[1] for s in samples)
pw = sorted(s[2] for s in samples)
cap = sum(1 for s in samples if s[3] & pynvml.nvmlClocksEventReasonSwPowerCap)
print(f"{len(samples)} samples: SM clock median {sm[len(sm) // 2]} MHz (min {sm[0]}, max {sm[-1]}); power median {pw[len(pw) // 2]:.0f} W, max {pw[-1]:.0f} W; "
      f"sw_power_cap in {cap}/{len(samples)} samples")
