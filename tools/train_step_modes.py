"""bench.train_step in both linear modes, fresh process, with allocator statistics and NVML clock / power samples.
Development tool."""
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
steps = int(os.environ.get("STEPS", "2"))
import pynvml
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)


class Sampler(threading.Thread):
    def __init__(self):
        super().__init__(daemon=True)
        self.stop = False
        self.samples = []

    def run(self):
        while not self.stop:
            self.samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1000,
                                 pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)))
            time.sleep(0.02)


for mode in sys.argv[1:] or ["fp32", "tf32x3", "fp32", "tf32x3"]:
    s = Sampler()
    s.start()
    r = bench.train_step(dev, 0, 1, linear_mode=mode, steps=steps)
    s.stop = True
    s.join()
    clk = sorted(c for c, _, _ in s.samples)
    pw = sorted(p for _, p, _ in s.samples)
    reasons = 0
    for _, _, rs in s.samples:
        reasons |= rs
    print(mode, r["ms_per_optimizer_step"], "ms", r["episodes_per_s"], "episodes/s", "mallocs in timed region", r["cudaMallocs_in_timed_region"],
          f"sm MHz min/med/max {clk[0]}/{clk[len(clk) // 2]}/{clk[-1]}  power W med/max {pw[len(pw) // 2]:.0f}/{pw[-1]:.0f}  throttle reasons {reasons:#x}",
          flush=True)
