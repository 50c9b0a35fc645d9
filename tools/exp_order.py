"""Developer experiment: does running the complex headline kernel first change the timing of the real one?"""
import os, sys, subprocess, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.argv = [sys.argv[0], "none"]
import bench_suite as bs
import torch
def clocks():
    return subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
for step in ("rfft:f", "cfft:f", "cfft:b", "rfft:f", "rfft:b", "rfft:f", "sleep", "rfft:f", "cfft:b", "rfft:f"):
    if step == "sleep":
        torch.cuda.synchronize(); time.sleep(2.0); print("slept 2 s"); continue
    bs.case(step.split(":")[0], 4096, 65536, d=step.split(":")[1])
    print("   ", clocks(), flush=True)
