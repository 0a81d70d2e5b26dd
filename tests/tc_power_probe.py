"""Profiling aid (not a test): filter-kernel loop for a few seconds with nvidia-smi power / clock samples beside it.
usage: python tests/tc_power_probe.py <mode> [seconds]"""
import subprocess, sys, time
import numpy as np
sys.path.insert(0, '.')
import reindexer_b200 as rx
from oracle import oracle as O

mode = int(sys.argv[1])
secs = float(sys.argv[2]) if len(sys.argv) > 2 else 3.0
rows, dim, nq, k = 10_000_000, 768, 1024, 10
gpu = rx.GpuBruteforceSearch(rx.IP, dim, rows)
gpu.append_synth(1, 0, rows)
q = O.synth_matrix(2, nq, dim)
gpu.set_tensor_core_filter(mode)
gpu.search_knn(q, k)
mon = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,enforced.power.limit,clocks_throttle_reasons.sw_power_cap,temperature.gpu",
                        "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
time.sleep(0.3)
t0 = time.time(); n = 0
rx.set_profile(1) if hasattr(rx, "set_profile") else None
while time.time() - t0 < secs:
    gpu.search_knn(q, k); n += 1
el = time.time() - t0
mon.terminate()
out = mon.communicate()[0].strip().splitlines()
print("mode", mode, "steps", n, "ms/step", 1000 * el / n, "q/s", n * nq / el)
vals = [l.split(", ") for l in out]
tail = vals[len(vals) // 3:]
print("samples", len(vals), "clock MHz", [v[0] for v in tail][:12], "\npower W", [v[1] for v in tail][:12], "limit", vals[-1][2], "sw_power_cap", [v[3] for v in tail][:12], "temp", vals[-1][4])
