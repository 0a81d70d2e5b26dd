"""Profiling aid (not a test): per-tile timestamps of CTA 0 of a filter-kernel variant at config-1 size.
usage: python tests/tc_trace_run.py <mode> [rows]   -> gpurun_out/trace_<mode>.txt + the analysis on stdout"""
import os, subprocess, sys
import numpy as np
sys.path.insert(0, '.')
import reindexer_b200 as rx
from oracle import oracle as O

mode = int(sys.argv[1])
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
dim, nq, k = 768, 1024, 10
gpu = rx.GpuBruteforceSearch(rx.IP, dim, rows)
gpu.append_synth(1, 0, rows)
q = O.synth_matrix(2, nq, dim)
gpu.set_tensor_core_filter(mode)
gpu.search_knn(q, k)
gpu.search_knn(q, k)
os.makedirs("gpurun_out", exist_ok=True)
path = f"gpurun_out/trace_{mode}.txt"
os.environ["RXGPU_TC_TRACE"] = path
os.environ["RXGPU_TC_TRACE_FIRST"] = "1000"
gpu.search_knn(q, k)
del os.environ["RXGPU_TC_TRACE"]
print("mode", mode, rx.last_search_stats())
subprocess.call([sys.executable, "tests/tc_trace_analyze.py", path])
