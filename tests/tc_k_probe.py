"""Profiling aid (not a test): config-1 index, 1024 queries, larger k -- the tensor-core filter path against the exact scan.
usage: python tests/tc_k_probe.py k [k ...]"""
import sys, time
import numpy as np
sys.path.insert(0, '.')
import reindexer_b200 as rx
from oracle import oracle as O

rows, dim, nq = 10_000_000, 768, 1024
gpu = rx.GpuBruteforceSearch(rx.IP, dim, rows)
gpu.append_synth(1, 0, rows)
q = O.synth_matrix(2, nq, dim)
for k in [int(x) for x in sys.argv[1:]]:
    gpu.set_tensor_core_filter(0)
    gpu.search_knn(q, k)
    t0 = time.time()
    d1, l1, c1 = gpu.search_knn(q, k)
    t_tc = time.time() - t0
    st = rx.last_search_stats()
    gpu.set_tensor_core_filter(2)
    t0 = time.time()
    d0, l0, c0 = gpu.search_knn(q[:64], k)
    t_ex = (time.time() - t0) * nq / 64
    same = (l0 == l1[:64]).all() and (d0.view(np.uint32) == d1[:64].view(np.uint32)).all()
    print(f"k={k}: filter path {nq / t_tc:.0f} queries/s ({t_tc * 1e3:.1f} ms per batch, tc_used={st['tc_used']}, candidates/query={st['tc_candidates'] / nq:.0f}, "
          f"fallbacks={st['tc_fallbacks']}); exact scan {nq / t_ex:.0f} queries/s (64-query sample); identical bits: {same}", flush=True)
