import sys, time, numpy as np
sys.path.insert(0, '.')
import reindexer_b200 as rx
from oracle import oracle as O
n, dim, nq, k = 20000, 128, 64, 10
gpu = rx.GpuBruteforceSearch(rx.IP, dim, n)
gpu.append_synth(1, 0, n)
q = O.synth_matrix(2, nq, dim)
gpu.set_tensor_core_filter(2)
d0, l0, c0 = gpu.search_knn(q, k)
print("exact ok", l0[0] >> np.uint64(32), flush=True)
gpu.set_tensor_core_filter(1)
t = time.time()
d1, l1, c1 = gpu.search_knn(q, k)
print("tc returned in", time.time() - t, rx.last_search_stats(), flush=True)
print("labels equal:", (l0 == l1).all(), "dist bits equal:", (d0.view(np.uint32) == d1.view(np.uint32)).all(), flush=True)
bad = np.argwhere(l0 != l1)
print("mismatch count", len(bad), bad[:10].tolist(), flush=True)
if len(bad):
    i = bad[0][0]; print(l0[i] >> np.uint64(32), l1[i] >> np.uint64(32), d0[i], d1[i])
