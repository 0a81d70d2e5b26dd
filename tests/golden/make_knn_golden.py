"""Generates tests/golden/knn_golden.npz from the REFERENCE's own code (oracle/_ref, built in place from /root/reference).

Run in the authoring container only:  python tests/golden/make_knn_golden.py
The reference's tests hold no golden vectors for the KNN path (property checks only), so these fixtures are
outputs of hnswlib::BruteforceSearch::SearchKnn / SearchRange themselves on seeded inputs.  Inputs that are not a pure
function of (seed, shape) -- the tie-heavy integer cases -- are stored in the file next to the outputs.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

assert O.ref_knn_available(), "build oracle/_ref first (make -C oracle ref)"

CASES = [  # (name, metric, n, dim, k, nq, seed)
    ("l2_small", O.L2, 2000, 128, 10, 8, 0x5EED0000),
    ("ip_small", O.IP, 2000, 96, 10, 8, 0x5EED0010),
    ("cos_small", O.COS, 2000, 100, 10, 8, 0x5EED0020),
    ("l2_odd_dim", O.L2, 1500, 17, 7, 8, 0x5EED0030),
    ("ip_768", O.IP, 3000, 768, 10, 8, 0x5EED0040),
    ("cos_k_gt_n", O.COS, 20, 32, 50, 4, 0x5EED0050),
]
out = {}
for name, metric, n, dim, k, nq, seed in CASES:
    vecs = O.synth_matrix(seed, n, dim)
    queries = O.synth_matrix(seed + 1, nq, dim)
    labels = O.row_labels(n)
    bf = O.RefBF(metric, dim, n)
    bf.add_batch(labels, vecs)
    D = np.zeros((nq, k), np.float32)
    L = np.zeros((nq, k), np.uint64)
    Cn = np.zeros(nq, np.uint32)
    for i in range(nq):
        q = queries[i] if metric != O.COS else O.normalize_copy(queries[i], True)[0]
        d, l = bf.search_knn(q, k)
        D[i, :len(d)], L[i, :len(d)], Cn[i] = d, l, len(d)
    out[f"{name}/meta"] = np.array([metric, n, dim, k, nq, seed], np.int64)
    out[f"{name}/dist"], out[f"{name}/label"], out[f"{name}/count"] = D, L, Cn
    # one range query per case: radius halfway between the 5th and 6th neighbour of query 0 (robust against fp noise)
    q0 = queries[0] if metric != O.COS else O.normalize_copy(queries[0], True)[0]
    radius = float((D[0, 4] + D[0, 5]) / 2) if Cn[0] > 5 else float(D[0, Cn[0] - 1] + 1.0)
    rd, rl = bf.search_range(q0, radius)
    out[f"{name}/range_radius"] = np.float32(radius)
    out[f"{name}/range_dist"], out[f"{name}/range_label"] = rd, rl

# tie-heavy integer-valued cases: every summation order gives the same fp32 sums, so ids AND distances are exact
rng = np.random.default_rng(20260922)
for name, metric, n, dim, k, nq, removes in [("tie_l2", O.L2, 1500, 8, 16, 12, 100), ("tie_ip", O.IP, 1500, 8, 16, 12, 0),
                                              ("dup_rows_cos", O.COS, 1200, 16, 9, 8, 50)]:
    if name.startswith("dup"):
        base = rng.integers(-3, 4, size=(n // 6, dim)).astype(np.float32)
        base[np.abs(base).sum(1) == 0, 0] = 1
        vecs = np.repeat(base, 6, axis=0)
        labels = ((np.repeat(np.arange(n // 6), 6).astype(np.uint64) << np.uint64(32)) | np.tile(np.arange(6), n // 6).astype(np.uint64))
    else:
        vecs = rng.integers(-2, 3, size=(n, dim)).astype(np.float32)
        labels = O.row_labels(n)
    perm = rng.permutation(n)
    vecs, labels = vecs[perm], labels[perm]
    queries = rng.integers(-2, 3, size=(nq, dim)).astype(np.float32)
    queries[np.abs(queries).sum(1) == 0, 0] = 1
    rm = labels[rng.choice(n, removes, replace=False)] if removes else np.zeros(0, np.uint64)
    bf = O.RefBF(metric, dim, n)
    bf.add_batch(labels, vecs)
    for l in rm:
        bf.remove(int(l))
    D = np.zeros((nq, k), np.float32)
    L = np.zeros((nq, k), np.uint64)
    for i in range(nq):
        q = queries[i] if metric != O.COS else O.normalize_copy(queries[i], True)[0]
        d, l = bf.search_knn(q, k)
        assert len(d) == k
        D[i], L[i] = d, l
    out[f"{name}/meta"] = np.array([metric, n, dim, k, nq, 0], np.int64)
    out[f"{name}/vecs"], out[f"{name}/labels"], out[f"{name}/queries"], out[f"{name}/removes"] = vecs, labels, queries, rm
    out[f"{name}/dist"], out[f"{name}/label"] = D, L

np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "knn_golden.npz"), **out)
print("wrote", len(out), "arrays")
