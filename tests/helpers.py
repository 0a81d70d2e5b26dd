"""Shared helpers for the parity tests (test infrastructure; may use oracle/)."""
import numpy as np

from oracle import oracle as O

RTOL = 1e-4  # BASELINE.json north_star: fp distances within 1e-4 relative
ATOL = 2e-6


def prep_query(metric, q, use_ref=None):
    return O.normalize_copy(q, use_ref)[0] if metric == O.COS else np.ascontiguousarray(q, np.float32)


def assert_same_knn(d_gpu, l_gpu, d_ref, l_ref, ctx=""):
    """Distances within RTOL; ids exact wherever neighbouring reference distances are separated by more than the fp noise
    (the reference itself is not bit-reproducible across its SSE/AVX/AVX-512 kernels, SURVEY.md §8a rule 6).  Inside a group
    of near-equal distances the order may differ; the last group may also trade members with rows just outside the top-k."""
    d_gpu, d_ref = np.asarray(d_gpu), np.asarray(d_ref)
    l_gpu, l_ref = np.asarray(l_gpu), np.asarray(l_ref)
    assert len(d_gpu) == len(d_ref), (ctx, len(d_gpu), len(d_ref))
    assert np.allclose(d_gpu, d_ref, rtol=RTOL, atol=ATOL), (ctx, d_gpu, d_ref)
    if (l_gpu == l_ref).all():
        return
    noise = RTOL * np.maximum(np.abs(d_ref), 1e-2)
    n = len(d_ref)
    i = 0
    while i < n:
        j = i
        while j + 1 < n and d_ref[j + 1] - d_ref[j] <= noise[j]:
            j += 1
        if j + 1 < n:
            assert set(l_ref[i:j + 1].tolist()) == set(l_gpu[i:j + 1].tolist()), (ctx, i, j, l_gpu, l_ref, d_ref)
        i = j + 1


def numpy_dists(metric, q, vecs, norm_coefs=None):
    """fp64 distances in map space (for property checks, not for bit parity)."""
    q64, v64 = q.astype(np.float64), vecs.astype(np.float64)
    if metric == O.L2:
        return ((v64 - q64) ** 2).sum(1)
    d = -(v64 @ q64)
    if metric == O.COS:
        d = d / np.sqrt((v64 ** 2).sum(1))
    return d
