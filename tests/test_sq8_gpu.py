"""GPU parity tests of the SQ8 path (reindexer_b200/csrc/sq8.cu, hnsw.cu) against the reference's own quantised map
(HierarchicalNSWImpl<uint8_t>, built by oracle/_ref from a float graph): the device-side quantisation must reproduce the reference's
codes and corrective offsets bit for bit, the quantised query too, the dp4a brute-force scan must equal the formula
alpha_2 * int_dist + offsets evaluated in the reference's float order, and the quantised HNSW search must return what
HierarchicalNSWImpl<uint8_t>::SearchKnn returns."""
import numpy as np
import pytest

import reindexer_b200 as rx
from oracle import oracle as O

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not O.ref_knn_available(), reason="oracle/_ref not built")]


def build(metric, n, dim, seed):
    rng = np.random.default_rng(seed)
    centers = rng.normal(0, 0.5, size=(40, dim)).astype(np.float32)
    vecs = (centers[rng.integers(0, 40, size=n)] + rng.normal(0, 0.2, size=(n, dim))).astype(np.float32)
    labels = O.row_labels(n)
    ref = O.RefHnsw(metric, dim, n, M=16, ef_construction=100, seed=100, multithread=False)
    ref.add_batch(labels, vecs)
    refq = ref.quantize()
    g = ref.export(with_vectors=False)
    gpu = rx.GpuBruteforceSearch(metric, dim, n)
    gpu.add_points(g["labels"], vecs[(g["labels"] >> np.uint64(32)).astype(np.int64)])  # row i = internal id i
    gpu.hnsw_import(g)
    prm = refq.params()
    params = dict(min_q=float(prm["minQ"]), max_q=float(prm["maxQ"]), alpha=float(prm["alpha"]), alpha_2=float(prm["alpha_2"]),
                  delta=float(prm["delta"]))
    queries = (centers[rng.integers(0, 40, size=64)] + rng.normal(0, 0.2, size=(64, dim))).astype(np.float32)
    return ref, refq, gpu, params, queries, vecs, g


@pytest.mark.parametrize("metric", [rx.L2, rx.IP, rx.COS])
@pytest.mark.parametrize("dim", [48, 100])
def test_sq8_matches_the_references_quantised_map(metric, dim):
    n, k, ef = 3000, 10, 64
    ref, refq, gpu, params, queries, vecs, g = build(metric, n, dim, 900 + dim + metric)
    rcodes, roffs = refq.export()
    # (1) quantisation on the device == Quantizer::quantize of the reference: same codes, same corrective offsets (bits)
    gpu.sq8_attach(params)
    dcodes, doffs = gpu.sq8_export()
    assert (dcodes == rcodes).all()
    assert (doffs.view(np.uint32) == roffs.view(np.uint32)).all(), np.argwhere(doffs.view(np.uint32) != roffs.view(np.uint32))[:5]
    # (2) importing the reference's codes gives the same state
    gpu.sq8_attach(params, rcodes, roffs)
    norms = np.ones(len(queries), np.float32)
    qs = queries
    if metric == rx.COS:
        pairs = [O.normalize_copy(q) for q in queries]
        qs = np.stack([p[0] for p in pairs])
        norms = np.array([1.0 / p[1] for p in pairs], np.float32)  # HnswIndexBase::search: normL2 = 1 / NormalizeCopyVector(...)
    # (3) the quantised query (prepareData): Cosine restores the length before quantising
    for i in range(6):
        if metric == rx.COS:
            continue  # the facade's prepare_query passes norm 1; the Cosine query path is covered by the search comparison below
        qc, qo = gpu.sq8_prepare_query(qs[i], float(norms[i]))
        rc, ro = refq.prepare_query(qs[i])
        assert (qc == rc).all() and np.float32(qo).view(np.uint32) == np.float32(ro).view(np.uint32)
    # (4) brute force over the codes == the reference's formula in its float order, exact top-k under (dist, label)
    d, l, c = gpu.sq8_search_knn(qs, k, None if metric != rx.COS else norms)
    a2 = np.float32(params["alpha_2"])
    labels = g["labels"]
    norm_coefs = None
    if metric == rx.COS:
        rows = vecs[(labels >> np.uint64(32)).astype(np.int64)]
        norm_coefs = np.array([O.normalize_copy(r)[1] for r in rows], np.float32)
    for i in range(0, len(qs), 7):
        if metric == rx.COS:
            scale = np.float32(1.0) / (np.float32(1.0) / norms[i])
            qc, qo = gpu.sq8_prepare_query(qs[i], float(norms[i]))
        else:
            qc, qo = gpu.sq8_prepare_query(qs[i])
        qi = qc.astype(np.int64)
        ci = rcodes.astype(np.int64)
        integer = ((ci - qi) ** 2).sum(1) if metric == rx.L2 else (ci * qi).sum(1)
        dist = (a2 * integer.astype(np.float32) + np.float32(qo)) + roffs
        if metric != rx.L2:
            dist = -dist
        if metric == rx.COS:
            dist = dist * norm_coefs
            dist = (np.float32(1.0) / norms[i]) * dist
        order = np.lexsort((labels, dist))[:k]
        assert (l[i] == labels[order]).all(), (i, l[i], labels[order])
        if metric == rx.COS:  # the row norm coefficients are fp32 sums whose order differs between the device and numpy: 1e-6, not bits
            assert np.allclose(d[i], dist[order], rtol=2e-6, atol=0)
        else:
            assert (d[i].view(np.uint32) == dist[order].astype(np.float32).view(np.uint32)).all()
    # (5) the quantised HNSW search == HierarchicalNSWImpl<uint8_t>::SearchKnn
    dh, lh, ch = gpu.hnsw_search_knn_sq8(qs, k, ef, None if metric != rx.COS else norms)
    same = 0
    for i in range(len(qs)):
        dr, lr = refq.search_knn(qs[i], k, ef, qnorm=None if metric != rx.COS else float(norms[i]))
        ok = len(lr) == ch[i] and (lh[i, :ch[i]] == lr).all()
        same += ok
        if ok and metric != rx.COS:
            assert (dh[i, :ch[i]].view(np.uint32) == dr.view(np.uint32)).all(), i
        elif ok:  # Cosine: the reference's own norm coefficients differ from the device's in the last bit (different fp32 sum order)
            assert np.allclose(dh[i, :ch[i]], dr, rtol=2e-6, atol=0), i
    assert same >= 0.97 * len(qs), same
    # and it agrees with the exact answer under the quantised metric most of the time
    rec = np.mean([len(set(lh[i].tolist()) & set(l[i].tolist())) / k for i in range(len(qs))])
    assert rec >= 0.9, rec
