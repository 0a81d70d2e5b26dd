"""CPU tests of the product's host-side logic (no GPU needed): shard merge + the closed-form replay of the reference's heap
tie rule (reindexer_b200/host/knn_select.h via rxgpu_merge_shards / rxgpu_tie_replay), checked against the oracle running the
reference's sequential algorithm on 1-D integer vectors (distance = (v - q)^2 exactly, so ties are everywhere)."""
import numpy as np
import pytest

import reindexer_b200 as rx
from oracle import oracle as O


def emulate_device_topk(dist, labels, k1):
    """What rxgpu_search_knn_device returns for one shard: top-k1 under (dist, internal idx)."""
    order = np.lexsort((np.arange(len(dist)), dist))[:k1]
    return dist[order], order.astype(np.uint32), labels[order], len(order)


def emulate_tie_rows(dist, labels, dstar, k):
    rows = np.nonzero(dist <= dstar)[0][:k]
    return dist[rows], rows.astype(np.uint32), labels[rows], len(rows)


def product_knn(dist, labels, k, nshards):
    """Full host pipeline on emulated device outputs; dist/labels in global internal order."""
    n = len(dist)
    k_eff = min(k, n)
    bounds = np.linspace(0, n, nshards + 1).astype(np.int64)
    k1 = k_eff + 1
    D = np.zeros((nshards, 1, k1), np.float32)
    I = np.zeros((nshards, 1, k1), np.uint32)
    L = np.zeros((nshards, 1, k1), np.uint64)
    Cn = np.zeros((nshards, 1), np.uint32)
    for s in range(nshards):
        sl = slice(bounds[s], bounds[s + 1])
        d, i, l, c = emulate_device_topk(dist[sl], labels[sl], k1)
        D[s, 0, :c], I[s, 0, :c], L[s, 0, :c], Cn[s, 0] = d, i, l, c
    od, og, ol, oc, nt = rx.merge_shards(k_eff, D, I, L, Cn, bounds[:-1].astype(np.uint64))
    c = int(oc[0])
    if not nt[0]:
        return od[0, :c], ol[0, :c], False
    dstar = od[0, c - 1]
    lower = od[0, :c] < dstar
    fd, fg, fl = [], [], []
    for s in range(nshards):
        sl = slice(bounds[s], bounds[s + 1])
        d, i, l, cc = emulate_tie_rows(dist[sl], labels[sl], dstar, k_eff)
        fd += d.tolist()
        fg += (i.astype(np.uint64) + np.uint64(bounds[s])).tolist()
        fl += l.tolist()
    o = np.argsort(np.array(fg, np.uint64), kind="stable")[:k_eff]
    first = (np.array(fd, np.float32)[o], np.array(fg, np.uint64)[o], np.array(fl, np.uint64)[o])
    rd, rl = rx.tie_replay(k_eff, float(dstar), (od[0, :c][lower], og[0, :c][lower], ol[0, :c][lower]), first)
    return rd, rl, True


@pytest.mark.parametrize("nshards", [1, 2, 3, 8])
def test_tie_rule_matches_sequential_reference(nshards):
    rng = np.random.default_rng(100 + nshards)
    replays = 0
    for trial in range(300):
        n = int(rng.integers(1, 60))
        k = int(rng.integers(1, 14))
        vals = rng.integers(-3, 4, size=n).astype(np.float32)
        labels = rng.permutation(n * 3)[:n].astype(np.uint64) << np.uint64(32) | rng.integers(0, 3, size=n).astype(np.uint64)
        labels = np.unique(labels)
        rng.shuffle(labels)
        n = len(labels)
        vals = vals[:n]
        bf = O.PortBF(O.L2, 1, n)
        assert bf.add_batch(labels, vals.reshape(-1, 1)) == 0
        q = np.array([float(rng.integers(-3, 4))], np.float32)
        d_ref, l_ref = bf.search_knn(q, k)
        dist = (vals - q[0]) ** 2
        d, l, replayed = product_knn(dist.astype(np.float32), labels, k, min(nshards, n))
        replays += replayed
        assert len(d) == len(d_ref)
        assert (l == l_ref).all(), (trial, n, k, dist, labels, l, l_ref)
        assert (d == d_ref).all()
    assert replays > 50  # the tie path was really exercised


def test_merge_orders_equal_distances_by_label():
    # two rows with equal distance inside the top-k (not at the boundary): device order is by index, reference drains by label
    D = np.array([[[1.0, 1.0, 2.0, 5.0]]], np.float32)
    I = np.array([[[0, 1, 2, 3]]], np.uint32)
    L = np.array([[[9 << 32, 4 << 32, 7 << 32, 1 << 32]]], np.uint64)
    od, og, ol, oc, nt = rx.merge_shards(3, D, I, L, np.array([[4]], np.uint32), np.array([0], np.uint64))
    assert oc[0] == 3 and not nt[0]
    assert ol[0].tolist() == [4 << 32, 9 << 32, 7 << 32]
    assert og[0].tolist() == [1, 0, 2]
