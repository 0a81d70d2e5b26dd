"""CPU test of the multi-GPU plumbing (world_size 2, gloo): ShardedBruteforceSearch's all-gather + merge + global tie replay, with
the per-shard device scans emulated in numpy, against the oracle running the reference's sequential algorithm over all rows."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_per, seed, out_q):
    sys.path.insert(0, ROOT)
    from reindexer_b200.sharded import ShardedBruteforceSearch

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(seed)
    n = n_per * world
    vals = rng.integers(-3, 4, size=n).astype(np.float32)          # 1-D integer vectors: exact distances, ties everywhere
    labels = (rng.permutation(n).astype(np.uint64) << np.uint64(32))
    lo, hi = rank * n_per, (rank + 1) * n_per
    my_vals, my_labels = vals[lo:hi], labels[lo:hi]

    def local_search(d_queries, k1):  # what rxgpu_search_knn_device returns for this shard
        q = d_queries.numpy()
        nq = q.shape[0]
        od, oi = torch.zeros((nq, k1), dtype=torch.float32), torch.zeros((nq, k1), dtype=torch.int32)
        ol, oc = torch.zeros((nq, k1), dtype=torch.int64), torch.zeros((nq,), dtype=torch.int32)
        for i in range(nq):
            d = (my_vals - q[i, 0]) ** 2
            order = np.lexsort((np.arange(len(d)), d))[:k1]
            c = len(order)
            od[i, :c] = torch.from_numpy(d[order])
            oi[i, :c] = torch.from_numpy(order.astype(np.int32))
            ol[i, :c] = torch.from_numpy(my_labels[order].view(np.int64))
            oc[i] = c
        return od, oi, ol, oc

    def local_tie_rows(d_query, dstar, k):
        d = (my_vals - float(d_query[0])) ** 2
        rows = np.nonzero(d <= dstar)[0][:k]
        od, oi = torch.zeros((k,), dtype=torch.float32), torch.zeros((k,), dtype=torch.int32)
        ol, oc = torch.zeros((k,), dtype=torch.int64), torch.zeros((1,), dtype=torch.int32)
        c = len(rows)
        od[:c], oi[:c] = torch.from_numpy(d[rows]), torch.from_numpy(rows.astype(np.int32))
        ol[:c] = torch.from_numpy(my_labels[rows].view(np.int64))
        oc[0] = c
        return od, oi, ol, oc

    sh = ShardedBruteforceSearch(None, n_per, device=torch.device("cpu"), local_search=local_search, local_tie_rows=local_tie_rows)
    queries = rng.integers(-3, 4, size=(9, 1)).astype(np.float32)
    results = {}
    for k in (1, 4, 10, n + 5):
        d, l, c = sh.search_knn(queries, k)
        results[k] = (d, l, c)
    if rank == 0:
        out_q.put((vals, labels, queries, results))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_search_world2_matches_sequential_reference():
    sys.path.insert(0, ROOT)
    from oracle import oracle as O

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    n_per = 23
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_per, 1234, q)) for r in range(2)]
    for p in procs:
        p.start()
    vals, labels, queries, results = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    bf = O.PortBF(O.L2, 1, len(vals))
    assert bf.add_batch(labels, vals.reshape(-1, 1)) == 0
    for k, (d, l, c) in results.items():
        for i in range(len(queries)):
            dr, lr = bf.search_knn(queries[i], k)
            assert c[i] == len(dr)
            assert (l[i, :c[i]] == lr).all(), (k, i, l[i, :c[i]], lr)
            assert (d[i, :c[i]] == dr).all()


def _hnsw_worker(rank, world, port, n_per, seed, out_q):
    sys.path.insert(0, ROOT)
    from reindexer_b200.sharded import ShardedHnswSearch

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(seed)
    n = n_per * world
    vecs = rng.normal(size=(n, 8)).astype(np.float32)
    labels = (rng.permutation(n).astype(np.uint64) << np.uint64(32))
    lo = rank * n_per
    my_vecs, my_labels = vecs[lo:lo + n_per], labels[lo:lo + n_per]
    # an APPROXIMATE local search (a graph search misses some rows): this shard only "sees" the rows with an even local index
    seen = np.arange(0, n_per, 2)

    def local_search(d_queries, k1):
        q = d_queries.numpy()
        nq = q.shape[0]
        od, oi = torch.zeros((nq, k1), dtype=torch.float32), torch.zeros((nq, k1), dtype=torch.int32)
        ol, oc = torch.zeros((nq, k1), dtype=torch.int64), torch.zeros((nq,), dtype=torch.int32)
        for i in range(nq):
            d = ((my_vecs[seen] - q[i]) ** 2).sum(axis=1).astype(np.float32)
            order = np.argsort(d, kind="stable")[:k1]
            c = len(order)
            od[i, :c] = torch.from_numpy(d[order])
            oi[i, :c] = torch.from_numpy(seen[order].astype(np.int32))
            ol[i, :c] = torch.from_numpy(my_labels[seen[order]].view(np.int64))
            oc[i] = c
        return od, oi, ol, oc

    sh = ShardedHnswSearch(None, n_per, ef=32, device=torch.device("cpu"), local_search=local_search)
    queries = rng.normal(size=(7, 8)).astype(np.float32)
    res = {k: sh.search_knn(queries, k) for k in (1, 5, 10)}
    if rank == 0:
        out_q.put((vecs, labels, queries, res))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_hnsw_merge_world2():
    """multi-GPU HNSW = independent per-shard searches + the same one-all-gather merge: the result is the global top-k of the union of
    what the shards returned (no tie replay)"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    n_per = 40
    procs = [ctx.Process(target=_hnsw_worker, args=(r, 2, port, n_per, 77, q)) for r in range(2)]
    for p in procs:
        p.start()
    vecs, labels, queries, res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    visible = np.concatenate([np.arange(0, n_per, 2), n_per + np.arange(0, n_per, 2)])
    for k, (d, l, c) in res.items():
        for i in range(len(queries)):
            dist_all = ((vecs[visible] - queries[i]) ** 2).sum(axis=1).astype(np.float32)
            order = np.argsort(dist_all, kind="stable")[:k]
            assert c[i] == k
            assert (l[i, :k] == labels[visible[order]]).all()
            assert np.allclose(d[i, :k], dist_all[order])
