#!/usr/bin/env python
"""Measurements of the other two hot-path rows (not the driver's bench line; see bench.py for that):

  python bench_extra.py hnsw [--rows 500000]   BASELINE configs[2] shape: HNSW M=16 efC=200, 768-dim, Cosine, ef=128, k=10
  python bench_extra.py ft   [--docs 50000000] BASELINE configs[3]: ft_fast BM25, 3-term OR (df 10% / 1% / 0.1%), top-100

Each prints one JSON line with our device number, the reference's own CPU code timed on this box (oracle/_ref) and the
algorithmic-bytes roofline figure of SURVEY.md §8d.  The HNSW graph is built by the reference's CPU code (multithreaded insert);
building 10M x 768 takes hours, so the default is 500k rows -- stated in the output.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return (json.load(open(p))["hbm_gbs"], "measured") if os.path.exists(p) else (6650.0, "fallback")


def lowrank(seed, n, dim, latent=32, noise=0.05):
    a = np.random.default_rng(99).normal(0, 1.0, size=(latent, dim)).astype(np.float32)
    rng = np.random.default_rng(seed)
    out = np.empty((n, dim), np.float32)
    for i in range(0, n, 100000):
        m = min(100000, n - i)
        out[i:i + m] = rng.normal(0, 1, size=(m, latent)).astype(np.float32) @ a + rng.normal(0, noise, size=(m, dim)).astype(np.float32)
    return out


def hnsw_record(rows, queries=4096, threads=None, sweep=False):
    """BASELINE configs[2] shape at `rows` rows: returns the sub-record bench.py embeds (workload, value, e2e, roofline, cpu_baseline)."""
    import reindexer_b200 as rx
    from oracle import oracle as O

    class A:
        pass

    args = A()
    args.rows, args.queries, args.sweep = rows, queries, sweep
    n, dim, k, ef, nq = args.rows, 768, 10, 128, args.queries
    threads = threads or (len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    vecs, labels = lowrank(1, n, dim), O.row_labels(n)
    t0 = time.perf_counter()
    ref = O.RefHnsw(O.COS, dim, n, M=16, ef_construction=200, seed=100, multithread=True)
    ref.add_batch(labels, vecs, threads=threads)
    build_s = time.perf_counter() - t0
    g = ref.export(with_vectors=False)
    gpu = rx.GpuBruteforceSearch(rx.COS, dim, n)  # multithreaded insert: internal id != insertion order, so rows follow the graph
    gpu.add_points(g["labels"], vecs[(g["labels"] >> np.uint64(32)).astype(np.int64)])  # row i of the index = internal id i
    gpu.hnsw_import(g)
    queries = np.stack([O.normalize_copy(q)[0] for q in lowrank(2, nq, dim)])
    import torch

    dq = torch.from_numpy(queries).cuda()
    od = torch.zeros((nq, k), dtype=torch.float32, device="cuda")
    oi = torch.zeros((nq, k), dtype=torch.int32, device="cuda")
    oc = torch.zeros((nq,), dtype=torch.int32, device="cuda")

    def device_qps(reps=5):  # queries and results resident in HBM, CUDA events around the calls
        for _ in range(2):
            gpu.hnsw_search_knn_device(nq, dq.data_ptr(), k, ef, od.data_ptr(), oi.data_ptr(), oc.data_ptr())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            gpu.hnsw_search_knn_device(nq, dq.data_ptr(), k, ef, od.data_ptr(), oi.data_ptr(), oc.data_ptr())
        e1.record()
        torch.cuda.synchronize()
        return reps * nq / (e0.elapsed_time(e1) / 1e3)

    sweep = {}
    if args.sweep:  # resident CTAs per SM (RXGPU_HNSW_CTAS_PER_SM is read at import time)
        for c in (2, 4, 6, 8, 12, 16):
            os.environ["RXGPU_HNSW_CTAS_PER_SM"] = str(c)
            gpu.hnsw_import(g)
            sweep[c] = device_qps()
        del os.environ["RXGPU_HNSW_CTAS_PER_SM"]
        gpu.hnsw_import(g)
    qps_device = device_qps()
    gpu.hnsw_search_knn(queries[:64], k, ef)  # warm-up
    reps = 5
    t0 = time.perf_counter()
    for _ in range(reps):
        d, l, c, st = gpu.hnsw_search_knn(queries, k, ef, with_stats=True)
    gpu_s = (time.perf_counter() - t0) / reps
    t0 = time.perf_counter()
    dr, lr, cr = ref.search_knn_batch(queries, k, ef, threads=threads)
    cpu_s = time.perf_counter() - t0
    db, lb, _ = gpu.search_knn(queries[:512], k)
    rec_gpu = float(np.mean([len(set(l[i]) & set(lb[i])) / k for i in range(512)]))
    rec_ref = float(np.mean([len(set(lr[i]) & set(lb[i])) / k for i in range(512)]))
    same = float(np.mean([(l[i] == lr[i]).all() for i in range(nq)]))
    ndist, nhops = float(st[:, 0].mean()), float(st[:, 1].mean())
    bytes_per_query = ndist * (dim * 4 + 4) + nhops * (4 + 8 * 16)
    peak, src = peak_hbm()
    gpu.close()
    return {
        "workload": f"HNSW float_vector, {n} x {dim} fp32, cosine, M=16 efC=200, ef_search={ef}, k={k}, batch={nq} (BASELINE configs[2] shape at "
                    f"{n} rows: the graph is built by the reference's CPU inserter, 10M rows would need hours)",
        "metric": "HNSW KNN queries/s", "value": qps_device, "unit": "queries/s", "higher_is_better": True,
        "e2e": {"value": nq / gpu_s, "unit": "queries/s", "h2d_bytes_per_step": nq * dim * 4, "d2h_bytes_per_step": nq * k * 12 + nq * 12},
        "qps_gpu_device_resident": qps_device, "qps_gpu_e2e": nq / gpu_s, "sweep_ctas_per_sm": sweep,
        "cpu_baseline": {"value": nq / cpu_s, "unit": "queries/s", "cores": threads, "kind": "reference",
                         "sample": f"{nq} queries, {threads} threads, HierarchicalNSW::SearchKnn on the same graph"},
        "speedup_e2e": cpu_s / gpu_s,
        "recall_at_10_gpu": rec_gpu, "recall_at_10_reference": rec_ref, "identical_top10_fraction": same,
        "dist_evals_per_query": ndist, "hops_per_query": nhops, "algorithmic_bytes_per_query": bytes_per_query,
        "roofline": {"bound": "hbm (random 3 KB row gathers)", "kernel": "hnsw_search_kernel", "achieved": bytes_per_query * qps_device / 1e9,
                     "peak": peak, "unit": "GB/s", "frac": bytes_per_query * qps_device / 1e9 / peak, "peak_source": src,
                     "note": "algorithmic gather bytes; at this row count part of the set is L2-resident", "traffic": None},
        "graph_build_s_reference_cpu": build_s, "data": "synthetic low-rank (latent 32) vectors"}


def run_hnsw(args):
    print(json.dumps(hnsw_record(args.rows, args.queries, sweep=args.sweep)))


def run_ivf(args):
    import reindexer_b200 as rx
    from oracle import oracle as O

    n, dim, nlist, nprobe, k, nq = args.rows, args.dim, args.nlist, args.nprobe, 10, args.queries
    vecs, labels = lowrank(11, n, dim, latent=min(32, dim)), O.row_labels(n)
    t0 = time.perf_counter()
    ref = O.RefIvf(O.L2, dim, nlist)
    ref.train_add(labels, vecs)
    build_s = time.perf_counter() - t0
    st = ref.export()
    gpu = rx.GpuBruteforceSearch(rx.L2, dim, n)
    gpu.add_points(st["labels"], st["vecs"])
    gpu.ivf_import(st["centroids"], st["list_sizes"])
    queries = lowrank(12, nq, dim, latent=min(32, dim))
    gpu.ivf_search_knn(queries[:64], k, nprobe)
    reps = 5
    t0 = time.perf_counter()
    for _ in range(reps):
        d, l, c = gpu.ivf_search_knn(queries, k, nprobe)
    gpu_s = (time.perf_counter() - t0) / reps
    ref.search_batch(queries[:64], k, nprobe)
    t0 = time.perf_counter()
    dr, lr = ref.search_batch(queries, k, nprobe)
    cpu_s = time.perf_counter() - t0
    same = float(np.mean([(l[i] == lr[i]).all() for i in range(nq)]))
    db, lb, _ = gpu.search_knn(queries[:256], k)
    recall = float(np.mean([len(set(l[i]) & set(lb[i])) / k for i in range(256)]))
    rows_scanned = float(np.sort(st["list_sizes"].astype(np.float64))[::-1][:nprobe].sum())  # upper bound; the mean is n * nprobe / nlist
    bytes_per_query = (n * nprobe / nlist) * dim * 4 + nlist * dim * 4
    peak, src = peak_hbm()
    print(json.dumps({
        "workload": f"IVF flat, {n} x {dim} fp32, L2, nlist={nlist}, nprobe={nprobe}, k={k}, batch={nq}; lists trained by the reference's FAISS",
        "qps_gpu_e2e": nq / gpu_s, "qps_reference_faiss_cpu": nq / cpu_s, "cpu_threads": os.cpu_count(), "speedup": cpu_s / gpu_s,
        "identical_top10_fraction": same, "recall_at_10_vs_exact": recall, "largest_probe_rows": rows_scanned,
        "algorithmic_bytes_per_query": bytes_per_query,
        "roofline": {"bound": "hbm / L2 (list scans)", "achieved": bytes_per_query * nq / gpu_s / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": bytes_per_query * nq / gpu_s / 1e9 / peak, "peak_source": src},
        "train_add_s_reference_cpu": build_s, "data": "synthetic low-rank vectors"}))


def ft_record(ndocs):
    """BASELINE configs[3] at `ndocs` documents: returns the sub-record bench.py embeds."""
    import reindexer_b200 as rx
    from ft_helpers import assert_same_merge

    from oracle import ft_oracle as F

    class A:
        pass

    args = A()
    args.docs = ndocs
    total = args.docs + 1
    rng = np.random.default_rng(7)
    words = (rng.poisson(100, size=total).astype(np.uint32) + 1).reshape(-1, 1)
    words[0] = 0
    p = F.FtProblem(total, words)
    npost = 0
    for df in (0.10, 0.01, 0.001):
        nd = int(df * args.docs)
        docs = np.unique(rng.integers(1, total, size=int(nd * 1.06), dtype=np.int64))[:nd].astype(np.uint32)
        npos = rng.integers(1, 4, size=len(docs)).astype(np.uint32)
        begin = np.concatenate([[0], np.cumsum(npos, dtype=np.int64)]).astype(np.uint32)
        first = (rng.random(len(docs)) * np.minimum(words[docs, 0], 60)).astype(np.uint32)
        pos = np.repeat(first, npos) + (np.arange(begin[-1], dtype=np.uint32) - np.repeat(begin[:-1], npos)) * 2  # ascending inside a doc
        p.add_term([(p.add_list_arrays(docs, begin, pos), 100.0)], op=F.OP_OR)
        npost += len(docs)
    ft = rx.GpuFtIndex(total, p.words, p.avg)
    ids = [ft.add_postings(d, b, q) for d, b, q in p.lists]
    terms = [dict(t, postings=[ids[int(x)] for x in t["postings"]]) for t in p.terms]
    res = ft.merge(p.cfg, p.field_cfg, terms)  # warm-up (allocates scratch)
    reps = 5
    t0 = time.perf_counter()
    dev_ms = 0.0
    for _ in range(reps):
        res = ft.merge(p.cfg, p.field_cfg, terms)
        dev_ms += ft.last_stats()["device_ms"]
    gpu_s = (time.perf_counter() - t0) / reps
    st = ft.last_stats()
    ref_res, ref_ns = F.ref_merge(p) if F.ref_available() else F.port_merge(p)
    assert_same_merge(ref_res, res, F.RANK_AND_ID)
    top_gpu, top_ref = F.after_select_order(res)[:100], F.after_select_order(ref_res)[:100]
    assert (top_gpu == top_ref).all()
    # the whole config-3 step in one call: merge + post-processing + (rank desc, id asc) top-100 on the device, 100 rows back
    sel_ids, sel_ranks, sel_n = ft.select(p.cfg, p.field_cfg, terms, 100)
    assert (sel_ids == top_ref["id"]).all() and (sel_ranks == top_ref["normalized_proc"].astype(np.float32)).all() and sel_n == len(ref_res)
    t0 = time.perf_counter()
    sel_dev_ms = 0.0
    for _ in range(reps):
        ft.select(p.cfg, p.field_cfg, terms, 100)
        sel_dev_ms += ft.last_stats()["device_ms"]
    sel_s = (time.perf_counter() - t0) / reps
    sel_st = ft.last_stats()
    peak, src = peak_hbm()
    return {
        "workload": f"ft_fast BM25 merge, {args.docs} docs, 3-term OR (df 10% / 1% / 0.1% = {npost} postings), merge_limit 20000, top-100 "
                    f"(BASELINE configs[3])",
        "metric": "ft_fast top-100 queries/s", "value": 1e3 / (sel_dev_ms / reps), "unit": "queries/s", "higher_is_better": True,
        "e2e": {"value": 1.0 / sel_s, "unit": "queries/s", "h2d_bytes_per_step": 512, "d2h_bytes_per_step": 100 * 8 + 16,
                "ms_per_query": sel_s * 1e3, "call": "rxgpu_ft_select (merge + postProcessResults + afterSelect + sortAfterSelect, top-100)"},
        "ms_per_query_gpu_e2e": sel_s * 1e3, "ms_per_query_gpu_device": sel_dev_ms / reps, "launches_select": sel_st["launches"],
        "merge_only": {"ms_per_query_gpu_e2e": gpu_s * 1e3, "ms_per_query_gpu_device": dev_ms / reps,
                       "call": "rxgpu_ft_merge (all merged documents back to the host)"},
        "cpu_baseline": {"value": 1e9 / ref_ns, "unit": "queries/s", "cores": 1, "kind": "reference" if F.ref_available() else "port",
                         "sample": "1 query, ft::Merger::Merge on one thread (the reference merges a query on one thread)",
                         "ms_per_query": ref_ns / 1e6},
        "speedup_vs_reference_merge": ref_ns / 1e9 / sel_s,
        "merged_docs": int(len(res)), "preselected": st["preselected"], "launches": st["launches"], "postings_scanned": st["postings_scanned"],
        "identical_to_reference": True,
        "roofline": {"bound": "hbm (posting streams + per-document gathers)", "kernel": "ft_rank_pass / ft_hist / ft_score_pass",
                     "achieved": st["algorithmic_bytes"] / (sel_dev_ms / reps * 1e-3) / 1e9,
                     "peak": peak, "unit": "GB/s", "frac": st["algorithmic_bytes"] / (sel_dev_ms / reps * 1e-3) / 1e9 / peak, "peak_source": src,
                     "algorithmic_bytes": st["algorithmic_bytes"], "traffic": None},
        "data": "synthetic postings, Poisson(100) document lengths"}


def ft_sharded_record(comm, rank, world, docs_per_rank, device):
    """BASELINE configs[3] sharded by docid range: every rank holds `docs_per_rank` documents of a world x docs_per_rank
    namespace; one collective rxgpu_sharded_ft_select per query (five small exchanges + the top-100 gather over NCCL).  Collective: every
    rank calls it; the record is meaningful on rank 0.  The answer is checked against the unsharded call in tests/ (bit-equal rows)."""
    import torch
    import torch.distributed as dist

    import reindexer_b200 as rx
    from oracle import ft_oracle as F

    rng = np.random.default_rng(700 + rank)
    words = (rng.poisson(100, size=docs_per_rank).astype(np.uint32) + 1).reshape(-1, 1)
    avg = np.asarray([101.0], np.float32)  # the namespace-wide average field length: the same on every shard
    ft = rx.GpuFtIndex(docs_per_rank, words, avg, device=device)
    p = F.FtProblem(2, np.zeros((2, 1), np.uint32))  # only its config dictionaries are used
    ids, npost = [], 0
    for df in (0.10, 0.01, 0.001):
        nd = int(df * docs_per_rank)
        docs = np.unique(rng.integers(0, docs_per_rank, size=int(nd * 1.06), dtype=np.int64))[:nd].astype(np.uint32)
        npos = rng.integers(1, 4, size=len(docs)).astype(np.uint32)
        begin = np.concatenate([[0], np.cumsum(npos, dtype=np.int64)]).astype(np.uint32)
        first = (rng.random(len(docs)) * np.minimum(words[docs, 0], 60)).astype(np.uint32)
        pos = np.repeat(first, npos) + (np.arange(begin[-1], dtype=np.uint32) - np.repeat(begin[:-1], npos)) * 2
        ids.append(ft.add_postings(docs, begin, pos))
        npost += len(docs)
    terms = [dict(op=F.OP_OR, boost=1.0, term_len_boost=1.0, field_boosts=np.ones(1, np.float32), postings=[i], procs=[100.0]) for i in ids]
    base = rank * docs_per_rank
    out = ft.sharded_select(comm, base, p.cfg, p.field_cfg, terms, 100)  # warm-up (allocates scratch)
    reps = 10
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    dev_ms = 0.0
    for _ in range(reps):
        out = ft.sharded_select(comm, base, p.cfg, p.field_cfg, terms, 100)
        dev_ms += ft.last_stats()["device_ms"]
    dist.barrier()
    sec = (time.perf_counter() - t0) / reps
    st = ft.last_stats()
    t = torch.tensor([sec, dev_ms / reps], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ft.close()
    sec, dev = float(t[0]), float(t[1])
    return {
        "workload": f"ft_fast BM25 merge over {world} docid-range shards, {docs_per_rank} docs and {npost} postings per GPU, 3-term OR "
                    f"(df 10% / 1% / 0.1%), merge_limit 20000, top-100 (BASELINE configs[3], one shard per GPU)",
        "metric": "ft_fast top-100 queries/s over all shards", "value": 1.0 / sec, "unit": "queries/s", "higher_is_better": True,
        "e2e": {"value": 1.0 / sec, "unit": "queries/s", "h2d_bytes_per_step": 512, "d2h_bytes_per_step": 100 * 8 * world + 16,
                "ms_per_query": sec * 1e3, "call": "rxgpu_sharded_ft_select (collective; max over ranks, host clock around the call)"},
        "ms_per_query_device_max_over_ranks": dev, "launches": st["launches"], "preselected": st["preselected"], "rows_total": int(out[2]),
        "top_rank": float(out[1][0]) if len(out[1]) else None, "total_docs": docs_per_rank * world, "scaling": "weak",
        "exchanges": "NCCL: all-reduce of posting counts, mask popcount, 65536-bin histogram (512 KB), max score, max rank; all-gather of the "
                     "threshold-document counts and of every shard's top-100 keys",
        "data": "synthetic postings, Poisson(100) document lengths"}


def run_ft(args):
    print(json.dumps(ft_record(args.docs)))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["hnsw", "ft", "ivf"])
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--nlist", type=int, default=256)
    ap.add_argument("--nprobe", type=int, default=16)
    ap.add_argument("--rows", type=int, default=500000)
    ap.add_argument("--queries", type=int, default=4096)
    ap.add_argument("--docs", type=int, default=50_000_000)
    ap.add_argument("--sweep", action="store_true", help="hnsw: also time several resident-CTA settings")
    a = ap.parse_args()
    {"hnsw": run_hnsw, "ft": run_ft, "ivf": run_ivf}[a.what](a)
