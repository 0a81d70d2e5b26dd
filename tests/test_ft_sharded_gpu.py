"""GPU tests of the ft_fast merge over docid-range shards (rxgpu_sharded_ft_select): the namespace is cut into contiguous
docid ranges, every shard holds its slice under local ids, and the collective call must return exactly what rxgpu_ft_select returns
for one index over all documents -- which tests/test_ft_gpu.py pins to the reference's Merger::Merge + afterSelect order.  The shards
here are the ranks of ONE process (rxgpu_comm_create_local: one thread per rank, all on cuda:0), so a one-GPU box runs every
cross-shard exchange: BM25's namespace-wide counts, the mask popcount, the preselect histogram, the ordered cut at the threshold score,
the uint8 normalisation maximum, the final gather.  tests/mp_sharded_nccl.py runs the same call over NCCL on two GPUs."""
import threading

import numpy as np
import pytest
from ft_helpers import random_problem

import reindexer_b200 as rx
from oracle import ft_oracle as F

pytestmark = pytest.mark.gpu


def upload(prob, device=0):
    ft = rx.GpuFtIndex(prob.total_docs, prob.words, prob.avg, prob.removed, device=device)
    ids = [ft.add_postings(d, b, p) for d, b, p in prob.lists]
    return ft, ids


def shard_of(prob, lo, hi, device=0):
    """documents [lo, hi) under local ids: word counts, removed flags and every posting list cut to the range"""
    ft = rx.GpuFtIndex(hi - lo, prob.words[lo:hi], prob.avg, None if prob.removed is None else prob.removed[lo:hi], device=device)
    ids = []
    for d, b, p in prob.lists:
        d, b, p = np.asarray(d, np.uint32), np.asarray(b, np.uint32), np.asarray(p, np.uint32)
        keep = np.nonzero((d >= lo) & (d < hi))[0]
        cnt = (b[keep + 1] - b[keep]).astype(np.uint32)
        nb = np.zeros(len(keep) + 1, np.uint32)
        nb[1:] = np.cumsum(cnt)
        pos = np.concatenate([p[b[i]:b[i + 1]] for i in keep]) if len(keep) else np.zeros(0, np.uint32)
        ids.append(ft.add_postings((d[keep] - lo).astype(np.uint32), nb, pos.astype(np.uint32)))
    return ft, ids


def run_sharded(prob, bounds, limit, rank_sort_type, row_tables=None, row_status=None):
    R = len(bounds) - 1
    comms = rx.ShardComm.local_group(R)
    shards = [shard_of(prob, bounds[r], bounds[r + 1]) for r in range(R)]
    results, errors = [None] * R, [None] * R

    def work(r):
        try:
            ft, ids = shards[r]
            lo, hi = bounds[r], bounds[r + 1]
            if row_tables is not None:
                ft.set_rows(*row_tables[r])
            terms = [dict(t, postings=[ids[int(x)] for x in t["postings"]]) for t in prob.terms]
            ex = None if prob.excluded is None else prob.excluded[lo:hi]
            results[r] = ft.sharded_select(comms[r], lo, prob.cfg, prob.field_cfg, terms, limit, excluded=ex, row_status=row_status,
                                           rank_sort_type=rank_sort_type)
        except Exception as e:  # noqa: BLE001 - reported by the main thread
            errors[r] = e

    threads = [threading.Thread(target=work, args=(r,)) for r in range(R)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=120)
    assert not any(t.is_alive() for t in threads), "a rank is stuck in a collective"
    for ft, _ in shards:
        ft.close()
    for c in comms:
        c.close()
    assert errors == [None] * R, errors
    return results


def run_whole(prob, limit, rank_sort_type, rows=None, row_status=None):
    ft, ids = upload(prob)
    if rows is not None:
        ft.set_rows(*rows)
    terms = [dict(t, postings=[ids[int(x)] for x in t["postings"]]) for t in prob.terms]
    out = ft.select(prob.cfg, prob.field_cfg, terms, limit, excluded=prob.excluded, row_status=row_status, rank_sort_type=rank_sort_type)
    st = ft.last_stats()
    ft.close()
    return out, st


def bounds_for(total, nshards, rng):
    cuts = np.sort(rng.choice(np.arange(1, total), size=nshards - 1, replace=False)) if nshards > 1 else np.zeros(0, np.int64)
    return [0] + [int(c) for c in cuts] + [total]


@pytest.mark.timeout(600)
@pytest.mark.parametrize("seed", range(12))
def test_sharded_select_equals_unsharded(seed):
    """random OR / AND / NOT problems, 2-5 shards of uneven size, with and without the preselect (merge_limit below / above the matches),
    both orders of sortAfterSelect, removed and excluded documents"""
    rng = np.random.default_rng(1000 + seed)
    total = int(rng.integers(900, 2600))
    small_limit = seed % 2 == 0  # merge_limit far below the matching documents: preselectMostRelevantDocs decides who survives
    prob = random_problem(seed * 7 + 3, total_docs=total, nfields=int(rng.integers(1, 3)), nterms=int(rng.integers(2, 5)), density=0.25,
                          merge_limit=int(rng.integers(40, 120)) if small_limit else 20000, removed_frac=0.05 if seed % 3 == 0 else 0.0,
                          excluded_frac=0.05 if seed % 4 == 1 else 0.0)
    prob.cfg["min_rank"] = int(rng.choice([0, 5, 30]))
    limit = int(rng.choice([10, 100, 5000]))
    for sort_type in (F.RANK_AND_ID, F.ID_ONLY):
        whole, st = run_whole(prob, limit, sort_type)
        if small_limit and all(t["op"] != F.OP_NOT for t in prob.terms[:1]):
            assert st["preselected"] in (0, 1)
        for nshards in (2, int(rng.integers(3, 6))):
            res = run_sharded(prob, bounds_for(total, nshards, rng), limit, sort_type)
            for r, (ids, ranks, n) in enumerate(res):
                assert n == whole[2], (seed, sort_type, nshards, r, n, whole[2])
                assert (ids == whole[0]).all() and (ranks == whole[1]).all(), (seed, sort_type, nshards, r)


@pytest.mark.timeout(300)
def test_sharded_preselect_ties_cut_in_global_id_order():
    """two dense single-list OR terms: only three distinct scores, merge_limit in the middle of a shard -- the documents kept AT the
    threshold score are the lowest GLOBAL ids (mergerimpl.h:448-462), i.e. the budget left for a shard depends on the shards below it"""
    total = 3000
    prob = random_problem(5, total_docs=total, nfields=1, nterms=2, max_sub=1, density=0.9, merge_limit=700, ops=[F.OP_OR, F.OP_OR])
    prob.cfg["min_rank"] = 0
    whole, st = run_whole(prob, 4000, F.ID_ONLY)
    assert st["preselected"] == 1 and whole[2] == 700
    for bounds in ([0, 1000, 2000, total], [0, 150, 300, 450, total], [0, 2900, total]):
        for ids, ranks, n in run_sharded(prob, bounds, 4000, F.ID_ONLY):
            assert n == 700 and (ids == whole[0]).all() and (ranks == whole[1]).all(), bounds


@pytest.mark.timeout(300)
def test_sharded_select_with_row_tables_and_statuses():
    """vdoc -> row ids through rxgpu_ft_set_rows (global row ids, some vdocs with two rows) and external row statuses"""
    rng = np.random.default_rng(77)
    total = 1500
    prob = random_problem(21, total_docs=total, nfields=2, nterms=3, density=0.3, merge_limit=20000)
    nrows_of = rng.integers(1, 3, size=total)
    row_begin = np.zeros(total + 1, np.uint32)
    row_begin[1:] = np.cumsum(nrows_of)
    row_ids = rng.permutation(int(row_begin[-1])).astype(np.int32)
    status = (rng.random(int(row_begin[-1])) < 0.8).astype(np.uint8)
    whole, _ = run_whole(prob, 200, F.RANK_AND_ID, rows=(row_begin, row_ids), row_status=status)
    bounds = [0, 400, 1100, total]
    tables = []
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        rb = (row_begin[lo:hi + 1] - row_begin[lo]).astype(np.uint32)
        tables.append((rb, row_ids[row_begin[lo]:row_begin[hi]]))
    for ids, ranks, n in run_sharded(prob, bounds, 200, F.RANK_AND_ID, row_tables=tables, row_status=status):
        assert n == whole[2] and (ids == whole[0]).all() and (ranks == whole[1]).all()


def test_sharded_select_rejects_what_it_does_not_serve():
    prob = random_problem(3, total_docs=300, nterms=2)
    (comm,) = rx.ShardComm.local_group(1)
    ft, ids = upload(prob)
    terms = [dict(t, postings=[ids[int(x)] for x in t["postings"]]) for t in prob.terms]
    terms[0] = dict(terms[0], phrase_num=1)
    terms[1] = dict(terms[1], phrase_num=1, distance=1)
    with pytest.raises(rx.RxGpuError):
        ft.sharded_select(comm, 0, prob.cfg, prob.field_cfg, terms, 10)
    # a single rank is the plain select
    terms = [dict(t, postings=[ids[int(x)] for x in t["postings"]]) for t in prob.terms]
    a = ft.sharded_select(comm, 0, prob.cfg, prob.field_cfg, terms, 50)
    b = ft.select(prob.cfg, prob.field_cfg, terms, 50)
    assert a[2] == b[2] and (a[0] == b[0]).all() and (a[1] == b[1]).all()
    ft.close()
    comm.close()
