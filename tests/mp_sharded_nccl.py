"""torchrun script (world >= 2, one GPU per rank): the C-ABI sharded search (rxgpu_sharded_search_knn over NCCL) must return, on every
rank, exactly what ONE index holding all rows returns through rxgpu_search_knn -- same labels, same order, same distance bits, the
reference's tie rule included -- for the exact-scan path (few queries) and the tensor-core filter path (a batch).  Then the ft_fast merge
over docid-range shards (rxgpu_sharded_ft_select, the same communicator) against rxgpu_ft_select over all documents."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import reindexer_b200 as rx  # noqa: E402
from reindexer_b200.sharded import ShardedBruteforceSearch  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    k = 10
    for case, (metric, dim, rows, nq, tc, ties) in enumerate([(rx.L2, 32, 15000, 5, 0, True), (rx.IP, 32, 20000, 128, 1, True),
                                                              (rx.IP, 96, 40000, 300, 1, False), (rx.COS, 64, 9000, 4, 0, False)]):
        rng = np.random.default_rng(100 + case)  # the same stream on every rank
        total = rows * world
        if ties:
            allv = rng.integers(-2, 3, size=(total, dim)).astype(np.float32)
            queries = rng.integers(-2, 3, size=(nq, dim)).astype(np.float32)
        else:
            allv = rng.normal(0, 0.25, size=(total, dim)).astype(np.float32)
            queries = rng.normal(0, 0.25, size=(nq, dim)).astype(np.float32)
        if metric == rx.COS:
            queries /= np.linalg.norm(queries, axis=1, keepdims=True)
        labels = (np.arange(total, dtype=np.uint64) << np.uint64(32))
        shard = rx.GpuBruteforceSearch(metric, dim, rows, device=local)
        shard.add_points(labels[rank * rows:(rank + 1) * rows], allv[rank * rows:(rank + 1) * rows])
        shard.set_tensor_core_filter(tc)
        s = ShardedBruteforceSearch(shard, rows)
        assert s.comm is not None
        d1, l1, c1 = s.search_knn(queries, k)
        st = rx.last_search_stats()
        dq = torch.from_numpy(queries).cuda()
        d2, l2, c2 = s.search_knn(dq, k)
        assert (l1 == l2).all() and (d1.view(np.uint32) == d2.view(np.uint32)).all()
        full = rx.GpuBruteforceSearch(metric, dim, total, device=local)
        full.add_points(labels, allv)
        full.set_tensor_core_filter(2)
        d0, l0, c0 = full.search_knn(queries, k)
        assert (c0 == c1).all(), (case, rank)
        assert (l0 == l1).all(), (case, rank, np.argwhere(l0 != l1)[:4])
        assert (d0.view(np.uint32) == d1.view(np.uint32)).all(), (case, rank)
        if ties:
            assert st["tie_replays"] > 0, st
            if tc:
                assert st["tie_from_lists"] == st["tie_replays"], st
        full.close()
        shard.close()
        dist.barrier()
    # ---- ft_fast merge over docid-range shards: NCCL all-reduces / all-gathers between the ranks (tests/test_ft_sharded_gpu.py has the
    # same checks with the ranks as threads of one process)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from ft_helpers import random_problem  # noqa: E402
    from test_ft_sharded_gpu import shard_of, upload  # noqa: E402
    from oracle import ft_oracle as F  # noqa: E402

    anchor = rx.GpuBruteforceSearch(rx.L2, 4, 8, device=local)  # the communicator of a sharded index on this device
    comm = ShardedBruteforceSearch(anchor, 8).comm
    for seed, merge_limit in ((1, 60), (2, 20000), (3, 150), (4, 20000)):
        totald = 2400 + 100 * seed
        prob = random_problem(40 + seed, total_docs=totald, nfields=1 + seed % 2, nterms=3, density=0.3, merge_limit=merge_limit,
                              removed_frac=0.04 if seed == 3 else 0.0)
        cuts = [0] + [totald * (r + 1) // world + (7 * r) % 13 for r in range(world - 1)] + [totald]
        for sort_type in (F.RANK_AND_ID, F.ID_ONLY):
            ft, ids = shard_of(prob, cuts[rank], cuts[rank + 1], device=local)
            terms = [dict(t, postings=[ids[int(x)] for x in t["postings"]]) for t in prob.terms]
            got = ft.sharded_select(comm, cuts[rank], prob.cfg, prob.field_cfg, terms, 300, rank_sort_type=sort_type)
            ft.close()
            whole, wids = upload(prob, device=local)
            wterms = [dict(t, postings=[wids[int(x)] for x in t["postings"]]) for t in prob.terms]
            want = whole.select(prob.cfg, prob.field_cfg, wterms, 300, rank_sort_type=sort_type)
            whole.close()
            assert got[2] == want[2] and (got[0] == want[0]).all() and (got[1] == want[1]).all(), ("ft", seed, sort_type, rank)
        dist.barrier()
    anchor.close()
    if rank == 0:
        print("mp_sharded_nccl ok", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
