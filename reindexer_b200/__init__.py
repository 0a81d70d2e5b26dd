"""reindexer_b200 -- B200-native (sm_100a) replacement for Reindexer's float_vector KNN / ft_fast BM25 hot path.

The product is the C-ABI library ``librxgpu.so`` (include/rxgpu.h) and the C++ adapter under ``host/``; this package is
the thin Python driver used by tests, the benchmark and the multi-GPU (one process per GPU, torch.distributed) plumbing.
"""
from .binding import (COS, FLAG_HOST_MIRROR, IP, L2, GpuBruteforceSearch, GpuFtIndex, RxGpuError, ShardComm, device_count, last_search_stats,
                      lib, merge_shards, tie_replay)

__all__ = ["L2", "IP", "COS", "FLAG_HOST_MIRROR", "GpuBruteforceSearch", "GpuFtIndex", "RxGpuError", "ShardComm", "device_count", "last_search_stats", "lib",
           "merge_shards", "tie_replay"]
