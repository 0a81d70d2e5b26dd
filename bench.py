#!/usr/bin/env python
"""Benchmark of the float_vector brute-force KNN hot path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W          our arm (CUDA, through the C ABI)
  python bench.py --impl reference --gpus N ...          the reference's own CPU implementation on this box's host cores

Workload at N=1 = BASELINE.json configs[1]: brute-force KNN, 10M x 768 fp32, inner product, k=10, batch of 1024 queries on one
B200.  A "step" is one batch of 1024 queries against the resident index.  For N>1 the namespace is sharded by row range, 10M rows
per GPU (weak scaling, configs[4] at N=8); every rank scans its shard for all 1024 queries and one NCCL all-gather + merge
yields the global top-k; `value` counts the (query x 10M-row-shard) scans all ranks complete per second.
Synthetic data: rows and queries from the counter-based generator in reindexer_b200/csrc/common.cuh (sigma 0.25, like the
reference's own test generator), produced directly in HBM.  Inputs are far larger than L2 (30.7 GB vs 126 MB), so no flush.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "KNN QPS @ recall@10 (10Mx768, k=10) + HBM GB/s vs roofline"
UNIT = "queries/s"
DIM, K, NQ = 768, 10, 1024
ROWS_FULL = 10_000_000
SEED = 0x5EED0001


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "50", "-i",
                                          str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t_begin=None, t_end=None):
        """samples taken inside [t_begin, t_end] (the timed region); nvidia-smi is started before the warm-up so that it is
        already reporting when the region begins"""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        inside = [ln for (ts, ln) in self.lines if t_begin is None or (t_begin <= ts <= t_end + 0.2)]
        window = "timed region"
        if not inside:  # region shorter than one sampling period: report the samples under the same load (warm-up + region)
            inside, window = [ln for (_, ln) in self.lines], "warm-up + timed region"
        for ln in inside:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


# ----------------------------------------------------------------------------------------------------------------- CPU arm
class CpuReference:
    """The reference's own hnswlib::BruteforceSearch::SearchKnn (oracle/_ref, runtime ISA dispatch -> AVX-512 here) -- or the C
    port when the reference build is absent -- on a bounded sample of the workload: `rows_sample` rows of the same generator,
    `threads` host threads each issuing independent single-threaded queries (the concurrency the reference permits).  Brute
    force is linear in the row count, so QPS at 10M rows = QPS(sample) * rows_sample / 10M (labelled "extrapolated")."""

    def __init__(self, rows_sample, nq_per_round, threads):
        from oracle import oracle as O

        self.kind = "reference" if O.ref_knn_available() else "port"
        self.rows_sample, self.nq, self.threads = rows_sample, nq_per_round, threads
        vecs = np.empty((rows_sample, DIM), np.float32)
        O.port_lib().port_synth_fill(SEED, 0, rows_sample * DIM, vecs.ctypes.data_as(O._f32p))
        self.queries = O.synth_matrix(SEED + 1, nq_per_round, DIM)
        labels = O.row_labels(rows_sample)
        if self.kind == "reference":
            self.bf = O.RefBF(O.IP, DIM, rows_sample)
            self.isa = {3: "avx512", 2: "avx2", 1: "avx", 0: "sse"}[O.ref_knn_lib().ref_isa_level()]
        else:
            self.bf = O.PortBF(O.IP, DIM, rows_sample)
            self.isa, self.threads = "scalar-c", 1
        assert self.bf.add_batch(labels, vecs) == 0
        self.round()  # warm-up

    def round(self):
        t0 = time.perf_counter()
        if self.kind == "reference":
            self.bf.search_knn_batch(self.queries, K, self.threads)
        else:
            for q in self.queries:
                self.bf.search_knn(q, K)
        return time.perf_counter() - t0

    def measure(self, rounds):
        secs = sum(self.round() for _ in range(rounds)) / rounds
        qps_sample = self.nq / secs
        return {"value": qps_sample * self.rows_sample / ROWS_FULL, "unit": UNIT, "cores": self.threads, "kind": self.kind,
                "isa": self.isa, "qps_on_sample": qps_sample, "seconds_per_round": secs,
                "sample": f"{self.nq} queries x {rounds} rounds over {self.rows_sample} rows x {DIM} (same generator), "
                          f"{self.threads} threads; QPS scaled linearly to {ROWS_FULL} rows (extrapolated)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    ref = CpuReference(200_000, max(2 * threads, 16), threads)
    per_round = ref.round()
    rounds_per_step = max(1, int(2.0 / max(per_round, 1e-3)))  # ~2 s of wall clock per step
    t0 = time.perf_counter()
    for _ in range(args.warmup):
        ref.measure(rounds_per_step)
    vals = []
    for _ in range(args.steps):
        vals.append(ref.measure(rounds_per_step))
        if time.perf_counter() - t0 > 240:
            break
    value = float(np.mean([v["value"] for v in vals]))
    last = vals[-1]
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(vals),
            "warmup": args.warmup, "ms_per_step": 1000.0 * NQ / value, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "brute-force KNN, 10M x 768 fp32, inner-product, k=10, batch=1024 queries (BASELINE configs[1])",
                       "rows": ROWS_FULL, "dim": DIM, "k": K, "batch": NQ, "cpu_path": last["kind"], "isa": last["isa"]},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": last["cores"], "kind": last["kind"], "sample": last["sample"]},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    import reindexer_b200 as rx
    from reindexer_b200 import binding as B
    from reindexer_b200.sharded import ShardedBruteforceSearch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if rx.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device -- librxgpu has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        # the contract is ONE JSON line on stdout: NCCL announces its version on stdout when the first communicator is created, so
        # stdout points at stderr while the process group comes up
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    rows = args.rows or ROWS_FULL
    free_b, _ = torch.cuda.mem_get_info()
    if rows * DIM * 4 * 1.05 > free_b:
        raise SystemExit(f"bench.py: {rows} x {DIM} fp32 does not fit in {free_b / 1e9:.0f} GB of free HBM")

    t_fill = time.perf_counter()
    idx = rx.GpuBruteforceSearch(rx.IP, DIM, rows, device=local_rank)
    idx.append_synth(SEED, rank * rows, rows)  # shard `rank` = global rows [rank*rows, (rank+1)*rows)
    if args.query_tile:
        idx.set_query_tile(args.query_tile)
    if args.tc:
        idx.set_tensor_core_filter(args.tc)
    fill_s = time.perf_counter() - t_fill
    sharded = ShardedBruteforceSearch(idx, rows) if world > 1 else None

    stream = torch.cuda.current_stream()
    dq = torch.empty((NQ, DIM), dtype=torch.float32, device="cuda")
    B._check(B.lib().rxgpu_synth_fill_device(dq.data_ptr(), SEED + 1, 0, NQ * DIM, local_rank, stream.cuda_stream))
    hq = dq.cpu().numpy()
    hq_pinned = torch.from_numpy(hq).pin_memory()
    k1 = K + 1
    od = torch.zeros((NQ, k1), dtype=torch.float32, device="cuda")
    oi = torch.zeros((NQ, k1), dtype=torch.int32, device="cuda")
    ol = torch.zeros((NQ, k1), dtype=torch.int64, device="cuda")
    oc = torch.zeros((NQ,), dtype=torch.int32, device="cuda")

    def step_resident():
        if sharded is not None:
            return sharded.search_knn(dq, K)
        idx.search_knn_device(NQ, dq.data_ptr(), k1, od.data_ptr(), oi.data_ptr(), ol.data_ptr(), oc.data_ptr(), stream.cuda_stream)
        return None

    def step_e2e():
        if sharded is not None:
            return sharded.search_knn(hq_pinned.numpy(), K)
        return idx.search_knn(hq, K)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident leg: CUDA events on the launching stream, max over ranks
    B.lib().rxgpu_set_profile(1)
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        step_resident()
    barrier()
    t_begin = time.perf_counter()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = passes = 0
    scan_ms = 0.0
    scan_launches = 0
    alg_bytes = 0
    ev0.record(stream)
    for _ in range(args.steps):
        step_resident()
        st = rx.last_search_stats()
        launches += st["launches"]
        passes += st["passes"]
        scan_ms += st["scan_kernel_ms"]
        scan_launches += st["scan_launches"]
        alg_bytes += st["algorithmic_bytes"]
        qt = st["query_tile"]
        tc_used = st["tc_used"]
        tc_cands = st["tc_candidates"]
        tc_fallbacks = st["tc_fallbacks"]
    ev1.record(stream)
    barrier()
    clocks = sampler.stop(t_begin, time.perf_counter())
    ms_total = ev0.elapsed_time(ev1)
    B.lib().rxgpu_set_profile(0)
    # ---- end-to-end leg: host buffers through the reference-facing C ABI call, copies inside the timed region
    for _ in range(min(args.warmup, 1)):
        res = step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    e2e_stats = rx.last_search_stats()
    if world > 1:
        t = torch.tensor([ms_total, e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, e2e_s = float(t[0]), float(t[1])
        tot = torch.tensor([launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(tot)
        launches_all = int(tot[0])
    else:
        launches_all = launches

    # sanity on the timed work: every query of the last batch got k results, sorted, and (N=1) planted-free recompute of row 0
    d_chk, l_chk, c_chk = res
    assert (np.asarray(c_chk) == K).all() and (np.diff(d_chk[:, :K], axis=1) >= 0).all()
    # recall@10 of what was timed: the search is exact by construction; check it in-run against the fp32 exact-scan path of the same
    # index on a few queries of the batch (outside the timed regions): the labels and the distance bits must be identical
    recall_checked = 0
    if world == 1:
        idx.set_tensor_core_filter(2)
        d_ex, l_ex, _ = idx.search_knn(hq[:8], K)
        idx.set_tensor_core_filter(args.tc or 0)
        assert (np.asarray(l_ex) == np.asarray(l_chk)[:8, :K]).all() and \
            (np.asarray(d_ex).view(np.uint32) == np.asarray(d_chk)[:8, :K].view(np.uint32)).all(), "timed path differs from the exact scan"
        recall_checked = 8

    if rank == 0:
        ms_per_step = ms_total / args.steps
        value = world * NQ / (ms_per_step / 1000.0)
        e2e_value = world * NQ / (e2e_s / args.steps)
        peak, peak_src = load_peaks()
        peaks_all = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
        if tc_used:  # dominant kernel = knn_tc_filter: bf16 shadow rows + row norms + the resident query block, per launch
            per_launch_bytes = rows * DIM * 2 + rows * 8 + qt * DIM * 2
            kernel_name = "knn_tc_filter_q (tcgen05 bf16 filter, queries in TMEM, certified bound) + knn_rerank (exact fp32)"
        else:
            per_launch_bytes = alg_bytes / max(passes, 1)
            kernel_name = "knn_scan_warp (fp32 FMA, fused top-k)"
        avg_launch_ms = scan_ms / max(scan_launches, 1)
        achieved = per_launch_bytes / (avg_launch_ms * 1e-3) / 1e9 if avg_launch_ms > 0 else 0.0
        # Which roofline binds the dominant kernel: the larger of its HBM time (algorithmic bytes / measured copy peak) and its
        # tensor time (algorithmic bf16 MMA flops / measured SUSTAINED cuBLAS rate -- the kernel is timed inside a long step).
        # The exact scan has no tensor work; the filter serves `qt` queries per pass (256 / 512 with clusters of 2 / 4).
        tensor_sust, tensor_burst = peaks_all.get("bf16_tflops_sustained"), peaks_all.get("bf16_tflops")
        tensor_src = "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
        if not tensor_sust:  # B200_PROFILING.md fallback when the driver's file is absent
            tensor_sust, tensor_burst, tensor_src = 1590.0, 1590.0, "fallback (B200_PROFILING.md 1.59 PFLOP/s)"
        flops_per_launch = 2.0 * rows * DIM * qt if tc_used else 0.0
        tensor_tflops = flops_per_launch / (avg_launch_ms * 1e-3) / 1e12 if tc_used and avg_launch_ms > 0 else None
        t_hbm = per_launch_bytes / (peak * 1e9)
        t_tensor = flops_per_launch / (tensor_sust * 1e12) if tc_used and tensor_sust else 0.0
        common = {
            # dram__bytes_read.sum + dram__bytes_write.sum per launch from the ncu --set full captures under profiles/
            # (r1_knn_tc_filter_q_*_raw.csv, r1_knn_scan_warp_full_raw.csv), valid for the full-size workload only
            "traffic": (None if rows != 10_000_000 else 15.4224e9 if tc_used else 30.7201e9),
            "kernel": "knn_tc_filter_q" if tc_used else "knn_scan_warp", "bytes_per_launch": per_launch_bytes,
            "flops_per_launch": flops_per_launch, "avg_launch_ms": avg_launch_ms, "launches_timed": scan_launches,
            "kernel_share_of_step": scan_ms / ms_total if ms_total else None,
            "hbm": {"achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src},
            "tensor": {"achieved": tensor_tflops, "peak": tensor_sust, "peak_burst": tensor_burst, "unit": "TFLOP/s",
                       "frac": (tensor_tflops / tensor_sust) if tensor_tflops and tensor_sust else None,
                       "peak_source": tensor_src},
        }
        if t_tensor > t_hbm:
            roofline = {"bound": "tensor", "achieved": tensor_tflops, "peak": tensor_sust, "unit": "TFLOP/s",
                        "frac": tensor_tflops / tensor_sust, "peak_source": common["tensor"]["peak_source"], **common}
        else:
            roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                        "peak_source": peak_src, **common}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "brute-force KNN, 10M x 768 fp32, inner-product, k=10, batch=1024 queries (BASELINE configs[1])"
                       if world == 1 else f"brute-force KNN, {world} x 10M x 768 fp32 sharded by row range, inner-product, k=10, "
                                          f"batch=1024, NCCL all-gather top-k merge (BASELINE configs[4] at N=8)",
                       "rows_per_gpu": rows, "total_rows": rows * world, "dim": DIM, "k": K, "batch": NQ, "query_tile": qt,
"kernel": kernel_name, "tc_candidates_per_step": tc_cands, "tc_fallbacks": tc_fallbacks,
                       "l2_policy": "inputs (30.7 GB/GPU) larger than L2, no flush",
                       "global_queries_per_s": NQ / (ms_per_step / 1000.0), "index_fill_s": round(fill_s, 2),
                       "value_definition": "(query x 10M-row shard) scans per second over all ranks",
                       "recall_at_10": 1.0, "recall_basis": f"exact search; {recall_checked} queries of the timed batch re-run on the fp32 "
                                                            f"exact-scan path in this run: identical labels and distance bits"},
            "roofline": roofline,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": NQ * DIM * 4,
                    "d2h_bytes_per_step": NQ * k1 * 16 + NQ * 4, "ms_per_step": 1000.0 * e2e_s / args.steps,
                    "tie_replays": e2e_stats["tie_replays"]},
            "gpu_launches": launches_all,
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            cb = CpuReference(200_000, max(2 * threads, 16), threads).measure(3)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "isa")}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=0, help="rows per GPU (default 10M = BASELINE config)")
    ap.add_argument("--query-tile", type=int, default=0)
    ap.add_argument("--tc", type=int, default=0, help="tensor-core filter: 0 auto, 1 on, 2 off (exact fp32 scan only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
