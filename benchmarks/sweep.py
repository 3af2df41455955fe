#!/usr/bin/env python
"""Batch-size sweep (BASELINE.json configs[4]): mixed insert/delete streaming batches of 1 K .. 100 M edges applied to
ONE evolving graph, a PageRank push step after every batch -- the GPU counterpart of the reference's benchmark
scripts (src/benchmarking/benchmark-strong-scaling.sh:83-156 vary the thread count; here the batch size varies).

  python benchmarks/sweep.py [--scale 20] [--max-batch 100000000] [--reps 5] [--workload mixed|insert]

Unlike bench.py the graph is NOT restored between batches: every batch meets the state the previous ones left
(steady state of a streaming system).  Prints one JSON line per batch size:
  {"batch": B, "updates_per_sec": ..., "ms_per_batch": ..., "ms_pagerank": ..., "stages_ms": {...}, "windows": ...}
Runs on one GPU, or under torchrun on several (vertex-range shards, the batch is split over the ranks).
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=int, default=20)
    ap.add_argument("--max-batch", type=int, default=100_000_000)
    ap.add_argument("--min-batch", type=int, default=1000)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--workload", default="mixed", choices=["mixed", "insert"])
    ap.add_argument("--no-pagerank", action="store_true")
    args = ap.parse_args()
    import torch

    pp = importlib.import_module("parallel-packed-csr_b200")
    synth = importlib.import_module("parallel-packed-csr_b200.synth")
    router = importlib.import_module("parallel-packed-csr_b200.router")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        dist.init_process_group("nccl", device_id=dev)
    scale, n = args.scale, 1 << args.scale
    total = 16 << scale
    lo, hi = rank * total // world, (rank + 1) * total // world
    cs, cd = synth.rmat(scale, lo, hi, 42, device=dev)
    cs, cd = cs.to(torch.int32), cd.to(torch.int32)
    starts = router.edge_balanced_starts(cs, n, world, dist, vertex_weight=None) if world > 1 else np.array([0, n], dtype=np.uint64)
    graph = router.ShardedGraph(n, starts, rank, world, local_rank, dist=dist,
                                peer_cap=max(hi - lo, args.max_batch // world), peer_values=True)
    graph.shard.bind_torch_stream()
    graph.apply(cs, cd)
    del cs, cd
    # a streaming system reaches its working size once: allocate it up front so that no batch pays cudaMalloc
    geo = graph.shard.geometry
    expect = sum((1 + (args.reps if B <= 10_000_000 else max(1, args.reps // 2))) * B
                 for B in (args.min_batch * 10 ** k for k in range(12)) if B <= args.max_batch)
    max_slots = int(geo.N)
    while max_slots < min(4 * (geo.N // 2 + expect // world), 1 << 31):
        max_slots *= 2
    graph.shard.reserve(max_slots=min(max_slots, 1 << 31), max_batch=int(args.max_batch // world * 1.3) + 1024)
    stream = torch.cuda.current_stream()
    vals = 1.0 + (torch.arange(n, device=dev, dtype=torch.float64) % 7)
    offset = 0  # position in the global update stream: every batch consumes fresh updates
    B = args.min_batch
    while B <= args.max_batch:
        b = B // world
        reps = args.reps if B <= 10_000_000 else max(1, args.reps // 2)
        t_upd, t_pr, stats = 0.0, 0.0, None
        for rep in range(reps + 1):  # the first batch of every size is a warm-up (allocations for the new size)
            base = offset + rank * b
            us, ud = synth.uniform(scale, base, base + b, 7, device=dev)
            uv = None
            if args.workload == "mixed":
                uv = synth.mixed_ops(base, base + b, 11, device=dev)
                idx = (synth._base_hash(torch.arange(base, base + b, device=dev), 5) * 2654435761) % total
                ds, dd = synth.rmat_at(scale, idx, 42)  # deletes aim at (possibly already deleted) core edges
                us, ud = torch.where(uv != 0, us, ds), torch.where(uv != 0, ud, dd)
                uv = uv.to(torch.int32).contiguous()
            us, ud = us.to(torch.int32).contiguous(), ud.to(torch.int32).contiguous()
            offset += b * world
            if dist is not None:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record(stream)
            stats = graph.apply(us, ud, uv)
            e1.record(stream)
            if not args.no_pagerank:
                graph.pagerank_step(vals)
            e2.record(stream)
            torch.cuda.synchronize()
            if rep > 0:
                t_upd += e0.elapsed_time(e1)
                t_pr += e1.elapsed_time(e2)
        t = torch.tensor([t_upd / reps, t_pr / reps], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        rep = graph.shard.check(check_lower=False)
        if rep.violations(False):
            raise SystemExit(f"PMA invariants violated at batch {B}: {rep.as_dict()}")
        if rank == 0:
            print(json.dumps({
                "batch": b * world, "n_gpus": world, "workload": args.workload, "scale": scale,
                "updates_per_sec": b * world / (float(t[0]) / 1e3), "ms_per_batch": float(t[0]),
                "ms_pagerank": float(t[1]),
                "stages_ms_rank0": {k: round(stats[k], 4) for k in ("ms_sort", "ms_locate", "ms_select", "ms_rebalance")},
                "windows_rank0": int(stats["n_windows"]), "whole_array_rank0": int(stats["whole_array"]),
                "slots_rank0": int(stats["slots_after"]), "kernel_launches_rank0": int(stats["kernel_launches"]),
                "sparse_path_rank0": int(stats.get("sparse_path", 0))}), flush=True)
        B *= 10
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
