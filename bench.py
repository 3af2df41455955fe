#!/usr/bin/env python
"""bench.py -- batched edge updates/s of the B200 Parallel Packed CSR engine (BASELINE.json metric).

A "step" is ONE pass of the hot path over one batch: the whole update batch is sorted, located, and
merged into the packed edge array (window selection + rebalance, array doubling folded in).

N = 1 (default): BASELINE.json configs[1] -- R-MAT scale-20 core (16.7 M raw edges) + 10 M uniform-random
edge insertions as one batch.  The shard is restored from a device snapshot before every step (untimed),
so every step is exactly that configuration, including the 2^25 -> 2^26 slot doubling.
N > 1 (torchrun): weak scaling -- the global graph has scale 20+log2(N), vertex-range shards (one per GPU,
edge-balanced boundaries), every rank contributes its own slice of 10 M updates per step; updates are
binned by owner on the device, exchanged with ONE NCCL all-to-all and applied by the owning shard.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload insert|delete]

`--impl reference` times the reference's own CPU implementation (oracle/_ref/ref_driver: the unmodified
reference sources driven through ThreadPoolPPPCSR, -pppcsrnuma, all host threads) on a bounded sample of
the same workload.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "edge_updates_per_sec"
UNIT = "updates/s"
SLOT_BYTES = 8  # this build: u32 dest + u32 value per slot (SoA)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="insert", choices=["insert", "delete", "skewed", "mixed"],
                    help="insert: uniform inserts (C2); delete: deletes sampled from the core (C3); skewed: R-MAT "
                         "inserts (C4); mixed: 3/4 uniform inserts + 1/4 deletes of core edges, per-update op (C5)")
    ap.add_argument("--strong", action="store_true",
                    help="strong scaling: --scale and --batch are the GLOBAL graph and batch, split over the ranks")
    ap.add_argument("--pagerank", action="store_true", help="one PageRank push step after every batch (timed apart)")
    ap.add_argument("--scale", type=int, default=20)
    ap.add_argument("--batch", type=int, default=10_000_000)
    ap.add_argument("--cpu-sample", type=int, default=2_000_000, help="updates in the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline
# ------------------------------------------------------------------------------------------------
def host_updates(scale, workload, lo, hi, synth, core=None):
    """Updates [lo, hi) of the workload's global stream on the host: (src, dst, value-or-array)."""
    total = 16 << scale
    if workload == "insert":
        us, ud = synth.uniform(scale, lo, hi, 7)
        return us, ud, 1
    if workload == "skewed":
        us, ud = synth.rmat(scale, lo, hi, 99)
        return us, ud, 1
    cs, cd = core if core is not None else synth.rmat(scale, 0, total, 42)
    idx = synth.sample_without_replacement(total, hi, 7)[lo:hi]
    if workload == "delete":
        return cs[idx], cd[idx], 0
    ops = synth.mixed_ops(lo, hi, 11)
    us, ud = synth.uniform(scale, lo, hi, 7)
    return np.where(ops != 0, us, cs[idx]), np.where(ops != 0, ud, cd[idx]), ops


def _write_inputs(tmp, scale, workload, sample, synth):
    n = 1 << scale
    cs, cd = synth.rmat(scale, 0, 16 << scale, 42)
    core = os.path.join(tmp, "core.bin")
    synth.write_triples(core, cs, cd, 1)
    upd = os.path.join(tmp, "upd.bin")
    us, ud, v = host_updates(scale, workload, 0, sample, synth, core=(cs, cd))
    synth.write_triples(upd, us, ud, v)
    return n, core, upd


def run_reference_once(n, core, upd, sample, threads, tmp):
    import oracle_py as O

    tpath = os.path.join(tmp, "timing.json")
    cmd = [O.REF_DRIVER, "--mode", "pppcsrnuma", "--api", "pool", "--threads", str(threads), "--ppd", "1",
           "--n", str(n), "--core", core, "--updates", upd, "--size", str(sample), "--timing", tpath]
    subprocess.run(cmd, stdout=subprocess.DEVNULL, check=True)
    return json.load(open(tpath))


def run_port_once(scale, workload, sample, synth):
    """No compiled reference on this box: time the C restatement (1 thread) instead."""
    import oracle_py as O

    n = 1 << scale
    cs, cd = synth.rmat(scale, 0, 16 << scale, 42)
    g = O.OraclePCSR(n)
    g.apply(cs, cd, 1)
    us, ud, v = host_updates(scale, workload, 0, sample, synth, core=(cs, cd))
    t0 = time.perf_counter()
    g.apply(us, ud, v)
    return {"update_ms": (time.perf_counter() - t0) * 1e3, "update_ops": sample}


def cpu_baseline(args, synth, steps=1, warmup=0):
    """Returns (cpu_baseline dict, ms_per_step).  kind 'reference' when oracle/_ref exists, else 'port'."""
    import oracle_py as O

    sample = min(args.cpu_sample, args.batch)
    what = (f"R-MAT scale-{args.scale} core loaded through the reference, then the first {sample} of the "
            f"{args.batch} {WORKLOAD_TEXT[args.workload]}; time = start()->stop() of the update phase")
    times = []
    if O.have_ref():
        threads = os.cpu_count() or 1
        with tempfile.TemporaryDirectory() as tmp:
            n, core, upd = _write_inputs(tmp, args.scale, args.workload, sample, synth)
            for i in range(warmup + steps):
                t = run_reference_once(n, core, upd, sample, threads, tmp)
                if i >= warmup:
                    times.append(t["update_ms"])
        kind, cores = "reference", threads
        what += f"; -pppcsrnuma -threads={threads} -partitions_per_domain=1, libnuma stubbed (1 domain)"
    else:
        for i in range(max(1, steps)):
            times.append(run_port_once(args.scale, args.workload, sample, synth)["update_ms"])
        kind, cores = "port", 1
    ms = float(np.mean(times))
    return {"value": sample / (ms / 1e3), "unit": UNIT, "cores": cores, "kind": kind, "sample": what}, ms


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    synth = importlib.import_module("parallel-packed-csr_b200.synth")
    base, ms = cpu_baseline(args, synth, steps=args.steps, warmup=min(args.warmup, 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if args.strong else "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


WORKLOAD_TEXT = {
    "insert": "uniform-random edge insertions",
    "delete": "edge deletions sampled from the core",
    "skewed": "skewed (R-MAT) edge insertions",
    "mixed": "mixed updates (3/4 uniform insertions, 1/4 deletions of core edges, per-update op)",
}


def global_shape(args, world):
    """(scale, updates per rank) of the run: weak scaling grows the graph with the ranks, strong splits it."""
    if args.strong:
        return args.scale, args.batch // world
    return args.scale + (world.bit_length() - 1), args.batch


def workload_config(args, world):
    scale, B = global_shape(args, world)
    return {
        "workload": f"R-MAT scale-{scale} core ({16 << scale} raw edges, a/b/c/d=.57/.19/.19/.05) + "
                    f"{B * world} {WORKLOAD_TEXT[args.workload]}, one batch of {B} per GPU per step",
        "batch_per_gpu": B, "scale": scale, "slot_bytes": SLOT_BYTES,
        "parallelism": "1 shard" if world == 1 else f"{world} vertex-range shards, updates routed to their owner "
                       "through NVLink peer memory (fused bin+scatter kernel; NCCL all-to-all when unavailable)",
        "l2": "shard state is restored from a device snapshot (>400 MB of writes, larger than the 126 MB L2) "
              "before every timed step; the working set (>=270 MB) also exceeds L2",
    }


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def main_b200(args):
    import torch

    pp = importlib.import_module("parallel-packed-csr_b200")
    synth = importlib.import_module("parallel-packed-csr_b200.synth")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 engine has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        # NCCL prints its version banner to stdout at NCCL_DEBUG=VERSION/WARN; rank 0 must print ONE JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

        dist = dist_mod
        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.current_stream()

    scale, B = global_shape(args, world)
    n = 1 << scale
    router = importlib.import_module("parallel-packed-csr_b200.router")

    # ---- core graph: every rank generates its slice of the global R-MAT stream, routes it to the owners
    core_total = 16 << scale
    lo, hi = rank * core_total // world, (rank + 1) * core_total // world
    cs, cd = synth.rmat(scale, lo, hi, 42, device=dev)
    cs, cd = cs.to(torch.int32), cd.to(torch.int32)
    if world > 1:
        # Shard cost model measured at N=1/2: ~0.13 us per routed update (sort + locate) and ~0.016 us per stored
        # item (window selection + rebalance).  A uniform stream sends B*world/n updates to every vertex, so a
        # vertex weighs ~8 * B*world/n "edges"; deletes follow the edge distribution instead.
        vw = 8.0 * B * world / n if args.workload in ("insert", "mixed") else 0.0
        starts = router.edge_balanced_starts(cs, n, world, dist, vertex_weight=vw)
    else:
        starts = np.array([0, n], dtype=np.uint64)
    peer_cap = 0 if os.environ.get("PPCSR_NO_PEER") else max(B, hi - lo)
    graph = router.ShardedGraph(n, starts, rank, world, local_rank, dist=dist, peer_cap=peer_cap,
                                peer_values=args.workload == "mixed")
    graph.shard.bind_torch_stream(stream)
    graph.apply(cs, cd, None, default_val=1)
    core_geo = graph.shard.geometry
    del cs, cd

    # ---- the update batch of this rank: its slice [rank*B, (rank+1)*B) of the workload's global stream
    uv = None  # per-update op (1 add / 0 delete) of the mixed stream
    default_val = 1
    if args.workload == "insert":
        us, ud = synth.uniform(scale, rank * B, (rank + 1) * B, 7, device=dev)
    elif args.workload == "skewed":
        us, ud = synth.rmat(scale, rank * B, (rank + 1) * B, 99, device=dev)
    else:
        # deletes: sampled without replacement from the raw core list; only the sampled edges are regenerated
        # (the stream is a pure function of the element index)
        idx = synth.sample_without_replacement(core_total, B * world, 7, device=dev)[rank * B:(rank + 1) * B]
        us, ud = synth.rmat_at(scale, idx, 42)
        del idx
        if args.workload == "delete":
            default_val = 0
        else:
            uv = synth.mixed_ops(rank * B, (rank + 1) * B, 11, device=dev)
            fs, fd = synth.uniform(scale, rank * B, (rank + 1) * B, 7, device=dev)
            us, ud = torch.where(uv != 0, fs, us), torch.where(uv != 0, fd, ud)
            uv = uv.to(torch.int32).contiguous()
            del fs, fd
    us, ud = us.to(torch.int32).contiguous(), ud.to(torch.int32).contiguous()
    graph.shard.reserve(max_slots=core_geo.N * 4, max_batch=int(B * 1.5) + 1024)
    graph.shard.snapshot()
    torch.cuda.synchronize()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing (value)
    stats_acc = []
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    for _ in range(args.warmup):
        graph.shard.restore()
        graph.apply(us, ud, uv, default_val=default_val)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for k in range(args.steps):
        graph.shard.restore()
        barrier()  # the untimed restore takes a different time on every shard: start the step together
        ev0[k].record(stream)
        st = graph.apply(us, ud, uv, default_val=default_val)
        ev1[k].record(stream)
        stats_acc.append(st)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    total_ms = sum(a.elapsed_time(b) for a, b in zip(ev0, ev1))
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = B * world * args.steps / (total_ms / 1e3)
    if world > 1 and os.environ.get("PPCSR_ROUTE_TIMING"):
        graph.route_timing = []
        for _ in range(3):
            graph.shard.restore()
            barrier()
            graph.apply(us, ud, uv, default_val=default_val)
        print(f"[rank {rank}] routing stages ms ([bin, counts, all-to-all, apply] over NCCL, [exchange, apply] over "
              f"peer memory): {graph.route_timing}", file=sys.stderr)
        graph.route_timing = None
    if world > 1:  # per-rank view (stderr): how many updates each shard received and where its time went
        s0 = stats_acc[-1]
        print(f"[rank {rank}] local ms/step {sum(a.elapsed_time(b) for a, b in zip(ev0, ev1)) / args.steps:.3f} "
              f"received {s0['batch_size']} apply {s0['ms_total']:.3f} ms (sort {s0['ms_sort']:.3f} locate "
              f"{s0['ms_locate']:.3f} select {s0['ms_select']:.3f} rebalance {s0['ms_rebalance']:.3f}) "
              f"windows {s0['n_windows']} N {s0['slots_before']}->{s0['slots_after']} "
              f"vertices {graph.n_local}", file=sys.stderr)

    # ---- end to end through the public host-buffer call: pinned host inputs, H2D inside the timed region
    hs = torch.empty(B, dtype=torch.int32).pin_memory()
    hd = torch.empty(B, dtype=torch.int32).pin_memory()
    hs.copy_(us)
    hd.copy_(ud)
    hv = None
    if uv is not None:
        hv = torch.empty(B, dtype=torch.int32).pin_memory()
        hv.copy_(uv)
    e2e_ms = 0.0
    e2e_steps = max(2, min(args.steps, 3))
    for k in range(1 + e2e_steps):
        graph.shard.restore()
        barrier()
        t0 = time.perf_counter()
        graph.apply_host(hs.numpy(), hd.numpy(), hv.numpy() if hv is not None else None, default_val=default_val)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) * 1e3
        if k > 0:
            e2e_ms += dt
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = B * world * e2e_steps / (float(t.item()) / 1e3)

    # ---- parity guard on the final state (cheap): invariants must hold
    lower = args.workload in ("delete", "mixed")
    rep = graph.shard.check(check_lower=lower)
    if rep.violations(lower):
        raise SystemExit(f"bench.py: PMA invariants violated after the timed steps: {rep.as_dict()}")

    # ---- optional edge scan: PageRank push steps over the updated graph (reference pagerank.h:16-29)
    pagerank = None
    if args.pagerank:
        vals = 1.0 + (torch.arange(n, device=dev, dtype=torch.float64) % 7)
        graph.pagerank_step(vals)
        barrier()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record(stream)
        for _ in range(args.steps):
            graph.pagerank_step(vals)
        p1.record(stream)
        barrier()
        pt = torch.tensor([p0.elapsed_time(p1) / args.steps], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(pt, op=dist.ReduceOp.MAX)
        geo = graph.shard.geometry
        pagerank = {"ms_per_step": float(pt.item()), "slots_rank0": int(geo.N),
                    "note": "one push step over every shard (+ one all-reduce of the fp64 vector when sharded)"}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    def mean(key):
        return float(np.mean([s[key] for s in stats_acc]))

    peak, peak_src = measured_peak()
    reb_bytes = mean("rebalance_bytes")
    reb_ms = mean("ms_rebalance_kernel")
    achieved = reb_bytes / (reb_ms / 1e3) / 1e9 if reb_ms > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            # ncu DRAM bytes of the roofline kernel, captured per workload (null where there is no capture)
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch", {}).get(
                f"{args.workload}/scale{args.scale}/batch{args.batch}") if world == 1 and not args.strong else None
        except Exception:
            traffic = None
    last = stats_acc[-1]
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong" if args.strong else "weak", "vs_baseline": None,
        "dtype": "u32", "data": "synthetic", "config": workload_config(args, world),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": (3 if hv is not None else 2) * 4 * B,
                "d2h_bytes_per_step": 2 * 128,
                "steps": e2e_steps},
        "gpu_launches": int(sum(s["kernel_launches"] for s in stats_acc)),
        "roofline": {"bound": "hbm", "kernel": "reb::k_rebalance_p", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": reb_bytes, "kernel_ms": reb_ms},
        "stages_ms": {k: mean(k) for k in ("ms_total", "ms_sort", "ms_locate", "ms_select", "ms_rebalance",
                                           "ms_rebalance_kernel")},
        "batch": {k: int(last[k]) for k in ("n_unique", "n_inserted", "n_overwritten", "n_deleted", "n_not_found",
                                            "n_windows", "window_slots", "slots_before", "slots_after", "resized",
                                            "whole_array")},
        "rebalance_bytes_per_update": reb_bytes / B,
        "hbm_roofline_updates_per_sec": peak * 1e9 / (16 + reb_bytes / B),
    }
    if pagerank is not None:
        line["pagerank"] = pagerank
    if world == 1 and not args.no_cpu_baseline:
        try:
            line["cpu_baseline"], _ = cpu_baseline(args, synth)
        except Exception as e:  # the baseline is reported context, never a reason to lose the GPU number
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {e}"}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


def _claim_stdout():
    """Rank 0 must print exactly ONE JSON line on stdout, but libraries (NCCL's version banner) write to fd 1
    directly.  Point fd 1 at stderr for the whole run and keep the real stdout for the result line."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real


if __name__ == "__main__":
    a = parse_args()
    _claim_stdout()
    rc = main_reference(a) if a.impl == "reference" else main_b200(a)
    sys.stdout.flush()
    sys.exit(rc)
