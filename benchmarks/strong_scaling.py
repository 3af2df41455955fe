#!/usr/bin/env python
"""GPU-count strong-scaling harness: the counterpart of the reference's
src/benchmarking/benchmark-strong-scaling.sh:83-156, with the GPU count in place of the core count.

The reference script runs its CLI for every core count / variant / partitions-per-domain / repetition, scrapes the
second `Elapsed wall clock time:` line (the update phase) and writes one CSV row per core count
    #CORES INS_<variant>0 .. INS_<variant>_Avg INS_<variant>_Stddev DEL_<variant>0 ..
plus a plot-data file `cores avg_ins stddev_ins avg_del stddev_del`.  Here every cell is one `bench.py` run
(a FIXED graph and batch split over N vertex-range shards, one process per GPU, launched exactly like the driver
launches bench.py), the scraped number is the time of the batch in milliseconds (`ms_per_step`), and the layout of
the two output files is the same, so the reference's gnuplot scripts read them unchanged:

    python benchmarks/strong_scaling.py --gpus 1 2 4 8 --reps 3 --scale 24 --batch 100000000 --out-prefix scaling

    scaling.csv :  #GPUS INS_SHARDS0 INS_SHARDS1 INS_SHARDS2 INS_SHARDS_Avg INS_SHARDS_Stddev DEL_SHARDS0 ...
    scaling.dat :  <gpus> <avg_ins> <stddev_ins> <avg_del> <stddev_del>

`--dry-run` prints the commands without running them.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def bench_command(gpus: int, workload: str, scale: int, batch: int, steps: int, warmup: int, port: int) -> list[str]:
    """The command line of one cell: plain python at one GPU, torchrun (one rank per GPU) above."""
    tail = [os.path.join(ROOT, "bench.py"), "--gpus", str(gpus), "--steps", str(steps), "--warmup", str(warmup),
            "--workload", workload, "--scale", str(scale), "--batch", str(batch), "--only-headline", "--no-cpu-baseline"]
    if gpus == 1:
        return [sys.executable] + tail
    return [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(gpus),
            "--master-addr", "127.0.0.1", "--master-port", str(port)] + tail


def scrape_ms(stdout: str) -> float:
    """ms of the batch from bench.py's JSON line (the last line of stdout that parses and carries `ms_per_step`)."""
    for line in reversed(stdout.strip().splitlines()):
        try:
            j = json.loads(line)
        except ValueError:
            continue
        if isinstance(j, dict) and "ms_per_step" in j:
            return float(j["ms_per_step"])
    raise ValueError("no bench.py JSON line in the output")


def avg_stddev(xs: list[float]) -> tuple[float, float]:
    """Mean and sample standard deviation, as the awk one-liner of the reference script (0 for one repetition)."""
    a = sum(xs) / len(xs)
    if len(xs) < 2:
        return a, 0.0
    return a, math.sqrt(sum((x - a) ** 2 for x in xs) / (len(xs) - 1))


def header(reps: int) -> str:
    cols = ["#GPUS"]
    for name in ("INS_SHARDS", "DEL_SHARDS"):
        cols += [f"{name}{r}" for r in range(reps)] + [f"{name}_Avg", f"{name}_Stddev"]
    return " ".join(cols)


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--gpus", type=int, nargs="+", default=[1, 2, 4, 8])
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--scale", type=int, default=24)
    ap.add_argument("--batch", type=int, default=100_000_000)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--out-prefix", default="strong_scaling")
    ap.add_argument("--port", type=int, default=29541)
    ap.add_argument("--dry-run", action="store_true")
    args = ap.parse_args(argv)

    rows_csv, rows_dat = [header(args.reps)], []
    for g in args.gpus:
        cells, dat = [], []
        for workload in ("insert", "delete"):
            times = []
            for r in range(args.reps):
                cmd = bench_command(g, workload, args.scale, args.batch, args.steps, args.warmup, args.port)
                print(f"[START]\t {workload}: repetition #{r + 1} on {g} GPUs: {' '.join(cmd)}", file=sys.stderr)
                if args.dry_run:
                    times.append(0.0)
                    continue
                out = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT)
                if out.returncode != 0:
                    sys.stderr.write(out.stderr[-2000:])
                    raise SystemExit(f"bench.py failed on {g} GPUs ({workload})")
                times.append(scrape_ms(out.stdout))
                print(f"[END]  \t {workload}: {times[-1]:.3f} ms", file=sys.stderr)
            a, sd = avg_stddev(times)
            cells += [f"{t:.4f}" for t in times] + [f"{a:.4f}", f"{sd:.4f}"]
            dat += [f"{a:.4f}", f"{sd:.4f}"]
        rows_csv.append(" ".join([str(g)] + cells))
        rows_dat.append(" ".join([str(g)] + dat))
    if args.dry_run:
        print("\n".join(rows_csv))
        return 0
    with open(args.out_prefix + ".csv", "w") as f:
        f.write("\n".join(rows_csv) + "\n")
    with open(args.out_prefix + ".dat", "w") as f:
        f.write("\n".join(rows_dat) + "\n")
    print("\n".join(rows_csv))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
