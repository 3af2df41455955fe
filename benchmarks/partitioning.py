#!/usr/bin/env python
"""Partitions-per-domain sweep: the counterpart of the reference's
src/benchmarking/benchmark-partitioning.sh:83-140.

The reference script runs its CLI for every partitions-per-domain value, both partitioned variants (-pppcsr,
-pppcsrnuma), insertions and deletions, REPETITIONS times; scrapes the SECOND `Elapsed wall clock time:` line (the
update phase, `sed -n '0~2p'`) and writes
    <base>_all_results.csv : #PARTITIONS INS_PPPCSR0 .. INS_PPPCSR_Avg INS_PPPCSR_Stddev DEL_PPPCSR0 .. INS_PPPCSR_NUMA0 ..
    <base>_plot_data.dat   : partitions ins del ins-NUMA del-NUMA
Here the executable is this repository's CLI (parallel-packed-csr_b200/host/parallel-packed-csr: same flags, same
stdout line), a "domain" is a GPU, and the files have the same layout, so the reference's gnuplot scripts read them
unchanged.  The wall-clock line has millisecond resolution, which a GPU batch often undercuts: a third file
`<base>_device_ms.csv` holds the device time of the update batch (the CLI's JSON line) in the same layout.

    python benchmarks/partitioning.py --scale 20 --size 1000000 --threads 8 --partitions 1 2 4 8 --reps 3
    python benchmarks/partitioning.py --core core.bin --insertions ins.bin --deletions del.bin --size 1000000 ...

Without input files an R-MAT core graph (seed 42), a uniform insertion stream (seed 7) and deletions sampled from the core
are written as binary pair files (`*.bin`) into --workdir.  `--dry-run` prints the commands.
"""
from __future__ import annotations

import argparse
import importlib
import json
import math
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VARIANTS = (("-pppcsr", "PPPCSR"), ("-pppcsrnuma", "PPPCSR_NUMA"))


def cli_command(exe: str, variant: str, delete: bool, threads: int, size: int, core: str, upd: str, ppd: int,
                extra: list[str]) -> list[str]:
    """Flag order as in the reference script (benchmark-partitioning.sh:106,122): -delete and -size come before the
    update file, which the CLI's parser needs (reference main.cpp:111-157)."""
    cmd = [exe]
    if delete:
        cmd.append("-delete")
    cmd += [f"-threads={threads}", variant, f"-size={size}", *extra, f"-core_graph={core}", f"-update_file={upd}",
            f"-partitions_per_domain={ppd}"]
    return cmd


def scrape(stdout: str) -> tuple[float, float]:
    """(wall ms of the update phase = the 2nd `Elapsed wall clock time:` line, device ms of the update batch)."""
    elapsed = [l.split(": ")[1] for l in stdout.splitlines() if l.startswith("Elapsed wall clock time: ")]
    if len(elapsed) < 2:
        raise ValueError("the CLI did not print two `Elapsed wall clock time` lines")
    dev = float("nan")
    for l in reversed(stdout.splitlines()):
        if l.startswith("{") and "device_ms" in l:
            dev = float(json.loads(l)["device_ms"])
            break
    return float(elapsed[1]), dev


def avg_stddev(xs: list[float]) -> tuple[float, float]:
    a = sum(xs) / len(xs)
    if len(xs) < 2:
        return a, 0.0
    return a, math.sqrt(sum((x - a) ** 2 for x in xs) / (len(xs) - 1))


def header(reps: int) -> str:
    cols = ["#PARTITIONS"]
    for _, name in VARIANTS:
        for op in ("INS", "DEL"):
            cols += [f"{op}_{name}{r}" for r in range(reps)] + [f"{op}_{name}_Avg", f"{op}_{name}_Stddev"]
    return " ".join(cols)


def write_inputs(workdir: str, scale: int, size: int) -> tuple[str, str, str]:
    import numpy as np

    synth = importlib.import_module("parallel-packed-csr_b200.synth")
    cs, cd = synth.rmat(scale, 0, 16 << scale, 42)
    us, ud = synth.uniform(scale, 0, size, 7)
    idx = synth.sample_without_replacement(16 << scale, size, 7)
    paths = [os.path.join(workdir, n) for n in ("core.bin", "insertions.bin", "deletions.bin")]
    for p, (s, d) in zip(paths, ((cs, cd), (us, ud), (cs[idx], cd[idx]))):
        np.stack([s, d], axis=1).astype("<u4").tofile(p)
    return tuple(paths)


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--exe", default=os.path.join(ROOT, "parallel-packed-csr_b200", "host", "parallel-packed-csr"))
    ap.add_argument("--core"), ap.add_argument("--insertions"), ap.add_argument("--deletions")
    ap.add_argument("--scale", type=int, default=20)
    ap.add_argument("--size", type=int, default=1_000_000)
    ap.add_argument("--threads", type=int, default=8)
    ap.add_argument("--partitions", type=int, nargs="+", default=[1, 2, 4, 8])
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--balanced", action="store_true", help="pass -balanced (edge-balanced partition boundaries)")
    ap.add_argument("--out-prefix", default="partitioning")
    ap.add_argument("--workdir", default=None)
    ap.add_argument("--dry-run", action="store_true")
    a = ap.parse_args(argv)
    extra = ["-balanced"] if a.balanced else []
    tmp = None
    core, ins, dele = a.core, a.insertions, a.deletions
    if not (core and ins and dele):
        if a.dry_run:
            core, ins, dele = "core.bin", "insertions.bin", "deletions.bin"
        else:
            tmp = tempfile.TemporaryDirectory(dir=a.workdir)
            core, ins, dele = write_inputs(tmp.name, a.scale, a.size)
    rows, rows_dev, dat = [header(a.reps)], [header(a.reps)], ["partitions ins del ins-NUMA del-NUMA"]
    for p in a.partitions:
        cells, cells_dev, avgs = [], [], []
        for flag, _ in VARIANTS:
            for delete, upd in ((False, ins), (True, dele)):
                wall, dev = [], []
                for r in range(a.reps):
                    cmd = cli_command(a.exe, flag, delete, a.threads, a.size, core, upd, p, extra)
                    print(f"[START]\t {flag[1:]} edge {'deletions' if delete else 'insertions'}: repetition #{r + 1}, "
                          f"{p} partitions per domain: {' '.join(cmd)}", file=sys.stderr)
                    if a.dry_run:
                        wall.append(0.0), dev.append(0.0)
                        continue
                    out = subprocess.run(cmd, capture_output=True, text=True)
                    if out.returncode != 0:
                        sys.stderr.write(out.stdout[-2000:] + out.stderr[-2000:])
                        raise SystemExit(f"the CLI failed ({' '.join(cmd)})")
                    w, d = scrape(out.stdout)
                    wall.append(w), dev.append(d)
                for xs, out_cells in ((wall, cells), (dev, cells_dev)):
                    m, sd = avg_stddev(xs)
                    out_cells += [f"{x:g}" for x in xs] + [f"{m:g}", f"{sd:g}"]
                avgs.append(f"{avg_stddev(wall)[0]:g}")
        rows.append(" ".join([str(p)] + cells))
        rows_dev.append(" ".join([str(p)] + cells_dev))
        dat.append(" ".join([str(p)] + avgs))
    if not a.dry_run:
        for suffix, content in (("_all_results.csv", rows), ("_plot_data.dat", dat), ("_device_ms.csv", rows_dev)):
            with open(a.out_prefix + suffix, "w") as f:
                f.write("\n".join(content) + "\n")
    print("\n".join(rows))
    if tmp is not None:
        tmp.cleanup()
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
