"""Generates tests/golden/c4_checksum.json: the checksum of the logical graph the UNMODIFIED reference
(oracle/_ref/ref_driver, built from /root/reference by oracle/Makefile) produces for BASELINE.json configs[3] at FULL
size -- R-MAT scale-24 core (268 435 456 raw edges, seed 42) + 100 000 000 skewed (R-MAT, seed 99) insertions --
through its own ThreadPoolPPPCSR (-pppcsrnuma).  bench.py compares the sum of the shards' ppcsr_checksum() with it
after the timed steps, at every GPU count.

  python tests/golden/make_c4_checksum.py [--threads 8]        (build container; ~10 min on 8 vCPUs, ~25 GB of RAM)

edges / edge_hash come from the reference run (get_neighbourhood of every vertex).  The reference's num_neighbors is
updated non-atomically and differs from run to run at threads > 1, so nn_hash_call_count is the
call-count rule itself (reference PCSR.cpp:1392: +1 per accepted add call, duplicates included) evaluated with numpy
over the two streams; the reference's own (racy) nn_hash is recorded beside it for information."""
import argparse
import importlib
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SCALE, CORE, BATCH, CORE_SEED, SKEW_SEED = 24, 16 << 24, 100_000_000, 42, 99


def call_count_hash(synth):
    n = 1 << SCALE
    nn = np.zeros(n, dtype=np.int64)
    for total, seed in ((CORE, CORE_SEED), (BATCH, SKEW_SEED)):
        for lo in range(0, total, 1 << 24):
            s, _ = synth.rmat(SCALE, lo, min(total, lo + (1 << 24)), seed)
            nn += np.bincount(s, minlength=n)
    with np.errstate(over="ignore"):
        h = int((nn.astype(np.uint64) * synth.mix64(np.arange(n, dtype=np.uint64))).sum(dtype=np.uint64))
    return h, int(nn.sum())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--from-run", default=None, help="directory holding sum.json / timing.json of a finished run")
    a = ap.parse_args()
    import oracle_py as O

    synth = importlib.import_module("parallel-packed-csr_b200.synth")
    if a.from_run:
        cs = json.load(open(os.path.join(a.from_run, "sum.json")))
        tm = json.load(open(os.path.join(a.from_run, "timing.json")))
    else:
        with tempfile.TemporaryDirectory() as td:
            sp, tp = os.path.join(td, "sum.json"), os.path.join(td, "timing.json")
            subprocess.run([O.REF_DRIVER, "--mode", "pppcsrnuma", "--api", "pool", "--threads", str(a.threads), "--ppd", "1",
                            "--n", str(1 << SCALE), "--synth-core", f"rmat:{SCALE}:0:{CORE}:{CORE_SEED}",
                            "--synth-updates", f"rmat:{SCALE}:0:{BATCH}:{SKEW_SEED}", "--timing", tp, "--checksum", sp],
                           check=True, stdout=subprocess.DEVNULL)
            cs, tm = json.load(open(sp)), json.load(open(tp))
    nn_hash, calls = call_count_hash(synth)
    assert calls == CORE + BATCH
    out = {
        "workload": {"scale": SCALE, "batch": BATCH, "stream": "skewed", "core_edges": CORE, "core_seed": CORE_SEED,
                     "update_seed": SKEW_SEED},
        "produced_by": f"oracle/_ref/ref_driver (unmodified reference, ThreadPoolPPPCSR -pppcsrnuma, {tm['threads']} "
                       f"threads, libnuma stubbed) via tests/golden/make_c4_checksum.py",
        "edges": cs["edges"], "edge_hash": cs["edge_hash"],
        "nn_hash_call_count": f"{nn_hash:016x}", "nn_hash_reference_run_racy": cs["nn_hash"],
        "reference_timing_ms": {"core": tm["core_ms"], "updates": tm["update_ms"], "threads": tm["threads"]},
    }
    with open(os.path.join(HERE, "c4_checksum.json"), "w") as f:
        json.dump(out, f, indent=1)
        f.write("\n")
    print(json.dumps(out))


if __name__ == "__main__":
    main()
