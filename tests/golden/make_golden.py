"""Generate the committed golden fixtures from the UNMODIFIED reference (oracle/_ref/ref_driver,
built from /root/reference by oracle/Makefile).  Run in the build container only:

    python tests/golden/make_golden.py

Each fixture is an .npz holding the inputs (core + update triples), the reference's resulting
logical graph (rowptr / col / num_neighbors), its physical geometry (N, logN, H) for sequential
runs, and one reference pagerank<T,double> push step.  The reference ships no golden vectors of
its own, so these are outputs of the reference itself run here.
"""
from __future__ import annotations

import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_py as O  # noqa: E402

_spec = importlib.util.spec_from_file_location("synth", os.path.join(ROOT, "parallel-packed-csr_b200", "synth.py"))
synth = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(synth)


def save(name, n, core, upd, ref, **extra):
    def trip(t):
        s, d, v = t
        s = np.asarray(s, dtype=np.uint32)
        return s, np.asarray(d, dtype=np.uint32), np.ascontiguousarray(np.broadcast_to(np.asarray(v, dtype=np.uint32), s.shape))

    cs, cd, cv = trip(core)
    us, ud, uv = trip(upd)
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"), n=np.uint64(n), core_src=cs, core_dst=cd, core_val=cv,
        upd_src=us, upd_dst=ud, upd_val=uv, rowptr=ref["rowptr"], col=ref["col"],
        num_neighbors=ref["num_neighbors"],
        pagerank=ref["pagerank"] if ref.get("pagerank") is not None else np.zeros(0),
        geometry=np.array([ref["N"], ref["logN"], ref["H"]], dtype=np.uint64), **extra)
    print(name, "n", n, "E", ref["E"], "geometry", ref["N"], ref["logN"], ref["H"])


def main():
    assert O.have_ref(), "build oracle/_ref first: make -C oracle ref"
    scale = 11
    n = 1 << scale
    s, d = synth.rmat(scale, 0, 16 << scale, 42)
    core = (s, d, 1)

    us, ud = synth.uniform(scale, 0, 6000, 7)
    save("rmat11_insert_uniform", n, core, (us, ud, 1), O.run_ref(n, core, (us, ud, 1), api="direct", pagerank=True))

    ks, kd = synth.rmat(scale, 0, 6000, 99)
    save("rmat11_insert_skewed", n, core, (ks, kd, 1), O.run_ref(n, core, (ks, kd, 1), api="direct", pagerank=True))

    di = synth.sample_without_replacement(16 << scale, 12000, 7)
    dele = (s[di], d[di], 0)
    save("rmat11_delete", n, core, dele, O.run_ref(n, core, dele, api="direct", pagerank=True))

    # mixed adds (value = op index) and deletes, 3:1, on 1000 vertices -- the shape of the reference's
    # add_remove_edge_random_2E4_seq test (test/DataStructureTest.cpp:122-144)
    m = 20000
    ms, md = synth.uniform(10, 0, m, 3)
    ms, md = ms % 1000, md % 1000
    mv = np.where(synth.mixed_ops(0, m, 3) != 0, np.arange(1, m + 1), 0)
    empty = (ms[:0], md[:0], 1)
    save("mixed1000_seq", 1000, empty, (ms, md, mv), O.run_ref(1000, empty, (ms, md, mv), api="direct", pagerank=True))

    # one hub vertex: 1e4 ascending inserts then 1e4 deletes (test/DataStructureTest.cpp:51-79)
    hs = np.zeros(10000, dtype=np.int64)
    hd = np.arange(1, 10001)
    save("hub_insert", 10, empty, (hs, hd, hd), O.run_ref(10, empty, (hs, hd, hd), api="direct"))
    save("hub_insert_delete_all", 10, (hs, hd, hd), (hs, hd, 0), O.run_ref(10, (hs, hd, hd), (hs, hd, 0), api="direct"))
    save("hub_insert_delete_half", 10, (hs, hd, hd), (hs[:5000], hd[::-1][:5000], 0),
         O.run_ref(10, (hs, hd, hd), (hs[:5000], hd[::-1][:5000], 0), api="direct"))

    # the thread pools (value always 1), multi-threaded and partitioned: logical graph only
    save("rmat11_pool_ppcsr_t8", n, core, (us, ud, 1),
         O.run_ref(n, core, (us, ud, 1), mode="ppcsr", api="pool", threads=8, pagerank=False))
    save("rmat11_pool_pppcsr_ppd4_t8", n, core, dele,
         O.run_ref(n, core, dele, mode="pppcsr", api="pool", threads=8, ppd=4, pagerank=False))


if __name__ == "__main__":
    main()
