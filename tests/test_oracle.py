"""The oracle (oracle/pcsr_oracle.c, a C restatement of the reference's sequential algorithm) is
pinned here against (a) the committed golden fixtures produced by the unmodified reference and
(b) live runs of the compiled reference when oracle/_ref exists (build container only)."""
import glob
import os

import numpy as np
import pytest

import oracle_py as O

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "*.npz")))


def _apply_fixture(fx):
    g = O.OraclePCSR(int(fx["n"]))
    g.apply(fx["core_src"], fx["core_dst"], fx["core_val"])
    g.apply(fx["upd_src"], fx["upd_dst"], fx["upd_val"])
    return g


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_matches_reference_golden(path):
    fx = np.load(path)
    g = _apply_fixture(fx)
    rowptr, col, nn = g.export()
    assert np.array_equal(rowptr, fx["rowptr"])
    assert np.array_equal(col, fx["col"])
    threaded = "_pool_" in path  # num_neighbors is racy in the reference at threads > 1
    if not threaded:
        assert np.array_equal(nn, fx["num_neighbors"])
        assert tuple(int(x) for x in fx["geometry"]) == g.geometry
    if fx["pagerank"].size:
        pr = g.pagerank(1.0 + (np.arange(int(fx["n"])) % 7))
        assert np.array_equal(pr, fx["pagerank"], equal_nan=True)
    assert g.check() == 0


def test_fixture_inputs_match_generator():
    """The synthetic generators are pure functions of (seed, index): fixtures must be reproducible."""
    import importlib

    synth = importlib.import_module("parallel-packed-csr_b200.synth")
    fx = np.load([p for p in GOLDEN if p.endswith("rmat11_insert_uniform.npz")][0])
    s, d = synth.rmat(11, 0, 16 << 11, 42)
    assert np.array_equal(s.astype(np.uint32), fx["core_src"]) and np.array_equal(d.astype(np.uint32), fx["core_dst"])
    us, ud = synth.uniform(11, 0, 6000, 7)
    assert np.array_equal(us.astype(np.uint32), fx["upd_src"]) and np.array_equal(ud.astype(np.uint32), fx["upd_dst"])


def test_geometry_formula():
    # reference src/pcsr/PCSR.cpp:68-73,777: N = 2 << bsr(max(2n,1024)), logN = 1 << bsr(2*bsr(N)+1)
    for n, geo in [(10, (2048, 16, 7)), (1000, (2048, 16, 7)), (1024, (4096, 16, 8)), (65536, (262144, 32, 13)), (1 << 20, (1 << 22, 32, 17))]:
        assert O.OraclePCSR(n).geometry == geo


def test_reference_unit_test_postconditions():
    """Restated expectations of reference test/DataStructureTest.cpp:12-49."""
    g = O.OraclePCSR(10)
    assert g.n == 10
    g.add_edge(11, 1, 1)  # no such vertex: silently ignored (:29)
    g.add_edge(0, 1, 1)
    assert g.edge_exists(0, 1) and len(g.neighbourhood(0)) == 1 and len(g.neighbourhood(2)) == 0
    g.remove_edge(0, 1)
    assert not g.edge_exists(0, 1)
    g.remove_edge(0, 1)  # absent: harmless
    assert g.not_found == 1
    e = O.OraclePCSR(0)
    assert e.n == 0
    e.add_node()
    assert e.n == 1 and len(e.neighbourhood(0)) == 0


def test_partition_and_domain_tables():
    # reference src/pppcsr/PPPCSR.cpp:13-34: equal vertex ranges, remainder to the last partition
    starts, sizes = O.partition_table(10, 4)
    assert starts.tolist() == [0, 2, 4, 6] and sizes.tolist() == [2, 2, 2, 4]
    assert [O.partition_owner(starts, v) for v in (0, 1, 2, 5, 6, 9, 100)] == [0, 0, 1, 2, 3, 3, 3]
    # reference test/SchedulerTest.cpp:11-58
    for d in range(1, 9):
        for t in range(1, 257):
            t2d, first, num = O.domain_table(t, d)
            assert num.sum() == t and first[0] == 0
            assert len(set(t2d.tolist())) == min(t, d) and max(t2d) < d
            assert 1 <= len(set(num.tolist())) <= 2


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_matches_live_reference_random_streams():
    rng = np.random.default_rng(5)
    for n, m in [(50, 4000), (700, 30000)]:
        src = rng.integers(0, n, m)
        dst = rng.integers(0, n, m)
        val = np.where(rng.integers(0, 3, m) != 0, rng.integers(1, 1000, m), 0)
        empty = (src[:0], dst[:0], 1)
        ref = O.run_ref(n, empty, (src, dst, val), api="direct", pagerank=True)
        g = O.OraclePCSR(n)
        g.apply(src, dst, val)
        rowptr, col, nn = g.export()
        assert np.array_equal(rowptr, ref["rowptr"]) and np.array_equal(col, ref["col"])
        assert np.array_equal(nn, ref["num_neighbors"])
        assert g.geometry == (ref["N"], ref["logN"], ref["H"])
        assert np.array_equal(g.pagerank(1.0 + (np.arange(n) % 7)), ref["pagerank"], equal_nan=True)


def test_graph_checksum_host_statement():
    """synth.graph_checksum (the host statement of ppcsr_checksum / ref_driver --checksum) on a tiny CSR, by hand."""
    import importlib

    synth = importlib.import_module("parallel-packed-csr_b200.synth")

    def mix(x):
        m = (1 << 64) - 1
        x ^= x >> 30
        x = (x * 0xBF58476D1CE4E5B9) & m
        x ^= x >> 27
        x = (x * 0x94D049BB133111EB) & m
        return x ^ (x >> 31)

    rowptr, col, nn = [0, 2, 2, 3], [1, 2, 0], [5, 0, 1]
    want_e = (mix((7 << 32) | 1) + mix((7 << 32) | 2) + mix((9 << 32) | 0)) & ((1 << 64) - 1)
    want_n = (5 * mix(7) + 1 * mix(9)) & ((1 << 64) - 1)
    got = synth.graph_checksum(rowptr, col, nn, vertex_offset=7)
    assert got == {"edges": 3, "edge_hash": want_e, "nn_hash": want_n}


def test_c4_golden_checksum_fixture():
    """tests/golden/c4_checksum.json (the reference's checksum of the full-size C4 graph, bench.py's parity anchor):
    the call-count hash is reproducible from the streams, the edge count is consistent with it."""
    import json

    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c4_checksum.json")))
    assert g["workload"] == {"scale": 24, "batch": 100_000_000, "stream": "skewed", "core_edges": 16 << 24,
                             "core_seed": 42, "update_seed": 99}
    assert 0 < g["edges"] < (16 << 24) + 100_000_000 and len(g["edge_hash"]) == 16 and len(g["nn_hash_call_count"]) == 16


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built")
def test_ref_driver_synthetic_streams_and_checksum(tmp_path):
    """ref_driver's built-in stream generator is bit-identical to synth.py, its --checksum equals the host statement
    over its own dump, and a --checkpoint dump equals a run that stops there."""
    import importlib
    import json
    import subprocess

    synth = importlib.import_module("parallel-packed-csr_b200.synth")
    for kind, scale, lo, hi, seed in (("rmat", 24, 5, 20005, 99), ("uniform", 20, 0, 20000, 7),
                                      ("rmat", 16, (1 << 32) - 10, (1 << 32) + 500, 42)):
        p = str(tmp_path / "emit.bin")
        subprocess.run([O.REF_DRIVER, "--threads", "3", "--synth-updates", f"{kind}:{scale}:{lo}:{hi}:{seed}",
                        "--emit-updates", p], check=True)
        a = np.fromfile(p, dtype="<u4").reshape(-1, 3)
        s, d = (synth.rmat if kind == "rmat" else synth.uniform)(scale, lo, hi, seed)
        assert np.array_equal(a[:, 0], s) and np.array_equal(a[:, 1], d) and (a[:, 2] == 1).all()
    d1, d2, d3, sp = (str(tmp_path / x) for x in ("ck.bin", "full.bin", "stop.bin", "sum.json"))
    base = [O.REF_DRIVER, "--mode", "pppcsrnuma", "--api", "pool", "--threads", "1", "--ppd", "2", "--n", str(1 << 12),
            "--synth-core", "rmat:12:0:65536:42", "--synth-updates", "rmat:12:0:20000:99"]
    subprocess.run(base + ["--checkpoint", "5000", d1, "--dump", d2, "--checksum", sp], check=True, stdout=subprocess.DEVNULL)
    subprocess.run(base + ["--size", "5000", "--dump", d3], check=True, stdout=subprocess.DEVNULL)
    ck, full, stop = O.read_dump(d1), O.read_dump(d2), O.read_dump(d3)
    assert np.array_equal(ck["rowptr"], stop["rowptr"]) and np.array_equal(ck["col"], stop["col"])
    cs = json.load(open(sp))
    want = synth.graph_checksum(full["rowptr"], full["col"], full["num_neighbors"])
    assert cs["edges"] == want["edges"] and int(cs["edge_hash"], 16) == want["edge_hash"]
    assert int(cs["nn_hash"], 16) == want["nn_hash"]  # threads = 1: the call counts are exact
    o = O.OraclePCSR(1 << 12)
    o.apply(*synth.rmat(12, 0, 65536, 42), 1)
    o.apply(*synth.rmat(12, 0, 20000, 99), 1)
    rp, col, nn = o.export()
    assert np.array_equal(rp, full["rowptr"]) and np.array_equal(col, full["col"]) and np.array_equal(nn, full["num_neighbors"])
