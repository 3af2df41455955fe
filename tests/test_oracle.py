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
