"""GPU tests of the host side: the reference's unit tests restated against the C++ classes (host/test_host),
the CLI's flag/stdout contract, the owner-binning kernel and -- with >= 2 GPUs -- the NCCL all-to-all path."""
import importlib
import os
import socket
import subprocess

import numpy as np
import pytest

import oracle_py as O

pp = importlib.import_module("parallel-packed-csr_b200")
build = importlib.import_module("parallel-packed-csr_b200.build")
router = importlib.import_module("parallel-packed-csr_b200.router")
synth = importlib.import_module("parallel-packed-csr_b200.synth")

pytestmark = pytest.mark.gpu


def test_reference_unit_tests_on_cpp_classes():
    build.build_host()
    r = subprocess.run([build.HOST_TEST, "--quick"], capture_output=True, text=True, timeout=900)
    tail = "\n".join(r.stdout.splitlines()[-15:])
    assert r.returncode == 0, tail + r.stderr[-2000:]
    assert "0 failed" in r.stdout


@pytest.mark.parametrize("mode,op", [("-ppcsr", "-insert"), ("-pppcsr", "-delete"), ("-pppcsrnuma", "-insert")])
def test_cli_contract(tmp_path, mode, op):
    """Same flags, same order sensitivity, same scraped stdout lines as reference src/main.cpp:111-189."""
    build.build_host()
    scale = 10
    cs, cd = synth.rmat(scale, 0, 16 << scale, 42)
    core, upd = str(tmp_path / "core.txt"), str(tmp_path / "upd.txt")
    synth.write_text(core, cs, cd)
    if op == "-insert":
        us, ud = synth.uniform(scale, 0, 3000, 7)
    else:
        idx = synth.sample_without_replacement(16 << scale, 3000, 7)
        us, ud = cs[idx], cd[idx]
    synth.write_text(upd, us, ud)
    cmd = [build.CLI, "-threads=8", op, "-size=2000", mode, "-partitions_per_domain=2", f"-core_graph={core}",
           f"-update_file={upd}", "-check"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    lines = r.stdout.splitlines()
    elapsed = [l for l in lines if l.startswith("Elapsed wall clock time: ")]
    assert len(elapsed) == 2 and all(l.split(": ")[1].isdigit() for l in elapsed)  # scripts keep the 2nd line
    assert f"Core graph size: {16 << scale}" in lines
    assert any(l.startswith("Edges: ") and " logN: " in l and " #count: " in l for l in lines)
    assert "PMA invariants: ok" in lines
    if mode != "-ppcsr":
        assert any(l.startswith("Number of partitions: ") for l in lines)
    # missing files -> the reference's messages and a non-zero exit
    r2 = subprocess.run([build.CLI, "-ppcsr"], capture_output=True, text=True)
    assert r2.returncode != 0 and "Core graph file not specified" in r2.stdout


def test_bin_by_owner_kernel_matches_torch():
    import torch

    dev = torch.device("cuda", 0)
    n, parts, count = 100_000, 5, 300_007
    rng = np.random.default_rng(1)
    starts = np.sort(np.concatenate([[0], rng.choice(np.arange(1, n), parts - 1, replace=False), [n]])).astype(np.uint64)
    src = torch.from_numpy(rng.integers(0, n, count).astype(np.int32)).to(dev)
    dst = torch.from_numpy(rng.integers(0, 1 << 31, count).astype(np.int32)).to(dev)
    val = torch.from_numpy(rng.integers(0, 9, count).astype(np.int32)).to(dev)
    starts_dev = torch.from_numpy(starts.astype(np.int64)).to(dev)
    got = router.CudaBinner(0)(starts_dev, parts, src, dst, val)
    torch.cuda.synchronize()
    want = router.TorchBinner()(starts_dev, parts, src, dst, val)
    assert got[3] == want[3] == router.split_counts_by_owner(src.cpu().numpy(), starts).tolist()
    for a, b in zip(got[:3], want[:3]):
        assert torch.equal(a, b)  # stable: submission order kept inside every owner's run


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _mixed_stream(scale, lo, hi, device=None):
    """Updates [lo, hi) of a mixed stream: 3/4 uniform inserts, 1/4 deletes of core edges, with the op as value."""
    total = 16 << scale
    ops = synth.mixed_ops(lo, hi, 11, device=device)
    fs, fd = synth.uniform(scale, lo, hi, 13, device=device)
    idx = synth.sample_without_replacement(total, hi, 5, device=device)[lo:hi]
    ds, dd = synth.rmat_at(scale, idx, 42)
    if device is None:
        return np.where(ops != 0, fs, ds), np.where(ops != 0, fd, dd), ops
    import torch

    return torch.where(ops != 0, fs, ds), torch.where(ops != 0, fd, dd), ops


def _nccl_worker(rank, world, port, scale, n_upd, out, transport):
    import torch
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        n = 1 << scale
        total = 16 << scale
        cs, cd = synth.rmat(scale, rank * total // world, (rank + 1) * total // world, 42, device=dev)
        cs, cd = cs.to(torch.int32), cd.to(torch.int32)
        starts = router.edge_balanced_starts(cs, n, world, dist)
        g = router.ShardedGraph(n, starts, rank, world, rank, dist=dist,
                                peer_cap=total // world if transport == "peer" else 0, peer_values=True)
        if transport == "peer":
            assert g.peer is not None, "symmetric memory (NVLink peer routing) is not available on this box"
        g.apply(cs, cd)
        us, ud = synth.uniform(scale, rank * n_upd // world, (rank + 1) * n_upd // world, 7, device=dev)
        g.apply(us.to(torch.int32), ud.to(torch.int32))
        ms, md, mv = _mixed_stream(scale, rank * n_upd // world, (rank + 1) * n_upd // world, device=dev)
        g.apply(ms.to(torch.int32), md.to(torch.int32), mv.to(torch.int32))
        used = (g.last_route or {}).get("transport", "nccl")
        rep = g.shard.check(False)
        rowptr, col = g.shard.export()
        vals = torch.from_numpy(1.0 + (np.arange(n) % 7)).to(dev)
        pr = g.pagerank_step(vals).cpu().numpy()
        np.savez(out.format(rank=rank), rowptr=rowptr, col=col, nn=g.shard.num_neighbors(), starts=starts,
                 bad=int(bool(rep.violations(False))), pr=pr, peer=int(used == "peer"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("transport", ["peer", "nccl"])
def test_two_gpu_all_to_all_vs_oracle(tmp_path, transport):
    """2 shards on 2 GPUs: updates routed to their owner (NVLink peer-memory scatter, or the NCCL all-to-all), the
    result compared with the sequential oracle on the unsharded graph: adjacency, num_neighbors, PageRank."""
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world, scale, n_upd = 2, 12, 40000
    out = str(tmp_path / "rank{rank}.npz")
    mp.spawn(_nccl_worker, args=(world, _free_port(), scale, n_upd, out, transport), nprocs=world, join=True)
    n = 1 << scale
    cs, cd = synth.rmat(scale, 0, 16 << scale, 42)
    us, ud = synth.uniform(scale, 0, n_upd, 7)
    o = O.OraclePCSR(n)
    o.apply(cs, cd, 1)
    o.apply(us, ud, 1)
    # the ranks' slices of the mixed stream are applied as ONE batch in rank order: keys are routed by source, and
    # the per-key order inside the batch is (rank, position), i.e. the order of the concatenated global stream
    ms, md, mv = _mixed_stream(scale, 0, n_upd)
    o.apply(ms, md, mv)
    rowptr, col, nn = o.export()
    opr = o.pagerank(1.0 + (np.arange(n) % 7))
    for r in range(world):
        z = np.load(out.format(rank=r))
        lo, hi = int(z["starts"][r]), int(z["starts"][r + 1])
        assert int(z["bad"]) == 0
        assert int(z["peer"]) == int(transport == "peer")
        assert np.array_equal(z["rowptr"], rowptr[lo:hi + 1] - rowptr[lo])
        assert np.array_equal(z["col"], col[int(rowptr[lo]):int(rowptr[hi])])
        assert np.array_equal(z["nn"], nn[lo:hi])
        fin = np.isfinite(opr)
        assert np.allclose(z["pr"][fin], opr[fin], rtol=1e-6, atol=0)


def _reference_read_input(text: str, default_val: int):
    """reference src/main.cpp:29-62 (read_input) restated: stoi / substr(pos + 1) / the op character."""
    import re

    num = re.compile(r"\s*[+-]?\d+")
    src, dst, val = [], [], []
    for line in text.split("\n"):
        m = num.match(line)
        if not m:
            continue
        m2 = num.match(line[m.end() + 1:])
        if not m2:
            continue
        pos, pos2 = m.end(), m2.end()
        v = default_val
        at = pos + 1 + pos2 + 1
        if at < len(line):
            v = 1 if line[at] == "1" else 0 if line[at] == "0" else default_val
        src.append(int(m.group()) & 0xFFFFFFFF)
        dst.append(int(m2.group()) & 0xFFFFFFFF)
        val.append(v)
    return np.array(src, dtype=np.uint32), np.array(dst, dtype=np.uint32), np.array(val, dtype=np.uint32)


def test_gpu_text_parser_matches_reference_reader():
    """ppcsr_parse_edge_list (GPU) against the reference reader's rules on separators, op columns, CRLF, blank and
    malformed lines, a missing final newline, and a large random file."""
    rng = np.random.default_rng(5)
    small = "1 2\n3\t4 1\n5,6,0\n\n7 8 x\n  9   10\n11 12 1\r\n13 14\r\nabc\n15\n16 17"
    big_s, big_d = rng.integers(0, 1 << 20, 300_000), rng.integers(0, 1 << 20, 300_000)
    big_o = rng.integers(0, 3, 300_000)
    big = "".join(f"{s} {d}\n" if o == 2 else f"{s} {d} {o}\n" for s, d, o in zip(big_s, big_d, big_o))
    for text, dv in ((small, 1), (small, 0), (big, 1), ("", 1), ("42 43", 0)):
        s, d, v, parsed, top = pp.parse_edge_list(text.encode(), default_val=dv)
        keep = s != 0xFFFFFFFF  # lines without a parsable pair come out as (SENT, SENT)
        rs, rd, rv = _reference_read_input(text, dv)
        assert parsed == rs.size and np.array_equal(s[keep], rs) and np.array_equal(d[keep], rd)
        assert np.array_equal(v[keep], rv)
        assert top == (max(int(rs.max()), int(rd.max())) if rs.size else 0)


def test_text_binary_and_array_inputs_give_the_same_graph(tmp_path):
    """the CLI fed with the text file (GPU parser), with the host parser, and with the binary
    pair file builds the same graph as the arrays through the C-ABI, which equals the oracle's."""
    build.build_host()
    scale = 11
    n = 1 << scale
    cs, cd = synth.rmat(scale, 0, 16 << scale, 42)
    us, ud = synth.rmat(scale, 0, 9000, 99)
    o = O.OraclePCSR(n)
    o.apply(cs, cd, 1)
    o.apply(us, ud, 1)
    rowptr, col, _ = o.export()
    # arrays vs interleaved pairs through the C-ABI
    a, b = pp.Shard(n), pp.Shard(n)
    a.apply(cs, cd, 1)
    a.apply(us, ud, 1)
    b.apply_pairs(np.stack([cs, cd], axis=1))
    b.apply_pairs(np.stack([us, ud], axis=1))
    for g in (a, b):
        rp, c = g.export()
        assert np.array_equal(rp, rowptr) and np.array_equal(c, col)
    want = a.checksum()
    # the CLI: text (GPU parse), text (host parse), binary pairs -- same checksum line
    core_t, upd_t = str(tmp_path / "core.txt"), str(tmp_path / "upd.txt")
    core_b, upd_b = str(tmp_path / "core.bin"), str(tmp_path / "upd.bin")
    synth.write_text(core_t, cs, cd)
    synth.write_text(upd_t, us, ud)
    np.stack([cs, cd], axis=1).astype("<u4").tofile(core_b)
    np.stack([us, ud], axis=1).astype("<u4").tofile(upd_b)
    outs = []
    for extra, core, upd in (([], core_t, upd_t), (["-host_parse"], core_t, upd_t), ([], core_b, upd_b)):
        cmd = [build.CLI, "-threads=4", "-insert", "-size=9000", "-ppcsr", *extra, f"-core_graph={core}",
               f"-update_file={upd}", "-check", "-checksum"]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        line = [l for l in r.stdout.splitlines() if l.startswith("Graph checksum: ")]
        assert len(line) == 1 and "PMA invariants: ok" in r.stdout
        outs.append(line[0])
    assert outs[0] == outs[1] == outs[2]
    assert outs[0] == f"Graph checksum: edges {want['edges']} edge_hash {want['edge_hash']:016x}"


@pytest.mark.parametrize("parts", [1, 3, 4])
def test_group_route_on_one_gpu(parts):
    """ppcsr_group_* (the C++ data plane of PPPCSR / ThreadPoolPPPCSR): several shards driven by one process, the
    batch binned on the device and stored into the owners' receive buffers.  All shards on GPU 0 here (the peer
    pointers are then plain device pointers); mixed stream with values; compared with the oracle."""
    n, m = 5000, 60000
    rng = np.random.default_rng(parts)
    src, dst = rng.integers(0, n, m), rng.integers(0, n, m)
    val = np.where(rng.integers(0, 4, m) != 0, rng.integers(1, 1000, m), 0)
    starts = np.array([0] + sorted(rng.choice(np.arange(1, n), parts - 1, replace=False).tolist()) + [n], dtype=np.uint64)
    g = pp.Group(n, starts, [0] * parts, region_cap=m // parts + 1, with_values=True)
    o = O.OraclePCSR(n)
    for lo in range(0, m, 20000):
        sl = slice(lo, lo + 20000)
        st = g.apply(src[sl], dst[sl], val[sl])
        assert sum(x["batch_size"] for x in st) == 20000
        o.apply(src[sl], dst[sl], val[sl])
    rowptr, col, nn = o.export()
    for r, sh in enumerate(g.shards):
        lo, hi = int(starts[r]), int(starts[r + 1])
        rp, c = sh.export()
        assert np.array_equal(rp, rowptr[lo:hi + 1] - rowptr[lo]), r
        assert np.array_equal(c, col[int(rowptr[lo]):int(rowptr[hi])]), r
        assert np.array_equal(sh.num_neighbors(), nn[lo:hi]), r
        assert not sh.check(True).violations(True)
        assert g.owner(lo) == r and g.owner(hi - 1) == r
    g.close()


def test_cli_pppcsr_device_route_matches_ppcsr(tmp_path):
    """-pppcsrnuma (partitions fed through the device-side group route, reference and -balanced boundaries) builds the
    same logical graph as -ppcsr: same checksum line."""
    build.build_host()
    scale = 12
    cs, cd = synth.rmat(scale, 0, 16 << scale, 42)
    idx = synth.sample_without_replacement(16 << scale, 20000, 7)
    core, upd = str(tmp_path / "core.bin"), str(tmp_path / "upd.bin")
    np.stack([cs, cd], axis=1).astype("<u4").tofile(core)
    np.stack([cs[idx], cd[idx]], axis=1).astype("<u4").tofile(upd)
    sums = []
    for mode in (["-ppcsr"], ["-pppcsrnuma", "-partitions_per_domain=3"], ["-pppcsr", "-partitions_per_domain=4", "-balanced"]):
        cmd = [build.CLI, "-threads=8", "-delete", "-size=20000", *mode, f"-core_graph={core}", f"-update_file={upd}",
               "-check", "-checksum"]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        assert "PMA invariants: ok" in r.stdout
        sums.append([l for l in r.stdout.splitlines() if l.startswith("Graph checksum: ")][0])
    assert sums[0] == sums[1] == sums[2]


def test_benchmark_harnesses_run_on_the_gpu(tmp_path):
    """the partitions sweep (reference benchmark-partitioning.sh layout) driven through the C++ CLI,
    and one cell of the GPU-count strong-scaling harness (bench.py launched as the driver launches it), at toy sizes."""
    import importlib.util
    import sys

    build.build_host()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("partitioning", os.path.join(root, "benchmarks", "partitioning.py"))
    pt = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(pt)
    prefix = str(tmp_path / "part")
    assert pt.main(["--scale", "12", "--size", "20000", "--partitions", "1", "3", "--reps", "2", "--out-prefix", prefix,
                    "--workdir", str(tmp_path)]) == 0
    rows = open(prefix + "_all_results.csv").read().strip().splitlines()
    assert rows[0] == pt.header(2) and [r.split()[0] for r in rows[1:]] == ["1", "3"]
    dev = open(prefix + "_device_ms.csv").read().strip().splitlines()
    assert all(float(x) > 0 for x in dev[1].split()[1:3])  # the device time of the insert batches was scraped
    assert open(prefix + "_plot_data.dat").read().splitlines()[0] == "partitions ins del ins-NUMA del-NUMA"
    r = subprocess.run([sys.executable, os.path.join(root, "benchmarks", "strong_scaling.py"), "--gpus", "1", "--reps", "1",
                        "--scale", "14", "--batch", "100000", "--steps", "2", "--warmup", "1", "--out-prefix",
                        str(tmp_path / "ss")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    row = open(str(tmp_path / "ss.csv")).read().strip().splitlines()[1].split()
    assert row[0] == "1" and float(row[1]) > 0 and float(row[4]) > 0
