"""Host-side logic of the multi-GPU path on CPU: shard tables, owner lookup, all-to-all split sizes and the
exchange itself over the gloo backend with world_size 2 and 3 (no GPU).  The shards are stood in for by the
oracle (test double) and the binning by the pure-torch TorchBinner; the product path uses the CUDA kernels."""
import importlib
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_py as O

router = importlib.import_module("parallel-packed-csr_b200.router")
synth = importlib.import_module("parallel-packed-csr_b200.synth")


def test_equal_vertex_split_matches_reference_table():
    # reference src/pppcsr/PPPCSR.cpp:13-34 via the oracle's restatement
    for n, parts in ((10, 4), (65537, 8), (1000, 3), (8, 8)):
        starts, sizes = O.partition_table(n, parts)
        mine = router.equal_vertex_starts(n, parts)
        assert mine[:-1].tolist() == starts.tolist() and int(mine[-1]) == n
        assert np.diff(mine.astype(np.int64)).tolist() == sizes.tolist()
        for v in (0, 1, n // 2, n - 1, n + 5):
            assert router.owner_of(mine, v) == O.partition_owner(starts, v)


def test_split_counts_and_scheduler_table():
    # the multi-GPU analogue of reference test/SchedulerTest.cpp: for shards in [1,8] every update lands on exactly
    # one owner, counts add up, and the owner is the shard whose range contains src
    rng = np.random.default_rng(0)
    for parts in range(1, 9):
        n = 1000 + parts
        starts = router.equal_vertex_starts(n, parts)
        src = rng.integers(0, n, 5000)
        counts = router.split_counts_by_owner(src, starts)
        assert counts.sum() == src.size and len(counts) == parts
        owners = np.array([router.owner_of(starts, int(v)) for v in src[:200]])
        lo, hi = starts[owners], starts[owners + 1]
        assert np.all((src[:200] >= lo) & (src[:200] < hi))


def test_edge_balanced_starts_balance_rmat():
    scale, parts = 12, 8
    s, _ = synth.rmat(scale, 0, 16 << scale, 42)
    starts = router.edge_balanced_starts(torch.from_numpy(s), 1 << scale, parts)
    assert starts[0] == 0 and starts[-1] == 1 << scale and np.all(np.diff(starts.astype(np.int64)) > 0)
    per = router.split_counts_by_owner(s, starts)
    eq = router.split_counts_by_owner(s, router.equal_vertex_starts(1 << scale, parts))
    assert per.max() < 1.35 * per.mean()          # edge-balanced ranges
    assert eq.max() > 2.5 * eq.mean()             # the reference's equal-vertex split is badly skewed on R-MAT


class OracleShard:
    """Test double with the Shard surface the router needs."""

    def __init__(self, n_local):
        self.g = O.OraclePCSR(n_local)

    def apply_device(self, d_src, d_dst, d_val, count, default_val):
        raise NotImplementedError


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, scale, n_upd, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 1 << scale
        starts = router.equal_vertex_starts(n, world)
        shard = OracleShard(int(starts[rank + 1] - starts[rank]))
        g = router.ShardedGraph(n, starts, rank, world, 0, dist=dist, shard_factory=lambda k: shard,
                                binner=router.TorchBinner())
        lo, hi = rank * n_upd // world, (rank + 1) * n_upd // world
        us, ud = synth.uniform(scale, lo, hi, 7)
        vals = (np.arange(lo, hi) % 5).astype(np.int32)  # zeros = deletes, travel through the third all-to-all
        r_src, r_dst, r_val = g.route(torch.from_numpy(us.astype(np.int32)), torch.from_numpy(ud.astype(np.int32)),
                                      torch.from_numpy(vals))
        shard.g.apply(r_src.numpy().astype(np.uint32), r_dst.numpy().astype(np.uint32), r_val.numpy().astype(np.uint32))
        rowptr, col, nn = shard.g.export()
        np.savez(out.format(rank=rank), rowptr=rowptr, col=col, nn=nn, send=np.array(g.last_route["send"]),
                 recv=np.array(g.last_route["recv"]), local_max=int(r_src.max()) if r_src.numel() else 0)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_all_to_all_routing_gloo(world, tmp_path):
    scale, n_upd = 9, 6000
    out = str(tmp_path / "rank{rank}.npz")
    mp.spawn(_worker, args=(world, _free_port(), scale, n_upd, out), nprocs=world, join=True)
    n = 1 << scale
    starts = router.equal_vertex_starts(n, world)
    # single-shard truth
    us, ud = synth.uniform(scale, 0, n_upd, 7)
    vals = np.arange(n_upd) % 5
    whole = O.OraclePCSR(n)
    whole.apply(us, ud, vals)
    rowptr, col, nn = whole.export()
    sends = []
    for r in range(world):
        z = np.load(out.format(rank=r))
        lo, hi = int(starts[r]), int(starts[r + 1])
        assert int(z["local_max"]) < hi - lo                      # sources were made shard-local
        assert np.array_equal(z["rowptr"], rowptr[lo:hi + 1] - rowptr[lo])
        assert np.array_equal(z["col"], col[int(rowptr[lo]):int(rowptr[hi])])   # dests stay global
        assert np.array_equal(z["nn"], nn[lo:hi])
        sends.append(z["send"])
        # what rank r sent must equal the host statement of the split
        mine = us[r * n_upd // world:(r + 1) * n_upd // world]
        assert np.array_equal(z["send"], router.split_counts_by_owner(mine, starts))
    sends = np.stack(sends)
    for r in range(world):
        assert np.array_equal(np.load(out.format(rank=r))["recv"], sends[:, r])
