"""Vertex-range shards, one per GPU, with NCCL all-to-all routing of update batches.

The reference's PPPCSR (src/pppcsr/PPPCSR.cpp:13-66) splits [0,n) into contiguous vertex ranges, one PCSR
per NUMA domain, and ThreadPoolPPPCSR::submit_* (src/thread_pool_pppcsr/thread_pool_pppcsr.cpp:96-118)
hands every op to a thread of the domain that owns `src`.  Here a range is a shard on one GPU (one process
per GPU, torch.distributed over NCCL for the plumbing) and that hand-over is: bin the batch by owner on the
device (C-ABI ppcsr_bin_by_owner), exchange the counts, ONE all-to-all of packed (src,dst) records, then
the owning shard applies what it received.  Shards never exchange edges afterwards (the reference's
migration hooks are empty stubs, src/pcsr/PCSR.cpp:1447-1468).

`binner` / `shard_factory` are injectable so the host-side logic (shard table, split sizes, exchange) is
testable on CPU with the gloo backend; the defaults are the CUDA kernels and fail loudly without a GPU.
"""
from __future__ import annotations

import ctypes as C

import numpy as np


def equal_vertex_starts(n: int, parts: int) -> np.ndarray:
    """The reference's split: parts equal vertex counts, remainder to the last (PPPCSR.cpp:20,27-29)."""
    size = n // parts
    starts = [p * size for p in range(parts)] + [n]
    return np.asarray(starts, dtype=np.uint64)


def owner_of(starts: np.ndarray, v: int) -> int:
    """reference PPPCSR::get_partiton (PPPCSR.cpp:58-66): first boundary > v, else the last partition."""
    parts = len(starts) - 1
    for i in range(1, parts):
        if starts[i] > v:
            return i - 1
    return parts - 1


def edge_balanced_starts(core_src, n: int, parts: int, dist=None, vertex_weight: float | None = 0.0) -> np.ndarray:
    """Contiguous vertex ranges with ~equal weight sum(deg(v) + vertex_weight) (R-MAT puts ~44 % of the sources in
    the first eighth of the id space).  vertex_weight = 0 balances the stored edges only;
    vertex_weight = None uses the average degree, which balances storage against a uniform update stream
    (a shard's share of uniform updates is proportional to its vertex count).
    `core_src`: this rank's slice of the core sources (torch)."""
    import torch

    hist = torch.bincount(core_src.long(), minlength=n)
    if dist is not None:
        dist.all_reduce(hist)
    if vertex_weight is None:
        vertex_weight = float(hist.sum().item()) / n
    if vertex_weight:
        hist = hist * 16 + int(round(vertex_weight * 16))
    csum = torch.cumsum(hist, 0)
    total = int(csum[-1].item())
    targets = torch.tensor([total * p // parts for p in range(1, parts)], device=csum.device, dtype=csum.dtype)
    cuts = torch.searchsorted(csum, targets, right=False).cpu().numpy().astype(np.int64) + 1
    starts = np.zeros(parts + 1, dtype=np.int64)
    starts[parts] = n
    for p in range(1, parts):
        starts[p] = min(max(int(cuts[p - 1]), starts[p - 1] + 1), n - (parts - p))
    return starts.astype(np.uint64)


def split_counts_by_owner(src: np.ndarray, starts: np.ndarray) -> np.ndarray:
    """Host statement of the all-to-all send counts (used by tests): updates per owning shard."""
    owners = np.searchsorted(np.asarray(starts[1:-1], dtype=np.uint64), np.asarray(src, dtype=np.uint64), side="right")
    return np.bincount(owners, minlength=len(starts) - 1).astype(np.int64)


class CudaBinner:
    """Device binning through the C-ABI (k_bin_count / k_bin_scatter): stable, src made shard-local."""

    def __init__(self, device_index: int):
        from . import load_library

        self.L = load_library()
        self.device_index = device_index

    def __call__(self, starts_dev, parts, src, dst, val):
        import torch

        count = src.numel()
        out_src = torch.empty_like(src)
        out_dst = torch.empty_like(dst)
        out_val = torch.empty_like(src) if val is not None else None
        counts = (C.c_uint64 * parts)()
        stream = torch.cuda.current_stream().cuda_stream or 1  # 0x1 = cudaStreamLegacy: torch's default stream
        rc = self.L.ppcsr_bin_by_owner(self.device_index, stream, starts_dev.data_ptr(), parts, src.data_ptr(),
                                       dst.data_ptr(), val.data_ptr() if val is not None else None, count,
                                       out_src.data_ptr(), out_dst.data_ptr(),
                                       out_val.data_ptr() if out_val is not None else None, counts)
        if rc != 0:
            raise RuntimeError(f"ppcsr_bin_by_owner failed: {self.L.ppcsr_last_error().decode()}")
        return out_src, out_dst, out_val, [int(c) for c in counts]

    def packed(self, starts_dev, parts, src, dst, val):
        """Same binning, emitting the packed (local_src << 32 | dst) records of the all-to-all."""
        import torch

        count = src.numel()
        out = torch.empty(count, dtype=torch.int64, device=src.device)
        out_val = torch.empty_like(src) if val is not None else None
        counts = (C.c_uint64 * parts)()
        stream = torch.cuda.current_stream().cuda_stream or 1
        rc = self.L.ppcsr_bin_by_owner_packed(self.device_index, stream, starts_dev.data_ptr(), parts, src.data_ptr(),
                                              dst.data_ptr(), val.data_ptr() if val is not None else None, count,
                                              out.data_ptr(), out_val.data_ptr() if out_val is not None else None, counts)
        if rc != 0:
            raise RuntimeError(f"ppcsr_bin_by_owner_packed failed: {self.L.ppcsr_last_error().decode()}")
        return out, out_val, [int(c) for c in counts]


class PeerExchange:
    """Routing fused with the exchange over NVLink / NVSwitch peer memory.

    Every rank owns a symmetric receive buffer (torch.distributed symmetric memory: the plumbing that maps each
    rank's allocation into every other rank's address space) of 2 x world regions of `cap` packed records; the
    binning kernel of rank s stores the records owned by rank d straight into region s of d's buffer (C-ABI
    ppcsr_bin_to_peers) together with the count, one device-side barrier publishes them, and the owner applies the
    regions in place (ppcsr_apply_batch_segments_device).  Compared with bin -> count exchange -> NCCL all-to-all
    there is no send buffer, no second copy and ONE host synchronisation (the world counts) instead of three.
    The two halves of the buffer alternate between batches: a peer can run at most one batch ahead (it cannot pass
    the next barrier alone), so it never overwrites records that are still being read.
    """

    def __init__(self, dist, device, world: int, rank: int, cap: int, with_values: bool = False):
        import torch
        import torch.distributed._symmetric_memory as symm_mem

        from . import load_library

        self.L = load_library()
        self.torch, self.dist = torch, dist
        self.world, self.rank, self.cap = world, rank, int(cap)
        self.device = device
        group = dist.group.WORLD
        self.rec = symm_mem.empty(2 * world * self.cap, dtype=torch.int64, device=device)
        self.cnt = symm_mem.empty(2 * world, dtype=torch.int64, device=device)
        self.rec_h = symm_mem.rendezvous(self.rec, group)
        self.cnt_h = symm_mem.rendezvous(self.cnt, group)
        self.val = self.val_h = None
        if with_values:
            self.val = symm_mem.empty(2 * world * self.cap, dtype=torch.int32, device=device)
            self.val_h = symm_mem.rendezvous(self.val, group)
        self.cnt.zero_()
        self.parity = 0
        # per parity: host arrays of the peers' base pointers (passed by value into the kernel's parameter block)
        u64 = C.c_uint64 * world
        self.rec_ptrs = [u64(*[int(b) + par * world * self.cap * 8 for b in self.rec_h.buffer_ptrs]) for par in (0, 1)]
        self.cnt_ptrs = [u64(*[int(b) + par * world * 8 for b in self.cnt_h.buffer_ptrs]) for par in (0, 1)]
        self.val_ptrs = None
        if with_values:
            self.val_ptrs = [u64(*[int(b) + par * world * self.cap * 4 for b in self.val_h.buffer_ptrs]) for par in (0, 1)]
        torch.cuda.synchronize(device)
        self.rec_h.barrier(channel=0)

    def exchange(self, starts_dev, src, dst, val):
        """Returns device pointers: this rank's regions, their values (or None), the per-sender counts.  No host
        synchronisation: the owner's key builder reads the counts on the device."""
        torch = self.torch
        if val is not None and self.val_ptrs is None:
            raise RuntimeError("PeerExchange was created without value buffers")
        par = self.parity
        self.parity ^= 1
        stream = torch.cuda.current_stream(self.device)
        rc = self.L.ppcsr_bin_to_peers(self.device.index, stream.cuda_stream or 1, starts_dev.data_ptr(), self.world, self.rank,
                                       src.data_ptr(), dst.data_ptr(), val.data_ptr() if val is not None else None,
                                       src.numel(), self.rec_ptrs[par],
                                       self.val_ptrs[par] if val is not None else None, self.cnt_ptrs[par], self.cap)
        if rc != 0:
            raise RuntimeError(f"ppcsr_bin_to_peers failed: {self.L.ppcsr_last_error().decode()}")
        self.rec_h.barrier(channel=par)  # every peer's stores (records and counts) are visible after this
        rec_ptr = self.rec.data_ptr() + par * self.world * self.cap * 8
        val_ptr = self.val.data_ptr() + par * self.world * self.cap * 4 if val is not None else None
        cnt_ptr = self.cnt.data_ptr() + par * self.world * 8
        return rec_ptr, val_ptr, cnt_ptr


class TorchBinner:
    """Pure-torch statement of the same binning (stable sort by owner).  TEST DOUBLE for the CPU/gloo tests;
    the product path uses CudaBinner."""

    def __call__(self, starts_dev, parts, src, dst, val):
        import torch

        owners = torch.searchsorted(starts_dev[1:parts].contiguous(), src.long(), right=True)
        order = torch.argsort(owners, stable=True)
        counts = torch.bincount(owners, minlength=parts).tolist()
        local = (src.long() - starts_dev[owners]).to(src.dtype)
        return local[order], dst[order], (val[order] if val is not None else None), counts


class ShardedGraph:
    """PPPCSR recast: rank r owns sources [starts[r], starts[r+1]) in a Shard on its GPU."""

    def __init__(self, n, starts, rank, world, device_index, dist=None, shard_factory=None, binner=None,
                 peer_cap: int = 0, peer_values: bool = False):
        import torch

        self.n, self.rank, self.world, self.dist = int(n), rank, world, dist
        self.starts = np.asarray(starts, dtype=np.uint64)
        assert len(self.starts) == world + 1 and self.starts[0] == 0 and self.starts[-1] == n
        self.n_local = int(self.starts[rank + 1] - self.starts[rank])
        if shard_factory is None:
            from . import Shard

            shard_factory = lambda n_local: Shard(n_local, device=device_index)  # noqa: E731
        self.shard = shard_factory(self.n_local)
        self.torch = torch
        on_gpu = torch.cuda.is_available() and not isinstance(binner, TorchBinner)
        self.dev = torch.device("cuda", device_index) if on_gpu else torch.device("cpu")
        self.binner = binner if binner is not None else (CudaBinner(device_index) if world > 1 else None)
        self.starts_dev = torch.from_numpy(self.starts.astype(np.int64)).to(self.dev)
        self.last_route = None
        # peer_cap > 0: route batches of up to peer_cap updates per rank through NVLink peer memory (PeerExchange);
        # larger batches, or a platform without symmetric memory, take the NCCL all-to-all.  Every rank makes the same
        # choice: apply() decides from the largest slice of ANY rank (see _largest_slice).
        self.peer = None
        self.peer_max_total = 0  # optional bound on what one rank can receive per batch (sizes the key array)
        if peer_cap and world > 1 and on_gpu and isinstance(self.binner, CudaBinner):
            try:
                self.peer = PeerExchange(dist, self.dev, world, rank, peer_cap, with_values=peer_values)
            except Exception as e:  # noqa: BLE001 -- symmetric memory is an optional transport
                import sys

                print(f"[rank {rank}] peer-memory routing unavailable ({type(e).__name__}: {e}); using NCCL all-to-all",
                      file=sys.stderr)
                ok = torch.zeros(1, device=self.dev)
            else:
                ok = torch.ones(1, device=self.dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)  # all or nothing: the ranks must agree on the transport
            if ok.item() == 0:
                self.peer = None
        self.route_timing = None  # set to [] to collect per-stage routing times (development aid)
        # The routed records are produced on torch's current stream (binning kernels, NCCL all-to-all): the shard
        # must consume them in stream order, not on its private stream.
        if on_gpu and hasattr(self.shard, "bind_torch_stream"):
            self.shard.bind_torch_stream(torch.cuda.current_stream(self.dev))

    def get_partition(self, v: int) -> int:
        return owner_of(self.starts, v)

    # ---- routing ----
    def route(self, src, dst, val=None):
        """Returns this rank's share of the global batch: (local_src, dst, val) tensors, after one all-to-all."""
        torch, dist = self.torch, self.dist
        if self.world == 1:
            return src, dst, val
        b_src, b_dst, b_val, send = self.binner(self.starts_dev, self.world, src, dst, val)
        send_t = torch.tensor(send, dtype=torch.int64, device=src.device)
        recv_t = torch.empty_like(send_t)
        dist.all_to_all_single(recv_t, send_t)
        recv = recv_t.tolist()
        packed = (b_src.long() << 32) | (b_dst.long() & 0xFFFFFFFF)
        got = torch.empty(sum(recv), dtype=torch.int64, device=src.device)
        dist.all_to_all_single(got, packed, output_split_sizes=recv, input_split_sizes=send)
        r_src = (got >> 32).to(src.dtype)
        r_dst = (got & 0xFFFFFFFF).to(dst.dtype)
        r_val = None
        if val is not None:
            r_val = torch.empty(sum(recv), dtype=val.dtype, device=src.device)
            dist.all_to_all_single(r_val, b_val, output_split_sizes=recv, input_split_sizes=send)
        self.last_route = {"send": send, "recv": recv}
        return r_src, r_dst, r_val

    def route_packed(self, src, dst, val=None):
        """Product path: bin straight into packed records, one all-to-all, no unpacking (the shard consumes
        packed records).  Returns (packed int64 tensor, val tensor or None)."""
        torch, dist = self.torch, self.dist
        packed, b_val, send = self.binner.packed(self.starts_dev, self.world, src, dst, val)
        recv = self._exchange_counts(send, src.device)
        got = torch.empty(sum(recv), dtype=torch.int64, device=src.device)
        dist.all_to_all_single(got, packed, output_split_sizes=recv, input_split_sizes=send)
        r_val = None
        if val is not None:
            r_val = torch.empty(sum(recv), dtype=val.dtype, device=src.device)
            dist.all_to_all_single(r_val, b_val, output_split_sizes=recv, input_split_sizes=send)
        self.last_route = {"send": send, "recv": recv}
        return got, r_val

    def _exchange_counts(self, send, device):
        """Every rank learns how many records each peer sends it: a tiny all-to-all through pinned staging."""
        torch, dist = self.torch, self.dist
        if getattr(self, "_cnt_send", None) is None:
            pin = device.type == "cuda"
            self._cnt_host = torch.empty(self.world, dtype=torch.int64, pin_memory=pin)
            self._cnt_back = torch.empty(self.world, dtype=torch.int64, pin_memory=pin)
            self._cnt_send = torch.empty(self.world, dtype=torch.int64, device=device)
            self._cnt_recv = torch.empty(self.world, dtype=torch.int64, device=device)
        self._cnt_host.copy_(torch.tensor(send, dtype=torch.int64))
        self._cnt_send.copy_(self._cnt_host, non_blocking=True)
        dist.all_to_all_single(self._cnt_recv, self._cnt_send)
        self._cnt_back.copy_(self._cnt_recv, non_blocking=True)
        if device.type == "cuda":
            torch.cuda.current_stream().synchronize()
        return self._cnt_back.tolist()

    def _largest_slice(self, count: int, global_max):
        """The largest per-rank slice of this batch, known to EVERY rank: the transport (peer memory or NCCL) must be
        the same on all ranks, so it may only depend on a value they all share -- the caller's `global_max` (no
        communication) or, without it, one all-reduce(MAX) of the slice sizes."""
        if global_max is not None:
            if count > global_max:
                raise ValueError(f"this rank's slice ({count}) exceeds the declared global_max ({global_max})")
            return int(global_max)
        if self.world == 1 or self.dist is None:
            return count
        t = self.torch.tensor([count], dtype=self.torch.int64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return int(t.item())

    def apply(self, src, dst, val=None, default_val=1, global_max=None):
        """Device tensors (int32 bit patterns of u32 ids) holding this rank's slice of the global batch.
        `global_max` (optional, identical on all ranks): an upper bound of the largest slice any rank holds; see
        _largest_slice."""
        use_peer = False
        if self.world > 1 and self.peer is not None and (val is None or self.peer.val_ptrs is not None):
            use_peer = self._largest_slice(src.numel(), global_max) <= self.peer.cap
        if use_peer:
            if self.route_timing is not None:
                import time

                self.torch.cuda.synchronize()
                t0 = time.perf_counter()
            rec_ptr, val_ptr, cnt_ptr = self.peer.exchange(self.starts_dev, src, dst, val)
            if self.route_timing is not None:
                self.torch.cuda.synchronize()
                t1 = time.perf_counter()
            st = self.shard.apply_segments_device(rec_ptr, val_ptr, self.peer.cap, cnt_ptr, self.world,
                                                  self.peer_max_total, default_val)
            if self.route_timing is not None:
                self.torch.cuda.synchronize()
                self.route_timing.append([round((t1 - t0) * 1e3, 3), round((time.perf_counter() - t1) * 1e3, 3)])
            self.last_route = {"transport": "peer", "received": st["batch_size"]}
            return st
        if self.world > 1 and hasattr(self.binner, "packed"):
            if self.route_timing is not None:
                return self._apply_timed(src, dst, val, default_val)
            got, r_val = self.route_packed(src, dst, val)
            return self.shard.apply_packed_device(got.data_ptr(), r_val.data_ptr() if r_val is not None else None,
                                                  got.numel(), default_val)
        r_src, r_dst, r_val = self.route(src, dst, val)
        r_src, r_dst = r_src.contiguous(), r_dst.contiguous()
        return self.shard.apply_device(r_src.data_ptr(), r_dst.data_ptr(),
                                       r_val.contiguous().data_ptr() if r_val is not None else None,
                                       r_src.numel(), default_val)

    def _apply_timed(self, src, dst, val, default_val):
        """Development aid (route_timing = []): wall-clock of each routing stage with a device sync after it."""
        import time

        torch, dist = self.torch, self.dist
        t = [time.perf_counter()]

        def mark():
            torch.cuda.synchronize()
            t.append(time.perf_counter())

        packed, b_val, send = self.binner.packed(self.starts_dev, self.world, src, dst, val)
        mark()
        recv = self._exchange_counts(send, src.device)
        mark()
        got = torch.empty(sum(recv), dtype=torch.int64, device=src.device)
        dist.all_to_all_single(got, packed, output_split_sizes=recv, input_split_sizes=send)
        mark()
        st = self.shard.apply_packed_device(got.data_ptr(), None, got.numel(), default_val)
        mark()
        self.route_timing.append([round((b - a) * 1e3, 3) for a, b in zip(t, t[1:])])
        return st

    def apply_host(self, src, dst, val=None, default_val=1, global_max=None):
        """Host (pinned) numpy arrays: the public end-to-end call. H2D happens inside."""
        if self.world == 1:
            return self.shard.apply(src, dst, val, default_val)
        torch = self.torch
        d_src = torch.from_numpy(src).to(self.dev, non_blocking=True)
        d_dst = torch.from_numpy(dst).to(self.dev, non_blocking=True)
        d_val = torch.from_numpy(val).to(self.dev, non_blocking=True) if val is not None else None
        return self.apply(d_src, d_dst, d_val, default_val, global_max=global_max)

    # ---- pipelined end-to-end submit: the H2D copy of batch i+1 runs under the compute of batch i ----
    def submit_host(self, src, dst, val=None, default_val=1, global_max=None):
        """Starts the host->device copy of this rank's slice of a batch (pinned numpy arrays) on a copy stream and
        returns a ticket for wait_host().  One shard: the C-ABI's ppcsr_submit_batch.  Sharded: two device staging
        slots filled on a side stream; the routing and the apply run in wait_host()."""
        if self.world == 1:
            return self.shard.submit(src, dst, val, default_val)
        torch = self.torch
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(self.dev)
            self._slots = [None, None]
            self._next_ticket = 1
        t = self._next_ticket
        self._next_ticket += 1
        slot = self._slots[t & 1]
        if slot is not None and slot.get("busy"):
            raise RuntimeError("two batches are already in flight; wait_host() for the older one first")
        n = src.shape[0]
        if slot is None or slot["src"].numel() < n or (val is not None and slot["val"] is None):
            slot = {"src": torch.empty(n, dtype=torch.int32, device=self.dev),
                    "dst": torch.empty(n, dtype=torch.int32, device=self.dev),
                    "val": torch.empty(n, dtype=torch.int32, device=self.dev) if val is not None else None}
            self._slots[t & 1] = slot
        with torch.cuda.stream(self._copy_stream):
            slot["src"][:n].copy_(torch.from_numpy(src.view(np.int32)), non_blocking=True)
            slot["dst"][:n].copy_(torch.from_numpy(dst.view(np.int32)), non_blocking=True)
            if val is not None:
                slot["val"][:n].copy_(torch.from_numpy(val.view(np.int32)), non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
        slot.update(busy=True, n=n, ev=ev, has_val=val is not None, default_val=default_val, host=(src, dst, val),
                    global_max=global_max)
        return t

    def wait_host(self, ticket):
        if self.world == 1:
            return self.shard.wait(ticket)
        slot = self._slots[ticket & 1]
        assert slot is not None and slot.get("busy"), "no such batch in flight"
        self.torch.cuda.current_stream(self.dev).wait_event(slot["ev"])
        n = slot["n"]
        st = self.apply(slot["src"][:n], slot["dst"][:n], slot["val"][:n] if slot["has_val"] else None,
                        slot["default_val"], global_max=slot.get("global_max"))
        slot["busy"] = False
        slot["host"] = None
        return st

    # ---- PageRank across shards: every shard pushes into a full-length vector, then one all-reduce ----
    def pagerank_step(self, values_global):
        """values_global: float64 device tensor of length n (replicated). Returns the summed push result."""
        torch = self.torch
        lo, hi = int(self.starts[self.rank]), int(self.starts[self.rank + 1])
        local_in = values_global[lo:hi].contiguous()
        out = torch.zeros(self.n, dtype=torch.float64, device=values_global.device)
        from . import _check

        _check(self.shard.L.ppcsr_pagerank_push_device(self.shard.h, local_in.data_ptr(), out.data_ptr(), self.n))
        self.shard.sync()
        if self.world > 1:
            self.dist.all_reduce(out)
        return out
