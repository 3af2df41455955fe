"""Deterministic synthetic edge streams (SURVEY.md §8d "Synthetic inputs").

Everything is a pure function of (seed, element index): a 32-bit counter hash evaluated in
int64 lanes, so numpy on the host and torch on the GPU produce bit-identical streams and any
slice [lo, hi) of a stream can be generated independently (each rank of a multi-GPU run
generates its own slice of the same global stream).

  rmat(scale, count, seed)      Graph500-style R-MAT, a,b,c,d = .57,.19,.19,.05, duplicates and
                                self loops kept, no id permutation.
  uniform(scale, count, seed)   src, dst iid uniform in [0, 2^scale).
  sample_without_replacement    indices for the delete stream (sampled from the raw core list).

The reference has no generator; its benchmark scripts take edge-list files
(reference src/benchmarking/benchmark-strong-scaling.sh:6-28, src/main.cpp:29-62).
"""
from __future__ import annotations

import numpy as np

_M32 = 0xFFFFFFFF
_A, _B, _C = 0.57, 0.19, 0.19
_T_A = int(_A * 2**32)
_T_AB = int((_A + _B) * 2**32)
_T_ABC = int((_A + _B + _C) * 2**32)


def _mix32(x):
    """murmur3 fmix32 on int64 lanes holding values in [0, 2^32). Works for numpy and torch."""
    x = x ^ (x >> 16)
    x = (x * 0x85EBCA6B) & _M32
    x = x ^ (x >> 13)
    x = (x * 0xC2B2AE35) & _M32
    x = x ^ (x >> 16)
    return x


def _arange(lo, hi, device):
    if device is None:
        return np.arange(lo, hi, dtype=np.int64)
    import torch

    return torch.arange(lo, hi, dtype=torch.int64, device=device)


def _base_hash(idx, seed):
    hi = _mix32(((idx >> 32) + (seed * 0x9E3779B1 & _M32) + 0x7F4A7C15) & _M32)
    return _mix32((idx & _M32) ^ hi)


def rmat(scale: int, lo: int, hi: int, seed: int, device=None):
    """Edges [lo, hi) of the R-MAT stream `seed`. Returns (src, dst) int64 arrays/tensors."""
    return rmat_at(scale, _arange(lo, hi, device), seed)


def rmat_at(scale: int, idx, seed: int):
    """Edges of the R-MAT stream `seed` at the given int64 element indices (numpy array or torch tensor)."""
    device = None if isinstance(idx, np.ndarray) else idx.device
    with np.errstate(over="ignore"):
        h0 = _base_hash(idx, seed)
        src = idx * 0
        dst = idx * 0
        for level in range(scale):
            r = _mix32((h0 + ((level + 1) * 0x9E3779B9 & _M32)) & _M32)
            sbit = r >= _T_AB
            dbit = ((r >= _T_A) & (r < _T_AB)) | (r >= _T_ABC)
            if device is None:
                src = (src << 1) | sbit.astype(np.int64)
                dst = (dst << 1) | dbit.astype(np.int64)
            else:
                src = (src << 1) | sbit.long()
                dst = (dst << 1) | dbit.long()
        return src, dst


def uniform(scale: int, lo: int, hi: int, seed: int, device=None):
    """Edges [lo, hi) of the uniform stream `seed`."""
    with np.errstate(over="ignore"):
        idx = _arange(lo, hi, device)
        h0 = _base_hash(idx, seed)
        mask = (1 << scale) - 1
        s = _mix32((h0 + 0x68E31DA4) & _M32)
        d = _mix32((h0 + 0xB5297A4D) & _M32)
        if scale > 32:
            raise ValueError("scale > 32 unsupported")
        return s & mask, d & mask


def sample_without_replacement(total: int, count: int, seed: int, device=None):
    """`count` distinct indices in [0, total), in a seed-determined order (unique sort keys)."""
    with np.errstate(over="ignore"):
        idx = _arange(0, total, device)
        key = ((_base_hash(idx, seed) >> 1) << 32) | idx
        if device is None:
            order = np.argsort(key, kind="stable")
        else:
            import torch

            order = torch.argsort(key)
        return order[:count]


def mixed_ops(lo: int, hi: int, seed: int, device=None):
    """1 = add with probability 3/4, 0 = delete (mirrors reference test/DataStructureTest.cpp:128)."""
    with np.errstate(over="ignore"):
        idx = _arange(lo, hi, device)
        r = _mix32((_base_hash(idx, seed) + 0x1B873593) & _M32)
        add = (r & 3) != 0
        return add.astype(np.uint32) if device is None else add.to(dtype=__import__("torch").int32)


def write_triples(path: str, src, dst, val) -> None:
    """Raw little-endian u32 (src, dst, value) triples; value 0 = delete (oracle/ref_driver.cpp)."""
    src = np.asarray(src, dtype=np.uint32)
    dst = np.asarray(dst, dtype=np.uint32)
    val = np.broadcast_to(np.asarray(val, dtype=np.uint32), src.shape)
    out = np.empty((src.shape[0], 3), dtype="<u4")
    out[:, 0] = src
    out[:, 1] = dst
    out[:, 2] = val
    out.tofile(path)


def write_text(path: str, src, dst, op=None) -> None:
    """The reference's text format: `src dst[ op]` per line (reference src/main.cpp:40-60)."""
    src = np.asarray(src)
    dst = np.asarray(dst)
    with open(path, "w") as f:
        if op is None:
            for s, d in zip(src.tolist(), dst.tolist()):
                f.write(f"{s} {d}\n")
        else:
            for s, d, o in zip(src.tolist(), dst.tolist(), np.asarray(op).tolist()):
                f.write(f"{s} {d} {o}\n")


_M64 = (1 << 64) - 1


def mix64(x: np.ndarray) -> np.ndarray:
    """splitmix64 finaliser on uint64 arrays (the hash of the graph checksum)."""
    x = np.asarray(x, dtype=np.uint64).copy()
    with np.errstate(over="ignore"):
        x ^= x >> np.uint64(30)
        x *= np.uint64(0xBF58476D1CE4E5B9)
        x ^= x >> np.uint64(27)
        x *= np.uint64(0x94D049BB133111EB)
        x ^= x >> np.uint64(31)
    return x


def graph_checksum(rowptr, col, num_neighbors=None, vertex_offset: int = 0) -> dict:
    """Host statement of ppcsr_checksum / ref_driver --checksum over a CSR: edges, edge_hash, nn_hash."""
    rowptr = np.asarray(rowptr, dtype=np.int64)
    n = rowptr.shape[0] - 1
    src = np.repeat(np.arange(n, dtype=np.uint64) + np.uint64(vertex_offset), np.diff(rowptr))
    with np.errstate(over="ignore"):
        eh = int(mix64((src << np.uint64(32)) | np.asarray(col, dtype=np.uint64)).sum(dtype=np.uint64))
        nh = 0
        if num_neighbors is not None:
            v = np.arange(n, dtype=np.uint64) + np.uint64(vertex_offset)
            nh = int((np.asarray(num_neighbors, dtype=np.uint64) * mix64(v)).sum(dtype=np.uint64))
    return {"edges": int(rowptr[-1]), "edge_hash": eh & _M64, "nn_hash": nh & _M64}
