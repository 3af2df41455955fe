// batch.cuh -- front half of the batch pipeline:
//   build 64-bit keys (src<<32 | dst) with the reference's guards  -> radix sort -> last-op-wins
//   -> num_neighbors call counts -> segmented search of every unique update inside its vertex range
//   -> per-leaf insert/delete counts and the compacted, key-ordered insert list.
// Replaces, per update, reference PCSR::add_edge_parallel / remove_edge up to and including
// PCSR::binary_search (src/pcsr/PCSR.cpp:1374-1424, 709-747, 427-502).
#pragma once
#include "common.cuh"
#include "primitives.cuh"

namespace batch {

enum : uint8_t { CLS_INSERT = 0, CLS_OVERWRITE = 1, CLS_DELETE = 2, CLS_MISS = 3 };

constexpr int BT = 256;

// ---- keys -------------------------------------------------------------------------------------
// Guards: add with value != 0 needs src < n (reference PCSR.cpp:1375) and dst != SENT; remove needs
// src < n (the reference would read out of bounds, PCSR.cpp:717).  Rejected updates get the key
// (n << 32), which sorts after every valid key, and are counted.
__global__ void __launch_bounds__(BT) k_build_keys(const uint32_t *__restrict__ src, const uint32_t *__restrict__ dst,
                                                   const uint32_t *__restrict__ val, uint32_t default_val,
                                                   size_t count, uint32_t n, uint64_t *__restrict__ keys,
                                                   uint32_t *__restrict__ pay, BatchScalars *sc) {
  __shared__ uint32_t s_or, s_bad;
  if (threadIdx.x == 0) {
    s_or = 0;
    s_bad = 0;
  }
  __syncthreads();
  uint32_t my_or = 0, my_bad = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
    const uint32_t s = src[i], d = dst[i];
    const uint32_t v = val ? val[i] : default_val;
    const bool ok = s < n && !(v != 0 && d == PPCSR_SENT);
    keys[i] = ok ? (((uint64_t)s << 32) | d) : ((uint64_t)n << 32);
    pay[i] = ok ? v : 0u;
    if (ok) my_or |= d;
    else my_bad++;
  }
  my_or = __reduce_or_sync(0xFFFFFFFFu, my_or);
  my_bad = __reduce_add_sync(0xFFFFFFFFu, my_bad);
  if (lane_id() == 0) {
    if (my_or) atomicOr(&s_or, my_or);
    if (my_bad) atomicAdd(&s_bad, my_bad);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (s_or) atomicOr(&sc->dst_or, s_or);
    if (s_bad) atomicAdd(&sc->n_ignored, (unsigned long long)s_bad);
  }
}

// ---- last op wins + call counts ----------------------------------------------------------------
struct InLastOfRun {
  const uint64_t *keys;
  size_t count;
  uint64_t invalid_key;
  __device__ uint32_t operator()(size_t i) const {
    const uint64_t k = keys[i];
    if (k >= invalid_key) return 0;
    return (i + 1 == count || keys[i + 1] != k) ? 1u : 0u;
  }
};
struct OutUnique {
  const uint64_t *keys;
  const uint32_t *pay;
  uint64_t invalid_key;
  uint64_t *ukey;
  uint32_t *uval;
  uint8_t *ufirst_del;  // 1 if the FIRST op of the key's run in this batch is a remove
  __device__ void operator()(size_t i, uint32_t ex, uint32_t own) const {
    const uint64_t k = keys[i];
    if (own) {
      ukey[ex] = k;
      uval[ex] = pay[i];
    }
    // `ex` = number of complete runs before i = index of i's run in the unique list
    if (k < invalid_key && (i == 0 || keys[i - 1] != k)) ufirst_del[ex] = pay[i] == 0 ? 1 : 0;
  }
};

// num_neighbors += (#add calls) - (#remove calls) per source over the WHOLE sorted batch, duplicates
// included (reference PCSR.cpp:1392 and :747).  Sorted by src => warp-aggregated atomics.
// Also counts the removes that the sequential reference would report as `not found` because the previous
// op on the same key in this batch was already a remove (reference PCSR.cpp:750-754).
__global__ void __launch_bounds__(BT) k_count_calls(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ pay,
                                                    size_t count, uint64_t invalid_key, uint32_t *__restrict__ nn,
                                                    BatchScalars *sc) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t s = 0xFFFFFFFFu;
  int delta = 0;
  bool dup_miss = false;
  if (i < count) {
    const uint64_t k = keys[i];
    if (k < invalid_key) {
      s = (uint32_t)(k >> 32);
      delta = pay[i] != 0 ? 1 : -1;
      dup_miss = delta < 0 && i > 0 && keys[i - 1] == k && pay[i - 1] == 0;
    }
  }
  const unsigned dm = __ballot_sync(0xFFFFFFFFu, dup_miss);
  if (dm && lane_id() == 0) atomicAdd(&sc->n_not_found, (unsigned long long)__popc(dm));
  const unsigned peers = __match_any_sync(0xFFFFFFFFu, s);
  const unsigned adds = __ballot_sync(0xFFFFFFFFu, delta > 0);
  const unsigned dels = __ballot_sync(0xFFFFFFFFu, delta < 0);
  if (s != 0xFFFFFFFFu && (peers & lanemask_lt()) == 0) {
    const int sum = __popc(peers & adds) - __popc(peers & dels);
    if (sum) atomicAdd(&nn[s], (uint32_t)sum);
  }
}

// ---- segmented search ----------------------------------------------------------------------------
// Finds (src,dst) inside vertex src's slot range (beg[src], beg[src+1]).  Two levels: a binary search
// over the LEAVES of the range on their first item (empty leaves are skipped to the right), then a
// lower_bound inside the chosen leaf's live prefix.  Returns true and the slot of the edge if it exists;
// otherwise false and the slot of its PREDECESSOR (the largest item with a smaller key -- at worst the
// vertex's own sentinel).  A new edge joins the leaf of its predecessor.
// Same contract as reference PCSR::binary_search (PCSR.cpp:427-502): position of the smallest element
// >= key, with the gap probing replaced by per-leaf live counts.
__device__ __forceinline__ bool find_edge(const uint32_t *__restrict__ dest, const uint32_t *__restrict__ leaf_cnt,
                                          uint32_t b, uint32_t e, uint32_t ls, uint32_t d, uint32_t *slot) {
  const uint32_t Lb = b >> ls, Le = (e - 1) >> ls;
  uint32_t lo = Lb, hi = Le + 1;
  while (hi - lo > 1) {
    const uint32_t mid = lo + ((hi - lo) >> 1);
    uint32_t m2 = mid;
    while (m2 < hi && leaf_cnt[m2] == 0) m2++;
    if (m2 == hi) {
      hi = mid;
      continue;
    }
    if (dest[(size_t)m2 << ls] <= d) lo = m2;
    else hi = mid;
  }
  const uint32_t base = lo << ls;
  const uint32_t cnt = leaf_cnt[lo];
  uint32_t f_lo = (lo == Lb) ? (b - base) + 1 : 0;
  uint32_t f_hi = cnt;
  if (lo == (e >> ls) && (e - base) < f_hi) f_hi = e - base;
  uint32_t x = f_lo, y = f_hi;  // lower_bound of d in dest[base + [f_lo, f_hi))
  while (x < y) {
    const uint32_t mid = (x + y) >> 1;
    if (dest[base + mid] < d) x = mid + 1;
    else y = mid;
  }
  if (x < f_hi && dest[base + x] == d) {
    *slot = base + x;
    return true;
  }
  *slot = base + x - 1;
  return false;
}

__global__ void __launch_bounds__(BT) k_locate(const uint64_t *__restrict__ ukey, const uint32_t *__restrict__ uval,
                                               const uint8_t *__restrict__ ufirst_del,
                                               const unsigned long long *__restrict__ n_unique,
                                               const uint32_t *__restrict__ dest, uint32_t *__restrict__ val,
                                               const uint32_t *__restrict__ leaf_cnt, const uint32_t *__restrict__ beg,
                                               uint32_t ls, uint32_t *__restrict__ uloc, uint8_t *__restrict__ ucls,
                                               uint32_t *__restrict__ ins_cnt, uint32_t *__restrict__ del_cnt,
                                               BatchScalars *sc) {
  __shared__ uint32_t s_stat[4];
  if (threadIdx.x < 4) s_stat[threadIdx.x] = 0;
  __syncthreads();
  const size_t U = (size_t)*n_unique;
  const size_t u = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t cls = 0xFFu, leaf = 0xFFFFFFFFu;
  bool first_miss = false;  // the key's first op is a remove and the key is absent: a sequential `not found`
  if (u < U) {
    const uint64_t k = ukey[u];
    const uint32_t s = (uint32_t)(k >> 32), d = (uint32_t)k, v = uval[u];
    uint32_t slot;
    const bool hit = find_edge(dest, leaf_cnt, beg[s], beg[s + 1], ls, d, &slot);
    if (v != 0) {
      cls = hit ? CLS_OVERWRITE : CLS_INSERT;
      if (hit) val[slot] = v;  // duplicate insert overwrites the value (reference PCSR.cpp:529-532)
    } else {
      cls = hit ? CLS_DELETE : CLS_MISS;
      if (hit) val[slot] = 0;  // tombstone; compacted away by the rebalance of this leaf (PCSR.cpp:605-606)
    }
    first_miss = ufirst_del[u] != 0 && !hit;
    uloc[u] = slot;
    ucls[u] = (uint8_t)cls;
    leaf = slot >> ls;
  }
  // per-leaf counts: the batch is key-sorted, so equal leaves are adjacent -> one atomic per warp run
  const unsigned lt = lanemask_lt();
  {
    const uint32_t key = (cls == CLS_INSERT) ? leaf : 0xFFFFFFFFu;
    const unsigned peers = __match_any_sync(0xFFFFFFFFu, key);
    if (key != 0xFFFFFFFFu && (peers & lt) == 0) atomicAdd(&ins_cnt[leaf], (uint32_t)__popc(peers));
  }
  {
    const uint32_t key = (cls == CLS_DELETE) ? leaf : 0xFFFFFFFFu;
    const unsigned peers = __match_any_sync(0xFFFFFFFFu, key);
    if (key != 0xFFFFFFFFu && (peers & lt) == 0) atomicAdd(&del_cnt[leaf], (uint32_t)__popc(peers));
  }
#pragma unroll
  for (uint32_t c = 0; c < 4; c++) {
    // CLS_MISS only says "nothing to remove physically"; the reported not-found count follows the
    // sequential rule: first op of the key is a remove of an absent edge (+ the repeats counted earlier)
    const unsigned m = __ballot_sync(0xFFFFFFFFu, c == CLS_MISS ? first_miss : cls == c);
    if (lane_id() == 0 && m) atomicAdd(&s_stat[c], (uint32_t)__popc(m));
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (s_stat[CLS_INSERT]) atomicAdd(&sc->n_inserted, (unsigned long long)s_stat[CLS_INSERT]);
    if (s_stat[CLS_OVERWRITE]) atomicAdd(&sc->n_overwritten, (unsigned long long)s_stat[CLS_OVERWRITE]);
    if (s_stat[CLS_DELETE]) atomicAdd(&sc->n_deleted, (unsigned long long)s_stat[CLS_DELETE]);
    if (s_stat[CLS_MISS]) atomicAdd(&sc->n_not_found, (unsigned long long)s_stat[CLS_MISS]);
  }
}

// ---- compacted insert list (key order preserved) ---------------------------------------------------
struct InIsInsert {
  const uint8_t *ucls;
  __device__ uint32_t operator()(size_t i) const { return ucls[i] == CLS_INSERT ? 1u : 0u; }
};
struct OutInsert {
  const uint64_t *ukey;
  const uint32_t *uval;
  const uint32_t *uloc;
  uint32_t *ins_dst, *ins_val, *ins_pred;
  __device__ void operator()(size_t i, uint32_t ex, uint32_t own) const {
    if (own) {
      ins_dst[ex] = (uint32_t)ukey[i];
      ins_val[ex] = uval[i];
      ins_pred[ex] = uloc[i];
    }
  }
};

// ---- owner binning for the multi-GPU all-to-all (reference PPPCSR::get_partiton, PPPCSR.cpp:58-66) ----
__device__ __forceinline__ uint32_t owner_of(const uint64_t *starts, uint32_t parts, uint64_t v) {
  uint32_t lo = 0, hi = parts;  // last p with starts[p] <= v
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (starts[mid] <= v) lo = mid;
    else hi = mid;
  }
  return lo;
}

constexpr int BIN_MAX_PARTS = 64;

__global__ void __launch_bounds__(BT) k_bin_count(const uint32_t *__restrict__ src, size_t count,
                                                  const uint64_t *__restrict__ starts, uint32_t parts,
                                                  uint32_t *__restrict__ block_hist, uint32_t nblocks) {
  __shared__ uint32_t s_h[BIN_MAX_PARTS];
  __shared__ uint64_t s_st[BIN_MAX_PARTS];
  if (threadIdx.x < parts) {
    s_h[threadIdx.x] = 0;
    s_st[threadIdx.x] = starts[threadIdx.x];
  }
  __syncthreads();
  const size_t base = (size_t)blockIdx.x * prim::SORT_TILE;
  for (int r = 0; r < prim::SORT_ROUNDS; r++) {
    const size_t i = base + (size_t)r * BT + threadIdx.x;
    if (i < count) atomicAdd(&s_h[owner_of(s_st, parts, src[i])], 1u);
  }
  __syncthreads();
  if (threadIdx.x < parts) block_hist[(size_t)threadIdx.x * nblocks + blockIdx.x] = s_h[threadIdx.x];
}

// stable scatter: same warp-synchronous ranking as the radix sort, digit = owner
__global__ void __launch_bounds__(BT) k_bin_scatter(const uint32_t *__restrict__ src, const uint32_t *__restrict__ dst,
                                                    const uint32_t *__restrict__ val, size_t count,
                                                    const uint64_t *__restrict__ starts, uint32_t parts,
                                                    const uint32_t *__restrict__ offs, uint32_t nblocks,
                                                    uint32_t *__restrict__ out_src, uint32_t *__restrict__ out_dst,
                                                    uint32_t *__restrict__ out_val) {
  __shared__ uint32_t s_cnt[prim::SORT_WARPS][BIN_MAX_PARTS];
  __shared__ uint64_t s_st[BIN_MAX_PARTS];
  for (int d = threadIdx.x; d < prim::SORT_WARPS * BIN_MAX_PARTS; d += BT) (&s_cnt[0][0])[d] = 0;
  if (threadIdx.x < parts) s_st[threadIdx.x] = starts[threadIdx.x];
  __syncthreads();
  const unsigned w = threadIdx.x >> 5, l = lane_id(), lt = lanemask_lt();
  const size_t wbase = (size_t)blockIdx.x * prim::SORT_TILE + (size_t)w * (32 * prim::SORT_ROUNDS);
  uint32_t own[prim::SORT_ROUNDS], rank[prim::SORT_ROUNDS];
#pragma unroll
  for (int r = 0; r < prim::SORT_ROUNDS; r++) {
    const size_t i = wbase + (size_t)r * 32 + l;
    const bool valid = i < count;
    const uint32_t d = valid ? owner_of(s_st, parts, src[i]) : 0x1FFu;
    own[r] = d;
    const unsigned peers = __match_any_sync(0xFFFFFFFFu, d);
    uint32_t base = 0;
    if (valid) base = s_cnt[w][d];
    __syncwarp();
    if (valid && (peers & lt) == 0) s_cnt[w][d] = base + __popc(peers);
    __syncwarp();
    rank[r] = base + __popc(peers & lt);
  }
  __syncthreads();
  if (threadIdx.x < parts) {
    uint32_t run = offs[(size_t)threadIdx.x * nblocks + blockIdx.x];
    for (int ww = 0; ww < prim::SORT_WARPS; ww++) {
      uint32_t t = s_cnt[ww][threadIdx.x];
      s_cnt[ww][threadIdx.x] = run;
      run += t;
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < prim::SORT_ROUNDS; r++) {
    const size_t i = wbase + (size_t)r * 32 + l;
    if (i < count) {
      const uint32_t d = own[r];
      const uint32_t pos = s_cnt[w][d] + rank[r];
      out_src[pos] = src[i] - (uint32_t)s_st[d];  // shard-local id (reference PPPCSR.cpp:46-52)
      out_dst[pos] = dst[i];
      if (out_val) out_val[pos] = val ? val[i] : 1u;
    }
  }
}

}  // namespace batch
