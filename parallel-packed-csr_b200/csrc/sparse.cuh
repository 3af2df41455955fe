// sparse.cuh -- the small-batch path: a batch that touches a few leaves of a large array costs O(touched * H)
// device work and ONE host synchronisation, instead of the O(leaves) passes (post-batch counts, tree rebuild, touched
// list, two offset scans) and the three host round trips of the general window path.
// Same decisions as the general path -- reference PCSR::insert / remove walk-ups (src/pcsr/PCSR.cpp:578-591,616-628)
// through win::k_select -- on the same data; only how the inputs of the selection are produced differs:
//   * touched leaves: appended by k_locate itself (the first update to reach a leaf in this batch, epoch stamps);
//   * count tree: the touched leaves add their (inserted - deleted) to their ancestors below the top TOP_LEVELS levels
//     (k_sp_tree); the top levels, where every leaf would hit the same few words, are summed again from the level below
//     by the last CTA to finish; the tree is never rebuilt;
//   * windows: the highest marked ancestor of every touched leaf, de-duplicated by a compare-and-swap on its mark
//     (k_sp_windows) instead of adjacency in a sorted list;
//   * rebalance: when every chosen window is small (<= reb::SMALL_MAX_LEAVES leaves, the overwhelmingly common case)
//     one warp per window computes the window-relative ranks and insert offsets on the fly (k_sp_rebalance) -- a prefix
//     over the window's leaf counts, and a binary search for the window's first insert: its inserts are one contiguous
//     run of the key-ordered list -- instead of global rank / insert-offset scans, rewrites the window in place and
//     restores the "all per-leaf batch counters are clear" state for the next batch.
// Anything else (a window larger than that, a root out of bounds, a dst wider than the speculated sort width) leaves the
// shard as the general path expects it and the host continues there.
#pragma once
#include "common.cuh"
#include "rebalance.cuh"

namespace sp {

constexpr int ST = 256;
constexpr uint32_t CLAIMED = 0x80000000u;  // mark[w] == epoch | CLAIMED: the window of node w has been emitted

constexpr uint32_t TOP_LEVELS = 8;         // tree levels 0 .. TOP_LEVELS - 1 (255 nodes) are summed, not incremented
constexpr uint32_t TOP_NODES = 1u << TOP_LEVELS;  // nodes of level TOP_LEVELS: heap indices TOP_NODES .. 2 TOP_NODES - 1

// Every touched leaf adds its net change to its ancestors from the leaf level up to level TOP_LEVELS and records what
// touched it for the invariant checker; the last CTA to finish sums levels TOP_LEVELS - 1 .. 0 from level TOP_LEVELS
// (with 100 K touched leaves the root alone would take 100 K serialised atomics).  Needs n_leaves >= 2 TOP_NODES.
__global__ void __launch_bounds__(ST) k_sp_tree(const uint32_t *__restrict__ touched, BatchScalars *sc,
                                                const uint32_t *__restrict__ ins_cnt,
                                                const uint32_t *__restrict__ del_cnt, uint32_t n_leaves,
                                                uint32_t *__restrict__ tree, uint32_t *__restrict__ touched_flags) {
  __shared__ uint32_t s_lvl[TOP_NODES];
  __shared__ bool s_last;
  if (sc->sparse_abort) return;
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < (size_t)sc->n_touched) {
    const uint32_t l = touched[t];
    const uint32_t ic = ins_cnt[l], dc = del_cnt[l];
    touched_flags[t] = (ic ? 1u : 0u) | (dc ? 2u : 0u);
    const uint32_t delta = ic - dc;  // modular: a net loss adds 2^32 - k
    if (delta != 0u)
      for (uint32_t node = n_leaves + l; node >= TOP_NODES; node >>= 1) atomicAdd(&tree[node], delta);
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(&sc->sp_blocks_done, 1u) + 1u == gridDim.x;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  static_assert(ST == TOP_NODES, "one thread per node of level TOP_LEVELS");
  s_lvl[threadIdx.x] = __ldcg(&tree[TOP_NODES + threadIdx.x]);
  __syncthreads();
  for (uint32_t w = TOP_NODES >> 1; w >= 1u; w >>= 1) {  // level of w nodes: heap indices w .. 2 w - 1
    uint32_t v = 0;
    if (threadIdx.x < w) v = s_lvl[2 * threadIdx.x] + s_lvl[2 * threadIdx.x + 1];
    __syncthreads();
    if (threadIdx.x < w) {
      s_lvl[threadIdx.x] = v;
      tree[w + threadIdx.x] = v;
    }
    __syncthreads();
  }
}

__device__ __forceinline__ uint32_t highest_marked_any(const uint32_t *__restrict__ mark, uint32_t epoch, uint32_t node) {
  uint32_t best = 0;
  while (node >= 1) {
    if ((mark[node] & ~CLAIMED) == epoch) best = node;  // claimed by a concurrent thread or not: it is marked
    node >>= 1;
  }
  return best;
}

// One window per maximal marked node: every touched leaf finds its highest marked ancestor; the first one to claim it
// (compare-and-swap on the mark) emits the descriptor.  The three counters are summed per warp first.
__global__ void __launch_bounds__(ST) k_sp_windows(const uint32_t *__restrict__ touched, uint32_t *mark, uint32_t epoch,
                                                   const uint32_t *__restrict__ tree, uint32_t n_leaves, uint32_t logN,
                                                   WindowDesc *windows, BatchScalars *sc) {
  if (sc->sparse_abort || sc->root_violation) return;
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool mine = false;
  uint32_t w = 0;
  if (t < (size_t)sc->n_touched) {
    w = highest_marked_any(mark, epoch, n_leaves + touched[t]);
    mine = atomicCAS(&mark[w], epoch, epoch | CLAIMED) == epoch;
  }
  const unsigned mm = __ballot_sync(0xFFFFFFFFu, mine);
  if (!mm) return;
  const uint32_t depth = mine ? 31u - (uint32_t)__clz(w) : 0u;
  const uint32_t m = mine ? n_leaves >> depth : 0u;
  const bool small = m <= (uint32_t)reb::SMALL_MAX_LEAVES;
  const uint32_t all_m = __reduce_add_sync(0xFFFFFFFFu, m);
  const uint32_t n_small = __reduce_add_sync(0xFFFFFFFFu, mine && small ? 1u : 0u);
  unsigned long long at = 0;
  const unsigned leader = (unsigned)__ffs(mm) - 1u;
  if (lane_id() == leader) {
    at = atomicAdd(&sc->n_windows, (unsigned long long)__popc(mm));
    atomicAdd(&sc->window_slots, (unsigned long long)all_m * logN);
    if (n_small) atomicAdd(&sc->n_small, (unsigned long long)n_small);
  }
  at = __shfl_sync(0xFFFFFFFFu, at, leader);
  if (mine) {
    WindowDesc d;
    d.node = w;
    d.m = m;
    d.leaf0 = (w - (1u << depth)) * m;
    d.items = tree[w];
    d.n_chunks = small ? 0u : 1u;  // != 0: not a one-warp window (the host falls back to the general path)
    d.chunk0 = 0;
    windows[at + __popc(mm & lanemask_lt())] = d;
  }
}

struct RebArgs {
  uint32_t *dest, *val;  // rebalanced in place
  uint32_t *leaf_cnt, *tree, *ins_cnt, *del_cnt;
  const uint32_t *ins_dst, *ins_val, *ins_pred;
  uint32_t *beg;
  const WindowDesc *windows;
  BatchScalars *sc;
  uint32_t n_leaves, ls;
};

// One warp per window of <= SMALL_MAX_LEAVES leaves, in place: reb::k_rebalance_small with the window-relative ranks
// (warp prefix over the post-batch leaf counts) and insert offsets (first insert of the window + prefix over the
// per-leaf insert counts: a window's inserts are one contiguous run of the key-ordered list) computed on the fly.
// Also rewrites leaf_cnt and the window's subtree of the count tree, and clears the batch counters of its leaves.
__global__ void __launch_bounds__(reb::RT) k_sp_rebalance(RebArgs A) {
  using namespace reb;
  __shared__ uint32_t s_dest[RWARPS][SMALL_MAX_SLOTS];
  __shared__ uint32_t s_val[RWARPS][SMALL_MAX_SLOTS];
  __shared__ uint32_t s_last[RWARPS][SMALL_MAX_LEAVES][32];
  __shared__ uint32_t s_mask[RWARPS][SMALL_MAX_LEAVES], s_rank[RWARPS][SMALL_MAX_LEAVES],
      s_ioff[RWARPS][SMALL_MAX_LEAVES + 1];
  const unsigned warp = threadIdx.x >> 5, lane = lane_id(), lt = lanemask_lt();
  const uint32_t wid = blockIdx.x * RWARPS + warp;
  // the whole grid agrees: this path only finishes batches whose windows are ALL small and whose root is in bounds
  const unsigned long long n_windows = A.sc->n_windows;
  const bool ok = !A.sc->sparse_abort && !A.sc->root_violation && n_windows == A.sc->n_small;
  if (wid == 0 && lane == 0) A.sc->sparse_done = ok ? 1u : 0u;
  if (!ok || wid >= n_windows) return;  // whole warp exits together
  const WindowDesc w = A.windows[wid];
  const uint32_t m = w.m, j = w.items, logN = 1u << A.ls;
  uint32_t *sd = s_dest[warp], *sv = s_val[warp];

  // per-leaf metadata: lane k < m owns leaf k
  uint32_t my_cnt = 0, my_ins = 0, my_new = 0;
  if (lane < m) {
    const uint32_t i = w.leaf0 + lane;
    my_cnt = A.leaf_cnt[i];
    my_ins = A.ins_cnt[i];
    my_new = my_cnt + my_ins - A.del_cnt[i];
  }
  uint32_t rank_incl = my_new, ins_incl = my_ins;
#pragma unroll
  for (int d = 1; d < SMALL_MAX_LEAVES; d <<= 1) {
    const uint32_t a = __shfl_up_sync(0xFFFFFFFFu, rank_incl, d), b = __shfl_up_sync(0xFFFFFFFFu, ins_incl, d);
    if ((int)lane >= d) {
      rank_incl += a;
      ins_incl += b;
    }
  }
  const uint32_t ins_total = __shfl_sync(0xFFFFFFFFu, ins_incl, SMALL_MAX_LEAVES - 1);
  // the window's inserts are one contiguous run of the key-ordered insert list (sorted by predecessor slot): its first
  // element is the first insert whose predecessor lies at or beyond the window's first slot
  uint32_t first = 0u;
  if (ins_total) {  // warp-uniform
    const uint32_t slot0 = w.leaf0 << A.ls;
    uint32_t lo = 0, hi = (uint32_t)A.sc->n_inserted;
    while (lo < hi) {
      const uint32_t mid = (lo + hi) >> 1;
      if (A.ins_pred[mid] < slot0) lo = mid + 1;
      else hi = mid;
    }
    first = lo;
  }
  if (lane < m) {
    s_rank[warp][lane] = rank_incl - my_new;
    s_ioff[warp][lane] = first + ins_incl - my_ins;
  }
  if (lane == 0) s_ioff[warp][m] = first + ins_total;
  for (uint32_t x = lane; x < m * 32; x += 32) (&s_last[warp][0][0])[x] = 0;
  uint32_t d[SMALL_MAX_LEAVES], v[SMALL_MAX_LEAVES];
#pragma unroll
  for (int k = 0; k < SMALL_MAX_LEAVES; k++) {
    d[k] = 0;
    v[k] = 0;
    const uint32_t cnt_k = __shfl_sync(0xFFFFFFFFu, my_cnt, k);
    if ((uint32_t)k < m && lane < cnt_k) {
      const size_t slot = ((size_t)(w.leaf0 + k) << A.ls) + lane;
      d[k] = A.dest[slot];
      v[k] = A.val[slot];
    }
  }
#pragma unroll
  for (int k = 0; k < SMALL_MAX_LEAVES; k++) {
    const unsigned mask = __ballot_sync(0xFFFFFFFFu, v[k] != 0u);
    if ((uint32_t)k < m && lane == 0) s_mask[warp][k] = mask;
  }
  __syncwarp();
  // inserts of the window
  const uint32_t q_end = s_ioff[warp][m];
  for (uint32_t q = s_ioff[warp][0] + lane; q < q_end; q += 32) {
    const uint32_t pred = A.ins_pred[q];
    const uint32_t k = (pred >> A.ls) - w.leaf0;
    const uint32_t f = pred & (logN - 1u);
    const uint32_t t = q - s_ioff[warp][k];
    const uint32_t r = s_rank[warp][k] + t + (uint32_t)__popc(s_mask[warp][k] & ((2u << f) - 1u));
    atomicMax(&s_last[warp][k][f], t + 1u);
    sd[r] = A.ins_dst[q];
    sv[r] = A.ins_val[q];
  }
  __syncwarp();
  // kept items
#pragma unroll
  for (int k = 0; k < SMALL_MAX_LEAVES; k++) {
    if ((uint32_t)k < m) {  // warp-uniform
      const unsigned mask = s_mask[warp][k];
      const uint32_t last = s_last[warp][k][lane];
      const unsigned hang = __ballot_sync(0xFFFFFFFFu, last != 0u) & lt;
      uint32_t ib = __shfl_sync(0xFFFFFFFFu, last, hang ? 31 - __clz(hang) : 0);
      if (!hang) ib = 0;
      if ((mask >> lane) & 1u) {
        const uint32_t r = s_rank[warp][k] + (uint32_t)__popc(mask & lt) + ib;
        sd[r] = d[k];
        sv[r] = v[k];
      }
    }
  }
  __syncwarp();
  // write-out (in place: every source slot of the window has been read above)
  const uint32_t out_slots = m << A.ls;
  const size_t slot0 = (size_t)w.leaf0 << A.ls;
  for (uint32_t x = lane * 4; x < out_slots; x += 32 * 4) {
    const uint32_t ol = x >> A.ls, f0 = x & (logN - 1u);
    const uint32_t a_o = (ol * j) / m, b_o = ((ol + 1) * j) / m;  // j <= m*(logN-1): 32-bit is plenty
    const uint32_t live_n = b_o - a_o > f0 ? min(4u, b_o - a_o - f0) : 0u;
    const uint32_t base = a_o + f0;
    uint4 dd = make_uint4(0u, 0u, 0u, 0u), vv = make_uint4(0u, 0u, 0u, 0u);
    if (live_n > 0) { dd.x = sd[base]; vv.x = sv[base]; }
    if (live_n > 1) { dd.y = sd[base + 1]; vv.y = sv[base + 1]; }
    if (live_n > 2) { dd.z = sd[base + 2]; vv.z = sv[base + 2]; }
    if (live_n > 3) { dd.w = sd[base + 3]; vv.w = sv[base + 3]; }
    if (dd.x == PPCSR_SENT) A.beg[vv.x - 1u] = (uint32_t)(slot0 + x);
    if (dd.y == PPCSR_SENT) A.beg[vv.y - 1u] = (uint32_t)(slot0 + x + 1);
    if (dd.z == PPCSR_SENT) A.beg[vv.z - 1u] = (uint32_t)(slot0 + x + 2);
    if (dd.w == PPCSR_SENT) A.beg[vv.w - 1u] = (uint32_t)(slot0 + x + 3);
    *reinterpret_cast<uint4 *>(A.dest + slot0 + x) = dd;
    *reinterpret_cast<uint4 *>(A.val + slot0 + x) = vv;
  }
  // new leaf counts, the window's subtree of the count tree (its root keeps its total), batch counters cleared
  uint32_t c = 0;
  if (lane < m) {
    c = ((lane + 1) * j) / m - (lane * j) / m;
    const uint32_t i = w.leaf0 + lane;
    A.leaf_cnt[i] = c;
    A.tree[A.n_leaves + i] = c;
    A.ins_cnt[i] = 0u;
    A.del_cnt[i] = 0u;
  }
  for (uint32_t s = 1, sh = 1; s < m; s <<= 1, sh++) {  // m is a power of two <= 8
    c += __shfl_down_sync(0xFFFFFFFFu, c, s);
    if (lane < m && (lane & (2u * s - 1u)) == 0u && 2u * s < m) A.tree[(A.n_leaves + w.leaf0 + lane) >> sh] = c;
  }
}

// density bounds on the paths touched by the last (sparse) batch: the checker's k_check_bounds over the touched list
__global__ void __launch_bounds__(ST) k_sp_check_bounds(const uint32_t *__restrict__ tree,
                                                        const uint32_t *__restrict__ touched,
                                                        const uint32_t *__restrict__ touched_flags, uint32_t n_touched,
                                                        uint32_t n_leaves, uint32_t logN, int H, int check_lower,
                                                        unsigned long long *bad_upper, unsigned long long *bad_lower) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_touched) return;
  const bool ins = touched_flags[t] & 1u, del = touched_flags[t] & 2u;
  uint32_t node = n_leaves + touched[t];
  uint64_t len = logN;
  unsigned up = 0, lo = 0;
  for (int depth = H; depth >= 0; depth--) {
    const uint32_t cnt = tree[node];
    if (ins && !window_ok_upper(cnt, len, logN, depth, H)) up++;
    if (del && check_lower && !window_ok_lower(cnt, len, depth, H)) lo++;
    node >>= 1;
    len <<= 1;
  }
  if (up) atomicAdd(bad_upper, (unsigned long long)up);
  if (lo) atomicAdd(bad_lower, (unsigned long long)lo);
}

}  // namespace sp
