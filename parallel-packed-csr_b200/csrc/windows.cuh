// windows.cuh -- the implicit PMA tree: live-count tree maintenance and the bottom-up choice of
// rebalance windows under the reference's density bounds.
// Replaces reference get_density / density_bound / the walks in PCSR::insert and PCSR::remove
// (src/pcsr/PCSR.cpp:126-133, 156-165, 578-591, 616-628) and the window pre-computation of
// acquire_insert_locks / acquire_remove_locks (PCSR.cpp:1012-1084, 1191-1231).
#pragma once
#include "common.cuh"
#include "primitives.cuh"

namespace win {

constexpr int WT = 256;

// post-batch live count of every leaf = old + inserted - deleted  -> leaf level of the tree
__global__ void __launch_bounds__(WT) k_leaf_new_counts(const uint32_t *__restrict__ leaf_cnt,
                                                        const uint32_t *__restrict__ ins_cnt,
                                                        const uint32_t *__restrict__ del_cnt, uint32_t n_leaves,
                                                        uint32_t *__restrict__ tree) {
  const uint32_t l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l < n_leaves) tree[n_leaves + l] = leaf_cnt[l] + ins_cnt[l] - del_cnt[l];
}

// Sums five tree levels per launch with warp shuffles: a warp loads 32 consecutive nodes of depth D
// (one coalesced 128-B line) and writes their ancestors at depths D-1 .. D-5.
__global__ void __launch_bounds__(WT) k_tree_up5(uint32_t *__restrict__ tree, uint32_t D) {
  const uint32_t level_nodes = 1u << D;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;  // position inside the level
  uint32_t v = (i < level_nodes) ? tree[level_nodes + i] : 0u;
  const unsigned l = lane_id();
#pragma unroll
  for (uint32_t k = 1; k <= 5; k++) {
    v += __shfl_down_sync(0xFFFFFFFFu, v, 1u << (k - 1));
    if (k <= D && (l & ((1u << k) - 1u)) == 0 && i < level_nodes) tree[(level_nodes + i) >> k] = v;
  }
}

// The levels above TOP_D (<= 2^TOP_D nodes) in ONE launch: a single CTA sums level by level from level TOP_D (every
// further k_tree_up5 launch would be pure launch latency: 4 more dependent launches at 2^21 leaves).
constexpr uint32_t TOP_D = 12;
__global__ void __launch_bounds__(1024) k_tree_top(uint32_t *__restrict__ tree, uint32_t D) {
  for (uint32_t d = D; d >= 1u; d--) {  // level d - 1 from level d
    const uint32_t nodes = 1u << (d - 1u);
    for (uint32_t i = threadIdx.x; i < nodes; i += blockDim.x) tree[nodes + i] = tree[2u * (nodes + i)] + tree[2u * (nodes + i) + 1u];
    __syncthreads();
  }
}

inline int tree_rebuild(ppcsr_shard *s, uint32_t *tree, uint32_t H) {
  int D = (int)H;
  for (; D > (int)TOP_D; D -= 5) {
    const uint32_t nodes = 1u << D;
    s->launches++;
    k_tree_up5<<<div_up(nodes, WT), WT, 0, s->stream>>>(tree, (uint32_t)D);
  }
  if (D > 0) {
    s->launches++;
    k_tree_top<<<1, 1024, 0, s->stream>>>(tree, (uint32_t)D);
  }
  CUDA_TRY(cudaGetLastError());
  return PPCSR_OK;
}

// touched leaves = leaves that gained or lost an item in this batch (key order == leaf order)
struct InTouched {
  const uint32_t *ins_cnt, *del_cnt;
  __device__ uint32_t operator()(size_t l) const { return (ins_cnt[l] | del_cnt[l]) ? 1u : 0u; }
};
struct OutTouched {
  uint32_t *touched;
  __device__ void operator()(size_t l, uint32_t ex, uint32_t own) const {
    if (own) touched[ex] = (uint32_t)l;
  }
};

// For one touched leaf: walk the whole path to the root with the post-batch counts.  Every ancestor
// is tested against the reference's bounds (upper if the leaf received inserts, lower if it lost
// items).  The window to rebalance is the parent of the HIGHEST violating node (the first node above
// which the path is within bounds); with no violation it is the leaf itself (the reference rewrites
// the leaf on every insert/remove, PCSR.cpp:552-560,609).  Checking all ancestors rather than stopping
// at the first in-bounds one is strictly tighter than the reference and gives invariant I4/I5 of
// SURVEY §8a on every touched path.  A violated root means double_list / half_list.
__global__ void __launch_bounds__(WT) k_select(const uint32_t *__restrict__ touched,
                                               const unsigned long long *__restrict__ n_touched,
                                               const uint32_t *__restrict__ ins_cnt,
                                               const uint32_t *__restrict__ del_cnt,
                                               const uint32_t *__restrict__ tree, uint32_t n_leaves, uint32_t logN,
                                               int H, uint32_t *__restrict__ mark, uint32_t epoch, BatchScalars *sc) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)*n_touched) return;
  const uint32_t l = touched[t];
  const bool ins = ins_cnt[l] != 0, del = del_cnt[l] != 0;
  uint32_t node = n_leaves + l;
  uint64_t len = logN;
  uint32_t top = 0;        // heap index of the highest violating node (0 = none)
  unsigned viol_bits = 0;  // what the highest violator violated
  for (int depth = H; depth >= 0; depth--) {
    const uint32_t cnt = tree[node];
    const bool bad_up = ins && !window_ok_upper(cnt, len, logN, depth, H);
    const bool bad_lo = del && !window_ok_lower(cnt, len, depth, H);
    if (bad_up || bad_lo) {
      top = node;
      viol_bits = bad_up ? 1u : 2u;
    }
    node >>= 1;
    len <<= 1;
  }
  uint32_t w;
  if (top == 0) {
    w = n_leaves + l;
  } else if (top == 1) {
    w = 1;
    atomicOr(&sc->root_violation, viol_bits);
  } else {
    w = top >> 1;
  }
  mark[w] = epoch;
}

__device__ __forceinline__ uint32_t highest_marked(const uint32_t *__restrict__ mark, uint32_t epoch, uint32_t node) {
  uint32_t best = 0;
  while (node >= 1) {
    if (mark[node] == epoch) best = node;
    node >>= 1;
  }
  return best;
}

// window (heap index of the highest marked ancestor) of every touched leaf; skipped when the root is violated
__global__ void __launch_bounds__(WT) k_touched_windows(const uint32_t *__restrict__ touched,
                                                        const unsigned long long *__restrict__ n_touched,
                                                        const uint32_t *__restrict__ mark, uint32_t epoch,
                                                        uint32_t n_leaves, const BatchScalars *__restrict__ sc,
                                                        uint32_t *__restrict__ touched_win) {
  if (sc->root_violation) return;
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)*n_touched) return;
  touched_win[t] = highest_marked(mark, epoch, n_leaves + touched[t]);
}

// One window per maximal marked node; touched leaves are sorted, so leaves of the same window are adjacent.
struct InWindowHead {
  const uint32_t *touched_win;
  __device__ uint32_t operator()(size_t t) const {
    if (t == 0) return 1u;
    return touched_win[t - 1] != touched_win[t] ? 1u : 0u;
  }
};
struct OutWindow {
  const uint32_t *touched_win, *tree;
  uint32_t n_leaves, chunk_leaves, small_max_leaves, logN;
  WindowDesc *windows;
  BatchScalars *sc;
  __device__ void operator()(size_t t, uint32_t ex, uint32_t own) const {
    if (!own) return;
    const uint32_t w = touched_win[t];
    const uint32_t depth = 31u - (uint32_t)__clz(w);
    const uint32_t m = n_leaves >> depth;
    WindowDesc d;
    d.node = w;
    d.m = m;
    d.leaf0 = (w - (1u << depth)) * m;
    d.items = tree[w];
    // small windows are rebalanced one per warp (k_rebalance_small) and take no CTA of the chunked kernel
    d.n_chunks = m <= small_max_leaves ? 0u : (m + chunk_leaves - 1) / chunk_leaves;
    d.chunk0 = 0;
    windows[ex] = d;
    // totals for the host's policy decision.  A 1 M-update batch opens ~220 K windows: three atomics each on the
    // same three words took 168 us of the 0.76 ms batch, so the lanes that are here together add up first (leaves,
    // not slots, so that 32 windows fit 32 bits) and one of them does the atomics.
    const unsigned peers = __activemask();
    const uint32_t all_m = __reduce_add_sync(peers, m);
    const uint32_t multi_m = __reduce_add_sync(peers, d.n_chunks > 1 ? m : 0u);
    const uint32_t small = __reduce_add_sync(peers, d.n_chunks == 0 ? 1u : 0u);
    if ((peers & ((1u << (threadIdx.x & 31u)) - 1u)) == 0u) {  // lowest lane of the group
      atomicAdd(&sc->window_slots, (unsigned long long)all_m * logN);
      if (multi_m) atomicAdd(&sc->multi_slots, (unsigned long long)multi_m * logN);
      if (small) atomicAdd(&sc->n_small, (unsigned long long)small);
    }
  }
};

struct InWinChunks {
  const WindowDesc *w;
  __device__ uint32_t operator()(size_t i) const { return w[i].n_chunks; }
};
struct OutWinChunk0 {
  WindowDesc *w;
  __device__ void operator()(size_t i, uint32_t ex, uint32_t) const { w[i].chunk0 = ex; }
};
}  // namespace win
