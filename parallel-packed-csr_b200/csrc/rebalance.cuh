// rebalance.cuh -- the rebalance kernel: merge the kept items of a window with the batch's inserts,
// drop the tombstones, and spread the result evenly over the window's leaves, fixing the vertex
// sentinels' back pointers in the same pass.
// Replaces reference PCSR::redistribute + fix_sentinel + slide_right/slide_left + double_list/half_list
// (src/pcsr/PCSR.cpp:222-249, 168-183, 326-390, 251-320).
//
// Formulation (gather, output-driven).  For a window of m source leaves the merged sequence is the
// concatenation over source leaves i of merge(kept(i), inserts(i)); its ranks are
//     rank(kept item at offset f of leaf i) = R[i] + #kept before f + #inserts of i whose predecessor < f
//     rank(q-th insert of leaf i)           = R[i] + q + #kept items of i at offsets <= offset(pred)
// with R = exclusive scan of the post-batch leaf counts.  Output leaf o of m_out receives the ranks
// [floor(o*j/m_out), floor((o+1)*j/m_out)) left-packed, the rest of the leaf is nulled.  Each CTA owns
// a chunk of CHUNK_SLOTS consecutive output slots: it finds the source leaves that feed its rank range
// by binary search in R, stages the items in shared memory at (rank - first rank) and writes the chunk
// with 16-byte stores.  Algorithmic traffic: every window slot is read once and written once.
#pragma once
#include "common.cuh"
#include "primitives.cuh"

namespace reb {

constexpr int RT = 256;
constexpr int RWARPS = RT / 32;
constexpr int CHUNK_SLOTS = 2048;  // output slots per CTA (16 KB of staging)
constexpr int TILE_LEAVES = 64;    // source leaves examined per inner iteration

struct Args {
  const uint32_t *src_dest, *src_val;  // source slots
  const uint32_t *leaf_cnt;            // source physical leaf counts (tombstones included)
  const uint32_t *rank_off;            // R[], n_leaves_src + 1
  const uint32_t *ins_off;             // first insert of each source leaf, n_leaves_src + 1
  const uint32_t *ins_dst, *ins_val, *ins_pred;
  uint32_t *out_dest_single, *out_val_single;  // target of single-CTA windows (in place)
  uint32_t *out_dest_multi, *out_val_multi;    // target of multi-CTA windows (out of place)
  uint32_t *tree_leaf_out;                     // post-rebalance leaf counts (leaf level of the tree)
  uint32_t *beg;
  const WindowDesc *windows;
  uint32_t n_windows;
  uint32_t ls_src, ls_dst;
  uint32_t m_dst_override;  // != 0: resize, the single window maps onto this many output leaves from leaf 0
  const ChunkPlan *plan;    // one entry per CTA
};

constexpr uint32_t CLAMP_THRESHOLD = 256;  // leaves with more inserts than this get their insert range clamped

__device__ __forceinline__ uint32_t upper_bound_u32(const uint32_t *a, uint32_t n, uint32_t key) {
  uint32_t lo = 0, hi = n;  // first index with a[idx] > key
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (a[mid] <= key) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}
__device__ __forceinline__ uint32_t lower_bound_u32(const uint32_t *a, uint32_t n, uint32_t key) {
  uint32_t lo = 0, hi = n;  // first index with a[idx] >= key
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (a[mid] < key) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

// One thread per chunk: which window the chunk belongs to and which source leaves feed its rank range.
__global__ void __launch_bounds__(RT) k_plan_chunks(const WindowDesc *__restrict__ windows, uint32_t n_windows,
                                                    const uint32_t *__restrict__ rank_off,
                                                    const uint32_t *__restrict__ ins_off, uint32_t ls_dst,
                                                    uint32_t m_dst_override, uint32_t n_chunks,
                                                    ChunkPlan *__restrict__ plan) {
  const uint32_t chunk = blockIdx.x * blockDim.x + threadIdx.x;
  if (chunk >= n_chunks) return;
  uint32_t lo = 0, hi = n_windows;  // last window with chunk0 <= chunk
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (windows[mid].chunk0 <= chunk) lo = mid;
    else hi = mid;
  }
  const WindowDesc w = windows[lo];
  const uint32_t m_dst = m_dst_override ? m_dst_override : w.m;
  const uint32_t CL = CHUNK_SLOTS >> ls_dst;
  const uint32_t o_lo = (chunk - w.chunk0) * CL;
  const uint32_t o_hi = min(o_lo + CL, m_dst);
  const uint64_t j = w.items;
  const uint32_t a = (uint32_t)rank_begin(o_lo, j, m_dst);
  const uint32_t b = (uint32_t)rank_begin(o_hi, j, m_dst);
  ChunkPlan p;
  p.win = lo;
  p.pad[0] = p.pad[1] = p.pad[2] = 0;
  if (b > a) {
    const uint32_t *R = rank_off + w.leaf0;
    const uint32_t R0 = R[0];
    p.i_lo = upper_bound_u32(R, w.m, R0 + a) - 1;      // source leaf holding rank a
    p.i_hi = upper_bound_u32(R, w.m, R0 + b - 1) - 1;  // source leaf holding rank b-1
    p.q_lo = ins_off[w.leaf0 + p.i_lo];
    p.q_hi = ins_off[w.leaf0 + p.i_hi + 1];
  } else {
    p.i_lo = 1;
    p.i_hi = 0;
    p.q_lo = p.q_hi = 0;
  }
  plan[chunk] = p;
}

constexpr int LEAVES_PER_WARP = TILE_LEAVES / RWARPS;
constexpr int MAX_CHUNK_LEAVES = CHUNK_SLOTS / 8;  // smallest leaf is 8 slots (N >= 32)

// ---- TMA (bulk async copy) + mbarrier helpers: sm_90+/sm_100a PTX ----------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// global -> shared bulk copy (SASS: UBLKCP); dst/src 16-byte aligned, bytes a non-zero multiple of 16
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

constexpr uint32_t INS_CAP = 1024;                      // inserts staged per round
constexpr uint32_t META_CAP = TILE_LEAVES + 8;          // leaf metadata entries staged (alignment slack included)
constexpr size_t REBALANCE_SMEM = (size_t)CHUNK_SLOTS * 8 /* staged output */ + (size_t)TILE_LEAVES * 32 * 8 /* source */ +
                                  (size_t)INS_CAP * 12 + (size_t)META_CAP * 12 + (size_t)TILE_LEAVES * 32 * 4 /* s_last */ +
                                  (size_t)(MAX_CHUNK_LEAVES + 4) * 4 + (size_t)TILE_LEAVES * 4 + 64;

// One CTA per chunk of CHUNK_SLOTS output slots.  One elected thread prefetches everything the chunk needs --
// the contiguous run of source leaves (dest[], val[]), their leaf_cnt / rank_off / ins_off entries and the insert
// run -- with bulk async copies (TMA) that signal one mbarrier, so the chunk pays ONE global-memory latency instead
// of a chain of dependent loads; all rank arithmetic then runs out of shared memory.
__global__ void __launch_bounds__(RT, 4) k_rebalance(Args A) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint32_t *s_dest = reinterpret_cast<uint32_t *>(smem_raw);              // staged items at (rank - a)
  uint32_t *s_val = s_dest + CHUNK_SLOTS;
  uint32_t *s_src_dest = s_val + CHUNK_SLOTS;                             // source leaves of the tile
  uint32_t *s_src_val = s_src_dest + TILE_LEAVES * 32;
  uint32_t *s_ins_pred = s_src_val + TILE_LEAVES * 32;                    // one round of inserts
  uint32_t *s_ins_dst = s_ins_pred + INS_CAP;
  uint32_t *s_ins_val = s_ins_dst + INS_CAP;
  uint32_t *s_cnt = s_ins_val + INS_CAP;                                  // metadata, 16-byte aligned windows
  uint32_t *s_rank = s_cnt + META_CAP;
  uint32_t *s_ioff = s_rank + META_CAP;
  uint32_t *s_last = s_ioff + META_CAP;                                   // [TILE_LEAVES][32]
  uint32_t *s_a = s_last + TILE_LEAVES * 32;                              // [MAX_CHUNK_LEAVES + 1]
  uint32_t *t_mask = s_a + MAX_CHUNK_LEAVES + 4;                          // [TILE_LEAVES]
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ uint32_t s_qb, s_qe;

  const uint32_t chunk = blockIdx.x;
  const ChunkPlan plan = A.plan[chunk];
  const WindowDesc w = A.windows[plan.win];
  const uint32_t logN_src = 1u << A.ls_src, logN_dst = 1u << A.ls_dst;
  const uint32_t m_dst = A.m_dst_override ? A.m_dst_override : w.m;
  const uint32_t dst_leaf0 = A.m_dst_override ? 0u : w.leaf0;
  const uint32_t CL = CHUNK_SLOTS >> A.ls_dst;  // output leaves per chunk
  const uint32_t o_lo = (chunk - w.chunk0) * CL;
  const uint32_t o_hi = min(o_lo + CL, m_dst);
  const uint32_t n_out = o_hi - o_lo;
  const uint64_t j = w.items;
  const bool multi = w.n_chunks > 1;
  uint32_t *out_dest = multi ? A.out_dest_multi : A.out_dest_single;
  uint32_t *out_val = multi ? A.out_val_multi : A.out_val_single;
  const unsigned warp = threadIdx.x >> 5, lane = lane_id(), lt = lanemask_lt();
  const uint32_t i_lo = plan.i_lo, i_hi = plan.i_hi;
  const bool has_items = i_lo <= i_hi;
  uint32_t phase = 0;

  // geometry of one tile's prefetch (all threads compute it; thread 0 issues)
  auto tile_leaves = [&](uint32_t tile) { return min((uint32_t)TILE_LEAVES, i_hi - tile + 1); };

  if (threadIdx.x == 0) {
    mbar_init(&s_bar, 1);
    fence_mbar_init();
  }
  __syncthreads();

  // stage 1 of a tile: source leaves + metadata (+ the insert run when it is known and fits: the common case)
  auto issue_tile = [&](uint32_t tile, bool with_inserts, uint32_t q0, uint32_t q1) {
    const uint32_t tl_n = tile_leaves(tile);
    const uint32_t leaf = w.leaf0 + tile;
    const uint32_t src_bytes = (tl_n << A.ls_src) * 4u;
    const uint32_t al = leaf & ~3u;                              // metadata windows start 16-byte aligned
    const uint32_t meta_n = ((leaf - al) + tl_n + 1 + 3) & ~3u;  // +1: ins_off of the leaf after the tile
    uint32_t bytes = 2 * src_bytes + 3 * meta_n * 4u;
    uint32_t qa = 0, qn = 0;
    if (with_inserts && q1 > q0) {
      qa = q0 & ~3u;
      qn = ((q1 - qa) + 3) & ~3u;
      bytes += 3 * qn * 4u;
    }
    mbar_expect_tx(&s_bar, bytes);
    bulk_g2s(s_src_dest, A.src_dest + ((size_t)leaf << A.ls_src), src_bytes, &s_bar);
    bulk_g2s(s_src_val, A.src_val + ((size_t)leaf << A.ls_src), src_bytes, &s_bar);
    bulk_g2s(s_cnt, A.leaf_cnt + al, meta_n * 4u, &s_bar);
    bulk_g2s(s_rank, A.rank_off + al, meta_n * 4u, &s_bar);
    bulk_g2s(s_ioff, A.ins_off + al, meta_n * 4u, &s_bar);
    if (qn) {
      bulk_g2s(s_ins_pred, A.ins_pred + qa, qn * 4u, &s_bar);
      bulk_g2s(s_ins_dst, A.ins_dst + qa, qn * 4u, &s_bar);
      bulk_g2s(s_ins_val, A.ins_val + qa, qn * 4u, &s_bar);
    }
  };
  auto issue_inserts = [&](uint32_t q0, uint32_t q1) {  // q1 - (q0 & ~3) <= INS_CAP
    const uint32_t qa = q0 & ~3u;
    const uint32_t qn = ((q1 - qa) + 3) & ~3u;
    mbar_expect_tx(&s_bar, 3 * qn * 4u);
    bulk_g2s(s_ins_pred, A.ins_pred + qa, qn * 4u, &s_bar);
    bulk_g2s(s_ins_dst, A.ins_dst + qa, qn * 4u, &s_bar);
    bulk_g2s(s_ins_val, A.ins_val + qa, qn * 4u, &s_bar);
  };

  // fast path: the chunk is fed by one tile of source leaves and its whole insert run fits one round
  const bool one_shot = has_items && (i_hi - i_lo) < TILE_LEAVES && (plan.q_hi - (plan.q_lo & ~3u)) <= INS_CAP;
  if (has_items && threadIdx.x == 0) issue_tile(i_lo, one_shot, plan.q_lo, plan.q_hi);

  // overlapped with the copies in flight: first rank of every output leaf.  One exact 64-bit division per CTA;
  // the others add floor((x*j + rem)/m_dst) whose numerator is < 2^40: double reciprocal + a +-1 fix-up is exact.
  {
    const uint64_t base_num = (uint64_t)o_lo * j;
    const uint64_t base_q = base_num / m_dst, base_r = base_num - base_q * m_dst;
    const double inv_m = 1.0 / (double)m_dst;
    for (uint32_t x = threadIdx.x; x <= n_out; x += RT) {
      const uint64_t num = (uint64_t)x * j + base_r;
      uint64_t q = (uint64_t)((double)num * inv_m);
      if (q * m_dst > num) q--;
      else if ((q + 1) * m_dst <= num) q++;
      s_a[x] = (uint32_t)(base_q + q);
    }
  }
  const uint32_t R0 = A.rank_off[w.leaf0];
  __syncthreads();
  const uint32_t a = s_a[0], b = s_a[n_out];

  if (has_items) {
    for (uint32_t tile = i_lo; tile <= i_hi; tile += TILE_LEAVES) {
      const uint32_t tl_n = tile_leaves(tile);
      const uint32_t tile_leaf0 = w.leaf0 + tile;
      const uint32_t mo = tile_leaf0 & 3u;  // offset of the tile's first leaf inside the aligned metadata window
      if (tile != i_lo) {
        __syncthreads();  // everyone is done with the previous tile's buffers
        if (threadIdx.x == 0) issue_tile(tile, false, 0, 0);
      }
      for (uint32_t x = threadIdx.x * 4; x < tl_n * 32; x += RT * 4)
        *reinterpret_cast<uint4 *>(s_last + x) = make_uint4(0u, 0u, 0u, 0u);
      mbar_wait(&s_bar, phase);
      phase ^= 1;
      // A1: one warp per leaf: the live prefix comes from shared memory now; publish the kept mask
      uint32_t d[LEAVES_PER_WARP], v[LEAVES_PER_WARP];
#pragma unroll
      for (int k = 0; k < LEAVES_PER_WARP; k++) {
        const uint32_t li = warp + k * RWARPS;
        if (li >= tl_n) break;  // warp-uniform early exit: a 2x expansion feeds a chunk from ~32 leaves, not 64
        d[k] = 0;
        v[k] = 0;
        if (lane < s_cnt[mo + li]) {
          d[k] = s_src_dest[(li << A.ls_src) + lane];
          v[k] = s_src_val[(li << A.ls_src) + lane];
        }
        const unsigned mask = __ballot_sync(0xFFFFFFFFu, v[k] != 0u);  // tombstones (val 0) drop out here
        if (lane == 0) t_mask[li] = mask;
      }
      __syncthreads();
      const uint32_t *t_ioff = s_ioff + mo;
      const uint32_t *t_rank = s_rank + mo;
      // Only the first and last source leaf of the chunk can straddle its rank range [a,b).  Inserts outside the
      // range are skipped by the test below, so clamping is only done for hub leaves whose insert run spans many
      // chunks (CTA-uniform condition).
      const bool clamp_lo = tile == i_lo && (t_ioff[1] - t_ioff[0]) > CLAMP_THRESHOLD;
      const bool clamp_hi = tile + tl_n - 1 == i_hi && (t_ioff[tl_n] - t_ioff[tl_n - 1]) > CLAMP_THRESHOLD;
      uint32_t q_begin = t_ioff[0], q_end = t_ioff[tl_n];
      if (clamp_lo || clamp_hi) {
        if (threadIdx.x == 0) {
          uint32_t qb = t_ioff[0], qe = t_ioff[tl_n];
          if (clamp_lo) {
            const uint32_t io = t_ioff[0], ie = t_ioff[1];
            const uint32_t mask = t_mask[0], Ri = t_rank[0] - R0;
            uint32_t lo = io, hi = ie;  // first q with rank(q) >= a
            while (lo < hi) {
              const uint32_t mid = (lo + hi) >> 1;
              const uint32_t f = A.ins_pred[mid] & (logN_src - 1u);
              const uint32_t r = Ri + (mid - io) + (uint32_t)__popc(mask & ((2u << f) - 1u));
              if (r < a) lo = mid + 1;
              else hi = mid;
            }
            qb = lo;
          }
          if (clamp_hi) {
            const uint32_t io = t_ioff[tl_n - 1], ie = t_ioff[tl_n];
            const uint32_t mask = t_mask[tl_n - 1], Ri = t_rank[tl_n - 1] - R0;
            uint32_t lo = io, hi = ie;  // first q with rank(q) >= b
            while (lo < hi) {
              const uint32_t mid = (lo + hi) >> 1;
              const uint32_t f = A.ins_pred[mid] & (logN_src - 1u);
              const uint32_t r = Ri + (mid - io) + (uint32_t)__popc(mask & ((2u << f) - 1u));
              if (r < b) lo = mid + 1;
              else hi = mid;
            }
            qe = lo;
          }
          s_qb = qb;
          s_qe = max(qb, qe);
        }
        __syncthreads();
        q_begin = s_qb;
        q_end = s_qe;
      }
      // B: the tile's inserts, one round of <= INS_CAP staged entries at a time (one round in the common case,
      // already in flight with the tile): rank = R[i] + index in the leaf's run + kept items up to the predecessor
      for (uint32_t q0 = q_begin; q0 < q_end;) {
        const uint32_t qa = q0 & ~3u;
        const uint32_t q1 = min(q_end, qa + INS_CAP);
        if (!(one_shot && q0 == q_begin)) {
          __syncthreads();  // previous round consumed
          if (threadIdx.x == 0) issue_inserts(q0, q1);
          mbar_wait(&s_bar, phase);
          phase ^= 1;
        }
        const uint32_t sa = one_shot ? (plan.q_lo & ~3u) : qa;  // global index of staged entry 0
        for (uint32_t q = q0 + threadIdx.x; q < q1; q += RT) {
          const uint32_t pred = s_ins_pred[q - sa];
          const uint32_t li = (pred >> A.ls_src) - tile_leaf0;
          const uint32_t f = pred & (logN_src - 1u);
          const uint32_t t = q - t_ioff[li];
          const uint32_t r = t_rank[li] - R0 + t + (uint32_t)__popc(t_mask[li] & ((2u << f) - 1u));
          atomicMax(&s_last[li * 32 + f], t + 1u);
          if (r >= a && r < b) {
            s_dest[r - a] = s_ins_dst[q - sa];
            s_val[r - a] = s_ins_val[q - sa];
          }
        }
        q0 = q1;
      }
      __syncthreads();
      // A2: kept items: rank = R[i] + kept before + inserts hanging on earlier offsets.  The inserts of a leaf are
      // ordered by predecessor, so that count is s_last of the nearest earlier offset that has any.
#pragma unroll
      for (int k = 0; k < LEAVES_PER_WARP; k++) {
        const uint32_t li = warp + k * RWARPS;
        if (li >= tl_n) break;  // warp-uniform
        const unsigned mask = t_mask[li];
        const uint32_t last = s_last[li * 32 + lane];
        const unsigned hang_all = __ballot_sync(0xFFFFFFFFu, last != 0u);
        uint32_t ib = 0;
        if (hang_all) {  // warp-uniform: most leaves of a sparse batch have no inserts at all
          const unsigned hang = hang_all & lt;
          ib = __shfl_sync(0xFFFFFFFFu, last, hang ? 31 - __clz(hang) : 0);
          if (!hang) ib = 0;
        }
        if ((mask >> lane) & 1u) {
          if ((li == 0 && clamp_lo) || (li == tl_n - 1 && clamp_hi)) {
            // clamped leaf: s_last only saw part of its inserts -> count them in the sorted list instead
            const uint32_t io = t_ioff[li], ic = t_ioff[li + 1] - io;
            ib = lower_bound_u32(A.ins_pred + io, ic, ((tile_leaf0 + li) << A.ls_src) + lane);
          }
          const uint32_t r = t_rank[li] - R0 + (uint32_t)__popc(mask & lt) + ib;
          if (r >= a && r < b) {
            s_dest[r - a] = d[k];
            s_val[r - a] = v[k];
          }
        }
      }
    }
  }
  __syncthreads();
  // write-out: 4 consecutive slots per thread, 16-byte stores to dest[] and val[]
  const uint32_t out_slots = n_out << A.ls_dst;
  const size_t chunk_slot0 = (size_t)(dst_leaf0 + o_lo) << A.ls_dst;
  for (uint32_t x = threadIdx.x * 4; x < out_slots; x += RT * 4) {
    const uint32_t ol = x >> A.ls_dst;
    const uint32_t f0 = x & (logN_dst - 1u);
    const uint32_t a_o = s_a[ol], b_o = s_a[ol + 1];
    const uint32_t base = a_o - a + f0;  // staging index of slot f0
    const uint32_t live_n = b_o - a_o > f0 ? min(4u, b_o - a_o - f0) : 0u;  // live slots among the 4
    uint4 dd = make_uint4(0u, 0u, 0u, 0u), vv = make_uint4(0u, 0u, 0u, 0u);
    if (live_n > 0) { dd.x = s_dest[base]; vv.x = s_val[base]; }
    if (live_n > 1) { dd.y = s_dest[base + 1]; vv.y = s_val[base + 1]; }
    if (live_n > 2) { dd.z = s_dest[base + 2]; vv.z = s_val[base + 2]; }
    if (live_n > 3) { dd.w = s_dest[base + 3]; vv.w = s_val[base + 3]; }
    // fix_sentinel (reference PCSR.cpp:168-183): a sentinel that lands here refreshes its vertex's back pointer
    if (max(max(dd.x, dd.y), max(dd.z, dd.w)) == PPCSR_SENT) {  // rare: ~1 slot in 16+ holds a sentinel
      if (dd.x == PPCSR_SENT) A.beg[vv.x - 1u] = (uint32_t)(chunk_slot0 + x);
      if (dd.y == PPCSR_SENT) A.beg[vv.y - 1u] = (uint32_t)(chunk_slot0 + x + 1);
      if (dd.z == PPCSR_SENT) A.beg[vv.z - 1u] = (uint32_t)(chunk_slot0 + x + 2);
      if (dd.w == PPCSR_SENT) A.beg[vv.w - 1u] = (uint32_t)(chunk_slot0 + x + 3);
    }
    *reinterpret_cast<uint4 *>(out_dest + chunk_slot0 + x) = dd;
    *reinterpret_cast<uint4 *>(out_val + chunk_slot0 + x) = vv;
  }
  for (uint32_t x = threadIdx.x; x < n_out; x += RT) A.tree_leaf_out[dst_leaf0 + o_lo + x] = s_a[x + 1] - s_a[x];
}

// ---------------------------------------------------------------------------------------------------------
// Small windows (<= SMALL_MAX_LEAVES leaves, the overwhelmingly common case of a steady-state batch: a touched
// leaf that stays within its bounds is its own window): ONE WARP per window, no block barriers, in place.
// Same rank arithmetic as k_rebalance; the window's items are staged in the warp's 2 KB slice of shared memory.
// ---------------------------------------------------------------------------------------------------------
constexpr int SMALL_MAX_LEAVES = 8;
constexpr int SMALL_MAX_SLOTS = SMALL_MAX_LEAVES * 32;

struct SmallArgs {
  uint32_t *dest, *val;          // rebalanced in place
  const uint32_t *leaf_cnt, *rank_off, *ins_off;
  const uint32_t *ins_dst, *ins_val, *ins_pred;
  uint32_t *tree_leaf_out, *beg;
  const WindowDesc *windows;
  uint32_t n_windows;            // one warp per window of the list; chunked (large) windows are skipped
  uint32_t ls;
};

__global__ void __launch_bounds__(RT) k_rebalance_small(SmallArgs A) {
  __shared__ uint32_t s_dest[RWARPS][SMALL_MAX_SLOTS];
  __shared__ uint32_t s_val[RWARPS][SMALL_MAX_SLOTS];
  __shared__ uint32_t s_last[RWARPS][SMALL_MAX_LEAVES][32];
  __shared__ uint32_t s_mask[RWARPS][SMALL_MAX_LEAVES], s_rank[RWARPS][SMALL_MAX_LEAVES],
      s_ioff[RWARPS][SMALL_MAX_LEAVES + 1];
  const unsigned warp = threadIdx.x >> 5, lane = lane_id(), lt = lanemask_lt();
  const uint32_t wid = blockIdx.x * RWARPS + warp;
  if (wid >= A.n_windows) return;  // whole warp exits together
  const WindowDesc w = A.windows[wid];
  if (w.n_chunks != 0) return;     // a large window: k_rebalance owns it
  const uint32_t m = w.m, j = w.items, logN = 1u << A.ls;
  const uint32_t R0 = A.rank_off[w.leaf0];
  uint32_t *sd = s_dest[warp], *sv = s_val[warp];

  // per-leaf metadata: lane k < m owns leaf k
  uint32_t my_cnt = 0;
  if (lane <= m) {
    const uint32_t i = w.leaf0 + lane;
    s_ioff[warp][lane] = A.ins_off[i];
    if (lane < m) {
      my_cnt = A.leaf_cnt[i];
      s_rank[warp][lane] = A.rank_off[i] - R0;
    }
  }
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
    if (f0 == 0) A.tree_leaf_out[w.leaf0 + ol] = b_o - a_o;
  }
}

// copy the chunks of multi-CTA windows back from the out-of-place target into the live array
__global__ void __launch_bounds__(RT) k_copy_back(const WindowDesc *__restrict__ windows,
                                                  const ChunkPlan *__restrict__ plan, uint32_t ls, const uint32_t *__restrict__ alt_dest,
                                                  const uint32_t *__restrict__ alt_val, uint32_t *__restrict__ dest,
                                                  uint32_t *__restrict__ val) {
  const uint32_t chunk = blockIdx.x;
  const WindowDesc w = windows[plan[chunk].win];
  if (w.n_chunks <= 1) return;
  const uint32_t CL = CHUNK_SLOTS >> ls;
  const uint32_t o_lo = (chunk - w.chunk0) * CL;
  const uint32_t o_hi = min(o_lo + CL, w.m);
  const size_t base = (size_t)(w.leaf0 + o_lo) << ls;
  const uint32_t slots = (o_hi - o_lo) << ls;
  for (uint32_t x = threadIdx.x * 4; x < slots; x += RT * 4) {
    *reinterpret_cast<uint4 *>(dest + base + x) = *reinterpret_cast<const uint4 *>(alt_dest + base + x);
    *reinterpret_cast<uint4 *>(val + base + x) = *reinterpret_cast<const uint4 *>(alt_val + base + x);
  }
}

// initial layout: src_n sentinels spread evenly (reference PCSR::PCSR, PCSR.cpp:796-837, in leaf-packed form)
__global__ void k_init_sentinels(uint32_t *__restrict__ dest, uint32_t *__restrict__ val, uint32_t *__restrict__ beg,
                                 uint32_t first_vertex, uint32_t count, uint32_t n_total, uint32_t m, uint32_t ls) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= count) return;
  // rank k of n_total items spread over m leaves: leaf = last o with floor(o*n/m) <= k
  const uint64_t o = (((uint64_t)k + 1) * m - 1) / n_total;
  const uint32_t f = k - (uint32_t)rank_begin(o, n_total, m);
  const size_t slot = ((size_t)o << ls) + f;
  dest[slot] = PPCSR_SENT;
  val[slot] = first_vertex + k + 1u;
  beg[first_vertex + k] = (uint32_t)slot;
}
__global__ void k_init_leaf_counts(uint32_t *__restrict__ leaf_cnt, uint32_t *__restrict__ tree, uint32_t n_total,
                                   uint32_t m) {
  const uint32_t o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= m) return;
  const uint32_t c = (uint32_t)(rank_begin((uint64_t)o + 1, n_total, m) - rank_begin(o, n_total, m));
  leaf_cnt[o] = c;
  tree[m + o] = c;
}
__global__ void k_set_u32(uint32_t *p, uint32_t v) { *p = v; }
__global__ void k_copy_u32(uint32_t *__restrict__ dst, const uint32_t *__restrict__ src, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
}

}  // namespace reb
