// primitives.cuh -- device-wide exclusive scan / stream compaction (the radix sort lives in sort.cuh).
// Hand-written (no CUB/Thrust), with functor input/output so the same code serves prefix sums, order-preserving
// selects and histogram offsets.  Two forms: up to SCAN_ONEPASS_MAX_TILES tiles ONE kernel -- tiles take tickets,
// publish their sum and resolve the prefix of all earlier tiles by decoupled look-back (a batch runs seven scans;
// at three launches each they were half of the launches of a small batch); beyond that the three-phase form (tile
// reduce -> spine -> tile scan), whose second read of the input is cheaper than the look-back ramp of the first wave
// of tiles (10 M-element compaction: 98 us against 114 us).
#pragma once
#include "common.cuh"

namespace prim {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;
constexpr unsigned SCAN_ONEPASS_MAX_TILES = 1024;  // 2 M elements

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t t = __shfl_up_sync(0xFFFFFFFFu, v, d);
    if ((int)lane_id() >= d) v += t;
  }
  return v;
}

// exclusive scan of one value per thread over the block; returns the exclusive prefix, *total = block sum
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t *total, uint32_t *s_warp /*[32]*/) {
  const unsigned w = threadIdx.x >> 5, l = lane_id();
  uint32_t inc = warp_incl_scan(v);
  if (l == 31) s_warp[w] = inc;
  __syncthreads();
  if (w == 0) {
    uint32_t x = (l < (blockDim.x >> 5)) ? s_warp[l] : 0;
    uint32_t xi = warp_incl_scan(x);
    s_warp[l] = xi - x;  // exclusive warp offsets
    if (l == 31) s_warp[32] = xi;
  }
  __syncthreads();
  uint32_t r = s_warp[w] + inc - v;
  *total = s_warp[32];
  __syncthreads();
  return r;
}

// ---- three-phase scan (tile reduce -> spine -> tile scan): used for LARGE inputs, see device_scan ----
template <class In>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_reduce(In in, size_t n, uint32_t *block_sums) {
  __shared__ uint32_t s_warp[33];
  const size_t base = (size_t)blockIdx.x * SCAN_TILE + threadIdx.x;  // coalesced: item k*256 + tid
  uint32_t sum = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    size_t i = base + (size_t)k * SCAN_THREADS;
    if (i < n) sum += in(i);
  }
  uint32_t total;
  block_excl_scan(sum, &total, s_warp);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// single block: exclusive scan of block_sums[0..nb) in place; total to *total32 / *total64 (nullable)
__global__ void __launch_bounds__(1024) k_scan_spine(uint32_t *block_sums, uint32_t nb, uint32_t *total32,
                                                     unsigned long long *total64) {
  __shared__ uint32_t s_warp[33];
  __shared__ uint32_t s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (uint32_t base = 0; base < nb; base += blockDim.x) {
    uint32_t i = base + threadIdx.x;
    uint32_t v = (i < nb) ? block_sums[i] : 0;
    uint32_t total;
    uint32_t ex = block_excl_scan(v, &total, s_warp);
    uint32_t carry = s_carry;
    if (i < nb) block_sums[i] = carry + ex;
    __syncthreads();
    if (threadIdx.x == 0) s_carry = carry + total;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (total32) *total32 = s_carry;
    if (total64) *total64 = s_carry;
  }
}

// padded index: thread t later reads the 8 consecutive items t*8..t*8+7 without bank conflicts
__device__ __forceinline__ uint32_t scan_pad(uint32_t j) { return j + (j >> 5); }

// Loads and stores are evaluated in coalesced order (item k*256 + tid); the tile is transposed through
// shared memory so that each thread scans 8 consecutive items.
template <class In, class Out>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(In in, Out out, size_t n, const uint32_t *block_sums) {
  __shared__ uint32_t s_warp[33];
  __shared__ uint32_t s_v[SCAN_TILE + SCAN_TILE / 32];
  __shared__ uint32_t s_ex[SCAN_TILE + SCAN_TILE / 32];
  const size_t base = (size_t)blockIdx.x * SCAN_TILE;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    const uint32_t j = k * SCAN_THREADS + threadIdx.x;
    const size_t i = base + j;
    s_v[scan_pad(j)] = (i < n) ? in(i) : 0u;
  }
  __syncthreads();
  uint32_t v[SCAN_ITEMS];
  uint32_t sum = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    v[k] = s_v[scan_pad(threadIdx.x * SCAN_ITEMS + k)];
    sum += v[k];
  }
  uint32_t total;
  uint32_t ex = block_excl_scan(sum, &total, s_warp) + block_sums[blockIdx.x];
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    s_ex[scan_pad(threadIdx.x * SCAN_ITEMS + k)] = ex;
    ex += v[k];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    const uint32_t j = k * SCAN_THREADS + threadIdx.x;
    const size_t i = base + j;
    if (i < n) out(i, s_ex[scan_pad(j)], s_v[scan_pad(j)]);
  }
}

__global__ void k_zero_totals(uint32_t *total32, unsigned long long *total64) {
  if (total32) *total32 = 0;
  if (total64) *total64 = 0;
}

// ---- single-pass scan: decoupled look-back --------------------------------------------------------------------
// One 64-bit word per tile: [63:62] status (1 = the tile's own sum, 2 = inclusive prefix over all tiles up to it),
// [61:32] epoch of the scan that wrote it, [31:0] the value.  The epoch makes words left behind by earlier scans
// read as "not there yet", so the array is never cleared between scans; tiles are numbered by a ticket counter that
// only ever grows (the host passes the first ticket of this scan), so every earlier tile is already running when a
// tile looks back.
constexpr unsigned long long SCAN_ST_AGG = 1ull << 62, SCAN_ST_INC = 2ull << 62;
__device__ __forceinline__ unsigned long long scan_word(unsigned long long status, uint32_t epoch, uint32_t value) {
  return status | ((unsigned long long)epoch << 32) | value;
}
template <class In, class Out>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_onepass(In in, Out out, size_t n, unsigned long long *tickets,
                                                               unsigned long long ticket_base,
                                                               unsigned long long *state, uint32_t epoch, uint32_t nb,
                                                               uint32_t *total32, unsigned long long *total64) {
  __shared__ uint32_t s_warp[33];
  __shared__ uint32_t s_v[SCAN_TILE + SCAN_TILE / 32];
  __shared__ uint32_t s_ex[SCAN_TILE + SCAN_TILE / 32];
  __shared__ uint32_t s_tile, s_prefix;
  if (threadIdx.x == 0) s_tile = (uint32_t)(atomicAdd(tickets, 1ull) - ticket_base);
  __syncthreads();
  const uint32_t tile = s_tile;
  const size_t base = (size_t)tile * SCAN_TILE;
  // loads in coalesced order (item k*256 + tid); the tile is transposed through shared memory so that each thread
  // scans 8 consecutive items
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    const uint32_t j = k * SCAN_THREADS + threadIdx.x;
    const size_t i = base + j;
    s_v[scan_pad(j)] = (i < n) ? in(i) : 0u;
  }
  __syncthreads();
  uint32_t v[SCAN_ITEMS];
  uint32_t sum = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    v[k] = s_v[scan_pad(threadIdx.x * SCAN_ITEMS + k)];
    sum += v[k];
  }
  uint32_t total;
  uint32_t ex = block_excl_scan(sum, &total, s_warp);
  if (threadIdx.x < 32) {  // warp 0: publish the tile's sum, then add up the tiles before it
    const unsigned lane = threadIdx.x;
    volatile unsigned long long *st = state;
    if (lane == 0) st[tile] = scan_word(tile == 0 ? SCAN_ST_INC : SCAN_ST_AGG, epoch, total);
    uint32_t prefix = 0;
    if (tile > 0) {
      int64_t j = (int64_t)tile - 1;  // lane l examines tile j - l
      for (;;) {
        const int64_t t = j - (int64_t)lane;
        // tiles before the first one count as an inclusive prefix of zero
        const unsigned long long w = t >= 0 ? st[t] : scan_word(SCAN_ST_INC, epoch, 0u);
        const bool ready = (uint32_t)((w >> 32) & 0x3FFFFFFFu) == epoch && (w >> 62) != 0ull;
        const unsigned inc = __ballot_sync(0xFFFFFFFFu, ready && (w >> 62) == 2ull);
        const unsigned not_ready = __ballot_sync(0xFFFFFFFFu, !ready);
        const unsigned first_inc = inc ? (unsigned)__ffs(inc) - 1u : 32u;
        const unsigned need = first_inc < 32u ? (2u << first_inc) - 1u : 0xFFFFFFFFu;  // lanes 0..first_inc
        if (not_ready & need) continue;  // a tile this window depends on has not published yet: look again
        uint32_t x = (need >> lane) & 1u ? (uint32_t)w : 0u;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(0xFFFFFFFFu, x, d);
        prefix += x;
        if (first_inc < 32u) break;
        j -= 32;
      }
      if (lane == 0) st[tile] = scan_word(SCAN_ST_INC, epoch, prefix + total);
    }
    if (lane == 0) {
      s_prefix = prefix;
      if (tile + 1 == nb) {
        if (total32) *total32 = prefix + total;
        if (total64) *total64 = prefix + total;
      }
    }
  }
  __syncthreads();
  ex += s_prefix;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    s_ex[scan_pad(threadIdx.x * SCAN_ITEMS + k)] = ex;
    ex += v[k];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    const uint32_t j = k * SCAN_THREADS + threadIdx.x;
    const size_t i = base + j;
    if (i < n) out(i, s_ex[scan_pad(j)], s_v[scan_pad(j)]);
  }
}

// look-back words for scans of up to `tiles` tiles (also callable ahead of time, so that a batch never allocates)
inline int reserve_scan_state(ppcsr_shard *s, size_t tiles) {
  const size_t cap0 = s->scan_state.cap;
  PPCSR_TRY(dev_reserve(s->scan_state, tiles + 1, s->stream));
  if (s->scan_state.cap != cap0 || s->scan_epoch + 1u >= (1u << 30)) {  // fresh memory, or the epoch wraps: no word may look valid
    CUDA_TRY(cudaMemsetAsync(s->scan_state.p, 0, s->scan_state.cap * sizeof(unsigned long long), s->stream));
    s->scan_epoch = 0;
  }
  if (!s->scan_ticket.p) {
    PPCSR_TRY(dev_reserve(s->scan_ticket, 1, s->stream));
    CUDA_TRY(cudaMemsetAsync(s->scan_ticket.p, 0, sizeof(unsigned long long), s->stream));
    s->scan_ticket_base = 0;
  }
  return PPCSR_OK;
}

// Exclusive scan of in(i), i in [0,n): out(i, exclusive_prefix, in(i)) is called once per element.
// The grand total goes to *d_total32 and/or *d_total64 (device pointers, nullable).
template <class In, class Out>
int device_scan(ppcsr_shard *s, In in, Out out, size_t n, uint32_t *d_total32, unsigned long long *d_total64) {
  if (n == 0) {
    if (d_total32 || d_total64) k_zero_totals<<<1, 1, 0, s->stream>>>(d_total32, d_total64);
    CUDA_TRY(cudaGetLastError());
    return PPCSR_OK;
  }
  const unsigned nb = div_up(n, SCAN_TILE);
  if (nb > SCAN_ONEPASS_MAX_TILES) {
    PPCSR_TRY(dev_reserve(s->block_tmp, (size_t)nb + 1, s->stream));
    s->launches += 3;
    k_scan_reduce<<<nb, SCAN_THREADS, 0, s->stream>>>(in, n, s->block_tmp.p);
    k_scan_spine<<<1, 1024, 0, s->stream>>>(s->block_tmp.p, nb, d_total32, d_total64);
    k_scan_apply<<<nb, SCAN_THREADS, 0, s->stream>>>(in, out, n, s->block_tmp.p);
    CUDA_TRY(cudaGetLastError());
    return PPCSR_OK;
  }
  PPCSR_TRY(reserve_scan_state(s, nb));
  s->scan_epoch++;
  s->launches += 1;
  k_scan_onepass<<<nb, SCAN_THREADS, 0, s->stream>>>(in, out, n, s->scan_ticket.p, s->scan_ticket_base, s->scan_state.p,
                                                     s->scan_epoch, nb, d_total32, d_total64);
  s->scan_ticket_base += nb;
  CUDA_TRY(cudaGetLastError());
  return PPCSR_OK;
}

// ---- common functors ----
struct InArray {
  const uint32_t *a;
  __device__ uint32_t operator()(size_t i) const { return a[i]; }
};
// writes the exclusive prefix to out[i] and the grand total to out[n]
struct OutPrefixWithTotal {
  uint32_t *out;
  size_t n;
  __device__ void operator()(size_t i, uint32_t ex, uint32_t own) const {
    out[i] = ex;
    if (i + 1 == n) out[n] = ex + own;
  }
};
struct OutNothing {
  __device__ void operator()(size_t, uint32_t, uint32_t) const {}
};
// list lengths that only the device knows: wrap a functor so elements at or beyond *n count as 0 / are skipped
// `skip` (nullable): a device flag that turns the whole pass into a no-op (e.g. the root is out of bounds,
// so no window list is needed).
template <class In>
struct BoundedIn {
  In f;
  const unsigned long long *n;
  const unsigned int *skip;
  __device__ uint32_t operator()(size_t i) const {
    if (skip && *skip) return 0u;
    return i < (size_t)*n ? f(i) : 0u;
  }
};
template <class Out>
struct BoundedOut {
  Out o;
  const unsigned long long *n;
  const unsigned int *skip;
  __device__ void operator()(size_t i, uint32_t ex, uint32_t own) const {
    if (skip && *skip) return;
    if (i < (size_t)*n) o(i, ex, own);
  }
};
template <class In>
BoundedIn<In> bounded_in(In f, const unsigned long long *n, const unsigned int *skip = nullptr) {
  return BoundedIn<In>{f, n, skip};
}
template <class Out>
BoundedOut<Out> bounded_out(Out o, const unsigned long long *n, const unsigned int *skip = nullptr) {
  return BoundedOut<Out>{o, n, skip};
}

// tile geometry of the owner-binning kernels (batch.cuh): 256 threads x 16 rounds, warp-contiguous sub-tiles
constexpr int SORT_THREADS = 256;
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int SORT_ROUNDS = 16;
constexpr int SORT_TILE = SORT_THREADS * SORT_ROUNDS;  // 4096 records per CTA

}  // namespace prim
