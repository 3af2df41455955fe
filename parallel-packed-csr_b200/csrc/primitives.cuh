// primitives.cuh -- device-wide exclusive scan / stream compaction (the radix sort lives in sort.cuh).
// Hand-written (no CUB/Thrust): three-phase scan (tile reduce -> spine -> tile scan) with functor
// input/output so the same code serves prefix sums, order-preserving selects and histogram offsets.
#pragma once
#include "common.cuh"

namespace prim {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

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
  PPCSR_TRY(dev_reserve(s->block_tmp, (size_t)nb + 1, s->stream));
  s->launches += 3;
  k_scan_reduce<<<nb, SCAN_THREADS, 0, s->stream>>>(in, n, s->block_tmp.p);
  k_scan_spine<<<1, 1024, 0, s->stream>>>(s->block_tmp.p, nb, d_total32, d_total64);
  k_scan_apply<<<nb, SCAN_THREADS, 0, s->stream>>>(in, out, n, s->block_tmp.p);
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
