// primitives.cuh -- device-wide exclusive scan / stream compaction and the LSD radix sort.
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

// ---------------------------------------------------------------------------------------------
// LSD radix sort of (u64 key, u32 payload), stable, digits of up to 11 bits.
// The key (src << 32 | dst) is compacted on the fly to (src << lo_bits | dst) so that the passes cover
// one contiguous field of lo_bits + hi_bits bits: 40 bits (R-MAT scale 20) sort in 4 passes of 10 bits.
// Per pass: per-tile digit histogram -> exclusive scan in digit-major order -> stable scatter with
// warp-synchronous ranking (__match_any_sync); the tile is reordered in shared memory first so that the
// global stores are coalesced runs per digit.  Equal keys keep their submission order: the LAST element
// of a run of equal keys is the last op submitted for that (src,dst).  HAS_PAY = false sorts keys only
// (all payloads equal: the thread pools' value-1 inserts or pure delete batches).
// ---------------------------------------------------------------------------------------------
constexpr int SORT_THREADS = 256;
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int SORT_ROUNDS = 16;
constexpr int SORT_TILE = SORT_THREADS * SORT_ROUNDS;  // 4096 keys per CTA
constexpr int RADIX_BITS_MAX = 11;
constexpr int RADIX_MAX = 1 << RADIX_BITS_MAX;

__device__ __forceinline__ uint32_t sort_digit(uint64_t key, int lo_bits, int shift, uint32_t mask) {
  const uint64_t ck = lo_bits >= 32 ? key : (((key >> 32) << lo_bits) | (uint32_t)key);
  return (uint32_t)(ck >> shift) & mask;
}

__global__ void __launch_bounds__(SORT_THREADS) k_radix_hist(const uint64_t *__restrict__ keys, size_t n, int lo_bits,
                                                             int shift, uint32_t mask, uint32_t *__restrict__ hist,
                                                             uint32_t nblocks) {
  __shared__ uint32_t s_hist[RADIX_MAX];
  const uint32_t radix = mask + 1;
  for (uint32_t d = threadIdx.x; d < radix; d += SORT_THREADS) s_hist[d] = 0;
  __syncthreads();
  const size_t base = (size_t)blockIdx.x * SORT_TILE;
#pragma unroll 4
  for (int r = 0; r < SORT_ROUNDS; r++) {
    size_t i = base + (size_t)r * SORT_THREADS + threadIdx.x;
    if (i < n) atomicAdd(&s_hist[sort_digit(keys[i], lo_bits, shift, mask)], 1u);
  }
  __syncthreads();
  for (uint32_t d = threadIdx.x; d < radix; d += SORT_THREADS) hist[(size_t)d * nblocks + blockIdx.x] = s_hist[d];
}

// dynamic shared memory layout of k_radix_scatter (bytes):
//   keys[SORT_TILE] u64 | pay[SORT_TILE] u32 (HAS_PAY) | dbase[radix] u32 | goff[radix] u32 | cnt[8][radix] u16
inline size_t scatter_smem_bytes(uint32_t radix, bool has_pay) {
  return (size_t)SORT_TILE * 8 + (has_pay ? (size_t)SORT_TILE * 4 : 0) + (size_t)radix * 8 +
         (size_t)SORT_WARPS * radix * 2;
}

template <bool HAS_PAY>
__global__ void __launch_bounds__(SORT_THREADS, 2) k_radix_scatter(const uint64_t *__restrict__ keys,
                                                                   const uint32_t *__restrict__ pay, size_t n,
                                                                   int lo_bits, int shift, uint32_t mask,
                                                                   const uint32_t *__restrict__ offs, uint32_t nblocks,
                                                                   uint64_t *__restrict__ out_keys,
                                                                   uint32_t *__restrict__ out_pay) {
  extern __shared__ __align__(16) unsigned char s_dyn[];
  const uint32_t radix = mask + 1;
  uint64_t *s_keys = reinterpret_cast<uint64_t *>(s_dyn);
  uint32_t *s_pay = reinterpret_cast<uint32_t *>(s_dyn + (size_t)SORT_TILE * 8);
  uint32_t *s_dbase = reinterpret_cast<uint32_t *>(s_dyn + (size_t)SORT_TILE * 8 + (HAS_PAY ? (size_t)SORT_TILE * 4 : 0));
  uint32_t *s_goff = s_dbase + radix;
  uint16_t *s_cnt = reinterpret_cast<uint16_t *>(s_goff + radix);  // [SORT_WARPS][radix]
  __shared__ uint32_t s_warp[33];
  for (uint32_t d = threadIdx.x; d < SORT_WARPS * radix / 2; d += SORT_THREADS) reinterpret_cast<uint32_t *>(s_cnt)[d] = 0;
  __syncthreads();
  const unsigned w = threadIdx.x >> 5, l = lane_id();
  const unsigned lt = lanemask_lt();
  uint16_t *my_cnt = s_cnt + (size_t)w * radix;
  // warp w owns the contiguous sub-tile [w*512, (w+1)*512): round r covers 32 consecutive keys
  const size_t tile0 = (size_t)blockIdx.x * SORT_TILE;
  const size_t wbase = tile0 + (size_t)w * (32 * SORT_ROUNDS);
  uint64_t k[SORT_ROUNDS];
  uint16_t rank[SORT_ROUNDS];
#pragma unroll
  for (int r = 0; r < SORT_ROUNDS; r++) {
    size_t i = wbase + (size_t)r * 32 + l;
    k[r] = (i < n) ? keys[i] : 0;
  }
#pragma unroll
  for (int r = 0; r < SORT_ROUNDS; r++) {
    size_t i = wbase + (size_t)r * 32 + l;
    const bool valid = i < n;
    const uint32_t d = valid ? sort_digit(k[r], lo_bits, shift, mask) : 0xFFFFu;
    const unsigned peers = __match_any_sync(0xFFFFFFFFu, d);
    uint32_t base = 0;
    if (valid) base = my_cnt[d];
    __syncwarp();
    if (valid && (peers & lt) == 0) my_cnt[d] = (uint16_t)(base + __popc(peers));
    __syncwarp();
    rank[r] = (uint16_t)(base + __popc(peers & lt));
  }
  __syncthreads();
  // per digit: exclusive prefix over warps; block-wide exclusive prefix over digits (radix/256 digits per thread)
  {
    const uint32_t per = radix / SORT_THREADS ? radix / SORT_THREADS : 1;
    const uint32_t d0 = threadIdx.x * per;
    uint32_t tot[RADIX_MAX / SORT_THREADS];
    uint32_t sum = 0;
#pragma unroll
    for (uint32_t x = 0; x < RADIX_MAX / SORT_THREADS; x++) {
      tot[x] = 0;
      const uint32_t d = d0 + x;
      if (x < per && d < radix) {
        uint32_t run = 0;
#pragma unroll
        for (int ww = 0; ww < SORT_WARPS; ww++) {
          const uint32_t t = s_cnt[(size_t)ww * radix + d];
          s_cnt[(size_t)ww * radix + d] = (uint16_t)run;
          run += t;
        }
        tot[x] = run;
        sum += run;
      }
    }
    uint32_t total;
    uint32_t dbase = block_excl_scan(sum, &total, s_warp);
#pragma unroll
    for (uint32_t x = 0; x < RADIX_MAX / SORT_THREADS; x++) {
      const uint32_t d = d0 + x;
      if (x < per && d < radix) {
        s_dbase[d] = dbase;
        s_goff[d] = offs[(size_t)d * nblocks + blockIdx.x] - dbase;
        dbase += tot[x];
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < SORT_ROUNDS; r++) {
    size_t i = wbase + (size_t)r * 32 + l;
    if (i < n) {
      const uint32_t d = sort_digit(k[r], lo_bits, shift, mask);
      const uint32_t pos = s_dbase[d] + my_cnt[d] + rank[r];
      s_keys[pos] = k[r];
      if (HAS_PAY) s_pay[pos] = pay[i];
    }
  }
  __syncthreads();
  const uint32_t tile_n = (uint32_t)min((size_t)SORT_TILE, n - tile0);
  for (uint32_t j = threadIdx.x; j < tile_n; j += SORT_THREADS) {
    const uint64_t key = s_keys[j];
    const uint32_t d = sort_digit(key, lo_bits, shift, mask);
    const uint32_t pos = s_goff[d] + j;
    out_keys[pos] = key;
    if (HAS_PAY) out_pay[pos] = s_pay[j];
  }
}

// Sorts n (key,payload) pairs by the key bits [0,lo_bits) and [32, 32+hi_bits).  Input in (ka,pa); the
// sorted result ends up in *rk / *rp which point at either buffer pair.  pa == nullptr sorts keys only.
inline int radix_sort_pairs(ppcsr_shard *s, uint64_t *ka, uint32_t *pa, uint64_t *kb, uint32_t *pb, size_t n,
                            int lo_bits, int hi_bits, uint64_t **rk, uint32_t **rp) {
  *rk = ka;
  *rp = pa;
  if (n <= 1) return PPCSR_OK;
  const bool has_pay = pa != nullptr;
  const unsigned nblocks = div_up(n, SORT_TILE);
  const int total = lo_bits + hi_bits;
  const int passes = (total + RADIX_BITS_MAX - 1) / RADIX_BITS_MAX;
  const int width = (total + passes - 1) / passes;
  PPCSR_TRY(dev_reserve(s->hist, ((size_t)1 << width) * nblocks + 1, s->stream));
  // > 48 KB of dynamic shared memory needs an explicit opt-in (per device, cheap)
  CUDA_TRY(cudaFuncSetAttribute(k_radix_scatter<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)scatter_smem_bytes(RADIX_MAX, true)));
  CUDA_TRY(cudaFuncSetAttribute(k_radix_scatter<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)scatter_smem_bytes(RADIX_MAX, false)));
  uint64_t *src_k = ka, *dst_k = kb;
  uint32_t *src_p = pa, *dst_p = pb;
  for (int done = 0; done < total; done += width) {
    const int w = (total - done) < width ? (total - done) : width;
    const uint32_t mask = (1u << w) - 1u;
    const uint32_t radix = mask + 1;
    const size_t hn = (size_t)radix * nblocks;
    s->launches += 2;
    k_radix_hist<<<nblocks, SORT_THREADS, 0, s->stream>>>(src_k, n, lo_bits, done, mask, s->hist.p, nblocks);
    PPCSR_TRY(device_scan(s, InArray{s->hist.p}, OutPrefixWithTotal{s->hist.p, hn}, hn, nullptr, nullptr));
    if (has_pay) {
      k_radix_scatter<true><<<nblocks, SORT_THREADS, scatter_smem_bytes(radix, true), s->stream>>>(
          src_k, src_p, n, lo_bits, done, mask, s->hist.p, nblocks, dst_k, dst_p);
    } else {
      k_radix_scatter<false><<<nblocks, SORT_THREADS, scatter_smem_bytes(radix, false), s->stream>>>(
          src_k, nullptr, n, lo_bits, done, mask, s->hist.p, nblocks, dst_k, nullptr);
    }
    CUDA_TRY(cudaGetLastError());
    uint64_t *tk = src_k;
    src_k = dst_k;
    dst_k = tk;
    uint32_t *tp = src_p;
    src_p = dst_p;
    dst_p = tp;
  }
  *rk = src_k;
  *rp = has_pay ? src_p : nullptr;
  return PPCSR_OK;
}

}  // namespace prim
