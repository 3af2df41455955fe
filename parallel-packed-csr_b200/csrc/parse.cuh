// parse.cuh -- the input path on the GPU: the reference's text edge-list reader (src/main.cpp:29-62, read_input:
// getline + stoi + substr per line, one thread) as two kernels over the raw bytes of the file.
//   line starts   one scan-compaction over the bytes: a line starts at byte 0 and after every '\n'
//   k_parse_lines one thread per line, the reference's rules restated:
//       src    = stoi(line, &pos)                  leading white space, optional sign, decimal digits
//       target = stoi(line.substr(pos + 1), &pos2) exactly ONE separator character is skipped, then as above
//       op     = line[pos + 1 + pos2 + 1] if that index lies inside the line: '1' add, '0' delete, else the default
//   A line without a parsable pair is dropped (the reference's stoi would throw): it is emitted as (0xFFFFFFFF,
//   0xFFFFFFFF), which the batch guards reject (src >= n), and does not count towards the largest vertex id.
// Values: add = 1 (the thread pools always insert value 1, reference thread_pool.cpp:44), delete = 0.
#pragma once
#include "common.cuh"
#include "primitives.cuh"

namespace parse {

constexpr int PT = 256;

struct InIsNewline {
  const char *text;
  __device__ uint32_t operator()(size_t i) const { return text[i] == '\n' ? 1u : 0u; }
};
// line k + 1 starts right after the k-th newline; line 0 starts at byte 0 (written by the host)
struct OutLineStart {
  unsigned long long *starts;
  __device__ void operator()(size_t i, uint32_t ex, uint32_t own) const {
    if (own) starts[ex + 1] = (unsigned long long)i + 1ull;
  }
};

__device__ __forceinline__ bool is_space(char c) { return c == ' ' || (c >= '\t' && c <= '\r'); }

// stoi on [p, end): returns false if no digits; *next = first byte after the number
__device__ __forceinline__ bool parse_int(const char *p, const char *end, long long *out, const char **next) {
  while (p < end && is_space(*p)) p++;
  bool neg = false;
  if (p < end && (*p == '+' || *p == '-')) {
    neg = *p == '-';
    p++;
  }
  if (p >= end || *p < '0' || *p > '9') return false;
  long long v = 0;
  while (p < end && *p >= '0' && *p <= '9') {
    v = v * 10 + (*p - '0');
    if (v > 0xFFFFFFFFll) v = 0xFFFFFFFFll;  // saturate: such an id is out of range anyway
    p++;
  }
  *out = neg ? -v : v;
  *next = p;
  return true;
}

__global__ void __launch_bounds__(PT) k_parse_lines(const char *__restrict__ text, unsigned long long bytes,
                                                    const unsigned long long *__restrict__ starts,
                                                    unsigned long long n_lines, uint32_t default_val,
                                                    uint32_t *__restrict__ src, uint32_t *__restrict__ dst,
                                                    uint32_t *__restrict__ val, unsigned int *max_id,
                                                    unsigned long long *n_valid) {
  const unsigned long long k = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t s = PPCSR_SENT, d = PPCSR_SENT, v = default_val;
  bool ok = false;
  if (k < n_lines) {
    const char *p = text + starts[k];
    // the line without its newline; the last line ends at the end of the file
    const char *end = text + (k + 1 < n_lines ? starts[k + 1] - 1ull : bytes);
    if (k + 1 == n_lines && end > p && end[-1] == '\n') end--;
    long long a = 0, b = 0;
    const char *q = nullptr, *r = nullptr;
    if (parse_int(p, end, &a, &q) && q < end && parse_int(q + 1, end, &b, &r)) {
      ok = true;
      s = (uint32_t)a;  // a negative id wraps to a huge one: rejected by the guards, like any id >= n
      d = (uint32_t)b;
      const char *opc = r + 1;  // == line + pos + 1 + pos2 + 1
      if (opc < end) {
        if (*opc == '1') v = 1u;
        else if (*opc == '0') v = 0u;
      }
    }
    src[k] = s;
    dst[k] = d;
    val[k] = v;
  }
  // largest id over the parsed lines (reference main.cpp:44: max(src, target), ints)
  unsigned int m = ok ? max(s, d) : 0u;
  if (ok && ((int)s < 0 || (int)d < 0)) m = 0u;  // negative ints never raise the reference's maximum
  m = __reduce_max_sync(0xFFFFFFFFu, m);
  const unsigned okm = __ballot_sync(0xFFFFFFFFu, ok);
  if (lane_id() == 0) {
    if (m) atomicMax(max_id, m);
    if (okm) atomicAdd(n_valid, (unsigned long long)__popc(okm));
  }
}

}  // namespace parse
