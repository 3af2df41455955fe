/*
 * TEST INFRASTRUCTURE ONLY.  Driver that links the UNMODIFIED reference sources
 * (compiled where they lie under /root/reference by oracle/Makefile -> oracle/_ref/)
 * and exposes what the parity tests and the CPU baseline need:
 *
 *   - apply a core graph and an update stream through the reference's own public API
 *       --api pool   : ThreadPool / ThreadPoolPPPCSR submit_* + start() + stop()
 *                      (reference src/thread_pool/thread_pool.h:17-29,
 *                       src/thread_pool_pppcsr/thread_pool_pppcsr.h:17-29), value is always 1
 *       --api direct : sequential PCSR/PPPCSR add_edge/remove_edge on one thread
 *                      (reference src/pcsr/PCSR.h:73-78), values preserved
 *   - dump the logical graph: per-vertex get_neighbourhood() + getNode().num_neighbors
 *   - one reference pagerank<T,double>() push step (reference src/utility/pagerank.h:16-29)
 *   - time start()->stop() exactly as the reference does (src/thread_pool/thread_pool.cpp:79,110)
 *
 * Nothing here is product code and nothing in the product links it.
 *
 * Input files: raw little-endian u32 triples (src, dst, value); value==0 means delete.  Instead of a file a
 * stream can be SYNTHESISED in place (--synth-core / --synth-updates kind:scale:lo:hi:seed, kind = rmat | uniform):
 * the same counter-hash streams as parallel-packed-csr_b200/synth.py (bit-identical, pinned by
 * tests/test_oracle.py), generated chunk by chunk so that a scale-24 run needs no multi-gigabyte input files.
 * --checksum <path> writes an order-independent 64-bit checksum of the logical graph (see graph_checksum below):
 * the full-size parity anchor of bench.py (tests/golden/c4_checksum.json).
 * Dump file  : u64 magic, u64 n, u64 E, u64 N, u64 logN, u64 H (geometry of partition 0 for PPPCSR),
 *              u64 rowptr[n+1], u32 col[E], u32 num_neighbors[n], u64 has_pagerank, double pr[n].
 */
#include <pagerank.h>

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <memory>
#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include "thread_pool/thread_pool.h"
#include "thread_pool_pppcsr/thread_pool_pppcsr.h"

struct Op {
  uint32_t src, dst, val;
};

// ---- synthetic streams: restatement of parallel-packed-csr_b200/synth.py (murmur3 fmix32 counter hash) ----
static inline uint32_t mix32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x85EBCA6Bu;
  x ^= x >> 13;
  x *= 0xC2B2AE35u;
  x ^= x >> 16;
  return x;
}
static inline uint32_t base_hash(uint64_t idx, uint32_t seed) {
  const uint32_t hi = mix32((uint32_t)(idx >> 32) + seed * 0x9E3779B1u + 0x7F4A7C15u);
  return mix32((uint32_t)idx ^ hi);
}
// thresholds int(.57 * 2^32), int((.57+.19) * 2^32), int((.57+.19+.19) * 2^32) as synth.py computes them
static const uint32_t T_A = 2448131358u, T_AB = 3264175144u, T_ABC = 4080218931u;
static inline Op rmat_op(int scale, uint64_t idx, uint32_t seed) {
  const uint32_t h0 = base_hash(idx, seed);
  uint32_t s = 0, d = 0;
  for (int level = 0; level < scale; level++) {
    const uint32_t r = mix32(h0 + (uint32_t)(level + 1) * 0x9E3779B9u);
    s = (s << 1) | (r >= T_AB ? 1u : 0u);
    d = (d << 1) | (((r >= T_A && r < T_AB) || r >= T_ABC) ? 1u : 0u);
  }
  return Op{s, d, 1u};
}
static inline Op uniform_op(int scale, uint64_t idx, uint32_t seed) {
  const uint32_t h0 = base_hash(idx, seed);
  const uint32_t mask = scale >= 32 ? 0xFFFFFFFFu : ((1u << scale) - 1u);
  return Op{mix32(h0 + 0x68E31DA4u) & mask, mix32(h0 + 0xB5297A4Du) & mask, 1u};
}

// A stream of ops: a file of triples or a synthetic stream; read in chunks.
struct Stream {
  enum Kind { NONE, FILE_, RMAT, UNIFORM } kind = NONE;
  std::vector<Op> file_ops;
  int scale = 0;
  uint64_t lo = 0, hi = 0;
  uint32_t seed = 0;
  size_t size() const { return kind == FILE_ ? file_ops.size() : (size_t)(hi - lo); }
  // ops [a, b) of the stream into out (multi-threaded for synthetic kinds)
  void fill(size_t a, size_t b, std::vector<Op> &out, int threads) const {
    out.resize(b - a);
    if (kind == FILE_) {
      std::copy(file_ops.begin() + a, file_ops.begin() + b, out.begin());
      return;
    }
    const int T = std::max(1, threads);
    std::vector<std::thread> ws;
    for (int t = 0; t < T; t++) {
      ws.emplace_back([&, t]() {
        const size_t x0 = a + (b - a) * t / T, x1 = a + (b - a) * (t + 1) / T;
        for (size_t x = x0; x < x1; x++)
          out[x - a] = kind == RMAT ? rmat_op(scale, lo + x, seed) : uniform_op(scale, lo + x, seed);
      });
    }
    for (auto &w : ws) w.join();
  }
};

static Stream file_stream(const std::string &path) {
  Stream st;
  if (path.empty()) return st;
  st.kind = Stream::FILE_;
  FILE *f = fopen(path.c_str(), "rb");
  if (!f) {
    fprintf(stderr, "ref_driver: cannot open %s\n", path.c_str());
    exit(2);
  }
  fseek(f, 0, SEEK_END);
  long sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  st.file_ops.resize(sz / sizeof(Op));
  if (!st.file_ops.empty() && fread(st.file_ops.data(), sizeof(Op), st.file_ops.size(), f) != st.file_ops.size()) {
    fprintf(stderr, "ref_driver: short read %s\n", path.c_str());
    exit(2);
  }
  fclose(f);
  return st;
}

// kind:scale:lo:hi:seed
static Stream synth_stream(const std::string &spec) {
  Stream st;
  char kind[16] = {0};
  unsigned long long lo = 0, hi = 0;
  unsigned seed = 0;
  int scale = 0;
  if (sscanf(spec.c_str(), "%15[a-z]:%d:%llu:%llu:%u", kind, &scale, &lo, &hi, &seed) != 5 || hi < lo) {
    fprintf(stderr, "ref_driver: bad synthetic stream spec %s\n", spec.c_str());
    exit(2);
  }
  st.kind = std::string(kind) == "rmat" ? Stream::RMAT : Stream::UNIFORM;
  st.scale = scale;
  st.lo = lo;
  st.hi = hi;
  st.seed = seed;
  return st;
}

// Order-independent checksum of the logical graph: number of edges, and the sums (mod 2^64) of
// mix64(src << 32 | dst) over all edges and of num_neighbors[v] * mix64(v) over all vertices (splitmix64 finaliser).
static inline uint64_t mix64(uint64_t x) {
  x ^= x >> 30;
  x *= 0xBF58476D1CE4E5B9ull;
  x ^= x >> 27;
  x *= 0x94D049BB133111EBull;
  x ^= x >> 31;
  return x;
}
template <typename G>
static void graph_checksum(G &g, const std::string &path) {
  const uint64_t n = g.get_n();
  uint64_t edges = 0, edge_hash = 0, nn_hash = 0;
  for (uint64_t v = 0; v < n; v++) {
    auto nb = g.get_neighbourhood((int)v);
    edges += nb.size();
    for (int d : nb) edge_hash += mix64((v << 32) | (uint32_t)d);
    nn_hash += (uint64_t)g.getNode((int)v).num_neighbors * mix64(v);
  }
  FILE *f = fopen(path.c_str(), "w");
  if (!f) {
    fprintf(stderr, "ref_driver: cannot write %s\n", path.c_str());
    exit(2);
  }
  fprintf(f, "{\"n\": %llu, \"edges\": %llu, \"edge_hash\": \"%016llx\", \"nn_hash\": \"%016llx\"}\n",
          (unsigned long long)n, (unsigned long long)edges, (unsigned long long)edge_hash,
          (unsigned long long)nn_hash);
  fclose(f);
}

static void geometry_of(PCSR &g, uint64_t geo[3]) {
  geo[0] = g.edges.N;
  geo[1] = (uint64_t)g.edges.logN;
  geo[2] = (uint64_t)g.edges.H;
}
static void geometry_of(PPPCSR &g, uint64_t geo[3]) {
  (void)g;  // partitions are private in the reference (src/pppcsr/PPPCSR.h:52-59)
  geo[0] = geo[1] = geo[2] = 0;
}

template <typename G>
static void dump_graph(G &g, const std::string &path, bool with_pr) {
  FILE *f = fopen(path.c_str(), "wb");
  if (!f) {
    fprintf(stderr, "ref_driver: cannot write %s\n", path.c_str());
    exit(2);
  }
  const uint64_t n = g.get_n();
  std::vector<uint64_t> rowptr(n + 1, 0);
  std::vector<uint32_t> col;
  std::vector<uint32_t> nn(n);
  for (uint64_t v = 0; v < n; v++) {
    auto nb = g.get_neighbourhood((int)v);
    for (int d : nb) col.push_back((uint32_t)d);
    rowptr[v + 1] = col.size();
    nn[v] = g.getNode((int)v).num_neighbors;
  }
  const uint64_t magic = 0x50504353524F5243ull;  // "PPCSRORC"
  const uint64_t E = col.size();
  fwrite(&magic, 8, 1, f);
  fwrite(&n, 8, 1, f);
  fwrite(&E, 8, 1, f);
  uint64_t geo[3];
  geometry_of(g, geo);
  fwrite(geo, 8, 3, f);
  fwrite(rowptr.data(), 8, n + 1, f);
  if (E) fwrite(col.data(), 4, E, f);
  if (n) fwrite(nn.data(), 4, n, f);
  uint64_t has_pr = with_pr ? 1 : 0;
  fwrite(&has_pr, 8, 1, f);
  if (with_pr) {
    std::vector<double> vals(n);
    for (uint64_t i = 0; i < n; i++) vals[i] = 1.0 + (double)(i % 7);
    std::vector<double> pr = pagerank<G, double>(g, vals);
    if (n) fwrite(pr.data(), 8, n, f);
  }
  fclose(f);
}

static const size_t CHUNK_OPS = 1u << 22;

template <typename Pool>
static double run_pool_phase(Pool &pool, const Stream &st, size_t count, int threads, size_t first = 0) {
  std::vector<Op> ops;
  for (size_t base = first; base < count; base += CHUNK_OPS) {
    const size_t top = std::min(count, base + CHUNK_OPS);
    st.fill(base, top, ops, threads);
    for (size_t i = base; i < top; i++) {
      const Op &o = ops[i - base];
      if (o.val != 0) {
        pool.submit_add((int)(i % threads), (int)o.src, (int)o.dst);
      } else {
        pool.submit_delete((int)(i % threads), (int)o.src, (int)o.dst);
      }
    }
  }
  auto t0 = std::chrono::steady_clock::now();
  pool.start(threads);
  pool.stop();
  auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double, std::milli>(t1 - t0).count();
}

template <typename G>
static void run_direct(G &g, const Stream &st, size_t count) {
  std::vector<Op> ops;
  for (size_t base = 0; base < count; base += CHUNK_OPS) {
    const size_t top = std::min(count, base + CHUNK_OPS);
    st.fill(base, top, ops, 1);
    for (const Op &o : ops) {
      if (o.val != 0) {
        g.add_edge(o.src, o.dst, o.val);
      } else {
        g.remove_edge(o.src, o.dst);
      }
    }
  }
}

int main(int argc, char **argv) {
  std::string mode = "ppcsr", api = "pool", core_path, upd_path, dump_path, timing_path, sum_path, core_synth, upd_synth,
      emit_path;
  int threads = 1, ppd = 1, add_nodes = 0;
  long size = -1, ckpt_at = 0;
  std::string ckpt_path;
  uint32_t n = 0;
  bool lock_search = true, with_pr = false;
  for (int i = 1; i < argc; i++) {
    std::string a = argv[i];
    auto next = [&]() -> std::string {
      if (i + 1 >= argc) {
        fprintf(stderr, "ref_driver: missing value for %s\n", a.c_str());
        exit(2);
      }
      return std::string(argv[++i]);
    };
    if (a == "--mode") mode = next();
    else if (a == "--api") api = next();
    else if (a == "--threads") threads = atoi(next().c_str());
    else if (a == "--ppd") ppd = atoi(next().c_str());
    else if (a == "--n") n = (uint32_t)strtoul(next().c_str(), nullptr, 10);
    else if (a == "--core") core_path = next();
    else if (a == "--updates") upd_path = next();
    else if (a == "--size") size = atol(next().c_str());
    else if (a == "--dump") dump_path = next();
    else if (a == "--timing") timing_path = next();
    else if (a == "--checksum") sum_path = next();
    else if (a == "--checkpoint") {  // --checkpoint <updates> <path>: extra dump after the first <updates> (pool api)
      ckpt_at = atol(next().c_str());
      ckpt_path = next();
    }
    else if (a == "--synth-core") core_synth = next();
    else if (a == "--synth-updates") upd_synth = next();
    else if (a == "--emit-updates") emit_path = next();
    else if (a == "--lock-free") lock_search = false;
    else if (a == "--pagerank") with_pr = true;
    else if (a == "--add-nodes") add_nodes = atoi(next().c_str());
    else {
      fprintf(stderr, "ref_driver: unknown arg %s\n", a.c_str());
      return 2;
    }
  }
  const Stream core = core_synth.empty() ? file_stream(core_path) : synth_stream(core_synth);
  const Stream upd = upd_synth.empty() ? file_stream(upd_path) : synth_stream(upd_synth);
  size_t upd_count = (size < 0 || (size_t)size > upd.size()) ? upd.size() : (size_t)size;
  if (!emit_path.empty()) {  // write the update stream as triples and leave (generator parity test)
    std::vector<Op> ops;
    upd.fill(0, upd_count, ops, threads);
    FILE *f = fopen(emit_path.c_str(), "wb");
    if (!f || (upd_count && fwrite(ops.data(), sizeof(Op), upd_count, f) != upd_count)) {
      fprintf(stderr, "ref_driver: cannot write %s\n", emit_path.c_str());
      return 2;
    }
    fclose(f);
    return 0;
  }
  double core_ms = 0, upd_ms = 0;

  if (api == "pool") {
    if (mode == "ppcsr") {
      ThreadPool pool(threads, lock_search, n, ppd);
      core_ms = run_pool_phase(pool, core, core.size(), threads);
      size_t done = 0;
      if (ckpt_at > 0 && (size_t)ckpt_at < upd_count) {  // dump the graph after the first ckpt_at updates, then go on
        upd_ms += run_pool_phase(pool, upd, (size_t)ckpt_at, threads);
        dump_graph(*pool.pcsr, ckpt_path, false);
        done = (size_t)ckpt_at;
      }
      upd_ms += run_pool_phase(pool, upd, upd_count, threads, done);
      if (!dump_path.empty()) dump_graph(*pool.pcsr, dump_path, with_pr);
      if (!sum_path.empty()) graph_checksum(*pool.pcsr, sum_path);
    } else {
      ThreadPoolPPPCSR pool(threads, lock_search, n, ppd, mode == "pppcsrnuma");
      core_ms = run_pool_phase(pool, core, core.size(), threads);
      size_t done = 0;
      if (ckpt_at > 0 && (size_t)ckpt_at < upd_count) {  // dump the graph after the first ckpt_at updates, then go on
        upd_ms += run_pool_phase(pool, upd, (size_t)ckpt_at, threads);
        dump_graph(*pool.pcsr, ckpt_path, false);
        done = (size_t)ckpt_at;
      }
      upd_ms += run_pool_phase(pool, upd, upd_count, threads, done);
      if (!dump_path.empty()) dump_graph(*pool.pcsr, dump_path, with_pr);
      if (!sum_path.empty()) graph_checksum(*pool.pcsr, sum_path);
    }
  } else {
    if (mode == "ppcsr") {
      PCSR g(n, n, lock_search, -1);
      for (int k = 0; k < add_nodes; k++) g.add_node();
      auto t0 = std::chrono::steady_clock::now();
      run_direct(g, core, core.size());
      auto t1 = std::chrono::steady_clock::now();
      run_direct(g, upd, upd_count);
      auto t2 = std::chrono::steady_clock::now();
      core_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
      upd_ms = std::chrono::duration<double, std::milli>(t2 - t1).count();
      if (!dump_path.empty()) dump_graph(g, dump_path, with_pr);
      if (!sum_path.empty()) graph_checksum(g, sum_path);
    } else {
      PPPCSR g(n, n, lock_search, 1, ppd, false);
      for (int k = 0; k < add_nodes; k++) g.add_node();
      auto t0 = std::chrono::steady_clock::now();
      run_direct(g, core, core.size());
      auto t1 = std::chrono::steady_clock::now();
      run_direct(g, upd, upd_count);
      auto t2 = std::chrono::steady_clock::now();
      core_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
      upd_ms = std::chrono::duration<double, std::milli>(t2 - t1).count();
      if (!dump_path.empty()) dump_graph(g, dump_path, with_pr);
      if (!sum_path.empty()) graph_checksum(g, sum_path);
    }
  }
  if (!timing_path.empty()) {
    FILE *f = fopen(timing_path.c_str(), "w");
    if (f) {
      fprintf(f, "{\"core_ms\": %.3f, \"update_ms\": %.3f, \"core_ops\": %zu, \"update_ops\": %zu, \"threads\": %d}\n",
              core_ms, upd_ms, core.size(), upd_count, threads);
      fclose(f);
    }
  }
  return 0;
}
