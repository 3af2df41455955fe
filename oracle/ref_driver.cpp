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
 * Input files: raw little-endian u32 triples (src, dst, value); value==0 means delete.
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
#include <string>
#include <vector>

#include "thread_pool/thread_pool.h"
#include "thread_pool_pppcsr/thread_pool_pppcsr.h"

struct Op {
  uint32_t src, dst, val;
};

static std::vector<Op> read_ops(const std::string &path) {
  std::vector<Op> v;
  if (path.empty()) return v;
  FILE *f = fopen(path.c_str(), "rb");
  if (!f) {
    fprintf(stderr, "ref_driver: cannot open %s\n", path.c_str());
    exit(2);
  }
  fseek(f, 0, SEEK_END);
  long sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  v.resize(sz / sizeof(Op));
  if (!v.empty() && fread(v.data(), sizeof(Op), v.size(), f) != v.size()) {
    fprintf(stderr, "ref_driver: short read %s\n", path.c_str());
    exit(2);
  }
  fclose(f);
  return v;
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

template <typename Pool>
static double run_pool_phase(Pool &pool, const std::vector<Op> &ops, size_t count, int threads) {
  for (size_t i = 0; i < count; i++) {
    if (ops[i].val != 0) {
      pool.submit_add((int)(i % threads), (int)ops[i].src, (int)ops[i].dst);
    } else {
      pool.submit_delete((int)(i % threads), (int)ops[i].src, (int)ops[i].dst);
    }
  }
  auto t0 = std::chrono::steady_clock::now();
  pool.start(threads);
  pool.stop();
  auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double, std::milli>(t1 - t0).count();
}

template <typename G>
static void run_direct(G &g, const std::vector<Op> &ops, size_t count) {
  for (size_t i = 0; i < count; i++) {
    if (ops[i].val != 0) {
      g.add_edge(ops[i].src, ops[i].dst, ops[i].val);
    } else {
      g.remove_edge(ops[i].src, ops[i].dst);
    }
  }
}

int main(int argc, char **argv) {
  std::string mode = "ppcsr", api = "pool", core_path, upd_path, dump_path, timing_path;
  int threads = 1, ppd = 1, add_nodes = 0;
  long size = -1;
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
    else if (a == "--lock-free") lock_search = false;
    else if (a == "--pagerank") with_pr = true;
    else if (a == "--add-nodes") add_nodes = atoi(next().c_str());
    else {
      fprintf(stderr, "ref_driver: unknown arg %s\n", a.c_str());
      return 2;
    }
  }
  std::vector<Op> core = read_ops(core_path);
  std::vector<Op> upd = read_ops(upd_path);
  size_t upd_count = (size < 0 || (size_t)size > upd.size()) ? upd.size() : (size_t)size;
  double core_ms = 0, upd_ms = 0;

  if (api == "pool") {
    if (mode == "ppcsr") {
      ThreadPool pool(threads, lock_search, n, ppd);
      core_ms = run_pool_phase(pool, core, core.size(), threads);
      upd_ms = run_pool_phase(pool, upd, upd_count, threads);
      if (!dump_path.empty()) dump_graph(*pool.pcsr, dump_path, with_pr);
    } else {
      ThreadPoolPPPCSR pool(threads, lock_search, n, ppd, mode == "pppcsrnuma");
      core_ms = run_pool_phase(pool, core, core.size(), threads);
      upd_ms = run_pool_phase(pool, upd, upd_count, threads);
      if (!dump_path.empty()) dump_graph(*pool.pcsr, dump_path, with_pr);
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
