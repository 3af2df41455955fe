// parallel-packed-csr (B200 build) -- command line driver with the reference's flags and stdout contract
// (reference src/main.cpp:111-189):
//   -threads= -size= -lock_free -insert -delete -pppcsrnuma -pppcsr -ppcsr -partitions_per_domain=
//   -core_graph= -update_file=
// Flags are matched by prefix in the same order as the reference, so the same ordering rules hold:
// -insert/-delete and -size= must come before -update_file=.  Input: one edge per line,
// `src<sep>dst[<sep>{1|0}]`, 1 = add, 0 = delete, absent = the default op (reference main.cpp:29-62).
// The core graph is loaded through the same insert path, then the first `size` updates are applied; each
// phase prints "Elapsed wall clock time: <ms>" -- benchmark scripts keep the second line
// (reference src/benchmarking/benchmark-strong-scaling.sh:116).
// Extra: -gpus=<k> limits the GPUs used as partitions' domains; -check verifies the PMA invariants.
#include <algorithm>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <memory>
#include <string>
#include <tuple>
#include <vector>

#include "thread_pool.h"
#include "thread_pool_pppcsr.h"

enum class Operation { READ, ADD, DELETE };
using EdgeOp = std::tuple<Operation, int, int>;

static bool starts_with(const std::string &s, const char *prefix) { return s.rfind(prefix, 0) == 0; }

// Parses an edge list; returns the ops and the largest vertex id seen.
static std::pair<std::vector<EdgeOp>, int> read_input(const std::string &filename, Operation default_op) {
  std::ifstream f(filename);
  if (!f.good()) {
    std::cerr << "Invalid file" << std::endl;
    std::exit(EXIT_FAILURE);
  }
  std::vector<EdgeOp> ops;
  int max_id = 0;
  std::string line;
  while (std::getline(f, line)) {
    if (line.empty()) continue;
    const char *p = line.c_str();
    char *end = nullptr;
    const long s = std::strtol(p, &end, 10);
    if (end == p) continue;
    const char *q = (*end != '\0') ? end + 1 : end;  // exactly one separator character
    char *end2 = nullptr;
    const long d = std::strtol(q, &end2, 10);
    Operation op = default_op;
    if (end2 != q && *end2 != '\0' && *(end2 + 1) != '\0') {
      const char c = *(end2 + 1);
      if (c == '1') op = Operation::ADD;
      else if (c == '0') op = Operation::DELETE;
      else std::cerr << "Invalid operation";
    }
    max_id = std::max(max_id, (int)std::max(s, d));
    ops.emplace_back(op, (int)s, (int)d);
  }
  return {std::move(ops), max_id};
}

template <typename Pool>
static void run_phase(const std::vector<EdgeOp> &ops, Pool *pool, int threads, int count) {
  for (int i = 0; i < count; i++) {
    switch (std::get<0>(ops[i])) {
      case Operation::ADD:
        pool->submit_add(i % threads, std::get<1>(ops[i]), std::get<2>(ops[i]));
        break;
      case Operation::DELETE:
        pool->submit_delete(i % threads, std::get<1>(ops[i]), std::get<2>(ops[i]));
        break;
      case Operation::READ:
        std::cerr << "Not implemented\n";
        break;
    }
  }
  pool->start(threads);
  pool->stop();
}

template <typename Pool>
static void execute(int threads, int size, const std::vector<EdgeOp> &core, const std::vector<EdgeOp> &updates,
                    std::unique_ptr<Pool> &pool) {
  run_phase(core, pool.get(), threads, (int)core.size());
  run_phase(updates, pool.get(), threads, size);
}

enum class Version { PPCSR, PPPCSR, PPPCSRNUMA };

int main(int argc, char *argv[]) {
  int threads = 8, size = 1000000, num_nodes = 0, partitions_per_domain = 1;
  bool lock_search = true, insert = true, check = false;
  Version v = Version::PPPCSRNUMA;
  std::vector<EdgeOp> core_graph, updates;
  for (int i = 1; i < argc; i++) {
    const std::string s(argv[i]);
    if (starts_with(s, "-threads=")) {
      threads = std::stoi(s.substr(9));
    } else if (starts_with(s, "-size=")) {
      size = std::stoi(s.substr(6));
    } else if (starts_with(s, "-lock_free")) {
      lock_search = false;
    } else if (starts_with(s, "-insert")) {
      insert = true;
    } else if (starts_with(s, "-delete")) {
      insert = false;
    } else if (starts_with(s, "-pppcsrnuma")) {
      v = Version::PPPCSRNUMA;
    } else if (starts_with(s, "-pppcsr")) {
      v = Version::PPPCSR;
    } else if (starts_with(s, "-ppcsr")) {
      v = Version::PPCSR;
    } else if (starts_with(s, "-partitions_per_domain=")) {
      partitions_per_domain = std::stoi(s.substr(23));
    } else if (starts_with(s, "-gpus=")) {
      setenv("PPCSR_GPUS", s.substr(6).c_str(), 1);
    } else if (starts_with(s, "-check")) {
      check = true;
    } else if (starts_with(s, "-core_graph=")) {
      int top = 0;
      std::tie(core_graph, top) = read_input(s.substr(12), Operation::ADD);
      num_nodes = std::max(num_nodes, top);
    } else if (starts_with(s, "-update_file=")) {
      const std::string name = s.substr(13);
      std::cout << name << std::endl;
      int top = 0;
      std::tie(updates, top) = read_input(name, insert ? Operation::ADD : Operation::DELETE);
      num_nodes = std::max(num_nodes, top);
      size = (int)std::min((size_t)size, updates.size());
    }
  }
  if (core_graph.empty()) {
    std::cout << "Core graph file not specified" << std::endl;
    return EXIT_FAILURE;
  }
  if (updates.empty()) {
    std::cout << "Updates file not specified" << std::endl;
    return EXIT_FAILURE;
  }
  std::cout << "Core graph size: " << core_graph.size() << std::endl;
  bool ok = true;
  if (v == Version::PPCSR) {
    auto pool = std::make_unique<ThreadPool>(threads, lock_search, num_nodes + 1, partitions_per_domain);
    execute(threads, size, core_graph, updates, pool);
    const ppcsr_batch_stats &st = pool->last_stats();
    std::cout << "{\"updates\": " << st.batch_size << ", \"device_ms\": " << st.ms_total
              << ", \"rebalance_bytes_per_update\": " << (st.batch_size ? (double)st.rebalance_bytes / st.batch_size : 0)
              << ", \"windows\": " << st.n_windows << ", \"slots\": " << st.slots_after << "}" << std::endl;
    if (check) ok = pool->pcsr->check_invariants(!insert);
  } else {
    auto pool = std::make_unique<ThreadPoolPPPCSR>(threads, lock_search, num_nodes + 1, partitions_per_domain,
                                                   v == Version::PPPCSRNUMA);
    execute(threads, size, core_graph, updates, pool);
    if (check) {
      for (std::size_t p = 0; p < pool->pcsr->partition_count(); p++)
        ok = pool->pcsr->partition(p).check_invariants(!insert) && ok;
    }
  }
  if (check) std::cout << "PMA invariants: " << (ok ? "ok" : "VIOLATED") << std::endl;
  return ok ? 0 : 2;
}
