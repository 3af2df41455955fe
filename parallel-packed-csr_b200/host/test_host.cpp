// The reference's GoogleTest suite restated against the B200 host classes (no gtest in this image):
// reference test/DataStructureTest.cpp:12-213 (both lock_search modes) and test/SchedulerTest.cpp:11-58.
// Needs a GPU; run by tests/test_gpu_host.py.
#include <cstdio>
#include <cstdlib>
#include <map>
#include <numeric>
#include <set>
#include <string>
#include <thread>
#include <vector>

#include "PPPCSR.h"
#include "bfs.h"
#include "pagerank.h"
#include "thread_pool.h"
#include "thread_pool_pppcsr.h"

static int g_failed = 0, g_checks = 0;
#define EXPECT(cond)                                                            \
  do {                                                                          \
    g_checks++;                                                                 \
    if (!(cond)) {                                                              \
      g_failed++;                                                               \
      std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);             \
    }                                                                           \
  } while (0)

static bool all_locks_free(PCSR &g) {
  for (uint64_t j = 0; j < g.edges.N / g.edges.logN; ++j)
    if (!g.edges.node_locks[j]->lockable()) return false;
  return g.edges.global_lock->lockable();
}

static void t_initialization(bool ls) {
  PPPCSR g(10, 10, ls, 1, 1, false);
  EXPECT(g.get_n() == 10);
}
static void t_add_node(bool ls) {
  PPPCSR g(0, 0, ls, 1, 1, false);
  EXPECT(g.get_n() == 0);
  g.add_node();
  EXPECT(g.get_n() == 1);
  EXPECT(g.get_neighbourhood(0).size() == 0);
}
static void t_add_edge(bool ls) {
  PPPCSR g(10, 10, ls, 1, 1, false);
  g.add_edge(11, 1, 1);  // no such vertex: ignored
  g.add_edge(0, 1, 1);
  EXPECT(g.edge_exists(0, 1));
  EXPECT(g.get_neighbourhood(0).size() == 1);
  EXPECT(g.get_n() == 10);
  EXPECT(g.get_neighbourhood(2).size() == 0);
}
static void t_remove_edge(bool ls) {
  PPPCSR g(10, 10, ls, 1, 1, false);
  g.add_node();
  g.remove_edge(0, 1);
  EXPECT(!g.edge_exists(0, 1));
  g.add_edge(0, 1, 1);
  EXPECT(g.edge_exists(0, 1));
  EXPECT(g.get_neighbourhood(0).size() == 1);
  g.remove_edge(0, 1);
  EXPECT(!g.edge_exists(0, 1));
  EXPECT(g.get_neighbourhood(2).size() == 0);
}
static void t_add_remove_seq(bool ls, int edge_count) {
  PCSR g(10, 10, ls, 0);
  g.print_not_found = false;
  for (int i = 1; i <= edge_count; ++i) {
    g.add_edge(0, i, i);
    EXPECT(g.edge_exists(0, i));
    EXPECT(all_locks_free(g));
  }
  EXPECT(g.get_n() == 10);
  EXPECT((int)g.getNode(0).num_neighbors == edge_count);
  for (int i = 1; i <= edge_count; ++i) {
    g.remove_edge(0, i);
    EXPECT(!g.edge_exists(0, i));
    EXPECT(all_locks_free(g));
  }
  EXPECT(g.get_neighbourhood(0).size() == 0);
  EXPECT(g.get_n() == 10);
  EXPECT(g.check_invariants(true));
}
// concurrent single-op calls from registered host threads (the reference uses OpenMP, :81-120)
static void t_add_remove_par(bool ls, int edge_count, int nthreads) {
  PCSR g(10, 10, ls, 0);
  g.print_not_found = false;
  auto worker = [&](int t, bool add) {
    g.edges.global_lock->registerThread();
    for (int i = 1 + t; i <= edge_count; i += nthreads) {
      if (add) {
        g.add_edge(0, i, i);
        EXPECT(g.edge_exists(0, i));
      } else {
        g.remove_edge(0, i);
        EXPECT(!g.edge_exists(0, i));
      }
    }
    g.edges.global_lock->unregisterThread();
  };
  for (int phase = 0; phase < 2; phase++) {
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++) th.emplace_back(worker, t, phase == 0);
    for (auto &x : th) x.join();
    EXPECT(all_locks_free(g));
    EXPECT(g.get_n() == 10);
    if (phase == 0) EXPECT((int)g.getNode(0).num_neighbors == edge_count);
  }
  EXPECT(g.get_neighbourhood(0).size() == 0);
}
static void t_random_seq(bool ls, int ops) {
  PCSR g(1000, 1000, ls, 0);
  g.print_not_found = false;
  for (int i = 1; i <= ops; ++i) {
    const int src = std::rand() % 1000, target = std::rand() % 1000;
    if (std::rand() % 4 != 0) {
      g.add_edge(src, target, i);
      EXPECT(g.edge_exists(src, target));
    } else {
      g.remove_edge(src, target);
      EXPECT(!g.edge_exists(src, target));
    }
    if (i % 64 == 0) EXPECT(all_locks_free(g));
  }
  EXPECT(g.check_invariants(true));
}
static void t_bfs_pagerank(bool ls, int edge_count) {
  PCSR g(1000, 1000, ls, 0);
  std::vector<uint32_t> s, d, v;
  for (int i = 1; i <= edge_count; ++i) {
    s.push_back(std::rand() % 1000);
    d.push_back(std::rand() % 1000);
    v.push_back(i);
  }
  g.apply_batch(s, d, v);
  auto levels = bfs(g, 0);
  EXPECT(levels.size() == 1000);
  EXPECT(levels[0] == 0);
  std::vector<float> w(g.get_n(), 1.0f);
  auto pr = pagerank(g, w);
  EXPECT(pr.size() == 1000);
  double total = 0;
  for (float x : pr) total += x;
  // every vertex with out-edges spreads exactly degree/num_neighbors of its weight
  EXPECT(total > 0 && total <= 1000.0 + 1e-3);
}
static void t_scheduler_table() {
  // reference test/SchedulerTest.cpp:11-58 against ThreadPoolPPPCSR's table; the number of domains is the
  // number of GPUs actually present, so sweep the thread count only
  for (int t = 1; t <= 64; t += 7) {
    ThreadPoolPPPCSR pool(t, true, 64, 1, true);
    const auto &t2d = pool.thread_to_domain();
    const auto &first = pool.first_thread_of_domain();
    const auto &num = pool.threads_of_domain();
    const int d = (int)num.size();
    EXPECT(std::accumulate(num.begin(), num.end(), 0) == t);
    EXPECT(first[0] == 0);
    std::set<int> domains(t2d.begin(), t2d.end());
    EXPECT((int)domains.size() == std::min(t, d));
    std::set<int> sizes(num.begin(), num.end());
    EXPECT(sizes.size() >= 1 && sizes.size() <= 2);
  }
}
static void t_pools() {
  // both schedulers end with the same logical graph on the same stream
  ThreadPool a(8, true, 200, 1);
  ThreadPoolPPPCSR b(8, true, 200, 3, false);
  std::srand(7);
  for (int i = 0; i < 5000; i++) {
    const int s = std::rand() % 200, d = std::rand() % 200;
    if (std::rand() % 4) {
      a.submit_add(i % 8, s, d);
      b.submit_add(i % 8, s, d);
    } else {
      a.submit_delete(i % 8, s, d);
      b.submit_delete(i % 8, s, d);
    }
  }
  a.start(8);
  a.stop();
  b.start(8);
  b.stop();
  bool same = true;
  for (int v = 0; v < 200; v++) same = same && a.pcsr->get_neighbourhood(v) == b.pcsr->get_neighbourhood(v);
  EXPECT(same);
  // 3 partitions per domain, one domain per GPU present (1 GPU: starts 0 / 66 / 132)
  const std::size_t parts = b.pcsr->partition_count(), share = 200 / parts;
  EXPECT(parts % 3 == 0);
  EXPECT(b.pcsr->get_partiton(0) == 0 && b.pcsr->get_partiton(199) == parts - 1 && b.pcsr->get_partiton(share) == 1 &&
         b.pcsr->get_partiton(share - 1) == 0);
  EXPECT(a.pcsr->getNode(5).num_neighbors == b.pcsr->getNode(5).num_neighbors);
}

int main(int argc, char **argv) {
  const bool quick = argc > 1 && std::string(argv[1]) == "--quick";
  for (int mode = 0; mode < 2; mode++) {
    const bool ls = mode == 1;
    t_initialization(ls);
    t_add_node(ls);
    t_add_edge(ls);
    t_remove_edge(ls);
    t_add_remove_seq(ls, quick ? 300 : 2000);
    t_add_remove_par(ls, quick ? 200 : 1000, 4);
    t_random_seq(ls, quick ? 500 : 3000);
    t_bfs_pagerank(ls, 50000);
  }
  t_scheduler_table();
  t_pools();
  std::printf("host tests: %d checks, %d failed\n", g_checks, g_failed);
  return g_failed ? 1 : 0;
}
