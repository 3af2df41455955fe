// ThreadPoolPPPCSR -- the reference's partition-aware scheduler surface
// (reference src/thread_pool_pppcsr/thread_pool_pppcsr.h:17-47).  submit_* route every op to the shard
// that owns `src` (the reference routes it to a thread of the owning NUMA domain,
// thread_pool_pppcsr.cpp:96-101); start() launches one batch per shard, each on its own GPU stream, so
// shards on different GPUs run concurrently; stop() joins them and prints the elapsed time.
#pragma once
#include <atomic>
#include <chrono>
#include <cstdint>
#include <vector>

#include "PPPCSR.h"
#include "task.h"

class ThreadPoolPPPCSR {
 public:
  PPPCSR *pcsr;

  explicit ThreadPoolPPPCSR(const int NUM_OF_THREADS, bool lock_search, uint32_t init_num_nodes,
                            int partitions_per_domain, bool use_numa);
  ~ThreadPoolPPPCSR();

  void submit_add(int thread_id, int src, int dest);
  void submit_delete(int thread_id, int src, int dest);
  void submit_read(int thread_id, int src);
  void start(int threads);
  void stop();

  // thread -> domain table of the reference (thread_pool_pppcsr.cpp:32-47), kept for the CPU-side reads
  const std::vector<int> &thread_to_domain() const { return threadToDomain; }
  const std::vector<int> &first_thread_of_domain() const { return firstThreadDomain; }
  const std::vector<int> &threads_of_domain() const { return numThreadsDomain; }

 private:
  struct Staged {
    std::vector<uint32_t> src, dst, val;
  };
  void stage(int src, int dest, uint32_t value);

  std::vector<Staged> staged_;  // one staging batch per partition, submission order inside
  std::vector<int> reads_;
  std::chrono::steady_clock::time_point t0_, t1_;
  std::atomic_bool finished_;
  uint64_t not_found_ = 0;

  const int available_nodes;  // GPUs used as "domains"
  int partitions_per_domain = 1;
  std::vector<int> threadToDomain;
  std::vector<int> firstThreadDomain;
  std::vector<int> numThreadsDomain;
};
