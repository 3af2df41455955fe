// ThreadPoolPPPCSR over per-GPU shards: see thread_pool_pppcsr.h.
#include "thread_pool_pppcsr.h"

#include <algorithm>
#include <cstdlib>
#include <iostream>
#include <thread>

static int usable_gpus(int threads) {
  int gpus = ppcsr_device_count();
  if (const char *e = std::getenv("PPCSR_GPUS")) gpus = std::min(gpus, std::max(1, std::atoi(e)));
  // the reference uses min(#NUMA nodes, #threads) domains (thread_pool_pppcsr.cpp:24)
  return std::max(1, std::min(gpus, threads));
}

ThreadPoolPPPCSR::ThreadPoolPPPCSR(const int NUM_OF_THREADS, bool lock_search, uint32_t init_num_nodes,
                                   int partitions_per_domain, bool use_numa)
    : finished_(false),
      available_nodes(usable_gpus(NUM_OF_THREADS)),
      partitions_per_domain(partitions_per_domain),
      threadToDomain(NUM_OF_THREADS),
      firstThreadDomain(available_nodes, 0),
      numThreadsDomain(available_nodes, 0) {
  pcsr = new PPPCSR(init_num_nodes, init_num_nodes, lock_search, available_nodes, partitions_per_domain, use_numa);
  staged_.resize(pcsr->partition_count());
  for (std::size_t p = 0; p < pcsr->partition_count(); p++) pcsr->partition(p).print_not_found = false;
  // threads are dealt to domains in contiguous blocks, the first (threads % domains) domains get one more
  const int base = NUM_OF_THREADS / available_nodes, extra = NUM_OF_THREADS % available_nodes;
  int t = 0;
  for (int d = 0; d < available_nodes && t < NUM_OF_THREADS; d++) {
    const int take = base + (d < extra ? 1 : 0);
    firstThreadDomain[d] = t;
    numThreadsDomain[d] = take;
    for (int k = 0; k < take; k++) threadToDomain[t++] = d;
  }
}

ThreadPoolPPPCSR::~ThreadPoolPPPCSR() { delete pcsr; }

void ThreadPoolPPPCSR::stage(int src, int dest, uint32_t value) {
  const std::size_t p = pcsr->get_partiton((size_t)src);
  Staged &s = staged_[p];
  s.src.push_back((uint32_t)src - (uint32_t)pcsr->partition_start(p));  // partition-local id (PPPCSR.cpp:46-52)
  s.dst.push_back((uint32_t)dest);
  s.val.push_back(value);
}

void ThreadPoolPPPCSR::submit_add(int thread_id, int src, int dest) {
  (void)thread_id;
  stage(src, dest, 1u);
}
void ThreadPoolPPPCSR::submit_delete(int thread_id, int src, int dest) {
  (void)thread_id;
  stage(src, dest, 0u);
}
void ThreadPoolPPPCSR::submit_read(int thread_id, int src) {
  (void)thread_id;
  reads_.push_back(src);
}

void ThreadPoolPPPCSR::start(int threads) {
  (void)threads;
  t0_ = std::chrono::steady_clock::now();
  finished_ = false;
  not_found_ = 0;
  // one host thread per shard so that shards living on different GPUs overlap
  std::vector<std::thread> workers;
  std::vector<ppcsr_batch_stats> stats(staged_.size());
  for (std::size_t p = 0; p < staged_.size(); p++) {
    std::cout << "Thread " << p << " has " << staged_[p].src.size() << " tasks, runs on domain "
              << pcsr->partition(p).device() << std::endl;
    if (staged_[p].src.empty()) continue;
    workers.emplace_back([this, p, &stats]() {
      pcsr->registerThread((int)p);
      pcsr->partition(p).apply_batch(staged_[p].src, staged_[p].dst, staged_[p].val, &stats[p]);
      pcsr->unregisterThread((int)p);
    });
  }
  for (auto &w : workers) w.join();
  for (auto &st : stats) not_found_ += st.n_not_found;
  for (int v : reads_) pcsr->read_neighbourhood(v);
}

void ThreadPoolPPPCSR::stop() {
  finished_ = true;
  for (std::size_t p = 0; p < pcsr->partition_count(); p++) {
    if (ppcsr_sync(pcsr->partition(p).handle()) != PPCSR_OK) {
      std::cout << "device synchronisation failed: " << ppcsr_last_error() << ". Abort\n";
      std::exit(EXIT_FAILURE);
    }
    std::cout << "Done" << std::endl;
  }
  t1_ = std::chrono::steady_clock::now();
  if (not_found_) std::cout << "not found " << not_found_ << " edges" << std::endl;
  std::cout << "Elapsed wall clock time: "
            << std::chrono::duration_cast<std::chrono::milliseconds>(t1_ - t0_).count() << std::endl;
  for (auto &s : staged_) {
    s.src.clear();
    s.dst.clear();
    s.val.clear();
  }
  reads_.clear();
}
