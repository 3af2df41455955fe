// PPPCSR -- vertex-range partitioned graph with the reference's class surface
// (reference src/pppcsr/PPPCSR.h:11-60): numDomain * partitionsPerDomain independent PCSR shards over
// contiguous source ranges; `src` is made partition-local, `dest` stays global.  A "domain" is a GPU here
// (the reference's NUMA domain): partition p lives on GPU (p / partitionsPerDomain) % #GPUs when
// use_numa is set, else on GPU 0.
#pragma once
#include <cstddef>
#include <vector>

#include "PCSR.h"

class PPPCSR {
 public:
  edge_list_t edges;  // present in the reference too, never initialised there (PPPCSR.h:14)

  PPPCSR(uint32_t init_n, uint32_t src_n, bool lock_search, int numDomain, int partitionsPerDomain, bool use_numa);
  // same, with explicit first vertices of the partitions (ascending, boundaries[0] == 0): the reference's
  // `distribution` vector admits any monotone boundaries (PPPCSR.h:57); edge-balanced ones keep a skewed graph from
  // putting 44 % of its edges on the first of eight GPUs
  PPPCSR(uint32_t init_n, bool lock_search, int partitionsPerDomain, bool use_numa,
         const std::vector<size_t> &boundaries);
  ~PPPCSR();
  PPPCSR(const PPPCSR &) = delete;
  PPPCSR &operator=(const PPPCSR &) = delete;

  bool edge_exists(uint32_t src, uint32_t dest);
  void add_node();
  void add_edge(uint32_t src, uint32_t dest, uint32_t value);
  void remove_edge(uint32_t src, uint32_t dest);
  void read_neighbourhood(int src);
  std::size_t get_partiton(size_t vertex_id) const;  // [sic] the reference's spelling
  std::vector<int> get_neighbourhood(int src) const;
  uint64_t get_n();
  node_t &getNode(int id);
  const node_t &getNode(int id) const;
  void registerThread(int par) { partitions[par].edges.global_lock->registerThread(); }
  void unregisterThread(int par) { partitions[par].edges.global_lock->unregisterThread(); }

  // ---- batched surface for ThreadPoolPPPCSR ----
  std::size_t partition_count() const { return partitions.size(); }
  PCSR &partition(std::size_t p) { return partitions[p]; }
  std::size_t partition_start(std::size_t p) const { return distribution[p]; }
  void pagerank_push(const std::vector<double> &in, std::vector<double> &out) const;
  // One batch of GLOBAL updates (value 0 = remove, value == nullptr = all adds): the batch is cut into one slice per
  // partition, every GPU bins its slice by owner ON THE DEVICE and stores the records straight into the owner's receive
  // buffer over NVLink peer memory, every partition applies what it received (C-ABI ppcsr_group_apply).  Replaces the
  // per-op hand-over of reference ThreadPoolPPPCSR::submit_* (thread_pool_pppcsr.cpp:96-118).  stats: one per partition.
  void apply_batch(const uint32_t *src, const uint32_t *dst, const uint32_t *value, size_t count,
                   std::vector<ppcsr_batch_stats> *stats = nullptr);

 private:
  void build(uint32_t init_n, bool lock_search, bool use_numa);
  std::vector<PCSR> partitions;
  std::vector<size_t> distribution;  // first vertex of every partition
  int partitionsPerDomain;
  ppcsr_group *group_ = nullptr;     // device-side router over the partitions, (re)created for the largest batch seen
  size_t group_cap_ = 0;
  bool group_values_ = false;
};
