// One PageRank push step with the reference's signature (reference src/utility/pagerank.h:16-29):
//   out[nbr] += values[v] / getNode(v).num_neighbors   for every edge (v, nbr)
// The edge scan runs on the GPU (leaf walk over the packed array, fp64 accumulation) through T::pagerank_push.
#pragma once
#include <cstdint>
#include <vector>

template <typename T, typename weight_t>
std::vector<weight_t> pagerank(T &graph, std::vector<weight_t> const &node_values) {
  const uint64_t n = graph.get_n();
  std::vector<double> in(node_values.begin(), node_values.end());
  in.resize(n, 0.0);
  std::vector<double> acc(n, 0.0);
  graph.pagerank_push(in, acc);
  return std::vector<weight_t>(acc.begin(), acc.end());
}
