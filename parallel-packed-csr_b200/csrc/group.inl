// group.inl -- (included by capi.cu, inside extern "C") one process driving several shards on several GPUs:
// the data plane of reference PPPCSR + ThreadPoolPPPCSR (src/pppcsr/PPPCSR.cpp:13-66,
// src/thread_pool_pppcsr/thread_pool_pppcsr.cpp:96-118) behind the C-ABI.
//
// The reference hands every op to a thread of the NUMA domain that owns `src`, one push per op.  Here a batch is
// cut into one contiguous slice per GPU (every PCIe link copies its slice at the same time), each GPU bins its slice
// by owner and stores the records STRAIGHT INTO THE OWNER'S RECEIVE BUFFER over NVLink peer memory (plain cudaMalloc
// buffers mapped with cudaDeviceEnablePeerAccess: one process, no IPC handles), stream events order "all senders have
// stored" before "the owner reads", and every owner applies what it received (ppcsr_apply_batch_segments_device) on
// its own host thread -- the apply has host round trips.  A shard never rebalances across GPUs.
struct ppcsr_group {
  uint32_t n = 0;
  std::vector<ppcsr_shard *> shards;
  std::vector<int> devices;
  std::vector<uint64_t> starts;          // n + 1 vertex boundaries (host)
  std::vector<uint64_t *> d_starts;      // per device copy
  uint64_t cap = 0;                      // records per (sender, receiver) region
  bool with_values = false;
  std::vector<uint64_t *> rec;           // per receiver: n regions of cap records
  std::vector<uint32_t *> val;           // per receiver: n regions of cap values (with_values)
  std::vector<uint64_t *> cnt;           // per receiver: n counts
  std::vector<DevBuf<uint32_t>> in_src, in_dst, in_val;  // per sender: staging of its slice
  std::vector<cudaEvent_t> binned;       // per sender: its records and counts are stored
  std::vector<cudaEvent_t> applied;      // per receiver: it has consumed its receive buffer
  bool first = true;
};

static void group_free(ppcsr_group *g) {
  if (!g) return;
  for (uint32_t r = 0; r < g->n; r++) {
    cudaSetDevice(g->devices[r]);
    if (r < g->rec.size() && g->rec[r]) cudaFree(g->rec[r]);
    if (r < g->val.size() && g->val[r]) cudaFree(g->val[r]);
    if (r < g->cnt.size() && g->cnt[r]) cudaFree(g->cnt[r]);
    if (r < g->d_starts.size() && g->d_starts[r]) cudaFree(g->d_starts[r]);
    if (r < g->in_src.size()) {
      dev_free(g->in_src[r]);
      dev_free(g->in_dst[r]);
      dev_free(g->in_val[r]);
    }
    if (r < g->binned.size() && g->binned[r]) cudaEventDestroy(g->binned[r]);
    if (r < g->applied.size() && g->applied[r]) cudaEventDestroy(g->applied[r]);
  }
  delete g;
}

int ppcsr_group_create(ppcsr_shard **shards, uint32_t n_shards, const uint64_t *starts, uint64_t region_cap,
                       int with_values, ppcsr_group **out) {
  if (!out) return PPCSR_ERR_ARG;
  *out = nullptr;
  if (!shards || !starts || n_shards == 0 || n_shards > batch::BIN_MAX_PARTS || region_cap == 0) return PPCSR_ERR_ARG;
  for (uint32_t r = 0; r < n_shards; r++) {
    if (!shards[r] || starts[r] > starts[r + 1] || starts[r + 1] - starts[r] != shards[r]->n) {
      g_ppcsr_error = "ppcsr_group_create: shard r must hold exactly the vertices [starts[r], starts[r+1])";
      return PPCSR_ERR_ARG;
    }
  }
  ppcsr_group *g = new ppcsr_group();
  g->n = n_shards;
  g->cap = region_cap;
  g->with_values = with_values != 0;
  g->shards.assign(shards, shards + n_shards);
  g->starts.assign(starts, starts + n_shards + 1);
  g->devices.resize(n_shards);
  g->rec.assign(n_shards, nullptr);
  g->val.assign(n_shards, nullptr);
  g->cnt.assign(n_shards, nullptr);
  g->d_starts.assign(n_shards, nullptr);
  g->in_src.resize(n_shards);
  g->in_dst.resize(n_shards);
  g->in_val.resize(n_shards);
  g->binned.assign(n_shards, nullptr);
  g->applied.assign(n_shards, nullptr);
  const int rc = [&]() -> int {
    for (uint32_t r = 0; r < n_shards; r++) g->devices[r] = shards[r]->device;
    for (uint32_t r = 0; r < n_shards; r++) {
      CUDA_TRY(cudaSetDevice(g->devices[r]));
      for (uint32_t p = 0; p < n_shards; p++) {  // every sender stores into every receiver's buffer
        if (g->devices[p] == g->devices[r]) continue;
        int can = 0;
        CUDA_TRY(cudaDeviceCanAccessPeer(&can, g->devices[r], g->devices[p]));
        if (!can) {
          g_ppcsr_error = "ppcsr_group_create: the GPUs of this group cannot address each other's memory";
          return PPCSR_ERR_NO_DEVICE;
        }
        const cudaError_t e = cudaDeviceEnablePeerAccess(g->devices[p], 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CUDA_TRY(e);
        cudaGetLastError();
      }
      CUDA_TRY(cudaMalloc((void **)&g->rec[r], (size_t)n_shards * region_cap * sizeof(uint64_t)));
      if (g->with_values) CUDA_TRY(cudaMalloc((void **)&g->val[r], (size_t)n_shards * region_cap * sizeof(uint32_t)));
      CUDA_TRY(cudaMalloc((void **)&g->cnt[r], (size_t)n_shards * sizeof(uint64_t)));
      CUDA_TRY(cudaMemset(g->cnt[r], 0, (size_t)n_shards * sizeof(uint64_t)));
      CUDA_TRY(cudaMalloc((void **)&g->d_starts[r], (size_t)(n_shards + 1) * sizeof(uint64_t)));
      CUDA_TRY(cudaMemcpy(g->d_starts[r], g->starts.data(), (size_t)(n_shards + 1) * sizeof(uint64_t),
                          cudaMemcpyHostToDevice));
      CUDA_TRY(cudaEventCreateWithFlags(&g->binned[r], cudaEventDisableTiming));
      CUDA_TRY(cudaEventCreateWithFlags(&g->applied[r], cudaEventDisableTiming));
    }
    return PPCSR_OK;
  }();
  if (rc != PPCSR_OK) {
    const std::string why = g_ppcsr_error;
    group_free(g);
    g_ppcsr_error = why;
    return rc;
  }
  *out = g;
  return PPCSR_OK;
}

void ppcsr_group_destroy(ppcsr_group *g) {
  if (!g) return;
  for (uint32_t r = 0; r < g->n; r++) {
    cudaSetDevice(g->devices[r]);
    cudaStreamSynchronize(g->shards[r]->stream);
  }
  group_free(g);
}

// reference PPPCSR::get_partiton (src/pppcsr/PPPCSR.cpp:58-66)
uint32_t ppcsr_group_owner(const ppcsr_group *g, uint64_t vertex) {
  if (!g) return 0;
  for (uint32_t p = 1; p < g->n; p++)
    if (g->starts[p] > vertex) return p - 1;
  return g->n - 1;
}

// One batch: host arrays (pinned or pageable) of GLOBAL (src, dst[, val]); stats[] has one entry per shard (nullable).
// Returns the first failing shard's status.
int ppcsr_group_apply(ppcsr_group *g, const uint32_t *src, const uint32_t *dst, const uint32_t *val, uint64_t count,
                      uint32_t default_val, ppcsr_batch_stats *stats) {
  if (!g || (count && (!src || !dst))) return PPCSR_ERR_ARG;
  if (val && !g->with_values) {
    g_ppcsr_error = "ppcsr_group_apply: the group was created without value buffers";
    return PPCSR_ERR_ARG;
  }
  const uint32_t n = g->n;
  const uint64_t slice = (count + n - 1) / n;
  if (slice > g->cap) {
    g_ppcsr_error = "ppcsr_group_apply: batch / n_shards exceeds the region capacity of the group";
    return PPCSR_ERR_CAPACITY;
  }
  std::vector<uint64_t> h_rec(n), h_val(n), h_cnt(n);
  for (uint32_t p = 0; p < n; p++) {
    h_rec[p] = reinterpret_cast<uint64_t>(g->rec[p]);
    h_val[p] = reinterpret_cast<uint64_t>(g->val[p]);
    h_cnt[p] = reinterpret_cast<uint64_t>(g->cnt[p]);
  }
  // 1. every sender: copy its slice, wait until every receiver has consumed the previous batch, bin + store
  for (uint32_t r = 0; r < n; r++) {
    const uint64_t lo = std::min<uint64_t>(count, (uint64_t)r * slice), hi = std::min<uint64_t>(count, lo + slice);
    const uint64_t c = hi - lo;
    ppcsr_shard *s = g->shards[r];
    CUDA_TRY(cudaSetDevice(g->devices[r]));
    PPCSR_TRY(dev_reserve(g->in_src[r], c, s->stream));
    PPCSR_TRY(dev_reserve(g->in_dst[r], c, s->stream));
    if (val) PPCSR_TRY(dev_reserve(g->in_val[r], c, s->stream));
    if (c) {
      CUDA_TRY(cudaMemcpyAsync(g->in_src[r].p, src + lo, c * 4, cudaMemcpyHostToDevice, s->stream));
      CUDA_TRY(cudaMemcpyAsync(g->in_dst[r].p, dst + lo, c * 4, cudaMemcpyHostToDevice, s->stream));
      if (val) CUDA_TRY(cudaMemcpyAsync(g->in_val[r].p, val + lo, c * 4, cudaMemcpyHostToDevice, s->stream));
    }
    if (!g->first)
      for (uint32_t p = 0; p < n; p++) CUDA_TRY(cudaStreamWaitEvent(s->stream, g->applied[p], 0));
    PPCSR_TRY(ppcsr_bin_to_peers(g->devices[r], s->stream, g->d_starts[r], n, r, g->in_src[r].p, g->in_dst[r].p,
                                 val ? g->in_val[r].p : nullptr, c, h_rec.data(), val ? h_val.data() : nullptr,
                                 h_cnt.data(), g->cap));
    CUDA_TRY(cudaSetDevice(g->devices[r]));
    CUDA_TRY(cudaEventRecord(g->binned[r], s->stream));
  }
  // 2. every receiver: wait for all senders, then apply its regions (one host thread each: the apply synchronises)
  for (uint32_t r = 0; r < n; r++) {
    CUDA_TRY(cudaSetDevice(g->devices[r]));
    for (uint32_t p = 0; p < n; p++) CUDA_TRY(cudaStreamWaitEvent(g->shards[r]->stream, g->binned[p], 0));
  }
  std::vector<int> rcs(n, PPCSR_OK);
  std::vector<std::string> errs(n);
  std::vector<ppcsr_batch_stats> st(n);
  auto work = [&](uint32_t r) {
    rcs[r] = ppcsr_apply_batch_segments_device(g->shards[r], g->rec[r], val ? g->val[r] : nullptr, g->cap, g->cnt[r], n,
                                               count, default_val, &st[r]);
    if (rcs[r] != PPCSR_OK) errs[r] = g_ppcsr_error;
    cudaSetDevice(g->devices[r]);
    cudaEventRecord(g->applied[r], g->shards[r]->stream);
  };
  if (n == 1) {
    work(0);
  } else {
    std::vector<std::thread> ts;
    for (uint32_t r = 0; r < n; r++) ts.emplace_back(work, r);
    for (auto &t : ts) t.join();
  }
  g->first = false;
  for (uint32_t r = 0; r < n; r++) {
    if (stats) stats[r] = st[r];
    if (rcs[r] != PPCSR_OK) {
      g_ppcsr_error = errs[r];
      return rcs[r];
    }
  }
  return PPCSR_OK;
}
