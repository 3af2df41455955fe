// HybridLock -- per-leaf lock + version counter of the reference (src/utility/hybridLock.h:12-49).
// Leaves are never locked individually here (a batch owns its windows by construction), so the lock is
// always free; the version counter counts the batches that rewrote the structure.
#pragma once
#include <atomic>

class HybridLock {
 public:
  HybridLock() = default;
  HybridLock(const HybridLock &) = delete;
  HybridLock &operator=(const HybridLock &) = delete;

  HybridLock &operator++() {
    ++version_;
    return *this;
  }
  HybridLock &operator--() {
    --version_;
    return *this;
  }
  void lock() {}
  void unlock() {}
  void lock_shared() {}
  void unlock_shared() {}
  int load() const { return version_.load(); }
  bool lockable() { return true; }

 private:
  std::atomic<int> version_{0};
};
