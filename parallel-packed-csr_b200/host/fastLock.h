// FastLock -- API-compatible stand-in for the reference's global safepoint lock
// (reference src/utility/fastLock.h:14-62).  The batched GPU engine has no per-operation locking: the
// kernel/stream boundary is the safepoint.  The type survives because the reference's tests and
// schedulers call lockable()/registerThread()/unregisterThread() on `edges.global_lock`
// (reference test/DataStructureTest.cpp:62,86,92, src/thread_pool/thread_pool.cpp:40,56).
// Single-op calls from many host threads are serialised on `mutex()` by the PCSR shell.
#pragma once
#include <atomic>
#include <mutex>

class FastLock {
 public:
  FastLock() = default;
  FastLock(const FastLock &) = delete;
  FastLock &operator=(const FastLock &) = delete;

  void lock() { host_.lock(); }
  void unlock() { host_.unlock(); }
  void lock_shared() {}
  void unlock_shared() {}
  void registerThread() { registered_.fetch_add(1); }
  void unregisterThread() { registered_.fetch_sub(1); }
  unsigned registered() const { return registered_.load(); }
  // true when no host thread is inside a device call on this shard
  bool lockable() {
    if (!host_.try_lock()) return false;
    host_.unlock();
    return true;
  }
  std::mutex &mutex() { return host_; }

 private:
  std::mutex host_;
  std::atomic<unsigned> registered_{0};
};
