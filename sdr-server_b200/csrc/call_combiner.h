/*
 * csrc/call_combiner.h -- combine synchronous calls made by many threads into
 * batches served by a few "launch lanes".  Host logic only (no CUDA in here: the
 * owner supplies the function that serves a batch, so tests/test_call_combiner.py
 * can exercise the synchronisation on a CPU).
 *
 * The reference runs one dsp thread per client, each calling process_* on its own
 * filter (src/dsp_worker.c:41-88).  Every such call that reaches the library while
 * a lane is free becomes the LEADER of a batch: it takes every call queued so far
 * (its own included), serves them with one invocation of the batch function, and
 * wakes exactly the callers it served plus one queued caller to lead the next
 * batch.  A lone caller is always its own leader: no thread hand-off on its path.
 *
 * Synchronisation is built so that hundreds of dsp threads arriving in the same
 * microsecond do not convoy on one mutex: the queue is guarded by a spinlock held for
 * tens of nanoseconds, every call sleeps on its OWN mutex + condition variable, and
 * there is no broadcast anywhere.
 */
#ifndef XLATING_B200_CALL_COMBINER_H_
#define XLATING_B200_CALL_COMBINER_H_

#include <stdint.h>

#include <atomic>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

namespace xl {

class SpinLock {
 public:
  void lock() {
    int spins = 0;
    while (flag_.test_and_set(std::memory_order_acquire)) {
      if (++spins < 128) {
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#endif
      } else {
        std::this_thread::yield();  // the holder may have been preempted
        spins = 0;
      }
    }
  }
  void unlock() { flag_.clear(std::memory_order_release); }

 private:
  std::atomic_flag flag_ = ATOMIC_FLAG_INIT;
};

// One per filter: a filter has at most one call in flight (include/xlating.h, threading).
struct CombinerCall {
  std::mutex m;
  std::condition_variable cv;
  bool done = false;    // guarded by m: the call has been served, `status` is valid
  bool lead = false;    // guarded by m: a lane became free, try to lead
  bool queued = false;  // guarded by the combiner's queue lock
  int status = 0;
  void *user = nullptr;
};

class CallCombiner {
 public:
  static constexpr int kMaxLanes = 8;
  // Serves batch[0..n) on `lane` (lanes are exclusive: at most one batch per lane at a
  // time) and returns the status given to every call in it.
  typedef int (*RunBatch)(void *ctx, int lane, CombinerCall *const *batch, int n);

  CallCombiner(int lanes, int max_batch, RunBatch fn, void *ctx)
      : n_lanes_(lanes < 1 ? 1 : (lanes > kMaxLanes ? kMaxLanes : lanes)), max_batch_(max_batch), fn_(fn), ctx_(ctx) {
    for (int i = 0; i < kMaxLanes; i++) {
      busy_[i] = false;
      lane_batch_[i].reserve((size_t)max_batch);
    }
  }
  int lanes() const { return n_lanes_; }

  // Queue `c` and return its status once it has been served -- by this thread as a
  // leader, or by another caller's batch.
  int run(CombinerCall *c) {
    c->done = c->lead = false;  // nobody else references c between calls
    q_.lock();
    c->queued = true;
    pending_.push_back(c);
    for (;;) {
      // invariant: q_ is held here
      int lane = -1;
      if (c->queued)
        for (int i = 0; i < n_lanes_; i++)
          if (!busy_[i]) {
            lane = i;
            break;
          }
      if (lane >= 0) {
        // leader: everything queued so far, in arrival order.  (With more than
        // max_batch calls queued ahead of its own, this thread serves those first and
        // its own call stays queued for its next round.)
        busy_[lane] = true;
        std::vector<CombinerCall *> &batch = lane_batch_[lane];
        batch.clear();
        while (!pending_.empty() && (int)batch.size() < max_batch_) {
          pending_.front()->queued = false;
          batch.push_back(pending_.front());
          pending_.pop_front();
        }
        batches_++;
        calls_ += batch.size();
        q_.unlock();
        const int rc = fn_(ctx_, lane, batch.data(), (int)batch.size());
        // Free the lane BEFORE waking the callers: every wake-up is a futex system call,
        // and a batch of 20 would otherwise keep the lane idle for 20 of them.  The
        // lane's vector goes with the lane, so keep the served calls in our own.
        static thread_local std::vector<CombinerCall *> served;
        if (served.capacity() < (size_t)max_batch_) served.reserve((size_t)max_batch_);  // the lane gets it back
        served.swap(batch);
        q_.lock();
        busy_[lane] = false;
        if (!pending_.empty() && pending_.front() != c) {
          // hand the free lane to the oldest queued caller.  It is still queued, hence
          // asleep or about to be, and alive; its mutex is only ever held for a few
          // instructions and never while taking q_, so locking it under q_ is safe.
          CombinerCall *next = pending_.front();
          std::lock_guard<std::mutex> g(next->m);
          next->lead = true;
          next->cv.notify_one();
        }
        q_.unlock();
        bool self_served = false;
        for (CombinerCall *b : served) {
          if (b == c) {
            self_served = true;
            continue;
          }
          // notify under b->m: the owner cannot return (and destroy b) before we let go
          std::lock_guard<std::mutex> g(b->m);
          b->status = rc;
          b->done = true;
          b->cv.notify_one();
        }
        served.clear();
        if (self_served) return rc;
        q_.lock();
        continue;
      }
      // no free lane, or another leader already took this call: sleep
      q_.unlock();
      {
        std::unique_lock<std::mutex> lk(c->m);
        c->cv.wait(lk, [c] { return c->done || c->lead; });
        if (c->done) return c->status;
        c->lead = false;
      }
      q_.lock();
    }
  }

  void stats(uint64_t *batches, uint64_t *calls) {
    q_.lock();
    *batches = batches_;
    *calls = calls_;
    q_.unlock();
  }

 private:
  const int n_lanes_, max_batch_;
  const RunBatch fn_;
  void *const ctx_;
  SpinLock q_;  // guards everything below and CombinerCall::queued
  std::deque<CombinerCall *> pending_;
  bool busy_[kMaxLanes];
  std::vector<CombinerCall *> lane_batch_[kMaxLanes];  // owned by the lane's current leader
  uint64_t batches_ = 0, calls_ = 0;
};

}  // namespace xl
#endif
