/*
 * csrc/stream_overlay.h -- "the filters of one SDR stream are one batch", discovered
 * behind the reference's per-filter ABI.  Host logic only (no CUDA in here: the owner
 * supplies the group operations, so tests/test_stream_overlay.py can exercise the
 * state machine on a CPU against the oracle).
 *
 * Why.  The reference gives every client a filter and a dsp thread; sdr_callback
 * memcpy's each SDR block into every client's queue (src/tcp_server.c:262-269 ->
 * src/queue.c:114) and every dsp thread calls process_* on its private copy
 * (src/dsp_worker.c:49-72).  The library is therefore handed the SAME block sequence C
 * times, by C threads, and the fast way to serve that is the batch engine
 * (include/xlating_group.h): one H2D copy, one fused launch for all clients.  This
 * class finds that structure at run time, without trusting anything but bytes:
 *
 *   * a LOG of the last R blocks of "the stream": pinned copies, in publication order;
 *     block k was submitted ONCE to a batch group for every member filter;
 *   * a filter that is a MEMBER expects block pos of the log next.  Its call compares its
 *     input with that entry (memcmp, always -- the sampled key only short-cuts
 *     mismatches), waits for the block's ticket and copies ITS row of the result.  The
 *     first member to arrive with a block the log does not hold yet PUBLISHES it: copies
 *     it into the log, submits it to the group for everybody, and wakes the others when
 *     the GPU is done;
 *   * a member whose input is NOT the expected block (its queue dropped a block,
 *     src/queue.c:90-94; it lagged more than R blocks; it is fed by another source) is
 *     DESYNCED: it leaves the group and is served by the per-filter engine from then on,
 *     starting from its own mirror of the state (the caller keeps that mirror: history
 *     tail, history_offset, oscillator as of the last block it really consumed) -- every
 *     group result computed for blocks it did not consume is simply never read;
 *   * a PRIVATE filter tracks the log too: once the blocks it has consumed are the log's
 *     newest ones (a run covering its history, or everything since it was created) it
 *     JOINS the group with its state (xlg_add_client_ex) and is a member from the next
 *     block on.
 *
 * Everything a filter returns is therefore either the private engine's result for the
 * bytes it was given, or the group's result for a block whose bytes were compared equal
 * to the ones it was given, computed from a state that consumed exactly the blocks the
 * filter consumed: results never depend on scheduling, only speed does.
 *
 * Threads.  The callers are the reference's dsp threads -- hundreds of them, and whichever
 * arrives first with a block publishes it, so the publisher is a different thread every
 * time.  The group operations (CUDA calls behind StreamOps) are therefore NOT made by the
 * callers: the first CUDA call of every new host thread costs milliseconds of per-thread
 * runtime setup (measured: 2.5-3.2 ms per published block with 256 dsp threads).  Two
 * service threads per stream make them: the SUBMITTER executes submit / add / remove
 * requests strictly in the order they were queued (that order is what makes a joining
 * filter part of exactly the blocks after the one it consumed last), the WAITER waits for
 * each submitted block's ticket and wakes the callers sleeping on the log entry.
 *
 * Publishing rule for private filters: only a filter that is "in step" (it consumed the
 * log's newest block) may append a block -- or anybody while the stream has no members
 * yet (bootstrap).  A filter fed by a different source can then never push a foreign
 * block between the members and their next block.
 */
#ifndef XLATING_B200_STREAM_OVERLAY_H_
#define XLATING_B200_STREAM_OVERLAY_H_

#include <errno.h>
#include <limits.h>
#include <sched.h>
#include <stdlib.h>
#include <linux/futex.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>
#include <sys/syscall.h>
#include <time.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

#include "block_cache.h"  // block_key()

namespace xl {

struct StreamOps {
  void *ctx;
  // page-locked memory for the log's block copies
  void *(*alloc_block)(void *ctx, size_t bytes);
  void (*free_block)(void *ctx, void *p);
  // submit one block (already in a log entry's pinned buffer) for ALL members; asynchronous
  int (*submit)(void *ctx, int fmt, const void *block, size_t elems, int64_t *ticket);
  // block until the ticket's results are readable
  int (*wait)(void *ctx, int64_t ticket);
  // attach `filter` (opaque) which has consumed the last `valid_history` samples of the stream
  int (*add)(void *ctx, void *filter, int64_t valid_history, int *client);
  int (*remove)(void *ctx, int client);
};

class AutoStream {
 public:
  struct Member {          // one per filter, owned by the filter; only its own thread touches it
    bool member = false;
    int client = -1;       // group client id while a member
    int64_t pos = -1;      // log index of the block this filter consumes next (-1: not tracking)
    int64_t run = 0;       // samples of CONSECUTIVE log blocks consumed, ending at pos-1
    int lane = -1;         // which of the kLanes futex words this filter sleeps on (assigned at first use)
  };
  // Hundreds of dsp threads wait for the same event (a block's results, a publisher's append).  On ONE
  // futex word every wait and the final wake-all go through one kernel hash bucket and its spinlock --
  // measured: the serialised cost per caller GROWS with the number of callers (4 us at 256 threads, 10 us
  // at 1000).  So every event has kLanes words on separate cache lines, a filter always sleeps on the same
  // lane, and whoever signals the event sets and wakes them all.
  static constexpr int kLanes = 16;
  struct alignas(64) Word {
    std::atomic<int> v{-1};
  };
  struct Served {
    int64_t ticket;
    int client;
  };
  struct Stats {
    uint64_t published, hits, desyncs, joins, private_matches;
    // where the callers' time goes (nanoseconds, summed over all threads): comparing a call's bytes
    // with the log entry; sleeping until the block's results are readable; and, for publishers, the
    // copy into the log, the group submit and the wait for the GPU
    uint64_t ns_compare, ns_wait, ns_pub_copy, ns_pub_submit, ns_pub_wait;
  };
  static uint64_t now_ns() {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (uint64_t)ts.tv_sec * 1000000000ull + (uint64_t)ts.tv_nsec;
  }

  AutoStream(const StreamOps &ops, int ring, size_t max_block_bytes)
      : ops_(ops), ring_(ring < 4 ? 4 : ring), max_bytes_(max_block_bytes), log_((size_t)(ring < 4 ? 4 : ring)) {
    if (getenv("XLATING_B200_SPIN_ITERS") != nullptr)
      spin_iters_ = std::min(std::max(atol(getenv("XLATING_B200_SPIN_ITERS")), 0l), 1000000l);
    two_level_ = getenv("XLATING_B200_WAKE") != nullptr && strcmp(getenv("XLATING_B200_WAKE"), "two_level") == 0;
    // the log's block copies are allocated here, by the creating thread, not lazily by whichever
    // dsp thread publishes first (page-locking memory from a fresh thread costs milliseconds)
    for (Entry &e : log_) e.host = ops_.alloc_block(ops_.ctx, max_bytes_ > 0 ? max_bytes_ : 1);
    submitter_ = std::thread([this] { submitter_main(); });
    waiter_ = std::thread([this] { waiter_main(); });
  }
  ~AutoStream() {
    {
      std::lock_guard<std::mutex> lk(q_mu_);
      Task t;
      t.kind = Task::kStop;
      tasks_.push_back(t);
    }
    q_cv_.notify_all();
    submitter_.join();
    {
      std::lock_guard<std::mutex> lk(w_mu_);
      waits_.push_back(Pending{-1, -1});
    }
    w_cv_.notify_all();
    waiter_.join();
    for (Entry &e : log_)
      if (e.host != nullptr) ops_.free_block(ops_.ctx, e.host);
  }
  AutoStream(const AutoStream &) = delete;
  AutoStream &operator=(const AutoStream &) = delete;

  int ring() const { return ring_; }
  int members() const { return n_members_.load(); }

  // ---- member path.  1: served (the ticket is complete, copy your row); 0: not served,
  // the filter is now private (it was desynced and has left the group) -- process the
  // block privately and call private_observe(); never fails otherwise.
  int member_call(Member &m, const void *input, size_t bytes, int fmt, size_t elems, Served *sv) {
    if (!m.member) return 0;
    const int64_t k = m.pos;
    bool mine = false;  // this caller appended block k itself: its bytes ARE the entry's
    for (;;) {
      const int64_t h = head_.load(std::memory_order_acquire);
      if (k <= h) {
        Entry &e = log_[(size_t)(k % ring_)];
        const uint64_t t0 = now_ns();
        const bool same = mine ? e.seq.load(std::memory_order_acquire) == k : matches(e, k, input, bytes, fmt);
        const uint64_t t1 = now_ns();
        Shard &sh = shards_[(size_t)lane_of(m)];
        sh.ns_compare.fetch_add(t1 - t0, std::memory_order_relaxed);
        if (!same) break;  // recycled or different bytes: desync
        const bool ready = wait_done(e, k, lane_of(m));
        sh.ns_wait.fetch_add(now_ns() - t1, std::memory_order_relaxed);
        if (!ready) break;
        const int64_t ticket = e.ticket;
        std::atomic_thread_fence(std::memory_order_acquire);
        if (e.seq.load() != k) break;
        sv->ticket = ticket;
        sv->client = m.client;
        m.pos = k + 1;
        m.run += (int64_t)(elems / 2);
        sh.hits.fetch_add(1, std::memory_order_relaxed);
        return 1;
      }
      // k == h + 1: nobody has brought this block yet
      const int rc = publish(k, h, input, bytes, fmt, elems, false, lane_of(m));
      if (rc < 0) break;  // the log cannot take the block: serve it privately
      mine = rc == 0;
      // rc == 0: published by this caller, rc == 1: by somebody else meanwhile -- either way the
      // block is (being) computed for every member: compare and wait like everybody else
    }
    leave(m);
    desyncs_.fetch_add(1, std::memory_order_relaxed);
    return 0;
  }

  // ---- private path bookkeeping: call for every cf32 block a non-member consumes, BEFORE it is
  // processed (so that members are not kept waiting for a block this caller brings first).  With
  // retry = true (call again after the private processing if the first call returned false) only
  // looks the block up.  Returns whether the block is now known to be block m.pos-1 of the log.
  bool private_observe(Member &m, const void *input, size_t bytes, int fmt, size_t elems, bool retry = false) {
    if (m.member) return false;
    const int64_t n = (int64_t)(elems / 2);
    const uint64_t key = bytes > 0 ? block_key(input, bytes) : 0;
    for (;;) {
      const int64_t h = head_.load(std::memory_order_acquire);
      // the expected next block first, then the newest few
      if (m.pos >= 0 && m.pos <= h && matches(log_[(size_t)(m.pos % ring_)], m.pos, input, bytes, fmt, &key)) {
        m.pos += 1;
        m.run += n;
        private_matches_.fetch_add(1, std::memory_order_relaxed);
        return true;
      }
      for (int64_t k = h; k >= 0 && k > h - 4; k--) {
        if (k == m.pos) continue;
        if (matches(log_[(size_t)(k % ring_)], k, input, bytes, fmt, &key)) {
          m.pos = k + 1;  // a new run starts with this block
          m.run = n;
          private_matches_.fetch_add(1, std::memory_order_relaxed);
          return true;
        }
      }
      if (retry) break;
      // not in the log.  May this caller append it?
      const bool in_step = m.pos == h + 1 && m.run > 0;
      if (!in_step && n_members_.load() > 0) break;
      const int rc = publish(h + 1, h, input, bytes, fmt, elems, /*bootstrap_only=*/!in_step, lane_of(m));
      if (rc == 1) continue;  // raced with another publisher: look again
      if (rc < 0) break;
      m.run = in_step ? m.run + n : n;
      m.pos = h + 2;
      return true;
    }
    if (retry) {
      m.pos = -1;
      m.run = 0;
    }
    return false;
  }

  // ---- after a private block: become a member from the next block on, if this filter's
  // consumed blocks are exactly the log's newest ones.  needed_history = samples of real history
  // the filter's next window can reach back (taps_len - 1); total_consumed = all samples it has
  // consumed since it was created.  The caller's state (history_offset, oscillator) must be that
  // after its last consumed block; `filter` is handed to StreamOps::add.
  bool try_join(Member &m, int64_t needed_history, int64_t total_consumed, void *filter) {
    if (m.member || m.pos < 0 || m.run <= 0) return false;
    if (m.run < needed_history && m.run != total_consumed) return false;
    std::lock_guard<std::mutex> lk(mu_);
    if (m.pos != head_.load() + 1) return false;  // not caught up: the newest block is not its last one
    // queued behind every submit so far, and no block can be published while mu_ is held: the filter
    // becomes a member of exactly the blocks after the one it consumed last
    int client = -1;
    Task t;
    t.kind = Task::kAdd;
    t.filter = filter;
    t.valid_history = m.run;
    if (run_sync(t, &client) != 0) return false;
    m.member = true;
    m.client = client;
    n_members_.fetch_add(1);
    joins_.fetch_add(1, std::memory_order_relaxed);
    return true;
  }

  // leave the group (desync, a Q15 call, destroy).  The filter keeps no claim on the log.
  void leave(Member &m) {
    if (m.member) {
      std::lock_guard<std::mutex> lk(mu_);
      Task t;
      t.kind = Task::kRemove;
      t.client = m.client;
      enqueue(t);  // in order with the submits; nobody waits for it
      n_members_.fetch_sub(1);
    }
    m.member = false;
    m.client = -1;
    m.pos = -1;
    m.run = 0;
  }

  Stats stats() const {
    uint64_t hits = 0, ns_compare = 0, ns_wait = 0;
    for (const Shard &sh : shards_) {
      hits += sh.hits.load();
      ns_compare += sh.ns_compare.load();
      ns_wait += sh.ns_wait.load();
    }
    return Stats{published_.load(), hits,    desyncs_.load(),     joins_.load(),         private_matches_.load(),
                 ns_compare,        ns_wait, ns_pub_copy_.load(), ns_pub_submit_.load(), ns_pub_wait_.load()};
  }

 private:
  struct Entry {
    std::atomic<int64_t> seq{-1};  // log index held, -1 while being (re)written
    uint64_t key = 0;
    size_t bytes = 0;
    int fmt = 0;
    size_t elems = 0;
    void *host = nullptr;          // pinned copy of the block
    int64_t ticket = -1;           // written by the submitter before the block is marked done
    // futex word: (log index << 2) | state, state 0 = pending, 1 = results readable, 2 = failed.
    // The index is part of the word because the thread that published block k may be
    // preempted between the GPU finishing and its store: by then the entry may hold
    // block k + R, and a plain "done = 1" would release that block's readers early.
    Word done[kLanes];
  };

  static void futex_wait(std::atomic<int> *w, int expected) {
    syscall(SYS_futex, reinterpret_cast<int *>(w), FUTEX_WAIT_PRIVATE, expected, nullptr, nullptr, 0);
  }
  static void futex_wait_us(std::atomic<int> *w, int expected, long us) {
    struct timespec ts = {us / 1000000, (us % 1000000) * 1000};
    syscall(SYS_futex, reinterpret_cast<int *>(w), FUTEX_WAIT_PRIVATE, expected, &ts, nullptr, 0);
  }
  static void futex_wake_n(std::atomic<int> *w, int n) {
    syscall(SYS_futex, reinterpret_cast<int *>(w), FUTEX_WAKE_PRIVATE, n, nullptr, nullptr, 0);
  }
  static void futex_wake_all(std::atomic<int> *w) {
    syscall(SYS_futex, reinterpret_cast<int *>(w), FUTEX_WAKE_PRIVATE, INT_MAX, nullptr, nullptr, 0);
  }

  // entry holds log block k and its bytes equal input's (seqlock: the entry may be recycled
  // by a publisher while we compare)
  bool matches(Entry &e, int64_t k, const void *input, size_t bytes, int fmt, const uint64_t *key = nullptr) {
    if (e.seq.load(std::memory_order_acquire) != k) return false;
    if (e.bytes != bytes || e.fmt != fmt) return false;
    if (key != nullptr && e.key != *key) return false;
    if (bytes > 0 && memcmp(e.host, input, bytes) != 0) return false;
    std::atomic_thread_fence(std::memory_order_acquire);
    return e.seq.load() == k;
  }

  static int done_tag(int64_t k) { return (int)((uint32_t)(k & 0x1fffffff) << 2); }
  int lane_of(Member &m) {
    if (m.lane < 0) m.lane = (int)(next_lane_.fetch_add(1, std::memory_order_relaxed) % kLanes);
    return m.lane;
  }
  // spin for a few microseconds first: by the time a caller has compared its 256 KiB the block is often
  // about to be ready, and a futex sleep + wake costs more than that
  // Waiting for an event that is a few hundred microseconds away (a block's results; a publisher's
  // append): poll for ~20 us -- by the time a caller has compared its 256 KiB the event is often about
  // to happen, and a futex sleep costs the sleeper a context switch and its waker 2-5 us of kernel time
  // -- then sleep.  Measured alternatives, all worse on a 128-thread host (256 dsp threads, MS/s in):
  // sleeping at once 100-110; a wake-up tree (every sleeper wakes two more) 84; two levels (one sleeper
  // per lane wakes its lane) 101-113; polling with sched_yield for 400 us 36; this, with the signaller
  // waking all sixteen lanes itself, 272.  XLATING_B200_SPIN_ITERS changes the polling length.
  void wait_word(std::atomic<int> *w, int expected, long timeout_us) {
    for (long i = 0; i < spin_iters_; i++) {
      if (w->load(std::memory_order_acquire) != expected) return;
#if defined(__x86_64__) || defined(__i386__)
      __builtin_ia32_pause();
#endif
    }
    if (timeout_us > 0)
      futex_wait_us(w, expected, timeout_us);
    else
      futex_wait(w, expected);
    if (two_level_ && w->load(std::memory_order_acquire) != expected) futex_wake_all(w);
  }

  bool wait_done(Entry &e, int64_t k, int lane) {
    const int tag = done_tag(k);
    std::atomic<int> *w = &e.done[lane].v;
    for (;;) {
      const int d = w->load(std::memory_order_acquire);
      if (e.seq.load() != k) return false;  // recycled: this filter lagged a whole ring
      if (d == (tag | 1)) return true;
      if (d == (tag | 2)) return false;
      if (d == tag) wait_word(w, tag, 0);
    }
  }

  // Append block k = h+1 and queue its submission.  0: appended by this caller; 1: head moved
  // (someone else appended); <0: error, nothing appended.
  int publish(int64_t k, int64_t h, const void *input, size_t bytes, int fmt, size_t elems, bool bootstrap_only = false,
              int lane = 0) {
    // When an SDR block lands, every dsp thread arrives here within microseconds with the
    // same new block.  One appends it; the others must NOT queue up on the mutex (255 hand-
    // offs of a contended lock, each a futex round trip, cost milliseconds per block): they
    // sleep until the head moves and then take the lock-free reader path.
    std::unique_lock<std::mutex> lk(mu_, std::try_to_lock);
    if (!lk.owns_lock()) {
      std::atomic<int> *w = &head_words_[(size_t)(lane >= 0 ? lane : 0)].v;
      if (w->load(std::memory_order_acquire) == (int)h) wait_word(w, (int)h, 200);
      return 1;
    }
    if (head_.load() != h) return 1;
    // a caller that is not in step may only start the log while the stream has no members
    // (membership changes under this same lock, so the check cannot go stale)
    if (bootstrap_only && n_members_.load() > 0) return -EBUSY;
    if (bytes > max_bytes_) return -EINVAL;
    Entry *e = &log_[(size_t)(k % ring_)];
    e->seq.store(-1);  // readers of the old block back off
    if (e->host == nullptr) {
      e->host = ops_.alloc_block(ops_.ctx, max_bytes_ > 0 ? max_bytes_ : 1);
      if (e->host == nullptr) return -ENOMEM;
    }
    const uint64_t t0 = now_ns();
    memcpy(e->host, input, bytes);
    e->key = bytes > 0 ? block_key(input, bytes) : 0;
    ns_pub_copy_.fetch_add(now_ns() - t0, std::memory_order_relaxed);
    e->bytes = bytes;
    e->fmt = fmt;
    e->elems = elems;
    e->ticket = -1;
    for (Word &w : e->done) w.v.store(done_tag(k));
    e->seq.store(k, std::memory_order_release);
    head_.store(k, std::memory_order_release);
    for (Word &w : head_words_) w.v.store((int)k, std::memory_order_release);
    Task t;
    t.kind = Task::kSubmit;
    t.k = k;
    enqueue(t);
    lk.unlock();
    for (Word &w : head_words_) futex_wake_all(&w.v);
    published_.fetch_add(1, std::memory_order_relaxed);
    return 0;
  }

  // ---- service threads ----
  struct Task {
    enum { kSubmit, kAdd, kRemove, kStop } kind = kSubmit;
    int64_t k = -1;            // kSubmit: log index
    void *filter = nullptr;    // kAdd
    int64_t valid_history = 0;
    int client = -1;           // kRemove
    int *out_client = nullptr; // kAdd: result
    int *out_rc = nullptr;
    std::atomic<int> *finished = nullptr;  // kAdd: futex word set to 1 when done
  };
  struct Pending {
    int64_t k, ticket;
  };

  void enqueue(const Task &t) {
    {
      std::lock_guard<std::mutex> lk(q_mu_);
      tasks_.push_back(t);
    }
    q_cv_.notify_one();
  }
  int run_sync(Task &t, int *client) {
    int rc = -EIO;
    std::atomic<int> fin{0};
    t.out_client = client;
    t.out_rc = &rc;
    t.finished = &fin;
    enqueue(t);
    while (fin.load(std::memory_order_acquire) == 0) futex_wait_us(&fin, 0, 1000);
    return rc;
  }
  void finish(Entry &e, int64_t k, bool ok) {
    for (Word &w : e.done) {
      int pending = done_tag(k);
      w.v.compare_exchange_strong(pending, done_tag(k) | (ok ? 1 : 2));  // fails if the entry was recycled
    }
    for (Word &w : e.done) {
      if (two_level_)
        futex_wake_n(&w.v, 1);  // that sleeper wakes the rest of its lane (wait_word)
      else
        futex_wake_all(&w.v);
    }
  }
  void submitter_main() {
    for (;;) {
      Task t;
      {
        std::unique_lock<std::mutex> lk(q_mu_);
        // blocks arrive back to back while a stream runs: poll briefly before sleeping
        for (int spin = 0; tasks_.empty() && spin < 2000; spin++) {
          lk.unlock();
#if defined(__x86_64__) || defined(__i386__)
          __builtin_ia32_pause();
#endif
          lk.lock();
        }
        q_cv_.wait(lk, [this] { return !tasks_.empty(); });
        t = tasks_.front();
        tasks_.pop_front();
      }
      if (t.kind == Task::kStop) return;
      if (t.kind == Task::kSubmit) {
        Entry &e = log_[(size_t)(t.k % ring_)];
        int64_t ticket = -1;
        const uint64_t t0 = now_ns();
        const int rc = ops_.submit(ops_.ctx, e.fmt, e.host, e.elems, &ticket);
        ns_pub_submit_.fetch_add(now_ns() - t0, std::memory_order_relaxed);
        if (rc != 0) {
          finish(e, t.k, false);  // every member falls back to its private engine for this block
          continue;
        }
        e.ticket = ticket;
        {
          std::lock_guard<std::mutex> lk(w_mu_);
          waits_.push_back(Pending{t.k, ticket});
        }
        w_cv_.notify_one();
      } else if (t.kind == Task::kAdd) {
        *t.out_rc = ops_.add(ops_.ctx, t.filter, t.valid_history, t.out_client);
        t.finished->store(1, std::memory_order_release);
        futex_wake_all(t.finished);
      } else {
        ops_.remove(ops_.ctx, t.client);
      }
    }
  }
  void waiter_main() {
    for (;;) {
      Pending p;
      {
        std::unique_lock<std::mutex> lk(w_mu_);
        w_cv_.wait(lk, [this] { return !waits_.empty(); });
        p = waits_.front();
        waits_.pop_front();
      }
      if (p.k < 0) return;
      const uint64_t t0 = now_ns();
      const int rc = ops_.wait(ops_.ctx, p.ticket);
      ns_pub_wait_.fetch_add(now_ns() - t0, std::memory_order_relaxed);
      finish(log_[(size_t)(p.k % ring_)], p.k, rc == 0);
    }
  }

  const StreamOps ops_;
  const int ring_;
  const size_t max_bytes_;
  std::vector<Entry> log_;
  std::mutex mu_;  // serialises publishers and membership changes
  // the service threads and their queues
  std::mutex q_mu_, w_mu_;
  std::condition_variable q_cv_, w_cv_;
  std::deque<Task> tasks_;
  std::deque<Pending> waits_;
  std::thread submitter_, waiter_;
  std::atomic<int64_t> head_{-1};
  Word head_words_[kLanes];  // low bits of head_: the futex words followers of a publisher sleep on
  std::atomic<unsigned> next_lane_{0};
  bool two_level_ = false;  // XLATING_B200_WAKE=two_level (measurement switch)
  long spin_iters_ = 400;  // PAUSE iterations (~20 us) a caller polls for an event before it sleeps
  std::atomic<int> n_members_{0};
  std::atomic<uint64_t> published_{0}, desyncs_{0}, joins_{0}, private_matches_{0};
  std::atomic<uint64_t> ns_pub_copy_{0}, ns_pub_submit_{0}, ns_pub_wait_{0};
  // per-call counters are sharded by lane: one shared counter would be one more cache line that every
  // caller of every block bounces between the sockets
  struct alignas(64) Shard {
    std::atomic<uint64_t> hits{0}, ns_compare{0}, ns_wait{0};
  };
  Shard shards_[kLanes];
};

}  // namespace xl
#endif
