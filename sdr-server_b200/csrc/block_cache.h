/*
 * csrc/block_cache.h -- content-addressed cache of recently submitted input blocks,
 * host logic only (no CUDA in here: the owner supplies the allocation and upload
 * callbacks, so tests/test_block_cache.py can exercise it on a CPU).
 *
 * Why: the reference hands every client a private memcpy of the SAME SDR block
 * (src/tcp_server.c:262-269 -> src/queue.c:114) and every dsp thread then calls
 * process_* on its copy.  Behind the per-filter ABI the library is not told that
 * those C inputs are identical -- but it can find out: a caller computes a key from
 * 16 sampled 32-byte pieces of its input (constant time), looks the key up here, and
 * byte-compares against every published block with that key (memcmp: exact, and it
 * stops at the first differing byte, so a key collision costs almost nothing).  On a
 * match it shares that block.  The block then crosses PCIe once instead of C times,
 * and the C-1 staging memcpys are replaced by C-1 memcmps -- one pass over the input
 * per call, as before.
 *
 * Entries are reference counted by the calls that use them and recycled LRU once
 * unreferenced.  Callers that find no free entry fall back to their private path.
 */
#ifndef XLATING_B200_BLOCK_CACHE_H_
#define XLATING_B200_BLOCK_CACHE_H_

#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <thread>

namespace xl {

struct BlockCacheOps {
  // allocate a host buffer (pinned, in production) and its device twin; 0 on success
  int (*alloc)(void *ctx, size_t bytes, void **host, void **dev);
  void (*release)(void *ctx, void *host, void *dev);
  // start the host -> device transfer of entry `slot` and make it waitable; 0 on success
  int (*upload)(void *ctx, int slot, const void *host, void *dev, size_t bytes);
  void *ctx;
};

// Key of a block: its length and 16 pieces of 32 bytes spread evenly over it (all of
// it when shorter than 512 bytes).  Only a filter in front of memcmp -- blocks that
// differ elsewhere get the same key and are told apart by the comparison.
inline uint64_t block_key(const void *data, size_t n) {
  const unsigned char *p = (const unsigned char *)data;
  const uint64_t k = 0xFF51AFD7ED558CCDull;
  uint64_t h[4] = {0x9E3779B97F4A7C15ull ^ n, 0xC2B2AE3D27D4EB4Full, 0x165667B19E3779F9ull, 0x27D4EB2F165667C5ull};
  auto mix32 = [&](const unsigned char *q) {
    uint64_t w[4];
    memcpy(w, q, 32);
    for (int l = 0; l < 4; l++) {
      h[l] = (h[l] ^ w[l]) * k;
      h[l] ^= h[l] >> 29;
    }
  };
  if (n >= 512) {
    for (int i = 0; i < 16; i++) mix32(p + (size_t)i * (n - 32) / 15);
  } else {
    size_t i = 0;
    for (; i + 32 <= n; i += 32) mix32(p + i);
    unsigned char tail[32] = {0};
    memcpy(tail, p + i, n - i);
    mix32(tail);
  }
  uint64_t r = h[0];
  for (int l = 1; l < 4; l++) r = (r ^ h[l]) * 0xC4CEB9FE1A85EC53ull + l;
  return r ^ (r >> 31);
}

class BlockCache {
 public:
  static constexpr int kSlots = 32;    // dsp threads drift apart by a few blocks; 32 x one block of memory is nothing
  static constexpr int kPrivate = -1;  // acquire(): no shared entry, use the caller's own path
  static_assert(kSlots <= 32, "acquire() keeps rejected entries in a 32-bit mask");

  explicit BlockCache(const BlockCacheOps &ops) : ops_(ops) {}
  ~BlockCache() {
    for (Entry &e : slots_)
      if (e.host != nullptr) ops_.release(ops_.ctx, e.host, e.dev);
  }
  BlockCache(const BlockCache &) = delete;
  BlockCache &operator=(const BlockCache &) = delete;

  // Returns the slot of an entry holding exactly input[0, bytes) -- found or newly
  // published -- with one reference taken, or kPrivate.
  //
  // The hit path takes no lock: when an SDR block lands, every dsp thread but one
  // arrives here within microseconds looking for the same entry, and a mutex (or a
  // condition-variable broadcast) would hand them through one at a time.  A reader
  // pins an entry with refs++ and then re-checks that it is still the entry it
  // wanted; the recycler closes an entry (state = filling) and then checks refs == 0.
  // Both sides use sequentially consistent operations, so at least one of them sees
  // the other (Dekker) and an entry is never recycled under a reader.
  int acquire(const void *input, size_t bytes) {
    const uint64_t hash = block_key(input, bytes);
    uint32_t rejected = 0;  // entries with this key whose bytes turned out to differ
    for (;;) {
      int filling = -1;
      bool raced = false;
      for (int i = 0; i < kSlots && !raced; i++) {
        Entry &e = slots_[i];
        const int st = e.state.load(std::memory_order_acquire);
        if (st == kEmpty || (rejected & (1u << i)) || e.hash.load(std::memory_order_relaxed) != hash ||
            e.bytes.load(std::memory_order_relaxed) != bytes)
          continue;
        if (st == kFilling) {  // another caller is publishing these bytes right now
          filling = i;
          continue;
        }
        e.refs.fetch_add(1);
        if (e.state.load() == kReady && e.hash.load() == hash && e.bytes.load() == bytes) {
          // pinned: host/dev/bytes cannot change while refs > 0
          e.stamp.store(clock_.fetch_add(1) + 1, std::memory_order_relaxed);
          if (memcmp(e.host, input, bytes) == 0) {
            hits_.fetch_add(1, std::memory_order_relaxed);
            return i;
          }
          e.refs.fetch_sub(1);  // same key, different bytes: another entry may still match
          rejected |= 1u << i;
          continue;
        }
        e.refs.fetch_sub(1);  // the entry was being recycled: look again
        raced = true;
      }
      if (raced) continue;
      if (filling >= 0) {
        // ~20 us (one memcpy + one async copy launch by the publisher); yield, do not sleep
        while (slots_[filling].state.load(std::memory_order_acquire) == kFilling) std::this_thread::yield();
        continue;
      }
      // ---- publish ----
      Entry *v = nullptr;
      {
        std::lock_guard<std::mutex> lk(mu_);  // one publisher at a time
        bool appeared = false;
        for (int i = 0; i < kSlots; i++) {
          Entry &e = slots_[i];
          if (!(rejected & (1u << i)) && e.state.load() != kEmpty && e.hash.load() == hash && e.bytes.load() == bytes)
            appeared = true;
        }
        if (appeared) continue;  // another publisher got there first: take the hit path
        for (Entry &e : slots_)
          if (e.state.load() == kEmpty) {
            v = &e;
            v->state.store(kFilling);
            break;
          }
        while (v == nullptr) {
          // least recently used unreferenced entry; close it, then make sure no reader got in
          Entry *cand = nullptr;
          for (Entry &e : slots_)
            if (e.state.load() == kReady && e.refs.load() == 0 && !e.skip &&
                (cand == nullptr || e.stamp.load() < cand->stamp.load()))
              cand = &e;
          if (cand == nullptr) break;
          cand->state.store(kFilling);
          if (cand->refs.load() == 0) {
            v = cand;
          } else {
            cand->state.store(kReady);
            cand->skip = true;  // busy after all: not this round
          }
        }
        for (Entry &e : slots_) e.skip = false;
        if (v == nullptr) return kPrivate;
        // v is closed (kFilling): readers of its old contents back off, nobody else writes it
        if (v->cap < bytes) {
          if (v->host != nullptr) ops_.release(ops_.ctx, v->host, v->dev);
          v->host = v->dev = nullptr;
          v->cap = 0;
          if (ops_.alloc(ops_.ctx, bytes, &v->host, &v->dev) != 0) {
            v->hash.store(0);
            v->bytes.store(0);
            v->state.store(kEmpty);
            return kPrivate;
          }
          v->cap = bytes;
        }
        v->hash.store(hash);
        v->bytes.store(bytes);
        v->refs.fetch_add(1);  // not store: late readers of the old contents may still be backing off
        v->stamp.store(clock_.fetch_add(1) + 1);
      }
      memcpy(v->host, input, bytes);
      const int slot = (int)(v - slots_);
      if (ops_.upload(ops_.ctx, slot, v->host, v->dev, bytes) != 0) {
        v->refs.fetch_sub(1);
        v->hash.store(0);
        v->bytes.store(0);
        v->state.store(kEmpty);
        return kPrivate;
      }
      publishes_.fetch_add(1, std::memory_order_relaxed);
      v->state.store(kReady);
      return slot;
    }
  }

  void release(int slot) { slots_[slot].refs.fetch_sub(1); }

  // valid while the caller holds a reference
  const void *device_ptr(int slot) const { return slots_[slot].dev; }
  const void *host_ptr(int slot) const { return slots_[slot].host; }

  void stats(uint64_t *hits, uint64_t *publishes) {
    *hits = hits_.load();
    *publishes = publishes_.load();
  }
  int referenced() {
    int n = 0;
    for (Entry &e : slots_) n += e.refs.load();
    return n;
  }

 private:
  enum { kEmpty = 0, kFilling = 1, kReady = 2 };
  struct Entry {
    void *host = nullptr, *dev = nullptr;  // written only by the publisher that closed the entry
    size_t cap = 0;
    bool skip = false;                     // publisher-private (under mu_)
    std::atomic<uint64_t> hash{0}, stamp{0};
    std::atomic<size_t> bytes{0};
    std::atomic<int> state{kEmpty}, refs{0};
  };
  BlockCacheOps ops_;
  std::mutex mu_;  // serialises publishers (victim choice, reallocation)
  Entry slots_[kSlots];
  std::atomic<uint64_t> clock_{0}, hits_{0}, publishes_{0};
};

}  // namespace xl
#endif
