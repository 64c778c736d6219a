// CPU harness for sdr-server_b200/csrc/block_cache.h (tests/test_block_cache.py): the
// "device" twin is a second malloc buffer and "upload" is a memcpy, so that the
// sharing / eviction / reference logic can be hammered without a GPU.
#include <stdlib.h>

#include <atomic>
#include <thread>
#include <vector>

#include "block_cache.h"

namespace {
std::atomic<long> g_allocs{0}, g_uploads{0};
int shim_alloc(void *, size_t bytes, void **host, void **dev) {
  *host = malloc(bytes);
  *dev = malloc(bytes);
  g_allocs++;
  return (*host && *dev) ? 0 : -1;
}
void shim_release(void *, void *host, void *dev) {
  free(host);
  free(dev);
}
int shim_upload(void *ctx, int, const void *host, void *dev, size_t bytes) {
  if (ctx != nullptr && ((std::atomic<int> *)ctx)->load() != 0) return -1;  // injected failure
  memcpy(dev, host, bytes);
  g_uploads++;
  return 0;
}
}  // namespace

extern "C" {

void *bc_new(void *fail_flag) {
  xl::BlockCacheOps ops = {shim_alloc, shim_release, shim_upload, fail_flag};
  return new xl::BlockCache(ops);
}
void bc_delete(void *c) { delete (xl::BlockCache *)c; }
int bc_acquire(void *c, const void *p, size_t n) { return ((xl::BlockCache *)c)->acquire(p, n); }
void bc_release(void *c, int slot) { ((xl::BlockCache *)c)->release(slot); }
int bc_matches(void *c, int slot, const void *p, size_t n) {
  return memcmp(((xl::BlockCache *)c)->device_ptr(slot), p, n) == 0;
}
void bc_stats(void *c, uint64_t *hits, uint64_t *publishes) { ((xl::BlockCache *)c)->stats(hits, publishes); }
int bc_referenced(void *c) { return ((xl::BlockCache *)c)->referenced(); }
long bc_uploads() { return g_uploads.load(); }
int bc_slots() { return xl::BlockCache::kSlots; }
uint64_t bc_key(const void *p, size_t n) { return xl::block_key(p, n); }

// n_threads "dsp threads" walk the same sequence of blocks (private copies of block
// b = pool[b % pool]), some of them lagging behind; every acquired entry must hold
// exactly the caller's bytes for as long as it is referenced.  Returns the number
// of violations; *shared / *priv count the outcomes.
long bc_stress(void *c, int n_threads, int n_iters, int pool, size_t bytes, long *shared, long *priv) {
  xl::BlockCache *cache = (xl::BlockCache *)c;
  std::vector<std::vector<unsigned char>> blocks((size_t)pool, std::vector<unsigned char>(bytes));
  uint64_t s = 88172645463325252ull;
  for (auto &b : blocks)
    for (auto &v : b) {
      s ^= s << 13;
      s ^= s >> 7;
      s ^= s << 17;
      v = (unsigned char)s;
    }
  std::atomic<long> bad{0}, n_shared{0}, n_priv{0};
  std::vector<std::thread> th;
  for (int t = 0; t < n_threads; t++)
    th.emplace_back([&, t]() {
      std::vector<unsigned char> own(bytes);
      for (int i = 0; i < n_iters; i++) {
        const int b = (i + (t % 3 == 2 ? 1 : 0)) % pool;  // a third of the threads is one block ahead
        own = blocks[(size_t)b];
        if (t % 7 == 3) own[bytes / 2] ^= (unsigned char)(1 + i % 255);  // a client with different data
        const int slot = cache->acquire(own.data(), bytes);
        if (slot < 0) {
          n_priv++;
          continue;
        }
        n_shared++;
        for (int rep = 0; rep < 3; rep++) {
          if (memcmp(cache->device_ptr(slot), own.data(), bytes) != 0) bad++;
          if (memcmp(cache->host_ptr(slot), own.data(), bytes) != 0) bad++;
          std::this_thread::yield();
        }
        cache->release(slot);
      }
    });
  for (auto &x : th) x.join();
  *shared = n_shared.load();
  *priv = n_priv.load();
  return bad.load();
}
}
