// CPU harness for sdr-server_b200/csrc/call_combiner.h (tests/test_call_combiner.py):
// the batch function is a stub that checks the combiner's guarantees and burns a
// little time, so the leader/follower synchronisation can be stressed without a GPU.
#include <stdlib.h>

#include <atomic>
#include <chrono>
#include <thread>
#include <vector>

#include "call_combiner.h"

namespace {
struct Ctx {
  std::atomic<long> violations{0}, served{0}, batches{0}, max_batch_seen{0};
  std::atomic<int> lane_busy[xl::CallCombiner::kMaxLanes];
  int max_batch = 0;
  int work_us = 0;
  int fail_every = 0;  // every n-th batch returns -5
};

struct Call {
  xl::CombinerCall cc;
  std::atomic<int> in_flight{0};
  long served = 0;
};

int stub_batch(void *p, int lane, xl::CombinerCall *const *batch, int n) {
  Ctx *ctx = (Ctx *)p;
  if (ctx->lane_busy[lane].fetch_add(1) != 0) ctx->violations++;  // lanes are exclusive
  if (n < 1 || n > ctx->max_batch) ctx->violations++;
  long m = ctx->max_batch_seen.load();
  while (n > m && !ctx->max_batch_seen.compare_exchange_weak(m, n)) {
  }
  for (int i = 0; i < n; i++) {
    Call *c = (Call *)batch[i]->user;
    if (c->in_flight.load() != 1) ctx->violations++;  // only calls that are actually waiting
    c->served++;
    for (int j = 0; j < i; j++)
      if (batch[j] == batch[i]) ctx->violations++;  // once per batch
  }
  if (ctx->work_us > 0) {
    const auto t0 = std::chrono::steady_clock::now();
    while (std::chrono::steady_clock::now() - t0 < std::chrono::microseconds(ctx->work_us)) {
    }
  }
  ctx->served += n;
  const long b = ctx->batches.fetch_add(1) + 1;
  ctx->lane_busy[lane].fetch_sub(1);
  return (ctx->fail_every > 0 && b % ctx->fail_every == 0) ? -5 : 0;
}
}  // namespace

extern "C" {

// n_threads callers (one call object each, like one filter per dsp thread) make n_calls
// calls each.  Returns the number of violations; outputs the batch statistics.
long cc_stress(int n_threads, int n_calls, int lanes, int max_batch, int work_us, int fail_every, long *batches,
               long *served, long *max_batch_seen, long *failed_calls) {
  Ctx ctx;
  for (auto &b : ctx.lane_busy) b = 0;
  ctx.max_batch = max_batch;
  ctx.work_us = work_us;
  ctx.fail_every = fail_every;
  xl::CallCombiner comb(lanes, max_batch, stub_batch, &ctx);
  std::vector<Call *> calls;
  for (int t = 0; t < n_threads; t++) {
    calls.push_back(new Call());
    calls.back()->cc.user = calls.back();
  }
  std::atomic<long> failed{0};
  std::vector<std::thread> th;
  for (int t = 0; t < n_threads; t++)
    th.emplace_back([&, t]() {
      Call *c = calls[(size_t)t];
      for (int i = 0; i < n_calls; i++) {
        c->in_flight = 1;
        const int rc = comb.run(&c->cc);
        c->in_flight = 0;
        if (rc != 0) failed++;
        if (c->served != i + 1) ctx.violations++;  // served exactly once per call, before run() returned
        if ((t + i) % 17 == 0) std::this_thread::yield();
      }
    });
  for (auto &x : th) x.join();
  uint64_t b = 0, c = 0;
  comb.stats(&b, &c);
  if ((long)b != ctx.batches.load() || (long)c != ctx.served.load()) ctx.violations++;
  *batches = (long)b;
  *served = (long)c;
  *max_batch_seen = ctx.max_batch_seen.load();
  *failed_calls = failed.load();
  for (Call *c2 : calls) delete c2;
  return ctx.violations.load();
}
}
