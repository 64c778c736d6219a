// tests/stream_overlay_shim.cpp -- model check of sdr-server_b200/csrc/stream_overlay.h on a CPU.
//
// The overlay's two back ends are replaced by the ORACLE (oracle/liboracle.so):
//   * the "group" is a set of oracle filters, one per member, fed every published block in
//     publication order; a member added with state starts from the group's OWN record of the
//     stream (samples older than valid_history read as zero), like xlg_add_client_ex;
//   * the "private engine" of a filter is an oracle filter that is re-seeded from the
//     filter's mirror (history tail, history_offset, oscillator) whenever the filter comes
//     back from the group, exactly what csrc/xlating_dropin.cu has to do.
// Every filter also owns an INDEPENDENT oracle filter that simply consumes the filter's own
// input sequence.  Whatever the threads, drops, lags, late joiners or second sources do, each
// call's output must equal that filter's bit for bit (all arithmetic is the oracle's).
//
// usage: stream_overlay_model <scenario> <seed>      prints one JSON line, exit 0 = all equal
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <atomic>
#include <mutex>
#include <random>
#include <thread>
#include <vector>

#include "oracle.h"
#include "stream_overlay.h"

using namespace xl;

namespace {

constexpr uint32_t kFs = 48000, kD = 5;
constexpr uint32_t kMaxIn = 8192;  // scalar elements per block at most
std::vector<float> g_taps;

float cvt_cu8(uint8_t u) { return ((float)u - 127.5f) / 128.0f; }

struct FakeGroup {
  struct Client {
    orc_xlating *f = nullptr;
    bool active = false;
  };
  struct Result {
    int64_t ticket = -1;
    std::vector<std::vector<float>> out;           // per client: interleaved re,im
    std::vector<size_t> hist_after;
    std::vector<float> ph_re, ph_im;
  };
  std::vector<Client> clients;
  std::vector<Result> ring;
  std::vector<float> stream;  // every submitted sample, interleaved re,im (the group's ring, unbounded here)
  int64_t next_ticket = 0;
  std::mutex mu;  // results ring vs readers (the real group guards its entries the same way)
  std::atomic<int> wait_jitter_us{0};
  explicit FakeGroup(int r) : ring((size_t)r) {}
  ~FakeGroup() {
    for (Client &c : clients)
      if (c.f) orc_xlating_destroy(c.f);
  }
};

struct MFilter {
  int id = 0;
  int32_t center = 0;
  orc_xlating *priv = nullptr;   // the "per-filter engine"
  orc_xlating *check = nullptr;  // independent truth
  AutoStream::Member m;
  // mirror of the state after the last consumed block
  size_t hist = 0;
  float ph_re = 1.f, ph_im = 0.f;
  std::vector<float> tail;  // last T-1 samples consumed, interleaved, zero-padded at the front
  int64_t total = 0;
  bool priv_stale = false;
  uint64_t served_by_group = 0, served_privately = 0, mismatches = 0;
};

FakeGroup *g_group = nullptr;

void *op_alloc(void *, size_t bytes) { return malloc(bytes); }
void op_free(void *, void *p) { free(p); }

int op_submit(void *, int fmt, const void *block, size_t elems, int64_t *ticket) {
  FakeGroup &g = *g_group;
  (void)fmt;
  const uint8_t *u = (const uint8_t *)block;
  for (size_t i = 0; i < elems; i++) g.stream.push_back(cvt_cu8(u[i]));
  FakeGroup::Result r;
  r.ticket = g.next_ticket++;
  const size_t nc = g.clients.size();
  r.out.resize(nc);
  r.hist_after.assign(nc, 0);
  r.ph_re.assign(nc, 1.f);
  r.ph_im.assign(nc, 0.f);
  for (size_t c = 0; c < nc; c++) {
    if (!g.clients[c].active) continue;
    const float *o = nullptr;
    const size_t n = orc_xlating_process_cf32(g.clients[c].f, ORC_FMT_CU8, block, elems, 1, &o);
    r.out[c].assign(o, o + 2 * n);
    r.hist_after[c] = orc_xlating_history(g.clients[c].f);
    orc_xlating_phase(g.clients[c].f, &r.ph_re[c], &r.ph_im[c]);
  }
  {
    std::lock_guard<std::mutex> lk(g.mu);
    g.ring[(size_t)(r.ticket % (int64_t)g.ring.size())] = std::move(r);
  }
  *ticket = g.next_ticket - 1;
  return 0;
}

int op_wait(void *, int64_t) {
  const int j = g_group->wait_jitter_us.load();
  if (j > 0) usleep((useconds_t)(rand() % j));
  return 0;
}

int op_add(void *, void *filter, int64_t valid_history, int *client) {
  FakeGroup &g = *g_group;
  MFilter *f = (MFilter *)filter;
  FakeGroup::Client c;
  if (orc_xlating_create(kD, g_taps.data(), g_taps.size(), f->center, kFs, kMaxIn, &c.f) != 0) return -1;
  // history from the GROUP's stream; older than valid_history reads as zero
  std::vector<float> h(2 * f->hist, 0.f);
  const int64_t have = (int64_t)(g.stream.size() / 2);
  for (int64_t i = 0; i < (int64_t)f->hist; i++) {  // i samples back from the end: 0 = newest
    if (i < valid_history && i < have) {
      h[2 * (f->hist - 1 - (size_t)i)] = g.stream[2 * (size_t)(have - 1 - i)];
      h[2 * (f->hist - 1 - (size_t)i) + 1] = g.stream[2 * (size_t)(have - 1 - i) + 1];
    }
  }
  orc_xlating_set_state(c.f, h.data(), f->hist, f->ph_re, f->ph_im);
  c.active = true;
  g.clients.push_back(c);
  *client = (int)g.clients.size() - 1;
  return 0;
}

int op_remove(void *, int client) {
  g_group->clients[(size_t)client].active = false;
  return 0;
}

// what xlating_dropin.cu does per call, with the oracle in place of the GPU
void update_tail(MFilter &f, const uint8_t *in, size_t elems) {
  const size_t T1 = f.tail.size() / 2, n = elems / 2;
  if (n >= T1) {
    for (size_t i = 0; i < 2 * T1; i++) f.tail[i] = cvt_cu8(in[2 * (n - T1) + i]);
  } else {
    memmove(f.tail.data(), f.tail.data() + 2 * n, sizeof(float) * 2 * (T1 - n));
    for (size_t i = 0; i < 2 * n; i++) f.tail[2 * (T1 - n) + i] = cvt_cu8(in[i]);
  }
}

bool filter_call(AutoStream &as, MFilter &f, const uint8_t *in, size_t elems) {
  const float *want = nullptr;
  const size_t n_want = orc_xlating_process_cf32(f.check, ORC_FMT_CU8, in, elems, 1, &want);
  std::vector<float> got;
  bool served = false;
  if (f.m.member) {
    AutoStream::Served sv;
    if (as.member_call(f.m, in, elems, 0, elems, &sv) == 1) {
      FakeGroup &g = *g_group;
      std::lock_guard<std::mutex> lk(g.mu);
      FakeGroup::Result &r = g.ring[(size_t)(sv.ticket % (int64_t)g.ring.size())];
      if (r.ticket == sv.ticket && (size_t)sv.client < r.out.size()) {
        got = r.out[(size_t)sv.client];
        f.hist = r.hist_after[(size_t)sv.client];
        f.ph_re = r.ph_re[(size_t)sv.client];
        f.ph_im = r.ph_im[(size_t)sv.client];
        served = true;
        f.priv_stale = true;
        f.served_by_group++;
      } else {
        as.leave(f.m);  // result ring recycled under us: as if desynced, mirror still valid
      }
    }
  }
  if (!served) {
    const bool seen = as.private_observe(f.m, in, elems, 0, elems);
    if (f.priv_stale) {
      orc_xlating_set_state(f.priv, f.tail.data() + (f.tail.size() - 2 * f.hist), f.hist, f.ph_re, f.ph_im);
      f.priv_stale = false;
    }
    const float *o = nullptr;
    const size_t n = orc_xlating_process_cf32(f.priv, ORC_FMT_CU8, in, elems, 1, &o);
    got.assign(o, o + 2 * n);
    f.hist = orc_xlating_history(f.priv);
    orc_xlating_phase(f.priv, &f.ph_re, &f.ph_im);
    f.served_privately++;
    if (!seen) as.private_observe(f.m, in, elems, 0, elems, /*retry=*/true);
  }
  update_tail(f, in, elems);
  f.total += (int64_t)(elems / 2);
  if (!served) as.try_join(f.m, (int64_t)g_taps.size() - 1, f.total, &f);
  const bool ok = got.size() == 2 * n_want && (n_want == 0 || memcmp(got.data(), want, sizeof(float) * 2 * n_want) == 0);
  if (!ok) f.mismatches++;
  return ok;
}

MFilter *make_filter(int id, int32_t center) {
  MFilter *f = new MFilter();
  f->id = id;
  f->center = center;
  orc_xlating_create(kD, g_taps.data(), g_taps.size(), center, kFs, kMaxIn, &f->priv);
  orc_xlating_create(kD, g_taps.data(), g_taps.size(), center, kFs, kMaxIn, &f->check);
  f->hist = g_taps.size() - 1;
  f->tail.assign(2 * (g_taps.size() - 1), 0.f);
  return f;
}

}  // namespace

int main(int argc, char **argv) {
  const char *scenario = argc > 1 ? argv[1] : "steady";
  const unsigned seed = argc > 2 ? (unsigned)atoi(argv[2]) : 1;
  float *taps = nullptr;
  size_t ntaps = 0;
  if (orc_lpf_design(1.0f, kFs, 4800, 2000, &taps, &ntaps) != 0) return 2;  // 57 taps (test_xlating.c shape)
  g_taps.assign(taps, taps + ntaps);
  free(taps);

  const bool drops = strcmp(scenario, "drops") == 0, lag = strcmp(scenario, "lag") == 0;
  const bool two = strcmp(scenario, "two_sources") == 0, late = strcmp(scenario, "late_joiners") == 0;
  const bool ragged = strcmp(scenario, "ragged") == 0;
  const int ring = lag ? 6 : 16;
  const int n_filters = 12, n_blocks = lag ? 120 : 80;
  FakeGroup group(ring);
  g_group = &group;
  group.wait_jitter_us = 30;
  StreamOps ops = {nullptr, op_alloc, op_free, op_submit, op_wait, op_add, op_remove};
  AutoStream as(ops, ring, kMaxIn);

  // the SDR streams: source 0 (and 1 for two_sources); block sizes vary in "ragged"
  std::mt19937 rng(seed);
  std::vector<std::vector<std::vector<uint8_t>>> src(2);
  for (int s = 0; s < 2; s++)
    for (int b = 0; b < n_blocks; b++) {
      size_t elems = 4096;
      if (ragged) {
        const size_t choices[] = {4096, 2, 38, 1234, 8192, 60, 4096, 200};
        elems = choices[rng() % 8];
      }
      std::vector<uint8_t> blk(elems);
      for (auto &v : blk) v = (uint8_t)rng();
      src[(size_t)s].push_back(blk);
    }

  std::vector<MFilter *> filters;
  for (int i = 0; i < n_filters; i++) filters.push_back(make_filter(i, -12000 + 997 * i));
  std::atomic<int> bad{0};
  // a barrier every `window` blocks keeps the threads within one reference queue depth of each
  // other (src/config.c:183), like the bounded queues of the real server
  const int window = lag ? 200 : 8;
  pthread_barrier_t bar;
  pthread_barrier_init(&bar, nullptr, (unsigned)n_filters);
  auto worker = [&](int i) {
    MFilter &f = *filters[(size_t)i];
    std::mt19937 r(seed * 7919u + (unsigned)i);
    const int s = two ? (i & 1) : 0;
    const int start = late ? (i % 4) * 10 : 0;  // late joiners: created "later" = they skip the first blocks
    for (int b = 0; b < n_blocks; b++) {
      if (b % window == 0) pthread_barrier_wait(&bar);
      if (b < start) continue;
      if (drops && i % 3 == 0 && b > 5 && r() % 11 == 0) continue;  // the client's queue dropped this block
      if (lag && i == 3 && b == 20) usleep(400000);                 // one client stalls for a long time
      if (r() % 4 == 0) usleep((useconds_t)(r() % 200));
      const std::vector<uint8_t> &blk = src[(size_t)s][(size_t)b];
      std::vector<uint8_t> own(blk);  // private copy, like queue_put's memcpy (src/queue.c:114)
      if (!filter_call(as, f, own.data(), own.size())) bad++;
    }
  };
  std::vector<std::thread> th;
  for (int i = 0; i < n_filters; i++) th.emplace_back(worker, i);
  for (auto &t : th) t.join();

  uint64_t by_group = 0, privately = 0;
  for (MFilter *f : filters) {
    by_group += f->served_by_group;
    privately += f->served_privately;
  }
  const AutoStream::Stats st = as.stats();
  printf("{\"scenario\": \"%s\", \"seed\": %u, \"mismatching_calls\": %d, \"served_by_group\": %llu, "
         "\"served_privately\": %llu, \"published\": %llu, \"hits\": %llu, \"desyncs\": %llu, \"joins\": %llu, "
         "\"private_matches\": %llu}\n",
         scenario, seed, bad.load(), (unsigned long long)by_group, (unsigned long long)privately,
         (unsigned long long)st.published, (unsigned long long)st.hits, (unsigned long long)st.desyncs,
         (unsigned long long)st.joins, (unsigned long long)st.private_matches);
  for (MFilter *f : filters) {
    as.leave(f->m);
    orc_xlating_destroy(f->priv);
    orc_xlating_destroy(f->check);
    delete f;
  }
  return bad.load() == 0 ? 0 : 1;
}
