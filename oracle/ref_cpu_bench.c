/*
 * oracle/ref_cpu_bench.c -- TEST / BASELINE INFRASTRUCTURE.  Times the reference's own
 * CPU implementation of the hot path (src/xlating.c + src/lpf.c, compiled UNMODIFIED
 * where they lie by oracle/Makefile, one binary per flag set) the way BASELINE.md
 * section 3 asks for: one pinned pthread per client (min(C, allowed CPUs) threads,
 * clients dealt round-robin), each client its own filter, all threads reading one
 * shared input block -- i.e. src/dsp_worker.c:41-88 without the socket/file write and
 * without the per-client queue memcpy (src/queue.c:114), which favours the CPU.
 * CLOCK_MONOTONIC wall time, not clock().
 *
 * usage: ref_cpu_bench <fs> <cu8|cs8|cs16> <block_elems> <blocks> <warmup_blocks>
 *                      <native|optimized> [max_threads [pin|nopin [main|local]]]   < plan
 *   pin|nopin   one thread per allowed CPU, pinned (default) / left to the scheduler
 *   main|local  filters created by the main thread, as the reference's acceptor thread
 *               does (src/tcp_server.c:327; default) / by the thread that runs them, which
 *               puts their working buffers on that thread's NUMA node (favours the CPU)
 * plan (stdin): one client per line "decimation cutoff transition_width center_offset"
 *
 * Prints one JSON line.  bench.py --impl reference and the cpu_baseline leg run it.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <sched.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "lpf.h"
#include "xlating.h"

extern const char *SIMD_STATUS;

typedef struct {
  uint32_t decimation, cutoff, tw;
  int32_t center;
  xlating *filter;
} client_t;

static size_t g_block_elems;
static size_t g_block_elems_for_create(void) { return g_block_elems; }
static client_t *g_clients;
static int g_n_clients, g_n_threads, g_blocks, g_warmup, g_fmt, g_optimized, g_local;
static uint32_t g_fs;
static size_t g_taps_min = (size_t)-1, g_taps_max = 0;
static pthread_mutex_t g_mu = PTHREAD_MUTEX_INITIALIZER;

static int create_client(client_t *c) {
  float *taps = NULL;
  size_t len = 0;
  if (create_low_pass_filter(1.0f, g_fs, c->cutoff, c->tw, &taps, &len) != 0) return 1;
  pthread_mutex_lock(&g_mu);
  if (len < g_taps_min) g_taps_min = len;
  if (len > g_taps_max) g_taps_max = len;
  pthread_mutex_unlock(&g_mu);
  return create_frequency_xlating_filter(c->decimation, taps, len, c->center, g_fs, (uint32_t)g_block_elems_for_create(),
                                         &c->filter);
}
static uint8_t *g_data[8];
static pthread_barrier_t g_barrier;
static struct timespec g_t0, g_t1;
static uint64_t g_outputs[1024];

static void process(client_t *c, const uint8_t *blk, size_t *n_out) {
  float complex *out = NULL;
  if (g_fmt == 0) {
    (g_optimized ? process_optimized_cu8_cf32 : process_native_cu8_cf32)(blk, g_block_elems, &out, n_out, c->filter);
  } else if (g_fmt == 1) {
    (g_optimized ? process_optimized_cs8_cf32 : process_native_cs8_cf32)((const int8_t *)blk, g_block_elems, &out, n_out,
                                                                        c->filter);
  } else {
    (g_optimized ? process_optimized_cs16_cf32 : process_native_cs16_cf32)((const int16_t *)blk, g_block_elems, &out,
                                                                          n_out, c->filter);
  }
}

static void *worker(void *arg) {
  const int tid = (int)(intptr_t)arg;
  uint64_t outputs = 0;
  if (g_local)
    for (int c = tid; c < g_n_clients; c += g_n_threads)
      if (create_client(&g_clients[c]) != 0) exit(1);
  for (int b = 0; b < g_warmup + g_blocks; b++) {
    if (b == g_warmup) {
      pthread_barrier_wait(&g_barrier);
      if (tid == 0) clock_gettime(CLOCK_MONOTONIC, &g_t0);
    }
    const uint8_t *blk = g_data[b % 8];
    for (int c = tid; c < g_n_clients; c += g_n_threads) {
      size_t n = 0;
      process(&g_clients[c], blk, &n);
      outputs += n;
    }
  }
  pthread_barrier_wait(&g_barrier);
  if (tid == 0) clock_gettime(CLOCK_MONOTONIC, &g_t1);
  g_outputs[tid % 1024] += outputs;
  return NULL;
}

int main(int argc, char **argv) {
  if (argc < 7) {
    fprintf(stderr, "usage: %s fs fmt block_elems blocks warmup native|optimized [max_threads] < plan\n", argv[0]);
    return 2;
  }
  g_fs = (uint32_t)strtoul(argv[1], NULL, 10);
  g_fmt = strcmp(argv[2], "cu8") == 0 ? 0 : (strcmp(argv[2], "cs8") == 0 ? 1 : 2);
  g_block_elems = (size_t)strtoul(argv[3], NULL, 10);
  g_blocks = atoi(argv[4]);
  g_warmup = atoi(argv[5]);
  g_optimized = strcmp(argv[6], "optimized") == 0;
  int max_threads = argc > 7 ? atoi(argv[7]) : 0;
  const int pin = !(argc > 8 && strcmp(argv[8], "nopin") == 0);
  g_local = argc > 9 && strcmp(argv[9], "local") == 0;

  int cap = 64;
  g_clients = (client_t *)calloc((size_t)cap, sizeof(client_t));
  for (;;) {
    client_t c;
    memset(&c, 0, sizeof(c));
    if (scanf("%u %u %u %d", &c.decimation, &c.cutoff, &c.tw, &c.center) != 4) break;
    if (g_n_clients == cap) {
      cap *= 2;
      g_clients = (client_t *)realloc(g_clients, (size_t)cap * sizeof(client_t));
    }
    g_clients[g_n_clients++] = c;
  }
  if (g_n_clients == 0) {
    fprintf(stderr, "empty plan\n");
    return 2;
  }
  if (!g_local)
    for (int i = 0; i < g_n_clients; i++)
      if (create_client(&g_clients[i]) != 0) return 1;
  /* synthetic blocks: xorshift64 bytes (cs16: 14-bit samples), the same in every build */
  const size_t bytes = g_block_elems * (g_fmt == 2 ? 2 : 1);
  uint64_t s = 0x9E3779B97F4A7C15ull;
  for (int b = 0; b < 8; b++) {
    g_data[b] = (uint8_t *)malloc(bytes);
    for (size_t i = 0; i < bytes; i++) {
      s ^= s << 13;
      s ^= s >> 7;
      s ^= s << 17;
      g_data[b][i] = (uint8_t)s;
    }
    if (g_fmt == 2) {
      int16_t *w = (int16_t *)g_data[b];
      for (size_t i = 0; i < g_block_elems; i++) w[i] = (int16_t)(w[i] >> 2);
    }
  }
  /* one thread per allowed CPU, at most one per client */
  cpu_set_t allowed;
  CPU_ZERO(&allowed);
  sched_getaffinity(0, sizeof(allowed), &allowed);
  int cpus[CPU_SETSIZE], n_cpus = 0;
  for (int i = 0; i < CPU_SETSIZE; i++)
    if (CPU_ISSET(i, &allowed)) cpus[n_cpus++] = i;
  g_n_threads = n_cpus < g_n_clients ? n_cpus : g_n_clients;
  if (max_threads > 0 && g_n_threads > max_threads) g_n_threads = max_threads;
  pthread_barrier_init(&g_barrier, NULL, (unsigned)g_n_threads);
  pthread_t *th = (pthread_t *)calloc((size_t)g_n_threads, sizeof(pthread_t));
  for (int t = 0; t < g_n_threads; t++) {
    pthread_attr_t attr;
    pthread_attr_init(&attr);
    cpu_set_t one;
    CPU_ZERO(&one);
    CPU_SET(cpus[t % n_cpus], &one);
    if (pin) pthread_attr_setaffinity_np(&attr, sizeof(one), &one);
    pthread_create(&th[t], &attr, worker, (void *)(intptr_t)t);
    pthread_attr_destroy(&attr);
  }
  for (int t = 0; t < g_n_threads; t++) pthread_join(th[t], NULL);
  const double dt = (double)(g_t1.tv_sec - g_t0.tv_sec) + 1e-9 * (double)(g_t1.tv_nsec - g_t0.tv_nsec);
  uint64_t outputs = 0;
  for (int t = 0; t < 1024; t++) outputs += g_outputs[t];
  const double in_samples = (double)g_blocks * (double)(g_block_elems / 2);
  printf("{\"bench\": \"ref_cpu_bench\", \"simd_status\": \"%s\", \"variant\": \"%s\", \"clients\": %d, \"threads\": %d, "
         "\"cpus_allowed\": %d, \"pinned\": %s, \"filters_created_by\": \"%s\", \"blocks\": %d, \"warmup_blocks\": %d, \"block_elems\": %zu, "
         "\"taps_min\": %zu, \"taps_max\": %zu, \"seconds\": %.6f, \"input_msps\": %.4f, \"client_msps\": %.2f, "
         "\"outputs\": %llu}\n",
         SIMD_STATUS, g_optimized ? "optimized" : "native", g_n_clients, g_n_threads, n_cpus, pin ? "true" : "false",
         g_local ? "the thread that runs them" : "the main thread", g_blocks, g_warmup, g_block_elems, g_taps_min, g_taps_max, dt, in_samples / dt / 1e6, in_samples / dt / 1e6 * g_n_clients,
         (unsigned long long)outputs);
  for (int i = 0; i < g_n_clients; i++) destroy_xlating(g_clients[i].filter);
  return 0;
}
